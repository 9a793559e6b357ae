#!/bin/bash
# long-form workload on one GPU: host gather vs device gather (the host also copies 42 MB of arriving audio into the rings every pass)
mkdir -p gpurun_out
for dg in 1 0; do
ASR_B200_DEVICE_GATHER=$dg timeout 900 python bench.py --workload longform --long-chunks 300 > gpurun_out/t44_dg$dg.json 2> gpurun_out/t44.err; echo "longform device_gather=$dg rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/t44_dg$dg.json"))
print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["e2e"]["wall_s"], d["chunk_latency_ms"]["p50"], d["clocks"]["sm_mhz"])
PY
done
