#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload longform --long-chunks 300 > gpurun_out/t11_longform300.json 2> gpurun_out/t11_longform300.err; echo "longform rc=$?"; tail -3 gpurun_out/t11_longform300.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/t11_longform300.json"))
print(d["value"], d["e2e"], d["chunk_latency_ms"], d["clocks"])
PY
for pdl in 1536 100000; do
ASR_B200_PDL_MAX_STREAMS=$pdl timeout 600 python bench.py --workload streams4096 --steps 10 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/t11_pdl$pdl.json 2> gpurun_out/t11_pdl$pdl.err
python - <<PY
import json
d=json.load(open("gpurun_out/t11_pdl$pdl.json"))
print("PDL_MAX=$pdl", d["ms_per_step"], d["value"], d["e2e"]["value"], d["kernel_families_ms_per_step"], d["clocks"])
PY
done
