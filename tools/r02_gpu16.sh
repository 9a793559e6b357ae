#!/bin/bash
# 8 ranks on one host: host gather vs device gather with pre-staged one-tick passes
mkdir -p gpurun_out
N=8
for dg in 0 1; do
ASR_B200_DEVICE_GATHER=$dg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$dg bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/t16_default_n8_dg$dg.json 2> gpurun_out/t16_default_n8_dg$dg.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/t16_default_n8_dg$dg.json"))
print("dg=$dg", d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["ragged"]["host_ms_per_tick"], d["ragged"]["gpu_busy_ms_per_pass"], d["chunk_latency_ms"])
PY
done
nproc
