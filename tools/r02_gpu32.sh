#!/bin/bash
# 8 ranks on one node with the final round-2 build: default workload (pre-staged device gather, speculative fbank)
mkdir -p gpurun_out
export ASR_B200_DEVICE_GATHER=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_ragged4096_8gpu.json 2> gpurun_out/t32_8gpu.err
echo "8gpu rc=$?"; tail -3 gpurun_out/t32_8gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_ragged4096_8gpu.json"))
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"], 3), d["e2e"], d["clocks"], d.get("ragged", {}).get("gpu_busy_ms_per_pass"))
PY
