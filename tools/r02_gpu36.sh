#!/bin/bash
# BASELINE configs[4] long form on 8 GPUs with the final round-2 build: 32,768 low-latency streams x 1875 chunks (10 minutes each)
mkdir -p gpurun_out
export ASR_B200_DEVICE_GATHER=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --workload longform --long-chunks 1875 > gpurun_out/r02_bench_longform_8gpu.json 2> gpurun_out/t36_longform.err
echo "longform rc=$?"; tail -3 gpurun_out/t36_longform.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_longform_8gpu.json"))
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"], 3), d["e2e"], d.get("chunk_latency_ms"), d["clocks"])
print({k: v for k, v in d.items() if k in ("longform", "ragged", "config")})
PY
