#!/bin/bash
# N ranks on one node with the final round-2 build (default workload): bash tools/r02_gpuN.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_ragged4096_${N}gpu.json 2> gpurun_out/tN_${N}gpu.err
echo "${N}gpu rc=$?"; tail -2 gpurun_out/tN_${N}gpu.err | cut -c1-200
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_ragged4096_${N}gpu.json"))
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"], 3), d["e2e"], d["clocks"], d.get("ragged", {}).get("batch_assembly"), d.get("ragged", {}).get("gpu_busy_ms_per_pass"))
PY
