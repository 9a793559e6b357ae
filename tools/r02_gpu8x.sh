#!/bin/bash
# 8-GPU node: default workload (configs[3] per GPU) and the long-form low-latency workload (configs[4]: 32k streams, 10 min each)
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/t13_default_n$N.json 2> gpurun_out/t13_default_n$N.err; echo "default rc=$?"
tail -2 gpurun_out/t13_default_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload longform --long-chunks 1875 > gpurun_out/t13_longform_n$N.json 2> gpurun_out/t13_longform_n$N.err; echo "longform rc=$?"
tail -2 gpurun_out/t13_longform_n$N.err
python - <<PY
import json
for f in ("t13_default_n$N", "t13_longform_n$N"):
    try:
        d=json.load(open("gpurun_out/%s.json" % f))
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"], d.get("chunk_latency_ms"), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
