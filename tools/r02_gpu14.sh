#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "prestaged or scheduler_ticks" 2>&1 | tail -15
timeout 900 python bench.py --no-cpu-baseline --no-sweep > gpurun_out/t14_bench.json 2> gpurun_out/t14_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/t14_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/t14_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d["ragged"], d["chunk_latency_ms"])
PY
