#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_guards_gpu.py -x -q > gpurun_out/t9_guards.log 2>&1; echo "rc=$?" >> gpurun_out/t9_guards.log
tail -25 gpurun_out/t9_guards.log
for rows in 4096 2048; do
ASR_BENCH_E2E_ROWS=$rows timeout 900 python bench.py --no-cpu-baseline --no-sweep > gpurun_out/t9_bench_rows$rows.json 2> gpurun_out/t9_bench_rows$rows.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/t9_bench_rows$rows.json"))
print($rows, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["ragged"]["host_ms_per_tick"], d["ragged"]["gpu_busy_ms_per_pass"], d["chunk_latency_ms"])
PY
done
