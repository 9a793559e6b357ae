#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/t8_all.log 2>&1; echo "rc=$?" >> gpurun_out/t8_all.log
tail -25 gpurun_out/t8_all.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/t8_bench_default.json 2> gpurun_out/t8_bench_default.err; echo "bench rc=$?"
tail -5 gpurun_out/t8_bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/t8_bench_default.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d.get("ragged"), d["chunk_latency_ms"], d.get("realtime_10240_streams",{}).get("chunk_latency_ms"))
PY
