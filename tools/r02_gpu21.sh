#!/bin/bash
# beam / CTC restructure: GPU tests, then the default bench's per-family times
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t21_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t21_pytest.log
python bench.py --steps 10 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/t21_bench.json 2> gpurun_out/t21_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/t21_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/t21_bench.json"))
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"])
print(d.get("kernel_families_ms_per_step"))
PY
