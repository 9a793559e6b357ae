#!/bin/bash
# decode-stage test (synthetic posteriors), full GPU suite, per-family times after the CTC changes
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t24_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/t24_pytest.log
python bench.py --steps 10 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/t24_bench.json 2> gpurun_out/t24_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/t24_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/t24_bench.json"))
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"])
print(d.get("kernel_families_ms_per_step"))
PY
