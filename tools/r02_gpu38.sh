#!/bin/bash
# EXACT attention with ldmatrix K fragments, CTC kernel unrolled over 26 columns per lane: tests + default bench (exact leg, decode families)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t38_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t38_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/t38_bench.json 2> gpurun_out/t38_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/t38_bench.json"))
print(round(d["ms_per_step"],3), round(d["value"]), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
print(d["kernel_families_ms_per_step"])
print("exact", d["exact"]["ms_per_step"], d["exact"]["value"], d["exact"]["kernel_families_ms_per_step"])
PY
