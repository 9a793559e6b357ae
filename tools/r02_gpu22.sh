#!/bin/bash
# two-group streaming attention: GPU tests, then A/B against the one-stream-at-a-time form on the same box
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t22_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t22_pytest.log
for v in new old new old; do
  if [ $v = old ]; then export ASR_B200_ATTN_V1=1; else unset ASR_B200_ATTN_V1; fi
  python bench.py --steps 10 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/t22_bench_$v.json 2> gpurun_out/t22_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/t22_bench_$v.json"))
print("$v", round(d["ms_per_step"],3), round(d["value"]), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
print(d.get("kernel_families_ms_per_step"))
PY
done
