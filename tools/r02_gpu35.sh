#!/bin/bash
# device gather: how many 128-thread gather CTAs next to the step's persistent kernels (one GPU, pre-staged ticks)
mkdir -p gpurun_out
for c in 148 74 37 296; do
  ASR_B200_DEVICE_GATHER=1 ASR_B200_GATHER_CTAS=$c python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/t35_bench.json 2> gpurun_out/t35_bench.err; echo "gather ctas=$c rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/t35_bench.json"))
print("   value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/pass", round(d["e2e"]["ms_per_step"], 3), "gpu busy", round(d["ragged"]["gpu_busy_ms_per_pass"], 3), d["ragged"]["host_ms_per_tick"])
PY
done
