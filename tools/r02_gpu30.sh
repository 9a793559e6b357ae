#!/bin/bash
# speculative fbank of the pre-staged chunks (runs in the collect -> submit gap): GPU tests, then A/B of the default bench's e2e leg
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t30_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t30_pytest.log
for v in spec nospec spec nospec; do
  if [ $v = nospec ]; then export ASR_B200_NO_SPEC_FBANK=1; else unset ASR_B200_NO_SPEC_FBANK; fi
  python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/t30_bench_$v.json 2> gpurun_out/t30_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/t30_bench_$v.json"))
print("$v value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e ms/pass", round(d["e2e"]["ms_per_step"], 3), "gpu busy", round(d["ragged"]["gpu_busy_ms_per_pass"], 3), d["clocks"]["sm_mhz"], d["ragged"]["host_ms_per_tick"])
PY
done
