#!/bin/bash
# FFN2's ragged last wave on a concurrent side launch (spare SMs): tests, then A/B on one box
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t40_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t40_pytest.log
for v in split nosplit split nosplit; do
  if [ $v = nosplit ]; then export ASR_B200_NO_LN_TAIL=1; else unset ASR_B200_NO_LN_TAIL; fi
  python bench.py --steps 10 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/t40_bench_$v.json 2> gpurun_out/t40_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/t40_bench_$v.json"))
f = d["kernel_families_ms_per_step"]
print("$v", round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["value"]), "ffn2", f["gemm_ffn2"], "ffn1", f["gemm_ffn1"], "out", f["gemm_out_proj"], "gpu busy", round(d["ragged"]["gpu_busy_ms_per_pass"],3), d["clocks"]["sm_mhz"], d["gpu_launches"])
PY
done
