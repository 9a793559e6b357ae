#!/bin/bash
# GEMM + LayerNorm epilogue through TMA stores: unit tests, stand-alone timing, then the whole suite and an A/B of the default bench
mkdir -p gpurun_out
python -m pytest tests/test_gemm_gpu.py -m gpu -x -q > gpurun_out/t29_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -6 gpurun_out/t29_gemm.log
python tools/gemm_ln_knock.py 2>&1 | tail -1
ASR_B200_NO_TMA_STORE=1 python tools/gemm_ln_knock.py 2>&1 | tail -1
python -m pytest tests -m gpu -x -q > gpurun_out/t29_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t29_pytest.log
for v in tma lsu tma lsu; do
  if [ $v = lsu ]; then export ASR_B200_NO_TMA_STORE=1; else unset ASR_B200_NO_TMA_STORE; fi
  python bench.py --steps 10 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/t29_bench_$v.json 2> gpurun_out/t29_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/t29_bench_$v.json"))
print("$v", round(d["ms_per_step"],3), round(d["value"]), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
print({k: v for k, v in d.get("kernel_families_ms_per_step").items() if k.startswith("gemm_")})
PY
done
