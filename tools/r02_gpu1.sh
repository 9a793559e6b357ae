#!/bin/bash
# round 2, GPU call 1: TMA-store epilogues — parity, stand-alone timing, in-step A/B
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -k "tma_store or tcgen05_gemm" > gpurun_out/t1_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/t1_gemm.log
tail -5 gpurun_out/t1_gemm.log
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "tma_store" > gpurun_out/t1_parity.log 2>&1; echo "rc=$?" >> gpurun_out/t1_parity.log
tail -5 gpurun_out/t1_parity.log
timeout 300 python tools/ares_time.py 81920 2048 512 > gpurun_out/t1_ares_ffn1.log 2>&1
cat gpurun_out/t1_ares_ffn1.log
for v in 0 1; do
  ASR_B200_NO_TMA_STORE=$v timeout 600 python bench.py --workload streams4096 --steps 10 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/t1_bench4096_notma$v.json 2> gpurun_out/t1_bench4096_notma$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/t1_bench4096_notma$v.json"))
print("NO_TMA_STORE=$v", d["ms_per_step"], d["value"], d["kernel_families_ms_per_step"], d["clocks"])
PY
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t1_all.log 2>&1; echo "rc=$?" >> gpurun_out/t1_all.log
tail -15 gpurun_out/t1_all.log
