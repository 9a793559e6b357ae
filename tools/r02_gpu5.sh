#!/bin/bash
# round 2, GPU call 5: warp-role placement (producer / MMA warps at the highest vs lowest warp ids)
mkdir -p gpurun_out
for lib in libasr_b200.so libasr_b200_rf.so; do
  export ASR_B200_LIB=$PWD/asr_streaming_b200/$lib
  echo "==== $lib"
  for shape in "81920 2048 512" "81920 1536 512" "81920 2048 2048"; do
    echo "== $shape"; timeout 300 python tools/ares_time.py $shape
  done
done 2>&1 | tee gpurun_out/t5_roles.log
unset ASR_B200_LIB
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_parity_gpu.py -x -q -k "tma_store" 2>&1 | tail -3
