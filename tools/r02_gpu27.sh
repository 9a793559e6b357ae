#!/bin/bash
# what bounds the GEMM + residual + LayerNorm kernels: knock-out builds (1 = no residual loads, 2 = no stores, 4 = no statistics exchange)
for k in "" 1 2 3 4 7 ""; do
  if [ -z "$k" ]; then unset ASR_B200_LIB; else export ASR_B200_LIB=$PWD/asr_streaming_b200/libasr_b200_knock$k.so; fi
  python tools/gemm_ln_knock.py 2>&1 | tail -1
done
