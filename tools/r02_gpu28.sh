#!/bin/bash
# knock-outs 8 / 16 / 24: the same store / load instructions, but to / from an L2-resident 4096-row window
for k in "" 8 16 24 ""; do
  if [ -z "$k" ]; then unset ASR_B200_LIB; else export ASR_B200_LIB=$PWD/asr_streaming_b200/libasr_b200_knock$k.so; fi
  python tools/gemm_ln_knock.py 2>&1 | tail -1
done
