#!/bin/bash
mkdir -p gpurun_out
export ASR_B200_LIB=$PWD/asr_streaming_b200/libasr_b200_dbg.so
for shape in "81920 2048 512" "81920 2048 2048"; do
  for d in 0 128 256 143 271; do
    echo -n "== shape $shape dbg $d  "
    ASR_EPI_DBG=$d timeout 120 python - $shape <<'PY'
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
from asr_streaming_b200 import _lib
lib = _lib.load_library()
M, N, K = (int(a) for a in sys.argv[1:4])
ms = C.c_float()
rc = lib.asr_debug_gemm_time(M, N, K, 0, 515, 2, 20, C.byref(ms), 0)
print("  bn 515 rc", rc, "%.1f us  %.0f TFLOP/s" % (ms.value * 1e3, 2.0 * M * N * K / ms.value / 1e9) if not rc else lib.asr_last_error(), flush=True)
PY
  done
done 2>&1 | tee gpurun_out/t7_l2_dbg.log
