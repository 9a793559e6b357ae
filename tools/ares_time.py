"""Times the cta_group::2 GEMM epilogue variants on one shape: python tools/ares_time.py M N K
   (bn 512 = LSU epilogue, 515 = TMA-store epilogue; epi 2 = bias + GELU -> bf16, 3 = accumulator dropped: the mainloop alone)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asr_streaming_b200 import _lib  # noqa: E402

lib = _lib.load_library()
M, N, K = (int(a) for a in sys.argv[1:4])
for bn, epi in ((512, 2), (515, 2), (512, 3), (515, 2), (512, 2)):
    ms = C.c_float()
    rc = lib.asr_debug_gemm_time(M, N, K, 0, bn, epi, 20, C.byref(ms), 0)
    print("bn", bn, "epi", epi, "rc", rc, (lib.asr_last_error() or b"").decode()[:300] if rc else "%.1f us  %.0f TFLOP/s" % (ms.value * 1e3, 2.0 * M * N * K / ms.value / 1e9), flush=True)
