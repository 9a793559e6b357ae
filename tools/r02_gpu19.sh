#!/bin/bash
export ASR_B200_LIB=$PWD/asr_streaming_b200/libasr_b200_mmat.so
python - <<'PY' 2>&1 | grep -v "^$" | tail -40
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
from asr_streaming_b200 import _lib
lib = _lib.load_library()
for (M,N,K,bn,epi) in ((81920,2048,512,515,2),(81920,2048,512,512,3),(81920,2048,2048,515,2),(81920,1536,512,515,2)):
    ms = C.c_float()
    rc = lib.asr_debug_gemm_time(M, N, K, 0, bn, epi, 3, C.byref(ms), 0)
    print("== bn", bn, "epi", epi, "N", N, "K", K, "rc", rc, "%.1f us" % (ms.value*1e3), flush=True)
import numpy as np
from asr_streaming_b200.engine import debug_gemm_ln
rng = np.random.default_rng(0)
M=81920
for K,pair,two in ((2048,2,True),(512,0,False)):
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((512, K)) / np.sqrt(K)).astype(np.float32)
    v = rng.standard_normal(512).astype(np.float32)
    res = rng.standard_normal((M, 512)).astype(np.float32)
    _, _, ms = debug_gemm_ln(A, W, v, res, v, v, v if two else None, v if two else None, iters=3, pair=pair)
    print("== gemm_ln K", K, "pair", pair, "%.1f us" % (ms*1e3), flush=True)
PY
