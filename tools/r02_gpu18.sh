#!/bin/bash
# does FFN2 (gemm_ln, K = 2048) speed up when its operands are L2-resident?  per-row time at several M (same launch repeated: small M stays in L2)
python - <<'PY'
import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from asr_streaming_b200.engine import debug_gemm_ln
rng = np.random.default_rng(0)
W = (rng.standard_normal((512, 2048)) / 45).astype(np.float32)
bias = rng.standard_normal(512).astype(np.float32)
g = np.ones(512, np.float32); b = np.zeros(512, np.float32)
for M in (81920, 34816, 17408, 8704):
    A = rng.standard_normal((M, 2048)).astype(np.float32)
    res = rng.standard_normal((M, 512)).astype(np.float32)
    for pair in (2, 0):
        _, _, ms = debug_gemm_ln(A, W, bias, res, g, b, g, b, iters=20, pair=pair)
        print(f"FFN2 shape M={M:6d} pair={pair}: {ms*1e3:8.1f} us  {ms*1e6/M:7.3f} ns/row  {2.0*M*512*2048/ms/1e9:7.0f} TFLOP/s", flush=True)
W = (rng.standard_normal((512, 512)) / 22).astype(np.float32)
for M in (81920, 17408):
    A = rng.standard_normal((M, 512)).astype(np.float32)
    res = rng.standard_normal((M, 512)).astype(np.float32)
    _, _, ms = debug_gemm_ln(A, W, bias, res, g, b, iters=20, pair=0)
    print(f"out_proj shape M={M:6d}: {ms*1e3:8.1f} us  {ms*1e6/M:7.3f} ns/row", flush=True)
PY
