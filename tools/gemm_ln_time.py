"""Stand-alone timing of the GEMM + residual + LayerNorm kernel at the engine's shapes: python tools/gemm_ln_time.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asr_streaming_b200.engine import debug_gemm_ln  # noqa: E402

rng = np.random.default_rng(0)
sizes = [int(a) for a in sys.argv[1:]] or [3200, 5120, 8000, 20480, 81920]
for M in sizes:
    for K, two in ((512, False), (2048, True)):
        A = rng.standard_normal((M, K)).astype(np.float32)
        W = (rng.standard_normal((512, K)) / np.sqrt(K)).astype(np.float32)
        v = rng.standard_normal(512).astype(np.float32)
        res = rng.standard_normal((M, 512)).astype(np.float32)
        for pair in (0, 1, 2, 3) if two else (0, 1, 3):
            _, _, ms = debug_gemm_ln(A, W, v, res, v, v, v if two else None, v if two else None, iters=20, pair=pair)
            print(f"M={M:6d} K={K:4d} {'two LN' if two else 'one LN'} {('cluster2', 'pair/cluster4', 'pair + 1-pass LN2', 'quad columns')[pair]:18s}: {1e3 * ms:8.1f} us  "
                  f"{2 * M * 512 * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
