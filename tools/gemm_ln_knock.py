"""Times the out_proj / FFN2 shapes of gemm_ln.cu for the library named by ASR_B200_LIB (diagnostic knock-out builds, -DASR_LN_KNOCK=n)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asr_streaming_b200.engine import debug_gemm_ln  # noqa: E402

rng = np.random.default_rng(0)
M = 81920
out = []
for K, pair, two in ((512, 0, False), (2048, 2, True)):
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((512, K)) / np.sqrt(K)).astype(np.float32)
    v = rng.standard_normal(512).astype(np.float32)
    res = rng.standard_normal((M, 512)).astype(np.float32)
    _, _, ms = debug_gemm_ln(A, W, v, res, v, v, v if two else None, v if two else None, iters=10, pair=pair)
    out.append(f"K={K} pair={pair}: {1e3 * ms:7.1f} us")
print(os.path.basename(os.environ.get("ASR_B200_LIB", "libasr_b200.so")), " | ".join(out), flush=True)
