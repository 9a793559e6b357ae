"""tcgen05/TMEM/TMA GEMM kernel against numpy on the same bf16-rounded operands (and the CUDA-core cross-check).
Runs in its own process on the GPU box: a descriptor bug traps the context."""
import numpy as np
import pytest

from oracle.lightspeech_oracle import to_bf16

pytestmark = pytest.mark.gpu


def _ref(A, B, bias, split):
    A = A.astype(np.float32); B = B.astype(np.float32)
    Ah, Bh = to_bf16(A).astype(np.float64), to_bf16(B).astype(np.float64)
    if not split:
        C = Ah @ Bh.T
    else:
        Al, Bl = to_bf16(A - Ah.astype(np.float32)).astype(np.float64), to_bf16(B - Bh.astype(np.float32)).astype(np.float64)
        C = Ah @ Bh.T + Al @ Bh.T + Ah @ Bl.T
    return (C + (bias if bias is not None else 0.0)).astype(np.float32)


CASES = [
    # M, N, K, bn, split
    (128, 128, 64, 128, 0),
    (128, 256, 128, 256, 0),
    (20, 128, 128, 128, 0),          # one stream: M tail inside a tile
    (333, 512, 512, 128, 0),         # ragged M
    (1000, 1536, 512, 256, 0),       # QKV shape, multiple tiles per CTA? (8 x 6 = 48 tiles)
    (5120, 2048, 512, 256, 0),       # FFN1 at 256 streams: 320 tiles > 148 CTAs -> persistent loop + TMEM double buffer
    (2560, 512, 2048, 256, 0),       # FFN2: long K, smem ring wraps many times
    (640, 804, 512, 256, 0),         # CTC vocab: ragged N, TMA zero fill
    (640, 804, 512, 64, 0),
    (777, 512, 512, 64, 0),
    (1000, 1536, 512, 256, 1),       # EXACT: three passes
    (512, 256, 64, 512, 0),          # bn = 512 selects the cta_group::2 CTA-pair kernel: one pair tile, one k-block
    (256, 512, 512, 512, 0),
    (5120, 1536, 512, 512, 0),       # QKV at 256 streams: 120 pair tiles on 74 pairs
    (5000, 2048, 512, 512, 0),       # ragged M inside a pair (last pair: second CTA partly / fully out of range)
    (2432, 512, 2048, 512, 0),       # odd number of 128-row tiles (19): the peer CTA of the last pair is all padding
    (1000, 1536, 512, 512, 1),       # pair kernel, EXACT passes
    (300, 804, 512, 128, 1),
]


@pytest.mark.parametrize("M,N,K,bn,split", CASES)
def test_tcgen05_gemm(M, N, K, bn, split):
    from asr_streaming_b200.engine import debug_gemm
    rng = np.random.default_rng(M * 7 + N + K + bn + split)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ref = _ref(A, B, bias, split)
    out = debug_gemm(A, B, bias, impl=0, split=split, bn=bn)
    err = np.abs(out - ref).max()
    assert err < 2e-4, f"tcgen05 GEMM max-abs {err}"


def test_simt_crosscheck_gemm():
    from asr_streaming_b200.engine import debug_gemm
    rng = np.random.default_rng(5)
    A = rng.standard_normal((300, 512)).astype(np.float32)
    B = (rng.standard_normal((804, 512)) / 22.0).astype(np.float32)
    for split in (0, 1):
        out = debug_gemm(A, B, None, impl=1, split=split)
        assert np.abs(out - _ref(A, B, None, split)).max() < 2e-4
