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


def _gelu(x):
    from math import erf
    return 0.5 * x * (1.0 + np.vectorize(erf)(x / np.sqrt(2.0)))


OPERAND_CASES = [
    # M, N, K, act (0 none, 1 GELU, 2 SiLU)
    (256, 256, 64, 0),               # one pair tile, one k-block
    (5120, 2048, 512, 1),            # FFN1 at 256 streams: 160 pair tiles on 74 pairs (staging box reused, accumulator ring wraps)
    (5000, 2048, 512, 1),            # ragged M: the last 32-row boxes hang over M (padded operand rows)
    (2432, 512, 512, 2),             # CTC1 shape, odd number of 128-row tiles: the peer CTA of the last pair is all padding
    (81920, 2048, 512, 1),           # FFN1 at 4096 streams
]


@pytest.mark.parametrize("M,N,K,act", OPERAND_CASES)
def test_tma_store_epilogue_equals_lsu_epilogue(M, N, K, act):
    """The bf16-operand epilogue of the cta_group::2 GEMM with TMA stores (bn 515: rows staged once in the 128B-swizzled box layout,
    one cp.async.bulk.tensor store per warp) against the LSU epilogue (bn 512) — same arithmetic, bit-identical — and against numpy."""
    from asr_streaming_b200.engine import debug_gemm_operand
    rng = np.random.default_rng(M + N + K + act)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    lsu = debug_gemm_operand(A, B, bias, act=act, bn=512)
    tma = debug_gemm_operand(A, B, bias, act=act, bn=515)
    assert np.isfinite(tma).all()
    assert np.array_equal(lsu, tma)
    if M <= 5120:
        v = _ref(A, B, bias, 0).astype(np.float64)
        ref = v if act == 0 else (_gelu(v) if act == 1 else v / (1.0 + np.exp(-v)))
        assert np.abs(tma - ref).max() < 0.03                         # bf16 rounding of values up to ~5


# ------------------------------------------------------------------------------------------------ GEMM + residual + LayerNorm epilogue
def _ln(x, g, b):
    x = x.astype(np.float64)
    mu = x.mean(1, keepdims=True)
    var = ((x - mu) ** 2).mean(1, keepdims=True)
    return (x - mu) / np.sqrt(var + 1e-5) * g + b


LN_CASES = [
    # M, K, split, mode ("a": out_proj form, "b": two LayerNorms, "c": last layer, compact segment rows), row mean offset
    (128, 64, 0, "a", 0.0),
    (20, 512, 0, "a", 0.0),            # one stream: a single, mostly empty tile
    (5120, 512, 0, "a", 0.0),          # out_proj at 256 streams: 40 row tiles
    (5120, 2048, 0, "b", 0.0),         # FFN2 at 256 streams
    (20000, 512, 0, "b", 3.0),         # more tiles than clusters (157 > 74): persistent loop, TMEM double buffering, stats ping-pong
    (1333, 2048, 0, "b", -7.5),        # ragged M, rows with a large common offset (variance must not cancel)
    (5120, 2048, 0, "c", 0.0),
    (1000, 512, 1, "a", 1.0),          # EXACT: three passes, hi|lo operand output
    (1000, 2048, 1, "b", 0.0),
    (1000, 2048, 1, "c", 0.0),
]


@pytest.mark.parametrize("pair", [0, 1, 2, 3], ids=["cluster2", "pair_cluster4", "pair_onepass_ln2", "quad_columns"])
@pytest.mark.parametrize("M,K,split,mode,offset", LN_CASES)
def test_gemm_residual_layernorm_epilogue(M, K, split, mode, offset, pair):
    import functools
    from asr_streaming_b200 import engine as E
    debug_gemm_ln = functools.partial(E.debug_gemm_ln, pair=pair)
    rng = np.random.default_rng(M + K + split + ord(mode))
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((512, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(512).astype(np.float32)
    res = (rng.standard_normal((M, 512)) + offset).astype(np.float32)
    g1, b1, g2, b2 = [(s * rng.standard_normal(512) + o).astype(np.float32) for s, o in ((0.2, 1.0), (0.3, 0.0), (0.2, 1.0), (0.3, 0.0))]
    v = _ref(A, W, bias, split).astype(np.float64) + res
    rows, seg = (20, 16) if mode == "c" else (0, 0)
    if mode == "a":
        out, op, _ = debug_gemm_ln(A, W, bias, res, g1, b1, split=split)
        ref_out, ref_op = v, _ln(v, g1, b1)
    elif mode == "b":
        out, op, _ = debug_gemm_ln(A, W, bias, res, g1, b1, g2, b2, split=split)
        ref_out = _ln(v, g1, b1)
        ref_op = _ln(ref_out, g2, b2)
    else:
        out, op, _ = debug_gemm_ln(A, W, bias, res, g1, b1, f32_normed=True, compact_rows=rows, compact_seg=seg, split=split)
        ref_out = _ln(v, g1, b1)
        keep = (np.arange(M) % rows) < seg
        ref_op = ref_out[keep][:op.shape[0]]
    tol = 3e-4 * max(1.0, abs(offset))
    assert np.abs(out - ref_out).max() < tol, f"fp32 output max-abs {np.abs(out - ref_out).max()}"
    op_tol = 2e-4 if split else 0.02                      # bf16 operand rounding (values up to ~5)
    assert np.abs(op - ref_op).max() < op_tol, f"operand output max-abs {np.abs(op - ref_op).max()}"


def test_gemm_layernorm_rows_do_not_depend_on_their_position():
    """The same input row must give bit-identical outputs wherever it sits in the batch (statistics are combined in a fixed order)."""
    from asr_streaming_b200.engine import debug_gemm_ln
    rng = np.random.default_rng(3)
    M, K = 1000, 512
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((512, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(512).astype(np.float32)
    res = rng.standard_normal((M, 512)).astype(np.float32)
    g = (1.0 + 0.2 * rng.standard_normal(512)).astype(np.float32); b = (0.3 * rng.standard_normal(512)).astype(np.float32)
    perm = rng.permutation(M)
    for pair in (0, 1, 2, 3):
        o1, p1, _ = debug_gemm_ln(A, W, bias, res, g, b, g, b, pair=pair)
        o2, p2, _ = debug_gemm_ln(A[perm], W, bias, res[perm], g, b, g, b, pair=pair)
        assert np.array_equal(o1[perm], o2) and np.array_equal(p1[perm], p2)
