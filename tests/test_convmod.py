"""Streaming convolution module (csrc/convmod.cu) against the reference's ConvolutionBlock (lightspeech/layers/block.py:129-171).

CPU: the numpy restatement (oracle/conv_oracle.py) against the fixture produced by the UNMODIFIED reference module
(tests/golden/convblock.npz, oracle/make_conv_goldens.py), and the cached streaming form against the full-sequence form.
GPU: the CUDA module through the C ABI against both, for ragged multi-session batches.
Stated tolerances: EXACT (split-bf16 GEMMs) max-abs <= 2e-4; FAST (bf16 operands) <= 3e-2 on outputs of magnitude ~1."""
import os

import numpy as np
import pytest

from oracle import conv_oracle as CO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "convblock.npz")


@pytest.fixture(scope="module")
def gold():
    g = dict(np.load(GOLD))
    g["W"] = CO.make_conv_weights(int(g["seed"]), int(g["d"]), int(g["k"]))
    return g


def _stream(x, W, T, k):
    """Chunked oracle run incl. the (k-1)/2 flush frames; returns the de-delayed output [len(x), d]."""
    pad = (k - 1) // 2
    n = x.shape[0]
    total = -(-(n + pad) // T) * T
    xp = np.concatenate([x, np.zeros((total - n, x.shape[1]), np.float32)])
    st, out = CO.init_conv_state(x.shape[1], k), []
    for i in range(0, total, T):
        y, st = CO.conv_block_stream(xp[i:i + T], st, W)
        out.append(y)
    return np.concatenate(out)[pad:pad + n]


def test_oracle_matches_reference_module_fixture(gold):
    assert np.abs(CO.conv_block_full(gold["x"], gold["W"]) - gold["y"]).max() < 5e-6


def test_streaming_form_equals_full_sequence_with_delay(gold):
    x, W, k = gold["x"], gold["W"], int(gold["k"])
    for T in (8, 16):
        y = _stream(x, W, T, k)
        # zeros fed after the utterance are not the reference's zero PADDING of the activations (pre_norm bias and SiLU(b1) != 0):
        # frames whose window reaches past the end differ by construction, all others are the reference's numbers
        pad = (k - 1) // 2
        assert np.abs(y[:-pad] - gold["y"][:-pad]).max() < 5e-6, T


def test_pack_order_and_size(gold):
    from asr_streaming_b200.convmod import PARAM_ORDER, pack_conv_weights
    d, k = int(gold["d"]), int(gold["k"])
    blob = pack_conv_weights(gold["W"], d, k)
    assert blob.size == 2 * d + d * d + d + d * k + d + 4 * d + d * d + d
    assert np.array_equal(blob[:d], gold["W"]["pre_norm.scale"]) and np.array_equal(blob[-d:], gold["W"]["pointwise_conv2.bias"])
    assert list(PARAM_ORDER) == [n for n in gold["W"]]                 # state_dict order of the reference module
    bad = dict(gold["W"]); bad["norm.bias"] = bad["norm.bias"][:-1]
    with pytest.raises(ValueError):
        pack_conv_weights(bad, d, k)
    import ctypes as C
    from asr_streaming_b200 import _lib
    n = C.c_uint64()
    assert _lib.load_library().asr_convmod_weights_count(d, k, C.byref(n)) == 0 and n.value == blob.size


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("exact", 2e-4), ("fast", 3e-2)])
def test_cuda_module_matches_reference_fixture(gold, precision, tol):
    from asr_streaming_b200 import PRECISION_EXACT, PRECISION_FAST
    from asr_streaming_b200.convmod import ConvModule, pack_conv_weights
    d, k, T = int(gold["d"]), int(gold["k"]), 16
    x, pad = gold["x"], (int(gold["k"]) - 1) // 2
    blob = pack_conv_weights(gold["W"], d, k)
    with ConvModule(d, k, T, blob, max_sessions=8, max_batch=8, precision=PRECISION_EXACT if precision == "exact" else PRECISION_FAST) as m:
        assert m.delay == pad
        n = x.shape[0]
        total = -(-(n + pad) // T) * T
        xp = np.concatenate([x, np.zeros((total - n, d), np.float32)])
        out = np.concatenate([m.step([3], xp[i:i + T][None])[0] for i in range(0, total, T)])
        y = out[pad:pad + n]
        err = np.abs(y[:-pad] - gold["y"][:-pad]).max()
        assert err < tol, f"{precision}: max-abs {err} vs the reference module"
        assert np.abs(y - _stream(x, gold["W"], T, k)).max() < tol     # incl. the flush frames, vs the streaming oracle


@pytest.mark.gpu
def test_cuda_module_ragged_sessions_and_reset(gold):
    """Sessions at different positions of different utterances in one step; a reset session restarts from a zero cache."""
    from asr_streaming_b200 import PRECISION_EXACT
    from asr_streaming_b200.convmod import ConvModule, pack_conv_weights
    d, k, T = int(gold["d"]), int(gold["k"]), 8
    W = gold["W"]
    rng = np.random.default_rng(12)
    xs = [rng.standard_normal((48, d)).astype(np.float32) for _ in range(3)]
    refs = [_stream(x, W, T, k) for x in xs]
    pad = (k - 1) // 2
    with ConvModule(d, k, T, pack_conv_weights(W, d, k), max_sessions=16, max_batch=4, precision=PRECISION_EXACT) as m:
        got = [[] for _ in xs]
        start = [0, 2, 5]                                    # session i joins at tick start[i]
        slots = [7, 1, 12]
        for tick in range(14):
            idx = [i for i in range(3) if 0 <= tick - start[i] < (48 + T) // T + 1]
            if not idx:
                continue
            batch = []
            for i in idx:
                c = tick - start[i]
                seg = xs[i][c * T:(c + 1) * T]
                batch.append(np.concatenate([seg, np.zeros((T - seg.shape[0], d), np.float32)]))
            y = m.step([slots[i] for i in idx], np.stack(batch))
            for j, i in enumerate(idx):
                got[i].append(y[j])
        for i in range(3):
            out = np.concatenate(got[i])[pad:pad + 48 - pad]
            assert np.abs(out - refs[i][:48 - pad]).max() < 2e-4, i
        # reset: session 7 replays utterance 0 from the start and must reproduce its first chunks exactly
        m.reset([7])
        again = np.concatenate([m.step([7], xs[0][c * T:(c + 1) * T][None])[0] for c in range(3)])
        assert np.array_equal(again, np.concatenate(got[0][:3]))
        with pytest.raises(Exception):
            m.step([1, 1], np.zeros((2, T, d), np.float32))
