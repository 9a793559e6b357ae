"""Oracle vs the live reference import (only where /root/reference exists, i.e. the build container)."""
import numpy as np
import pytest

from oracle import lightspeech_oracle as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")


def test_oracle_tracks_live_reference(oracle_weights):
    import torch
    R, m = ref_import.build_reference_model(oracle_weights)
    rng = np.random.Generator(np.random.PCG64(99))
    pcm = (0.05 * rng.standard_normal(O.CANONICAL.segment_length * 3)).astype(np.float32)
    state, st = m.init_state(), O.init_state()
    for ch in O.chunk_windows(pcm):
        em, _, sts = m.stream([torch.from_numpy(ch)[None]], 16000, [state])
        state = sts[0]
        eo, st = O.stream_chunk(ch, st, oracle_weights)
        assert np.abs(eo - em[0].numpy()).max() < 5e-5
    assert R.greedy_search(em[0])[0] == O.greedy_search(eo, R.vocab)[0]


def test_product_vocab_and_weights_match_oracle(oracle_weights):
    from asr_streaming_b200 import random_weights
    w = random_weights(1234)
    assert set(w) == set(oracle_weights)
    for k in w:
        assert np.array_equal(w[k], oracle_weights[k]), k


def test_torch_port_is_bit_identical_to_reference(oracle_weights):
    """oracle/torch_ref_port.py (the CPU arm bench.py times on the GPU box) == the real reference, bit for bit."""
    import torch
    from oracle.torch_ref_port import TorchRefPort, greedy_search
    R, m = ref_import.build_reference_model(oracle_weights)
    port = TorchRefPort(oracle_weights)
    rng = np.random.Generator(np.random.PCG64(17))
    pcm = (0.05 * rng.standard_normal(O.CANONICAL.segment_length * 3)).astype(np.float32)
    s_ref, s_port = m.init_state(), port.init_state()
    acc = torch.zeros(0, 804)
    for ch in O.chunk_windows(pcm):
        x = torch.from_numpy(ch)[None]
        a, la, sa = m.stream([x], 16000, [s_ref])
        b, lb, sb = port.stream([x], 16000, [s_port])
        s_ref, s_port = sa[0], sb[0]
        assert torch.equal(a, b) and torch.equal(la, lb)
        acc = torch.cat((acc, a[0]))
        assert R.greedy_search(acc) == greedy_search(acc, R.vocab)
