"""CPU properties of the prefix-beam oracle (parity-unpinned vs the reference, see oracle/ctc_beam_oracle.py)."""
import itertools
import math

import numpy as np

from oracle import ctc_beam_oracle as B


def _rand_logp(rng, T, V, peak=3.0):
    z = peak * rng.standard_normal((T, V))
    z = z - z.max(1, keepdims=True)
    return (z - np.log(np.exp(z).sum(1, keepdims=True))).astype(np.float64)


def _brute_force(logp):
    T, V = logp.shape
    tot = {}
    for path in itertools.product(range(V), repeat=T):
        p = sum(logp[t, c] for t, c in enumerate(path))
        pre = tuple(c for i, c in enumerate(path) if c != 0 and (i == 0 or c != path[i - 1]))
        tot[pre] = B.logaddexp(tot.get(pre, -math.inf), p)
    return tot


def test_wide_beam_is_exact_on_small_problems():
    rng = np.random.default_rng(0)
    for _ in range(5):
        logp = _rand_logp(rng, 5, 4, peak=1.5)
        tot = _brute_force(logp)
        best_pre = max(tot, key=tot.get)
        st = B.beam_step(B.BeamState(), logp, beam=200, cand_k=3)
        pre, score = B.best(st)
        assert tuple(pre) == best_pre
        assert abs(score - tot[best_pre]) < 1e-9
        for p, pb, pnb in st.entries:                         # every kept prefix carries its exact total probability
            assert abs(B.logaddexp(pb, pnb) - tot[p]) < 1e-9


def test_scores_nonpositive_sorted_and_deterministic():
    rng = np.random.default_rng(1)
    logp = _rand_logp(rng, 40, 50)
    a = B.beam_step(B.BeamState(), logp, beam=10, cand_k=8)
    b = B.beam_step(B.BeamState(), logp, beam=10, cand_k=8)
    assert a.entries == b.entries
    scores = [B.logaddexp(pb, pnb) for _, pb, pnb in a.entries]
    assert all(s <= 1e-12 for s in scores) and scores == sorted(scores, reverse=True)
    assert len({p for p, _, _ in a.entries}) == len(a.entries) <= 10


def test_chunked_equals_one_shot_and_reset():
    rng = np.random.default_rng(2)
    logp = _rand_logp(rng, 48, 30)
    one = B.beam_step(B.BeamState(), logp)
    st = B.BeamState()
    for k in range(0, 48, 16):
        st = B.beam_step(st, logp[k:k + 16])
    assert st.entries == one.entries
    assert B.best(B.BeamState()) == ([], 0.0)


def test_beam_at_least_as_good_as_greedy_on_peaky_input():
    rng = np.random.default_rng(3)
    logp = _rand_logp(rng, 60, 40, peak=6.0)
    idx = logp.argmax(1)
    greedy = [int(c) for i, c in enumerate(idx) if c != 0 and (i == 0 or c != idx[i - 1])]
    pre, score = B.best(B.beam_step(B.BeamState(), logp, beam=10, cand_k=8))
    greedy_path_score = float(logp[np.arange(60), idx].sum())
    assert score >= greedy_path_score - 1e-9                 # the best prefix sums at least the greedy alignment
    assert pre == greedy                                      # peaky posteriors: both decoders agree


def test_max_len_stops_extension():
    rng = np.random.default_rng(4)
    logp = _rand_logp(rng, 30, 10, peak=5.0)
    st = B.beam_step(B.BeamState(), logp, beam=4, cand_k=4, max_len=3)
    assert all(len(p) <= 3 for p, _, _ in st.entries)
