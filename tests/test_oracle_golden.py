"""The numpy oracle (oracle/lightspeech_oracle.py) against the fixtures produced by the UNMODIFIED reference
(oracle/make_goldens.py).  CPU only.  Tolerance: the oracle and the reference are both fp32 and differ only in
summation order; observed max-abs 3e-6 on log-probs, bound 5e-5."""
import numpy as np
import pytest

from oracle import lightspeech_oracle as O

TOL = 5e-5


def _run(case, W, geo, meta_case, max_chunks=None):
    pcm = case["pcm"].astype(np.float32) / np.float32(32768.0)
    chunks = O.chunk_windows(pcm, geo)
    st = O.init_state(geo)
    ems = []
    for k, ch in enumerate(chunks):
        if k in meta_case["reset_before"]:
            st = O.init_state(geo)
        if k in meta_case["skip"]:
            continue
        em, st = O.stream_chunk(ch, st, W, geo)
        ems.append(em)
        if max_chunks and len(ems) >= max_chunks:
            break
    return np.stack(ems), st


@pytest.mark.parametrize("name,max_chunks", [("synth_noise", 3), ("testwav", 2), ("edge_silence", 2), ("edge_fullscale", 2),
                                             ("edge_dc", 2), ("seq_reset_skip", None), ("synth_tone", 2)])
def test_emission_matches_reference(name, max_chunks, golden, meta, oracle_weights):
    case, mc = golden(name), meta["cases"][name]
    em, st = _run(case, oracle_weights, O.CANONICAL, mc, max_chunks)
    ref = case["emission"][:em.shape[0]]
    assert np.abs(em - ref).max() < TOL
    assert (em.argmax(2) == case["argmax"][:em.shape[0]]).all()
    if max_chunks is None:
        assert st[0].past_length == mc["past_length"]
        assert np.abs(st[0].k - case["state_k_l0"]).max() < TOL
        assert np.abs(st[-1].v - case["state_v_l19"]).max() < TOL


def test_low_latency_geometry(golden, meta, oracle_weights):
    case, mc = golden("lowlat_noise"), meta["cases"]["lowlat_noise"]
    em, _ = _run(case, oracle_weights, O.LOW_LATENCY, mc, 3)
    assert em.shape[1] == 8
    assert np.abs(em - case["emission"][:3]).max() < TOL


def test_melspec128_matches_extract_filterbank(golden):
    fb = golden("fbank")
    pcm = fb["melspec_pcm"].astype(np.float32) / np.float32(32768.0)
    out = O.melspec128(pcm, dtype=np.float64)
    assert out.shape == (80, 128)
    assert np.abs(out - fb["melspec128"]).max() < 2e-4          # reference FFT is fp32


def test_greedy_matches_reference_texts(golden, meta):
    for name in ("synth_noise", "testwav", "seq_reset_skip"):
        case, mc = golden(name), meta["cases"][name]
        em = case["emission"]
        acc = np.zeros((0, em.shape[2]), np.float32)
        j = 0
        # replay the caller: emission accumulates, is cleared at a reset, skipped chunks are not in the fixture
        kept = [k for k in range(mc["n_chunks"] + len(mc["skip"])) if k not in mc["skip"]]
        for j, k in enumerate(kept):
            if k in mc["reset_before"]:
                acc = np.zeros((0, em.shape[2]), np.float32)
            acc = np.concatenate([acc, em[j]])
            text, last_blank = O.greedy_search(acc, meta["vocab"])
            assert text == mc["texts"][j]
            ids, lb, _ = O.greedy_ids(acc)
            tok = np.nonzero(acc.argmax(1) > 1)[0]
            exp = float(np.float32(len(acc) - 1 - tok[-1]) * np.float32(0.04)) if tok.size else 0.04 * len(acc)
            assert abs(case["last_blank"][j] - exp) < 1e-6


def test_incremental_greedy_equals_rescan(golden):
    """Carrying (prev_id, n_frames, last_tok_frame) across chunks == greedy over the concatenated emission."""
    em = golden("testwav")["emission"]
    prev, nf, lt, toks = -1, 0, -1, []
    for ch in em:
        for idx in ch.argmax(1):
            idx = int(idx)
            if idx != prev and idx != 0:
                toks.append(idx)
            if idx > 1:
                lt = nf
            prev = idx
            nf += 1
        ids, lb, _ = O.greedy_ids(em[: nf // em.shape[1]].reshape(-1, em.shape[2]))
        assert ids == toks
        assert abs(lb - ((nf - 1 - lt) * 0.04 if lt >= 0 else 0.04 * nf)) < 1e-9


def test_precision_models_meet_stated_tolerances(golden, meta, oracle_weights):
    """Derives the tolerances the GPU tests use: bf16 operands (FAST) <= 1e-2, split-bf16 (EXACT) <= 1e-4 on log-probs."""
    case = golden("synth_noise")
    pcm = case["pcm"].astype(np.float32) / np.float32(32768.0)
    ch = O.chunk_windows(pcm)[0]
    ref = case["emission"][0]
    fast, _ = O.stream_chunk(ch, O.init_state(), oracle_weights, mm=O.mm_bf16)
    exact, _ = O.stream_chunk(ch, O.init_state(), oracle_weights, mm=O.mm_bf16x3)
    assert np.abs(fast - ref).max() < 1e-2
    assert np.abs(exact - ref).max() < 1e-4
