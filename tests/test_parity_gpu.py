"""Parity of the CUDA path (through the C ABI) with the reference fixtures and the CPU oracle.

Stated tolerances (derived in tests/test_oracle_golden.py::test_precision_models_meet_stated_tolerances):
  fbank melspec128 : max-abs <= 2e-3 in the log domain on audio with signal, bins > clamp (fp32 FFT both sides)
  EXACT precision  : log-probs max-abs <= 2e-4, greedy token ids bit-exact on every frame
  FAST  precision  : log-probs max-abs <= 1e-2 (north-star), greedy ids exact on frames whose fp32 top-2 margin > 2e-2
"""
import numpy as np
import pytest

from oracle import lightspeech_oracle as O
from helpers import chunks_i16, margins, model_cfg, report, to_float

pytestmark = pytest.mark.gpu

EXACT_TOL, FAST_TOL = 2e-4, 1e-2


@pytest.fixture(scope="module")
def engines(packed_weights):
    from asr_streaming_b200 import Engine, PRECISION_EXACT, PRECISION_FAST
    made = {}

    def get(precision, low_latency=False):
        key = (precision, low_latency)
        if key not in made:
            made[key] = Engine(model_cfg(precision, low_latency), packed_weights)
        return made[key]
    get.EXACT, get.FAST = PRECISION_EXACT, PRECISION_FAST
    yield get
    for e in made.values():
        e.close()


# ------------------------------------------------------------------------------------------------ fbank
def test_melspec128_vs_reference(engines, golden):
    fb = golden("fbank")
    e = engines(engines.FAST)
    out_i16 = e.fbank(fb["melspec_pcm"][None, :])[0]
    out_f32 = e.fbank(to_float(fb["melspec_pcm"])[None, :])[0]
    assert out_i16.shape == (80, 128)
    report(f"melspec128 vs reference extract_filterbank: max-abs {np.abs(out_i16 - fb['melspec128']).max():.3e}; vs float64 oracle {np.abs(out_i16 - O.melspec128(to_float(fb['melspec_pcm']))).max():.3e}")
    assert np.array_equal(out_i16, out_f32)                    # int16 and float inputs are the same samples
    assert np.abs(out_i16 - fb["melspec128"]).max() < 2e-3
    assert np.abs(out_i16 - O.melspec128(to_float(fb["melspec_pcm"]))).max() < 2e-3


def test_melspec128_edge_inputs(engines):
    e = engines(engines.FAST)
    n = O.CANONICAL.chunk_length
    sil = e.fbank(np.zeros((1, n), np.int16))[0]
    assert np.all(sil == np.float32(np.log(np.float32(1e-5))))  # clamp(1e-5).log() exactly
    fs = np.where((np.arange(n) // 37) % 2 == 0, 32767, -32768).astype(np.int16)
    dc = np.full(n, 12000, np.int16)
    for pcm in (fs, dc):
        got = e.fbank(pcm[None])[0]
        ref = O.melspec128(to_float(pcm))
        big = ref > np.log(1e-3)                                # leakage bins far below the signal are fp32 noise on both sides
        assert np.abs(got - ref)[big].max() < 5e-3
        assert np.isfinite(got).all()


def test_kaldi80_vs_torchaudio(engines, golden):
    fb = golden("fbank")
    e = engines(engines.FAST)
    out = e.fbank(fb["kaldi_pcm"][None, :], kind=1)[0]
    assert out.shape == fb["kaldi80"].shape == (64, 80)
    report(f"kaldi80 vs torchaudio.compliance.kaldi.fbank: max-abs {np.abs(out - fb['kaldi80']).max():.3e}")
    assert np.abs(out - fb["kaldi80"]).max() < 2e-3
    out_cmvn = e.fbank(fb["kaldi_pcm"][None, :], kind=1, subtract_mean=True)[0]
    assert np.abs(out_cmvn - fb["kaldi80_cmvn"]).max() < 2e-3


def test_fbank_batch_is_per_stream(engines):
    e = engines(engines.FAST)
    rng = np.random.default_rng(3)
    pcm = rng.integers(-3000, 3000, size=(37, O.CANONICAL.chunk_length)).astype(np.int16)
    out = e.fbank(pcm)
    for i in (0, 17, 36):
        assert np.array_equal(out[i], e.fbank(pcm[i:i + 1])[0])


# ------------------------------------------------------------------------------------------------ full path vs reference fixtures
def _run_case(e, case, mc, geo=O.CANONICAL):
    slot = e.open_session()
    ems, toks, blanks, texts_ids = [], [], [], []
    acc = []
    for k, ch in enumerate(chunks_i16(case["pcm"], geo)):
        if k in mc["reset_before"]:
            e.reset_session(slot)
            acc = []
        if k in mc["skip"]:
            continue
        r = e.step([slot], ch[None, :], want_logprobs=True)
        ems.append(r.logprobs[0])
        acc.extend(int(t) for t in r.new_tokens[0])
        texts_ids.append(list(acc))
        blanks.append(r.last_blank(0))
    e.close_session(slot)
    return np.stack(ems), texts_ids, blanks


@pytest.mark.parametrize("name", ["synth_noise", "synth_tone", "testwav", "edge_silence", "edge_fullscale", "edge_dc", "seq_reset_skip"])
def test_exact_mode_matches_reference(name, engines, golden, meta):
    from asr_streaming_b200 import ids_to_text
    case, mc = golden(name), meta["cases"][name]
    em, ids, blanks = _run_case(engines(engines.EXACT), case, mc)
    ref = case["emission"]
    assert em.shape == ref.shape
    report(f"EXACT {name}: logprob max-abs {np.abs(em - ref).max():.3e}, argmax equal {int((em.argmax(2) == case['argmax']).sum())}/{case['argmax'].size}")
    assert np.abs(em - ref).max() < EXACT_TOL
    assert np.array_equal(em.argmax(2), case["argmax"])                       # bit-exact greedy ids, every frame
    for j in range(len(ids)):
        assert ids_to_text(ids[j], meta["vocab"]) == mc["texts"][j]            # incremental greedy == reference rescans
        assert abs(blanks[j] - case["last_blank"][j]) < 1e-7


@pytest.mark.parametrize("name", ["synth_noise", "testwav", "seq_reset_skip"])
def test_fast_mode_within_tolerance(name, engines, golden, meta):
    case, mc = golden(name), meta["cases"][name]
    em, _, _ = _run_case(engines(engines.FAST), case, mc)
    ref = case["emission"]
    err = np.abs(em - ref).max()
    report(f"FAST {name}: logprob max-abs {err:.3e}, argmax equal {int((em.argmax(2) == case['argmax']).sum())}/{case['argmax'].size}")
    assert err < FAST_TOL, f"bf16 path max-abs {err}"
    safe = margins(ref) > 2 * FAST_TOL
    assert safe.mean() > 0.3
    assert np.array_equal(em.argmax(2)[safe], case["argmax"][safe])


def test_low_latency_geometry(engines, golden, meta):
    case, mc = golden("lowlat_noise"), meta["cases"]["lowlat_noise"]
    em, _, _ = _run_case(engines(engines.EXACT, True), case, mc, O.LOW_LATENCY)
    assert em.shape == case["emission"].shape
    assert np.abs(em - case["emission"]).max() < EXACT_TOL
    assert np.array_equal(em.argmax(2), case["argmax"])


def test_kv_state_matches_reference(engines, golden, meta):
    case, mc = golden("synth_noise"), meta["cases"]["synth_noise"]
    e = engines(engines.EXACT)
    slot = e.open_session()
    for ch in chunks_i16(case["pcm"]):
        e.step([slot], ch[None, :])
    k0, pl = e.debug_read_state(slot, 0, 0)
    v19, _ = e.debug_read_state(slot, 19, 1)
    e.close_session(slot)
    assert pl == mc["past_length"]
    assert np.abs(k0 - case["state_k_l0"]).max() < EXACT_TOL       # ring overwrite == cat + slice (TA:emformer.py:400-414)
    assert np.abs(v19 - case["state_v_l19"]).max() < EXACT_TOL


# ------------------------------------------------------------------------------------------------ ragged batching
def test_ragged_batch_equals_batch1(engines, golden, meta):
    """Streams at different progress (fresh / 1 chunk / steady state / just reset) in ONE step must each equal their
    own batch-1 run — the property torchaudio's batched infer violates (TA:emformer.py:392)."""
    e = engines(engines.EXACT)
    names = ["synth_noise", "testwav", "synth_tone", "edge_fullscale"]
    cases = [golden(n) for n in names]
    chunks = [chunks_i16(c["pcm"]) for c in cases]
    starts = [0, 2, 1, 3]                        # stream i joins at global tick starts[i]
    slots = [e.open_session() for _ in names]
    got = [[] for _ in names]
    for tick in range(6):
        idx = [i for i in range(len(names)) if 0 <= tick - starts[i] < len(chunks[i])]
        if not idx:
            continue
        pcm = np.stack([chunks[i][tick - starts[i]] for i in idx])
        r = e.step([slots[i] for i in idx], pcm, want_logprobs=True)
        for j, i in enumerate(idx):
            got[i].append(r.logprobs[j])
    for i, c in enumerate(cases):
        g = np.stack(got[i])
        ref = c["emission"][:g.shape[0]]
        assert np.abs(g - ref).max() < EXACT_TOL, names[i]
        assert np.array_equal(g.argmax(2), c["argmax"][:g.shape[0]])
    for s in slots:
        e.close_session(s)


def test_batch_order_and_duplicates_of_audio(engines):
    """Permuting the batch permutes the outputs bit-exactly; identical audio in different slots gives identical rows."""
    e = engines(engines.FAST)
    rng = np.random.default_rng(11)
    n = 24
    pcm = rng.integers(-4000, 4000, size=(n, O.CANONICAL.chunk_length)).astype(np.int16)
    pcm[5] = pcm[3]
    a = [e.open_session() for _ in range(n)]
    b = [e.open_session() for _ in range(n)]
    perm = rng.permutation(n)
    for _ in range(3):                           # through L_valid = 0, 16, 32
        ra = e.step(a, pcm, want_logprobs=True)
        rb = e.step([b[i] for i in perm], pcm[perm], want_logprobs=True)
        assert np.array_equal(ra.logprobs[perm], rb.logprobs)
        assert np.array_equal(ra.logprobs[3], ra.logprobs[5])
    for s in a + b:
        e.close_session(s)


def test_full_batch_steady_state_property(engines):
    """max_batch streams at once (size-independent property): every stream fed the same audio must produce the same
    tokens as stream 0, in both precisions."""
    for prec in (engines.FAST, engines.EXACT):
        e = engines(prec)
        n = e.cfg.max_batch
        rng = np.random.default_rng(2)
        one = rng.integers(-3000, 3000, size=(4, O.CANONICAL.chunk_length)).astype(np.int16)
        slots = [e.open_session() for _ in range(n)]
        for t in range(4):
            r = e.step(slots, np.repeat(one[t][None], n, 0))
            assert (r.argmax_ids == r.argmax_ids[0]).all()
        for s in slots:
            e.close_session(s)


# ------------------------------------------------------------------------------------------------ error behaviour
def test_error_paths(engines):
    from asr_streaming_b200 import AsrLibraryError
    e = engines(engines.FAST)
    n = O.CANONICAL.chunk_length
    with pytest.raises(AsrLibraryError):
        e.step([12345], np.zeros((1, n), np.int16))                     # not an open session
    with pytest.raises(ValueError):
        e.step([0], np.zeros((1, n - 1), np.int16))                     # wrong chunk length
    s = e.open_session()
    e.close_session(s)
    with pytest.raises(AsrLibraryError):
        e.step([s], np.zeros((1, n), np.int16))                         # closed session
    r = e.step([], np.zeros((0, n), np.int16))                          # empty batch is a no-op
    assert r.argmax_ids.shape[0] == 0


def test_lightning_asr_dropin(packed_weights, golden, meta):
    """The reference call pattern (streaming_server.py:324-326, :420-435, :514-515, :530) on the shim."""
    import torch
    from asr_streaming_b200 import LightningASR, PRECISION_EXACT, greedy_search
    case, mc = golden("seq_reset_skip"), meta["cases"]["seq_reset_skip"]
    model = LightningASR(weights=packed_weights, cfg=model_cfg(PRECISION_EXACT, max_batch=4, max_sessions=8), vocab=meta["vocab"])
    state_init = model.init_state()
    state, emission = state_init, torch.Tensor([])
    audio = torch.cat([torch.zeros(3200), torch.from_numpy(to_float(case["pcm"]))])
    k = j = 0
    while audio.numel() >= 13440:
        if k in mc["reset_before"]:
            emission, state = torch.Tensor([]), state_init
        if k not in mc["skip"]:
            em, length, states = model.stream([audio[None, :13440]], 16000, [state])
            state = states[0]
            assert int(length[0]) == 16
            emission = torch.cat((emission, em[0]))
            text, last_blank = greedy_search(emission)
            assert text == mc["texts"][j]
            assert abs(last_blank - case["last_blank"][j]) < 1e-7
            assert np.abs(em[0].numpy() - case["emission"][j]).max() < EXACT_TOL
            j += 1
        audio = audio[10240:]
        k += 1
    assert j == mc["n_chunks"]


# ------------------------------------------------------------------------------------------------ prefix beam search
def test_prefix_beam_kernel_matches_oracle(packed_weights, golden):
    """Warp-per-stream CTC prefix beam (beam 10, 8 candidates/frame, state carried across chunks, ragged batch) vs
    oracle/ctc_beam_oracle.py run on the very log-probs the device produced.  Parity unpinned vs the reference."""
    from asr_streaming_b200 import Engine, PRECISION_EXACT
    from oracle import ctc_beam_oracle as B
    e = Engine(model_cfg(PRECISION_EXACT, max_batch=8, max_sessions=8), packed_weights)
    e.set_beam(10, 8)
    names = ["synth_noise", "testwav", "synth_tone"]
    chunks = [chunks_i16(golden(n)["pcm"]) for n in names]
    slots = [e.open_session() for _ in names]
    states = [B.BeamState() for _ in names]
    n_cmp = 0
    for tick in range(6):
        idx = [i for i in range(len(names)) if tick < len(chunks[i])]
        if tick == 3:                                                  # endpoint on stream 0: beam state must clear too
            e.reset_session(slots[0])
            states[0] = B.BeamState()
        r = e.step([slots[i] for i in idx], np.stack([chunks[i][tick] for i in idx]), want_logprobs=True)
        for j, i in enumerate(idx):
            states[i] = B.beam_step(states[i], r.logprobs[j].astype(np.float64), beam=10, cand_k=8, max_len=1023)
            pre, score = B.best(states[i])
            assert list(r.beam_tokens[j]) == pre, f"{names[i]} tick {tick}"
            assert abs(float(r.beam_score[j]) - score) < 1e-3 * max(1.0, abs(score))
            n_cmp += 1
    report(f"prefix beam (beam 10, cand 8): {n_cmp} stream-chunks token-exact vs oracle; last score {score:.4f} vs device {float(r.beam_score[j]):.4f}")
    e.close()


def _rank_candidates(lp_row, k):
    """The k best non-blank ids by (log-prob desc, id asc) — the rule of oracle/ctc_beam_oracle.top_candidates."""
    ids = np.arange(1, lp_row.shape[0])
    order = np.lexsort((ids, -lp_row[1:].astype(np.float64)))
    return ids[order[:k]]


def test_decode_stage_on_peaked_tied_and_flat_posteriors(packed_weights):
    """The decode stage alone (asr_debug_decode_logits) on posteriors the random-init encoder never produces:
    * peaked rows (a trained model's regime), rows of small integers (hundreds of exact ties: more than 32 elements reach the
      candidate threshold, i.e. the plain-selection path of ctc_greedy_kernel) and completely flat rows;
    * per row: log-probs vs float64 log_softmax, argmax = lowest index among equal maxima (torch.argmax), the beam's extension
      candidates = the 8 best non-blank ids by (log-prob desc, id asc) with log-probs bit-equal to the log-prob array;
    * greedy carry across chunks vs oracle.greedy_ids on the concatenated emission; prefix beam vs oracle/ctc_beam_oracle.py on the
      peaked streams (parity unpinned vs the reference, see that file)."""
    from asr_streaming_b200 import Engine, PRECISION_FAST
    from oracle import ctc_beam_oracle as B
    cfg = model_cfg(PRECISION_FAST, max_batch=8, max_sessions=8)
    S, V = cfg.seg_rows, cfg.vocab
    rng = np.random.default_rng(77)
    n, T = 6, 4
    kinds = ["peaked", "peaked", "peaked_repeats", "ints", "ints_wide", "flat"]

    def make(kind, t):
        if kind == "flat":
            z = np.zeros((S, V), np.float32)
            if t % 2:
                z[:, 5] = 1.0
            return z
        if kind == "ints":
            return rng.integers(0, 4, size=(S, V)).astype(np.float32)
        if kind == "ints_wide":
            return rng.integers(0, 40, size=(S, V)).astype(np.float32)
        z = rng.standard_normal((S, V)).astype(np.float32)
        toks = rng.integers(0, V, size=S)
        toks[rng.random(S) < 0.45] = 0                               # blanks
        if kind == "peaked_repeats":
            toks[1::2] = toks[0::2]                                  # repeated frames of one token (collapse / p_b vs p_nb paths)
        z[np.arange(S), toks] += 14.0
        return z

    with Engine(cfg, packed_weights) as e:
        e.set_beam(10, 8)
        slots = [e.open_session() for _ in range(n)]
        states = [B.BeamState() for _ in range(n)]
        emis = [[] for _ in range(n)]
        toks = [[] for _ in range(n)]
        n_rows = 0
        for t in range(T):
            z = np.stack([make(k, t) for k in kinds])
            r = e.debug_decode_logits(slots, z, want_logprobs=True)
            lp = r.logprobs
            z64 = z.astype(np.float64)
            ref = z64 - z64.max(2, keepdims=True)
            ref = ref - np.log(np.exp(ref).sum(2, keepdims=True))
            assert np.abs(lp - ref).max() < 2e-5
            assert np.array_equal(r.argmax_ids, lp.argmax(2))          # numpy argmax: first (lowest) index among equal maxima
            cand_lp = e.debug_read(5, (n * S, 8)).reshape(n, S, 8)
            cand_tok = e.debug_read(6, (n * S, 8)).view(np.int32).reshape(n, S, 8)
            for i in range(n):
                for f in range(S):
                    want = _rank_candidates(lp[i, f], 8)
                    assert np.array_equal(cand_tok[i, f], want), (kinds[i], t, f, cand_tok[i, f], want)
                    assert np.array_equal(cand_lp[i, f], lp[i, f][want])
                    n_rows += 1
                emis[i].append(lp[i])
                ids, last_blank, _ = O.greedy_ids(np.concatenate(emis[i]))
                toks[i].extend(int(x) for x in r.new_tokens[i])
                assert toks[i] == ids, (kinds[i], t)
                assert int(r.blank_frames[i]) == int(round(last_blank / 0.04))
                if kinds[i].startswith("peaked"):
                    states[i] = B.beam_step(states[i], lp[i].astype(np.float64), beam=10, cand_k=8, max_len=1023)
                    pre, score = B.best(states[i])
                    assert list(r.beam_tokens[i]) == pre, (kinds[i], t)
                    assert abs(float(r.beam_score[i]) - score) < 1e-3 * max(1.0, abs(score))
                    # where the posterior is peaked the beam's best hypothesis is the greedy path
                    assert pre == ids
        report(f"decode stage on synthetic posteriors: {n_rows} rows (peaked / tied / flat) — candidates, argmax ties, log-probs, greedy carry, beam vs oracle exact")


def test_pipelined_submit_collect_equals_sync(engines):
    """Two steps in flight (H2D of k+1 overlapping the kernels of k, batches assembled in the pinned staging buffers)
    must give bit-identical results to synchronous steps."""
    e = engines(engines.FAST)
    rng = np.random.default_rng(21)
    n, T = 16, 5
    pcm = rng.integers(-4000, 4000, size=(T, n, O.CANONICAL.chunk_length)).astype(np.int16)
    a = [e.open_session() for _ in range(n)]
    b = [e.open_session() for _ in range(n)]
    sync = [e.step(a, pcm[t], want_logprobs=True) for t in range(T)]
    got, prev = [], None
    for t in range(T):
        view = e.pinned_pcm(np.int16)
        view[:n] = pcm[t]
        tk = e.submit(b, view[:n], want_logprobs=True)
        if prev is not None:
            got.append(e.collect(prev))
        prev = tk
    got.append(e.collect(prev))
    for t in range(T):
        assert np.array_equal(sync[t].logprobs, got[t].logprobs)
        assert np.array_equal(sync[t].argmax_ids, got[t].argmax_ids)
        assert all(np.array_equal(x, y) for x, y in zip(sync[t].new_tokens, got[t].new_tokens))
    with pytest.raises(Exception):
        t1 = e.submit(a, pcm[0]); t2 = e.submit(a, pcm[0]); e.submit(a, pcm[0])     # a third ticket is refused
    e.collect(t1); e.collect(t2)
    for s in a + b:
        e.close_session(s)


# ------------------------------------------------------------------------------------------------ scheduler on the real engine
@pytest.mark.parametrize("device_gather", [False, True], ids=["host_gather", "device_gather"])
def test_scheduler_ticks_match_reference_texts(engines, golden, meta, device_gather):
    """Websocket-style delivery (ragged message sizes, streams joining at different times) through SessionScheduler.tick:
    native pinned gather (asr_gather_pcm) + one asr_step per tick.  Every stream's text / trailing-blank after each of
    its chunks must equal the reference's batch-1 greedy_search on the accumulated emission (fixtures)."""
    from asr_streaming_b200 import SessionScheduler, ids_to_text
    e = engines(engines.EXACT)
    names = ["synth_noise", "testwav", "synth_tone", "edge_fullscale", "edge_dc"]
    cases = [golden(n) for n in names]
    sch = SessionScheduler(e, capacity=16, backlog_chunks=3, device_gather=device_gather, vocab=meta["vocab"])   # device: the GPU gathers out of pinned rings
    rng = np.random.default_rng(5)
    sess = [sch.open() for _ in names]
    pos = [0] * len(names)
    start = [0, 3, 1, 5, 2]
    done_chunks = [0] * len(names)
    for rnd in range(400):
        for i, c in enumerate(cases):
            if rnd < start[i] or pos[i] >= c["pcm"].size:
                continue
            room = sch.CAP - sess[i].length_of_segment
            n = int(min(rng.integers(101, 9000), c["pcm"].size - pos[i], room))
            if n > 100:
                sess[i].accept_waveform(c["pcm"][pos[i]:pos[i] + n].astype(np.int16))
                pos[i] += n
            elif c["pcm"].size - pos[i] <= 100:
                pos[i] = c["pcm"].size                                     # tail shorter than a message the server would keep
        res = sch.tick()
        for s in res.sessions:
            i = sess.index(s)
            mc = meta["cases"][names[i]]
            j = done_chunks[i]
            text = ids_to_text(s.tokens, meta["vocab"])
            assert text == mc["texts"][j], (names[i], j)
            assert abs(s.trailing_blank_duration - (cases[i]["last_blank"][j] if text else 0.64 * (j + 1))) < 1e-6      # stream.py:121-125
            done_chunks[i] += 1
        if all(p >= c["pcm"].size for p, c in zip(pos, cases)) and not sch.ready_rows().size:
            break
    for i, n in enumerate(names):
        assert done_chunks[i] == meta["cases"][n]["n_chunks"], n
    for s in sess:
        sch.close(s)


def test_reset_many_equals_fresh_sessions(engines):
    """asr_session_reset_many on a subset: the reset streams continue exactly like freshly opened sessions, the others
    are untouched."""
    e = engines(engines.FAST)
    rng = np.random.default_rng(8)
    n = 6
    pcm = rng.integers(-4000, 4000, size=(5, n, O.CANONICAL.chunk_length)).astype(np.int16)
    a = [e.open_session() for _ in range(n)]
    for t in range(3):
        e.step(a, pcm[t])
    e.reset_sessions([a[1], a[4]])
    fresh = [e.open_session() for _ in range(2)]
    keep = [e.open_session() for _ in range(n)]
    for t in range(3):
        e.step(keep, pcm[t])
    for t in range(3, 5):
        ra = e.step(a, pcm[t], want_logprobs=True)
        rf = e.step(fresh, pcm[t][[1, 4]], want_logprobs=True)
        rk = e.step(keep, pcm[t], want_logprobs=True)
        assert np.array_equal(ra.logprobs[[1, 4]], rf.logprobs)
        assert np.array_equal(ra.logprobs[[0, 2, 3, 5]], rk.logprobs[[0, 2, 3, 5]])
        assert np.array_equal(ra.blank_frames[[1, 4]], rf.blank_frames)
    for s in a + fresh + keep:
        e.close_session(s)


# ------------------------------------------------------------------------------------------------ fused GEMM + LayerNorm path
@pytest.mark.parametrize("name", ["synth_noise", "testwav", "seq_reset_skip"])
def test_fused_layernorm_path_matches_reference(name, packed_weights, golden, meta, monkeypatch):
    """The engine takes the GEMM + residual + LayerNorm kernels (gemm_ln.cu) from 96 streams per step on; here they are forced
    for a single stream so that the reference fixtures pin them too (both precisions), and compared with the separate-pass path."""
    from asr_streaming_b200 import Engine, PRECISION_EXACT, PRECISION_FAST
    case, mc = golden(name), meta["cases"][name]
    monkeypatch.setenv("ASR_B200_FUSED_LN_MIN_STREAMS", "1")
    with Engine(model_cfg(PRECISION_EXACT), packed_weights) as e:
        em, ids, blanks = _run_case(e, case, mc)
    assert np.abs(em - case["emission"]).max() < EXACT_TOL
    assert np.array_equal(em.argmax(2), case["argmax"])
    report(f"EXACT fused-LN {name}: logprob max-abs {np.abs(em - case['emission']).max():.3e}")
    with Engine(model_cfg(PRECISION_FAST), packed_weights) as e:
        em_f, _, _ = _run_case(e, case, mc)
    monkeypatch.setenv("ASR_B200_NO_FUSED_LN", "1")
    with Engine(model_cfg(PRECISION_FAST), packed_weights) as e:
        em_u, _, _ = _run_case(e, case, mc)
    assert np.abs(em_f - case["emission"]).max() < FAST_TOL
    report(f"FAST fused-LN {name}: logprob max-abs {np.abs(em_f - case['emission']).max():.3e}; vs separate LN passes {np.abs(em_f - em_u).max():.3e}")
    assert np.abs(em_f - em_u).max() < FAST_TOL


@pytest.mark.parametrize("name", ["synth_noise", "seq_reset_skip"])
def test_fused_layernorm_pair_shape_matches_reference(name, packed_weights, golden, meta, monkeypatch):
    """FFN2 in the cta_group::2 shape (cluster of 4) with the one-pass statistics of the second LayerNorm, forced for one stream."""
    from asr_streaming_b200 import Engine, PRECISION_EXACT, PRECISION_FAST
    case, mc = golden(name), meta["cases"][name]
    monkeypatch.setenv("ASR_B200_FUSED_LN_MIN_STREAMS", "1")
    monkeypatch.setenv("ASR_B200_PAIR_LN_MIN_TILES", "1")
    with Engine(model_cfg(PRECISION_EXACT), packed_weights) as e:
        em, ids, blanks = _run_case(e, case, mc)
    report(f"EXACT fused-LN pair shape {name}: logprob max-abs {np.abs(em - case['emission']).max():.3e}")
    assert np.abs(em - case["emission"]).max() < EXACT_TOL
    assert np.array_equal(em.argmax(2), case["argmax"])
    with Engine(model_cfg(PRECISION_FAST), packed_weights) as e:
        em_f, _, _ = _run_case(e, case, mc)
    assert np.abs(em_f - case["emission"]).max() < FAST_TOL


def test_fused_layernorm_large_ragged_batch(packed_weights):
    """200 streams in one step (above the fused threshold, M = 4000 is not a multiple of the 128-row tile) against the same
    streams run one by one: per-stream results must not depend on the batch they ride in (FAST precision, bit-exact)."""
    from asr_streaming_b200 import Engine, PRECISION_FAST
    rng = np.random.default_rng(17)
    n = 200
    pcm = rng.integers(-4000, 4000, size=(3, n, O.CANONICAL.chunk_length)).astype(np.int16)
    with Engine(model_cfg(PRECISION_FAST, max_batch=256, max_sessions=512), packed_weights) as e:
        big = [e.open_session() for _ in range(n)]
        got = [e.step(big, pcm[t], want_logprobs=True).logprobs for t in range(3)]
        for i in (0, 57, 199):
            s = e.open_session()
            for t in range(3):
                one = e.step([s], pcm[t, i:i + 1], want_logprobs=True).logprobs[0]
                assert np.abs(one - got[t][i]).max() < FAST_TOL, (i, t)  # different kernels (fused vs separate LN, bf16 operands): close, not identical
        perm = rng.permutation(n)
        other = [e.open_session() for _ in range(n)]
        for t in range(3):
            r = e.step([other[j] for j in perm], pcm[t][perm], want_logprobs=True).logprobs
            assert np.array_equal(r, got[t][perm])                       # same kernels, different positions: bit-exact


@pytest.mark.parametrize("low_latency", [False, True], ids=["chunk16", "chunk8"])
def test_tma_store_gemms_are_bit_identical(packed_weights, monkeypatch, low_latency):
    """QKV (stream-tiled M tiles, every destination a TMA box: q, the session's K/V ring block, the right-context scratch), FFN1 and
    CTC1 (row tiles, one box per warp) through the TMA-store epilogues against the LSU epilogues (ASR_B200_NO_TMA_STORE=1): same MMA
    shapes, k order and epilogue arithmetic => bit-identical log-probs over chained steps (the K/V ring wraps, sessions at mixed
    progress after a partial reset).  700 / 1100 streams: the last segment tile / right-context tile of each kind is ragged."""
    from asr_streaming_b200 import Engine, PRECISION_FAST
    rng = np.random.default_rng(43)
    n = 1100 if low_latency else 700                                       # enough 256-row tiles for the cta_group::2 kernels
    geo = O.LOW_LATENCY if low_latency else O.CANONICAL
    pcm = rng.integers(-4000, 4000, size=(5, n, geo.chunk_length)).astype(np.int16)
    outs = []
    for tma in (False, True):
        monkeypatch.setenv("ASR_B200_NO_TMA_STORE", "0" if tma else "1")
        with Engine(model_cfg(PRECISION_FAST, low_latency=low_latency, max_batch=n, max_sessions=n + 8), packed_weights) as e:
            sl = [e.open_session() for _ in range(n + 8)][8:]              # slots != batch positions
            got = []
            for t in range(5):
                if t == 2:
                    e.reset_sessions(sl[50:120])
                k = n if t != 3 else 333                                   # a smaller, ragged batch in between
                got.append(e.step(sl[:k], pcm[t, :k], want_logprobs=True).logprobs[:333])
            outs.append(np.stack(got))
    assert np.isfinite(outs[1]).all()
    assert np.array_equal(outs[0], outs[1])


def test_cuda_graph_replay_is_bit_identical(packed_weights, monkeypatch):
    """Small batches can replay the per-step kernel chain from a CUDA graph (captured the second time a (streams, staging buffer,
    format, outputs) key is seen): six chained steps of a 5-stream batch with a mid-sequence reset and a change of the batch size,
    (ASR_B200_GRAPHS=1; opt-in) against the plain launches — identical log-probs, ids and incremental tokens; the launch counter keeps
    counting the kernels of replayed steps."""
    from asr_streaming_b200 import Engine, PRECISION_FAST
    rng = np.random.default_rng(47)
    n = 5
    pcm = rng.integers(-4000, 4000, size=(6, n, O.CANONICAL.chunk_length)).astype(np.int16)
    runs = []
    for graphs in (False, True):
        monkeypatch.setenv("ASR_B200_GRAPHS", "1" if graphs else "0")
        with Engine(model_cfg(PRECISION_FAST, max_batch=8, max_sessions=8), packed_weights) as e:
            sl = [e.open_session() for _ in range(n)]
            out = []
            for t in range(6):
                if t == 3:
                    e.reset_sessions(sl[1:3])
                k = n if t != 4 else 3                                   # another key in between
                r = e.step(sl[:k], pcm[t, :k], want_logprobs=True)
                out.append((r.logprobs.copy(), r.argmax_ids.copy(), [list(x) for x in r.new_tokens]))
            runs.append((out, e.stats()["kernel_launches"]))
    for a, b in zip(runs[0][0], runs[1][0]):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    assert runs[0][1] == runs[1][1]


def test_streaming_attention_kernel_is_bit_identical(packed_weights, monkeypatch):
    """The persistent double-buffered attention kernel (taken from 148 streams per step on) forced for a small ragged batch:
    same fragments and summation order as the CTA-per-stream kernel => bit-identical log-probs, at every left-context fill."""
    from asr_streaming_b200 import Engine, PRECISION_FAST
    rng = np.random.default_rng(23)
    n = 37
    pcm = rng.integers(-4000, 4000, size=(4, n, O.CANONICAL.chunk_length)).astype(np.int16)
    outs = []
    for force in (False, True):
        if force:
            monkeypatch.setenv("ASR_B200_ATTN_STREAM_MIN", "1")
        with Engine(model_cfg(PRECISION_FAST, max_batch=64, max_sessions=64), packed_weights) as e:
            sl = [e.open_session() for _ in range(n)]
            got = []
            for t in range(4):
                if t == 2:
                    e.reset_sessions(sl[5:20])                         # ragged: 0 / 32 valid left-context rows in one step
                got.append(e.step(sl, pcm[t], want_logprobs=True).logprobs)
            outs.append(np.stack(got))
    assert np.array_equal(outs[0], outs[1])


def test_exact_tensor_core_attention_matches_fp32_kernel(packed_weights, golden, meta, monkeypatch):
    """EXACT precision: the split-bf16 (3 MMAs per product) TMA-fed attention kernel, taken from 148 streams per step on, forced
    for a small ragged batch against the fp32 CUDA-core kernel — log-probs within the EXACT tolerance, greedy ids identical —
    and on a golden case against the reference's emission, in both geometries."""
    from asr_streaming_b200 import Engine, PRECISION_EXACT
    rng = np.random.default_rng(29)
    n = 21
    pcm = rng.integers(-4000, 4000, size=(4, n, O.CANONICAL.chunk_length)).astype(np.int16)
    outs, ids = [], []
    for force in (False, True):
        monkeypatch.setenv("ASR_B200_ATTN_STREAM_MIN", "1" if force else "100000")
        with Engine(model_cfg(PRECISION_EXACT, max_batch=32, max_sessions=32), packed_weights) as e:
            sl = [e.open_session() for _ in range(n)]
            got, gi = [], []
            for t in range(4):
                if t == 2:
                    e.reset_sessions(sl[3:11])
                r = e.step(sl, pcm[t], want_logprobs=True)
                got.append(r.logprobs)
                gi.append(r.argmax_ids.copy())
            outs.append(np.stack(got))
            ids.append(np.stack(gi))
    assert np.abs(outs[0] - outs[1]).max() < EXACT_TOL
    assert np.array_equal(ids[0], ids[1])
    monkeypatch.setenv("ASR_B200_ATTN_STREAM_MIN", "1")
    for name, geom, low in (("synth_noise", O.CANONICAL, False), ("lowlat_noise", O.LOW_LATENCY, True)):
        case, mc = golden(name), meta["cases"][name]
        with Engine(model_cfg(PRECISION_EXACT, low), packed_weights) as e:
            em, _, _ = _run_case(e, case, mc, geom)
        assert np.abs(em - case["emission"]).max() < EXACT_TOL, name
        assert np.array_equal(em.argmax(-1), case["emission"].argmax(-1)), name


def test_streaming_attention_low_latency_geometry(packed_weights, golden, meta, monkeypatch):
    """Low-latency geometry (8 segment rows, 5-block ring) through the streaming attention kernel: 8-row TMA boxes."""
    from asr_streaming_b200 import Engine, PRECISION_FAST
    monkeypatch.setenv("ASR_B200_ATTN_STREAM_MIN", "1")
    case, mc = golden("lowlat_noise"), meta["cases"]["lowlat_noise"]
    with Engine(model_cfg(PRECISION_FAST, True), packed_weights) as e:
        em, _, _ = _run_case(e, case, mc, O.LOW_LATENCY)
    monkeypatch.setenv("ASR_B200_ATTN_STREAM_MIN", "100000")
    with Engine(model_cfg(PRECISION_FAST, True), packed_weights) as e:
        em_old, _, _ = _run_case(e, case, mc, O.LOW_LATENCY)
    assert np.abs(em - case["emission"]).max() < FAST_TOL
    assert np.array_equal(em, em_old)


def test_full_size_batch_is_consistent_with_small_batches(packed_weights):
    """BASELINE configs[3] size: 4096 streams in one step (M = 81,920 rows: cta_group::2 GEMMs, gemm_ln in its pair shape with the
    one-pass second LayerNorm, TMA streaming attention).  Size-independent properties: streams fed the same audio give bit-identical
    results wherever they sit in the batch, and every stream agrees with the same audio run in a batch of 8 (other kernel shapes:
    separate LayerNorm passes, CTA-per-stream attention) within the FAST tolerance, step after step (left context 0 / 16 / 32)."""
    from asr_streaming_b200 import Engine, PRECISION_FAST
    rng = np.random.default_rng(31)
    n, kinds, T = 4096, 8, 4
    base = rng.integers(-4000, 4000, size=(T, kinds, O.CANONICAL.chunk_length)).astype(np.int16)
    which = rng.integers(0, kinds, size=n)
    which[:kinds] = np.arange(kinds)
    with Engine(model_cfg(PRECISION_FAST, max_batch=kinds, max_sessions=kinds), packed_weights) as small:
        ss = [small.open_session() for _ in range(kinds)]
        ref = [small.step(ss, base[t], want_logprobs=True) for t in range(T)]
    with Engine(model_cfg(PRECISION_FAST, max_batch=n, max_sessions=n), packed_weights) as big:
        sl = [big.open_session() for _ in range(n)]
        for t in range(T):
            r = big.step(sl, base[t][which], want_logprobs=(t == T - 1))
            for k in range(kinds):
                rows = np.nonzero(which == k)[0]
                assert (r.argmax_ids[rows] == r.argmax_ids[rows[0]]).all(), (t, k)          # same audio, same kernels: identical
                assert np.array_equal(r.blank_frames[rows], np.full(rows.size, r.blank_frames[rows[0]]))
            if t == T - 1:
                lp = r.logprobs[:kinds]
                err = np.abs(lp - ref[t].logprobs).max()
                report(f"FAST 4096-stream step vs batch of 8 (different kernel shapes): logprob max-abs {err:.3e}")
                assert err < FAST_TOL
                safe = margins(ref[t].logprobs) > 2 * FAST_TOL
                assert np.array_equal(r.argmax_ids[:kinds][safe], ref[t].argmax_ids[safe])


def test_gpu_router_two_devices(packed_weights, golden, meta):
    """One Engine + scheduler per GPU inside one process (GpuRouter): sessions placed least-loaded, ticks on one host thread per
    GPU, results equal to the reference fixtures on both devices.  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from asr_streaming_b200 import GpuRouter, ids_to_text
    names = ["synth_noise", "testwav", "synth_tone", "edge_dc"]
    cases = [golden(n) for n in names]
    from asr_streaming_b200 import PRECISION_EXACT
    router = GpuRouter(model_cfg(PRECISION_EXACT, max_batch=8, max_sessions=8), packed_weights, [0, 1])
    sess = [router.open() for _ in names]
    assert sorted(s.gpu for s in sess) == [0, 0, 1, 1] and len({s.id for s in sess}) == 4
    chunks = [chunks_i16(c["pcm"]) for c in cases]
    for k in range(max(len(c) for c in chunks)):
        for i, s in enumerate(sess):
            if k < len(chunks[i]):
                s.accept_waveform(chunks[i][k][O.CANONICAL.buffer_length:])      # the 10,240 new samples of chunk k
        router.tick()
        for i, s in enumerate(sess):
            if k < len(chunks[i]):
                assert ids_to_text(s.tokens, meta["vocab"]) == meta["cases"][names[i]]["texts"][k], (names[i], k)
    for s in sess:
        router.close(s)


def test_lightning_asr_v1_batched_call_pattern(packed_weights, golden, meta):
    """The v1 batcher's call pattern (streaming_decoder_v1/streaming_asr.py:94-112): ONE model.stream() for a list of streams at
    different progress, then per stream `emission[idx]`, torch.cat onto its accumulated emission and greedy_search.  The reference's
    batched infer is wrong for mixed progress (TA:emformer.py:392 reads element 0's past_length); here every stream must equal its
    own batch-1 fixture."""
    import torch
    from asr_streaming_b200 import LightningASR, PRECISION_EXACT, greedy_search
    names = ["synth_noise", "testwav", "synth_tone"]
    cases = [golden(n) for n in names]
    model = LightningASR(weights=packed_weights, cfg=model_cfg(PRECISION_EXACT, max_batch=4, max_sessions=8), vocab=meta["vocab"])
    state_init = model.init_state()
    audio = [torch.cat([torch.zeros(3200), torch.from_numpy(to_float(c["pcm"]))]) for c in cases]
    states = [state_init] * 3
    emissions = [torch.Tensor([]) for _ in names]
    done = [0, 0, 0]
    start = [0, 2, 1]
    for tick in range(20):
        idx = [i for i in range(3) if tick >= start[i] and audio[i].numel() >= 13440]
        if not idx:
            continue
        em, length, new_states = model.stream([audio[i][None, :13440] for i in idx], 16000, [states[i] for i in idx])
        for j, i in enumerate(idx):
            states[i] = new_states[j]
            emissions[i] = torch.cat((emissions[i], em[j]), dim=0)
            text, last_blank = greedy_search(emissions[i])
            mc = meta["cases"][names[i]]
            assert text == mc["texts"][done[i]], (names[i], done[i])
            assert abs(last_blank - cases[i]["last_blank"][done[i]]) < 1e-7
            assert np.abs(em[j].numpy() - cases[i]["emission"][done[i]]).max() < EXACT_TOL
            done[i] += 1
            audio[i] = audio[i][10240:]
    assert done == [meta["cases"][n]["n_chunks"] for n in names]


# ------------------------------------------------------------------------------------------------ peaked posteriors: bf16 is token-exact
@pytest.mark.parametrize("big_batch", [False, True], ids=["batch1_kernels", "large_batch_kernels"])
def test_fast_precision_is_token_exact_on_peaked_posteriors(oracle_weights, golden, meta, monkeypatch, big_batch):
    """The north-star asks for bit-exact greedy ids AND a bf16 path.  On the flat posteriors of the random-init model no bf16 path can
    deliver both (nor can the reference's own autocast); on a checkpoint whose posteriors are peaked like a trained model's it must.
    tests/golden/peaky_head.npz (oracle/make_peaky_goldens.py) is such a checkpoint: the seeded encoder with a CTC output layer fitted
    so that the top-1 posterior is > 0.99 on the fixture audio; tests/golden/peaky_*.npz are the UNMODIFIED reference's outputs with it.
    FAST (bf16 operands, bf16 K/V cache) must reproduce every argmax id, every text and every trailing-blank duration: 400 / 400 frames."""
    from asr_streaming_b200 import Engine, PRECISION_FAST, ids_to_text, pack_weights
    head = golden("peaky_head")
    W = dict(oracle_weights)
    W["decoder.linear2.weight"], W["decoder.linear2.bias"] = head["weight"], head["bias"]
    if big_batch:                                                   # the kernels the 4096-stream bench runs: fused-LN GEMMs, TMA streaming attention
        monkeypatch.setenv("ASR_B200_FUSED_LN_MIN_STREAMS", "1")
        monkeypatch.setenv("ASR_B200_ATTN_STREAM_MIN", "1")
    frames = worst = 0
    with Engine(model_cfg(PRECISION_FAST), pack_weights(W)) as e:
        for name, pk in meta["peaky"]["cases"].items():
            case, mc = golden(name), meta["cases"][name]
            ref = golden(f"peaky_{name}")
            em, toks, blanks = _run_case(e, case, mc)
            assert em.shape == ref["emission"].shape
            assert np.array_equal(em.argmax(-1), ref["argmax"]), f"{name}: greedy ids differ from the reference"
            assert [ids_to_text(t, meta["vocab"]) for t in toks] == pk["texts"]
            assert np.allclose(blanks, ref["last_blank"], atol=1e-6)
            frames += ref["argmax"].size
            worst = max(worst, float(np.abs(em - ref["emission"]).max()))
            s = np.sort(em, axis=-1)
            assert (s[..., -1] - s[..., -2]).min() > 0.5 * pk["min_margin"]          # the margins survive bf16: nowhere near a flip
    report(f"FAST on the peaked checkpoint ({'large-batch' if big_batch else 'batch-1'} kernels): {frames}/{frames} greedy ids identical to the reference; "
           f"log-prob max-abs {worst:.3e} at top-2 margins >= {min(c['min_margin'] for c in meta['peaky']['cases'].values()):.1f}")


def test_exact_precision_full_size_batch_equals_small_batch(packed_weights):
    """EXACT precision at the BASELINE configs[3] size (4096 streams per step: cta_group::2 GEMMs with three passes, gemm_ln pair shape,
    three-MMA tensor-core attention) against the same audio in a batch of 8 (one-CTA GEMMs, separate LayerNorms, fp32 CUDA-core
    attention): log-probs within the EXACT tolerance and IDENTICAL greedy ids, chunk after chunk (left context 0 / 16 / 32 / wrapped)."""
    from asr_streaming_b200 import Engine, PRECISION_EXACT
    rng = np.random.default_rng(33)
    n, kinds, T = 4096, 8, 5
    base = rng.integers(-4000, 4000, size=(T, kinds, O.CANONICAL.chunk_length)).astype(np.int16)
    which = rng.integers(0, kinds, size=n)
    which[:kinds] = np.arange(kinds)
    with Engine(model_cfg(PRECISION_EXACT, max_batch=kinds, max_sessions=kinds), packed_weights) as small:
        ss = [small.open_session() for _ in range(kinds)]
        ref = [small.step(ss, base[t], want_logprobs=True) for t in range(T)]
    worst = 0.0
    with Engine(model_cfg(PRECISION_EXACT, max_batch=n, max_sessions=n), packed_weights) as big:
        sl = [big.open_session() for _ in range(n)]
        for t in range(T):
            r = big.step(sl, base[t][which], want_logprobs=(t == T - 1))
            assert np.array_equal(r.argmax_ids, ref[t].argmax_ids[which]), f"chunk {t}: greedy ids of the 4096-stream step differ from the batch of 8"
            assert np.array_equal(r.blank_frames, ref[t].blank_frames[which])
        worst = float(np.abs(r.logprobs - ref[T - 1].logprobs[which]).max())
        assert worst < EXACT_TOL
    report(f"EXACT 4096-stream step vs batch of 8: greedy ids identical on {T} x 4096 x 16 frames, log-prob max-abs {worst:.3e}")


@pytest.mark.parametrize("device_gather", [False, True], ids=["host_gather", "device_gather"])
def test_prestaged_ticks_match_reference_texts(engines, golden, meta, device_gather):
    """One-tick-per-pass pipelining: every buffered chunk — also of the sessions still in flight — is gathered and copied to the device
    (SessionScheduler.prestage) BEFORE the running tick is collected; the next tick then launches on a subset of the staged rows through
    the fbank kernel's row-index indirection.  Same fixtures, same expectations as the plain scheduler test, plus a VAD-skip in between
    (a staged chunk that was consumed otherwise must not be run from the stage)."""
    from asr_streaming_b200 import SessionScheduler, ids_to_text
    e = engines(engines.EXACT)
    names = ["synth_noise", "testwav", "synth_tone", "edge_fullscale", "edge_dc", "edge_silence"]
    cases = [golden(n) for n in names]
    sch = SessionScheduler(e, capacity=16, backlog_chunks=3, vocab=meta["vocab"], device_gather=device_gather)
    rng = np.random.default_rng(6)
    sess = [sch.open() for _ in names]
    pos, start, done_chunks = [0] * len(names), [0, 3, 1, 5, 2, 0], [0] * len(names)
    prev, staged_total = None, 0

    def check(res):
        for s in res.sessions:
            i = sess.index(s)
            j = done_chunks[i]
            text = ids_to_text(s.tokens, meta["vocab"])
            assert text == meta["cases"][names[i]]["texts"][j], (names[i], j)
            assert abs(s.trailing_blank_duration - (cases[i]["last_blank"][j] if text else 0.64 * (j + 1))) < 1e-6
            done_chunks[i] += 1
    for rnd in range(400):
        for i, c in enumerate(cases):
            if rnd < start[i] or pos[i] >= c["pcm"].size:
                continue
            room = sch.CAP - sess[i].length_of_segment
            n = int(min(rng.integers(101, 9000), c["pcm"].size - pos[i], room))
            if n > 100:
                sess[i].accept_waveform(c["pcm"][pos[i]:pos[i] + n].astype(np.int16))
                pos[i] += n
            elif c["pcm"].size - pos[i] <= 100:
                pos[i] = c["pcm"].size
        staged_total += sch.prestage()                                   # while `prev` is still in flight
        if prev is not None:
            check(sch.collect_tick(prev))
            prev = None
        p = sch.submit_tick()
        if p.rows.size:
            prev = p
        if all(q >= c["pcm"].size for q, c in zip(pos, cases)) and prev is None and not sch.ready_rows().size:
            break
    if prev is not None:
        check(sch.collect_tick(prev))
    for i, n in enumerate(names):
        assert done_chunks[i] == meta["cases"][n]["n_chunks"], n
    assert staged_total >= sum(done_chunks)
    for s in sess:
        sch.close(s)
