"""C-ABI misuse is rejected loudly, overflow is reported, and the long-form low-latency workload (BASELINE configs[4]) stays exact."""
import numpy as np
import pytest

from oracle import lightspeech_oracle as O
from helpers import model_cfg, report

pytestmark = pytest.mark.gpu


def test_duplicate_slots_and_in_flight_misuse_are_rejected(packed_weights):
    from asr_streaming_b200 import Engine, PRECISION_FAST
    from asr_streaming_b200._lib import AsrLibraryError
    rng = np.random.default_rng(1)
    with Engine(model_cfg(PRECISION_FAST, max_batch=8, max_sessions=8), packed_weights) as e:
        a, b = e.open_session(), e.open_session()
        pcm = rng.integers(-3000, 3000, size=(3, O.CANONICAL.chunk_length)).astype(np.int16)
        with pytest.raises(AsrLibraryError, match="twice"):                       # two rows of one slot would race on its K/V ring
            e.step([a, b, a], pcm)
        ref = e.step([a, b], pcm[:2], want_logprobs=True).logprobs                # the rejected call changed nothing
        t = e.submit([a], pcm[:1])
        with pytest.raises(AsrLibraryError, match="rides a submitted step"):
            e.close_session(a)
        with pytest.raises(AsrLibraryError, match="in flight"):
            e.stage([b], pcm[:1])
        with pytest.raises(AsrLibraryError, match="in flight"):
            e.fbank(pcm[:1])
        e.close_session(b)                                                        # not in flight: fine
        e.collect(t)
        e.close_session(a)
        c, d = e.open_session(), e.open_session()                                 # open resets asynchronously (no host sync): state must be fresh
        got = e.step([c, d], pcm[:2], want_logprobs=True).logprobs
        assert np.array_equal(got, ref)


def test_has_text_follows_the_silent_id_mask(packed_weights):
    """AsrStepOut.has_text = the `if text:` of Stream.update_stream (stream.py:121): ids that render to "" do not count."""
    from asr_streaming_b200 import Engine, PRECISION_FAST
    rng = np.random.default_rng(2)
    pcm = rng.integers(-3000, 3000, size=(2, 4, O.CANONICAL.chunk_length)).astype(np.int16)
    with Engine(model_cfg(PRECISION_FAST, max_batch=4, max_sessions=4), packed_weights) as e:
        sl = [e.open_session() for _ in range(4)]
        r = e.step(sl, pcm[0])
        assert r.has_token.all() and np.array_equal(r.has_text, r.has_token)      # random-init model: tokens everywhere, default mask {0, 1}
        seen = np.unique(r.argmax_ids)
        e.reset_sessions(sl)
        e.set_silent_ids(list(range(O.CANONICAL.vocab)))                          # everything renders to "": no text, ever
        r = e.step(sl, pcm[0])
        assert r.has_token.all() and not r.has_text.any()
        e.reset_sessions(sl)
        e.set_silent_ids([0, 1] + [int(i) for i in seen if i > 1][1:])             # all but one of the ids this audio produces
        r = e.step(sl, pcm[0])
        keep = [int(i) for i in seen if i > 1][0]
        assert np.array_equal(r.has_text, (r.argmax_ids == keep).any(axis=1))
        r2 = e.step(sl, pcm[1])                                                    # the flag is carried over the segment
        assert (r2.has_text >= r.has_text).all()


def test_beam_truncation_and_token_overflow_are_flagged(packed_weights):
    """The random-init model emits a token every other frame: without endpoint rules a segment passes ASR_BEAM_MAX_LEN - 1 = 1023 beam tokens
    (and MAX_TOKENS = 1024 greedy tokens) after ~145 chunks.  The step flags the truncation, the scheduler reports it; nothing is overwritten."""
    from asr_streaming_b200 import Engine, PRECISION_FAST, SessionScheduler
    from asr_streaming_b200.engine import BEAM_MAX_LEN, FLAG_BEAM_TRUNCATED
    from asr_streaming_b200.scheduler import MAX_TOKENS
    rng = np.random.default_rng(3)
    cfg = model_cfg(PRECISION_FAST, max_batch=2, max_sessions=2)
    with Engine(cfg, packed_weights) as e:
        e.set_beam(10, 8)
        sch = SessionScheduler(e, capacity=2, backlog_chunks=2)                   # no endpoint rules: the segment never ends
        s = sch.open()
        first_flag = first_over = None
        lens = []
        for k in range(200):
            s.accept_waveform(rng.integers(-4000, 4000, size=cfg.segment_length).astype(np.int16))
            res = sch.tick()
            assert len(res) == 1
            lens.append(int(res.step.beam_len[0]))
            if first_flag is None and res.step.flags[0] & FLAG_BEAM_TRUNCATED:
                first_flag = k
            if first_over is None and res.overflow[0]:
                first_over = k
            if res.overflow[0]:
                assert res.step.flags[0] & FLAG_BEAM_TRUNCATED or len(s.tokens) == MAX_TOKENS
            if first_flag is not None and k >= first_flag + 3:
                break
        assert max(lens) == BEAM_MAX_LEN - 1 and first_flag is not None and lens[first_flag] == BEAM_MAX_LEN - 1
        assert first_over is not None and first_over <= first_flag
        assert all(l < BEAM_MAX_LEN - 1 for l in lens[:first_flag])
        hyp = res.beam_row(0)
        assert hyp.size == BEAM_MAX_LEN - 1 and (hyp > 0).all()
        report(f"beam truncation flagged at chunk {first_flag} (hypothesis {lens[first_flag]} tokens), scheduler overflow from chunk {first_over}")


@pytest.mark.timeout(1500)
def test_long_form_low_latency_soak(packed_weights):
    """BASELINE configs[4] per-GPU share, long form: 256 low-latency streams (chunk_size = 8, 320 ms chunks) x 1875 chunks = 10 minutes of
    audio each, driven by the native scheduler with two ticks in flight, the energy gate and the endpoint rules (forced rule4 endpoints at
    40 s, asr-online.yaml:103-107).  The K/V ring wraps ~1250 times per session, the audio rings compact hundreds of times, segments end
    and restart.  At the end three sampled sessions are replayed through the batch-1 reference port (torchaudio Emformer + the reference
    glue, oracle/torch_ref_port.py) from their last endpoint: tokens, trailing silence and counters must be EXACT."""
    import torch
    from asr_streaming_b200 import Engine, PRECISION_EXACT, SessionScheduler
    from asr_streaming_b200.endpoint import EndpointRules
    from asr_streaming_b200.scheduler import native_energy_gate
    from oracle.torch_ref_port import TorchRefPort
    geo = O.LOW_LATENCY
    n, n_chunks, seg = 256, 1875, geo.segment_length
    cfg = model_cfg(PRECISION_EXACT, low_latency=True, max_batch=n, max_sessions=n)
    rng = np.random.default_rng(77)
    # per stream: a speech / silence pattern in whole chunks (speech = noise + tone at 10 % of full scale, silence = zeros) from a small pool
    pool = (0.1 * 32768 * rng.standard_normal((8, 64 * seg)) + 1500 * np.sin(np.arange(64 * seg) * 0.17)).clip(-32768, 32767).astype(np.int16)
    speech = rng.random((n, n_chunks)) < 0.8
    run = rng.random((n, n_chunks)) < 0.9
    for k in range(1, n_chunks):                                                  # sticky pattern: runs of speech / silence
        speech[:, k] = np.where(run[:, k], speech[:, k - 1], speech[:, k])
    shift = rng.integers(0, 60 * seg, size=n)

    def chunk_audio(i, k):                                                        # the seg new samples of chunk k of stream i
        if not speech[i, k]:
            return np.zeros(seg, np.int16)
        o = (shift[i] + k * seg) % (60 * seg)
        return pool[i % 8, o:o + seg]

    sample = [3, 100, 255]
    log = {i: [] for i in sample}                                                 # per sampled stream: (chunk index, run?, final?)
    with Engine(cfg, packed_weights) as e:
        sch = SessionScheduler(e, capacity=n, backlog_chunks=3, endpoint_rules=EndpointRules())
        ss = [sch.open() for _ in range(n)]
        gate = native_energy_gate()
        done = np.zeros(n, np.int64)
        n_final = n_skip = 0
        prev = None

        def note(res):
            nonlocal n_final, n_skip
            for r in res.skipped_rows:
                if int(r) in log:
                    log[int(r)].append((int(done[r]), False, ss[int(r)].id in res.final_tokens))
                done[r] += 1
            n_skip += int(res.skipped_rows.size)
            for j, r in enumerate(res.rows):
                if int(r) in log:
                    log[int(r)].append((int(done[r]), True, bool(res.final[j])))
                done[r] += 1
            n_final += len(res.final_tokens)
        for k in range(n_chunks):
            block = np.stack([chunk_audio(i, k) for i in range(n)])
            for i in range(n):                                                    # websocket-style delivery (asr_sched_accept: compaction path)
                ss[i].accept_waveform(block[i])
            while True:
                p = sch.submit_tick(gate=gate)
                if prev is not None:
                    note(sch.collect_tick(prev))
                    prev = None
                if p.rows.size:
                    prev = p
                else:
                    note(p.res)
                    if not sch.ready_rows().size:
                        break
        if prev is not None:
            note(sch.collect_tick(prev))
        assert (done == n_chunks).all()
        assert n_final > n * 10 and n_skip > n * 20, (n_final, n_skip)          # endpoints and VAD skips really happened
        W = O.make_weights(1234)
        port = TorchRefPort(W, geo)
        vocab = ["-", "|"] + [f"<{i}>" for i in range(2, geo.vocab)]
        from oracle.torch_ref_port import greedy_search
        checked = 0
        for i in sample:
            ev = log[i]
            assert [c for c, _, _ in ev] == list(range(n_chunks))
            last_end = max([j for j, (_, _, f) in enumerate(ev) if f], default=-1)
            seg_events = ev[last_end + 1:]
            state, em, trailing, contain, processed = port.init_state(), torch.zeros(0, geo.vocab), 0.0, False, 0
            for c, ran, _ in seg_events:
                processed += 1
                if not ran:
                    trailing = round(trailing + 0.32, 2)
                    continue
                lo = c * seg - geo.buffer_length
                win = np.concatenate([chunk_audio(i, c - 1)[lo - (c - 1) * seg:] if c > 0 else np.zeros(geo.buffer_length, np.int16), chunk_audio(i, c)])
                x = torch.from_numpy(win.astype(np.float32) / np.float32(32768.0))[None]
                out, _, st = port.stream([x], 16000, [state])
                state = st[0]
                em = torch.cat((em, out[0]))
                text, last_blank = greedy_search(em, vocab)
                if text:
                    trailing, contain = last_blank, True
                else:
                    trailing += 0.32
                trailing = round(trailing, 2)
            s = ss[i]
            ids = torch.unique_consecutive(torch.argmax(em, dim=1)) if em.numel() else torch.zeros(0, dtype=torch.long)
            assert s.tokens == [int(t) for t in ids if t != 0], f"stream {i}: tokens of the last segment ({len(seg_events)} chunks) differ"
            assert abs(s.trailing_blank_duration - trailing) < 1e-9 and s.is_contain_token == contain and s.chunk_processed == processed
            checked += len(seg_events)
        report(f"long-form soak: {n} low-latency streams x {n_chunks} chunks, {n_final} endpoints, {n_skip} VAD skips; {checked} chunks of 3 sampled sessions replayed through the reference port: exact")
        for s in ss:
            sch.close(s)
