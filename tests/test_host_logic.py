"""CPU-only tests: the C-ABI library loads and exports every symbol include/asr_b200.h declares (no compute calls),
geometry / weight packing, and the ragged session scheduler + per-GPU partitioning (host logic, fake engine)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import asr_streaming_b200 as A
from asr_streaming_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "asr_b200.h")).read()
    return sorted(set(re.findall(r"ASR_API\s+[\w\s\*]+?\b(asr_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    decl = _declared_symbols()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/asr_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == decl, "ctypes binding and header disagree"
    assert lib.asr_abi_version() == _lib.ABI_VERSION


def test_config_geometry_matches_reference_audio_config():
    c = _lib.AsrConfigC()
    lib = _lib.load_library()
    assert lib.asr_default_config(C.byref(c), 0) == 0
    ch, seg, rows = C.c_int32(), C.c_int32(), C.c_int32()
    assert lib.asr_chunk_geometry(C.byref(c), C.byref(ch), C.byref(seg), C.byref(rows)) == 0
    ac = A.AudioConfig()                                  # reference arithmetic, utils.py:9-23
    assert (ch.value, seg.value, rows.value) == (ac.chunk_length, ac.segment_length, 16) == (13440, 10240, 16)
    assert ac.buffer_length == 3200 and ac.hop_length == 160
    mc = A.ModelConfig()
    assert (mc.chunk_length, mc.frames, mc.rows, mc.seg_rows, mc.rc_rows) == (13440, 80, 20, 16, 4)
    ll = A.ModelConfig(segment_size=32)
    assert (ll.chunk_length, ll.frames, ll.rows, ll.seg_rows) == (8320, 48, 12, 8)
    assert lib.asr_default_config(C.byref(c), 1) == 0 and c.segment_size == 32


def test_weights_count_and_packing(oracle_weights):
    lib = _lib.load_library()
    c = _lib.AsrConfigC()
    lib.asr_default_config(C.byref(c), 0)
    n = C.c_uint64()
    assert lib.asr_weights_count(C.byref(c), C.byref(n)) == 0
    blob = A.pack_weights(oracle_weights)
    assert blob.dtype == np.float32 and blob.size == n.value == 63_759_652
    # layout spot checks: input_linear first, q rows before kv rows inside Wqkv
    assert np.array_equal(blob[:128 * 128], oracle_weights["encoder.input_linear.weight"].reshape(-1))
    p = "encoder.encoder_layers.emformer_layers.0."
    off = 128 * 128
    assert np.array_equal(blob[off:off + 512 * 512], oracle_weights[p + "attention.emb_to_query.weight"].reshape(-1))
    assert np.array_equal(blob[off + 512 * 512:off + 3 * 512 * 512], oracle_weights[p + "attention.emb_to_key_value.weight"].reshape(-1))
    assert np.array_equal(blob[-804:], oracle_weights["decoder.linear2.bias"])
    bad = dict(oracle_weights)
    bad["decoder.linear2.bias"] = bad["decoder.linear2.bias"][:-1]
    with pytest.raises(ValueError):
        A.pack_weights(bad)
    del bad["decoder.linear2.bias"]
    with pytest.raises(KeyError):
        A.pack_weights(bad)


def test_bad_config_is_rejected_without_gpu():
    lib = _lib.load_library()
    c = _lib.AsrConfigC()
    lib.asr_default_config(C.byref(c), 0)
    c.segment_size = 62                                    # 78 fbank frames != 19 rows * stride 4
    n = C.c_uint64()
    assert lib.asr_weights_count(C.byref(c), C.byref(n)) != 0
    assert b"geometry" in lib.asr_last_error()
    c.segment_size = 64
    c.abi_version = 99
    assert lib.asr_weights_count(C.byref(c), C.byref(n)) != 0


def test_engine_create_fails_loudly_without_gpu(packed_weights):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(A.AsrLibraryError) as ei:
        A.Engine(A.ModelConfig(max_batch=2, max_sessions=2), packed_weights)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)
    with pytest.raises(RuntimeError):
        A.LightningASR(weights=packed_weights, device="cpu")


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(A.AsrLibraryError):
        _lib.load_library(str(tmp_path / "nope.so"))


def test_ids_to_text_matches_reference_rules():
    vocab = ["-", "|", "a", "b", "<<", ">>", "c"]
    assert A.ids_to_text([2, 1, 1, 3, 4, 6, 5, 1], vocab) == "a bc"
    assert A.ids_to_text([], vocab) == ""


# ------------------------------------------------------------------------------------------------ scheduler (fake engine)
class FakeEngine:
    """Stands in for Engine on the CPU with the interface the scheduler drives (open / close / reset_sessions / set_silent_ids / submit /
    collect): records every step's (slots, pcm) and emits one deterministic token per chunk."""

    def __init__(self, cfg):
        self.cfg, self.calls, self._next, self.open_slots, self.resets = cfg, [], 0, set(), []
        self.tickets, self.busy, self.silent = {}, set(), None

    def open_session(self):
        s = self._next
        self._next += 1
        self.open_slots.add(s)
        return s

    def close_session(self, s):
        self.open_slots.remove(s)

    def reset_sessions(self, slots):
        self.resets.extend(int(s) for s in slots)

    def set_silent_ids(self, ids):
        self.silent = set(int(i) for i in ids)

    def step(self, slots, pcm, want_logprobs=False):
        slots = list(slots)
        assert len(set(slots)) == len(slots), "a session may appear at most once per step"
        assert len(slots) <= self.cfg.max_batch and pcm.shape == (len(slots), self.cfg.chunk_length)
        self.calls.append((slots, pcm.copy()))
        n, S = len(slots), self.cfg.seg_rows
        new = [np.array([2 + (int(pcm[i, -1]) % 7)], np.int32) for i in range(n)]
        return A.StepResult(np.zeros((n, S), np.int32), new, np.full(n, 3, np.int32), np.ones(n, bool), None)

    # pipelined form (results computed at submit, delivered at collect), at most two tickets in flight
    def submit(self, slots, pcm, want_logprobs=False):
        slots = [int(x) for x in slots]
        assert len(self.tickets) < 2, "more than two steps in flight"
        assert not (self.busy & set(slots)), "a session was submitted again before its previous chunk was collected"
        self.busy |= set(slots)
        t = len(self.calls)
        self.tickets[t] = (slots, self.step(slots, pcm, want_logprobs))
        return t

    def collect(self, t):
        slots, out = self.tickets.pop(t)
        self.busy -= set(slots)
        return out


def test_scheduler_ragged_batching_and_buffer_semantics():
    cfg = A.ModelConfig(max_batch=4, max_sessions=16)
    eng = FakeEngine(cfg)
    sch = A.SessionScheduler(eng)
    ss = [sch.open() for _ in range(6)]
    rng = np.random.default_rng(0)
    audio = [rng.integers(-1000, 1000, size=n).astype(np.int16) for n in (10240, 30000, 5000, 10240 * 3, 50, 20480)]
    for s, a in zip(ss, audio):
        s.accept_waveform(a)
    assert ss[4].length_of_segment == cfg.buffer_length            # <= 100 samples dropped (stream.py:82)
    ready = [s.id for s in sch.ready_sessions()]
    assert ready == [0, 1, 3, 5]                                     # 3200 + n >= 13440, capped at max_batch
    out = sch.tick()
    assert [s.id for s, _, _ in out] == [0, 1, 3, 5]
    slots, pcm = eng.calls[-1]
    # chunk k = [k*10240 - 3200, k*10240 + 10240) with 3200 leading zeros (stream.py:23, :159)
    assert np.array_equal(pcm[0][:3200], np.zeros(3200, np.int16)) and np.array_equal(pcm[0][3200:], audio[0][:10240])
    assert ss[0].length_of_segment == 3200 and not ss[0].ready()
    out = sch.tick()                                                  # second tick: only streams with another full chunk
    assert [s.id for s, _, _ in out] == [1, 3, 5]                    # 5: 3200 + 20480 - 10240 = 13440, exactly one more chunk
    slots, pcm = eng.calls[-1]
    assert np.array_equal(pcm[0], audio[1][10240 - 3200:20480])
    assert ss[1].tokens and ss[1].chunk_processed == 2 and ss[1].n_frames == 32
    assert abs(ss[1].trailing_blank_duration - float(np.float32(3) * np.float32(0.04))) < 1e-9
    sch.reset(ss[1])
    assert eng.resets == [ss[1].slot] and ss[1].tokens == [] and ss[1].segment == 1
    sch.close(ss[2])
    assert ss[2].slot not in eng.open_slots


def test_scheduler_round_robin_fairness_and_vad_gate():
    cfg = A.ModelConfig(max_batch=2, max_sessions=8)
    eng = FakeEngine(cfg)
    sch = A.SessionScheduler(eng)
    ss = [sch.open() for _ in range(5)]
    for s in ss:
        s.accept_waveform(np.ones(10240 * 3, np.int16))
    served = []
    for _ in range(5):
        served += [s.id for s, _, _ in sch.tick()]
    assert served[:5] == [0, 1, 2, 3, 4]                              # nobody starves behind a backlog
    assert sorted(served) == sorted(served[:5] * 2)
    # VAD gate: gated-out chunk advances the buffer and the silence clock, not the encoder (stream.py:183-189)
    s = sch.open()
    s.accept_waveform(np.zeros(10240, np.int16))
    out = []
    for _ in range(3):                                                # the newcomer queues behind the remaining backlog
        out += list(sch.tick(gate=lambda sess, chunk: bool(np.abs(chunk).max() > 0)))
    assert s.id not in [x.id for x, _, _ in out] and s.chunk_processed == 1 and abs(s.trailing_blank_duration - 0.64) < 1e-9
    assert s.length_of_segment == cfg.buffer_length


def _scalar_rule_activated(rule, trailing_silence, utterance_length, relative_cost):
    """Scalar restatement of online_endpoint.py:42-66 (checker for the vectorised table)."""
    contains_nonsilence = utterance_length > trailing_silence
    return ((contains_nonsilence or not rule.must_contain_nonsilence) and trailing_silence >= rule.min_trailing_silence
            and relative_cost < rule.max_relative_cost and utterance_length >= rule.min_utterance_length)


def test_endpoint_rules_vectorised_equals_scalar_first_match():
    from asr_streaming_b200.endpoint import DEFAULT_RULES, EndpointRules, OnlineEndpointRule, detect_endpointing, load_endpointing_rule
    rules = dict(DEFAULT_RULES)
    rules["free"] = OnlineEndpointRule(False, 5.0, 0.0, float("inf"))          # a rule that fires on pure silence
    tab = EndpointRules(rules)
    rng = np.random.default_rng(3)
    n = 4000
    utt = np.round(rng.integers(0, 70, n) * 0.64, 2)
    sil = np.round(np.minimum(utt, rng.integers(0, 30, n) * 0.04 + rng.integers(0, 10, n) * 0.64), 2)
    cost = rng.choice([0.5, 1.99, 2.0, 4.9, 5.0, 7.5, 8.0, 10.0, float("inf")], n)
    fired, which = tab.detect(utt, sil, cost)
    names = list(rules)
    for i in range(n):
        exp = next((k for k, name in enumerate(names) if _scalar_rule_activated(rules[name], sil[i], utt[i], cost[i])), -1)
        assert which[i] == exp and fired[i] == (exp >= 0)
    assert fired.any() and not fired.all() and len(set(which.tolist())) > 4
    ok, name, over = detect_endpointing(DEFAULT_RULES, 6.4, 1.0, 10.0)
    assert (ok, name) == (True, "rule1.1") and abs(over) < 1e-12
    assert detect_endpointing(DEFAULT_RULES, 6.4, 0.96, 10.0) == (False, None, None)
    assert detect_endpointing(DEFAULT_RULES, 6.4, 0.92, 7.9)[1] == "rule1.2"
    assert detect_endpointing(DEFAULT_RULES, 40.32, 0.0, 10.0)[1] == "rule4"
    assert detect_endpointing(DEFAULT_RULES, 1.28, 1.28, 0.0) == (False, None, None)     # only silence so far
    y = {"a": dict(must_contain_nonsilence=True, min_trailing_silence=1, min_utterance_length=0.0, max_relative_cost=float("inf"))}
    assert load_endpointing_rule(y)["a"] == DEFAULT_RULES["rule1.1"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/streaming_decoder"), reason="reference tree not present")
def test_endpoint_rules_match_live_reference():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_online_endpoint", "/root/reference/streaming_decoder/online_endpoint.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from asr_streaming_b200.endpoint import DEFAULT_RULES, EndpointRules
    import yaml
    y = yaml.safe_load(open("/root/reference/streaming_decoder/config/asr-online.yaml"))["Endpointing_rules"]["DEFAULT"]
    ref_rules = ref.load_endpointing_rule(y)
    assert list(ref_rules) == list(DEFAULT_RULES)
    for k in ref_rules:
        for f in ("must_contain_nonsilence", "min_trailing_silence", "min_utterance_length", "max_relative_cost"):
            assert float(getattr(ref_rules[k], f)) == float(getattr(DEFAULT_RULES[k], f)), (k, f)
    tab = EndpointRules()
    rng = np.random.default_rng(5)
    for _ in range(3000):
        utt = round(int(rng.integers(0, 70)) * 0.64, 2)
        sil = round(min(utt, int(rng.integers(0, 40)) * 0.04), 2)
        cost = float(rng.choice([0.5, 2.0, 4.9, 5.0, 7.5, 8.0, 10.0]))
        d, name, _ = ref.detect_endpointing(ref_rules, utt, sil, cost)
        fired, which = tab.detect([utt], [sil], [cost])
        assert bool(fired[0]) == bool(d) and (tab.names[which[0]] if d else None) == name


class TokenEngine(FakeEngine):
    """Fake engine whose chunk decodes to a token iff the chunk's last sample is positive; blank frames = 16 - (sample % 17)."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.frames = {}

    def step(self, slots, pcm, want_logprobs=False):
        slots = [int(s) for s in slots]
        self.calls.append((slots, pcm.copy()))
        n, S = len(slots), self.cfg.seg_rows
        last = pcm[:, -1].astype(np.int64)
        nnew = (last > 0).astype(np.int32)
        newtok = np.zeros((n, S), np.int32)
        newtok[:, 0] = 2 + last % 5
        blank, has = np.zeros(n, np.int32), np.zeros(n, bool)
        for i, s in enumerate(slots):
            st = self.frames.setdefault(s, dict(n=0, last=-1))
            if s in self.resets:
                st["n"], st["last"] = 0, -1
                self.resets = [x for x in self.resets if x != s]
            if nnew[i]:
                st["last"] = st["n"] + S - 1 - int(last[i] % S)
            st["n"] += S
            has[i] = st["last"] >= 0
            blank[i] = st["n"] - 1 - st["last"] if has[i] else st["n"]
        new = [newtok[i, :nnew[i]].copy() for i in range(n)]
        return A.StepResult(np.zeros((n, S), np.int32), new, blank, has, None, n_new=nnew, new_tokens_padded=newtok)


def _stream_py_oracle(chunks_last, S=16):
    """Per-session scalar restatement of the reference loop (stream.py:110-163 + online_endpoint.py) fed with the TokenEngine's
    decode rule; returns [(chunk index, rule name)] of the endpoints."""
    from asr_streaming_b200.endpoint import DEFAULT_RULES
    out, chunk_processed, trailing, n, last_tok = [], 0, 0.0, 0, -1
    for k, last in enumerate(chunks_last):
        if last > 0:
            last_tok = n + S - 1 - int(last % S)
        n += S
        text = last_tok >= 0
        chunk_processed += 1
        if text:
            trailing = float(np.float32(n - 1 - last_tok) * np.float32(0.04))
        else:
            trailing += 0.64
        utt = chunk_processed * 10240 / 16000
        trailing = round(trailing, 2)
        name = next((nm for nm, r in DEFAULT_RULES.items() if _scalar_rule_activated(r, trailing, utt, 10.0)), None)
        if name:
            out.append((k, name))
            chunk_processed, trailing, n, last_tok = 0, 0.0, 0, -1
    return out


def test_scheduler_endpointing_matches_scalar_stream_loop():
    from asr_streaming_b200.endpoint import EndpointRules
    cfg = A.ModelConfig(max_batch=64, max_sessions=64)
    eng = TokenEngine(cfg)
    sch = A.SessionScheduler(eng, endpoint_rules=EndpointRules())
    rng = np.random.default_rng(11)
    n_sess, n_chunks = 24, 80
    ss = [sch.open() for _ in range(n_sess)]
    lasts = np.where(rng.random((n_sess, n_chunks)) < 0.35, rng.integers(1, 3000, (n_sess, n_chunks)), -rng.integers(0, 3000, (n_sess, n_chunks)))
    lasts[3] = -5                                        # a session that never says anything: never endpoints (must_contain_nonsilence)
    lasts[4] = 7                                         # talks all the time: rule4 at 40.32 s (63 chunks)
    got = {s.id: [] for s in ss}
    for k in range(n_chunks):
        for i, s in enumerate(ss):
            a = np.zeros(cfg.segment_length, np.int16)
            a[-1] = lasts[i, k]
            s.accept_waveform(a)
        res = sch.tick()
        assert len(res) == n_sess
        for j, s in enumerate(res.sessions):
            if res.final[j]:
                got[s.id].append((k, res.final_rule[j]))
                assert s.id in res.final_tokens and s.tokens == [] and s.chunk_processed == 0
    for i, s in enumerate(ss):
        assert got[s.id] == _stream_py_oracle(lasts[i]), i
    assert got[ss[3].id] == [] and got[ss[4].id] == [(62, "rule4")]
    assert sum(len(v) for v in got.values()) > 20


def test_scheduler_vectorised_gate_and_backlog_compaction():
    from asr_streaming_b200.scheduler import energy_gate
    cfg = A.ModelConfig(max_batch=8, max_sessions=8)
    eng = TokenEngine(cfg)
    sch = A.SessionScheduler(eng, backlog_chunks=2)
    a, b = sch.open(), sch.open()
    loud = np.full(cfg.segment_length, 1000, np.int16)
    quiet = np.full(cfg.segment_length, 10, np.int16)
    for k in range(12):                                   # ring compaction: many more samples than CAP flow through
        a.accept_waveform(loud + k)
        b.accept_waveform(quiet)
        res = sch.tick(gate=energy_gate())
        assert [s.id for s in res.sessions] == [a.id] and [s.id for s in res.skipped] == [b.id]
        assert np.array_equal(eng.calls[-1][1][0, cfg.buffer_length:], loud + k)
        if k:
            assert np.array_equal(eng.calls[-1][1][0, :cfg.buffer_length], (loud + k - 1)[-cfg.buffer_length:])
    assert b.chunk_processed == 12 and abs(b.trailing_blank_duration - 12 * 0.64) < 1e-9 and b.n_frames == 0
    with pytest.raises(BufferError):
        for _ in range(5):
            b.accept_waveform(quiet)


def test_pipelined_ticks_equal_synchronous_ticks():
    """Two ticks in flight with a backlog larger than max_batch: per session the decoded tokens, endpoints and counters must be
    exactly those of synchronous ticking, and no session may ride two in-flight steps."""
    from asr_streaming_b200.endpoint import EndpointRules
    cfg = A.ModelConfig(max_batch=8, max_sessions=32)
    rng = np.random.default_rng(4)
    n_sess, n_chunks = 20, 70
    lasts = np.where(rng.random((n_sess, n_chunks)) < 0.3, rng.integers(1, 3000, (n_sess, n_chunks)), -rng.integers(0, 3000, (n_sess, n_chunks)))
    audio = np.zeros((n_sess, n_chunks * cfg.segment_length), np.int16)
    audio[:, cfg.segment_length - 1::cfg.segment_length] = lasts

    def run(pipelined):
        eng = TokenEngine(cfg)
        sch = A.SessionScheduler(eng, endpoint_rules=EndpointRules(), backlog_chunks=n_chunks + 1)
        ss = [sch.open() for _ in range(n_sess)]
        for i, s in enumerate(ss):
            s.accept_waveform(audio[i])
        log = {s.id: [] for s in ss}

        def note(res):
            for j, s in enumerate(res.sessions):
                log[s.id].append((tuple(int(t) for t in res.new_tokens[j, :res.n_new[j]]), bool(res.final[j]), res.final_rule[j]))
        prev = None
        for _ in range(10000):
            if pipelined:
                p = sch.submit_tick()
                if prev is not None:
                    note(sch.collect_tick(prev))
                prev = p
                if not len(p.rows) and not sch.ready_rows().size:
                    break
            else:
                r = sch.tick()
                note(r)
                if not len(r):
                    break
        if prev is not None:
            note(sch.collect_tick(prev))
        return log, [(s.chunk_processed, s.segment, s.tokens) for s in ss]

    log_sync, end_sync = run(False)
    log_pipe, end_pipe = run(True)
    assert log_sync == log_pipe and end_sync == end_pipe
    assert all(len(v) == n_chunks for v in log_sync.values())
    assert sum(f for v in log_sync.values() for _, f, _ in v) > 10


def test_native_pcm_peaks_matches_numpy():
    import ctypes as C
    lib = _lib.load_library()
    rng = np.random.default_rng(9)
    audio = rng.integers(-32768, 32768, size=(300, 30000)).astype(np.int16)
    audio[7] = 0
    audio[8, 5000:] = -32768
    rows = rng.permutation(300)[:257].astype(np.int32)
    offs = rng.integers(0, 30000 - 13440, size=rows.size).astype(np.int64)
    out = np.empty(rows.size, np.int32)
    assert lib.asr_pcm_peaks(int(rows.size), audio.ctypes.data, audio.shape[1], rows.ctypes.data, offs.ctypes.data, 3200, 13440, out.ctypes.data) == 0
    exp = np.array([np.abs(audio[r, o + 3200:o + 13440].astype(np.int32)).max() for r, o in zip(rows, offs)])
    assert np.array_equal(out, exp)
    assert lib.asr_pcm_peaks(0, audio.ctypes.data, audio.shape[1], None, None, 0, 0, None) == 0


def test_result_messages_have_the_reference_wire_format():
    from asr_streaming_b200 import results as R
    vocab = ["-", "|", "xin", "chào", "<<", ">>"]
    m = R.interim_message("xin chào")
    assert m == ('{"id": "", "status": 0, "msg": 0, "segment": 0, "result": {"hypotheses": [{"transcript": "xin chào", '
                 '"transcript_normalized": "xin chào", "confidence": 0.0, "likelihood": 1.0, "word_alignment": []}], "final": false}, '
                 '"segment_start": 0.0, "segment_length": 0.0, "total_length": 0.0, "message_type": 0, "word_start": 0.0, "word_end": 0.0, '
                 '"snr": 0.0, "vol_noise": 0.0, "vol_speech": 0.0, "is_speaker": false}')
    assert R.interim_message("  ") is None
    f = R.final_message("7", 2, 6.4, 19.2, "xin chào")
    assert f.startswith('{"id": "7", "status": 0, "msg": 0, "segment": 2, "result": {"hypotheses": [{"transcript": "xin chào"') and '"final": true' in f
    assert '"segment_length": 6.4, "total_length": 19.2' in f and R.final_message("7", 2, 6.4, 19.2, "") is None
    if os.path.isdir("/root/reference/streaming_decoder"):       # field-for-field against the reference source (importing it has side effects)
        import ast, dataclasses
        tree = ast.parse(open("/root/reference/streaming_decoder/utils.py").read())
        cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == "DecodedResult")
        ref_fields = [n.target.id for n in cls.body if isinstance(n, ast.AnnAssign)]
        assert ref_fields == [f.name for f in dataclasses.fields(R.DecodedResult)]
        fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "create_hypotheses")
        keys = [n.targets[0].slice.value for n in fn.body if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Subscript)]
        assert keys == list(R.create_hypotheses("x"))
    # through the scheduler: interim after a decoded chunk, final when a rule fires
    from asr_streaming_b200.endpoint import EndpointRules
    cfg = A.ModelConfig(max_batch=4, max_sessions=4)
    sch = A.SessionScheduler(TokenEngine(cfg), endpoint_rules=EndpointRules())
    s = sch.open()
    seq = [5] + [-1] * 3                       # a token in the first chunk, then silence-like chunks until rule1.x fires
    msgs = []
    for last in seq:
        a = np.zeros(cfg.segment_length, np.int16); a[-1] = last
        s.accept_waveform(a)
        res = sch.tick()
        msgs += R.tick_messages(sch, res, vocab).get(s.id, [])
    assert len(msgs) >= 2 and '"final": false' in msgs[0] and '"final": true' in msgs[-1] and '"segment": 0' in msgs[-1]
    assert s.segment == 1


def test_partition_streams_covers_everything_once():
    from asr_streaming_b200.scheduler import partition_streams
    for n, w in ((32768, 8), (10, 3), (7, 8), (4096, 2)):
        parts = [partition_streams(n, w, r) for r in range(w)]
        flat = [i for p in parts for i in p]
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_scheduler_chunk_windows_property():
    """Property (hypothesis): for ANY split of an utterance into websocket messages, chunk k handed to the engine is the window
    [k*10240 - 3200, k*10240 + 10240) of the utterance with 3200 leading zeros (stream.py:23, :159), messages of
    <= 100 samples are dropped (stream.py:82), and the ring buffer compaction never corrupts a window."""
    from hypothesis import given, settings, strategies as st
    cfg = A.ModelConfig(max_batch=4, max_sessions=4)

    @settings(max_examples=40, deadline=None)
    @given(st.lists(st.integers(min_value=1, max_value=20000), min_size=1, max_size=30), st.integers(min_value=0, max_value=2**31 - 1))
    def run(sizes, seed):
        eng = FakeEngine(cfg)
        sch = A.SessionScheduler(eng, backlog_chunks=3)
        s = sch.open()
        rng = np.random.default_rng(seed)
        kept = [np.zeros(cfg.buffer_length, np.int16)]
        for n in sizes:
            msg = rng.integers(-3000, 3000, size=n).astype(np.int16)
            room = sch.CAP - s.length_of_segment
            if n > room:                                   # the server would apply back-pressure; drain first
                while sch.tick().rows.size:
                    pass
            s.accept_waveform(msg)
            if n > 100:
                kept.append(msg)
            if rng.random() < 0.5:
                sch.tick()
        while sch.tick().rows.size:
            pass
        utt = np.concatenate(kept)
        n_chunks = (utt.size - cfg.chunk_length) // cfg.segment_length + 1 if utt.size >= cfg.chunk_length else 0
        assert len(eng.calls) == n_chunks and s.chunk_processed == n_chunks
        for k, (_, pcm) in enumerate(eng.calls):
            assert np.array_equal(pcm[0], utt[k * cfg.segment_length:k * cfg.segment_length + cfg.chunk_length])

    run()


def test_weights_from_checkpoint_reads_the_reference_layout(tmp_path, oracle_weights):
    """A checkpoint file in the reference's layout ({'hyper_parameters', 'state_dict': {'encoder', 'decoder'}},
    recognition.py:149-159) packs to the same blob as the state dict itself; when the reference tree is present the file is also
    produced by the reference's own modules' state_dict()."""
    import torch
    enc = {k[len("encoder."):]: torch.from_numpy(v) for k, v in oracle_weights.items() if k.startswith("encoder.")}
    dec = {k[len("decoder."):]: torch.from_numpy(v) for k, v in oracle_weights.items() if k.startswith("decoder.")}
    path = tmp_path / "model.ckpt"
    torch.save({"hyper_parameters": {"encoder": {}, "decoder": {}}, "state_dict": {"encoder": enc, "decoder": dec}}, path)
    blob = A.weights_from_checkpoint(str(path), A.ModelConfig())
    assert np.array_equal(blob, A.pack_weights(oracle_weights))
    from oracle import ref_import
    if ref_import.available():
        ref_import.load_reference()
        from lightspeech.modules.encoder import StreamingAcousticEncoder      # encoder.py:73-147
        from lightspeech.modules.decoder import CTCDecoder                     # decoder.py:60-70
        e = StreamingAcousticEncoder(input_dim=128, d_model=512, segment_length=64, left_context_length=128, right_context_length=16,
                                     ffn_dim=2048, num_layers=20, subsampling_factor=4, num_heads=8, dropout=0.1, activation="gelu",
                                     max_memory_size=0, tanh_on_mem=True)
        d = CTCDecoder(512, 512, 804)
        e.load_state_dict(enc, strict=True)
        d.load_state_dict(dec, strict=True)
        torch.save({"hyper_parameters": {}, "state_dict": {"encoder": e.state_dict(), "decoder": d.state_dict()}}, path)
        assert np.array_equal(A.weights_from_checkpoint(str(path), A.ModelConfig()), blob)


# ------------------------------------------------------------------------------------------------ update_stream / endpoint goldens of the live reference
class ScriptedIdsEngine(FakeEngine):
    """Replays per-frame argmax ids through the bookkeeping ctc_greedy_kernel does on the device (layers.cu): unique_consecutive +
    blank drop with a carry across chunks, last frame with id > 1, and `has_text` = an id outside the silent set was seen."""

    def __init__(self, cfg, scripts):
        super().__init__(cfg)
        self.scripts, self.pos, self.carry = scripts, {}, {}

    def reset_sessions(self, slots):
        super().reset_sessions(slots)
        for s in slots:
            self.carry.pop(int(s), None)

    def step(self, slots, pcm, want_logprobs=False):
        slots = [int(s) for s in slots]
        self.calls.append((slots, pcm.copy()))
        n, S = len(slots), self.cfg.seg_rows
        nnew, newtok = np.zeros(n, np.int32), np.zeros((n, S), np.int32)
        blank, has, text = np.zeros(n, np.int32), np.zeros(n, bool), np.zeros(n, bool)
        for i, s in enumerate(slots):
            k = self.pos.get(s, 0)
            self.pos[s] = k + 1
            c = self.carry.setdefault(s, dict(prev=-1, nf=0, lt=-1, ht=False))
            for tid in self.scripts[s][k]:
                if tid != c["prev"] and tid != 0:
                    newtok[i, nnew[i]] = tid
                    nnew[i] += 1
                if tid > 1:
                    c["lt"] = c["nf"]
                c["ht"] |= tid not in self.silent
                c["prev"] = tid
                c["nf"] += 1
            has[i], text[i] = c["lt"] >= 0, c["ht"]
            blank[i] = c["nf"] - 1 - c["lt"] if has[i] else c["nf"]
        return A.StepResult(np.zeros((n, S), np.int32), None, blank, has, None, n_new=nnew, new_tokens_padded=newtok, has_text=text)


def test_update_stream_and_endpoints_match_live_reference_goldens():
    """tests/golden/stream_update.json (oracle/make_stream_goldens.py: the unmodified greedy_search + Stream.update_stream +
    Stream.endpoint_detected) vs the native scheduler: segments whose only tokens are '<<' / '>>' have an id > 1 but empty text, so
    trailing_blank_duration keeps growing by 0.64 and is_contain_token stays False (stream.py:121-125)."""
    import json
    import os
    from asr_streaming_b200.endpoint import EndpointRules
    from asr_streaming_b200.recognition import silent_ids
    with open(os.path.join(os.path.dirname(__file__), "golden", "stream_update.json")) as f:
        G = json.load(f)
    vocab = ["-", "|"] + [f"t{i}" for i in range(2, G["vocab_size"])]
    vocab[792], vocab[793] = "<<", ">>"
    assert silent_ids(vocab) == G["silent_ids"]
    cfg = A.ModelConfig(max_batch=8, max_sessions=8)
    names = list(G["cases"])
    eng = ScriptedIdsEngine(cfg, {i: [r["ids"] for r in G["cases"][nm]] for i, nm in enumerate(names)})
    sch = A.SessionScheduler(eng, endpoint_rules=EndpointRules(), vocab=vocab)
    ss = [sch.open() for _ in names]
    assert [s.slot for s in ss] == list(range(len(names)))
    for k in range(max(len(G["cases"][nm]) for nm in names)):
        live = [s for s, nm in zip(ss, names) if k < len(G["cases"][nm])]
        for s in live:
            s.accept_waveform(np.ones(cfg.segment_length, np.int16))
        res = sch.tick()
        assert len(res) == len(live)
        for j, s in enumerate(res.sessions):
            g = G["cases"][names[ss.index(s)]][k]
            assert bool(res.final[j]) == g["detected"], (names[ss.index(s)], k)
            assert abs(s.trailing_blank_duration - g["trailing_after"]) < 1e-9, (names[ss.index(s)], k, s.trailing_blank_duration, g)
            assert s.segment == g["segment"]
            if not g["detected"]:
                assert s.is_contain_token == g["contain"] and s.chunk_processed == g["chunk_processed"]
            else:
                assert abs(res.final_utt_length[s.id] - g["utt"]) < 1e-9


def test_token_overflow_is_reported_not_silent():
    """A segment longer than MAX_TOKENS: the tick says so (TickResult.overflow) instead of overwriting the last token."""
    from asr_streaming_b200.scheduler import MAX_TOKENS
    cfg = A.ModelConfig(max_batch=2, max_sessions=2)
    S = cfg.seg_rows
    n_chunks = MAX_TOKENS // S + 2
    script = [[2 + (k * S + i) % 700 for i in range(S)] for k in range(n_chunks)]           # 16 distinct tokens per chunk, never blank
    eng = ScriptedIdsEngine(cfg, {0: script})
    sch = A.SessionScheduler(eng)                                                              # no endpoint rules: the segment never ends
    s = sch.open()
    seen = []
    for k in range(n_chunks):
        s.accept_waveform(np.ones(cfg.segment_length, np.int16))
        res = sch.tick()
        seen.append(bool(res.overflow[0]))
    assert seen[:MAX_TOKENS // S] == [False] * (MAX_TOKENS // S) and all(seen[MAX_TOKENS // S:])
    assert len(s.tokens) == MAX_TOKENS and s.tokens == [t for ch in script for t in ch][:MAX_TOKENS]
    sch.reset(s)
    s.accept_waveform(np.ones(cfg.segment_length, np.int16))
    eng.pos[0] = 0
    assert not sch.tick().overflow[0]


def test_failed_step_releases_its_sessions_and_nothing_is_applied_before_submit():
    cfg = A.ModelConfig(max_batch=4, max_sessions=4)

    class Flaky(FakeEngine):
        fail_submit = fail_collect = False

        def submit(self, slots, pcm, want_logprobs=False):
            if self.fail_submit:
                raise RuntimeError("submit failed")
            return super().submit(slots, pcm, want_logprobs)

        def collect(self, t):
            out = super().collect(t)
            if self.fail_collect:
                raise RuntimeError("device fault")
            return out
    eng = Flaky(cfg)
    sch = A.SessionScheduler(eng)
    a, b = sch.open(), sch.open()
    for s in (a, b):
        s.accept_waveform(np.ones(cfg.segment_length * 2, np.int16))
    b.accept_waveform(np.zeros(1, np.int16))
    eng.fail_submit = True
    with pytest.raises(RuntimeError):
        sch.tick(gate=lambda sess, chunk: sess is a)                 # b would be VAD-skipped: must not be applied when the submit fails
    assert a.chunk_processed == 0 and b.chunk_processed == 0 and b.trailing_blank_duration == 0.0 and not sch.inflight.any()
    assert a.length_of_segment == cfg.buffer_length + 2 * cfg.segment_length
    eng.fail_submit, eng.fail_collect = False, True
    with pytest.raises(RuntimeError):
        sch.tick()
    assert not sch.inflight.any()                                    # the sessions of the lost step are eligible again
    eng.fail_collect = False
    assert len(sch.tick()) == 2
    sch.close(a)                                                     # close / reset no longer raise forever


def test_relative_cost_callback_reaches_the_rules():
    """rule1.2 (trailing silence >= 0.9 s, relative cost < 8) can only fire when a language model supplies the cost (utils.py:126-139)."""
    from asr_streaming_b200.endpoint import EndpointRules
    cfg = A.ModelConfig(max_batch=2, max_sessions=2)
    script = [[5] + [0] * 15, [0] * 16, [0] * 16]                    # token at frame 0: last_blank 0.6, 1.24 after the next chunk
    fired = {}
    for cost in (10.0, 3.0):
        eng = ScriptedIdsEngine(cfg, {0: [[5] + [0] * 9 + [0] * 6, [0] * 7 + [0] * 9, [0] * 16]})
        sch = A.SessionScheduler(eng, endpoint_rules=EndpointRules(), cost_fn=lambda sc, rows, c=cost: np.full(len(rows), c))
        s = sch.open()
        out = []
        for k in range(2):
            s.accept_waveform(np.ones(cfg.segment_length, np.int16))
            res = sch.tick()
            out.append(res.final_rule[0])
        fired[cost] = out
    # chunk 0: last_blank = 15 frames = 0.6 s: no rule.  chunk 1: 31 frames = 1.24 s >= 1.0: rule1.1 for any cost
    assert fired[10.0] == [None, "rule1.1"] and fired[3.0] == [None, "rule1.1"]
    for cost, want in ((10.0, None), (3.0, "rule1.3")):              # 0.84 s of trailing silence: only cost < 5 ends the utterance (rule1.3: >= 0.8 s)
        eng = ScriptedIdsEngine(cfg, {0: [[0] * 10 + [5] + [0] * 5, [0] * 16]})
        sch = A.SessionScheduler(eng, endpoint_rules=EndpointRules(), cost_fn=lambda sc, rows, c=cost: np.full(len(rows), c))
        s = sch.open()
        for k in range(2):
            s.accept_waveform(np.ones(cfg.segment_length, np.int16))
            res = sch.tick()
        assert res.final_rule[0] == want, (cost, res.final_rule, s.trailing_blank_duration)


def test_step_less_ticks_do_not_consume_a_pipeline_slot():
    """A tick whose ready sessions were all VAD-skipped has nothing to collect: it must not use up one of the two in-flight slots."""
    cfg = A.ModelConfig(max_batch=1, max_sessions=4)
    eng = FakeEngine(cfg)
    sch = A.SessionScheduler(eng)
    a, b, c = sch.open(), sch.open(), sch.open()
    loud, quiet = np.full(cfg.segment_length, 900, np.int16), np.zeros(cfg.segment_length, np.int16)
    a.accept_waveform(loud); b.accept_waveform(quiet); c.accept_waveform(loud)
    gate = lambda sess, chunk: bool(np.abs(chunk).max() > 0)
    p1 = sch.submit_tick(gate=gate)                                  # a runs (max_batch 1)
    p2 = sch.submit_tick(gate=gate)                                  # b: skipped only
    assert [s.id for s in p1.res.sessions] == [a.id] and p2.rows.size == 0 and [s.id for s in p2.res.skipped] == [b.id]
    p3 = sch.submit_tick(gate=gate)                                  # c runs: second slot, while p1 is still in flight
    assert [s.id for s in p3.res.sessions] == [c.id]
    assert len(sch.collect_tick(p1)) == 1 and len(sch.collect_tick(p2)) == 0 and len(sch.collect_tick(p3)) == 1
    assert b.chunk_processed == 1 and abs(b.trailing_blank_duration - 0.64) < 1e-9 and not sch.inflight.any()
