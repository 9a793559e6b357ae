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
    """Stands in for Engine on the CPU: records every step's (slots, pcm) and emits one deterministic token per chunk."""

    def __init__(self, cfg):
        self.cfg, self.calls, self._next, self.open_slots, self.resets = cfg, [], 0, set(), []

    def open_session(self):
        s = self._next
        self._next += 1
        self.open_slots.add(s)
        return s

    def close_session(self, s):
        self.open_slots.remove(s)

    def reset_session(self, s):
        self.resets.append(s)

    def step(self, slots, pcm, want_logprobs=False):
        slots = list(slots)
        assert len(set(slots)) == len(slots), "a session may appear at most once per step"
        assert len(slots) <= self.cfg.max_batch and pcm.shape == (len(slots), self.cfg.chunk_length)
        self.calls.append((slots, pcm.copy()))
        n, S = len(slots), self.cfg.seg_rows
        new = [np.array([2 + (int(pcm[i, -1]) % 7)], np.int32) for i in range(n)]
        return A.StepResult(np.zeros((n, S), np.int32), new, np.full(n, 3, np.int32), np.ones(n, bool), None)


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
    n_calls = len(eng.calls)
    sch._rr.rotate(1)                                                 # put the new session first
    out = sch.tick(gate=lambda sess, chunk: bool(np.abs(chunk).max() > 0))
    assert s.id not in [x.id for x, _, _ in out] and s.chunk_processed == 1 and abs(s.trailing_blank_duration - 0.64) < 1e-9
    assert s.length_of_segment == cfg.buffer_length


def test_partition_streams_covers_everything_once():
    from asr_streaming_b200.scheduler import partition_streams
    for n, w in ((32768, 8), (10, 3), (7, 8), (4096, 2)):
        parts = [partition_streams(n, w, r) for r in range(w)]
        flat = [i for p in parts for i in p]
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
