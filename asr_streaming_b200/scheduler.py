"""Ragged session batching: packs every stream that has a full chunk buffered into ONE engine step per tick.

Python face of the native session scheduler (csrc/sched.cu), which mirrors the reference's per-connection loop
(streaming_decoder/streaming_server.py:367-546), the buffer semantics of ``Stream`` (streaming_decoder/stream.py:23-26 initial
zero buffer, :78-87 accept_waveform, :110-125 update_stream, :127-163 endpoint_detected, :159-160 advance by segment_length,
:166-189 VAD skip), the rule evaluation of online_endpoint.py:42-94 and the v1 cross-stream batcher ``StreamingE2E.process``
(streaming_decoder_v1/streaming_asr.py:41-119).  Streams progress independently: a tick may mix first chunks (no left context),
steady-state chunks and streams that were just reset by an endpoint; the device applies per-stream left-context validity.

Scale: session state is struct-of-arrays owned by the library; the attributes below (``rd``, ``wr``, ``tok``, ``trailing``, ...) are
numpy views of it.  With an ``Engine`` a tick is two native calls — ``asr_sched_submit`` (ready scan, energy gate, skip
bookkeeping, batch assembly into pinned memory, launch) and ``asr_sched_collect`` (wait, update_stream for every served session
straight out of the pinned result area, endpoint rules, one stream-ordered reset launch) — with up to two ticks in flight.
With any other engine object (the CPU tests' scripted engines: ``open_session / close_session / reset_sessions / submit /
collect``) the same native bookkeeping runs around that object's submit / collect.  Sessions never interact, so a multi-GPU box
partitions them per GPU (``GpuRouter`` or one process per GPU) with no collective.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from .config import ModelConfig
from .endpoint import EndpointRules
from .engine import BEAM_MAX_LEN, Engine, StepResult
from .recognition import get_vocab, ids_to_text, silent_ids

MAX_TOKENS = 1024         # greedy tokens kept per utterance segment: rule4 force-ends an utterance at 40 s = 1000 frames (asr-online.yaml:103-107)
                          # and CTC emits at most one token per frame; a segment that still overflows is reported (TickResult.overflow)


def _view(ptr, dtype, shape):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


class StreamSession:
    """Handle of one websocket session (row of the scheduler's tables); exposes the fields of ``Stream`` the hot path uses."""

    __slots__ = ("sched", "row", "id", "gpu")

    def __init__(self, sched: "SessionScheduler", row: int, sid: int):
        self.sched, self.row, self.id, self.gpu = sched, row, sid, 0

    # -- audio in (stream.py:78-87) -------------------------------------------------------------------
    def accept_waveform(self, pcm: np.ndarray) -> None:
        self.sched.accept(self, pcm)

    def ready(self) -> bool:
        s = self.sched
        return bool(s.wr[self.row] - s.rd[self.row] >= s.cfg.chunk_length)        # streaming_server.py:371

    # -- state the server reads ------------------------------------------------------------------------
    @property
    def slot(self) -> int:
        return int(self.sched.slot[self.row])

    @property
    def length_of_segment(self) -> int:
        return int(self.sched.wr[self.row] - self.sched.rd[self.row])

    @property
    def tokens(self) -> List[int]:
        s = self.sched
        return [int(t) for t in s.tok[self.row, :s.ntok[self.row]]]

    @property
    def text(self) -> str:
        return ids_to_text(self.tokens, self.sched.vocab)

    @property
    def n_frames(self) -> int:
        return int(self.sched.n_frames[self.row])

    @property
    def chunk_processed(self) -> int:
        return int(self.sched.chunk_processed[self.row])

    @property
    def trailing_blank_duration(self) -> float:
        return float(self.sched.trailing[self.row])

    @property
    def is_contain_token(self) -> bool:
        return bool(self.sched.contain_token[self.row])

    @property
    def segment(self) -> int:
        return int(self.sched.segment[self.row])


class TickResult:
    """Outcome of one tick, struct-of-arrays (n = streams run through the model this tick)."""

    def __init__(self, sched: "SessionScheduler"):
        self._sched = sched
        self.rows = np.zeros(0, np.int64)                 # [n] scheduler rows, in batch order
        self.n_new = np.zeros(0, np.int32)                # [n]
        self.new_tokens = np.zeros((0, 0), np.int32)      # [n, S], row i valid in [:n_new[i]]
        self.logprobs: Optional[np.ndarray] = None        # [n, S, V] when requested
        self.step: Optional[StepResult] = None            # the step's per-stream outputs (argmax ids, blank frames, beam hypotheses, ...)
        self.skipped_rows = np.zeros(0, np.int64)         # VAD-gated chunks (consumed, not run)
        self.final = np.zeros(0, bool)                    # [n] an endpoint fired after this chunk
        self.final_rule_index = np.zeros(0, np.int32)     # [n] index into the rule table, -1 if not final
        self.overflow = np.zeros(0, bool)                 # [n] the segment lost greedy tokens (MAX_TOKENS) or its beam hypotheses were truncated
        self.final_tokens: Dict[int, List[int]] = {}      # session id -> greedy tokens of the finished segment (served AND skipped sessions)
        self.final_utt_length: Dict[int, float] = {}      # session id -> seconds decoded in the finished segment (stream.py:132-134)
        self.final_beam: Dict[int, np.ndarray] = {}       # session id -> best prefix-beam hypothesis of the finished segment (beam enabled)
        self._beam_view = None                            # (tokens [n, L] int16, len [n]) in pinned memory + the scheduler generation it is valid for
        self._n_commit_finals = 0

    @property
    def skipped(self) -> List[StreamSession]:
        return [self._sched._by_row[int(r)] for r in self.skipped_rows]

    @property
    def final_rule(self) -> List[Optional[str]]:
        names = self._sched.endpoint_rules.names if self._sched.endpoint_rules is not None else []
        return [names[k] if k >= 0 else None for k in self.final_rule_index]

    @property
    def beam_tokens(self) -> Optional[List[np.ndarray]]:
        """Best prefix-beam hypothesis per served stream (None without beam); read out of the step's result buffer on access."""
        if self.step is not None and self.step.beam_tokens is not None:
            return self.step.beam_tokens
        if self._beam_view is None:
            return None
        return [self.beam_row(j) for j in range(len(self))]

    def beam_row(self, j: int) -> Optional[np.ndarray]:
        """Best prefix-beam hypothesis of served stream j.  For non-final streams this reads the step's pinned result buffer, which
        lives until two more ticks have been submitted."""
        sid = self._sched._by_row[int(self.rows[j])].id if j < len(self) else None
        if sid in self.final_beam:
            return self.final_beam[sid]
        if self.step is not None and self.step.beam_tokens_padded is not None:
            return self.step.beam_tokens_padded[j, :self.step.beam_len[j]].astype(np.int32)
        if self._beam_view is None:
            return None
        tok, ln, gen = self._beam_view
        if self._sched._generation - gen >= 2:
            raise RuntimeError("beam_row: the step's result buffer was reused (two ticks were submitted since); read hypotheses right after collect_tick")
        return tok[j, :ln[j]].astype(np.int32)

    @property
    def sessions(self) -> List[StreamSession]:
        """The n sessions run through the model this tick, in batch order."""
        return [self._sched._by_row[int(r)] for r in self.rows]

    def __iter__(self):
        """(session, new token ids, logprobs | None) — the shape of the first-generation API."""
        for i, s in enumerate(self.sessions):
            yield s, [int(t) for t in self.new_tokens[i, :self.n_new[i]]], (self.logprobs[i] if self.logprobs is not None else None)

    def __len__(self):
        return int(self.rows.size)


class PendingTick:
    """A submitted, not yet collected tick."""

    def __init__(self, res: TickResult, rows: np.ndarray, tick: int, ticket=None, want_logprobs: bool = False):
        self.res, self.rows, self.tick, self.ticket, self.want_logprobs = res, rows, tick, ticket, want_logprobs


def energy_gate(threshold: int = 328):
    """Vectorised stand-in for the WebRTC / Silero VAD gate (both absent here): speech iff the new 640 ms of the
    chunk peaks above ``threshold`` int16 units (default 1 % of full scale).  Evaluated in numpy on the gathered chunks."""
    def gate(chunks: np.ndarray, buffer_length: int) -> np.ndarray:
        return np.abs(chunks[:, buffer_length:]).max(axis=1) >= threshold
    gate.vectorised = True
    return gate


def native_energy_gate(threshold: int = 328):
    """Same decision as ``energy_gate`` taken inside the native tick (asr_pcm_peaks over the session rings, multi-threaded): no
    Python between the ready scan and the launch."""
    def gate(sched: "SessionScheduler", rows: np.ndarray) -> np.ndarray:
        cfg = sched.cfg
        a = sched.audio
        return np.array([np.abs(a[r, sched.rd[r] + cfg.buffer_length:sched.rd[r] + cfg.chunk_length]).max() >= threshold for r in rows], bool)
    gate.native = True
    gate.threshold = int(threshold)
    return gate


class SessionScheduler:
    """One engine (one GPU).  ``tick()`` = one launch chain over all ready streams (up to max_batch).

    relative_cost: the LM relative cost fed to the endpoint rules (utils.py:126-139; the ARPA LM is absent, so a constant);
    cost_fn(scheduler, rows) -> [len(rows)] floats, when given, is called for the served sessions of every tick after their
    transcripts were updated and before the rules run (a language model plugs in here: rules x.2-x.4 need cost < 8 / 5 / 2)."""

    def __init__(self, engine, capacity: Optional[int] = None, backlog_chunks: int = 4, endpoint_rules: Optional[EndpointRules] = None,
                 relative_cost: float = 10.0, device_gather: Optional[bool] = None, cost_fn: Optional[Callable] = None,
                 vocab: Optional[Sequence[str]] = None):
        self.engine, self.cfg = engine, engine.cfg
        cfg = self.cfg
        self.lib = _lib.load_library()
        self.capacity = capacity or cfg.max_sessions
        self.vocab = list(vocab) if vocab is not None else get_vocab(cfg.vocab)
        self._real = isinstance(engine, Engine)
        # Two ways to assemble a tick's batch (measured on B200, 4096 sessions, two ticks of <= 2048 in flight):
        #   host gather   (default) multi-threaded memcpy into the pinned staging buffer + one DMA: best single-GPU throughput (the copy
        #                 engine is free)
        #   device gather the GPU reads the chunks straight out of pinned rings over PCIe: less host time per tick, but the gather kernel
        #                 shares the SMs with the previous tick's kernels.  The choice when the host is the bottleneck (many GPUs per host).
        if device_gather is None:
            device_gather = os.environ.get("ASR_B200_DEVICE_GATHER") == "1"
        self._rings_pinned = bool(device_gather) and self._real
        c = _lib.AsrSchedConfigC(self.capacity, cfg.chunk_length, cfg.segment_length, cfg.buffer_length, cfg.seg_rows, cfg.sample_rate,
                                 cfg.max_batch, backlog_chunks, MAX_TOKENS, int(self._rings_pinned), float(relative_cost))
        h = C.c_void_p()
        _lib.check(self.lib, self.lib.asr_sched_create(C.byref(c), engine._h if self._real else None, C.byref(h)), "asr_sched_create")
        self._h = h
        a = _lib.AsrSchedArraysC()
        _lib.check(self.lib, self.lib.asr_sched_arrays(self._h, C.byref(a)), "asr_sched_arrays")
        n = self.capacity
        self.CAP = int(a.audio_row_samples)
        self.audio = _view(a.audio, np.int16, (n, self.CAP))
        self.rd, self.wr = _view(a.rd, np.int64, (n,)), _view(a.wr, np.int64, (n,))
        self.active, self.inflight = _view(a.active, np.bool_, (n,)), _view(a.inflight, np.bool_, (n,))
        self.slot = _view(a.slot, np.int32, (n,))
        self.tok, self.ntok = _view(a.tok, np.int32, (n, MAX_TOKENS)), _view(a.ntok, np.int32, (n,))
        self.n_frames = _view(a.n_frames, np.int64, (n,))
        self.chunk_processed = _view(a.chunk_processed, np.int64, (n,))
        self.chunk_processed_total = _view(a.chunk_processed_total, np.int64, (n,))
        self.trailing = _view(a.trailing, np.float64, (n,))
        self.contain_token = _view(a.contain_token, np.bool_, (n,))
        self.segment, self.last_served = _view(a.segment, np.int64, (n,)), _view(a.last_served, np.int64, (n,))
        self.relative_costs = _view(a.relative_cost, np.float64, (n,))
        self.seg_overflow = _view(a.overflow, np.bool_, (n,))
        self.relative_cost = relative_cost
        self.cost_fn = cost_fn
        self._free = list(range(n - 1, -1, -1))
        self._next_id = 0
        self._generation = 0                                # ticks submitted (lifetime of the pinned result views)
        self.sessions: Dict[int, StreamSession] = {}        # id -> handle
        self._by_row: Dict[int, StreamSession] = {}
        self.endpoint_rules = None
        self.set_endpoint_rules(endpoint_rules)
        engine.set_silent_ids(silent_ids(self.vocab))        # `if text:` of Stream.update_stream (stream.py:121) on rendered text

    def close_scheduler(self) -> None:
        if getattr(self, "_h", None):
            self.lib.asr_sched_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close_scheduler()
        except Exception:
            pass

    def set_endpoint_rules(self, rules: Optional[EndpointRules]) -> None:
        self.endpoint_rules = rules
        if rules is None or len(rules) == 0:
            _lib.check(self.lib, self.lib.asr_sched_set_rules(self._h, 0, None, None, None, None), "asr_sched_set_rules")
            return
        must = np.ascontiguousarray(rules.must, np.uint8)
        sil, utt, cost = (np.ascontiguousarray(x, np.float64) for x in (rules.min_sil, rules.min_utt, rules.max_cost))
        _lib.check(self.lib, self.lib.asr_sched_set_rules(self._h, len(rules), must.ctypes.data, sil.ctypes.data, utt.ctypes.data, cost.ctypes.data),
                   "asr_sched_set_rules")

    # ------------------------------------------------------------------ session lifecycle
    def open(self) -> StreamSession:
        if not self._free:
            raise RuntimeError(f"scheduler is full ({self.capacity} sessions)")
        row = self._free.pop()
        slot = -1 if self._real else int(self.engine.open_session())
        try:
            _lib.check(self.lib, self.lib.asr_sched_open(self._h, row, slot), "asr_sched_open")
        except Exception:
            self._free.append(row)
            raise
        s = StreamSession(self, row, self._next_id)
        self._next_id += 1
        self.sessions[s.id] = s
        self._by_row[row] = s
        return s

    def close(self, s: StreamSession) -> None:
        slot = int(self.slot[s.row])
        try:
            _lib.check(self.lib, self.lib.asr_sched_close(self._h, s.row), "asr_sched_close")
        except _lib.AsrLibraryError as e:
            raise RuntimeError(str(e)) from e
        if not self._real:
            self.engine.close_session(slot)
        self.sessions.pop(s.id, None)
        self._by_row.pop(s.row, None)
        self._free.append(s.row)

    def reset(self, s: StreamSession) -> None:
        """Endpoint decided by the caller: emission := [], state := init (streaming_server.py:514-515, :530; stream.py:152-157)."""
        self.reset_rows(np.array([s.row]))

    def reset_rows(self, rows: np.ndarray) -> None:
        """Endpoint decided by the caller (e.g. a final-pass decoder or the client's EOS) for many sessions at once."""
        rows = np.ascontiguousarray(rows, np.int32)
        if rows.size == 0:
            return
        if not self._real:
            if self.inflight[rows].any():
                raise RuntimeError("reset_rows: a session with a chunk in flight cannot be reset before its tick is collected")
            self.engine.reset_sessions(self.slot[rows])
        try:
            _lib.check(self.lib, self.lib.asr_sched_reset_rows(self._h, int(rows.size), rows.ctypes.data), "asr_sched_reset_rows")
        except _lib.AsrLibraryError as e:
            raise RuntimeError(str(e)) from e

    # ------------------------------------------------------------------ audio in
    def accept(self, s: StreamSession, pcm: np.ndarray) -> None:
        """stream.py:78-87 (messages of <= 100 samples are dropped).  int16, or float in [-1, 1) (converted)."""
        pcm = np.asarray(pcm).reshape(-1)
        if pcm.dtype != np.int16:
            pcm = np.clip(np.round(pcm.astype(np.float32) * 32768.0), -32768, 32767).astype(np.int16)
        pcm = np.ascontiguousarray(pcm)
        rc = self.lib.asr_sched_accept(self._h, s.row, pcm.ctypes.data, int(pcm.size))
        if rc == 1:
            raise BufferError(f"session {s.id}: " + (self.lib.asr_last_error() or b"").decode("utf-8", "replace"))
        _lib.check(self.lib, rc, "asr_sched_accept")

    def accept_block(self, rows: np.ndarray, block: np.ndarray) -> None:
        """Bulk ingest: ``block[i]`` (int16, equal lengths) is appended to session row ``rows[i]`` with accept()'s semantics, in one native call."""
        rows = np.ascontiguousarray(rows, np.int32)
        block = np.ascontiguousarray(block)
        assert block.dtype == np.int16 and block.ndim == 2 and block.shape[0] == rows.size
        rc = self.lib.asr_sched_accept_block(self._h, int(rows.size), rows.ctypes.data, block.ctypes.data, int(block.shape[1]))
        if rc == 1:
            raise BufferError((self.lib.asr_last_error() or b"").decode("utf-8", "replace"))
        _lib.check(self.lib, rc, "asr_sched_accept_block")

    # ------------------------------------------------------------------ the tick
    def ready_rows(self, max_rows: Optional[int] = None) -> np.ndarray:
        """Sessions with a full chunk buffered and no chunk in flight; under backlog the longest-waiting first (nobody starves)."""
        p, n = C.c_void_p(), C.c_int32()
        _lib.check(self.lib, self.lib.asr_sched_ready(self._h, int(max_rows or 0), C.byref(p), C.byref(n)), "asr_sched_ready")
        return _view(p.value, np.int32, (n.value,)).astype(np.int64)

    def ready_sessions(self) -> List[StreamSession]:
        return [self._by_row[int(r)] for r in self.ready_rows()]

    def skip(self, s: StreamSession) -> None:
        """VAD said no speech (stream.py:183-189): the session's next chunk is consumed without touching encoder state."""
        r = s.row
        self.trailing[r] += self.cfg.segment_length / self.cfg.sample_rate
        self.chunk_processed[r] += 1
        self.chunk_processed_total[r] += 1
        self.rd[r] += self.cfg.segment_length

    def tick(self, want_logprobs: bool = False, gate: Optional[Callable] = None, max_rows: Optional[int] = None) -> TickResult:
        """One step over the ready streams.  ``gate``: ``native_energy_gate()`` (inside the native tick), ``energy_gate()``
        (vectorised numpy) or ``gate(session, chunk) -> bool`` (streaming_server.py:374-379); it is consulted only for streams
        without text in the current segment, and gated-out chunks are skipped.  Endpoint rules, when configured, are evaluated
        after the step for every served stream; fired endpoints reset encoder state and are reported in ``TickResult.final*``."""
        return self.collect_tick(self.submit_tick(want_logprobs, gate, max_rows))

    def _gate_mask(self, gate, max_rows):
        """keep-mask of a Python gate for the current ready set (None: no rows)."""
        cfg = self.cfg
        rows = self.ready_rows(max_rows)
        if rows.size == 0:
            return None
        keep = np.ones(rows.size, np.uint8)
        need = ~self.contain_token[rows]
        if need.any():
            idx = rows[need]
            if getattr(gate, "vectorised", False):
                chunks = np.stack([self.audio[r, self.rd[r]:self.rd[r] + cfg.chunk_length] for r in idx])
                keep[need] = gate(chunks, cfg.buffer_length)
            elif getattr(gate, "native", False):
                keep[need] = gate(self, idx)
            else:
                keep[need] = [bool(gate(self._by_row[int(r)], self.audio[r, self.rd[r]:self.rd[r] + cfg.chunk_length])) for r in idx]
        return keep

    # Pipelined form: ``p1 = submit_tick(); p2 = submit_tick(); r1 = collect_tick(p1); ...`` keeps up to two ticks in flight, so
    # batch assembly + H2D of tick k+1 overlap the kernels of tick k.  A session with a chunk in flight is not eligible for
    # the next tick (its endpoint decision needs the results first), which preserves the reference's per-stream order
    # chunk -> update_stream -> endpoint_detected -> next chunk exactly.  Nothing of a tick is applied before its step was
    # enqueued successfully; a step that fails at collect releases its sessions (their chunk is lost) and raises.
    def prestage(self, gate: Optional[Callable] = None) -> int:
        """One-tick-per-pass pipelining (real engine, host gather): every buffered chunk — also of the sessions whose previous chunk is still
        in flight — is gathered and its H2D copy started now, overlapping the running tick; the next ``submit_tick`` (called after that tick
        was collected) decides which of them run.  Returns the number of staged chunks."""
        if not self._real:
            return 0
        n = C.c_int32()
        thr = gate.threshold if (gate is not None and getattr(gate, "native", False)) else -1
        _lib.check(self.lib, self.lib.asr_sched_prestage(self._h, thr, C.byref(n)), "asr_sched_prestage")
        return n.value

    def submit_tick(self, want_logprobs: bool = False, gate: Optional[Callable] = None, max_rows: Optional[int] = None) -> PendingTick:
        res = TickResult(self)
        thr, keep = -1, None
        if gate is not None:
            if self._real and getattr(gate, "native", False):
                thr = gate.threshold
            else:
                keep = self._gate_mask(gate, max_rows)
        keep_p = keep.ctypes.data if keep is not None else None
        r = _lib.AsrSchedResultC()
        if self._real:
            t = C.c_int32()
            _lib.check(self.lib, self.lib.asr_sched_submit(self._h, int(max_rows or 0), thr, keep_p, int(want_logprobs), C.byref(r), C.byref(t)), "asr_sched_submit")
            self._generation += 1
            self._fill_commit(res, r)
            return PendingTick(res, res.rows, t.value, None, want_logprobs)
        plan = _lib.AsrSchedPlanC()
        _lib.check(self.lib, self.lib.asr_sched_plan(self._h, int(max_rows or 0), thr, keep_p, C.byref(plan)), "asr_sched_plan")
        ticket = None
        if plan.n:
            rows = _view(plan.rows, np.int32, (plan.n,))
            offs = _view(plan.offsets, np.int64, (plan.n,))
            L = self.cfg.chunk_length
            pcm = np.stack([self.audio[r_, o:o + L] for r_, o in zip(rows, offs)])
            ticket = self.engine.submit(_view(plan.slots, np.int32, (plan.n,)).copy(), pcm, want_logprobs)    # raises: nothing was applied yet
        _lib.check(self.lib, self.lib.asr_sched_commit(self._h, plan.tick, C.byref(r)), "asr_sched_commit")
        self._generation += 1
        self._fill_commit(res, r)
        return PendingTick(res, res.rows, plan.tick, ticket, want_logprobs)

    def _fill_commit(self, res: TickResult, r) -> None:
        res.rows = _view(r.rows, np.int32, (r.n,)).astype(np.int64)
        res.skipped_rows = _view(r.skipped, np.int32, (r.n_skipped,)).astype(np.int64)
        res.final = np.zeros(r.n, bool)
        res.final_rule_index = np.full(r.n, -1, np.int32)
        res.overflow = np.zeros(r.n, bool)
        res._n_commit_finals = int(r.n_final)               # endpoints of VAD-skipped sessions (fired at submit)
        self._fill_finals(res, r)
        if not self._real and r.n_final:
            self.engine.reset_sessions(self.slot[_view(r.final_rows, np.int32, (r.n_final,))])

    def _fill_finals(self, res: TickResult, r) -> None:
        if not r.n_final:
            return
        rows = _view(r.final_rows, np.int32, (r.n_final,))
        ntok, off = _view(r.final_ntok, np.int32, (r.n_final,)), _view(r.final_tok_off, np.int32, (r.n_final,))
        utt = _view(r.final_utt, np.float64, (r.n_final,))
        tok = _view(r.final_tok, np.int32, (int(off[-1] + ntok[-1]),))
        for i in range(r.n_final):
            sid = self._by_row[int(rows[i])].id
            res.final_tokens[sid] = [int(t) for t in tok[off[i]:off[i] + ntok[i]]]
            res.final_utt_length[sid] = float(utt[i])

    def collect_tick(self, pend: PendingTick) -> TickResult:
        res = pend.res
        n = int(pend.rows.size)
        if n == 0:
            return res
        S, V = self.cfg.seg_rows, self.cfg.vocab
        r = _lib.AsrSchedResultC()
        if self._real:
            _lib.check(self.lib, self.lib.asr_sched_collect(self._h, pend.tick, int(self.cost_fn is None), C.byref(r)), "asr_sched_collect")
            if self.cost_fn is not None:
                self.relative_costs[pend.rows] = np.asarray(self.cost_fn(self, pend.rows), np.float64)
                _lib.check(self.lib, self.lib.asr_sched_endpoints(self._h, pend.tick, C.byref(r)), "asr_sched_endpoints")
            step = StepResult(_view(r.argmax_ids, np.int32, (n, S)).copy(), None, _view(r.blank_frames, np.int32, (n,)).copy(),
                              _view(r.has_token, np.int32, (n,)).astype(bool), None, n_new=None, new_tokens_padded=None,
                              has_text=_view(r.has_text, np.int32, (n,)).astype(bool), flags=_view(r.flags, np.int32, (n,)).copy())
            if r.beam_len:
                step.beam_len = _view(r.beam_len, np.int32, (n,)).copy()
                step.beam_score = _view(r.beam_score, np.float32, (n,)).copy()
                res._beam_view = (_view(r.beam_tokens, np.int16, (n, BEAM_MAX_LEN)), step.beam_len, self._generation)
            if pend.want_logprobs and r.logprobs:
                res.logprobs = _view(r.logprobs, np.float32, (n, S, V)).copy()
                step.logprobs = res.logprobs
        else:
            try:
                out = self.engine.collect(pend.ticket)
            except Exception:
                self.lib.asr_sched_abort(self._h, pend.tick)          # the sessions of the lost step become eligible again
                raise
            o = _lib.AsrStepOutC()
            keepalive = []

            def arr(x, dt):
                a = np.ascontiguousarray(x, dt)
                keepalive.append(a)
                return a.ctypes.data
            if out.n_new is not None:
                n_new, new_tok = np.asarray(out.n_new, np.int32), np.asarray(out.new_tokens_padded, np.int32)
            else:
                n_new = np.array([len(t) for t in out.new_tokens], np.int32)
                new_tok = np.zeros((n, S), np.int32)
                for i, t in enumerate(out.new_tokens):
                    new_tok[i, :len(t)] = t
            o.n_new, o.new_tokens = arr(n_new, np.int32), arr(new_tok, np.int32)
            o.blank_frames, o.has_token = arr(out.blank_frames, np.int32), arr(np.asarray(out.has_token).astype(np.int32), np.int32)
            o.has_text = arr(np.asarray(out.has_text).astype(np.int32), np.int32)
            if out.flags is not None:
                o.flags = arr(out.flags, np.int32)
            _lib.check(self.lib, self.lib.asr_sched_update(self._h, pend.tick, C.byref(o)), "asr_sched_update")
            if self.cost_fn is not None:
                self.relative_costs[pend.rows] = np.asarray(self.cost_fn(self, pend.rows), np.float64)
            _lib.check(self.lib, self.lib.asr_sched_endpoints(self._h, pend.tick, C.byref(r)), "asr_sched_endpoints")
            fin = _view(r.final_rows, np.int32, (r.n_final,))[res._n_commit_finals:]      # endpoints fired by this chunk
            if fin.size:
                self.engine.reset_sessions(self.slot[fin])
            step, res.logprobs = out, out.logprobs
        res.n_new = _view(r.n_new, np.int32, (n,)).copy()
        res.new_tokens = _view(r.new_tokens, np.int32, (n, S)).copy()
        res.final = _view(r.final_flags, np.uint8, (n,)).astype(bool)
        res.final_rule_index = _view(r.final_rule, np.int32, (n,)).copy()
        res.overflow = _view(r.overflow, np.uint8, (n,)).astype(bool)
        res.step = step
        step.n_new, step.new_tokens_padded = res.n_new, res.new_tokens
        self._fill_finals(res, r)
        if res._beam_view is not None and res.final.any():            # hypotheses of the finished segments: copied out of the pinned buffer now
            tok, ln, _ = res._beam_view
            for j in np.nonzero(res.final)[0]:
                res.final_beam[self._by_row[int(res.rows[j])].id] = tok[j, :ln[j]].astype(np.int32)
        return res


class GpuRouter:
    """Partitions sessions across the GPUs of one box inside ONE process: one Engine + SessionScheduler per GPU, least-loaded placement
    at open(), independent ticks on one host thread per GPU (the native tick releases the GIL for its whole duration).  No collectives:
    the path has no cross-session term.  The measured deployment form is one process per GPU (bench.py under torchrun)."""

    def __init__(self, cfg: ModelConfig, weights: np.ndarray, devices: Sequence[int], **sched_kw):
        self.schedulers = [SessionScheduler(Engine(cfg, weights, d), **sched_kw) for d in devices]
        for g, sc in enumerate(self.schedulers):
            sc._next_id = g << 32                      # session ids stay unique across the GPUs of the box

    def open(self) -> StreamSession:
        g = min(range(len(self.schedulers)), key=lambda i: len(self.schedulers[i].sessions))
        s = self.schedulers[g].open()
        s.gpu = g
        return s

    def close(self, s: StreamSession) -> None:
        self.schedulers[s.gpu].close(s)

    def reset(self, s: StreamSession) -> None:
        self.schedulers[s.gpu].reset(s)

    def tick(self, want_logprobs: bool = False, gate=None) -> List[TickResult]:
        results: List[Optional[TickResult]] = [None] * len(self.schedulers)

        def run(i):
            results[i] = self.schedulers[i].tick(want_logprobs, gate)

        threads = [threading.Thread(target=run, args=(i,)) for i in range(len(self.schedulers))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        return results


def partition_streams(n_streams: int, world_size: int, rank: int) -> range:
    """Static partition used by the multi-process launch (one process per GPU): contiguous, sizes differ by <= 1."""
    base, rem = divmod(n_streams, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))
