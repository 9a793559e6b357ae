"""Ragged session batching: packs every stream that has a full chunk buffered into ONE engine step per tick.

Host-side mirror of the reference's per-connection loop (streaming_decoder/streaming_server.py:367-546) and of the v1
cross-stream batcher ``StreamingE2E.process`` (streaming_decoder_v1/streaming_asr.py:41-119), with the buffer semantics of
``Stream`` (streaming_decoder/stream.py:23-26 initial zero buffer, :78-87 accept_waveform, :110-125 update_stream,
:127-163 endpoint_detected, :159-160 advance by segment_length, :166-189 VAD skip).  Streams progress independently: a tick
may mix first chunks (no left context), steady-state chunks and streams that were just reset by an endpoint; the device
applies per-stream left-context validity.

Scale: session state is struct-of-arrays numpy (one row per session), so a tick over thousands of sessions is a handful
of vectorised operations plus one native multi-threaded gather into the engine's pinned staging buffer
(``asr_gather_pcm``); endpoint rules (online_endpoint.py:42-94) are evaluated for all sessions at once and the endpoints
of a tick are one ``asr_session_reset_many`` launch.  Sessions never interact, so a multi-GPU box partitions them per GPU
(``GpuRouter``) with no collective.
"""
from __future__ import annotations

import os
import threading
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .config import ModelConfig
from .endpoint import EndpointRules
from .engine import FRAMERATE
from .recognition import ids_to_text

MAX_TOKENS = 512          # tokens kept per utterance segment (an utterance is force-ended at 40 s, asr-online.yaml:103-107)


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
        return ids_to_text(self.tokens)

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


@dataclass
class TickResult:
    """Outcome of one tick, struct-of-arrays (n = streams run through the model this tick)."""
    rows: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))           # [n] scheduler rows, in batch order
    _sched: Optional["SessionScheduler"] = None
    n_new: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))          # [n]
    new_tokens: np.ndarray = field(default_factory=lambda: np.zeros((0, 0), np.int32))  # [n, S], valid [:n_new]
    logprobs: Optional[np.ndarray] = None                            # [n, S, V] when requested
    step: object = None                                              # the engine's StepResult (argmax ids, beam hypotheses, ...)
    skipped: List[StreamSession] = field(default_factory=list)       # VAD-gated chunks (not run)
    final: np.ndarray = field(default_factory=lambda: np.zeros(0, bool))              # [n] endpoint fired after this chunk
    final_rule: List[Optional[str]] = field(default_factory=list)    # rule name per session (None if not final)
    final_tokens: Dict[int, List[int]] = field(default_factory=dict)  # session id -> tokens of the finished segment
    final_utt_length: Dict[int, float] = field(default_factory=dict)  # session id -> seconds decoded in the finished segment (stream.py:132-134)

    @property
    def beam_tokens(self) -> Optional[List[np.ndarray]]:
        """Best prefix-beam hypothesis per served stream (None without beam); built on first access."""
        return self.step.beam_tokens if self.step is not None else None

    def beam_row(self, j: int) -> Optional[np.ndarray]:
        """Best prefix-beam hypothesis of served stream j without materialising the others."""
        st = self.step
        if st is None:
            return None
        if getattr(st, "beam_tokens_padded", None) is not None:
            return st.beam_tokens_padded[j, :st.beam_len[j]]
        bt = st.beam_tokens
        return None if bt is None else bt[j]

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


@dataclass
class PendingTick:
    """A submitted, not yet collected tick."""
    res: TickResult
    rows: np.ndarray
    ticket: object
    out: object
    want_logprobs: bool


def energy_gate(threshold: int = 328):
    """Vectorised stand-in for the WebRTC / Silero VAD gate (both absent here): speech iff the new 640 ms of the
    chunk peaks above ``threshold`` int16 units (default 1 % of full scale).  Marks itself as vectorised for the scheduler."""
    def gate(chunks: np.ndarray, buffer_length: int) -> np.ndarray:
        return np.abs(chunks[:, buffer_length:]).max(axis=1) >= threshold
    gate.vectorised = True
    return gate


def native_energy_gate(threshold: int = 328):
    """Same decision as ``energy_gate`` computed by the library's multi-threaded host helper (asr_pcm_peaks) straight from the
    scheduler's audio rings — no per-session Python work, no chunk copies."""
    def gate(sched: "SessionScheduler", rows: np.ndarray) -> np.ndarray:
        cfg = sched.cfg
        return sched.engine.pcm_peaks(sched.audio, rows, sched.rd[rows], cfg.buffer_length, cfg.chunk_length) >= threshold
    gate.native = True
    return gate


class SessionScheduler:
    """One engine (one GPU).  ``tick()`` = one launch chain over all ready streams (up to max_batch)."""

    def __init__(self, engine, capacity: Optional[int] = None, backlog_chunks: int = 4,
                 endpoint_rules: Optional[EndpointRules] = None, relative_cost: float = 10.0, device_gather: Optional[bool] = None):
        self.engine, self.cfg = engine, engine.cfg
        cfg = self.cfg
        self.capacity = capacity or cfg.max_sessions
        self.CAP = cfg.chunk_length + backlog_chunks * cfg.segment_length
        n = self.capacity
        # audio rings in pinned, device-mapped memory when the engine offers it: the GPU then gathers each tick's chunks itself
        # Two ways to assemble a tick's batch (measured on B200, 4096 sessions, two ticks of <= 2048 in flight):
        #   host gather   (default) multi-threaded memcpy into the pinned staging buffer + one DMA: 2.5 ms of host time per tick, best
        #                 throughput (the copy engine is free): 122.6 k audio-s/s end to end, tick p99 24.6 ms
        #   device gather the GPU reads the chunks straight out of pinned rings over PCIe: 1.1 ms of host time per tick, tick p99
        #                 22.1 ms, but the gather kernel shares the SMs with the previous tick's kernels: 117.2 k audio-s/s.
        #                 The choice when the host is the bottleneck (many GPUs per host) or latency matters more than throughput.
        if device_gather is None:
            device_gather = os.environ.get("ASR_B200_DEVICE_GATHER") == "1"
        self._rings_pinned = bool(device_gather) and hasattr(engine, "host_alloc") and hasattr(engine, "submit_rings")
        self.audio = engine.host_alloc((n, self.CAP), np.int16) if self._rings_pinned else np.zeros((n, self.CAP), np.int16)
        self.rd = np.zeros(n, np.int64)
        self.wr = np.zeros(n, np.int64)
        self.active = np.zeros(n, bool)
        self.inflight = np.zeros(n, bool)                  # a chunk of this session is in a submitted, uncollected tick
        self.slot = np.full(n, -1, np.int32)
        self.tok = np.zeros((n, MAX_TOKENS), np.int32)
        self.ntok = np.zeros(n, np.int32)
        self.n_frames = np.zeros(n, np.int64)
        self.chunk_processed = np.zeros(n, np.int64)
        self.chunk_processed_total = np.zeros(n, np.int64)
        self.trailing = np.zeros(n, np.float64)
        self.contain_token = np.zeros(n, bool)
        self.segment = np.zeros(n, np.int64)
        self.last_served = np.zeros(n, np.int64)          # service sequence number of the last service (strict LRU under backlog)
        self._seq = 0
        self._tick = 0
        self._free = list(range(n - 1, -1, -1))
        self._next_id = 0
        self.sessions: Dict[int, StreamSession] = {}       # id -> handle
        self._by_row: Dict[int, StreamSession] = {}
        self.endpoint_rules = endpoint_rules
        self.relative_cost = relative_cost                 # LM relative cost fed to the rules when no ARPA LM is loaded (utils.py:126-139)
        self._chunk_s = cfg.segment_length / cfg.sample_rate   # 0.64 s (0.32 in low-latency mode)
        self._fallback_pack = None if hasattr(engine, "gather_pcm") else np.empty((cfg.max_batch, cfg.chunk_length), np.int16)

    # ------------------------------------------------------------------ session lifecycle
    def open(self) -> StreamSession:
        if not self._free:
            raise RuntimeError(f"scheduler is full ({self.capacity} sessions)")
        row = self._free.pop()
        s = StreamSession(self, row, self._next_id)
        self._next_id += 1
        self.sessions[s.id] = s
        self._by_row[row] = s
        self.slot[row] = self.engine.open_session()
        self.audio[row, :self.cfg.buffer_length] = 0            # stream.py:23 (buffer_length leading zeros)
        self.rd[row], self.wr[row] = 0, self.cfg.buffer_length
        self.active[row] = True
        self._clear_segment(np.array([row]))
        self.chunk_processed_total[row] = 0
        self.segment[row] = 0
        self.last_served[row] = self._seq
        self._seq += 1
        return s

    def close(self, s: StreamSession) -> None:
        if self.inflight[s.row]:
            raise RuntimeError("close: the session has a chunk in flight; collect its tick first")
        self.engine.close_session(int(self.slot[s.row]))
        self.active[s.row] = False
        self.slot[s.row] = -1
        self.sessions.pop(s.id, None)
        self._by_row.pop(s.row, None)
        self._free.append(s.row)

    def _clear_segment(self, rows: np.ndarray) -> None:
        self.ntok[rows] = 0
        self.n_frames[rows] = 0
        self.chunk_processed[rows] = 0
        self.contain_token[rows] = False
        self.trailing[rows] = 0.0

    def reset(self, s: StreamSession) -> None:
        """Endpoint: emission := [], state := init (streaming_server.py:514-515, :530; stream.py:152-157)."""
        if self.inflight[s.row]:
            raise RuntimeError("reset: the session has a chunk in flight; collect its tick first")
        self.engine.reset_session(int(self.slot[s.row]))
        self._clear_segment(np.array([s.row]))
        self.segment[s.row] += 1

    # ------------------------------------------------------------------ audio in
    def accept(self, s: StreamSession, pcm: np.ndarray) -> None:
        """stream.py:78-87 (messages of <= 100 samples are dropped).  int16, or float in [-1, 1) (converted)."""
        pcm = np.asarray(pcm).reshape(-1)
        if pcm.dtype != np.int16:
            pcm = np.clip(np.round(pcm.astype(np.float32) * 32768.0), -32768, 32767).astype(np.int16)
        n = pcm.size
        if n <= 100:
            return
        r = s.row
        if self.wr[r] + n > self.CAP:                                  # compact: move the unread tail to the front
            live = int(self.wr[r] - self.rd[r])
            if live + n > self.CAP:
                raise BufferError(f"session {s.id}: backlog of {live + n} samples exceeds the {self.CAP}-sample buffer")
            if self._rings_pinned and self.inflight[r]:
                self.engine.wait_inputs()                              # the GPU may still be reading this session's chunk out of the ring
            self.audio[r, :live] = self.audio[r, self.rd[r]:self.wr[r]]
            self.rd[r], self.wr[r] = 0, live
        self.audio[r, self.wr[r]:self.wr[r] + n] = pcm
        self.wr[r] += n

    def accept_block(self, rows: np.ndarray, block: np.ndarray) -> None:
        """Bulk ingest: ``block[i]`` (int16, equal lengths) is appended to session row ``rows[i]`` (no compaction: must fit)."""
        rows = np.asarray(rows, np.int64)
        block = np.asarray(block)
        assert block.dtype == np.int16 and block.ndim == 2 and block.shape[0] == rows.size
        n = block.shape[1]
        if (self.wr[rows] + n > self.CAP).any():
            raise BufferError(f"accept_block: {n} samples do not fit behind the write pointer of every session (CAP {self.CAP})")
        for w in np.unique(self.wr[rows]):
            m = self.wr[rows] == w
            self.audio[rows[m], w:w + n] = block[m]
        self.wr[rows] += n

    # ------------------------------------------------------------------ the tick
    def ready_rows(self, max_rows: Optional[int] = None) -> np.ndarray:
        cap = self.cfg.max_batch if max_rows is None else min(int(max_rows), self.cfg.max_batch)
        rows = np.nonzero(self.active & ~self.inflight & (self.wr - self.rd >= self.cfg.chunk_length))[0]
        if rows.size > cap:                                           # backlog: longest-waiting first, nobody starves
            order = np.argsort(self.last_served[rows], kind="stable")
            rows = rows[order[:cap]]
        return rows

    def ready_sessions(self) -> List[StreamSession]:
        return [self._by_row[int(r)] for r in self.ready_rows()]

    def _advance(self, rows: np.ndarray) -> None:
        self.rd[rows] += self.cfg.segment_length                      # stream.py:159-160
        self.last_served[rows] = self._seq + np.arange(rows.size)
        self._seq += int(rows.size)

    def skip(self, s: StreamSession) -> None:
        self._skip_rows(np.array([s.row]))

    def _skip_rows(self, rows: np.ndarray) -> None:
        """VAD said no speech (stream.py:183-189): the chunk is consumed without touching encoder state."""
        self.trailing[rows] += self._chunk_s
        self.chunk_processed[rows] += 1
        self.chunk_processed_total[rows] += 1
        self._advance(rows)

    def tick(self, want_logprobs: bool = False, gate: Optional[Callable] = None, max_rows: Optional[int] = None) -> TickResult:
        """One step over the ready streams.  ``gate``: ``energy_gate()`` (vectorised) or ``gate(session, chunk) -> bool``
        (streaming_server.py:374-379); it is consulted only for streams without a token in the current segment, and
        gated-out chunks are skipped.  Endpoint rules, when configured, are evaluated after the step for every served
        stream; fired endpoints reset encoder state and are reported in ``TickResult.final*``."""
        return self.collect_tick(self.submit_tick(want_logprobs, gate, max_rows))

    # Pipelined form: ``p1 = submit_tick(); p2 = submit_tick(); r1 = collect_tick(p1); ...`` keeps up to two ticks in flight, so
    # batch assembly + H2D of tick k+1 overlap the kernels of tick k.  A session with a chunk in flight is not eligible for
    # the next tick (its endpoint decision needs the results first), which preserves the reference's per-stream order
    # chunk -> update_stream -> endpoint_detected -> next chunk exactly.
    def submit_tick(self, want_logprobs: bool = False, gate: Optional[Callable] = None, max_rows: Optional[int] = None) -> "PendingTick":
        self._tick += 1
        cfg = self.cfg
        rows = self.ready_rows(max_rows)
        res = TickResult(_sched=self)
        pend = PendingTick(res, rows[:0], None, None, want_logprobs)
        if rows.size == 0:
            return pend
        # ---- VAD gate
        if gate is not None:
            need = ~self.contain_token[rows]
            keep = np.ones(rows.size, bool)
            if need.any():
                idx = rows[need]
                if getattr(gate, "native", False):
                    keep[need] = gate(self, idx)
                elif getattr(gate, "vectorised", False):
                    chunks = np.stack([self.audio[r, self.rd[r]:self.rd[r] + cfg.chunk_length] for r in idx])
                    keep[need] = gate(chunks, cfg.buffer_length)
                else:
                    for j in np.nonzero(need)[0]:
                        r = int(rows[j])
                        keep[j] = bool(gate(self._by_row[r], self.audio[r, self.rd[r]:self.rd[r] + cfg.chunk_length]))
            skipped = rows[~keep]
            if skipped.size:
                res.skipped = [self._by_row[int(r)] for r in skipped]
                self._skip_rows(skipped)
                self._endpoints(skipped, res, rows[:0])
            rows = rows[keep]
            if rows.size == 0:
                return pend
        n = int(rows.size)
        if self._rings_pinned:
            # ---- batch assembly on the GPU: a gather kernel reads the chunks straight out of the pinned rings
            pend.ticket = self.engine.submit_rings(self.slot[rows], self.audio, rows, self.rd[rows], want_logprobs)
        else:
            # ---- batch assembly straight into the pinned staging buffer of the next step
            if self._fallback_pack is None:
                pcm = self.engine.gather_pcm(self.audio, rows, self.rd[rows])
            else:
                pcm = self._fallback_pack[:n]
                for i, r in enumerate(rows):
                    pcm[i] = self.audio[r, self.rd[r]:self.rd[r] + cfg.chunk_length]
            if hasattr(self.engine, "submit"):
                pend.ticket = self.engine.submit(self.slot[rows], pcm, want_logprobs)
            else:
                pend.out = self.engine.step(self.slot[rows], pcm, want_logprobs)
        pend.rows = rows
        self.inflight[rows] = True
        self._advance(rows)
        return pend

    def collect_tick(self, pend: "PendingTick") -> TickResult:
        res, rows = pend.res, pend.rows
        n = int(rows.size)
        if n == 0:
            return res
        out = pend.out if pend.out is not None else self.engine.collect(pend.ticket)
        self.inflight[rows] = False
        # ---- vectorised bookkeeping (update_stream, stream.py:110-125)
        S = self.cfg.seg_rows
        if out.n_new is not None:
            n_new = np.asarray(out.n_new, np.int32)
            new_tok = np.where(np.arange(S)[None, :] < n_new[:, None], out.new_tokens_padded, 0).astype(np.int32)
        else:
            n_new = np.array([len(t) for t in out.new_tokens], np.int32)
            new_tok = np.zeros((n, S), np.int32)
            for i, t in enumerate(out.new_tokens):
                new_tok[i, :len(t)] = t
        for j in range(int(n_new.max())):
            m = n_new > j
            dst = np.minimum(self.ntok[rows[m]] + j, MAX_TOKENS - 1)
            self.tok[rows[m], dst] = new_tok[m, j]
        self.ntok[rows] = np.minimum(self.ntok[rows] + n_new, MAX_TOKENS)
        self.n_frames[rows] += S
        self.chunk_processed[rows] += 1
        self.chunk_processed_total[rows] += 1
        has = np.asarray(out.has_token, bool)              # a frame with id > 1 exists in the segment  <=>  non-empty text
        blank = np.asarray(out.blank_frames)
        lb = np.where(has, (blank.astype(np.float32) * np.float32(FRAMERATE)).astype(np.float64), FRAMERATE * blank)
        self.trailing[rows] = np.where(has, lb, self.trailing[rows] + self._chunk_s)
        self.contain_token[rows] |= has
        res.rows = rows
        res.n_new, res.new_tokens, res.logprobs, res.step = n_new, new_tok, out.logprobs, out
        res.final = np.zeros(n, bool)
        res.final_rule = [None] * n
        self._endpoints(rows, res, rows)
        return res

    def reset_rows(self, rows: np.ndarray) -> None:
        """Endpoint decided by the caller (e.g. a final-pass decoder or the client's EOS) for many sessions at once."""
        rows = np.asarray(rows, np.int64)
        if rows.size == 0:
            return
        if self.inflight[rows].any():
            raise RuntimeError("reset_rows: a session with a chunk in flight cannot be reset before its tick is collected")
        if hasattr(self.engine, "reset_sessions"):
            self.engine.reset_sessions(self.slot[rows])
        else:
            for r in rows:
                self.engine.reset_session(int(self.slot[r]))
        self._clear_segment(rows)
        self.segment[rows] += 1

    # ------------------------------------------------------------------ endpointing (stream.py:127-163, online_endpoint.py)
    def _endpoints(self, rows: np.ndarray, res: TickResult, run_rows: np.ndarray) -> None:
        if self.endpoint_rules is None or rows.size == 0:
            return
        utt = self.chunk_processed[rows] * self.cfg.segment_length / self.cfg.sample_rate
        trailing = np.round(self.trailing[rows], 2)
        self.trailing[rows] = trailing
        rc = np.full(rows.size, float(self.relative_cost))
        fired, which = self.endpoint_rules.detect(utt, trailing, rc)
        if not fired.any():
            return
        frows = rows[fired]
        for r, w in zip(frows, which[fired]):
            s = self._by_row[int(r)]
            res.final_tokens[s.id] = [int(t) for t in self.tok[r, :self.ntok[r]]]
            res.final_utt_length[s.id] = float(self.chunk_processed[r] * self.cfg.segment_length / self.cfg.sample_rate)
            pos = np.nonzero(run_rows == r)[0]
            if pos.size:
                res.final[pos[0]] = True
                res.final_rule[pos[0]] = self.endpoint_rules.names[w]
        self.reset_rows(frows)


class GpuRouter:
    """Partitions sessions across the GPUs of one box: one Engine + SessionScheduler per GPU, least-loaded placement at
    open(), independent ticks (one host thread per GPU; ctypes releases the GIL during the step).  No collectives:
    the path has no cross-session term."""

    def __init__(self, cfg: ModelConfig, weights: np.ndarray, devices: Sequence[int], **sched_kw):
        from .engine import Engine
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
