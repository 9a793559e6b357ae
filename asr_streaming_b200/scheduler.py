"""Ragged session batching: packs every stream that has a full chunk buffered into ONE engine step per tick.

Host-side mirror of the reference's per-connection loop (streaming_decoder/streaming_server.py:367-470) and of the v1
cross-stream batcher ``StreamingE2E.process`` (streaming_decoder_v1/streaming_asr.py:41-119), with the buffer semantics of
``Stream`` (streaming_decoder/stream.py:23-26 initial zero buffer, :78-87 accept_waveform, :159-160 advance by
segment_length).  Streams progress independently: a tick may mix first chunks (no left context), steady-state chunks,
and streams that were just reset by an endpoint; the device applies per-stream left-context validity.
Sessions never interact, so a multi-GPU box partitions them per GPU (GpuRouter) with no collective.
"""
from __future__ import annotations

import threading
from collections import deque
from typing import Deque, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .config import ModelConfig
from .engine import Engine, FRAMERATE, StepResult
from .recognition import ids_to_text


class StreamSession:
    """Per-websocket state the hot path needs (subset of ``Stream``, stream.py:10-64)."""

    def __init__(self, sid: int, slot: int, cfg: ModelConfig):
        self.id, self.slot, self.cfg = sid, slot, cfg
        self.audio = np.zeros(cfg.buffer_length, np.int16)          # stream.py:23 (buffer_length leading zeros)
        self.length_of_segment = cfg.buffer_length                  # stream.py:26
        self.tokens: List[int] = []
        self.n_frames = 0
        self.chunk_processed = 0
        self.chunk_processed_total = 0
        self.trailing_blank_duration = 0.0
        self.is_contain_token = False
        self.segment = 0
        self.gpu = 0

    def accept_waveform(self, pcm: np.ndarray) -> None:
        """stream.py:78-87 (messages of <= 100 samples are dropped).  int16, or float in [-1, 1) (converted)."""
        if pcm.dtype != np.int16:
            pcm = np.clip(np.round(pcm.astype(np.float32) * 32768.0), -32768, 32767).astype(np.int16)
        if pcm.size > 100:
            self.audio = np.concatenate([self.audio, pcm.reshape(-1)])
            self.length_of_segment += pcm.size

    def ready(self) -> bool:
        return self.length_of_segment >= self.cfg.chunk_length        # streaming_server.py:371

    def chunk(self) -> np.ndarray:
        return self.audio[:self.cfg.chunk_length]                     # streaming_server.py:384

    def advance(self) -> None:
        self.audio = self.audio[self.cfg.segment_length:]             # stream.py:159-160
        self.length_of_segment -= self.cfg.segment_length

    @property
    def text(self) -> str:
        return ids_to_text(self.tokens)


class SessionScheduler:
    """One engine (one GPU).  ``tick()`` = one launch chain over all ready streams (up to max_batch)."""

    def __init__(self, engine: Engine):
        self.engine, self.cfg = engine, engine.cfg
        self.sessions: Dict[int, StreamSession] = {}
        self._next_id = 0
        self._rr: Deque[int] = deque()            # round-robin order so a backlog cannot starve old sessions
        self._fallback_pack = None if hasattr(engine, "pinned_pcm") else np.empty((self.cfg.max_batch, self.cfg.chunk_length), np.int16)

    def open(self) -> StreamSession:
        s = StreamSession(self._next_id, self.engine.open_session(), self.cfg)
        self._next_id += 1
        self.sessions[s.id] = s
        self._rr.append(s.id)
        return s

    def close(self, s: StreamSession) -> None:
        self.engine.close_session(s.slot)
        self.sessions.pop(s.id, None)
        try:
            self._rr.remove(s.id)
        except ValueError:
            pass

    def reset(self, s: StreamSession) -> None:
        """Endpoint: emission := [], state := init (streaming_server.py:514-515, :530; stream.py:152-157)."""
        self.engine.reset_session(s.slot)
        s.tokens, s.n_frames = [], 0
        s.chunk_processed, s.is_contain_token, s.trailing_blank_duration = 0, False, 0.0
        s.segment += 1

    def skip(self, s: StreamSession) -> None:
        """VAD said no speech (stream.py:183-189): the chunk is consumed without touching encoder state."""
        s.trailing_blank_duration += 0.64 * self.cfg.segment_size / 64
        s.chunk_processed += 1
        s.chunk_processed_total += 1
        s.advance()

    def ready_sessions(self) -> List[StreamSession]:
        out = []
        for sid in list(self._rr):
            s = self.sessions[sid]
            if s.ready():
                out.append(s)
                if len(out) == self.cfg.max_batch:
                    break
        return out

    def tick(self, want_logprobs: bool = False, gate=None) -> List[Tuple[StreamSession, List[int], Optional[np.ndarray]]]:
        """Runs one step over the ready streams.  ``gate(session, chunk) -> bool`` (optional) is the VAD decision
        (streaming_server.py:374-379); gated-out streams are skipped.  Returns (session, new token ids, logprobs|None)."""
        batch = self.ready_sessions()
        # batch assembly happens directly in the engine's pinned staging buffer of the next step (no second host copy)
        self._pack = self.engine.pinned_pcm(np.int16) if self._fallback_pack is None else self._fallback_pack
        run: List[StreamSession] = []
        for s in batch:
            if gate is not None and not s.is_contain_token and not gate(s, s.chunk()):
                self.skip(s)
            else:
                self._pack[len(run)] = s.chunk()
                run.append(s)
        if not run:
            return []
        n = len(run)
        res: StepResult = self.engine.step([s.slot for s in run], self._pack[:n], want_logprobs)
        out = []
        for i, s in enumerate(run):
            new = [int(t) for t in res.new_tokens[i]]
            s.tokens.extend(new)
            s.n_frames += self.cfg.seg_rows
            # update_stream (stream.py:110-125)
            s.chunk_processed += 1
            s.chunk_processed_total += 1
            if s.tokens:
                s.trailing_blank_duration = res.last_blank(i)
                s.is_contain_token = True
            else:
                s.trailing_blank_duration += 0.64 * self.cfg.segment_size / 64
            s.advance()
            self._rr.remove(s.id)
            self._rr.append(s.id)
            out.append((s, new, res.logprobs[i] if res.logprobs is not None else None))
        return out


class GpuRouter:
    """Partitions sessions across the GPUs of one box: one Engine + SessionScheduler per GPU, least-loaded placement at
    open(), independent ticks (one host thread per GPU; ctypes releases the GIL during the step).  No collectives:
    the path has no cross-session term."""

    def __init__(self, cfg: ModelConfig, weights: np.ndarray, devices: Sequence[int]):
        self.schedulers = [SessionScheduler(Engine(cfg, weights, d)) for d in devices]

    def open(self) -> StreamSession:
        g = min(range(len(self.schedulers)), key=lambda i: len(self.schedulers[i].sessions))
        s = self.schedulers[g].open()
        s.gpu = g
        return s

    def close(self, s: StreamSession) -> None:
        self.schedulers[s.gpu].close(s)

    def reset(self, s: StreamSession) -> None:
        self.schedulers[s.gpu].reset(s)

    def tick(self, want_logprobs: bool = False):
        results: List[list] = [[] for _ in self.schedulers]

        def run(i):
            results[i] = self.schedulers[i].tick(want_logprobs)

        threads = [threading.Thread(target=run, args=(i,)) for i in range(len(self.schedulers))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        return [r for rs in results for r in rs]


def partition_streams(n_streams: int, world_size: int, rank: int) -> range:
    """Static partition used by the multi-process launch (one process per GPU): contiguous, sizes differ by <= 1."""
    base, rem = divmod(n_streams, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))
