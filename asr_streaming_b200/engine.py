"""Slot-based batched front of the C ABI: one Engine per GPU, device-resident K/V rings and greedy carry per session."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .config import ModelConfig

FRAMERATE = 0.04            # recognition.py:30
PCM_I16, PCM_F32 = 0, 1
FBANK_MELSPEC128, FBANK_KALDI80 = 0, 1
BEAM_MAX_LEN = 1024          # ASR_BEAM_MAX_LEN
FLAG_BEAM_TRUNCATED = 1      # ASR_FLAG_BEAM_TRUNCATED


class StepResult:
    """Per-step outputs for n streams (S = segment rows).  ``new_tokens`` / ``beam_tokens`` (lists of per-stream arrays) are built on
    first access from the padded arrays the device returned: a tick over thousands of sessions never pays for them unless asked."""

    def __init__(self, argmax_ids, new_tokens, blank_frames, has_token, logprobs, beam_tokens=None, beam_score=None, n_new=None,
                 new_tokens_padded=None, beam_tokens_padded=None, beam_len=None, has_text=None, flags=None):
        self.argmax_ids = argmax_ids                 # [n, S] int32
        self._new_tokens = new_tokens                # n arrays of the ids appended this chunk (collapsed, blank-free) or None (lazy)
        self.blank_frames = blank_frames             # [n] int32
        self.has_token = has_token                   # [n] bool: an id > 1 exists in the segment (recognition.py:40-41)
        self.has_text = has_token if has_text is None else has_text   # [n] bool: the rendered text of the segment is non-empty (stream.py:121)
        self.flags = flags                           # [n] int32 FLAG_* bits (None from engines that have none)
        self.logprobs = logprobs                     # [n, S, V] float32 or None
        self._beam_tokens = beam_tokens              # n arrays: best prefix-beam hypothesis of the utterance so far, or None (lazy / no beam)
        self.beam_score = beam_score                 # [n] log-probability of that hypothesis
        self.n_new = n_new                           # [n] int32: len(new_tokens[i])
        self.new_tokens_padded = new_tokens_padded   # [n, S] int32, row i valid in [:n_new[i]] (struct-of-arrays form)
        self.beam_tokens_padded = beam_tokens_padded  # [n, BEAM_MAX_LEN] int16, row i valid in [:beam_len[i]] (the rest is not written)
        self.beam_len = beam_len

    @property
    def new_tokens(self) -> List[np.ndarray]:
        if self._new_tokens is None:
            self._new_tokens = [self.new_tokens_padded[i, :self.n_new[i]].copy() for i in range(len(self.n_new))]
        return self._new_tokens

    @property
    def beam_tokens(self) -> Optional[List[np.ndarray]]:
        if self._beam_tokens is None and self.beam_tokens_padded is not None:
            self._beam_tokens = [self.beam_tokens_padded[i, :self.beam_len[i]].astype(np.int32) for i in range(len(self.beam_len))]
        return self._beam_tokens

    def last_blank(self, i: int) -> float:
        """The reference's ``last_blank`` (recognition.py:38-43): python float 0.04*T when the segment has no
        token yet, else the float32 product tensor(int64) * 0.04 -> .item()."""
        if self.has_token[i]:
            return float(np.float32(self.blank_frames[i]) * np.float32(FRAMERATE))
        return FRAMERATE * int(self.blank_frames[i])


def _cfg_to_c(cfg: ModelConfig) -> _lib.AsrConfigC:
    c = _lib.AsrConfigC()
    c.abi_version = _lib.ABI_VERSION
    for name in ("sample_rate", "hop", "n_fft", "win", "n_mels", "segment_size", "context_size", "bias", "stride", "d_model",
                 "n_heads", "ffn_dim", "n_layers", "left_context", "ctc_hidden", "vocab", "precision", "max_sessions", "max_batch"):
        setattr(c, name, int(getattr(cfg, name)))
    return c


class Engine:
    """Owns the device state of one GPU.  ``weights`` is the flat fp32 blob from ``pack_weights``."""

    def __init__(self, cfg: ModelConfig, weights: np.ndarray, device: int = 0):
        self.lib = _lib.load_library()
        self.cfg = cfg
        self.device = device
        self._c = _cfg_to_c(cfg)
        n = C.c_uint64()
        _lib.check(self.lib, self.lib.asr_weights_count(C.byref(self._c), C.byref(n)), "asr_weights_count")
        w = np.ascontiguousarray(weights, dtype=np.float32).reshape(-1)
        if w.size != n.value:
            raise ValueError(f"weights blob has {w.size} floats, config needs {n.value}")
        h = C.c_void_p()
        _lib.check(self.lib, self.lib.asr_engine_create(C.byref(self._c), w.ctypes.data, w.size, device, C.byref(h)), "asr_engine_create")
        self._h = h
        self.S = cfg.seg_rows
        self.beam = 0
        self._host_blocks = []

    def set_beam(self, beam: int, cand_k: int = 8) -> None:
        """Enable CTC prefix beam search (beam <= 16, cand_k <= 8) as part of every step; 0 disables."""
        _lib.check(self.lib, self.lib.asr_set_beam(self._h, beam, cand_k), "asr_set_beam")
        self.beam = beam

    # ------------------------------------------------------------------ lifecycle
    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.asr_engine_destroy(self._h)
            self._h = None
            for p in getattr(self, "_host_blocks", []):
                self.lib.asr_host_free(p)
            self._host_blocks = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ sessions
    def open_session(self) -> int:
        s = C.c_int32()
        _lib.check(self.lib, self.lib.asr_session_open(self._h, C.byref(s)), "asr_session_open")
        return s.value

    def set_silent_ids(self, ids: Sequence[int]) -> None:
        """Vocabulary ids that render to "" in greedy_search (see ``recognition.silent_ids``): they never make ``has_text`` true."""
        a = np.ascontiguousarray(ids, dtype=np.int32)
        _lib.check(self.lib, self.lib.asr_set_silent_ids(self._h, int(a.size), a.ctypes.data), "asr_set_silent_ids")

    def reset_session(self, slot: int) -> None:
        _lib.check(self.lib, self.lib.asr_session_reset(self._h, slot), "asr_session_reset")

    def reset_sessions(self, slots: Sequence[int]) -> None:
        """Endpoint on many sessions in one launch."""
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        if sl.size:
            _lib.check(self.lib, self.lib.asr_session_reset_many(self._h, int(sl.size), sl.ctypes.data), "asr_session_reset_many")

    def gather_pcm(self, audio: np.ndarray, rows: np.ndarray, offsets: np.ndarray) -> np.ndarray:
        """Native batch assembly: chunk i = audio[rows[i], offsets[i] : offsets[i] + chunk_length] -> row i of the pinned staging
        buffer of the next step.  Returns the [n, chunk_length] view to pass to step()/submit()."""
        assert audio.dtype == np.int16 and audio.flags.c_contiguous and audio.ndim == 2
        rows = np.ascontiguousarray(rows, np.int32)
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = int(rows.size)
        out = C.c_void_p()
        _lib.check(self.lib, self.lib.asr_gather_pcm(self._h, n, audio.ctypes.data, audio.shape[1], rows.ctypes.data, offsets.ctypes.data,
                                                     C.byref(out)), "asr_gather_pcm")
        L = self.cfg.chunk_length
        buf = (C.c_char * (n * L * 2)).from_address(out.value)
        return np.frombuffer(buf, dtype=np.int16, count=n * L).reshape(n, L)

    def pcm_peaks(self, audio: np.ndarray, rows: np.ndarray, offsets: np.ndarray, start: int, stop: int) -> np.ndarray:
        """max |x| over audio[rows[i], offsets[i] + start : offsets[i] + stop] for every i (host helper, multi-threaded)."""
        assert audio.dtype == np.int16 and audio.flags.c_contiguous and audio.ndim == 2
        rows = np.ascontiguousarray(rows, np.int32)
        offsets = np.ascontiguousarray(offsets, np.int64)
        out = np.empty(rows.size, np.int32)
        _lib.check(self.lib, self.lib.asr_pcm_peaks(int(rows.size), audio.ctypes.data, audio.shape[1], rows.ctypes.data, offsets.ctypes.data,
                                                    int(start), int(stop), out.ctypes.data), "asr_pcm_peaks")
        return out

    def close_session(self, slot: int) -> None:
        _lib.check(self.lib, self.lib.asr_session_close(self._h, slot), "asr_session_close")

    # ------------------------------------------------------------------ the hot path
    def _pcm(self, pcm: np.ndarray, n: int):
        a = np.ascontiguousarray(pcm)
        if a.dtype == np.int16:
            fmt = PCM_I16
        elif a.dtype == np.float32:
            fmt = PCM_F32
        else:
            raise TypeError(f"pcm must be int16 or float32, got {a.dtype}")
        if a.size != n * self.cfg.chunk_length:
            raise ValueError(f"pcm has {a.size} samples, expected {n} x chunk_length {self.cfg.chunk_length}")
        return a, fmt

    def _alloc_out(self, n: int, want_logprobs: bool):
        S, V = self.S, self.cfg.vocab
        bufs = dict(argmax=np.empty((n, S), np.int32), newtok=np.empty((n, S), np.int32), nnew=np.empty(n, np.int32),
                    blank=np.empty(n, np.int32), hastok=np.empty(n, np.int32), hastext=np.empty(n, np.int32), flags=np.empty(n, np.int32),
                    logprobs=np.empty((n, S, V), np.float32) if want_logprobs else None)
        o = _lib.AsrStepOutC(bufs["argmax"].ctypes.data, bufs["newtok"].ctypes.data, bufs["nnew"].ctypes.data, bufs["blank"].ctypes.data,
                             bufs["hastok"].ctypes.data, bufs["logprobs"].ctypes.data if want_logprobs else None)
        o.has_text, o.flags = bufs["hastext"].ctypes.data, bufs["flags"].ctypes.data
        if self.beam:
            bufs["btok"] = np.empty((n, BEAM_MAX_LEN), np.int16)
            bufs["blen"] = np.empty(n, np.int32)
            bufs["bscore"] = np.empty(n, np.float32)
            o.beam_tokens, o.beam_len, o.beam_score = bufs["btok"].ctypes.data, bufs["blen"].ctypes.data, bufs["bscore"].ctypes.data
        return bufs, o

    @staticmethod
    def _result(bufs, n) -> StepResult:
        r = StepResult(bufs["argmax"], None, bufs["blank"], bufs["hastok"].astype(bool), bufs["logprobs"],
                       n_new=bufs["nnew"], new_tokens_padded=bufs["newtok"], has_text=bufs["hastext"].astype(bool), flags=bufs["flags"])
        if "btok" in bufs:
            r.beam_tokens_padded, r.beam_len, r.beam_score = bufs["btok"], bufs["blen"], bufs["bscore"]
        return r

    def step(self, slots: Sequence[int], pcm: np.ndarray, want_logprobs: bool = False) -> StepResult:
        """One chunk for each of ``slots`` (any mix of progress).  pcm: [n, chunk_length] int16 or float32."""
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        n = int(sl.size)
        a, fmt = self._pcm(pcm, n)
        bufs, o = self._alloc_out(n, want_logprobs)
        _lib.check(self.lib, self.lib.asr_step(self._h, n, sl.ctypes.data, a.ctypes.data, fmt, C.byref(o)), "asr_step")
        return self._result(bufs, n)

    # pipelined form: up to two steps in flight, H2D of step k+1 overlaps the kernels of step k
    def submit(self, slots: Sequence[int], pcm: np.ndarray, want_logprobs: bool = False):
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        n = int(sl.size)
        a, fmt = self._pcm(pcm, n)
        t = C.c_int32()
        _lib.check(self.lib, self.lib.asr_submit(self._h, n, sl.ctypes.data, a.ctypes.data, fmt, int(want_logprobs), C.byref(t)), "asr_submit")
        return (t.value, n, want_logprobs)

    def submit_rings(self, slots: Sequence[int], audio: np.ndarray, rows: np.ndarray, offsets: np.ndarray, want_logprobs: bool = False):
        """submit() with the batch assembled by the GPU: chunk i = audio[rows[i], offsets[i] : offsets[i] + chunk_length], read by a
        gather kernel straight out of ``audio`` (which must come from ``host_alloc``) — no host-side copy."""
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        rows = np.ascontiguousarray(rows, np.int32)
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = int(sl.size)
        assert audio.dtype == np.int16 and audio.ndim == 2 and audio.flags.c_contiguous and rows.size == n and offsets.size == n
        t = C.c_int32()
        _lib.check(self.lib, self.lib.asr_submit_rings(self._h, n, sl.ctypes.data, audio.ctypes.data, audio.shape[1], rows.ctypes.data,
                                                       offsets.ctypes.data, int(want_logprobs), C.byref(t)), "asr_submit_rings")
        return (t.value, n, want_logprobs)

    def wait_inputs(self) -> None:
        """Returns once every submitted step has read its inputs (the host memory they came from may be rewritten)."""
        _lib.check(self.lib, self.lib.asr_wait_inputs(self._h), "asr_wait_inputs")

    def host_alloc(self, shape, dtype=np.int16) -> np.ndarray:
        """Zero-filled numpy array in pinned, device-mapped host memory (for the scheduler's audio rings); freed with the engine."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = self.lib.asr_host_alloc(nbytes)
        if not ptr:
            raise _lib.AsrLibraryError("asr_host_alloc failed: " + (self.lib.asr_last_error() or b"").decode("utf-8", "replace"))
        self._host_blocks.append(ptr)
        buf = (C.c_char * max(nbytes, 1)).from_address(ptr)
        a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        a[...] = 0
        return a

    def collect(self, ticket) -> StepResult:
        t, n, want = ticket
        bufs, o = self._alloc_out(n, want)
        _lib.check(self.lib, self.lib.asr_collect(self._h, t, C.byref(o)), "asr_collect")
        return self._result(bufs, n)

    # split form (device-resident timing)
    def stage(self, slots: Sequence[int], pcm: np.ndarray) -> int:
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        a, fmt = self._pcm(pcm, int(sl.size))
        _lib.check(self.lib, self.lib.asr_stage(self._h, int(sl.size), sl.ctypes.data, a.ctypes.data, fmt), "asr_stage")
        return int(sl.size)

    def run_staged(self, n: int, want_logprobs: bool = False) -> None:
        _lib.check(self.lib, self.lib.asr_run_staged(self._h, n, int(want_logprobs)), "asr_run_staged")

    def fetch(self, n: int, want_logprobs: bool = False) -> StepResult:
        bufs, o = self._alloc_out(n, want_logprobs)
        _lib.check(self.lib, self.lib.asr_fetch(self._h, n, C.byref(o)), "asr_fetch")
        return self._result(bufs, n)

    def sync(self) -> None:
        _lib.check(self.lib, self.lib.asr_sync(self._h), "asr_sync")

    def pinned_pcm(self, dtype=np.int16) -> np.ndarray:
        """numpy view [max_batch, chunk_length] of the engine's pinned staging buffer: fill rows [0, n) in place and pass
        ``view[:n]`` to step()/stage() — the host copy is skipped and the H2D DMA reads what the caller wrote."""
        cap = C.c_uint64()
        ptr = self.lib.asr_pinned_pcm(self._h, C.byref(cap))
        if not ptr:
            raise _lib.AsrLibraryError("asr_pinned_pcm failed")
        n = self.cfg.max_batch * self.cfg.chunk_length
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(self.cfg.max_batch, self.cfg.chunk_length)

    @property
    def cuda_stream(self) -> int:
        return int(self.lib.asr_stream_handle(self._h) or 0)

    # ------------------------------------------------------------------ fbank
    def fbank(self, pcm: np.ndarray, kind: int = FBANK_MELSPEC128, subtract_mean: bool = False) -> np.ndarray:
        """pcm [n, n_samples] int16/float32 -> [n, frames, mels] float32 (extract_filterbank, audio.py:9-30 / Kaldi fbank)."""
        a = np.ascontiguousarray(pcm)
        if a.ndim != 2:
            raise ValueError("pcm must be [n, n_samples]")
        fmt = PCM_I16 if a.dtype == np.int16 else PCM_F32
        if a.dtype not in (np.int16, np.float32):
            raise TypeError("pcm must be int16 or float32")
        n, ns = a.shape
        if kind == FBANK_MELSPEC128:
            frames, mels = self.cfg.frames, self.cfg.n_mels
        else:
            frames, mels = 1 + (ns - 400) // 160, 80
        out = np.empty((n, frames, mels), np.float32)
        _lib.check(self.lib, self.lib.asr_fbank(self._h, kind, n, a.ctypes.data, fmt, ns, int(subtract_mean), out.ctypes.data), "asr_fbank")
        return out

    def stage_raw(self, pcm: np.ndarray) -> None:
        a = np.ascontiguousarray(pcm)
        _lib.check(self.lib, self.lib.asr_stage_raw(self._h, a.ctypes.data, a.nbytes), "asr_stage_raw")

    def fbank_staged(self, kind: int, n: int, fmt: int, n_samples: int) -> None:
        _lib.check(self.lib, self.lib.asr_fbank_staged(self._h, kind, n, fmt, n_samples), "asr_fbank_staged")

    # ------------------------------------------------------------------ stats / diagnostics
    def stats(self) -> dict:
        s = _lib.AsrStatsC()
        _lib.check(self.lib, self.lib.asr_get_stats(self._h, C.byref(s)), "asr_get_stats")
        return {k: getattr(s, k) for k, _ in _lib.AsrStatsC._fields_}

    def pipeline_gpu_time(self, reset: bool = True):
        """(total device ms, steps) of the pipelined steps collected so far (kernel chain + result D2H on the engine stream)."""
        ms, n = C.c_double(), C.c_uint64()
        _lib.check(self.lib, self.lib.asr_pipeline_gpu_time(self._h, C.byref(ms), C.byref(n), int(reset)), "asr_pipeline_gpu_time")
        return ms.value, int(n.value)

    PROF_NAMES = ("fbank", "gemm_input_linear", "layernorm", "gemm_qkv", "attention", "gemm_out_proj", "gemm_ffn1", "gemm_ffn2",
                  "gemm_ctc1", "gemm_ctc2", "ctc_greedy", "beam")

    def profile_enable(self, on: bool) -> None:
        _lib.check(self.lib, self.lib.asr_profile_enable(self._h, int(on)), "asr_profile_enable")

    def profile_read(self) -> dict:
        """{family: (total_ms, launches)} accumulated since the last read (device time, CUDA events on the engine stream)."""
        n = len(self.PROF_NAMES)
        ms = (C.c_double * n)()
        cnt = (C.c_uint64 * n)()
        _lib.check(self.lib, self.lib.asr_profile_read(self._h, ms, cnt), "asr_profile_read")
        return {self.PROF_NAMES[i]: (ms[i], int(cnt[i])) for i in range(n)}

    def debug_step_partial(self, slots: Sequence[int], pcm: np.ndarray, n_layers: int) -> None:
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        a, fmt = self._pcm(pcm, int(sl.size))
        _lib.check(self.lib, self.lib.asr_debug_step_partial(self._h, int(sl.size), sl.ctypes.data, a.ctypes.data, fmt, n_layers), "asr_debug_step_partial")

    def debug_decode_logits(self, slots: Sequence[int], logits: np.ndarray, want_logprobs: bool = False) -> StepResult:
        """The decode stage alone (log_softmax, greedy collapse, prefix beam) on caller-supplied CTC logits [n, seg_rows, vocab]."""
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        n = int(sl.size)
        z = np.ascontiguousarray(logits, np.float32)
        assert z.shape == (n, self.cfg.seg_rows, self.cfg.vocab), z.shape
        bufs, o = self._alloc_out(n, want_logprobs)
        _lib.check(self.lib, self.lib.asr_debug_decode_logits(self._h, n, sl.ctypes.data, z.ctypes.data, C.byref(o)), "asr_debug_decode_logits")
        return self._result(bufs, n)

    def debug_read(self, which: int, shape) -> np.ndarray:
        out = np.empty(shape, np.float32)
        _lib.check(self.lib, self.lib.asr_debug_read(self._h, which, out.ctypes.data, out.size), "asr_debug_read")
        return out

    def debug_read_state(self, slot: int, layer: int, which: int):
        out = np.empty((self.cfg.left_context, self.cfg.d_model), np.float32)
        pl = C.c_int32()
        _lib.check(self.lib, self.lib.asr_debug_read_state(self._h, slot, layer, which, out.ctypes.data, C.byref(pl)), "asr_debug_read_state")
        return out, pl.value


def debug_gemm(A: np.ndarray, B: np.ndarray, bias: Optional[np.ndarray] = None, impl: int = 0, split: int = 0, bn: int = 128,
               device: int = 0) -> np.ndarray:
    """C = A @ B.T (+bias) through the tcgen05 GEMM."""
    lib = _lib.load_library()
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    M, K = A.shape
    N = B.shape[0]
    Cm = np.empty((M, N), np.float32)
    b = np.ascontiguousarray(bias, np.float32) if bias is not None else None
    _lib.check(lib, lib.asr_debug_gemm(impl, M, N, K, split, bn, A.ctypes.data, B.ctypes.data, b.ctypes.data if b is not None else None,
                                       Cm.ctypes.data, device), "asr_debug_gemm")
    return Cm


def debug_gemm_operand(A: np.ndarray, B: np.ndarray, bias: np.ndarray, act: int = 1, bn: int = 515, device: int = 0) -> np.ndarray:
    """bf16(act(A @ B.T + bias)) as fp32 through the EpiOperand epilogue (bn 512: LSU stores, 515: TMA stores; act 0 none, 1 GELU, 2 SiLU)."""
    lib = _lib.load_library()
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    bias = np.ascontiguousarray(bias, np.float32)
    M, K = A.shape
    N = B.shape[0]
    out = np.empty((M, N), np.float32)
    _lib.check(lib, lib.asr_debug_gemm_operand(M, N, K, bn, act, A.ctypes.data, B.ctypes.data, bias.ctypes.data, out.ctypes.data, device),
               "asr_debug_gemm_operand")
    return out


def debug_gemm_ln(A: np.ndarray, W: np.ndarray, bias: np.ndarray, res: np.ndarray, g1: np.ndarray, b1: np.ndarray,
                  g2: Optional[np.ndarray] = None, b2: Optional[np.ndarray] = None, f32_normed: bool = False, compact_rows: int = 0,
                  compact_seg: int = 0, split: int = 0, iters: int = 1, pair: int = 0, device: int = 0):
    """(out_f32 [M,512], out_op [M or compacted,512] as fp32, ms per launch) of the GEMM + residual + LayerNorm kernel (gemm_ln.cu).
    pair: 0 = cluster of 2, 1 = cta_group::2 shape (cluster of 4), 2 = that shape with the one-pass second-LN statistics,
    3 = cluster of 4 column quarters (small M)."""
    lib = _lib.load_library()
    f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    A, W, bias, res, g1, b1, g2, b2 = map(f, (A, W, bias, res, g1, b1, g2, b2))
    M, K = A.shape
    assert W.shape == (512, K) and res.shape == (M, 512)
    Mo = M // compact_rows * compact_seg if compact_rows else M
    out, op = np.empty((M, 512), np.float32), np.empty((Mo, 512), np.float32)
    ms = C.c_float()
    p = lambda a: None if a is None else a.ctypes.data
    _lib.check(lib, lib.asr_debug_gemm_ln(M, K, split, p(A), p(W), p(bias), p(res), p(g1), p(b1), p(g2), p(b2), int(f32_normed), compact_rows,
                                          compact_seg, p(out), p(op), iters, C.byref(ms), int(pair), device), "asr_debug_gemm_ln")
    return out, op, ms.value
