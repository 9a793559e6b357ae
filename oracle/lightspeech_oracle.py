"""CPU oracle for the lightspeech per-chunk compute path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``asr_streaming_b200``) never imports anything
from ``oracle/`` and fails loudly when the CUDA library is missing.

It is a plain-numpy restatement of the algorithm the reference runs for one stream-chunk
(batch 1 — the only numerically correct mode of the reference for ragged progress, see
SURVEY.md §0).  Each function cites the reference lines it follows.  ``path:line`` is
relative to the reference repo root; ``TA:`` is the un-vendored third-party dependency
**torchaudio** (reference pins no version — ``Dockerfile:39`` installs ``torch torchaudio``
unpinned; this container has torchaudio 2.11.0), whose published algorithm
(``torchaudio.models.Emformer``, ``torchaudio.transforms.MelSpectrogram``) is restated here.

Parity pin: the reference has no golden vectors or tests for this path, so
the oracle is pinned against outputs of *the reference itself*, imported in the build
container through ``oracle/ref_import.py`` and stored as fixtures under ``tests/golden/``
by ``oracle/make_goldens.py`` (committed).  ``tests/test_oracle_golden.py`` checks this
file against those fixtures.

Precision: every function takes ``dtype`` (np.float32 mirrors the reference's fp32 math up
to summation order; np.float64 is the "true value" used when measuring error budgets).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# Geometry (streaming_decoder/utils.py:9-23, config/asr-online.yaml:112-118)
# --------------------------------------------------------------------------------------


@dataclass(frozen=True)
class Geometry:
    """Chunk geometry.  Mirrors ``AudioConfig`` (utils.py:9-23) + model hyper-parameters
    (recognition.py:207-217, encoder.py:73-117; ffn_dim / ctc hidden are ASSUMED)."""

    sample_rate: int = 16000
    hop: int = 160                 # int(0.01 * sr)            utils.py:16
    n_fft: int = 800               # int(0.05 * sr)            audio.py:17
    win: int = 400                 # int(0.025 * sr)           audio.py:18
    n_mels: int = 128              # audio.py:20
    segment_size: int = 64         # frames                    asr-online.yaml:115
    context_size: int = 16         # frames                    asr-online.yaml:116
    bias: int = 4                  # frames                    asr-online.yaml:117
    stride: int = 4                # subsampling_factor        encoder.py:82
    d_model: int = 512
    n_heads: int = 8
    ffn_dim: int = 2048
    n_layers: int = 20
    left_context: int = 32         # rows (after stride)       recognition.py:211
    ctc_hidden: int = 512
    vocab: int = 804
    framerate_sec: float = 0.04    # FRAMERATE                 recognition.py:30

    @property
    def segment_length(self) -> int:      # utils.py:18
        return self.segment_size * self.hop

    @property
    def buffer_length(self) -> int:       # utils.py:21
        return (self.context_size + self.bias) * self.hop

    @property
    def chunk_length(self) -> int:        # utils.py:22
        return self.segment_length + self.buffer_length

    @property
    def frames_per_chunk(self) -> int:    # torch.stft, center=False
        return 1 + (self.chunk_length - self.n_fft) // self.hop

    @property
    def seg_rows(self) -> int:            # Emformer segment_length // stride  encoder.py:109
        return self.segment_size // self.stride

    @property
    def rc_rows(self) -> int:             # right_context_length // stride     encoder.py:112
        return self.context_size // self.stride

    @property
    def rows(self) -> int:
        return self.seg_rows + self.rc_rows


CANONICAL = Geometry()
LOW_LATENCY = Geometry(segment_size=32)   # config #5 "chunk_size=8": 8 output rows, T = 12


# --------------------------------------------------------------------------------------
# a3: extract_filterbank  (lightspeech/datas/audio.py:9-30)
# --------------------------------------------------------------------------------------


def hann_periodic(n: int, dtype=np.float64) -> np.ndarray:
    """torch.hann_window(n, periodic=True) — TA:transforms/_transforms.py:109 (Spectrogram window_fn)."""
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(dtype)


def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int,
                        dtype=np.float32) -> np.ndarray:
    """TA:functional/functional.py:492-587 (melscale_fbanks, norm=None, mel_scale='htk').

    Evaluated in float32 like torchaudio does (torch.linspace default dtype), so that the
    triangle weights are the same numbers the reference multiplies by."""
    f32 = np.float32
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs).astype(f32)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = np.linspace(m_min, m_max, n_mels + 2).astype(f32)
    f_pts = (f32(700.0) * (np.power(f32(10.0), m_pts / f32(2595.0)) - f32(1.0))).astype(f32)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(f32(0.0), np.minimum(down, up))
    return fb.astype(dtype)                                   # [n_freqs, n_mels]


def melspec128(pcm: np.ndarray, geo: Geometry = CANONICAL, dtype=np.float64) -> np.ndarray:
    """audio.py:15-26: MelSpectrogram(n_fft=800, win=400, hop=160, n_mels=128, center=False)
    -> clamp(1e-5).log() -> transpose.  ``pcm`` is float in [-1, 1) of length >= n_fft.
    Returns [frames, n_mels]."""
    x = np.asarray(pcm, dtype=dtype)
    n_frames = 1 + (x.shape[0] - geo.n_fft) // geo.hop
    # torch.stft pads the window (win_length < n_fft) with zeros on both sides, centred.
    w = np.zeros(geo.n_fft, dtype=dtype)
    left = (geo.n_fft - geo.win) // 2
    w[left:left + geo.win] = hann_periodic(geo.win, dtype)
    idx = np.arange(geo.n_fft)[None, :] + geo.hop * np.arange(n_frames)[:, None]
    frames = x[idx] * w[None, :]
    spec = np.fft.rfft(frames.astype(np.float64), axis=1)
    power = (spec.real ** 2 + spec.imag ** 2).astype(dtype)    # power=2.0
    fb = melscale_fbanks_htk(geo.n_fft // 2 + 1, 0.0, geo.sample_rate / 2.0, geo.n_mels,
                             geo.sample_rate, dtype)
    mel = power @ fb
    return np.log(np.maximum(mel, dtype(1e-5))).astype(dtype)


# --------------------------------------------------------------------------------------
# Weights: names follow the reference checkpoint's state_dict["encoder"/"decoder"] layout
# (recognition.py:151-157 + torchaudio Emformer parameter names).
# --------------------------------------------------------------------------------------


def layer_prefix(i: int) -> str:
    return f"encoder_layers.emformer_layers.{i}."


def weight_shapes(geo: Geometry = CANONICAL) -> Dict[str, Tuple[int, ...]]:
    d, f = geo.d_model, geo.ffn_dim
    shapes: Dict[str, Tuple[int, ...]] = {"encoder.input_linear.weight": (d // geo.stride, geo.n_mels)}
    for i in range(geo.n_layers):
        p = "encoder." + layer_prefix(i)
        shapes[p + "attention.emb_to_key_value.weight"] = (2 * d, d)
        shapes[p + "attention.emb_to_key_value.bias"] = (2 * d,)
        shapes[p + "attention.emb_to_query.weight"] = (d, d)
        shapes[p + "attention.emb_to_query.bias"] = (d,)
        shapes[p + "attention.out_proj.weight"] = (d, d)
        shapes[p + "attention.out_proj.bias"] = (d,)
        shapes[p + "pos_ff.0.weight"] = (d,)
        shapes[p + "pos_ff.0.bias"] = (d,)
        shapes[p + "pos_ff.1.weight"] = (f, d)
        shapes[p + "pos_ff.1.bias"] = (f,)
        shapes[p + "pos_ff.4.weight"] = (d, f)
        shapes[p + "pos_ff.4.bias"] = (d,)
        shapes[p + "layer_norm_input.weight"] = (d,)
        shapes[p + "layer_norm_input.bias"] = (d,)
        shapes[p + "layer_norm_output.weight"] = (d,)
        shapes[p + "layer_norm_output.bias"] = (d,)
    shapes["decoder.linear1.weight"] = (geo.ctc_hidden, d)
    shapes["decoder.linear1.bias"] = (geo.ctc_hidden,)
    shapes["decoder.linear2.weight"] = (geo.vocab, geo.ctc_hidden)
    shapes["decoder.linear2.bias"] = (geo.vocab,)
    return shapes


def make_weights(seed: int, geo: Geometry = CANONICAL, ctc_gain: float = 1.0) -> Dict[str, np.ndarray]:
    """Seeded random-init weights (no checkpoint ships with the reference).

    Distribution: uniform(+-1/sqrt(fan_in)) for Linear weight/bias (torch.nn.Linear default),
    layer-norm gamma ~ 1 + 0.1 N(0,1), beta ~ 0.1 N(0,1) so that the affine terms are exercised.
    Generated with numpy's PCG64 so the same seed gives the same bytes on every box; the
    reference model is loaded with exactly these arrays when goldens are made.
    ``ctc_gain`` scales decoder.linear2 (weight and bias)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out: Dict[str, np.ndarray] = {}
    for name, shape in weight_shapes(geo).items():
        if ("layer_norm" in name or "pos_ff.0" in name):
            if name.endswith("weight"):
                a = 1.0 + 0.1 * rng.standard_normal(shape)
            else:
                a = 0.1 * rng.standard_normal(shape)
        else:
            fan_in = shape[-1] if len(shape) == 2 else None
            if fan_in is None:                      # bias: fan_in of its weight
                wname = name[:-4] + "weight"
                fan_in = weight_shapes(geo)[wname][-1]
            bound = 1.0 / math.sqrt(fan_in)
            a = rng.uniform(-bound, bound, size=shape)
        if name.startswith("decoder.linear2"):
            a = a * ctc_gain
        out[name] = a.astype(np.float32)
    return out


# --------------------------------------------------------------------------------------
# a4..a8: encoder
# --------------------------------------------------------------------------------------

MatMul = Callable[[np.ndarray, np.ndarray], np.ndarray]   # (x[M,K], w[N,K]) -> x @ w.T


def _mm(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    return x @ w.T


def layer_norm(x: np.ndarray, g: np.ndarray, b: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """torch.nn.LayerNorm(512) (TA:emformer.py:362-365, :375-376): biased variance, eps 1e-5."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * g + b


def gelu_erf(x: np.ndarray) -> np.ndarray:
    """torch.nn.GELU() default = exact erf form (TA:emformer.py:42-43)."""
    erf = np.vectorize(math.erf, otypes=[np.float64])
    return (0.5 * x * (1.0 + erf(x.astype(np.float64) / math.sqrt(2.0)))).astype(x.dtype)


def silu(x: np.ndarray) -> np.ndarray:
    return x / (1.0 + np.exp(-x))


@dataclass
class LayerState:
    """Per-layer streaming state — TA:emformer.py:384-389 (_init_state): fixed 32-row K/V left
    context (right-aligned) + past_length.  Memory bank is size 0 (recognition.py:210)."""
    k: np.ndarray
    v: np.ndarray
    past_length: int = 0


def init_state(geo: Geometry = CANONICAL, dtype=np.float32) -> List[LayerState]:
    """recognition.py:207-217 (LightningASR.init_state)."""
    return [LayerState(np.zeros((geo.left_context, geo.d_model), dtype),
                       np.zeros((geo.left_context, geo.d_model), dtype), 0)
            for _ in range(geo.n_layers)]


def emformer_layer_infer(utt: np.ndarray, rc: np.ndarray, st: LayerState, W: Dict[str, np.ndarray],
                         p: str, geo: Geometry, mm: MatMul = _mm,
                         taps: Optional[Dict[str, np.ndarray]] = None
                         ) -> Tuple[np.ndarray, np.ndarray, LayerState]:
    """One ``_EmformerLayer.infer`` (TA:emformer.py:542-588) with max_memory_size = 0, batch 1.

    utt [S,d], rc [R,d]  ->  (utt_out [S,d], rc_out [R,d], new state)."""
    S, R, d, H = utt.shape[0], rc.shape[0], geo.d_model, geo.n_heads
    dh = d // H
    dt = utt.dtype
    x = np.concatenate([rc, utt], axis=0)                               # :431  cat([right_context, utterance])
    ln = layer_norm(x, W[p + "layer_norm_input.weight"], W[p + "layer_norm_input.bias"]).astype(dt)
    # _unpack_state :391-398 — valid left context = min(left_context_length, past_length)
    lv = min(geo.left_context, st.past_length)
    lc_k = st.k[geo.left_context - lv:]
    lc_v = st.v[geo.left_context - lv:]
    # attention.infer :256-316 -> _forward_impl :146-217 (summary empty, mems empty)
    q = (mm(ln, W[p + "attention.emb_to_query.weight"]) + W[p + "attention.emb_to_query.bias"]).astype(dt)   # :161
    kv = (mm(ln, W[p + "attention.emb_to_key_value.weight"]) + W[p + "attention.emb_to_key_value.bias"]).astype(dt)  # :164
    k_new, v_new = kv[:, :d], kv[:, d:]                                  # chunk(2, dim=2)
    key = np.concatenate([k_new[:R], lc_k, k_new[R:]], axis=0)           # :166-173
    val = np.concatenate([v_new[:R], lc_v, v_new[R:]], axis=0)           # :174-181
    scaling = dt.type(dh ** -0.5)                                        # :108
    attn = np.empty((R + S, d), dtype=dt)
    for h in range(H):
        sl = slice(h * dh, (h + 1) * dh)
        s = ((q[:, sl] * scaling) @ key[:, sl].T).astype(np.float32 if dt == np.float32 else np.float64)  # :188
        s = s - s.max(axis=1, keepdims=True)                             # softmax fp32 :133-143 (mask all-False)
        e = np.exp(s)
        prob = (e / e.sum(axis=1, keepdims=True)).astype(dt)
        attn[:, sl] = prob @ val[:, sl]                                  # :196
    out = (mm(attn, W[p + "attention.out_proj.weight"]) + W[p + "attention.out_proj.bias"]).astype(dt)       # :207
    # _pack_state :400-414 — cached K/V = [left ctx, utterance rows]; rc rows are NOT cached (:314-315)
    next_k = np.concatenate([lc_k, k_new[R:]], axis=0)
    next_v = np.concatenate([lc_v, v_new[R:]], axis=0)
    new_k = np.concatenate([st.k, next_k], axis=0)[-geo.left_context:]
    new_v = np.concatenate([st.v, next_v], axis=0)[-geo.left_context:]
    new_st = LayerState(new_k, new_v, st.past_length + S)
    # _process_attention_output :416-425
    r = out + x                                                          # residual = pre-LN input
    h1 = layer_norm(r, W[p + "pos_ff.0.weight"], W[p + "pos_ff.0.bias"]).astype(dt)
    h2 = gelu_erf((mm(h1, W[p + "pos_ff.1.weight"]) + W[p + "pos_ff.1.bias"]).astype(dt))
    ff = (mm(h2, W[p + "pos_ff.4.weight"]) + W[p + "pos_ff.4.bias"]).astype(dt)
    r2 = ff + r
    y = layer_norm(r2, W[p + "layer_norm_output.weight"], W[p + "layer_norm_output.bias"]).astype(dt)
    if taps is not None:
        taps["ln_in"] = ln; taps["q"] = q; taps["kv"] = kv; taps["attn"] = attn
        taps["x1"] = r; taps["x2"] = r2; taps["y"] = y
    return y[R:], y[:R], new_st


def encoder_infer(fbank: np.ndarray, states: List[LayerState], W: Dict[str, np.ndarray],
                  geo: Geometry = CANONICAL, mm: MatMul = _mm,
                  taps: Optional[Dict[str, np.ndarray]] = None
                  ) -> Tuple[np.ndarray, List[LayerState]]:
    """StreamingAcousticEncoder.infer (encoder.py:134-147) for one stream.
    fbank [frames, n_mels] -> enc_out [seg_rows, d_model]."""
    dt = fbank.dtype
    x = mm(fbank, W["encoder.input_linear.weight"].astype(dt)).astype(dt)     # encoder.py:142 (no bias)
    t = x.shape[0]
    assert t % geo.stride == 0
    x = x.reshape(t // geo.stride, geo.d_model)                               # time_reduction common.py:118-119
    assert x.shape[0] == geo.rows, "Emformer.infer size check (TA:emformer.py:775-780)"
    rc = x[geo.seg_rows:]                                                     # TA:emformer.py:782-784
    utt = x[:geo.seg_rows]
    new_states = []
    for i in range(geo.n_layers):                                             # TA:emformer.py:793-801
        lt = {} if (taps is not None) else None
        Wc = W if dt == np.float32 else _CastView(W, dt)
        utt, rc, ns = emformer_layer_infer(utt, rc, states[i], Wc, "encoder." + layer_prefix(i), geo, mm, lt)
        new_states.append(ns)
        if taps is not None:
            for kname, v in lt.items():
                taps[f"L{i}.{kname}"] = v
    if taps is not None:
        taps["input_linear"] = x
    return utt, new_states


class _CastView(dict):
    """Lazily casts fp32 weights to ``dt`` (float64 runs)."""

    def __init__(self, base: Dict[str, np.ndarray], dt):
        super().__init__()
        self._b, self._dt = base, dt

    def __getitem__(self, k):
        if k not in self.keys():
            super().__setitem__(k, self._b[k].astype(self._dt))
        return super().__getitem__(k)


def ctc_head(enc: np.ndarray, W: Dict[str, np.ndarray], mm: MatMul = _mm) -> np.ndarray:
    """CTCDecoder.forward (lightspeech/modules/decoder.py:66-70) -> log-probs [rows, V]."""
    dt = enc.dtype
    h = silu((mm(enc, W["decoder.linear1.weight"].astype(dt)) + W["decoder.linear1.bias"].astype(dt)).astype(dt))
    z = (mm(h.astype(dt), W["decoder.linear2.weight"].astype(dt)) + W["decoder.linear2.bias"].astype(dt)).astype(dt)
    z = z - z.max(axis=1, keepdims=True)
    return (z - np.log(np.exp(z).sum(axis=1, keepdims=True))).astype(dt)


def stream_chunk(pcm_chunk: np.ndarray, states: List[LayerState], W: Dict[str, np.ndarray],
                 geo: Geometry = CANONICAL, dtype=np.float32, mm: MatMul = _mm,
                 taps: Optional[Dict[str, np.ndarray]] = None
                 ) -> Tuple[np.ndarray, List[LayerState]]:
    """LightningASR.stream (recognition.py:191-204) for ONE stream (batch 1):
    pcm [chunk_length] float -> (log-probs [seg_rows, V], new states)."""
    fb = melspec128(pcm_chunk, geo, dtype=np.float64).astype(dtype)
    if taps is not None:
        taps["fbank"] = fb
    enc, ns = encoder_infer(fb, states, W, geo, mm, taps)
    if taps is not None:
        taps["enc_out"] = enc
    return ctc_head(enc, W, mm), ns


# --------------------------------------------------------------------------------------
# a13: greedy_search  (recognition.py:33-57)
# --------------------------------------------------------------------------------------


def greedy_ids(emission: np.ndarray, framerate: float = 0.04) -> Tuple[List[int], float, np.ndarray]:
    """argmax -> unique_consecutive -> drop blank(0).  Returns (token ids, last_blank seconds, argmax path).
    ``last_blank`` uses ``indices > 1`` (ids 0 '-' and 1 '|' are not tokens), recognition.py:38-43."""
    idx = np.argmax(emission, axis=1).astype(np.int64)
    n = idx.shape[0]
    last_blank = framerate * n
    tok = np.nonzero(idx > 1)[0]
    if tok.size:
        last_blank = float((n - 1 - tok[-1]) * framerate)
    if n:
        keep = np.concatenate([[True], idx[1:] != idx[:-1]])
        uniq = idx[keep]
    else:
        uniq = idx
    ids = [int(i) for i in uniq if i != 0]
    return ids, last_blank, idx


def ids_to_text(ids: Sequence[int], vocab: Sequence[str]) -> str:
    """recognition.py:47-52."""
    text = "".join(vocab[i] for i in ids)
    text = text.replace("<<", "").replace(">>", "")
    text = text.replace("-", "").replace("|", " ")
    return re.sub(r"\s+", " ", text).strip()


def greedy_search(emission: np.ndarray, vocab: Sequence[str]) -> Tuple[str, float]:
    ids, last_blank, _ = greedy_ids(emission)
    return ids_to_text(ids, vocab), last_blank


# --------------------------------------------------------------------------------------
# Session driver: the caller's buffer semantics (stream.py:23-26, :78-87, :159-160;
# streaming_server.py:371, :384, :420-435)
# --------------------------------------------------------------------------------------


def chunk_windows(audio: np.ndarray, geo: Geometry = CANONICAL) -> List[np.ndarray]:
    """Chunk k = samples [k*segment_length - buffer_length, k*segment_length + segment_length) with
    ``buffer_length`` leading zeros (stream.py:23), advancing by segment_length (stream.py:159)."""
    buf = np.concatenate([np.zeros(geo.buffer_length, dtype=audio.dtype), audio])
    out = []
    while buf.shape[0] >= geo.chunk_length:                       # streaming_server.py:371
        out.append(buf[:geo.chunk_length].copy())                 # :384
        buf = buf[geo.segment_length:]
    return out


def run_stream(audio: np.ndarray, W: Dict[str, np.ndarray], geo: Geometry = CANONICAL, dtype=np.float32,
               mm: MatMul = _mm, max_chunks: Optional[int] = None) -> np.ndarray:
    """All chunks of one stream, no VAD gating, no endpoint -> emission [n_chunks*seg_rows, V]."""
    st = init_state(geo, dtype)
    ems = []
    for i, ch in enumerate(chunk_windows(audio, geo)):
        if max_chunks is not None and i >= max_chunks:
            break
        em, st = stream_chunk(ch, st, W, geo, dtype, mm)
        ems.append(em)
    return np.concatenate(ems, axis=0) if ems else np.zeros((0, geo.vocab), dtype)


# --------------------------------------------------------------------------------------
# Operand-precision emulation (used by tests to derive the stated tolerances)
# --------------------------------------------------------------------------------------


def to_bf16(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def mm_bf16(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """bf16 operands, fp32 accumulate (the product's FAST mode)."""
    return to_bf16(x).astype(np.float32) @ to_bf16(w).astype(np.float32).T


def mm_bf16x3(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Split-bf16 (hi+lo) operands, 3 products, fp32 accumulate (the product's EXACT mode)."""
    x = x.astype(np.float32); w = w.astype(np.float32)
    xh = to_bf16(x); xl = to_bf16(x - xh)
    wh = to_bf16(w); wl = to_bf16(w - wh)
    return xh @ wh.T + (xl @ wh.T + xh @ wl.T)
