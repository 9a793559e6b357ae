"""Geometry of the hot path.  AudioConfig mirrors the reference class of the same name
(streaming_decoder/utils.py:9-23, values from config/asr-online.yaml:112-118)."""
from __future__ import annotations

from dataclasses import dataclass

PRECISION_FAST = 0    # bf16 operands, fp32 accumulate
PRECISION_EXACT = 1   # split-bf16 (hi+lo), 3 tcgen05 products, fp32 K/V cache


class AudioConfig(object):
    """Same attribute names and arithmetic as the reference AudioConfig (utils.py:9-23).
    ``config`` may be any object / mapping with sample_rate, hop_length (seconds), segment_size,
    context_size, bias, framerate; defaults are the reference's ``audio:`` block."""

    def __init__(self, config=None):
        get = (lambda k, d: (config.get(k, d) if isinstance(config, dict) else getattr(config, k, d))) if config is not None else (lambda k, d: d)
        self.sample_rate = get("sample_rate", 16000)
        self.hop_length = int(get("hop_length", 0.01) * self.sample_rate)          # utils.py:16
        self.segment_size = get("segment_size", 64)
        self.segment_length = self.segment_size * self.hop_length                  # utils.py:18
        self.context_size = get("context_size", 16)
        self.bias = get("bias", 4)
        self.buffer_length = int((self.context_size + self.bias) * self.hop_length)  # utils.py:21
        self.chunk_length = self.segment_length + self.buffer_length               # utils.py:22
        self.framerate = get("framerate", 4)


@dataclass(frozen=True)
class ModelConfig:
    """Model hyper-parameters.  ffn_dim / ctc_hidden are not in the reference repo (they live in
    the absent checkpoint's hyper_parameters) and default to the torchaudio emformer_rnnt_base values."""
    sample_rate: int = 16000
    hop: int = 160
    n_fft: int = 800
    win: int = 400
    n_mels: int = 128
    segment_size: int = 64
    context_size: int = 16
    bias: int = 4
    stride: int = 4
    d_model: int = 512
    n_heads: int = 8
    ffn_dim: int = 2048
    n_layers: int = 20
    left_context: int = 32
    ctc_hidden: int = 512
    vocab: int = 804
    precision: int = PRECISION_FAST
    max_sessions: int = 1024
    max_batch: int = 256

    @property
    def chunk_length(self) -> int:
        return (self.segment_size + self.context_size + self.bias) * self.hop

    @property
    def segment_length(self) -> int:
        return self.segment_size * self.hop

    @property
    def buffer_length(self) -> int:
        return (self.context_size + self.bias) * self.hop

    @property
    def seg_rows(self) -> int:
        return self.segment_size // self.stride

    @property
    def rc_rows(self) -> int:
        return self.context_size // self.stride

    @property
    def rows(self) -> int:
        return self.seg_rows + self.rc_rows

    @property
    def frames(self) -> int:
        return 1 + (self.chunk_length - self.n_fft) // self.hop
