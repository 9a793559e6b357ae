"""Streaming convolution module with per-session cache: host front of ``asr_convmod_*`` (csrc/convmod.cu).

Streaming form of the reference's ``ConvolutionBlock`` (lightspeech/layers/block.py:129-171).  ``step`` returns the block's output
on the whole utterance delayed by ``(kernel - 1) // 2`` frames: chunk by chunk, with the last ``kernel - 1`` activated frames of
every session cached on the device.  Not used by the Emformer path (which has no convolution module); provided for encoders built
from ``SqueezeformerBlock`` (block.py:9-75)."""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Sequence

import numpy as np

from . import _lib
from .config import PRECISION_FAST

# reference parameter names of ConvolutionBlock.state_dict() in blob order (convmod.cu: asr_convmod_create)
PARAM_ORDER = ("pre_norm.scale", "pre_norm.bias", "pointwise_conv1.weight", "pointwise_conv1.bias", "depthwise_conv.weight",
               "depthwise_conv.bias", "norm.weight", "norm.bias", "norm.running_mean", "norm.running_var", "pointwise_conv2.weight",
               "pointwise_conv2.bias")


def pack_conv_weights(state: Mapping[str, np.ndarray], d_model: int, kernel: int) -> np.ndarray:
    """ConvolutionBlock state_dict (numpy arrays; Conv1d weights [d, d, 1] / [d, 1, k]) -> flat fp32 blob."""
    shapes = {"pointwise_conv1.weight": d_model * d_model, "pointwise_conv2.weight": d_model * d_model, "depthwise_conv.weight": d_model * kernel}
    parts = []
    for name in PARAM_ORDER:
        a = np.asarray(state[name], np.float32).reshape(-1)
        want = shapes.get(name, d_model)
        if a.size != want:
            raise ValueError(f"{name}: {a.size} values, expected {want}")
        parts.append(a)
    return np.ascontiguousarray(np.concatenate(parts))


class ConvModule:
    def __init__(self, d_model: int, kernel: int, rows_per_chunk: int, weights: np.ndarray, max_sessions: int = 64, max_batch: int = 64,
                 precision: int = PRECISION_FAST, device: int = 0):
        self.lib = _lib.load_library()
        self.d, self.k, self.T = d_model, kernel, rows_per_chunk
        n = C.c_uint64()
        _lib.check(self.lib, self.lib.asr_convmod_weights_count(d_model, kernel, C.byref(n)), "asr_convmod_weights_count")
        w = np.ascontiguousarray(weights, np.float32).reshape(-1)
        if w.size != n.value:
            raise ValueError(f"weights blob has {w.size} floats, the module needs {n.value}")
        h = C.c_void_p()
        _lib.check(self.lib, self.lib.asr_convmod_create(d_model, kernel, rows_per_chunk, max_sessions, max_batch, precision, w.ctypes.data, w.size,
                                                         device, C.byref(h)), "asr_convmod_create")
        self._h = h

    @property
    def delay(self) -> int:
        """Output latency in frames: y[t] of ``step`` is the reference block's output at frame t - delay."""
        return (self.k - 1) // 2

    def step(self, slots: Sequence[int], x: np.ndarray) -> np.ndarray:
        """x: [n, rows_per_chunk, d_model] float32 -> same shape."""
        sl = np.ascontiguousarray(slots, np.int32)
        a = np.ascontiguousarray(x, np.float32)
        if a.shape != (sl.size, self.T, self.d):
            raise ValueError(f"x has shape {a.shape}, expected {(sl.size, self.T, self.d)}")
        y = np.empty_like(a)
        _lib.check(self.lib, self.lib.asr_convmod_step(self._h, int(sl.size), sl.ctypes.data, a.ctypes.data, y.ctypes.data), "asr_convmod_step")
        return y

    def reset(self, slots: Sequence[int]) -> None:
        sl = np.ascontiguousarray(slots, np.int32)
        _lib.check(self.lib, self.lib.asr_convmod_reset(self._h, int(sl.size), sl.ctypes.data), "asr_convmod_reset")

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.asr_convmod_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
