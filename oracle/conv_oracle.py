"""CPU restatement of the reference's ConvolutionBlock (lightspeech/layers/block.py:129-171) and of its streaming (cached) form.
TEST INFRASTRUCTURE ONLY: imported by tests/ and oracle/make_conv_goldens.py, never by the product.

Pin: ``tests/golden/convblock.npz`` holds the output of the UNMODIFIED reference module (imported in the build container by
oracle/make_conv_goldens.py) on a seeded input with the seeded weights below; tests check ``conv_block_full`` against it, the
streaming form against ``conv_block_full`` (delay (k-1)/2), and the CUDA module against both."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def make_conv_weights(seed: int, d: int = 512, k: int = 31) -> Dict[str, np.ndarray]:
    """Deterministic, non-trivial parameters in the reference's state_dict layout (ConvolutionBlock.__init__, block.py:130-152)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = np.float32
    return {
        "pre_norm.scale": (1.0 + 0.1 * rng.standard_normal(d)).astype(f), "pre_norm.bias": (0.1 * rng.standard_normal(d)).astype(f),
        "pointwise_conv1.weight": (rng.standard_normal((d, d, 1)) / np.sqrt(d)).astype(f), "pointwise_conv1.bias": (0.1 * rng.standard_normal(d)).astype(f),
        "depthwise_conv.weight": (rng.standard_normal((d, 1, k)) / np.sqrt(k)).astype(f), "depthwise_conv.bias": (0.1 * rng.standard_normal(d)).astype(f),
        "norm.weight": (1.0 + 0.2 * rng.standard_normal(d)).astype(f), "norm.bias": (0.1 * rng.standard_normal(d)).astype(f),
        "norm.running_mean": (0.2 * rng.standard_normal(d)).astype(f), "norm.running_var": (0.5 + rng.random(d)).astype(f),
        "pointwise_conv2.weight": (rng.standard_normal((d, d, 1)) / np.sqrt(d)).astype(f), "pointwise_conv2.bias": (0.1 * rng.standard_normal(d)).astype(f),
    }


def _silu(x):
    return x / (1.0 + np.exp(-x))


def _front(x: np.ndarray, W) -> np.ndarray:
    """pre_norm -> pointwise_conv1 -> SiLU (block.py:155-158), per frame.  x: [T, d] -> [T, d]."""
    u = W["pre_norm.scale"] * x + W["pre_norm.bias"]                                   # normalization.py:15-19
    h = u @ W["pointwise_conv1.weight"][:, :, 0].T + W["pointwise_conv1.bias"]
    return _silu(h)


def _back(z: np.ndarray, W) -> np.ndarray:
    """BatchNorm1d (eval) -> SiLU -> pointwise_conv2 (block.py:163-166), per frame."""
    a = W["norm.weight"] / np.sqrt(W["norm.running_var"] + 1e-5)
    y = _silu((z - W["norm.running_mean"]) * a + W["norm.bias"])
    return y @ W["pointwise_conv2.weight"][:, :, 0].T + W["pointwise_conv2.bias"]


def conv_block_full(x: np.ndarray, W) -> np.ndarray:
    """The reference block on a whole utterance (masks all False): zero padding (k-1)/2 on both sides (block.py:137-143)."""
    x = x.astype(np.float64)
    W = {k: v.astype(np.float64) for k, v in W.items()}
    h = _front(x, W)
    w = W["depthwise_conv.weight"][:, 0, :]                                             # [d, k]
    k = w.shape[1]
    pad = (k - 1) // 2
    hp = np.concatenate([np.zeros((pad, h.shape[1])), h, np.zeros((pad, h.shape[1]))])
    z = np.stack([(hp[t:t + k] * w.T).sum(0) for t in range(h.shape[0])]) + W["depthwise_conv.bias"]
    return _back(z, W).astype(np.float32)


def init_conv_state(d: int = 512, k: int = 31) -> np.ndarray:
    return np.zeros((k - 1, d), np.float64)


def conv_block_stream(x: np.ndarray, state: np.ndarray, W) -> Tuple[np.ndarray, np.ndarray]:
    """One chunk [T, d] with the k-1 cached activated frames: causal window => the full-sequence output delayed by (k-1)/2 frames."""
    W64 = {k: v.astype(np.float64) for k, v in W.items()}
    h = _front(x.astype(np.float64), W64)
    w = W64["depthwise_conv.weight"][:, 0, :]
    k = w.shape[1]
    cat = np.concatenate([state, h])
    z = np.stack([(cat[t:t + k] * w.T).sum(0) for t in range(h.shape[0])]) + W64["depthwise_conv.bias"]
    return _back(z, W64).astype(np.float32), cat[-(k - 1):]
