"""Weight packing: reference checkpoint layout -> the flat fp32 blob asr_engine_create expects.

The reference loads ``torch.load(ckpt)["state_dict"]["encoder"|"decoder"]`` into StreamingAcousticEncoder /
CTCDecoder (lightspeech/models/recognition.py:149-159, lightspeech/utils/common.py:127-139); parameter names are
torchaudio Emformer's.  Blob order (all fp32, row-major, nn.Linear [out, in]):

    encoder.input_linear.weight                              [d/stride, n_mels]
    per layer i (prefix encoder.encoder_layers.emformer_layers.{i}.):
        Wqkv = cat(attention.emb_to_query.weight, attention.emb_to_key_value.weight)   [3d, d]
        bqkv = cat(attention.emb_to_query.bias,   attention.emb_to_key_value.bias)     [3d]
        attention.out_proj.weight [d, d], .bias [d]
        layer_norm_input.weight, .bias            [d], [d]
        pos_ff.0.weight, .bias  (FFN LayerNorm)   [d], [d]
        pos_ff.1.weight [ffn, d], .bias [ffn]
        pos_ff.4.weight [d, ffn], .bias [d]
        layer_norm_output.weight, .bias           [d], [d]
    decoder.linear1.weight [H, d], .bias [H], decoder.linear2.weight [V, H], .bias [V]
"""
from __future__ import annotations

import math
from typing import Dict, Mapping

import numpy as np

from .config import ModelConfig


def _layer_prefix(i: int) -> str:
    return f"encoder.encoder_layers.emformer_layers.{i}."


def _np(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().float().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


def pack_weights(sd: Mapping[str, "np.ndarray"], cfg: ModelConfig = ModelConfig()) -> np.ndarray:
    """``sd`` keys are 'encoder.<name>' / 'decoder.<name>' (numpy arrays or torch tensors)."""
    d, f = cfg.d_model, cfg.ffn_dim
    parts = []

    def take(name, shape):
        if name not in sd:
            raise KeyError(f"missing weight {name}")
        a = _np(sd[name])
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name}: shape {a.shape} != expected {shape}")
        return a

    parts.append(take("encoder.input_linear.weight", (d // cfg.stride, cfg.n_mels)))
    for i in range(cfg.n_layers):
        p = _layer_prefix(i)
        parts.append(take(p + "attention.emb_to_query.weight", (d, d)))
        parts.append(take(p + "attention.emb_to_key_value.weight", (2 * d, d)))
        parts.append(take(p + "attention.emb_to_query.bias", (d,)))
        parts.append(take(p + "attention.emb_to_key_value.bias", (2 * d,)))
        parts.append(take(p + "attention.out_proj.weight", (d, d)))
        parts.append(take(p + "attention.out_proj.bias", (d,)))
        parts.append(take(p + "layer_norm_input.weight", (d,)))
        parts.append(take(p + "layer_norm_input.bias", (d,)))
        parts.append(take(p + "pos_ff.0.weight", (d,)))
        parts.append(take(p + "pos_ff.0.bias", (d,)))
        parts.append(take(p + "pos_ff.1.weight", (f, d)))
        parts.append(take(p + "pos_ff.1.bias", (f,)))
        parts.append(take(p + "pos_ff.4.weight", (d, f)))
        parts.append(take(p + "pos_ff.4.bias", (d,)))
        parts.append(take(p + "layer_norm_output.weight", (d,)))
        parts.append(take(p + "layer_norm_output.bias", (d,)))
    parts.append(take("decoder.linear1.weight", (cfg.ctc_hidden, d)))
    parts.append(take("decoder.linear1.bias", (cfg.ctc_hidden,)))
    parts.append(take("decoder.linear2.weight", (cfg.vocab, cfg.ctc_hidden)))
    parts.append(take("decoder.linear2.bias", (cfg.vocab,)))
    return np.concatenate([a.reshape(-1) for a in parts]).astype(np.float32)


def weights_from_checkpoint(path: str, cfg: ModelConfig = ModelConfig()) -> np.ndarray:
    """Reads a reference checkpoint ({"hyper_parameters", "state_dict": {"encoder", "decoder"}}, recognition.py:151-157)."""
    import torch
    ck = torch.load(path, map_location="cpu", weights_only=False)
    sd = ck["state_dict"]
    flat: Dict[str, np.ndarray] = {}
    for part in ("encoder", "decoder"):
        for k, v in sd[part].items():
            flat[f"{part}.{k}"] = v
    return pack_weights(flat, cfg)


def random_weights(seed: int, cfg: ModelConfig = ModelConfig()) -> Dict[str, np.ndarray]:
    """Seeded random-init state dict with torch.nn.Linear-like scales (synthetic benchmarks; no checkpoint ships
    with the reference).  Same generator and draw order as the test oracle's make_weights, so a seed names one model."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d, f = cfg.d_model, cfg.ffn_dim
    shapes = {"encoder.input_linear.weight": (d // cfg.stride, cfg.n_mels)}
    for i in range(cfg.n_layers):
        p = _layer_prefix(i)
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
    shapes["decoder.linear1.weight"] = (cfg.ctc_hidden, d)
    shapes["decoder.linear1.bias"] = (cfg.ctc_hidden,)
    shapes["decoder.linear2.weight"] = (cfg.vocab, cfg.ctc_hidden)
    shapes["decoder.linear2.bias"] = (cfg.vocab,)
    out: Dict[str, np.ndarray] = {}
    for name, shape in shapes.items():
        if "layer_norm" in name or "pos_ff.0" in name:
            a = (1.0 + 0.1 * rng.standard_normal(shape)) if name.endswith("weight") else 0.1 * rng.standard_normal(shape)
        else:
            fan_in = shape[-1] if len(shape) == 2 else shapes[name[:-4] + "weight"][-1]
            bound = 1.0 / math.sqrt(fan_in)
            a = rng.uniform(-bound, bound, size=shape)
        out[name] = a.astype(np.float32)
    return out
