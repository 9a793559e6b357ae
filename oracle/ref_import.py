"""Import the UNMODIFIED reference (``/root/reference``) in the build container.  TEST INFRASTRUCTURE ONLY.

Used only by ``oracle/make_goldens.py`` and ``tests/test_reference_live.py`` (skipped when
``/root/reference`` is absent, e.g. on the GPU box).  Nothing in the product, ``-m gpu`` tests,
``smoke()`` or ``bench.py`` reads ``/root/reference`` at run time.

Recipe: stub the python packages the reference imports but this image
lacks (omegaconf, hydra, importlib_resources, flashlight-backed torchaudio.models.decoder),
put ``streaming_decoder`` on sys.path, then build the model through the reference's own classes
(``StreamingAcousticEncoder`` encoder.py:73-147, ``CTCDecoder`` decoder.py:60-70) and attach them
to a ``LightningASR`` (recognition.py:136-217) created without its checkpoint loader."""
from __future__ import annotations

import importlib.resources
import os
import sys
import types
from typing import Dict

REFERENCE_ROOT = os.environ.get("ASR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "streaming_decoder", "lightspeech"))


def _install_stubs() -> None:
    if "omegaconf" not in sys.modules:
        om = types.ModuleType("omegaconf")
        om.DictConfig = dict
        om.OmegaConf = object
        sys.modules["omegaconf"] = om
    if "hydra" not in sys.modules:
        hy = types.ModuleType("hydra")
        hu = types.ModuleType("hydra.utils")
        hu.instantiate = lambda *a, **k: None
        hy.utils = hu
        sys.modules["hydra"] = hy
        sys.modules["hydra.utils"] = hu
    if "importlib_resources" not in sys.modules:
        ir = types.ModuleType("importlib_resources")
        ir.files = importlib.resources.files
        sys.modules["importlib_resources"] = ir
    import torchaudio  # noqa: F401
    dec = types.ModuleType("torchaudio.models.decoder")
    dec.ctc_decoder = None
    dec.CTCHypothesis = object
    sys.modules["torchaudio.models.decoder"] = dec


def load_reference():
    """Returns the reference's ``lightspeech.models.recognition`` module."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    p = os.path.join(REFERENCE_ROOT, "streaming_decoder")
    if p not in sys.path:
        sys.path.insert(0, p)
    from lightspeech.models import recognition as R
    return R


def build_reference_model(weights: Dict[str, "np.ndarray"], geo=None):
    """LightningASR instance (CPU) holding ``weights`` (oracle naming: 'encoder.*' / 'decoder.*')."""
    import torch
    R = load_reference()
    from lightspeech.modules.encoder import StreamingAcousticEncoder
    from lightspeech.modules.decoder import CTCDecoder
    from oracle.lightspeech_oracle import CANONICAL
    geo = geo or CANONICAL
    enc = StreamingAcousticEncoder(
        input_dim=geo.n_mels, d_model=geo.d_model, segment_length=geo.segment_size,
        left_context_length=geo.left_context * geo.stride, right_context_length=geo.context_size,
        ffn_dim=geo.ffn_dim, num_layers=geo.n_layers, subsampling_factor=geo.stride,
        num_heads=geo.n_heads, dropout=0.1, activation="gelu", max_memory_size=0, tanh_on_mem=True).eval()
    ctc = CTCDecoder(geo.d_model, geo.ctc_hidden, geo.vocab).eval()
    enc_sd = {k[len("encoder."):]: torch.from_numpy(v.copy()) for k, v in weights.items() if k.startswith("encoder.")}
    dec_sd = {k[len("decoder."):]: torch.from_numpy(v.copy()) for k, v in weights.items() if k.startswith("decoder.")}
    enc.load_state_dict(enc_sd, strict=True)
    ctc.load_state_dict(dec_sd, strict=True)
    m = R.LightningASR.__new__(R.LightningASR)
    m.device = "cpu"
    m.encoder = enc
    m.decoder = ctc
    m.blank = 0
    m.vocab = R.vocab
    return R, m
