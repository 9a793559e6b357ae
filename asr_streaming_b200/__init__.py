"""B200-native per-chunk compute path of the lightspeech streaming decoder.

Public surface (mirrors the reference's lightspeech.models.recognition, file:line in each docstring):
    LightningASR.init_state / .stream, greedy_search           -- drop-in for stream.py / streaming_server.py
    Engine                                                     -- batched, slot-based front of the C ABI
    SessionScheduler, GpuRouter                                -- ragged session batching, per-GPU partitioning
The CUDA library (libasr_b200.so, sm_100a) is mandatory: there is no CPU fallback.
"""
from .config import AudioConfig, ModelConfig, PRECISION_FAST, PRECISION_EXACT  # noqa: F401
from ._lib import AsrLibraryError, load_library  # noqa: F401
from .engine import Engine, StepResult  # noqa: F401
from .weights import pack_weights, weights_from_checkpoint, random_weights  # noqa: F401
from .recognition import LightningASR, greedy_search, ids_to_text, set_vocab, SessionState  # noqa: F401
from .scheduler import SessionScheduler, StreamSession, GpuRouter  # noqa: F401
from .endpoint import EndpointRules, OnlineEndpointRule, detect_endpointing, load_endpointing_rule  # noqa: F401
from .results import DecodedResult, create_hypotheses, interim_message, final_message, tick_messages  # noqa: F401

__all__ = ["AudioConfig", "ModelConfig", "Engine", "StepResult", "LightningASR", "greedy_search", "SessionScheduler",
           "StreamSession", "GpuRouter", "pack_weights", "weights_from_checkpoint", "random_weights", "load_library",
           "AsrLibraryError", "PRECISION_FAST", "PRECISION_EXACT"]
