"""ctypes binding of include/asr_b200.h.  Fails loudly when libasr_b200.so is missing: no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libasr_b200.so")

ABI_VERSION = 2


class AsrLibraryError(RuntimeError):
    pass


class AsrConfigC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "sample_rate", "hop", "n_fft", "win", "n_mels", "segment_size", "context_size", "bias", "stride",
        "d_model", "n_heads", "ffn_dim", "n_layers", "left_context", "ctc_hidden", "vocab", "precision", "max_sessions",
        "max_batch")]


class AsrStepOutC(C.Structure):
    _fields_ = [("argmax_ids", C.c_void_p), ("new_tokens", C.c_void_p), ("n_new", C.c_void_p), ("blank_frames", C.c_void_p),
                ("has_token", C.c_void_p), ("logprobs", C.c_void_p), ("beam_tokens", C.c_void_p), ("beam_len", C.c_void_p),
                ("beam_score", C.c_void_p), ("has_text", C.c_void_p), ("flags", C.c_void_p)]


class AsrSchedConfigC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("capacity", "chunk_length", "segment_length", "buffer_length", "seg_rows", "sample_rate", "max_batch",
                                         "backlog_chunks", "max_tokens", "device_gather")] + [("relative_cost", C.c_double)]


class AsrSchedArraysC(C.Structure):
    _fields_ = [("audio", C.c_void_p), ("audio_row_samples", C.c_int64), ("rd", C.c_void_p), ("wr", C.c_void_p), ("active", C.c_void_p),
                ("inflight", C.c_void_p), ("slot", C.c_void_p), ("tok", C.c_void_p), ("ntok", C.c_void_p), ("n_frames", C.c_void_p),
                ("chunk_processed", C.c_void_p), ("chunk_processed_total", C.c_void_p), ("trailing", C.c_void_p), ("contain_token", C.c_void_p),
                ("segment", C.c_void_p), ("last_served", C.c_void_p), ("relative_cost", C.c_void_p), ("overflow", C.c_void_p)]


class AsrSchedPlanC(C.Structure):
    _fields_ = [("tick", C.c_int32), ("n", C.c_int32), ("rows", C.c_void_p), ("slots", C.c_void_p), ("offsets", C.c_void_p),
                ("n_skipped", C.c_int32), ("skipped", C.c_void_p)]


class AsrSchedResultC(C.Structure):
    _fields_ = [("n", C.c_int32), ("rows", C.c_void_p), ("n_new", C.c_void_p), ("new_tokens", C.c_void_p), ("final_flags", C.c_void_p),
                ("final_rule", C.c_void_p), ("overflow", C.c_void_p), ("n_skipped", C.c_int32), ("skipped", C.c_void_p), ("n_final", C.c_int32),
                ("final_rows", C.c_void_p), ("final_rule_of", C.c_void_p), ("final_ntok", C.c_void_p), ("final_tok_off", C.c_void_p),
                ("final_utt", C.c_void_p), ("final_tok", C.c_void_p), ("argmax_ids", C.c_void_p), ("blank_frames", C.c_void_p),
                ("has_token", C.c_void_p), ("has_text", C.c_void_p), ("flags", C.c_void_p), ("beam_tokens", C.c_void_p), ("beam_len", C.c_void_p),
                ("beam_score", C.c_void_p), ("logprobs", C.c_void_p)]


class AsrStatsC(C.Structure):
    _fields_ = [("steps", C.c_uint64), ("stream_chunks", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("step_ms_p50", C.c_double), ("step_ms_p99", C.c_double), ("step_ms_max", C.c_double), ("step_ms_mean", C.c_double)]


# name -> (restype, argtypes); must list every symbol include/asr_b200.h declares (tests check this).
SIGNATURES = {
    "asr_last_error": (C.c_char_p, []),
    "asr_abi_version": (C.c_int, []),
    "asr_default_config": (C.c_int, [C.POINTER(AsrConfigC), C.c_int]),
    "asr_weights_count": (C.c_int, [C.POINTER(AsrConfigC), C.POINTER(C.c_uint64)]),
    "asr_chunk_geometry": (C.c_int, [C.POINTER(AsrConfigC), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "asr_engine_create": (C.c_int, [C.POINTER(AsrConfigC), C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]),
    "asr_engine_destroy": (C.c_int, [C.c_void_p]),
    "asr_session_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "asr_session_reset": (C.c_int, [C.c_void_p, C.c_int32]),
    "asr_session_close": (C.c_int, [C.c_void_p, C.c_int32]),
    "asr_session_reset_many": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "asr_set_silent_ids": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "asr_gather_pcm": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "asr_pcm_peaks": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "asr_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(AsrStepOutC)]),
    "asr_submit": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "asr_collect": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(AsrStepOutC)]),
    "asr_submit_rings": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "asr_wait_inputs": (C.c_int, [C.c_void_p]),
    "asr_host_alloc": (C.c_void_p, [C.c_uint64]),
    "asr_host_free": (C.c_int, [C.c_void_p]),
    "asr_stage": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "asr_run_staged": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "asr_fetch": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(AsrStepOutC)]),
    "asr_sync": (C.c_int, [C.c_void_p]),
    "asr_stream_handle": (C.c_void_p, [C.c_void_p]),
    "asr_pinned_pcm": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "asr_fbank": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "asr_fbank_staged": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "asr_stage_raw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "asr_set_beam": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "asr_get_stats": (C.c_int, [C.c_void_p, C.POINTER(AsrStatsC)]),
    "asr_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "asr_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "asr_debug_step_partial": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "asr_debug_decode_logits": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(AsrStepOutC)]),
    "asr_debug_read": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64]),
    "asr_debug_read_state": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]),
    "asr_debug_gemm": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int]),
    "asr_debug_gemm_operand": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "asr_convmod_weights_count": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]),
    "asr_convmod_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_int32, C.POINTER(C.c_void_p)]),
    "asr_convmod_destroy": (C.c_int, [C.c_void_p]),
    "asr_convmod_reset": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "asr_convmod_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "asr_debug_gemm_ln": (C.c_int, [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 8 + [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.POINTER(C.c_float), C.c_int32, C.c_int]),
    "asr_pipeline_gpu_time": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int32]),
    "asr_sched_create": (C.c_int, [C.POINTER(AsrSchedConfigC), C.c_void_p, C.POINTER(C.c_void_p)]),
    "asr_sched_destroy": (C.c_int, [C.c_void_p]),
    "asr_sched_arrays": (C.c_int, [C.c_void_p, C.POINTER(AsrSchedArraysC)]),
    "asr_sched_set_rules": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "asr_sched_open": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "asr_sched_close": (C.c_int, [C.c_void_p, C.c_int32]),
    "asr_sched_reset_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "asr_sched_accept": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]),
    "asr_sched_accept_block": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]),
    "asr_sched_ready": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]),
    "asr_sched_plan": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(AsrSchedPlanC)]),
    "asr_sched_commit": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(AsrSchedResultC)]),
    "asr_sched_update": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(AsrStepOutC)]),
    "asr_sched_endpoints": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(AsrSchedResultC)]),
    "asr_sched_abort": (C.c_int, [C.c_void_p, C.c_int32]),
    "asr_sched_submit": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(AsrSchedResultC), C.POINTER(C.c_int32)]),
    "asr_sched_collect": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(AsrSchedResultC)]),
    "asr_sched_prestage": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "asr_debug_gemm_time": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.c_int]),
}

_lib = None
_lock = threading.Lock()


def load_library(path: str = None) -> C.CDLL:
    """dlopen libasr_b200.so and type every entry point.  Raises AsrLibraryError if it is missing or incomplete."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or os.environ.get("ASR_B200_LIB", LIB_PATH)
        if not os.path.exists(p):
            raise AsrLibraryError(
                f"{p} not found: build it with `python -m asr_streaming_b200.build` (nvcc, sm_100a). "
                "This package has no CPU fallback.")
        try:
            lib = C.CDLL(p)
        except OSError as e:
            raise AsrLibraryError(f"cannot load {p}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise AsrLibraryError(f"{p} does not export {name}; rebuild the library") from e
            fn.restype = res
            fn.argtypes = args
        if lib.asr_abi_version() != ABI_VERSION:
            raise AsrLibraryError(f"ABI mismatch: library {lib.asr_abi_version()} vs binding {ABI_VERSION}")
        if path is None:
            _lib = lib
        return lib


def check(lib: C.CDLL, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.asr_last_error()
        raise AsrLibraryError(f"{what} failed: {msg.decode('utf-8', 'replace') if msg else rc}")
