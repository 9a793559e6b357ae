import numpy as np

from oracle import lightspeech_oracle as O


def to_float(pcm_i16):
    return (pcm_i16.astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def chunks_i16(pcm_i16, geo=O.CANONICAL):
    """Chunk windows as int16 (what the websocket delivers), same framing as the caller (stream.py:23, :159)."""
    return O.chunk_windows(pcm_i16.astype(np.int16), geo)


def model_cfg(precision, low_latency=False, max_batch=64, max_sessions=128):
    from asr_streaming_b200 import ModelConfig
    return ModelConfig(segment_size=32 if low_latency else 64, precision=precision, max_batch=max_batch, max_sessions=max_sessions)


def margins(emission):
    s = np.sort(emission, axis=-1)
    return s[..., -1] - s[..., -2]


def report(line):
    """Appends a measured-error line to gpurun_out/parity_report.txt (kept as evidence under profiles/)."""
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_report.txt"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass
