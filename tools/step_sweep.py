"""ms per step of the whole per-chunk path vs batch size (device-resident inputs, CUDA events on the engine stream):
   python tools/step_sweep.py [streams ...]        (environment switches such as ASR_B200_NO_FUSED_LN=1 apply)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asr_streaming_b200 import Engine, ModelConfig, pack_weights, random_weights  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [64, 256, 512, 1024, 2048, 4096]
cfg0 = ModelConfig()
blob = pack_weights(random_weights(1234, cfg0), cfg0)
rng = np.random.default_rng(0)
for n in sizes:
    cfg = ModelConfig(max_batch=n, max_sessions=n)
    eng = Engine(cfg, blob, 0)
    ext = torch.cuda.ExternalStream(eng.cuda_stream)
    slots = [eng.open_session() for _ in range(n)]
    eng.stage(slots, rng.integers(-3000, 3000, size=(n, cfg.chunk_length)).astype(np.int16))
    for _ in range(4):
        eng.run_staged(n)
    eng.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20 if n <= 1024 else 8
    a.record(ext)
    for _ in range(iters):
        eng.run_staged(n)
    b.record(ext)
    eng.sync()
    ms = a.elapsed_time(b) / iters
    print(f"streams {n:5d}: {ms:8.3f} ms/step  {n * 0.64 / ms * 1e3:10.0f} audio-s/s", flush=True)
    eng.close()
