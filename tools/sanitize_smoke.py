"""Smallest run that touches every kernel family (for compute-sanitizer --tool memcheck): 3 streams, 2 chunks, large-batch kernels forced."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASR_B200_FUSED_LN_MIN_STREAMS"] = "1"
os.environ["ASR_B200_ATTN_STREAM_MIN"] = "1"
from asr_streaming_b200 import Engine, ModelConfig, pack_weights, random_weights  # noqa: E402

cfg = ModelConfig(max_batch=4, max_sessions=4)
e = Engine(cfg, pack_weights(random_weights(1234, cfg), cfg), 0)
e.set_beam(10, 8)
sl = [e.open_session() for _ in range(3)]
pcm = np.random.default_rng(0).integers(-3000, 3000, size=(3, cfg.chunk_length)).astype(np.int16)
for _ in range(2):
    r = e.step(sl, pcm, want_logprobs=True)
e.reset_sessions(sl[:2])
r = e.step(sl, pcm)
print("ok", r.argmax_ids[0][:8], e.fbank(pcm[:, :10480], kind=1).shape)
e.close()
