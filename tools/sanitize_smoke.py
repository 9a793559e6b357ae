"""Smallest run that touches every kernel family: 3 streams, 2 chunks, large-batch kernels forced; the decode stage on tied logits; pre-staged
scheduler ticks with the device gather and the speculative fbank.  Written for a memory checker (closed on this GPU pool: run plain there)."""
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
# the decode stage alone on tied logits (plain-selection path of the candidate extraction) and the pre-staged scheduler tick
# (device gather out of pinned rings, speculative fbank, compaction of the rows that run)
z = np.random.default_rng(1).integers(0, 3, size=(3, cfg.seg_rows, cfg.vocab)).astype(np.float32)
r = e.debug_decode_logits(sl, z, want_logprobs=True)
print("decode ok", r.beam_tokens[0][:4])
e.close()
from asr_streaming_b200 import SessionScheduler  # noqa: E402
os.environ["ASR_B200_DEVICE_GATHER"] = "1"
e = Engine(cfg, pack_weights(random_weights(1234, cfg), cfg), 0)
sch = SessionScheduler(e, capacity=4)
ss = [sch.open() for _ in range(3)]
audio = np.random.default_rng(2).integers(-3000, 3000, size=(3, 4 * cfg.segment_length)).astype(np.int16)
prev = None
for k in range(4):
    for i, s_ in enumerate(ss):
        s_.accept_waveform(audio[i, k * cfg.segment_length:(k + 1) * cfg.segment_length])
    sch.prestage()
    if prev is not None:
        sch.collect_tick(prev)
    p = sch.submit_tick()
    prev = p if p.rows.size else None
if prev is not None:
    sch.collect_tick(prev)
print("scheduler ok", [len(s_.tokens) for s_ in ss])
e.close()
