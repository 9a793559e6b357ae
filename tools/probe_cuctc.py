"""Probe of torchaudio's cuda_ctc_decoder on this box (which argument combinations run): python tools/probe_cuctc.py B T V thr"""
import sys
import torch
from torchaudio.models.decoder import cuda_ctc_decoder

B, T, V, thr = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
torch.manual_seed(0)
lp = torch.log_softmax(3.0 * torch.randn(B, T, V, device="cuda"), dim=-1).contiguous()
dec = cuda_ctc_decoder([str(i) for i in range(V)], nbest=1, beam_size=10, blank_skip_threshold=thr)
h = dec(lp, torch.full((B,), T, dtype=torch.int32, device="cuda"))
torch.cuda.synchronize()
print("ok", B, T, V, thr, h[0][0].tokens[:8], float(h[0][0].score))
