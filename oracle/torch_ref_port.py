"""Torch/torchaudio port of the reference's CPU path.  TEST INFRASTRUCTURE ONLY (bench.py cpu_baseline / --impl reference).

``/root/reference`` cannot travel to the GPU box, but the arithmetic of its hot path lives in third-party
torchaudio (``torchaudio.models.Emformer``, ``torchaudio.transforms.MelSpectrogram``), which the box has.  This
file re-assembles the reference's own glue around those *unmodified* library modules, line for line in behaviour:

    StreamingAcousticEncoder.infer   lightspeech/modules/encoder.py:73-147
    CTCDecoder.forward               lightspeech/modules/decoder.py:60-70
    extract_filterbank               lightspeech/datas/audio.py:9-30      (MelSpectrogram is rebuilt EVERY call, as there)
    time_reduction                   lightspeech/utils/common.py:110-124
    pack_input / unpack_states       lightspeech/models/recognition.py:60-92
    LightningASR.stream / init_state lightspeech/models/recognition.py:191-217
    greedy_search                    lightspeech/models/recognition.py:33-57

so timing it is timing the reference's CPU implementation (kind = "port").  tests/test_reference_live.py checks it
bit-for-bit against the real reference where that is importable.
"""
from __future__ import annotations

import re
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
import torchaudio
from torchaudio.models import Emformer

from oracle.lightspeech_oracle import CANONICAL, Geometry

FRAMERATE = 0.04


class _Encoder(torch.nn.Module):
    def __init__(self, geo: Geometry):
        super().__init__()
        self.stride = geo.stride
        self.input_linear = torch.nn.Linear(geo.n_mels, geo.d_model // geo.stride, bias=False)
        self.encoder_layers = Emformer(
            input_dim=geo.d_model, num_heads=geo.n_heads, ffn_dim=geo.ffn_dim, num_layers=geo.n_layers,
            segment_length=geo.segment_size // geo.stride, dropout=0.1, activation="gelu",
            left_context_length=geo.left_context, right_context_length=geo.context_size // geo.stride,
            max_memory_size=0, weight_init_scale_strategy="depthwise", tanh_on_mem=True)

    def infer(self, xs, x_lens, states):
        xs = self.input_linear(xs)
        b, t, d = xs.shape
        n = t + (self.stride - t % self.stride) % self.stride
        xs = F.pad(xs, (0, 0, 0, n - t)).reshape(b, n // self.stride, d * self.stride).contiguous()
        x_lens = (torch.div(x_lens - 1, self.stride, rounding_mode="trunc") + 1).type(torch.long)
        return self.encoder_layers.infer(xs, x_lens, states)


class _Ctc(torch.nn.Module):
    def __init__(self, geo: Geometry):
        super().__init__()
        self.linear1 = torch.nn.Linear(geo.d_model, geo.ctc_hidden)
        self.linear2 = torch.nn.Linear(geo.ctc_hidden, geo.vocab)

    def forward(self, enc_outs):
        return self.linear2(F.silu(self.linear1(enc_outs))).log_softmax(2)


class TorchRefPort:
    def __init__(self, weights: Dict[str, np.ndarray], geo: Geometry = CANONICAL, device: str = "cpu"):
        self.geo, self.device = geo, device
        self.encoder, self.decoder = _Encoder(geo).eval(), _Ctc(geo).eval()
        self.encoder.load_state_dict({k[len("encoder."):]: torch.from_numpy(v.copy()) for k, v in weights.items() if k.startswith("encoder.")})
        self.decoder.load_state_dict({k[len("decoder."):]: torch.from_numpy(v.copy()) for k, v in weights.items() if k.startswith("decoder.")})

    def init_state(self):
        g = self.geo
        return [[torch.zeros(0, 1, g.d_model), torch.zeros(g.left_context, 1, g.d_model), torch.zeros(g.left_context, 1, g.d_model),
                 torch.zeros(1, 1, dtype=torch.int32)] for _ in range(g.n_layers)]

    def _extract_filterbank(self, waveform, sample_rate):
        tr = torchaudio.transforms.MelSpectrogram(sample_rate=sample_rate, n_fft=int(0.05 * sample_rate), win_length=int(0.025 * sample_rate),
                                                  hop_length=int(0.01 * sample_rate), n_mels=self.geo.n_mels, center=False)
        fb = torch.transpose(tr(waveform).clamp(1e-5).log(), 2, 1)
        return fb, torch.tensor([x.size(0) for x in fb])

    @torch.inference_mode()
    def stream(self, speeches: List[torch.Tensor], sample_rate: int, states: List):
        L, B = self.geo.n_layers, len(states)
        packed = [[torch.cat([states[i][l][j] for i in range(B)], dim=1) for j in range(4)] for l in range(L)]
        xs, x_lens = self._extract_filterbank(torch.cat(speeches), sample_rate)
        enc, enc_lens, packed = self.encoder.infer(xs, x_lens, packed)
        out_states = [[[packed[l][0][:, i, :].unsqueeze(1), packed[l][1][:, i, :].unsqueeze(1), packed[l][2][:, i, :].unsqueeze(1),
                        packed[l][3][:, i].unsqueeze(1)] for l in range(L)] for i in range(B)]
        return self.decoder(enc).cpu(), enc_lens.cpu(), out_states


def greedy_search(emission: torch.Tensor, vocab: Sequence[str]) -> Tuple[str, float]:
    indices = torch.argmax(emission, dim=1)
    last_blank = FRAMERATE * len(emission)
    tokens_idx = (indices > 1).nonzero(as_tuple=True)[0]
    if len(tokens_idx):
        last_blank = ((len(indices) - 1 - tokens_idx[-1]) * FRAMERATE).item()
    indices = torch.unique_consecutive(indices, dim=0)
    indices = torch.masked_select(indices, indices != 0)
    text = "".join(vocab[idx] for idx in indices if idx != 0)
    text = text.replace("<<", "").replace(">>", "").replace("-", "").replace("|", " ")
    text = re.sub(r"\s+", " ", text).strip()
    _ = (torch.amax(emission, dim=1).sum() / max(indices.size(0), 1)).exp().item()      # computed and discarded, :54-55
    return text, last_blank
