"""CPU oracle for CTC prefix beam search.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference has no CTC prefix beam search (its final-pass decoder is the
flashlight-text lexicon/KenLM decoder behind torchaudio.models.decoder.ctc_decoder, recognition.py:220-300 —
third-party native code, absent here, no LM / lexicon files in the repo).  The north-star names "prefix-beam decode
(beam=10)"; this file restates the standard algorithm (Hannun et al. 2014, "First-pass large vocabulary continuous
speech recognition using bi-directional recurrent DNNs", Alg. 1, without LM) with the pruning and tie-break rules the
CUDA kernel implements, so the kernel can be checked token-exactly against it:

  * per frame, extension candidates are the ``cand_k`` best NON-blank tokens (ties: lower id first);
  * every beam entry always gets its blank continuation and its repeat-of-last-token continuation;
  * candidates are enumerated as [stay(0..B-1), ext(parent 0, cand 0..K-1), ext(parent 1, ...), ...]; an extension
    whose prefix equals an existing beam entry's prefix is merged into that entry; the next beam is the ``beam`` best
    by logaddexp(p_blank, p_nonblank), ties to the lower enumeration index;
  * prefixes longer than ``max_len`` are not extended;
  * state (beam entries) is carried across chunks; reset at an endpoint.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

NEG_INF = -math.inf


def logaddexp(a: float, b: float) -> float:
    if a == NEG_INF:
        return b
    if b == NEG_INF:
        return a
    m, n = (a, b) if a > b else (b, a)
    return m + math.log1p(math.exp(n - m))


@dataclass
class BeamState:
    entries: List[Tuple[Tuple[int, ...], float, float]] = field(default_factory=lambda: [((), 0.0, NEG_INF)])


def top_candidates(logp: np.ndarray, k: int, blank: int = 0) -> List[int]:
    order = sorted((i for i in range(len(logp)) if i != blank), key=lambda i: (-float(logp[i]), i))
    return order[:k]


def beam_step(state: BeamState, logp_frames: np.ndarray, beam: int = 10, cand_k: int = 8, max_len: int = 256, blank: int = 0) -> BeamState:
    entries = state.entries
    for logp in logp_frames:
        cands = top_candidates(logp, cand_k, blank)
        B = len(entries)
        prefix_index: Dict[Tuple[int, ...], int] = {e[0]: j for j, e in enumerate(entries)}
        # stay candidates
        pb_new = [NEG_INF] * B
        pnb_new = [NEG_INF] * B
        for j, (pre, pb, pnb) in enumerate(entries):
            ptot = logaddexp(pb, pnb)
            pb_new[j] = ptot + float(logp[blank])
            if pre:
                pnb_new[j] = pnb + float(logp[pre[-1]])
        # extension candidates, merged into a stay entry when the extended prefix is already in the beam
        ext: List[Tuple[int, Tuple[int, ...], float]] = []          # (enumeration index, prefix, pnb)
        for i, (pre, pb, pnb) in enumerate(entries):
            ptot = logaddexp(pb, pnb)
            for kk, c in enumerate(cands):
                if len(pre) >= max_len:
                    continue
                val = (pb if (pre and c == pre[-1]) else ptot) + float(logp[c])
                new = pre + (c,)
                j = prefix_index.get(new)
                if j is not None:
                    pnb_new[j] = logaddexp(pnb_new[j], val)
                else:
                    ext.append((B + i * cand_k + kk, new, val))
        allc = [(j, entries[j][0], pb_new[j], pnb_new[j]) for j in range(B)] + [(idx, pre, NEG_INF, v) for idx, pre, v in ext]
        allc.sort(key=lambda t: (-logaddexp(t[2], t[3]), t[0]))
        entries = [(pre, pb, pnb) for _, pre, pb, pnb in allc[:beam] if logaddexp(pb, pnb) > NEG_INF]
    return BeamState(entries)


def best(state: BeamState) -> Tuple[List[int], float]:
    pre, pb, pnb = state.entries[0]
    return list(pre), logaddexp(pb, pnb)
