"""Interim / final result messages in the reference's wire format.

Mirrors ``DecodedResult`` (streaming_decoder/utils.py:24-42), ``create_hypotheses`` (utils.py:142-151) and the message
formation of ``handle_connection_impl`` (streaming_server.py:470-546) + ``_update_decoded_result`` (:588-598): an interim
message after every decoded chunk whose greedy text is non-empty, a final message when an endpoint fires.  The reference's
final pass re-decodes the segment with the third-party flashlight lexicon decoder + KenLM (absent here); this path
fills the final hypothesis from the on-GPU decode (prefix beam when enabled, else greedy) and leaves ``word_alignment`` empty,
which is exactly the shape ``get_hypotheses`` (utils.py:154-182) produces for an empty alignment.  Audio statistics and
speaker verification (streaming_server.py:535-537) are out of scope and keep their dataclass defaults.
"""
from __future__ import annotations

import dataclasses
import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .recognition import ids_to_text


@dataclass
class DecodedResult:
    """utils.py:24-42 — same field names, order and defaults (the JSON key order is part of the wire format)."""
    id: str = field(default_factory=str)
    status: int = field(default_factory=int)
    msg: int = field(default_factory=int)
    segment: int = field(default_factory=int)
    result: Dict[str, float] = field(default_factory=str)
    segment_start: float = field(default_factory=float)
    segment_length: float = field(default_factory=float)
    total_length: float = field(default_factory=float)
    message_type: int = field(default_factory=int)
    word_start: float = field(default_factory=float)
    word_end: float = field(default_factory=float)
    snr: float = 0.0
    vol_noise: float = 0.0
    vol_speech: float = 0.0
    is_speaker: bool = False


def create_hypotheses(transcript: str) -> dict:
    """utils.py:142-151."""
    return {"transcript": transcript, "transcript_normalized": transcript, "confidence": 0.0, "likelihood": 1.0, "word_alignment": []}


def final_hypotheses(transcript: str, confidence: float = 0) -> dict:
    """Shape of ``get_hypotheses`` (utils.py:154-182) for a decode without word alignment (no lexicon decoder here)."""
    return {"transcript": transcript, "transcript_normalized": transcript, "confidence": confidence, "word_alignment": []}


def to_json(r: DecodedResult) -> str:
    return json.dumps(dataclasses.asdict(r), ensure_ascii=False)          # streaming_server.py:489, :542


def interim_message(text: str) -> Optional[str]:
    """streaming_server.py:476-491: sent after a chunk when it is not final and the greedy text is not blank."""
    if text.strip() == "":
        return None
    r = DecodedResult()
    r.result = {"hypotheses": [create_hypotheses(text)], "final": False}
    return to_json(r)


def final_message(stream_id: str, segment: int, utt_length: float, total_length: float, transcript: str) -> Optional[str]:
    """streaming_server.py:505-545 + _update_decoded_result (:588-598) for language 'vi'; None when the transcript is blank
    (the reference sends nothing then, :533)."""
    r = DecodedResult()
    r.id = stream_id
    r.segment_length = utt_length
    r.segment = segment
    r.result = {"hypotheses": [final_hypotheses(transcript)], "final": True}
    r.total_length = total_length
    if transcript.strip() == "":
        return None
    return to_json(r)


def tick_messages(sched, res, vocab=None) -> Dict[int, List[str]]:
    """Messages of one scheduler tick, keyed by session id, in the order the reference would send them."""
    out: Dict[int, List[str]] = {}
    seg_s = sched.cfg.segment_length / sched.cfg.sample_rate
    for j, r in enumerate(res.rows):
        s = sched._by_row[int(r)]
        if res.final.size and res.final[j]:
            toks = res.final_tokens.get(s.id, [])
            bt = res.beam_row(j)
            if bt is not None:
                toks = [int(t) for t in bt]
            utt = float(res.final_utt_length.get(s.id, 0.0))
            total = float(sched.chunk_processed_total[r]) * seg_s
            m = final_message(str(s.id), int(sched.segment[r]) - 1, utt, total, ids_to_text(toks, vocab))
        else:
            m = interim_message(ids_to_text(sched.tok[r, :sched.ntok[r]], vocab))
        if m is not None:
            out.setdefault(s.id, []).append(m)
    return out
