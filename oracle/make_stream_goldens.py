"""Golden sequences of Stream.update_stream / endpoint_detected from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Drives the reference's own ``greedy_search`` (lightspeech/models/recognition.py:33-57), ``Stream.update_stream`` (stream.py:110-125)
and ``Stream.endpoint_detected`` (stream.py:127-163, with online_endpoint.detect_endpointing and the DEFAULT rule table of
config/asr-online.yaml:31-104) on scripted per-frame argmax ids, including segments whose only tokens are the markers '<<' / '>>'
(ids 792 / 793 of corpus/vocab.txt): they have an id > 1 but render to "", so ``if text:`` takes the else branch.
Writes tests/golden/stream_update.json.  Run in the build container: python -m oracle.make_stream_goldens"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

from oracle import ref_import

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_stream_class():
    R = ref_import.load_reference()
    for name in ("webrtcvad",):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    os.environ.setdefault("NORM_PORT", "0")
    sd = os.path.join(ref_import.REFERENCE_ROOT, "streaming_decoder")
    if sd not in sys.path:
        sys.path.insert(0, sd)
    import logging
    import tempfile
    cwd, tmp = os.getcwd(), tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "logs"))
    os.chdir(tmp)                                        # utils.py:90 opens logs/debug.log relative to the working directory at import
    try:
        import online_endpoint  # noqa: F401
        import stream as S
    finally:
        os.chdir(cwd)
    root = logging.getLogger()
    for h in list(root.handlers):
        root.removeHandler(h)
    root.setLevel(logging.WARNING)
    return R, S


CASES = {
    # per chunk: 16 argmax ids (0 blank '-', 1 silence '|', 792 '<<', 793 '>>', others sub-syllables)
    "markers_only_then_speech": [
        [0] * 16,
        [0, 0, 792, 792, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],          # only '<<': id > 1 exists, text ""
        [0, 0, 0, 0, 0, 793, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],            # '<<' '>>' : still ""
        [0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],              # silence token: text " " -> ""
        [0, 0, 0, 57, 57, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0],            # first real token
        [0] * 16, [0] * 16, [0] * 16,
    ],
    "speech_then_long_silence": [
        [5, 5, 0, 6, 0, 0, 7, 1, 0, 0, 0, 0, 0, 0, 0, 0],
        [0] * 16, [0] * 16, [0] * 16,
    ],
    "markers_between_utterances": [
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 9, 9, 10, 0],
        [0] * 16, [0] * 16,
        [792, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],            # after the endpoint: a marker-only segment
        [0] * 16,
        [0, 0, 0, 0, 0, 0, 0, 0, 11, 0, 0, 0, 0, 0, 0, 793],
        [0] * 16, [0] * 16,
    ],
}


def main():
    R, S = load_stream_class()
    from online_endpoint import load_endpointing_rule
    from asr_streaming_b200.endpoint import DEFAULT_RULES
    rules = load_endpointing_rule({k: dict(must_contain_nonsilence=r.must_contain_nonsilence, min_trailing_silence=r.min_trailing_silence,
                                          min_utterance_length=r.min_utterance_length, max_relative_cost=r.max_relative_cost)
                                   for k, r in DEFAULT_RULES.items()})
    S.compute_relative_cost = lambda *a, **k: 10.0                       # the ARPA LM is absent: the constant the product uses
    out = {"vocab_size": len(R.vocab), "silent_ids": [i for i, t in enumerate(R.vocab) if i == 0 or R.greedy_search(_onehot([i], len(R.vocab)))[0] == ""], "cases": {}}
    for name, chunks in CASES.items():
        st = S.Stream.__new__(S.Stream)
        # the attributes update_stream / endpoint_detected read (Stream.__init__ needs OmegaConf + webrtcvad, absent here)
        st.language, st.emission, st.chunk_processed, st.chunk_processed_total = "vi", torch.Tensor([]), 0, 0
        st.segment_size, st.bias, st.context_size, st.framerate = 64, 4, 16, 4
        st.transcript_internal, st.transcript, st.trailing_blank_duration, st.is_contain_token = "", "", 0, False
        st.segment_length, st.sample_rate, st.sw_model, st.id, st.segment = 10240, 16000, "GENERAL", "g", 0
        st.mapping_endpointing_rule, st.EndpointingRule = {"GENERAL": "DEFAULT"}, {"DEFAULT": rules}
        st.audio_stream, st.length_of_segment, st.segment_end = torch.zeros(13440 * 4), 13440 * 4, 0.0
        rec = []
        for ids in chunks:
            st.emission = torch.cat((st.emission, _onehot(ids, len(R.vocab))))          # streaming_server.py:428-431
            text, last_blank = R.greedy_search(st.emission)                                # :433
            st.update_stream(text, last_blank)                                             # :435
            before = dict(text=text, last_blank=float(last_blank), trailing=float(st.trailing_blank_duration), contain=bool(st.is_contain_token),
                          chunk_processed=int(st.chunk_processed))
            detected, utt = st.endpoint_detected(None, None)                               # :470
            if detected:
                st.emission = torch.Tensor([])                                             # :514-515
            rec.append(dict(ids=ids, **before, detected=bool(detected), utt=float(utt), trailing_after=float(st.trailing_blank_duration),
                            segment=int(st.segment)))
        out["cases"][name] = rec
    p = os.path.join(ROOT, "tests", "golden", "stream_update.json")
    with open(p, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", p, {k: sum(r["detected"] for r in v) for k, v in out["cases"].items()})


def _onehot(ids, V):
    e = torch.full((len(ids), V), -20.0)
    for i, t in enumerate(ids):
        e[i, t] = -0.01
    return e


if __name__ == "__main__":
    main()
