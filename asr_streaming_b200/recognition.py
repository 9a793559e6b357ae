"""Drop-in mirror of the reference's model facade (lightspeech/models/recognition.py):

    model = LightningASR(...)                      # recognition.py:136-159
    state = model.init_state()                     # recognition.py:207-217
    emission, lengths, states = model.stream(speeches, sample_rate, states)     # recognition.py:191-204
    text, last_blank = greedy_search(stream.emission)                            # recognition.py:33-57

so that stream.py / streaming_server.py (:324-326, :420-435, :478, :530) run unmodified.  Everything numeric runs
on the B200 through the C ABI; this file only maps the reference's value-style ``state`` lists onto device-resident
session slots and formats text.

State mapping.  The reference hands the *same* ``state_init`` object to every new stream and re-assigns it at every
endpoint (streaming_server.py:326, :530).  Here ``init_state()`` returns a fresh-marker ``SessionState``; ``stream()``
opens a slot the first time it sees the marker and returns a bound ``SessionState``; when the caller drops a bound
state (re-assigning ``state_init``) its slot is released.  Unlike the reference, a bound state is consumed by
``stream()`` (the K/V ring is updated in place), which is how the server uses it.

greedy_search.  ``stream()`` returns the emission as an ``Emission`` tensor (a torch.Tensor subclass) tagged with its
session; the tag survives ``emission[0]`` and ``torch.cat((stream.emission, emission[0]))``, so
``greedy_search(stream.emission)`` returns the tokens the device already decoded incrementally (identical to
re-scanning the accumulated emission, tests/test_parity_gpu.py) instead of re-running argmax on the host.
"""
from __future__ import annotations

import re
import weakref
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .config import ModelConfig
from .engine import Engine, FRAMERATE, StepResult
from .weights import weights_from_checkpoint

SILENCE = "|"

_vocab: Optional[List[str]] = None


def set_vocab(vocab: Sequence[str]) -> None:
    """Install the id -> sub-syllable table (reference: build_vocab(), lightspeech/datas/text.py:27-30, 804 entries)."""
    global _vocab
    _vocab = list(vocab)


def get_vocab(n: int = 804) -> List[str]:
    global _vocab
    if _vocab is None:
        try:                                        # running inside the reference tree: use its corpus/vocab.txt
            from lightspeech.datas.text import build_vocab  # type: ignore
            _vocab = list(build_vocab())
        except Exception:
            _vocab = ["-", "|"] + [f"<{i}>" for i in range(2, n)]
    return _vocab


def ids_to_text(ids: Sequence[int], vocab: Optional[Sequence[str]] = None) -> str:
    """recognition.py:47-52: join sub-syllables, strip '<<' '>>' '-', '|' -> space, collapse whitespace."""
    v = vocab if vocab is not None else get_vocab()
    text = "".join(v[int(i)] for i in ids if int(i) != 0)
    text = text.replace("<<", "").replace(">>", "")
    text = text.replace("-", "").replace("|", " ")
    return re.sub(r"\s+", " ", text).strip()


def silent_ids(vocab: Optional[Sequence[str]] = None) -> List[int]:
    """Ids whose vocabulary string renders to "" under recognition.py:47-52 ('-', '|', '<<', '>>' in the reference's
    corpus/vocab.txt: ids 0, 1, 792, 793).  A segment whose tokens are all silent has empty text: ``Stream.update_stream`` then takes
    the ``else`` branch (stream.py:121-125).  Every other entry must keep a character that the stripping cannot remove (otherwise
    emptiness would depend on neighbouring tokens and could not be decided per id): checked here."""
    v = list(vocab) if vocab is not None else get_vocab()
    out = []
    for i, t in enumerate(v):
        if ids_to_text([i], v) == "" or i == 0:
            out.append(i)
        elif not any(c not in "<>-| \t\r\n" for c in t):
            raise ValueError(f"vocabulary entry {i} ({t!r}) consists of strippable characters only: whether a segment's text is empty "
                             "would depend on token order; this table is not supported by the per-id silent mask")
    return out


class SessionState:
    """Opaque replacement of the reference's 20x4 state tensor list."""

    def __init__(self, engine: Optional[Engine] = None, slot: Optional[int] = None):
        self.engine, self.slot = engine, slot
        self.tokens: List[int] = []          # collapsed, blank-free ids of the current utterance segment
        self.n_frames = 0
        self.blank_frames = 0
        self.has_token = False
        self.has_text = False
        if engine is not None and slot is not None:
            self._fin = weakref.finalize(self, _release, weakref.ref(engine), slot)

    @property
    def fresh(self) -> bool:
        return self.slot is None

    def last_blank(self) -> float:
        if self.has_token:
            return float(np.float32(self.blank_frames) * np.float32(FRAMERATE))
        return FRAMERATE * self.n_frames


def _release(engine_ref, slot):
    e = engine_ref()
    if e is not None and getattr(e, "_h", None):
        try:
            e.close_session(slot)
        except Exception:
            pass


class Emission(torch.Tensor):
    """Log-probs [.., 804] on the CPU, tagged with the session(s) that produced them."""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        out = super().__torch_function__(func, types, args, kwargs)
        if not isinstance(out, torch.Tensor):
            return out
        if func is torch.Tensor.__getitem__ and isinstance(args[1], int):
            states = getattr(args[0], "_asr_states", None)
            if states is not None:
                out._asr_state = states[args[1]]
        elif func is torch.cat:
            seq = args[0] if args else kwargs.get("tensors", ())
            tag = None
            for t in seq:
                tag = getattr(t, "_asr_state", tag)
            if tag is not None:
                out = out.as_subclass(Emission)
                out._asr_state = tag
        return out


def greedy_search(emission) -> Tuple[str, float]:
    """recognition.py:33-57.  Accepts what the server passes (the accumulated ``stream.emission``) or a SessionState."""
    st = emission if isinstance(emission, SessionState) else getattr(emission, "_asr_state", None)
    if st is None:
        if isinstance(emission, torch.Tensor) and emission.numel() == 0:
            return "", 0.0
        raise TypeError("greedy_search needs the Emission returned by LightningASR.stream (or its SessionState): the argmax / "
                        "collapse runs on the GPU with the chunk; a plain tensor carries no session to look the result up")
    if isinstance(emission, torch.Tensor) and emission.dim() == 2 and emission.size(0) != st.n_frames:
        raise ValueError(f"emission has {emission.size(0)} frames but its session decoded {st.n_frames}: the accumulated emission "
                         "must be reset together with the state (streaming_server.py:514-515, :530)")
    return ids_to_text(st.tokens), st.last_blank()


class LightningASR:
    """Same constructor positional arguments and methods as the reference class (recognition.py:136-217)."""

    def __init__(self, filepath: Optional[str] = None, model_dir: Optional[str] = None, device: Optional[str] = "cuda:0", *,
                 weights: Optional[np.ndarray] = None, cfg: ModelConfig = ModelConfig(), vocab: Optional[Sequence[str]] = None):
        import os
        self.blank = 0
        self.silence = SILENCE
        if vocab is not None:
            set_vocab(vocab)
        self.vocab = get_vocab(cfg.vocab)
        dev = str(device or "cuda:0")
        if not dev.startswith("cuda"):
            raise RuntimeError(f"device={device!r}: the B200 path has no CPU implementation; pass 'cuda:N'")
        self.device = dev
        index = int(dev.split(":")[1]) if ":" in dev else 0
        if weights is None:
            if filepath is None:
                raise ValueError("need a checkpoint path (filepath, model_dir) or a packed `weights` blob")
            weights = weights_from_checkpoint(os.path.join(model_dir or "", filepath), cfg)
        self.cfg = cfg
        self.engine = Engine(cfg, weights, index)
        self.engine.set_silent_ids(silent_ids(self.vocab))

    def init_state(self) -> SessionState:
        """recognition.py:207-217."""
        return SessionState()

    @torch.inference_mode()
    def stream(self, speeches: List[torch.Tensor], sample_rate: int, states: List[SessionState]):
        """recognition.py:191-204.  speeches: list of [1, chunk_length] float tensors in [-1, 1); states: one per stream."""
        if sample_rate != self.cfg.sample_rate:
            raise ValueError(f"sample_rate {sample_rate} != {self.cfg.sample_rate}")
        n = len(speeches)
        if n != len(states):
            raise ValueError("len(speeches) != len(states)")
        pcm = torch.cat([s.reshape(1, -1) for s in speeches]).to(torch.float32).contiguous().numpy()
        out_states: List[SessionState] = []
        for st in states:
            if st.fresh:
                st = SessionState(self.engine, self.engine.open_session())
            out_states.append(st)
        res: StepResult = self.engine.step([s.slot for s in out_states], pcm, want_logprobs=True)
        for i, st in enumerate(out_states):
            st.tokens.extend(int(t) for t in res.new_tokens[i])
            st.n_frames += self.cfg.seg_rows
            st.blank_frames = int(res.blank_frames[i])
            st.has_token = bool(res.has_token[i])
            st.has_text = bool(res.has_text[i])
        em = torch.from_numpy(res.logprobs).as_subclass(Emission)
        em._asr_states = out_states
        lengths = torch.full((n,), self.cfg.seg_rows, dtype=torch.long)
        return em, lengths, out_states
