"""Endpoint rules evaluated for all sessions of a tick at once.

Mirrors ``streaming_decoder/online_endpoint.py``: ``OnlineEndpointRule`` (:4-21), ``_rule_activated`` (:42-66) and
``detect_endpointing`` (:69-94, first activated rule in declaration order wins).  ``DEFAULT_RULES`` restates the
``Endpointing_rules: DEFAULT`` table of ``config/asr-online.yaml:31-104``.  The relative cost fed to the rules comes from
an ARPA language model in the reference (``utils.py:126-139``, file absent); callers pass it per session.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Mapping, Optional, Tuple

import numpy as np

INF = float("inf")


@dataclass(frozen=True)
class OnlineEndpointRule:
    """online_endpoint.py:4-21 (same field names)."""
    must_contain_nonsilence: bool
    min_trailing_silence: float
    min_utterance_length: float
    max_relative_cost: float


def _r(sil: float, utt: float, cost: float) -> OnlineEndpointRule:
    return OnlineEndpointRule(True, sil, utt, cost)


# asr-online.yaml:31-104, declaration order preserved (dict order is evaluation order, online_endpoint.py:89)
DEFAULT_RULES: Dict[str, OnlineEndpointRule] = {
    "rule1.1": _r(1.0, 0.0, INF), "rule1.2": _r(0.9, 0.0, 8), "rule1.3": _r(0.8, 0.0, 5), "rule1.4": _r(0.7, 0.0, 2),
    "rule2.1": _r(1.0, 10.0, INF), "rule2.2": _r(0.9, 10.0, 8), "rule2.3": _r(0.7, 10.0, 5), "rule2.4": _r(0.6, 10.0, 2),
    "rule3.1": _r(0.9, 20.0, INF), "rule3.2": _r(0.8, 20.0, 8), "rule3.3": _r(0.7, 20.0, 5), "rule3.4": _r(0.6, 20.0, 2),
    "rule4": _r(0.0, 40.0, INF),
}


def load_endpointing_rule(args: Mapping[str, Mapping]) -> Dict[str, OnlineEndpointRule]:
    """online_endpoint.py:24-39: ``{name: {field: value}}`` (the parsed YAML block) -> ``{name: rule}``."""
    return {name: OnlineEndpointRule(a["must_contain_nonsilence"], a["min_trailing_silence"], a["min_utterance_length"],
                                     a["max_relative_cost"]) for name, a in args.items()}


def detect_endpointing(rule: Mapping[str, OnlineEndpointRule], utterance_length: float, trailing_silence: float,
                       relative_cost: float) -> Tuple[bool, Optional[str], Optional[float]]:
    """Scalar form with the reference's signature and return triple (online_endpoint.py:69-94)."""
    fired, which = EndpointRules(rule).detect(np.array([utterance_length]), np.array([trailing_silence]), np.array([relative_cost]))
    if not fired[0]:
        return False, None, None
    name = list(rule)[int(which[0])]
    return True, name, trailing_silence - rule[name].min_trailing_silence


class EndpointRules:
    """A rule table as arrays: ``detect`` evaluates every rule for every session in one shot."""

    def __init__(self, rules: Optional[Mapping[str, OnlineEndpointRule]] = None):
        rules = DEFAULT_RULES if rules is None else rules
        self.names = list(rules)
        rs = [rules[k] for k in self.names]
        self.must = np.array([r.must_contain_nonsilence for r in rs], bool)
        self.min_sil = np.array([r.min_trailing_silence for r in rs], np.float64)
        self.min_utt = np.array([r.min_utterance_length for r in rs], np.float64)
        self.max_cost = np.array([r.max_relative_cost for r in rs], np.float64)

    def __len__(self) -> int:
        return len(self.names)

    def detect(self, utterance_length, trailing_silence, relative_cost):
        """[n] float arrays -> (fired [n] bool, which [n] int = index of the first activated rule, -1 if none)."""
        utt = np.asarray(utterance_length, np.float64)[:, None]
        sil = np.asarray(trailing_silence, np.float64)[:, None]
        cost = np.asarray(relative_cost, np.float64)[:, None]
        nonsil = utt > sil                                             # online_endpoint.py:59
        act = ((nonsil | ~self.must[None, :]) & (sil >= self.min_sil[None, :]) & (cost < self.max_cost[None, :])
               & (utt >= self.min_utt[None, :]))                       # :60-65
        if act.shape[1] == 0:
            return np.zeros(act.shape[0], bool), np.full(act.shape[0], -1, np.int64)
        fired = act.any(axis=1)
        which = np.where(fired, act.argmax(axis=1), -1)
        return fired, which
