import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _cuda_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def meta():
    with open(os.path.join(GOLD, "meta.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLD, f"{name}.npz")))
    return load


@pytest.fixture(scope="session")
def oracle_weights(meta):
    from oracle import lightspeech_oracle as O
    return O.make_weights(meta["weight_seed"])


@pytest.fixture(scope="session")
def packed_weights(oracle_weights):
    from asr_streaming_b200 import pack_weights
    return pack_weights(oracle_weights)
