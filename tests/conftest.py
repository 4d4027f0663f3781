import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    import json
    import numpy as np
    import torch
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    t = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta_json"}
    return t, meta


_PARAM_CACHE = {}


def oracle_params(knobs):
    """random_params is deterministic in its knobs; cache (full-size tables take seconds)."""
    from oracle import pipeline as pl
    key = tuple(sorted(knobs.items()))
    if key not in _PARAM_CACHE:
        if len(_PARAM_CACHE) > 2:
            _PARAM_CACHE.clear()
        _PARAM_CACHE[key] = pl.random_params(**knobs)
    return _PARAM_CACHE[key]
