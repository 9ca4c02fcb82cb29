import gzip
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
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100a device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden_meta():
    return json.load(open(os.path.join(GOLD, "meta.json")))


def load_model_case(name):
    z = np.load(os.path.join(GOLD, f"model_{name}.npz"))
    return {k: z[k] for k in z.files}


def load_gz_json(name):
    with gzip.open(os.path.join(GOLD, name), "rb") as f:
        return json.loads(f.read().decode())


def case_inputs(meta, name, device="cpu"):
    """(cfg, rgb, flow) of a golden case, regenerated from seeds and verified against the stored hashes."""
    import hashlib

    import torch

    from prego_b200 import synthetic

    c = meta["cases"][name]
    cfg = dict(getattr(synthetic, c["cfg"]), **c["overrides"])
    rgb, flow = synthetic.feature_batch(c["stream_ids"], c["T"], "cpu", c["zero_flow"])
    sha = lambda t: hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()
    assert sha(rgb) == c["rgb_sha256"] and sha(flow) == c["flow_sha256"], "synthetic feature generator drifted"
    return cfg, rgb.to(device), flow.to(device)


def seeded_weights_checked(meta, name, device=None):
    """Model built under seed 20 whose ten tensors hash to what the reference produced."""
    import hashlib

    from prego_b200 import synthetic

    c = meta["cases"][name]
    cfg = dict(getattr(synthetic, c["cfg"]), **c["overrides"])
    m = synthetic.seeded_model(cfg, seed=meta["seed"])
    for k, v in m.state_dict().items():
        h = hashlib.sha256(v.contiguous().numpy().tobytes()).hexdigest()
        assert h == meta["weights_sha256"][c["weights"]][k], f"seeded weight {k} differs from the reference's"
    return m.to(device) if device is not None else m
