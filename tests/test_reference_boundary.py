"""The drop-in boundary proven against the reference's OWN registry: the INTEGRATION.md snippet
is executed verbatim in a subprocess that has /root/reference/step_recognition on sys.path -- register
``prego_b200.MROAD`` into the reference's ``META_ARCHITECTURES`` (model/model_builder.py:5-9), build it through the
reference's ``build_model(cfg, device)``, strict-load a ``state_dict`` saved from the reference's ``MROAD``
(main.py:48,107), and check key order / shapes / seeded init.  Module construction needs no GPU; the forward is checked
to fail loudly on CPU tensors (no fallback).  Skips where the reference checkout is not mounted (the GPU box)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference/step_recognition"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, io
import torch
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
import yaml
# ---- what a PREGO maintainer adds (INTEGRATION.md section 2), verbatim
from model.model_builder import META_ARCHITECTURES, build_model     # the reference's own Registry / builder
from prego_b200.model import MROAD as MROAD_B200
META_ARCHITECTURES.register("MiniROAD_B200", MROAD_B200)
# ----
assert set(META_ARCHITECTURES) >= {"MiniROAD", "MiniROAD_B200"}
for cfg_file, K in (("configs/miniroad_assembly101-O.yaml", 86), ("configs/miniroad_epic-tent-O.yaml", 12)):
    cfg = yaml.load(open(REF + "/" + cfg_file), Loader=yaml.FullLoader)
    cfg.update(no_rgb=False, no_flow=False, eval=None)               # argparse keys merged in main.py:28-30
    torch.manual_seed(20)
    ref = build_model(dict(cfg), "cpu")                              # the reference's MROAD (rnn.py:18-71)
    torch.manual_seed(20)
    mine = build_model(dict(cfg, model="MiniROAD_B200"), "cpu")      # ours, through the reference's builder
    assert type(mine).__module__ == "prego_b200.model" and cfg["num_classes"] == K
    sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
    assert list(sd_ref) == list(sd_mine), (list(sd_ref), list(sd_mine))
    for k in sd_ref:
        assert sd_ref[k].shape == sd_mine[k].shape and sd_ref[k].dtype == sd_mine[k].dtype, k
        assert torch.equal(sd_ref[k], sd_mine[k]), f"seeded init differs: {k}"
    # a checkpoint written by the reference (main.py:107: torch.save(model.state_dict(), ...)) loads strictly (main.py:48)
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(0.25)
    buf = io.BytesIO()
    torch.save(ref.state_dict(), buf)
    buf.seek(0)
    res = mine.load_state_dict(torch.load(buf), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert all(torch.equal(mine.state_dict()[k], ref.state_dict()[k]) for k in sd_ref)
    # and the other way round: our state_dict loads strictly into the reference module
    assert not ref.load_state_dict(mine.state_dict(), strict=True).missing_keys
    assert sum(p.numel() for p in mine.parameters()) == sum(p.numel() for p in ref.parameters())
    assert isinstance(mine.h0, torch.Tensor) and "h0" not in sd_mine and tuple(mine.h0.shape) == (1, 1, cfg["hidden_dim"])
    mine.eval()
    try:
        mine(torch.zeros(1, 4, 2048), torch.zeros(1, 4, 2048))
    except RuntimeError as e:
        assert "CUDA" in str(e)                                       # no CPU fallback
    else:
        raise AssertionError("CPU tensors must be rejected")
# duplicate registration asserts, exactly like the reference's registry (utils/registry.py:1-3)
try:
    META_ARCHITECTURES.register("MiniROAD_B200", MROAD_B200)
except AssertionError:
    pass
else:
    raise AssertionError("duplicate registration must assert")
print("BOUNDARY_OK")
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only mounted in the build container")
def test_integration_snippet_against_the_reference_registry(tmp_path):
    script = tmp_path / "boundary.py"
    script.write_text(f"REF = {REF!r}\nROOT = {ROOT!r}\n" + SCRIPT)
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "BOUNDARY_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
