"""The ATen-based CPU port used as bench.py's cpu_baseline reproduces the reference's golden vectors."""
import numpy as np
import pytest

from conftest import case_inputs, load_model_case, seeded_weights_checked
from oracle.miniroad_torch_cpu import CpuMiniROAD


@pytest.mark.parametrize("name", ["epic_b1_t300", "asm_b1_t64_zeroflow", "epic_b1_t96_rgbonly"])
def test_cpu_port_matches_reference(golden_meta, name):
    gold = load_model_case(name)
    cfg, rgb, flow = case_inputs(golden_meta, name)
    sd = seeded_weights_checked(golden_meta, name).state_dict()
    port = CpuMiniROAD(sd, use_rgb=not cfg["no_rgb"], use_flow=not cfg["no_flow"])
    probs = port.forward(rgb, flow).numpy()
    assert np.abs(probs - gold["probs"]).max() <= 1e-6  # same ATen kernels: equal up to thread-count effects
    assert np.array_equal(port.labels(rgb, flow), gold["probs"].argmax(-1))
