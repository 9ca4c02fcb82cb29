"""Host-side logic of bench.py that needs no GPU: the argument contract and the clock-sample windows."""
import importlib.util
import os
import sys

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_argument_contract(monkeypatch):
    b = _bench()
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = b.parse()
    assert (a.gpus, a.impl) == (1, "ours") and a.steps >= 1 and a.warmup >= 3  # defaults: N = 1, W >= 3
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "8", "--steps", "20", "--warmup", "5", "--impl", "reference"])
    a = b.parse()
    assert (a.gpus, a.steps, a.warmup, a.impl) == (8, 20, 5, "reference")


def test_clock_sampler_windows():
    b = _bench()
    s = b.ClockSampler(0)
    row = lambda sm, cap: [str(sm), "1965", "700.0", "Not Active", "Not Active", "Not Active", cap]
    s.rows = [(10.0, row(1965, "Not Active")), (10.5, row(1400, "Active")), (10.6, row(1300, "Active")), (10.7, row(1350, "Active")),
              (11.5, row(1965, "Not Active")), (11.6, ["garbage"])]
    burst = s.stop(9.9, 10.1)
    assert burst == {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 1, "power_w_max": 700.0}
    load = s.stop(10.4, 10.8)
    assert load["sm_mhz"] == 1350.0 and load["reasons"] == ["sw_power_cap"] and load["samples"] == 3
    assert s.stop(20.0, 21.0)["sm_mhz"] is None


def test_peaks_come_from_the_driver_file():
    b = _bench()
    p = b.peaks()
    assert p["source"] in ("measured", "fallback") and p["hbm_gbs"] > 1000 and p["bf16_tflops_sustained"] <= p["bf16_tflops"]
