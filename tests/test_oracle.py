"""The oracle (CPU restatement) against the golden vectors produced by the reference itself."""
import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLD, ROOT, case_inputs, load_gz_json, load_model_case, seeded_weights_checked
from oracle import aggregate_np, miniroad_np

MODEL_CASES = ["epic_b1_t300", "asm_b2_t160", "asm_b1_t64_zeroflow", "epic_b1_t96_rgbonly"]


@pytest.mark.parametrize("name", MODEL_CASES)
def test_numpy_restatement_matches_reference(golden_meta, name):
    gold = load_model_case(name)
    cfg, rgb, flow = case_inputs(golden_meta, name)
    sd = seeded_weights_checked(golden_meta, name).state_dict()
    probs, logits, h_last = miniroad_np.forward(sd, rgb.numpy(), flow.numpy(), use_rgb=not cfg["no_rgb"],
                                                use_flow=not cfg["no_flow"], return_all=True)
    # fp32 restatement vs ATen fp32: different summation order only
    assert np.abs(logits - gold["logits"]).max() <= 1e-4 * np.abs(gold["logits"]).max()
    assert np.abs(probs - gold["probs"]).max() <= 2e-6
    assert np.abs(h_last - gold["h_last"]).max() <= 2e-5
    ref_labels = gold["probs"].argmax(-1)
    mine = miniroad_np.labels_from_probs(probs)
    bad = mine != ref_labels
    margin = miniroad_np.top2_margin(gold["logits"])
    assert bad.mean() <= 1e-3 and np.all(margin[bad] < 1e-4)


def test_numpy_restatement_fp64_is_tighter(golden_meta):
    name = "epic_b1_t96_rgbonly"
    gold = load_model_case(name)
    cfg, rgb, flow = case_inputs(golden_meta, name)
    sd = seeded_weights_checked(golden_meta, name).state_dict()
    probs = miniroad_np.forward(sd, rgb.numpy(), flow.numpy(), use_flow=False, dtype=np.float64)
    assert np.abs(probs - gold["probs"]).max() <= 1e-6


def test_streaming_equals_whole_sequence(golden_meta):
    """Carried-state chunking (online inference) reproduces the whole-sequence result."""
    name = "epic_b1_t96_rgbonly"
    cfg, rgb, flow = case_inputs(golden_meta, name)
    sd = seeded_weights_checked(golden_meta, name).state_dict()
    whole = miniroad_np.forward(sd, rgb.numpy(), flow.numpy(), use_flow=False)
    h, parts = None, []
    for s in range(0, 96, 25):
        p, _, h = miniroad_np.forward(sd, rgb.numpy()[:, s:s + 25], flow.numpy()[:, s:s + 25], use_flow=False,
                                      h0=h, return_all=True)
        parts.append(p)
    assert np.abs(np.concatenate(parts, 1) - whole).max() <= 1e-6


# ------------------------------------------------------------------------------ aggregate
def test_aggregate_np_golden_pair_byte_exact(golden_meta):
    data = load_gz_json("aggregate_input_epic_tent.json.gz")
    out = json.dumps(aggregate_np.aggregate_dict(data)).encode()
    assert hashlib.sha256(out).hexdigest() == golden_meta["aggregate"]["expected_sha256"]
    assert out == open(os.path.join(GOLD, "aggregate_expected_epic_tent.json"), "rb").read()


def test_aggregate_np_kats():
    for kat in load_gz_json("aggregate_kats.json.gz"):
        assert aggregate_np.aggregate_video(kat["pred"], kat["gt"]) == kat["expected"]


def test_aggregate_empty_raises():
    with pytest.raises(IndexError):
        aggregate_np.aggregate_video([], [])


@pytest.fixture(scope="module")
def c_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libaggregate_oracle.so"))
    lib.oracle_aggregate_pred.restype = C.c_int64
    lib.oracle_rle.restype = C.c_int64
    return lib


def _c_collapse(lib, seq, window=None, num_labels=None):
    a = np.ascontiguousarray(seq, dtype=np.int32)
    vals = np.empty(max(len(a), 1), np.int32)
    chg = np.empty(max(len(a), 1), np.int64)
    p = lambda x, t: x.ctypes.data_as(C.POINTER(t))
    if window is None:
        n = lib.oracle_rle(p(a, C.c_int32), C.c_int64(len(a)), p(vals, C.c_int32), p(chg, C.c_int64))
    else:
        n = lib.oracle_aggregate_pred(p(a, C.c_int32), C.c_int64(len(a)), C.c_int32(window), C.c_int32(num_labels),
                                      p(vals, C.c_int32), p(chg, C.c_int64))
    return n, vals[:max(n, 0)].tolist(), chg[:max(n, 0)].tolist()


def test_aggregate_c_oracle(c_oracle):
    data = load_gz_json("aggregate_input_epic_tent.json.gz")
    expected = json.load(open(os.path.join(GOLD, "aggregate_expected_epic_tent.json")))
    for vid, v in data.items():
        n, vals, chg = _c_collapse(c_oracle, v["pred"], 200, max(v["pred"]) + 1)
        assert vals == expected[vid]["pred"] and chg == expected[vid]["changes_pred"]
        n, vals, chg = _c_collapse(c_oracle, v["gt"])
        assert vals == expected[vid]["gt"] and chg == expected[vid]["changes_gt"]
    for kat in load_gz_json("aggregate_kats.json.gz"):
        n, vals, chg = _c_collapse(c_oracle, kat["pred"], 200, max(kat["pred"]) + 1)
        assert vals == kat["expected"]["pred"] and chg == kat["expected"]["changes_pred"]
    assert _c_collapse(c_oracle, [])[0] == -1


# ------------------------------------------------------------------ property tests (hypothesis)
def _reference_aggregate_module():
    """The reference's own utils/aggregate.py, imported live when the checkout is mounted (the build container);
    None on machines without it (the GPU box never sees /root/reference)."""
    import importlib.util
    path = "/root/reference/utils/aggregate.py"
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("ref_aggregate_live", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_aggregate_restatements_agree_on_random_sequences(c_oracle, tmp_path):
    """numpy restatement == plain-C restatement (== the reference's own functions, when the checkout is present) on
    random label sequences: short and long runs, lengths around the 200-frame window, one-element sequences."""
    from hypothesis import given, settings, strategies as st
    ref = _reference_aggregate_module()
    runs = st.lists(st.tuples(st.integers(0, 11), st.integers(1, 260)), min_size=1, max_size=12)

    @settings(max_examples=120, deadline=None)
    @given(runs, runs)
    def check(pred_runs, gt_runs):
        pred = [l for l, n in pred_runs for _ in range(n)]
        gt = [l for l, n in gt_runs for _ in range(n)]
        want = aggregate_np.aggregate_video(pred, gt)
        n, vals, chg = _c_collapse(c_oracle, pred, 200, 12)
        assert (vals, chg) == (want["pred"], want["changes_pred"])
        n, vals, chg = _c_collapse(c_oracle, gt)
        assert (vals, chg) == (want["gt"], want["changes_gt"])
        if ref is not None:
            out = tmp_path / "o.json"
            ref.aggregate({"v": {"pred": pred, "gt": gt}}, str(out))
            assert json.load(open(out))["v"] == want

    check()


# ------------------------------------------------------------------------------ long sequences (reference's own eval shape)
@pytest.mark.parametrize("name", ["epic_b1_t12531", "asm_b1_t9507"])
def test_numpy_restatement_matches_reference_on_whole_videos(name):
    """Goldens made by the live reference at its real evaluation lengths (oracle/gen_golden_long.py): the fp32
    restatement must still agree at the END of 10^4 recurrence steps."""
    from prego_b200 import synthetic
    meta = json.load(open(os.path.join(GOLD, "meta_long.json")))
    c = meta["cases"][name]
    z = np.load(os.path.join(GOLD, f"long_{name}.npz"))
    cfg = dict(getattr(synthetic, c["cfg"]))
    rgb, flow = synthetic.feature_batch([c["stream_id"]], c["T"], "cpu", False)
    sha = lambda t: hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()
    assert sha(rgb) == c["rgb_sha256"] and sha(flow) == c["flow_sha256"]
    sd = synthetic.seeded_model(cfg, seed=meta["seed"]).state_dict()
    probs, logits, h_last = miniroad_np.forward(sd, rgb.numpy(), flow.numpy(), return_all=True)
    fr = z["frames"]
    assert np.abs(logits[0][fr] - z["logits"]).max() <= 1e-4 * float(z["max_abs_logit"])
    assert np.abs(probs[0][fr] - z["probs"]).max() <= 2e-6
    assert np.abs(h_last[0] - z["h_last"]).max() <= 5e-5
    bad = miniroad_np.labels_from_probs(probs[0]) != z["labels"]
    assert bad.mean() <= 1e-3 and np.all(z["margin"][bad] < 1e-4)
    assert hashlib.sha256(z["labels"].astype(np.int16).tobytes()).hexdigest() == c["labels_sha256"]
