"""SURVEY 8f rank 4 on the CPU: the oracle restatements (MiniROADA forward, sklearn average precision) against the
golden vectors the reference itself produced (oracle/gen_golden_rank4.py), and the host-side contract of the
MROADA module / ANT evaluator mirror (no compute without a GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import metrics_np, miniroad_np
from oracle.map_cases import map_cases, one_hot

import prego_b200
from prego_b200 import synthetic


@pytest.fixture(scope="module")
def meta4():
    return json.load(open(os.path.join(GOLD, "meta_rank4.json")))


def ant_case(meta4, name):
    """(cfg, seeded MROADA whose tensors hash to the reference's seeded init, rgb, flow, golden arrays)."""
    c = meta4["anticipation"][name]
    cfg = dict(getattr(synthetic, c["cfg"]), model="MiniROADA", **c["overrides"])
    torch.manual_seed(meta4["seed"])
    m = prego_b200.MROADA(cfg)
    sha = lambda t: hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()
    # (the key ORDER is asserted against the live reference by the generator; the JSON stores them sorted)
    assert sorted(m.state_dict().keys()) == sorted(c["weights_sha256"].keys()), "state_dict keys differ from the reference"
    for k, v in m.state_dict().items():
        assert sha(v) == c["weights_sha256"][k], f"seeded weight {k} differs from the reference's"
    rgb, flow = synthetic.feature_batch(c["stream_ids"], c["T"], "cpu", False)
    assert sha(rgb) == c["rgb_sha256"] and sha(flow) == c["flow_sha256"], "synthetic feature generator drifted"
    z = np.load(os.path.join(GOLD, f"anticipation_{name}.npz"))
    return cfg, m, rgb, flow, {k: z[k] for k in z.files}


ANT_NAMES = ["epic_a4_b2_t40", "asm_a2_b3_t24_act", "asm_a3_b20_t6"]


@pytest.mark.parametrize("name", ANT_NAMES)
def test_anticipation_restatement_matches_reference(meta4, name):
    cfg, m, rgb, flow, gold = ant_case(meta4, name)
    probs, ant_probs, logits, ant_logits = miniroad_np.forward_anticipation(
        m.state_dict(), rgb.numpy(), flow.numpy(), cfg["anticipation_length"])
    assert ant_probs.shape == gold["ant_probs"].shape == (rgb.shape[0], rgb.shape[1], cfg["anticipation_length"], cfg["num_classes"])
    # fp32 restatement vs ATen fp32: summation order only
    assert np.abs(logits - gold["logits"]).max() <= 1e-4 * np.abs(gold["logits"]).max()
    assert np.abs(ant_logits - gold["ant_logits"]).max() <= 1e-4 * np.abs(gold["ant_logits"]).max()
    assert np.abs(probs - gold["probs"]).max() <= 2e-6
    assert np.abs(ant_probs - gold["ant_probs"]).max() <= 2e-6


def test_mroada_module_contract(meta4):
    cfg, m, _, _, _ = ant_case(meta4, "asm_a2_b3_t24_act")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes["anticipation_layer.0.weight"] == (2 * 1024, 1024) and shapes["anticipation_layer.0.bias"] == (2048,)
    assert shapes["f_actionness.0.weight"] == (1, 1024)  # cfg actionness: container only, as in the reference
    assert len(shapes) == 14
    assert isinstance(prego_b200.build_model(cfg, None), prego_b200.MROADA)
    assert prego_b200.EVAL["ANTICIPATION"] is prego_b200.ANT_Evaluate
    x = torch.zeros(1, 4, 2048)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.eval()(x, x)
    with pytest.raises(RuntimeError, match="inference-only"):
        m.train()(x, x)


def test_map_case_inputs_regenerate_bit_identically(meta4):
    for name, scores, labels in map_cases():
        c = meta4["map"][name]
        assert hashlib.sha256(scores.tobytes()).hexdigest() == c["scores_sha256"], name
        assert hashlib.sha256(labels.tobytes()).hexdigest() == c["labels_sha256"], name


def test_average_precision_restatement_matches_reference(meta4):
    gold = np.load(os.path.join(GOLD, "map_cases.npz"))
    for name, scores, labels in map_cases():
        K = scores.shape[1]
        r = metrics_np.perframe_average_precision(scores, one_hot(labels, K), [str(i) for i in range(K)])
        want = gold[f"{name}.ap"]
        assert [int(k) for k in r["per_class_AP"]] == [k for k in range(1, K) if not np.isnan(want[k])], name
        for k, v in r["per_class_AP"].items():
            assert abs(v - want[int(k)]) <= 1e-12, (name, k)
        assert abs(r["mean_AP"] - float(gold[f"{name}.mean_ap"])) <= 1e-12
        assert len(r["per_class_AP"]) == meta4["map"][name]["classes_scored"]


def test_metrics_need_cuda():
    if torch.cuda.is_available():
        pytest.skip("no-GPU failure mode")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        prego_b200.perframe_average_precision(torch.rand(8, 3), torch.zeros(8, dtype=torch.int32), ["a", "b", "c"])


def test_average_precision_restatement_matches_sklearn_live_on_multilabel():
    """Beyond the committed golden cases: the restatement against the dependency itself (scikit-learn, the library the
    reference calls at utils/metrics.py:43,55) on multi-hot targets, all-positive classes, N = 1."""
    from sklearn.metrics import average_precision_score
    for N, K in [(1, 1), (1, 7), (31, 3), (4097, 2), (12289, 33)]:
        rs = np.random.RandomState(N * 131 + K)
        scores = (rs.randint(0, 1 << 12, (N, K)).astype(np.float32) / np.float32(1 << 12))
        scores[rs.rand(N, K) < 0.05] = 1.0
        scores[rs.rand(N, K) < 0.05] = 0.0
        targets = (rs.rand(N, K) < 0.3).astype(np.float32)
        if K > 2:
            targets[:, 1] = 1.0
            targets[:, 2] = 0.0
        for k in range(K):
            if targets[:, k].any():
                assert abs(metrics_np.average_precision(targets[:, k], scores[:, k]) - average_precision_score(targets[:, k], scores[:, k])) <= 1e-12


def test_c_restatement_of_average_precision_matches_golden_and_numpy():
    """oracle/metrics_oracle.c (plain C, qsort) against the reference-made golden values and the numpy restatement."""
    import ctypes as C
    import subprocess
    from conftest import ROOT
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libmetrics_oracle.so"))
    lib.oracle_average_precision.restype = C.c_double
    lib.oracle_average_precision.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    gold = np.load(os.path.join(GOLD, "map_cases.npz"))
    for name, scores, labels in map_cases():
        want = gold[f"{name}.ap"]
        for k in range(scores.shape[1]):
            col = np.ascontiguousarray(scores[:, k])
            pos = np.ascontiguousarray((labels == k).astype(np.int32))
            got = lib.oracle_average_precision(col.ctypes.data, pos.ctypes.data, col.size)
            if k == 0:  # background: not in the golden dict, compare with the numpy restatement instead
                ref = metrics_np.average_precision(pos, col) if pos.any() else float("nan")
            else:
                ref = want[k]
            assert (np.isnan(got) and np.isnan(ref)) or abs(got - ref) <= 1e-12, (name, k, got, ref)


_LIVE_CHECK = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, "/root/reference/step_recognition")
from model import build_model                      # the reference's registry (model/model_builder.py:7-9)
from oracle import miniroad_np
from prego_b200 import synthetic
torch.set_num_threads(4)
for seed, (B, T), A, K in ((1, (2, 9), 2, 12), (2, (1, 33), 5, 86), (3, (5, 4), 1, 30)):
    cfg = dict(synthetic.EPIC_TENT_O, num_classes=K, model="MiniROADA", anticipation_length=A, actionness=bool(seed % 2))
    torch.manual_seed(seed)
    ref = build_model(cfg, "cpu").eval()
    g = torch.Generator().manual_seed(seed)
    rgb, flow = torch.randn(B, T, 2048, generator=g).abs(), torch.randn(B, T, 2048, generator=g).abs()
    with torch.no_grad():
        out = ref(rgb, flow)
    p, ap, _, _ = miniroad_np.forward_anticipation(ref.state_dict(), rgb.numpy(), flow.numpy(), A)
    assert np.abs(p - out["logits"].numpy()).max() <= 2e-6 and np.abs(ap - out["anticipation_logits"].numpy()).max() <= 2e-6
    cfg = dict(synthetic.EPIC_TENT_O, num_classes=K, model="MiniROAD")
    torch.manual_seed(seed)
    ref = build_model(cfg, "cpu").eval()
    with torch.no_grad():
        want = ref(rgb, flow)["logits"].numpy()
    assert np.abs(miniroad_np.forward(ref.state_dict(), rgb.numpy(), flow.numpy()) - want).max() <= 2e-6
print("live reference == restatement")
"""


@pytest.mark.skipif(not os.path.exists("/root/reference/step_recognition/model"), reason="reference checkout not mounted")
def test_restatements_against_the_live_reference_modules(tmp_path):
    """Beyond the committed golden vectors: MROAD and MROADA imported live from the reference checkout (build container
    only; a subprocess so that the reference's top-level packages do not leak into this one), random seeds / shapes."""
    import subprocess
    import sys
    from conftest import ROOT
    script = tmp_path / "live_check.py"
    script.write_text(_LIVE_CHECK)
    out = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "live reference == restatement" in out.stdout, out.stderr[-2000:]
