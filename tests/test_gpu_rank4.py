"""SURVEY 8f rank 4 on the B200, through the C ABI: the MiniROADA anticipation head against the reference-made
golden vectors, and the device per-frame mAP against the reference's (sklearn) values.

Tolerances: the ones of tests/test_gpu_parity.py for logits (fp32 1e-4, fp16 2e-3, bf16 1e-2, relative to max|logit|;
label flips only on near-ties); average precision is float64 arithmetic over exactly-ordered integers: 1e-12 absolute.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import metrics_np, miniroad_np
from oracle.map_cases import map_cases, one_hot
from test_oracle_rank4 import ANT_NAMES, ant_case

pytestmark = pytest.mark.gpu

REL = {"fp32": 1e-4, "fp16x3": 1e-4, "bf16": 1e-2, "fp16": 2e-3}
AP_TOL = 1e-12


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need the B200"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def meta4():
    return json.load(open(os.path.join(GOLD, "meta_rank4.json")))


def _check_logits(got, ref, rel, what):
    err = np.abs(got - ref).max()
    assert np.isfinite(got).all() and err <= rel * np.abs(ref).max(), f"{what}: max |dlogit| {err:.3e} vs {rel} * {np.abs(ref).max():.3f}"
    return err


def _check_labels(labels, ref_logits, err, what):
    bad = labels != ref_logits.argmax(-1)
    margin = miniroad_np.top2_margin(ref_logits)
    assert np.all(margin[bad] < 4 * err + 1e-6), f"{what}: label flip away from a near-tie"


@pytest.mark.parametrize("prec", ["fp32", "fp16x3", "fp16", "bf16"])
@pytest.mark.parametrize("name", ANT_NAMES)
def test_anticipation_forward_matches_reference(dev, meta4, name, prec):
    cfg, m, rgb, flow, gold = ant_case(meta4, name)
    m = m.to(dev).eval()
    out = m.infer(rgb.to(dev), flow.to(dev), want_probs=True, want_logits=True, precision=prec, want_anticipation=True,
                  want_anticipation_logits=True)
    torch.cuda.synchronize()
    e0 = _check_logits(out["logits"].cpu().numpy(), gold["logits"], REL[prec], "trunk")
    e1 = _check_logits(out["anticipation_logits"].cpu().numpy(), gold["ant_logits"], REL[prec], "anticipation")
    _check_labels(out["labels"].cpu().numpy(), gold["logits"], e0, "trunk")
    _check_labels(out["anticipation_labels"].cpu().numpy(), gold["ant_logits"], e1, "anticipation")
    ap = out["anticipation_probs"].cpu().numpy()
    assert ap.shape == gold["ant_probs"].shape and np.abs(ap.sum(-1) - 1).max() < 1e-5
    assert np.abs(ap - gold["ant_probs"]).max() <= {"fp32": 2e-6, "fp16x3": 1e-5, "fp16": 2e-3, "bf16": 1e-2}[prec]
    # probabilities are the softmax of the returned logits; labels are their first maximum
    assert np.array_equal(out["anticipation_labels"].cpu().numpy(), ap.argmax(-1))
    assert m.device_error() == 0


def test_module_forward_returns_reference_dict(dev, meta4):
    cfg, m, rgb, flow, gold = ant_case(meta4, "epic_a4_b2_t40")
    m.precision = "fp32"
    m = m.to(dev).eval()
    with torch.no_grad():
        out = m(rgb.to(dev), flow.to(dev))
    assert set(out) == {"logits", "anticipation_logits"}
    assert np.abs(out["logits"].cpu().numpy() - gold["probs"]).max() <= 2e-6
    assert np.abs(out["anticipation_logits"].cpu().numpy() - gold["ant_probs"]).max() <= 2e-6
    assert tuple(m.last_anticipation_labels.shape) == (2, 40, 4)
    # in-place weight update is picked up (the anticipation tensors are re-packed)
    with torch.no_grad():
        m.anticipation_layer[0].bias.add_(0.5)
        out2 = m(rgb.to(dev), flow.to(dev))
    assert (out2["anticipation_logits"] - out["anticipation_logits"]).abs().max() > 1e-4
    assert torch.equal(out2["logits"], out["logits"])


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
def test_anticipation_slabs_and_time_chunks_are_invisible(dev, meta4, prec):
    """Row slabs of the [rows, A*H] activation and time chunking with carried state give the one-pass result."""
    cfg, m, rgb, flow, _ = ant_case(meta4, "asm_a3_b20_t6")
    m = m.to(dev).eval()
    rgb, flow = rgb.to(dev).repeat(1, 20, 1), flow.to(dev).repeat(1, 20, 1)  # 20 streams x 120 frames = 2400 rows
    kw = dict(want_logits=True, precision=prec, want_anticipation=True, want_anticipation_logits=True)
    whole = m.infer(rgb, flow, **kw)
    m.anticipation_slab_rows = 128
    m._ant_workspace = None
    slabbed = m.infer(rgb, flow, **kw)
    for k in ("anticipation_logits", "anticipation_probs", "anticipation_labels", "logits"):
        assert torch.equal(whole[k], slabbed[k]), k
    m.anticipation_slab_rows = 384
    m._ant_workspace = None
    chunked = m.infer(rgb, flow, chunk_T=7, **kw)
    if prec == "fp32":
        assert (chunked["anticipation_logits"] - whole["anticipation_logits"]).abs().max() <= 1e-5
    else:
        assert (chunked["anticipation_logits"] - whole["anticipation_logits"]).abs().max() <= 2e-3 * whole["anticipation_logits"].abs().max()
    assert (chunked["anticipation_labels"] != whole["anticipation_labels"]).float().mean() <= 1e-3


def test_anticipation_abi_errors(dev, meta4):
    from prego_b200 import synthetic
    plain = synthetic.seeded_model(dict(synthetic.EPIC_TENT_O), seed=20, device=dev)
    rgb, flow = synthetic.feature_batch([0], 4, dev)
    with pytest.raises(RuntimeError, match="MROADA"):
        plain.infer(rgb, flow, want_anticipation=True)


# ------------------------------------------------------------------------------ per-frame mAP
@pytest.mark.parametrize("fmt", ["onehot", "labels"])
def test_perframe_ap_matches_reference_golden(dev, meta4, fmt):
    from prego_b200 import perframe_average_precision
    gold = np.load(os.path.join(GOLD, "map_cases.npz"))
    for name, scores, labels in map_cases():
        K = scores.shape[1]
        tgt = torch.from_numpy(one_hot(labels, K)) if fmt == "onehot" else torch.from_numpy(labels)
        r = perframe_average_precision(torch.from_numpy(scores).to(dev), tgt.to(dev), [str(i) for i in range(K)])
        want = gold[f"{name}.ap"]
        assert [int(k) for k in r["per_class_AP"]] == [k for k in range(1, K) if not np.isnan(want[k])], name
        for k, v in r["per_class_AP"].items():
            assert abs(v - want[int(k)]) <= AP_TOL, (name, k, v, want[int(k)])
        assert abs(r["mean_AP"] - float(gold[f"{name}.mean_ap"])) <= AP_TOL, name
        # metrics.py:58: '[true: <positives>, pred:<int(sum of the class's scores)>, AP:<percent, 1 decimal>]'
        assert r["num"] == {str(k): f"[true: {int((labels == k).sum())}, pred:{int(scores[:, k].astype(np.float64).sum())}, AP:{r['per_class_AP'][str(k)] * 100:.1f}]"
                            for k in range(1, K) if (labels == k).any()}


def test_perframe_ap_vs_oracle_on_model_probabilities(dev):
    """Real softmax outputs of the CUDA path (fp16), 9 600 frames: device mAP == oracle mAP on the same numbers."""
    from prego_b200 import perframe_average_precision, synthetic
    cfg = dict(synthetic.ASSEMBLY101_O)
    model = synthetic.seeded_model(cfg, seed=20, device=dev)
    rgb, flow = synthetic.device_features(32, 300, dev, seed=9)
    probs = model.infer(rgb, flow, want_probs=True)["probs"].reshape(-1, 86)
    labels = torch.stack([synthetic.targets(500 + s, 300, 86) for s in range(32)]).argmax(-1).reshape(-1).to(torch.int32)
    r = perframe_average_precision(probs, labels.to(dev), [str(i) for i in range(86)])
    want = metrics_np.perframe_average_precision(probs.cpu().numpy(), one_hot(labels.numpy(), 86), [str(i) for i in range(86)])
    assert list(r["per_class_AP"]) == list(want["per_class_AP"])
    for k, v in want["per_class_AP"].items():
        assert abs(r["per_class_AP"][k] - v) <= AP_TOL
    assert abs(r["mean_AP"] - want["mean_AP"]) <= AP_TOL


def test_perframe_ap_full_size_properties(dev):
    """4 M frames x 86 classes (beyond what the oracle finishes in seconds): size-independent properties."""
    from prego_b200.metrics import average_precision_per_class
    N, K = 1 << 22, 86
    g = torch.Generator(device=dev).manual_seed(3)
    labels = torch.randint(0, K, (N,), generator=g, device=dev, dtype=torch.int32)
    scores = torch.rand(N, K, generator=g, device=dev)
    ap, npos = average_precision_per_class(scores, labels)
    assert np.array_equal(npos, np.bincount(labels.cpu().numpy(), minlength=K))
    assert np.all(np.abs(ap - 1.0 / K) < 2e-3)                      # uninformative scores: AP ~ prevalence
    perm = torch.randperm(N, generator=g, device=dev)                # frame order is irrelevant
    ap2, _ = average_precision_per_class(scores[perm].contiguous(), labels[perm].contiguous())
    assert np.abs(ap - ap2).max() <= AP_TOL
    onehot = torch.nn.functional.one_hot(labels.long(), K).float()
    ap3, _ = average_precision_per_class(onehot * 0.5 + 0.25, labels)  # perfect ranking, two thresholds
    assert np.abs(ap3 - 1.0).max() <= AP_TOL
    ap4, npos4 = average_precision_per_class(torch.full((N, K), 0.5, device=dev), labels)  # one threshold: AP = P / N
    assert np.abs(ap4 - npos4 / N).max() <= AP_TOL


def test_perframe_ap_rejects_non_probabilities(dev):
    from prego_b200 import perframe_average_precision
    s = torch.rand(100, 4, device=dev)
    s[17, 2] = 1.5
    with pytest.raises(ValueError, match="probabilities"):
        perframe_average_precision(s, torch.zeros(100, dtype=torch.int32, device=dev), list("abcd"))
    s[17, 2] = float("nan")
    with pytest.raises(ValueError, match="probabilities"):
        perframe_average_precision(s, torch.zeros(100, dtype=torch.int32, device=dev), list("abcd"))


def test_ant_evaluate_matches_oracle(dev, meta4):
    """ANT_Evaluate (eval.py:85-163 mirror): OAD mAP + per-step anticipation mAPs equal the oracle's on the same outputs."""
    from prego_b200 import build_eval, synthetic
    cfg, m, _, _, _ = ant_case(meta4, "epic_a4_b2_t40")
    cfg = dict(cfg, task="ANTICIPATION", metric="AP")
    m.precision = "fp32"
    m = m.to(dev).eval()
    A, K = 4, 12
    loader, all_p, all_ap, all_t, all_at = [], [], [], [], []
    for i, T in enumerate([150, 97, 260]):
        rgb, flow = synthetic.feature_batch([300 + i], T, "cpu")
        tgt = synthetic.targets(300 + i, T + A, K)  # [T + A, K]; ant_target[t, a] = target[t + 1 + a] (dataset.py:213-214)
        ant = torch.stack([tgt[t + 1:t + 1 + A] for t in range(T)])
        loader.append((rgb, flow, tgt[:T].unsqueeze(0), ant.unsqueeze(0)))
        out = m.infer(rgb.to(dev), flow.to(dev), precision="fp32", want_anticipation=True)
        all_p.append(out["probs"][0].cpu().numpy()); all_ap.append(out["anticipation_probs"][0].cpu().numpy())
        all_t.append(tgt[:T].numpy()); all_at.append(ant.numpy())
    ev = build_eval(cfg)
    got = ev(m, loader, None, dev)
    names = [str(i) for i in range(K)]
    p, t, ap, at = map(np.concatenate, (all_p, all_t, all_ap, all_at))
    want_oad = metrics_np.perframe_average_precision(p, t, names)["mean_AP"]
    want_steps = [metrics_np.perframe_average_precision(ap[:, a], at[:, a], names)["mean_AP"] for a in range(A)]
    assert abs(ev.last_result["mean_AP"] - want_oad) <= AP_TOL
    for a in range(A):
        assert abs(ev.last_result[f"anticipation_{a+1}"]["mean_AP"] - want_steps[a]) <= AP_TOL
    assert abs(got - np.mean(want_steps)) <= AP_TOL


def test_main_entry_miniroada_synthetic(dev, tmp_path, monkeypatch):
    """python -m prego_b200.main --config configs/miniroada_synthetic.yaml --eval synthetic --synthetic 3: registry ->
    MROADA -> ANT_Evaluate end to end (zero-flow dummy recognised on the host), mean anticipation mAP == the oracle's on
    the same outputs."""
    from conftest import ROOT
    from prego_b200 import main as pmain, synthetic
    import yaml
    monkeypatch.chdir(tmp_path)
    cfg_path = os.path.join(ROOT, "configs", "miniroada_synthetic.yaml")
    got = pmain.main(["--config", cfg_path, "--eval", "synthetic", "--synthetic", "3", "--device", "cuda:0", "--precision", "fp32"])
    cfg = yaml.safe_load(open(cfg_path))
    cfg.update(no_rgb=False, no_flow=False, precision="fp32")
    pmain.set_seed(20)
    model = __import__("prego_b200").build_model(cfg, dev).eval()
    ds = pmain.SyntheticAnticipation(cfg, 3)
    A, K = 4, 86
    ps, ts = [], []
    for i in range(3):
        rgb, flow, _t, ant = ds[i]
        out = model.infer(rgb.unsqueeze(0).to(dev), flow.unsqueeze(0).to(dev), precision="fp32", want_anticipation=True)
        ps.append(out["anticipation_probs"][0].cpu().numpy())
        ts.append(ant.numpy())
    p, t = np.concatenate(ps), np.concatenate(ts)
    names = [str(i) for i in range(K)]
    want = np.mean([metrics_np.perframe_average_precision(p[:, a], t[:, a], names)["mean_AP"] for a in range(A)])
    assert abs(got - want) <= 1e-9
    with pytest.raises(RuntimeError, match="inference-only"):
        pmain.main(["--config", cfg_path, "--synthetic", "2"])


@pytest.mark.parametrize("N,K", [(1, 1), (1, 7), (31, 3), (4097, 2), (12289, 33)])
def test_perframe_ap_multilabel_and_tiny_shapes_vs_oracle(dev, N, K):
    """Multi-hot targets (several positive classes per frame, classes that are all-positive or empty), N = 1, K = 1,
    tile and slice boundaries: the device result equals the oracle (sklearn restatement) class by class."""
    from prego_b200.metrics import average_precision_per_class
    rs = np.random.RandomState(N * 131 + K)
    scores = (rs.randint(0, 1 << 12, (N, K)).astype(np.float32) / np.float32(1 << 12))
    scores[rs.rand(N, K) < 0.05] = 1.0
    scores[rs.rand(N, K) < 0.05] = 0.0
    targets = (rs.rand(N, K) < 0.3).astype(np.float32)
    if K > 2:
        targets[:, 1] = 1.0   # every frame positive
        targets[:, 2] = 0.0   # no positive: skipped by the reference (metrics.py:54)
    ap, npos = average_precision_per_class(torch.from_numpy(scores).to(dev), torch.from_numpy(targets).to(dev))
    assert np.array_equal(npos, targets.sum(0).astype(np.int64))
    for k in range(K):
        if targets[:, k].any():
            assert abs(ap[k] - metrics_np.average_precision(targets[:, k], scores[:, k])) <= AP_TOL, (k, ap[k])
        else:
            assert np.isnan(ap[k])


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
@pytest.mark.parametrize("over", [{"num_classes": 100, "anticipation_length": 2}, {"no_flow": True, "anticipation_length": 3},
                                  {"num_classes": 128, "anticipation_length": 5, "actionness": True}])
def test_anticipation_non_shipped_shapes_vs_oracle(dev, prec, over):
    """Shapes no golden case covers (the 128-column head tile, rgb-only input, odd A): CUDA path vs the numpy oracle,
    which the golden cases pin on the shipped shapes."""
    from prego_b200 import MROADA, synthetic
    cfg = dict(synthetic.EPIC_TENT_O, model="MiniROADA", actionness=False)
    cfg.update(over)
    torch.manual_seed(7)
    m = MROADA(cfg).to(dev).eval()
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    B, T = 18, 5  # B > 16: batched tcgen05 recurrence
    rgb, flow = synthetic.feature_batch(list(range(900, 900 + B)), T, "cpu")
    _, _, ref_l, ref_al = miniroad_np.forward_anticipation(sd, rgb.numpy(), flow.numpy(), cfg["anticipation_length"],
                                                           use_flow=not cfg["no_flow"])
    out = m.infer(rgb.to(dev), None if cfg["no_flow"] else flow.to(dev), want_logits=True, precision=prec,
                  want_anticipation=True, want_anticipation_logits=True)
    e0 = _check_logits(out["logits"].cpu().numpy(), ref_l, REL[prec], "trunk")
    e1 = _check_logits(out["anticipation_logits"].cpu().numpy(), ref_al, REL[prec], "anticipation")
    _check_labels(out["anticipation_labels"].cpu().numpy(), ref_al, e1, "anticipation")
    ap = out["anticipation_probs"].cpu().numpy()
    assert ap.shape == (B, T, cfg["anticipation_length"], cfg["num_classes"]) and np.abs(ap.sum(-1) - 1).max() < 1e-5
    assert e0 >= 0
