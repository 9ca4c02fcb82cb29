"""Generate the committed golden fixtures under tests/golden/ (TEST INFRASTRUCTURE).

Runs ONLY in the build container, where the reference checkout is mounted read-only at
/root/reference.  It imports the reference's own MROAD (step_recognition/model/rnn/rnn.py)
and aggregate (utils/aggregate.py) unmodified, runs them on CPU on seeded synthetic inputs
(prego_b200.synthetic) and stores their outputs; the GPU box never sees /root/reference, it
only sees these fixtures.

  python oracle/gen_golden.py            # rewrites tests/golden/*
"""
from __future__ import annotations

import gzip
import hashlib
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from prego_b200 import synthetic  # noqa: E402


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def load_reference_model_pkg():
    sys.path.insert(0, os.path.join(REF, "step_recognition"))
    from model import build_model  # the reference's registry/builder (model/model_builder.py:7-9)
    return build_model


def ref_forward_all(model, rgb, flow):
    """probs via the reference forward; logits via the reference's own sub-modules."""
    with torch.no_grad():
        probs = model(rgb, flow)["logits"]
        if model.use_rgb and model.use_flow:
            x = torch.cat((rgb, flow), 2)
        elif model.use_rgb:
            x = rgb
        else:
            x = flow
        x = model.layer1(x)
        h0 = model.h0.expand(-1, x.shape[0], -1)
        ht, hT = model.gru(x, h0)
        logits = model.f_classification(model.relu(ht))
    return probs.numpy(), logits.numpy(), hT[0].numpy()


CASES = [
    # name, base cfg, cfg overrides, stream ids, T, zero_flow
    ("epic_b1_t300", "EPIC_TENT_O", {}, [0], 300, False),
    ("asm_b2_t160", "ASSEMBLY101_O", {}, [1, 2], 160, False),
    ("asm_b1_t64_zeroflow", "ASSEMBLY101_O", {}, [3], 64, True),
    ("epic_b1_t96_rgbonly", "EPIC_TENT_O", {"no_flow": True}, [4], 96, False),
    ("asm_b40_t24", "ASSEMBLY101_O", {}, list(range(10, 50)), 24, False),
]


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    build_model = load_reference_model_pkg()
    meta = {"torch": torch.__version__, "seed": 20, "cases": {}, "weights_sha256": {}}
    for name, base, over, sids, T, zero_flow in CASES:
        cfg = dict(getattr(synthetic, base), **over)
        torch.manual_seed(20)
        ref = build_model(cfg, "cpu").eval()
        mine = synthetic.seeded_model(cfg, seed=20)
        wkey = f"{base}{'_rgbonly' if over.get('no_flow') else ''}"
        shas = {}
        for k, v in ref.state_dict().items():
            assert torch.equal(v, mine.state_dict()[k]), f"seeded init differs from the reference for {k}"
            shas[k] = sha(v)
        meta["weights_sha256"][wkey] = shas
        rgb, flow = synthetic.feature_batch(sids, T, "cpu", zero_flow)
        probs, logits, hT = ref_forward_all(ref, rgb, flow)
        np.savez_compressed(os.path.join(GOLD, f"model_{name}.npz"), probs=probs.astype(np.float32),
                            logits=logits.astype(np.float32), h_last=hT.astype(np.float32))
        meta["cases"][name] = {"cfg": base, "overrides": over, "weights": wkey, "stream_ids": sids, "T": T,
                               "zero_flow": zero_flow, "rgb_sha256": sha(rgb), "flow_sha256": sha(flow)}
        print(name, probs.shape, "labels", np.bincount(probs.argmax(-1).ravel())[:6])

    # ---- training step: reference module + reference OadLoss, autograd gradients (dropout 0 for parity)
    from criterions import build_criterion  # the reference's criterions/loss_builder.py
    from oracle.miniroad_torch_cpu import TorchRefMROAD, oad_loss
    cfg = dict(synthetic.EPIC_TENT_O, dropout=0.0, loss="NONUNIFORM")
    torch.manual_seed(20)
    ref = build_model(cfg, "cpu").train()
    crit = build_criterion(cfg, "cpu")
    sids, T = [60, 61, 62], 10
    rgb, flow = synthetic.feature_batch(sids, T, "cpu", False)
    target = torch.stack([synthetic.targets(s, T, 12) for s in sids])
    loss = crit(ref(rgb, flow), target)
    loss.backward()
    port = TorchRefMROAD(4096, 2048, 1024, 12, 0.0).train()
    port.load_state_dict(ref.state_dict())
    ploss = oad_loss(port(rgb, flow)["logits"], target)
    ploss.backward()
    assert abs(float(ploss) - float(loss)) < 1e-6
    grads = {}
    for (k, p_ref), (_, p_port) in zip(ref.named_parameters(), port.named_parameters()):
        assert torch.allclose(p_ref.grad, p_port.grad, rtol=1e-5, atol=1e-8), k
        g = p_ref.grad.reshape(-1)
        grads[k + ".head"] = g[:64].numpy().copy()
        grads[k + ".stats"] = np.array([g.sum().item(), g.abs().sum().item(), g.norm().item(), g.abs().max().item()])
    np.savez_compressed(os.path.join(GOLD, "train_epic_b3_t10.npz"), loss=np.array(float(loss)), **grads)
    meta["train_case"] = {"cfg": "EPIC_TENT_O", "dropout": 0.0, "stream_ids": sids, "T": T, "loss": float(loss)}
    print("train case loss", float(loss))

    # aggregate golden pair (the reference's only known-answer vectors)
    src_in = os.path.join(REF, "output_miniRoad", "output_miniROAD.json")
    src_out = os.path.join(REF, "data", "output", "aggregated_data.json")
    data = json.load(open(src_in))
    with gzip.GzipFile(os.path.join(GOLD, "aggregate_input_epic_tent.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(data).encode())
    expected = open(src_out, "rb").read()
    open(os.path.join(GOLD, "aggregate_expected_epic_tent.json"), "wb").write(expected)
    meta["aggregate"] = {"expected_sha256": hashlib.sha256(expected).hexdigest(),
                         "videos": len(data), "frames": sum(len(v["pred"]) for v in data.values())}
    # re-run the reference function itself to confirm the pair (and build KATs from it)
    spec = importlib.util.spec_from_file_location("ref_aggregate", os.path.join(REF, "utils", "aggregate.py"))
    ref_agg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_agg)
    tmp = "/tmp/_agg_check.json"
    ref_agg.aggregate(data, tmp)
    assert open(tmp, "rb").read() == expected, "reference aggregate() does not reproduce its own fixture"
    kats = []
    kat_inputs = [
        ([1] * 150 + [2] * 150, None),
        ([3] * 100 + [1] * 100, None),
        ([5], [0]),
        ([0] * 200 + [7], [0] * 201),
        ([2] * 199 + [9] + [2] * 200 + [9] * 3, [2, 2, 9, 9, 2]),
        (list(np.random.RandomState(7).randint(0, 5, 1234)), list(np.random.RandomState(8).randint(0, 3, 1234))),
        (list(np.random.RandomState(9).randint(0, 86, 9507)), list(np.repeat(np.arange(40), 238)[:9507])),
    ]
    for pred, gt in kat_inputs:
        gt = pred if gt is None else gt
        ref_agg.aggregate({"v": {"pred": [int(x) for x in pred], "gt": [int(x) for x in gt]}}, tmp)
        kats.append({"pred": [int(x) for x in pred], "gt": [int(x) for x in gt], "expected": json.load(open(tmp))["v"]})
    with gzip.GzipFile(os.path.join(GOLD, "aggregate_kats.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(kats).encode())
    json.dump(meta, open(os.path.join(GOLD, "meta.json"), "w"), indent=1, sort_keys=True)
    print("wrote", sorted(os.listdir(GOLD)))


if __name__ == "__main__":
    main()
