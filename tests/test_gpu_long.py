"""Parity evidence at the sizes the reference really runs (VERDICT r01 "missing" 1, "weak" 1).

* whole videos of 9 507 / 12 531 / 31 114 frames as ONE sequence (``test_batch_size: 1``, configs/*.yaml:17;
  datasets/dataset.py:120-123; model/rnn/rnn.py:60-61) against goldens made by the live reference
  (``oracle/gen_golden_long.py``): drift of the recurrence over 10^4 steps is pinned at the END of the sequence,
  both on the few-stream kernel (B = 1) and on the batched tcgen05 recurrence, which re-rounds ``h`` to 16 bits every
  step (the same video replicated over 32 streams so that ``B > 16`` selects it);
* label agreement over > 10^5 frames (128 streams x 1 024 frames) against the fp32 numpy oracle, reported with and
  without the near-tie carve-out;
* the actual bench shape (4 096 streams x 64-frame chunks, carried state) against the oracle on 256 sampled streams.

Tolerances are the ones stated in tests/test_gpu_parity.py; "near-tie" = reference top-2 logit margin below
4 x max|delta logit| of that run.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import miniroad_np

pytestmark = pytest.mark.gpu

REL = {"fp32": 1e-4, "fp16x3": 1e-4, "bf16": 1e-2, "fp16": 2e-3}
LONG_CASES = ["epic_b1_t12531", "epic_b1_t31114", "asm_b1_t9507"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need the B200"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def meta_long():
    return json.load(open(os.path.join(GOLD, "meta_long.json")))


def _long_case(meta, name, dev):
    import hashlib

    from prego_b200 import synthetic
    c = meta["cases"][name]
    z = np.load(os.path.join(GOLD, f"long_{name}.npz"))
    gold = {k: z[k] for k in z.files}
    cfg = dict(getattr(synthetic, c["cfg"]))
    rgb, flow = synthetic.feature_batch([c["stream_id"]], c["T"], "cpu", False)
    sha = lambda t: hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()
    assert sha(rgb) == c["rgb_sha256"] and sha(flow) == c["flow_sha256"], "synthetic feature generator drifted"
    model = synthetic.seeded_model(cfg, seed=meta["seed"], device=dev)
    return gold, model, rgb.to(dev), flow.to(dev)


def _check_long(out, h_last, gold, prec, row=0):
    fr = gold["frames"]
    logits = out["logits"][row].cpu().numpy()
    labels = out["labels"][row].cpu().numpy()
    scale = float(gold["max_abs_logit"])
    d = np.abs(logits[fr] - gold["logits"])
    dmax, dend = float(d.max()), float(d[-64:].max())
    assert np.isfinite(logits).all()
    assert dmax <= REL[prec] * scale, f"logit error {dmax:.3e} > {REL[prec]} * {scale:.3f}"
    ref = gold["labels"].astype(np.int64)
    bad = labels != ref
    eps = max(4 * dmax, 1e-5)
    assert np.all(gold["margin"][bad] < eps), f"label differs away from a near-tie (margins {gold['margin'][bad][:5]}, eps {eps:.2e})"
    dh = float(np.abs(h_last - gold["h_last"]).max())
    assert dh <= 2 * REL[prec], f"final state error {dh:.3e}"
    return dmax / scale, dend / scale, 1.0 - bad.mean(), int(bad.sum()), dh


@pytest.mark.parametrize("name", LONG_CASES)
@pytest.mark.parametrize("prec", ["fp32", "fp16x3", "fp16"])
def test_whole_video_single_stream_vs_reference(dev, meta_long, name, prec):
    """B = 1, the reference's own evaluation shape: one whole video per forward."""
    gold, model, rgb, flow = _long_case(meta_long, name, dev)
    h = torch.zeros(1, 1024, device=dev)
    out = model.infer(rgb, flow, h_state=h, want_probs=False, want_logits=True, precision=prec)
    torch.cuda.synchronize()
    assert model.device_error() == 0
    rel, rel_end, agree, nbad, dh = _check_long(out, h.cpu().numpy()[0], gold, prec)
    print(f"[{prec} {name} B=1] rel logit err {rel:.2e} (last 64 frames {rel_end:.2e}), |dh_T| {dh:.2e}, "
          f"labels {agree:.5f} ({nbad} flips, all near-ties)")
    if prec in ("fp32", "fp16x3"):
        assert agree >= 0.9999


@pytest.mark.parametrize("name", LONG_CASES)
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_whole_video_on_the_batched_recurrence_vs_reference(dev, meta_long, name, prec):
    """The same video on 32 streams: B > 16 selects the tcgen05 recurrence, whose operand copy of h is re-rounded to
    16 bits at every one of the T steps (time-chunked, fp32 master state carried).  Every row must reproduce the
    reference's result for that video within the stated tolerance at the END of the sequence."""
    gold, model, rgb, flow = _long_case(meta_long, name, dev)
    B = 32
    rgb_b, flow_b = rgb.expand(B, -1, -1).contiguous(), flow.expand(B, -1, -1).contiguous()
    h = torch.zeros(B, 1024, device=dev)
    out = model.infer(rgb_b, flow_b, h_state=h, want_probs=False, want_logits=True, precision=prec)
    torch.cuda.synchronize()
    assert model.device_error() == 0
    for row in (0, B - 1):
        rel, rel_end, agree, nbad, dh = _check_long(out, h.cpu().numpy()[row], gold, prec, row)
        print(f"[{prec} {name} B={B} row {row}] rel logit err {rel:.2e} (last 64 frames {rel_end:.2e}), |dh_T| {dh:.2e}, "
              f"labels {agree:.5f} ({nbad} flips, all near-ties)")
    assert torch.equal(out["labels"][0], out["labels"][B - 1]), "identical streams must give identical labels"
    if prec == "fp16":
        assert agree >= 0.999


def _agreement(labels, logits, ref_logits, ref_probs):
    ref = ref_probs.argmax(-1)
    bad = labels != ref
    dmax = float(np.abs(logits - ref_logits).max())
    margin = miniroad_np.top2_margin(ref_logits)
    eps = 4 * dmax
    near = margin < eps
    return {"frames": int(ref.size), "flips": int(bad.sum()), "raw": 1.0 - bad.mean(),
            "flips_away_from_near_ties": int((bad & ~near).sum()), "near_tie_frames": int(near.sum()),
            "carved": 1.0 - (bad & ~near).sum() / max(int((~near).sum()), 1), "dmax": dmax,
            "rel": dmax / float(np.abs(ref_logits).max()), "eps": eps}


def test_label_agreement_over_100k_frames(dev):
    """131 072 frames (128 streams x 1 024 frames, K = 86, time-chunked with carried state) against the fp32 numpy
    oracle: the 99.9 % bar of north_star measured on a sample where it means something."""
    from prego_b200 import synthetic
    cfg = dict(synthetic.ASSEMBLY101_O)
    model = synthetic.seeded_model(cfg, seed=20, device=dev)
    B, T = 128, 1024
    rgb, flow = synthetic.device_features(B, T, dev, seed=4242)
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    ref_probs, ref_logits, _ = miniroad_np.forward(sd, rgb.cpu().numpy(), flow.cpu().numpy(), return_all=True)
    report = {}
    for prec in ("fp16", "bf16", "fp32", "fp16x3"):
        out = model.infer(rgb, flow, want_probs=False, want_logits=True, precision=prec, chunk_T=256)
        torch.cuda.synchronize()
        assert model.device_error() == 0
        r = _agreement(out["labels"].cpu().numpy(), out["logits"].cpu().numpy(), ref_logits, ref_probs)
        report[prec] = r
        print(f"[{prec} {r['frames']} frames] rel logit err {r['rel']:.2e}; raw label agreement {r['raw']:.5f} ({r['flips']} flips); "
              f"{r['near_tie_frames']} near-tie frames (margin < {r['eps']:.2e}); away from near-ties {r['carved']:.6f} "
              f"({r['flips_away_from_near_ties']} flips)")
        assert r["rel"] <= REL[prec]
        assert r["flips_away_from_near_ties"] == 0
    assert report["fp32"]["raw"] >= 0.9999 and report["fp16x3"]["raw"] >= 0.9999
    assert report["fp16"]["raw"] >= 0.999, "north_star: >= 99.9 % identical labels on the default path"
    assert report["bf16"]["raw"] >= 0.99


def test_bench_shape_vs_oracle(dev):
    """BASELINE configs[2] as bench.py runs it: 4 096 streams, two 64-frame steps with the GRU state carried between
    them, fp16 operands; 256 sampled streams re-computed by the fp32 oracle over the whole 128 frames."""
    from prego_b200 import synthetic
    cfg = dict(synthetic.ASSEMBLY101_O)
    model = synthetic.seeded_model(cfg, seed=20, device=dev)
    B, Tc = 4096, 64
    h = torch.zeros(B, 1024, device=dev)
    steps = [synthetic.device_features(B, Tc, dev, seed=1234 + i) for i in range(2)]
    outs = [model.infer(r, f, h_state=h, want_probs=False, want_logits=True, precision="fp16", chunk_T=Tc) for r, f in steps]
    torch.cuda.synchronize()
    assert model.device_error() == 0
    pick = torch.randperm(B, generator=torch.Generator().manual_seed(3))[:256].sort().values
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    rgb = torch.cat([r[pick.to(dev)] for r, _ in steps], 1).cpu().numpy()
    flow = torch.cat([f[pick.to(dev)] for _, f in steps], 1).cpu().numpy()
    ref_probs, ref_logits, ref_h = miniroad_np.forward(sd, rgb, flow, return_all=True)
    logits = torch.cat([o["logits"][pick.to(dev)] for o in outs], 1).cpu().numpy()
    labels = torch.cat([o["labels"][pick.to(dev)] for o in outs], 1).cpu().numpy()
    r = _agreement(labels, logits, ref_logits, ref_probs)
    print(f"[fp16 bench shape, 256 of 4096 streams x 128 frames] rel logit err {r['rel']:.2e}; raw label agreement {r['raw']:.5f} "
          f"({r['flips']} flips, {r['flips_away_from_near_ties']} away from near-ties)")
    assert r["rel"] <= REL["fp16"] and r["flips_away_from_near_ties"] == 0 and r["raw"] >= 0.999
    assert np.abs(h[pick.to(dev)].cpu().numpy() - ref_h).max() <= 2 * REL["fp16"]
