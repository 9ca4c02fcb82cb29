"""Training step on the CUDA path (BASELINE configs[4]): gradients of all ten tensors
against the reference-made golden gradients and against autograd through stock torch layers."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle.miniroad_torch_cpu import TorchRefMROAD, oad_loss
from prego_b200 import OadLoss, synthetic, train_one_step

pytestmark = pytest.mark.gpu
GRAD_REL = 2e-4  # exact-fp32 kernels vs ATen autograd: relative to max|grad| of the tensor
# The ReLU after LayerNorm is a hard gate: an activation within fp32 rounding of 0 (|yn| ~ 1e-7) can gate
# differently in two correct fp32 implementations (observed: 1 element of 4.2 M at B=16, T=128), which
# moves layer1 gradients by up to ~1e-2 * max in one row.  Large cases are therefore judged in the
# Frobenius norm, with a loose max-norm bound; small cases (no tie) element-wise at GRAD_REL.
FRO_REL, MAX_REL_LOOSE = 2e-3, 5e-2


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _grads_cuda(cfg, rgb, flow, target, dev):
    cfg = dict(cfg, train_precision="fp32")  # the exact CUDA-core mode: the gradient-parity bounds below are its bounds
    model = synthetic.seeded_model(cfg, seed=20, device=dev).train()
    out = model(rgb.to(dev), flow.to(dev))
    assert out["logits"].requires_grad and tuple(out["logits"].shape) == (rgb.shape[0], rgb.shape[1], cfg["num_classes"])
    loss = OadLoss(cfg)(out, target.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    return model, float(loss), {k: p.grad.detach().cpu() for k, p in model.named_parameters()}, out["logits"].detach().cpu()


def test_gradients_match_reference_golden(dev, golden_meta):
    gold = np.load(os.path.join(GOLD, "train_epic_b3_t10.npz"))
    c = golden_meta["train_case"]
    cfg = dict(synthetic.EPIC_TENT_O, dropout=0.0)
    rgb, flow = synthetic.feature_batch(c["stream_ids"], c["T"], "cpu", False)
    target = torch.stack([synthetic.targets(s, c["T"], 12) for s in c["stream_ids"]])
    _, loss, grads, _ = _grads_cuda(cfg, rgb, flow, target, dev)
    assert abs(loss - float(gold["loss"])) <= 1e-5
    for k, g in grads.items():
        g = g.reshape(-1)
        scale = float(gold[k + ".stats"][3])
        assert np.abs(g[:64].numpy() - gold[k + ".head"]).max() <= GRAD_REL * scale, k
        assert abs(g.abs().sum().item() - gold[k + ".stats"][1]) <= 1e-3 * gold[k + ".stats"][1], k
        assert abs(g.norm().item() - gold[k + ".stats"][2]) <= 1e-3 * gold[k + ".stats"][2], k


@pytest.mark.parametrize("B,T,K", [(5, 17, 86), (12, 9, 86), (16, 128, 86)])
def test_gradients_match_torch_autograd(dev, B, T, K):
    cfg = dict(synthetic.ASSEMBLY101_O, dropout=0.0, num_classes=K, train_precision="fp32")
    M_rows = B * T
    rgb, flow = synthetic.feature_batch(list(range(100, 100 + B)), T, "cpu", False)
    # a loss that touches EVERY frame (the reference loss only touches the last one)
    wts = torch.randn(B, T, K, generator=torch.Generator().manual_seed(1))
    model = synthetic.seeded_model(cfg, seed=20, device=dev).train()
    logits = model(rgb.to(dev), flow.to(dev))["logits"]
    (logits * wts.to(dev)).sum().backward()
    torch.cuda.synchronize()
    port = TorchRefMROAD(4096, 2048, 1024, K, 0.0).train()
    port.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    ref_logits = port(rgb, flow)["logits"]
    (ref_logits * wts).sum().backward()
    assert (logits.detach().cpu() - ref_logits.detach()).abs().max().item() <= 1e-4 * ref_logits.abs().max().item()
    for (k, p), (_, q) in zip(model.named_parameters(), port.named_parameters()):
        d = p.grad.cpu() - q.grad
        err, fro = d.abs().max().item(), d.norm().item() / q.grad.norm().item()
        if M_rows <= 256:
            assert err <= GRAD_REL * q.grad.abs().max().item(), f"{k}: {err} vs max {q.grad.abs().max().item()}"
        assert fro <= FRO_REL and err <= MAX_REL_LOOSE * q.grad.abs().max().item(), f"{k}: fro {fro}, max {err}"


@pytest.mark.parametrize("B,T", [(16, 128), (128, 12)])
def test_gradients_tf32_mode(dev, B, T):
    """train_precision = 'tf32': the large projections and their weight / input gradients on tcgen05 kind::tf32
    (10-bit-mantissa operands, fp32 accumulate).  Judged in the Frobenius norm against ATen fp32 autograd at a TF32-class
    tolerance; the exact mode above is the parity mode.  Measured on B200 (scripts/diag_train_tf32.py): logits 5.9e-4
    relative; every gradient that flows through the 128-step recurrence 1.8e-2 .. 2.1e-2 -- including the bias
    gradients, whose own reductions are exact: the BPTT chain amplifies the 6e-4 forward perturbation, as it does for
    any TF32 forward -- and 5.9e-4 for the classifier weight, which does not.  (128, 12): the wide-batch variant, whose
    per-step recurrent GEMMs run on the tensor cores as well."""
    K = 86
    cfg = dict(synthetic.ASSEMBLY101_O, dropout=0.0, num_classes=K, train_precision="tf32")
    rgb, flow = synthetic.feature_batch(list(range(100, 100 + B)), T, "cpu", False)
    wts = torch.randn(B, T, K, generator=torch.Generator().manual_seed(1))
    model = synthetic.seeded_model(cfg, seed=20, device=dev).train()
    assert model.train_precision == "tf32"
    logits = model(rgb.to(dev), flow.to(dev))["logits"]
    (logits * wts.to(dev)).sum().backward()
    torch.cuda.synchronize()
    port = TorchRefMROAD(4096, 2048, 1024, K, 0.0).train()
    port.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    ref_logits = port(rgb, flow)["logits"]
    (ref_logits * wts).sum().backward()
    rel = (logits.detach().cpu() - ref_logits.detach()).abs().max().item() / ref_logits.abs().max().item()
    assert 1e-7 < rel <= 3e-3, rel  # > 0: the tensor-core path really ran
    for (k, p), (_, q) in zip(model.named_parameters(), port.named_parameters()):
        fro = (p.grad.cpu() - q.grad).norm().item() / q.grad.norm().item()
        assert fro <= (5e-2 if not k.startswith("f_classification") else 3e-3), f"{k}: fro {fro}"
    assert model.device_error() == 0


@pytest.mark.parametrize("B,T", [(16, 128), (128, 12)])
def test_gradients_tf32x3_mode(dev, B, T):
    """train_precision = 'tf32x3': the same tensor-core products with every operand split into TF32 hi + lo, the three
    leading terms contracted per 1 024-column chunk (small terms first) and the chunks added in fp32.  Forward: fp32-class
    (logits 2.3e-6 of ATen's on the CPU; stock torch fp32 on the same GPU: 1.6e-6; plain TF32: 5.9e-4).  Gradients that pass
    the hard gates of this graph (relu(h_t) in front of the classifier, ReLU after the LayerNorm) differ by the few gates
    that flip for values within the forward error of 0, so their error goes with the square root of the forward error:
    measured 2.2e-3 .. 2.8e-3 (Frobenius) for a 2e-6 forward error, 1.9e-2 for plain TF32's 6e-4, 1e-6 only when the
    forward agrees to rounding (exact mode).  Table: profiles/r02_train_modes.txt (scripts/diag_train_modes.py).
    (128, 12): per-step recurrent products split as well."""
    K = 86
    cfg = dict(synthetic.ASSEMBLY101_O, dropout=0.0, num_classes=K, train_precision="tf32x3")
    rgb, flow = synthetic.feature_batch(list(range(100, 100 + B)), T, "cpu", False)
    wts = torch.randn(B, T, K, generator=torch.Generator().manual_seed(1))
    model = synthetic.seeded_model(cfg, seed=20, device=dev).train()
    logits = model(rgb.to(dev), flow.to(dev))["logits"]
    (logits * wts.to(dev)).sum().backward()
    torch.cuda.synchronize()
    port = TorchRefMROAD(4096, 2048, 1024, K, 0.0).train()
    port.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    ref_logits = port(rgb, flow)["logits"]
    (ref_logits * wts).sum().backward()
    rel = (logits.detach().cpu() - ref_logits.detach()).abs().max().item() / ref_logits.abs().max().item()
    worst = 0.0
    for (k, p), (_, q) in zip(model.named_parameters(), port.named_parameters()):
        d = p.grad.cpu() - q.grad
        fro, err = d.norm().item() / q.grad.norm().item(), d.abs().max().item() / q.grad.abs().max().item()
        worst = max(worst, fro)
        assert fro <= (6e-3 if not k.startswith("f_classification") else 1e-5) and err <= MAX_REL_LOOSE, f"{k}: fro {fro}, max {err}"
    print(f"[tf32x3 B={B} T={T}] logits rel {rel:.2e}, worst gradient Frobenius error {worst:.2e}")
    assert 1e-8 < rel <= 1e-5, rel
    assert model.device_error() == 0


def test_dropout_mask_and_training_loop(dev):
    cfg = dict(synthetic.EPIC_TENT_O, dropout=0.2)
    B, T, K = 8, 32, 12
    rgb, flow = synthetic.feature_batch(list(range(B)), T, dev, True)  # zero flow, as the reference loader feeds it
    target = torch.stack([synthetic.targets(s, T, K) for s in range(B)]).to(dev)
    model = synthetic.seeded_model(cfg, seed=20, device=dev).train()
    torch.manual_seed(1)
    a = model(rgb, flow)["logits"].detach()
    torch.manual_seed(1)
    b = model(rgb, flow)["logits"].detach()
    c = model(rgb, flow)["logits"].detach()
    assert torch.equal(a, b) and not torch.equal(a, c)      # mask stream follows torch's RNG state
    model.eval()
    p = model(rgb, flow)["logits"]
    assert (p.sum(-1) - 1).abs().max().item() < 1e-5          # eval path unaffected
    crit = OadLoss(cfg)
    opt = torch.optim.AdamW([{"params": model.parameters(), "initial_lr": 1e-3}], lr=1e-3, weight_decay=0.05)
    losses = [float(train_one_step(model, crit, opt, rgb, flow, target)) for _ in range(12)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    # the same loop on the one-launch AdamW follows the same trajectory (the packed weights must track the raw-pointer updates)
    from prego_b200 import FusedAdamW
    model2 = synthetic.seeded_model(cfg, seed=20, device=dev).train()
    opt2 = FusedAdamW([{"params": list(model2.parameters()), "initial_lr": 1e-3}], lr=1e-3, weight_decay=0.05)
    torch.manual_seed(7)
    l2 = [float(train_one_step(model2, crit, opt2, rgb, flow, target)) for _ in range(12)]
    model3 = synthetic.seeded_model(cfg, seed=20, device=dev).train()
    opt3 = torch.optim.AdamW([{"params": model3.parameters(), "initial_lr": 1e-3}], lr=1e-3, weight_decay=0.05)
    torch.manual_seed(7)
    l3 = [float(train_one_step(model3, crit, opt3, rgb, flow, target)) for _ in range(12)]
    assert l2[-1] < l2[0] and max(abs(a - b) for a, b in zip(l2, l3)) < 5e-3 * max(l3), (l2, l3)


def test_training_packs_the_fp32_set_only_and_inference_repacks(dev):
    """A training step re-packs only the fp32 operand set (prego_model_load_weights_ex(PREGO_PACK_F32)); the 16-bit and split
    copies go stale, the C ABI refuses to run on a stale format, and the next inference call of the module re-packs what it
    needs -- its result must follow the UPDATED weights (checked against the numpy oracle on those weights)."""
    import ctypes as C

    from oracle import miniroad_np
    from prego_b200 import _lib
    cfg = dict(synthetic.EPIC_TENT_O, dropout=0.0)
    B, T, K = 8, 16, 12
    rgb, flow = synthetic.feature_batch(list(range(B)), T, dev, False)
    target = torch.stack([synthetic.targets(s, T, K) for s in range(B)]).to(dev)
    model = synthetic.seeded_model(cfg, seed=20, device=dev)
    before = model.infer(rgb, flow, want_probs=False, want_logits=True, precision="fp16")["logits"].clone()
    assert model._packed_formats == _lib.PACK_ALL
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=0.0)
    model.train()
    train_one_step(model, OadLoss(cfg), opt, rgb, flow, target)
    train_one_step(model, OadLoss(cfg), opt, rgb, flow, target)   # second step: packs the weights the first step wrote
    assert model._packed_formats == _lib.PACK_F32
    lib = _lib.load()
    ws = torch.empty(lib.prego_workspace_bytes(model._handle, B, T, _lib.PREC_F16) + 1024, dtype=torch.uint8, device=dev)
    labels = torch.empty(B, T, dtype=torch.int32, device=dev)
    args = _lib.ForwardArgs()
    args.rgb, args.flow, args.B, args.T, args.chunk_T = rgb.data_ptr(), flow.data_ptr(), B, T, T
    args.labels = labels.data_ptr()
    args.workspace, args.workspace_bytes = ws.data_ptr() + (-ws.data_ptr()) % 1024, ws.numel() - 1024
    args.precision = _lib.PREC_F16
    rc = lib.prego_forward(model._handle, C.byref(args), torch.cuda.current_stream(dev).cuda_stream)
    assert rc == 4, (rc, lib.prego_last_error())  # PREGO_ERR_STATE
    model.eval()
    out = model.infer(rgb, flow, want_probs=False, want_logits=True, precision="fp16")
    torch.cuda.synchronize()
    assert model._packed_formats == _lib.PACK_ALL
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    _, ref_logits, _ = miniroad_np.forward(sd, rgb.cpu().numpy(), flow.cpu().numpy(), return_all=True)
    scale = float(np.abs(ref_logits).max())
    assert np.abs(out["logits"].cpu().numpy() - ref_logits).max() <= 2e-3 * scale
    assert (out["logits"] - before).abs().max().item() > 1e-2 * scale, "the update must be visible"


def test_fused_adamw_matches_torch(dev):
    """prego_adamw_step (one launch for all tensors) against torch.optim.AdamW over 5 steps, ragged tensor sizes."""
    from prego_b200 import FusedAdamW
    g = torch.Generator().manual_seed(3)
    shapes = [(3072, 64), (2048,), (86, 1024), (1,), (7, 3)]
    pa = [torch.randn(s, generator=g).to(dev).requires_grad_() for s in shapes]
    pb = [p.detach().clone().requires_grad_() for p in pa]
    oa = FusedAdamW(pa, lr=1e-3, weight_decay=0.05)
    ob = torch.optim.AdamW(pb, lr=1e-3, weight_decay=0.05)
    for step in range(5):
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, generator=g).to(dev)
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    for x, y in zip(pa, pb):
        assert (x - y).abs().max().item() <= 2e-6 * max(1.0, y.abs().max().item())
    # grad_scale: a SUM all-reduced gradient of world 4 scaled on the fly equals stepping on the mean
    for x, y in zip(pa, pb):
        gr = torch.randn(x.shape, generator=g).to(dev)
        x.grad, y.grad = gr * 4, gr.clone()
    v0 = pa[0]._version
    oa.step(grad_scale=0.25)
    ob.step()
    assert pa[0]._version > v0   # in-place update is visible to autograd and to MROAD's weight re-pack check
    for x, y in zip(pa, pb):
        assert (x - y).abs().max().item() <= 2e-6 * max(1.0, y.abs().max().item())


def test_main_training_entry_synthetic(dev, tmp_path):
    """python -m prego_b200.main --config ... (no --eval): the reference's training loop (main.py:59-115) end to end on
    synthetic videos: one epoch, evaluation, best checkpoint written and renamed, loss finite, weights changed."""
    import glob
    import yaml
    from prego_b200 import main as pmain
    cfg = dict(synthetic.EPIC_TENT_O, window_size=32, stride=16, batch_size=4, test_batch_size=1, num_epoch=1, lr=1e-4, weight_decay=0.05,
               optimizer="AdamW", loss="NONUNIFORM", num_workers=0, video_list_path="", root_path="", annotation_type="target_perframe")
    cfg_path = tmp_path / "cfg.yaml"
    yaml.safe_dump(cfg, open(cfg_path, "w"))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        mAP = pmain.main(["--config", str(cfg_path), "--synthetic", "2", "--device", "cuda:0", "--output_path", str(tmp_path / "out")])
    finally:
        os.chdir(cwd)
    ck = glob.glob(str(tmp_path / "out" / "ckpts" / "best_*.pth"))
    assert len(ck) == 1 and 0.0 <= mAP <= 1.0
    sd = torch.load(ck[0], map_location="cpu")
    ref = synthetic.seeded_model(dict(synthetic.EPIC_TENT_O), seed=20, device="cpu").state_dict()
    assert set(sd) == set(ref) and all(not torch.equal(sd[k], ref[k]) for k in sd)
    assert all(torch.isfinite(v).all() for v in sd.values())


def test_two_forwards_before_backward_keep_their_own_activations(dev):
    """Gradient accumulation / a loss over several batches: a second train-mode forward -- larger, so it
    would also have re-allocated a shared workspace -- before the first backward must not disturb the first one's
    saved activations.  The accumulated gradients must equal the sum of the two separate runs."""
    cfg = dict(synthetic.EPIC_TENT_O, dropout=0.0)
    crit = OadLoss(cfg)
    r1, f1 = synthetic.feature_batch([70, 71], 9, "cpu", False)
    r2, f2 = synthetic.feature_batch([72, 73, 74, 75, 76], 23, "cpu", False)
    t1 = torch.stack([synthetic.targets(s, 9, 12) for s in (70, 71)])
    t2 = torch.stack([synthetic.targets(s, 23, 12) for s in (72, 73, 74, 75, 76)])
    _, _, g1, _ = _grads_cuda(cfg, r1, f1, t1, dev)
    _, _, g2, _ = _grads_cuda(cfg, r2, f2, t2, dev)
    model = synthetic.seeded_model(dict(cfg, train_precision="fp32"), seed=20, device=dev).train()
    l1 = crit(model(r1.to(dev), f1.to(dev)), t1.to(dev))
    l2 = crit(model(r2.to(dev), f2.to(dev)), t2.to(dev))   # second forward BEFORE the first backward
    (l1 + l2).backward()
    torch.cuda.synchronize()
    for k, p in model.named_parameters():
        want = g1[k] + g2[k]
        assert (p.grad.cpu() - want).abs().max().item() <= 1e-5 * max(want.abs().max().item(), 1e-6), k


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (runs under `gpurun --gpus 2`; the single-GPU box skips it)")
def test_overlapped_allreduce_equals_mean_of_local_gradients_nccl():
    """scripts/ddp_check.py under torchrun on 2 GPUs: the bucketed all-reduce that runs inside the CUDA backward gives every
    rank exactly the mean of the ranks' local gradients (evidence of the round: worst relative difference 0.0)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(root, "scripts", "ddp_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and '"overlapped_allreduce_equals_mean_of_local_gradients": true' in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
