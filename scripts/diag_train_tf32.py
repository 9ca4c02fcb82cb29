#!/usr/bin/env python
"""Per-tensor gradient error of the tf32 training mode (and the exact mode) vs ATen fp32 autograd on CPU."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.miniroad_torch_cpu import TorchRefMROAD  # noqa: E402
from prego_b200 import synthetic  # noqa: E402

dev = torch.device("cuda:0")
B, T, K = 16, 128, 86
rgb, flow = synthetic.feature_batch(list(range(100, 100 + B)), T, "cpu", False)
wts = torch.randn(B, T, K, generator=torch.Generator().manual_seed(1))
ref = None
for prec in ("fp32", "tf32"):
    cfg = dict(synthetic.ASSEMBLY101_O, dropout=0.0, num_classes=K, train_precision=prec)
    model = synthetic.seeded_model(cfg, seed=20, device=dev).train()
    logits = model(rgb.to(dev), flow.to(dev))["logits"]
    (logits * wts.to(dev)).sum().backward()
    torch.cuda.synchronize()
    if ref is None:
        port = TorchRefMROAD(4096, 2048, 1024, K, 0.0).train()
        port.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
        ref_logits = port(rgb, flow)["logits"]
        (ref_logits * wts).sum().backward()
        ref = {k: q.grad.clone() for k, q in port.named_parameters()}
    print(prec, "logits rel", ((logits.detach().cpu() - ref_logits.detach()).abs().max() / ref_logits.abs().max()).item())
    for k, p in model.named_parameters():
        d = p.grad.cpu() - ref[k]
        print(f"  {k:32s} fro {d.norm().item() / ref[k].norm().item():.3e}  max {d.abs().max().item() / ref[k].abs().max().item():.3e}")
