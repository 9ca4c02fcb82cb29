import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.miniroad_torch_cpu import TorchRefMROAD
from prego_b200 import synthetic
dev = torch.device("cuda:0")
for (B, T, K) in [(16, 32, 86), (4, 128, 86), (16, 128, 86)]:
    cfg = dict(synthetic.ASSEMBLY101_O, dropout=0.0, num_classes=K)
    rgb, flow = synthetic.feature_batch(list(range(100, 100 + B)), T, "cpu", False)
    wts = torch.randn(B, T, K, generator=torch.Generator().manual_seed(1))
    model = synthetic.seeded_model(cfg, seed=20, device=dev).train()
    logits = model(rgb.to(dev), flow.to(dev))["logits"]
    (logits * wts.to(dev)).sum().backward()
    torch.cuda.synchronize()
    res = {}
    for dt in (torch.float32, torch.float64):
        port = TorchRefMROAD(4096, 2048, 1024, K, 0.0).train().to(dt)
        port.load_state_dict({k: v.cpu().to(dt) for k, v in model.state_dict().items()})
        ref_logits = port(rgb.to(dt), flow.to(dt))["logits"]
        (ref_logits * wts.to(dt)).sum().backward()
        res[dt] = {k: q.grad.double() for k, q in port.named_parameters()}
    print(f"--- B={B} T={T}")
    for k, p in model.named_parameters():
        g = p.grad.cpu().double()
        r32, r64 = res[torch.float32][k], res[torch.float64][k]
        s = r64.abs().max().item()
        print(f"{k:28s} max|g| {s:9.3e}  ours-vs-f64 {(g - r64).abs().max().item() / s:8.2e}  aten32-vs-f64 {(r32 - r64).abs().max().item() / s:8.2e}")
