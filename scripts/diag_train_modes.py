"""Gradient error of each training precision against ATen fp32 autograd on the CPU (B = 16, T = 128, loss over every frame):
relative logit error and, per tensor, the Frobenius-relative and max-relative gradient error.  Last block: the SAME stock
torch modules run in fp32 on the GPU (cuBLAS / cuDNN with TF32 off) against the CPU run -- how far two correct fp32
implementations are apart on this graph (hard ReLU gates on h_t and after the LayerNorm flip for values within rounding of 0)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from oracle.miniroad_torch_cpu import TorchRefMROAD
from prego_b200 import synthetic

dev = torch.device("cuda:0")
B, T, K = 16, 128, 86
rgb, flow = synthetic.feature_batch(list(range(100, 100 + B)), T, "cpu", False)
wts = torch.randn(B, T, K, generator=torch.Generator().manual_seed(1))
ref = None
for prec in ("fp32", "tf32x3", "tf32"):
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
        ref_logits = ref_logits.detach()
    rel = (logits.detach().cpu() - ref_logits).abs().max().item() / ref_logits.abs().max().item()
    print(f"{prec}: logits rel {rel:.2e}")
    for k, p in model.named_parameters():
        d = p.grad.cpu() - ref[k]
        print(f"   {k:28s} fro {d.norm().item() / ref[k].norm().item():.2e}   max {d.abs().max().item() / ref[k].abs().max().item():.2e}")

for label, cudnn_tf32 in (("stock torch fp32 on the GPU, TF32 off everywhere", False),
                          ("stock torch on the GPU with torch's DEFAULT flags (fp32 cuBLAS, cuDNN GRU free to use TF32) = the reference as shipped", True)):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = cudnn_tf32
    gport = TorchRefMROAD(4096, 2048, 1024, K, 0.0).to(dev).train()
    gport.load_state_dict(port.state_dict())
    glogits = gport(rgb.to(dev), flow.to(dev))["logits"]
    (glogits * wts.to(dev)).sum().backward()
    torch.cuda.synchronize()
    rel = (glogits.detach().cpu() - ref_logits).abs().max().item() / ref_logits.abs().max().item()
    print(f"{label} vs the CPU run: logits rel {rel:.2e}")
    for k, p in gport.named_parameters():
        d = p.grad.cpu() - ref[k]
        print(f"   {k:28s} fro {d.norm().item() / ref[k].norm().item():.2e}   max {d.abs().max().item() / ref[k].abs().max().item():.2e}")
