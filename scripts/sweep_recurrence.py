"""Per-phase device time vs number of streams (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prego_b200 import synthetic

dev = torch.device("cuda:0")
model = synthetic.seeded_model(dict(synthetic.ASSEMBLY101_O), seed=20, device=dev)
T = 16
for B in (128, 256, 512, 1024, 2048, 4096, 8192):
    rgb, flow = synthetic.device_features(B, T, dev, seed=1)
    h = torch.zeros(B, 1024, device=dev)
    for _ in range(2):
        model.infer(rgb, flow, h_state=h, want_probs=False, precision="fp16", chunk_T=T)
    torch.cuda.synchronize()
    model.profile_begin()
    for _ in range(4):
        model.infer(rgb, flow, h_state=h, want_probs=False, precision="fp16", chunk_T=T)
    p = model.profile_end()
    steps = 4 * T
    print(f"B={B:5d}  rec {p['recurrence']['ms'] / steps * 1e3:7.1f} us/step   gemm1 {p['gemm1']['ms'] / 4:7.3f} ms  gemm2 {p['gemm2']['ms'] / 4:7.3f} ms  "
          f"stage {p['stage']['ms'] / 4:6.3f}  ln {p['layernorm']['ms'] / 4:6.3f}  head {p['head']['ms'] / 4:6.3f}", flush=True)
    del rgb, flow
