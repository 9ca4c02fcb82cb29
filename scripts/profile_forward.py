"""Small driver for ncu captures: one warm-up forward, then one profiled forward (B streams x T frames)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prego_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfgname = sys.argv[3] if len(sys.argv) > 3 else "ASSEMBLY101_O"
dev = torch.device("cuda:0")
model = synthetic.seeded_model(dict(getattr(synthetic, cfgname)), seed=20, device=dev)
rgb, flow = synthetic.device_features(B, T, dev, seed=1)
h = torch.zeros(B, 1024, device=dev)
for _ in range(2):
    model.infer(rgb, flow, h_state=h, want_probs=False, precision="fp16", chunk_T=T)
torch.cuda.synchronize()
print("done")
