#!/usr/bin/env python
"""Phase breakdown of the fused per-frame kernel from its SM-clock stamps (PREGO_ONLINE_TRACE=1)."""
import os
import sys

os.environ["PREGO_ONLINE_TRACE"] = "1"
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prego_b200 import synthetic  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
cfg = dict(synthetic.EPIC_TENT_O)
model = synthetic.seeded_model(cfg, seed=20, device=dev)
rgb, flow = synthetic.device_features(1, 64, dev, seed=3)
sess = model.online_session(1, dev, sys.argv[1] if len(sys.argv) > 1 else "fp16")
names = ["weights req + x/h slice staged", "phase A math + reduce + LN partials", "grid barrier", "LayerNorm", "phase B math + gates",
         "logit partials", "arrive (fence + atomic)", "finalize (last CTA)"]
acc = []
for t in range(64):
    sess.step(rgb[0, t].contiguous(), flow[0, t].contiguous())
    torch.cuda.synchronize()
    if t >= 16:
        acc.append(sess.trace())
tr = np.stack(acc).astype(np.float64)  # [frames, ctas, 16]
mhz = 1965.0
d = np.diff(tr[:, :, :9], axis=2) / mhz  # us
print("phase (us)                      CTA0-median   all-CTA median   all-CTA max(median over frames)")
for i, nm in enumerate(names):
    col = d[:, :, i]
    valid = (tr[:, :, i + 1] > tr[:, :, i]) & (tr[:, :, i] > 0)
    col = np.where(valid, col, np.nan)
    c0 = np.nanmedian(col[:, 0])
    allm = np.nanmedian(col)
    mx = np.nanmedian(np.nanmax(col, axis=1))
    print(f"{nm:38s} {c0:10.2f} {allm:14.2f} {mx:14.2f}")
span = (tr[:, :, 10].max(axis=1) - tr[:, :, 9].min(axis=1)) / 1e3
skew = (tr[:, :, 9].max(axis=1) - tr[:, :, 9].min(axis=1)) / 1e3
print("globaltimer: first CTA entry -> last exit, us (median):", np.median(span), " entry skew across CTAs:", np.median(skew))
