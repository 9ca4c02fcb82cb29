#!/usr/bin/env python
"""Kernel-level breakdown of one training step (torch profiler, CUDA activities)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prego_b200 import OadLoss, synthetic, train_one_step  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = 128
cfg = dict(synthetic.ASSEMBLY101_O)
model = synthetic.seeded_model(cfg, seed=20, device=dev)
crit = OadLoss(cfg)
opt = torch.optim.AdamW([{"params": model.parameters(), "initial_lr": 1e-4}], lr=1e-4, weight_decay=0.05)
rgb, flow = synthetic.device_features(B, T, dev, seed=7, zero_flow=True)
target = torch.nn.functional.one_hot(torch.randint(0, 86, (B, T), device=dev), 86).float()
for _ in range(3):
    train_one_step(model, crit, opt, rgb, flow, target)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        train_one_step(model, crit, opt, rgb, flow, target)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
