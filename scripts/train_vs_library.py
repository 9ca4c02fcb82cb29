"""This repo's training step against the stock-torch step on the same GPU, same process (the two legs bench.py reports as
`train_step` and `library_baseline.train_step`), without the rest of the bench:
    python scripts/train_vs_library.py > gpurun_out/train_vs_library.txt
PREGO_TRAIN_GROUP=4|8 selects the earlier generations of the persistent recurrence kernels for A/B runs."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.nn as nn

import bench


class Ref(nn.Module):
    """The reference's module structure on stock layers (rnn.py:38-47)."""

    def __init__(self):
        super().__init__()
        self.gru = nn.GRU(2048, 1024, 1, batch_first=True)
        self.layer1 = nn.Sequential(nn.Linear(4096, 2048), nn.LayerNorm(2048), nn.ReLU(), nn.Dropout(0.2))
        self.fc = nn.Linear(1024, 86)


dev = torch.device("cuda:0")
print(json.dumps(bench.library_train_step(dev, Ref), indent=0))
ours = bench.training_leg(dev, 1)
ours.pop("note")
print(json.dumps(ours, indent=0))
