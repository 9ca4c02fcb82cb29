#!/usr/bin/env python
"""Training-step leg of bench.py on its own (BASELINE configs[4])."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    print(json.dumps(bench.training_leg(dev, 1)))
