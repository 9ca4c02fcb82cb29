#!/usr/bin/env python
"""Single-stream latency leg of bench.py on its own (whole video + strict per-frame stepping).
Usage: python scripts/online_latency.py [fp16|bf16]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    print(json.dumps(bench.latency_leg(dev, prec)))
