"""Times prego_perframe_ap (csrc/metrics.cuh) with CUDA events: N frames x K classes of random probabilities.
  python scripts/ap_bench.py [N] [K]
Algorithmic bytes: build 4 B read + 4 B write, sort 4 passes x (4 B hist read + 4 B read + 4 B write), scan 8 B  = 64 B per (frame, class)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from prego_b200.metrics import average_precision_per_class

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 86
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
scores = torch.softmax(torch.randn(N, K, generator=g, device=dev) * 3, -1)
labels = torch.randint(0, K, (N,), generator=g, device=dev, dtype=torch.int32)
for _ in range(2):
    average_precision_per_class(scores, labels)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
reps = 5
for _ in range(reps):
    ap, npos = average_precision_per_class(scores, labels)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"perframe_ap N={N} K={K}: {ms:.3f} ms per call, {N * K * 64 / ms / 1e6:.1f} GB/s algorithmic, mean AP {np.nanmean(ap[1:]):.6f}")
t = time.perf_counter()
from sklearn.metrics import average_precision_score  # host baseline (what the reference calls), on a few classes
s_h, l_h = scores[:, :4].cpu().numpy(), labels.cpu().numpy()
for k in range(1, 4):
    average_precision_score(l_h == k, s_h[:, k])
print(f"sklearn on the host: {(time.perf_counter() - t) / 3 * 1e3:.1f} ms per class ({K - 1} classes per call in the reference)")
