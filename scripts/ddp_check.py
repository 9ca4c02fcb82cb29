#!/usr/bin/env python
"""Data-parallel training check under NCCL (torchrun, N >= 2): the bucketed all-reduce that runs INSIDE the CUDA backward
(prego_b200.training.enable_overlapped_allreduce: gru / classifier bucket on a side stream under the layer1 backward) must
give every rank exactly the mean of the ranks' local gradients.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/ddp_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prego_b200 import OadLoss, enable_overlapped_allreduce, synthetic  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
cfg = dict(synthetic.ASSEMBLY101_O, dropout=0.0)
model = synthetic.seeded_model(cfg, seed=20, device=dev).train()
crit = OadLoss(cfg)
B, T = 16, 128
rgb, flow = synthetic.device_features(B, T, dev, seed=100 + rank, zero_flow=True)
target = torch.nn.functional.one_hot(torch.randint(0, 86, (B, T), device=dev, generator=torch.Generator(device=dev).manual_seed(rank)), 86).float()


def grads(overlapped):
    model._dp_group = None
    if overlapped:
        assert enable_overlapped_allreduce(model)
    model.zero_grad(set_to_none=True)
    crit(model(rgb, flow), target).backward()
    torch.cuda.synchronize()
    return [p.grad.detach().clone() for p in model.parameters()]


local_g = grads(False)
red_g = grads(True)
worst = 0.0
for lg, rg in zip(local_g, red_g):
    parts = [torch.empty_like(lg) for _ in range(world)]
    dist.all_gather(parts, lg)
    mean = torch.stack(parts).mean(0)
    worst = max(worst, float((rg - mean).abs().max() / mean.abs().max().clamp_min(1e-12)))
ok = worst <= 1e-6
t = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(t)
if rank == 0:
    print(json.dumps({"world": world, "overlapped_allreduce_equals_mean_of_local_gradients": int(t) == 0, "worst_relative_difference": worst}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t) == 0 else 1)
