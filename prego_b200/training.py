"""Training step on the CUDA path: autograd bridge, criterion mirror and the data-parallel gradient all-reduce.

Reference: ``trainer/train.py:5-29`` (``out = model(rgb, flow); loss = criterion(out, target);
optimizer.zero_grad(); loss.backward(); optimizer.step()``), ``criterions/loss.py:6-37`` (``OadLoss``,
registered 'NONUNIFORM'), ``main.py:62-67`` (AdamW).  The model's train-mode forward and the whole BPTT
backward run in ``libprego_b200.so`` (``prego_train_forward`` / ``prego_train_backward``); the tiny [B, K]
criterion and the optimizer stay stock torch, exactly as a user of the reference would keep them.
Multi-GPU: one process per GPU, streams (windows) sharded by rank, gradients summed with one NCCL
all-reduce of the flat 17.9 M-float buffer (``allreduce_gradients``) -- the only collective of this path.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .registry import Registry

CRITERIONS = Registry()


@CRITERIONS.register("NONUNIFORM")
class OadLoss(nn.Module):
    """criterions/loss.py:6-37: CE on the LAST frame only, target L2-normalised, mean over the batch."""

    def __init__(self, cfg, reduction="mean"):
        super().__init__()
        self.reduction = reduction
        self.num_classes = cfg["num_classes"]

    def forward(self, out_dict, target):
        logits = out_dict["logits"][:, -1, :].contiguous()
        target = target[:, -1, :].contiguous()
        output = torch.sum(-F.normalize(target) * F.log_softmax(logits, dim=-1), dim=1)
        return output.mean() if self.reduction == "mean" else output.sum()


def build_criterion(cfg, device=None):
    return CRITERIONS[cfg["loss"]](cfg).to(device)


class MiniROADTrainFn(torch.autograd.Function):
    """logits = f(rgb, flow; ten parameters) with the forward and backward in the C-ABI library."""

    @staticmethod
    def forward(ctx, module, rgb, flow, seed, *params):
        lib = _lib.load()
        ref = rgb if module.use_rgb else flow
        device = ref.device
        B, T = int(ref.shape[0]), int(ref.shape[1])
        with torch.cuda.device(device):
            module._ensure_handle(device)
            module._sync_weights(lib, device)
            need = lib.prego_train_workspace_bytes(module._handle, B, T)
            ws = module._train_ws
            if ws is None or ws.device != device or ws.numel() < need + 1024:
                module._train_ws = None
                ws = module._train_ws = torch.empty(need + 1024, dtype=torch.uint8, device=device)
            ws_ptr = ws.data_ptr() + (-ws.data_ptr()) % 1024
            rgb_c = rgb.contiguous() if module.use_rgb else None
            flow_c = flow.contiguous() if module.use_flow else None
            logits = torch.empty(B, T, module.out_dim, dtype=torch.float32, device=device)
            args = _lib.TrainArgs(rgb_c.data_ptr() if rgb_c is not None else None,
                                  flow_c.data_ptr() if flow_c is not None else None, B, T, logits.data_ptr(), None, None,
                                  ws_ptr, need, float(module.layer1[3].p), int(seed),
                                  _lib.TRAIN_PRECISIONS[getattr(module, "train_precision", "fp32")])
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(lib.prego_train_forward(module._handle, C.byref(args), stream), "prego_train_forward")
        ctx.module, ctx.rgb, ctx.flow, ctx.seed = module, rgb_c, flow_c, int(seed)
        ctx.prec = args.precision
        ctx.ws_ptr, ctx.need, ctx.B, ctx.T = ws_ptr, need, B, T
        ctx.shapes = [tuple(p.shape) for p in params]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = _lib.load()
        module = ctx.module
        device = dlogits.device
        dlogits = dlogits.contiguous().float()
        with torch.cuda.device(device):
            grads = [torch.empty(s, dtype=torch.float32, device=device) for s in ctx.shapes]
            g = _lib.Grads(*[t.data_ptr() for t in grads])
            args = _lib.TrainArgs(ctx.rgb.data_ptr() if ctx.rgb is not None else None,
                                  ctx.flow.data_ptr() if ctx.flow is not None else None, ctx.B, ctx.T, None,
                                  dlogits.data_ptr(), C.pointer(g), ctx.ws_ptr, ctx.need, float(module.layer1[3].p), ctx.seed,
                                  ctx.prec)
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(lib.prego_train_backward(module._handle, C.byref(args), stream), "prego_train_backward")
        return (None, None, None, None, *grads)


def train_forward(module, rgb, flow):
    """Train-mode forward of ``prego_b200.MROAD`` (raw logits [B, T, K], dropout active, autograd-tracked)."""
    ref = rgb if module.use_rgb else flow
    if not isinstance(ref, torch.Tensor) or not ref.is_cuda:
        raise RuntimeError("prego_b200.MROAD runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # drawn from torch's RNG: reproducible under manual_seed
    return MiniROADTrainFn.apply(module, rgb, flow, seed, *module._param_tensors())


def allreduce_gradients(module, group=None):
    """Data-parallel gradient averaging: ONE all-reduce of the flat gradient buffer (17.9 M floats for
    K = 86) over NCCL / NVLink, then scatter back.  No-op for a single process."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def train_one_step(model, criterion, optimizer, rgb, flow, target, group=None):
    """One iteration of trainer/train.py:8-24 (+ the gradient all-reduce when run data-parallel)."""
    model.train()
    out = model(rgb, flow)
    loss = criterion(out, target)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    allreduce_gradients(model, group)
    optimizer.step()
    return loss.detach()
