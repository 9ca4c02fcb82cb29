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
            module._sync_weights(lib, device, _lib.PACK_F32)  # the training entry points read the fp32 set only
            prec = _lib.TRAIN_PRECISIONS[getattr(module, "train_precision", "tf32x3")]
            need = lib.prego_train_workspace_bytes_ex(module._handle, B, T, prec)
            # the saved activations (gates, h_t, e, y, masks) belong to THIS forward: one workspace per call, kept alive
            # by ctx until its backward ran, so a second train-mode forward before the first backward (gradient
            # accumulation, two views, a loss over several batches) cannot overwrite or free them; torch's caching
            # allocator makes the steady-state cost of the per-call allocation nil
            ws = torch.empty(need + 1024, dtype=torch.uint8, device=device)
            ws_ptr = ws.data_ptr() + (-ws.data_ptr()) % 1024
            rgb_c = rgb.contiguous() if module.use_rgb else None
            flow_c = flow.contiguous() if module.use_flow else None
            logits = torch.empty(B, T, module.out_dim, dtype=torch.float32, device=device)
            args = _lib.TrainArgs(rgb_c.data_ptr() if rgb_c is not None else None,
                                  flow_c.data_ptr() if flow_c is not None else None, B, T, logits.data_ptr(), None, None,
                                  ws_ptr, need, float(module.layer1[3].p), int(seed),
                                  _lib.TRAIN_PRECISIONS[getattr(module, "train_precision", "tf32x3")])
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(lib.prego_train_forward(module._handle, C.byref(args), stream), "prego_train_forward")
        ctx.module, ctx.rgb, ctx.flow, ctx.seed = module, rgb_c, flow_c, int(seed)
        ctx.prec = args.precision
        ctx.ws, ctx.ws_ptr, ctx.need, ctx.B, ctx.T = ws, ws_ptr, need, B, T
        ctx.shapes = [tuple(p.shape) for p in params]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = _lib.load()
        module = ctx.module
        device = dlogits.device
        dlogits = dlogits.contiguous().float()
        group = getattr(module, "_dp_group", None)
        overlapped = group is not None
        with torch.cuda.device(device):
            # ten gradients as views of ONE flat buffer laid out in the order the backward finishes them
            # (gradient_buckets): a bucket is a contiguous slice, so its all-reduce needs no gather / scatter passes
            buckets, offsets, total = gradient_buckets(ctx.shapes)
            flat = torch.empty(total, dtype=torch.float32, device=device)
            grads = [flat[o:o + _numel(s)].view(s) for o, s in zip(offsets, ctx.shapes)]
            g = _lib.Grads(*[t.data_ptr() for t in grads])
            main = torch.cuda.current_stream(device)
            ev = torch.cuda.Event() if overlapped else None
            if ev is not None:
                ev.record(main)  # materialises the cudaEvent_t; re-recorded by the library at the bucket boundary
            args = _lib.TrainArgs(ctx.rgb.data_ptr() if ctx.rgb is not None else None,
                                  ctx.flow.data_ptr() if ctx.flow is not None else None, ctx.B, ctx.T, None,
                                  dlogits.data_ptr(), C.pointer(g), ctx.ws_ptr, ctx.need, float(module.layer1[3].p), ctx.seed,
                                  ctx.prec, 0, ev.cuda_event if ev is not None else None)
            _lib.check(lib.prego_train_backward(module._handle, C.byref(args), main.cuda_stream), "prego_train_backward")
            ctx.ws.record_stream(main)  # freed by autograd right after this returns
            if overlapped:
                import torch.distributed as dist
                world = dist.get_world_size(group)
                (a0, a1), (b0, b1) = buckets
                comm = _comm_stream(device)
                with torch.cuda.stream(comm):
                    comm.wait_event(ev)  # gru.* / f_classification.* gradients final: reduce them under the layer1 backward
                    w_a = dist.all_reduce(flat[a0:a1], op=dist.ReduceOp.SUM, group=group, async_op=True)
                flat.record_stream(comm)
                w_b = dist.all_reduce(flat[b0:b1], op=dist.ReduceOp.SUM, group=group, async_op=True)  # after the whole backward
                w_a.wait()
                w_b.wait()
                flat.mul_(1.0 / world)
                module._grads_reduced = True
        return (None, None, None, None, *grads)


def _numel(shape):
    n = 1
    for d in shape:
        n *= int(d)
    return n


# _param_tensors() order: layer1.0.{weight, bias}, layer1.1.{weight, bias}, gru.{weight_ih, weight_hh, bias_ih, bias_hh},
# f_classification.0.{weight, bias}.  The backward (csrc/train_api.inc) finishes the classifier and GRU gradients first and
# the layer1 gradients (LayerNorm backward, dW1: the largest GEMM) last.
_BUCKET_A = (4, 5, 6, 7, 8, 9)   # final when prego_train_args_t.gru_grads_event fires
_BUCKET_B = (0, 1, 2, 3)          # final when prego_train_backward's last kernel is done


def gradient_buckets(shapes):
    """Layout of the flat gradient buffer: ((a0, a1), (b0, b1)) element ranges of the two all-reduce buckets, the offset of
    each of the ten tensors (in ``_param_tensors()`` order), and the total length.  Offsets are 16-byte aligned so the
    library's vectorised stores and NCCL see aligned slices."""
    offsets = [0] * len(shapes)
    pos = 0
    bounds = []
    for bucket in (_BUCKET_A, _BUCKET_B):
        start = pos
        for i in bucket:
            offsets[i] = pos
            pos += (_numel(shapes[i]) + 3) // 4 * 4
        bounds.append((start, pos))
    return tuple(bounds), offsets, pos


_COMM_STREAMS = {}


def _comm_stream(device):
    key = torch.device(device).index
    if key not in _COMM_STREAMS:
        _COMM_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COMM_STREAMS[key]


def train_forward(module, rgb, flow):
    """Train-mode forward of ``prego_b200.MROAD`` (raw logits [B, T, K], dropout active, autograd-tracked)."""
    ref = rgb if module.use_rgb else flow
    if not isinstance(ref, torch.Tensor) or not ref.is_cuda:
        raise RuntimeError("prego_b200.MROAD runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # drawn from torch's RNG: reproducible under manual_seed
    return MiniROADTrainFn.apply(module, rgb, flow, seed, *module._param_tensors())


def enable_overlapped_allreduce(module, group=None):
    """Data-parallel training on the CUDA path: reduce the gradients INSIDE the backward, in two buckets over the flat
    gradient buffer -- the gru.* / f_classification.* bucket (9.5 M floats for K = 86) starts on a side stream as soon as
    the library signals it final (``gru_grads_event``) and overlaps the layer1 backward (LayerNorm backward, dW1), the
    layer1 bucket (8.4 M floats) follows the last kernel.  ``allreduce_gradients`` becomes a no-op for
    gradients reduced this way.  No-op itself for a single process."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        module._dp_group = None
        return False
    module._dp_group = group if group is not None else dist.group.WORLD
    return True


def allreduce_gradients(module, group=None):
    """Data-parallel gradient averaging after the backward: ONE all-reduce of the flat gradient buffer (17.9 M floats
    for K = 86), then scatter back.  Generic path (any module, gloo in the CPU tests); skipped when the CUDA backward
    already reduced its gradients (``enable_overlapped_allreduce``).  No-op for a single process."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if getattr(module, "_grads_reduced", False):
        module._grads_reduced = False
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


class FusedAdamW(torch.optim.Optimizer):
    """``torch.optim.AdamW`` (amsgrad = False) with the whole step -- decoupled weight decay, moment updates, bias
    correction, parameter update of every tensor -- in ONE kernel launch through ``prego_adamw_step`` (the reference
    builds ``torch.optim.AdamW(lr=1e-4, weight_decay=0.05)``, main.py:62-67).  ``grad_scale`` multiplies the gradients on
    the fly (1 / world_size after a SUM all-reduce).  CUDA fp32 parameters only; no CPU path."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            for s0 in range(0, len(ps), _lib.ADAMW_MAX_TENSORS):
                chunk = ps[s0:s0 + _lib.ADAMW_MAX_TENSORS]
                a = _lib.AdamWArgs()
                for i, p in enumerate(chunk):
                    if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                        raise RuntimeError("FusedAdamW takes contiguous fp32 CUDA parameters (there is no CPU path)")
                    st = self.state[p]
                    if not st:
                        st["step"] = 0
                        st["exp_avg"] = torch.zeros_like(p)
                        st["exp_avg_sq"] = torch.zeros_like(p)
                    st["step"] += 1
                    a.params[i], a.grads[i] = p.data_ptr(), p.grad.data_ptr()
                    a.exp_avg[i], a.exp_avg_sq[i], a.numel[i] = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()
                a.num_tensors, a.step = len(chunk), self.state[chunk[0]]["step"]
                a.lr, (a.beta1, a.beta2), a.eps = group["lr"], group["betas"], group["eps"]
                a.weight_decay, a.grad_scale = group["weight_decay"], grad_scale
                with torch.cuda.device(chunk[0].device):
                    _lib.check(lib.prego_adamw_step(C.byref(a), torch.cuda.current_stream(chunk[0].device).cuda_stream), "prego_adamw_step")
                for p in chunk:  # the kernel wrote through raw pointers: tell autograd / the weight re-pack check (MROAD._sync_weights)
                    torch.autograd.graph.increment_version(p)
        return loss


TRAINER = Registry()


@TRAINER.register("OAD")
def train_one_epoch(trainloader, model, criterion, optimizer, scaler, epoch, device, writer=None, scheduler=None):
    """trainer/train.py:5-29 with the same signature.  ``scaler`` (the reference's ``--amp`` GradScaler) is accepted and
    ignored: the reduced-precision mode of this implementation is ``cfg['train_precision'] = 'tf32'`` (fp32 storage, no
    loss scaling needed).  Returns the summed loss of the epoch like the reference."""
    epoch_loss = torch.zeros((), device=device)
    for it, (rgb_input, flow_input, target, vid, start, end) in enumerate(trainloader):
        rgb_input, flow_input, target = rgb_input.to(device, non_blocking=True), flow_input.to(device, non_blocking=True), target.to(device, non_blocking=True)
        loss = train_one_step(model, criterion, optimizer, rgb_input, flow_input, target)
        epoch_loss += loss  # stays on the device: no per-iteration sync (the reference calls loss.item() every step)
        if scheduler is not None:
            scheduler.step()
        if writer is not None:
            writer.add_scalar("Train Loss", loss.item(), it + epoch * len(trainloader))
    if hasattr(model, "check_device"):
        model.check_device(f"training epoch {epoch}")   # a watchdog trip in the persistent recurrence / BPTT kernels = garbage gradients
    return float(epoch_loss)


def build_trainer(cfg):
    """trainer/train_builder.py:9-11."""
    return TRAINER[cfg["task"]]


class WindowDataset(torch.utils.data.Dataset):
    """Train-mode view of the reference dataset (datasets/dataset.py:45-135): every video gets ``window_size - 1``
    all-zero dummy frames (features AND targets) in front (dataset.py:53-55,77-82), then is cut into windows of
    ``window_size`` frames every ``stride`` frames, starting at a random offset in [0, stride) that is re-drawn by
    ``_init_features()`` (the reference calls it after every epoch, main.py:101).  The loss reads the LAST frame of a
    window (criterions/loss.py:15-34), so with the front pad every real frame from the very first one is a training
    target, seen with a short (zero-padded) context, and a video shorter than ``window_size`` still yields windows.
    Items: ``(rgb[W, Dr], flow[W, Df], target[W, K], vid, start, end)`` fp32 with ``start`` / ``end`` in the PADDED
    frame numbering, like ``THUMOSDataset.__getitem__`` (dataset.py:125-132).  The pad is virtual: it is written into
    the item, never concatenated to the stored arrays.

    ``videos``: ``{vid: (rgb[T, Dr], flow[T, Df] | None, target[T, K])}`` numpy / torch arrays; a ``None`` flow is the
    all-zero dummy of dataset.py:63-69.  ``front_pad=False`` gives plain windows over the stored frames."""

    def __init__(self, videos, window_size: int, stride: int, d_flow: int = 2048, rng=None, front_pad: bool = True):
        import numpy as np
        self.videos, self.window_size, self.stride, self.d_flow = videos, int(window_size), int(stride), d_flow
        self.pad = self.window_size - 1 if front_pad else 0
        self.rng = rng if rng is not None else np.random
        self.inputs = []
        self._init_features()

    def _init_features(self):
        self.inputs = []
        for vid, (rgb, flow, target) in self.videos.items():
            n = int(target.shape[0]) + self.pad
            seed = int(self.rng.randint(self.stride))
            for start, end in zip(range(seed, n, self.stride), range(seed + self.window_size, n + 1, self.stride)):
                self.inputs.append((vid, start, end))

    def __len__(self):
        return len(self.inputs)

    def __getitem__(self, index):
        vid, start, end = self.inputs[index]
        rgb, flow, target = self.videos[vid]
        lo, hi = max(start - self.pad, 0), end - self.pad   # stored frames covered by the window
        z = (end - start) - (hi - lo)                       # leading dummy frames

        def cut(a, width):
            out = torch.zeros(end - start, width, dtype=torch.float32)
            if a is not None and hi > lo:
                out[z:] = torch.as_tensor(a[lo:hi]).to(torch.float32)
            return out

        return cut(rgb, int(rgb.shape[1])), cut(flow, self.d_flow if flow is None else int(flow.shape[1])), \
            cut(target, int(target.shape[1])), vid, start, end


def train_one_step(model, criterion, optimizer, rgb, flow, target, group=None):
    """One iteration of trainer/train.py:8-24 (+ the gradient all-reduce when run data-parallel: bucketed and overlapped
    with the backward on the CUDA path, see ``enable_overlapped_allreduce``)."""
    model.train()
    if hasattr(model, "_param_tensors") and not hasattr(model, "_dp_group"):
        enable_overlapped_allreduce(model, group)
    out = model(rgb, flow)
    loss = criterion(out, target)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    allreduce_gradients(model, group)
    optimizer.step()
    return loss.detach()


def build_optimizer(cfg, model, fused: bool = True):
    """main.py:62-67: AdamW (or Adam) over all parameters with ``lr`` / ``weight_decay`` from the config.  ``fused``
    selects the one-launch CUDA AdamW; Adam (coupled L2) stays on torch."""
    if cfg.get("optimizer", "AdamW") == "AdamW" and fused:
        return FusedAdamW([{"params": list(model.parameters()), "initial_lr": cfg["lr"]}], lr=cfg["lr"], weight_decay=cfg["weight_decay"])
    optim = torch.optim.AdamW if cfg.get("optimizer", "AdamW") == "AdamW" else torch.optim.Adam
    return optim([{"params": model.parameters(), "initial_lr": cfg["lr"]}], lr=cfg["lr"], weight_decay=cfg["weight_decay"])
