"""Stream sharding across the GPUs of one box.

Video streams are independent (the batch dimension is only ever a GEMM row dimension), so the
multi-GPU path is pure data parallelism: rank r owns a contiguous block of stream ids, weights are
replicated, and there is NO collective during compute.  The only exchange is one gather of the
results at the end -- per-frame int32 labels, or (better) the already-collapsed step sequences.
One process per GPU, ``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_streams: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of stream ids owned by ``rank``; sizes differ by at most 1."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_streams, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_ragged(local: torch.Tensor, dst: int = 0, group=None) -> List[torch.Tensor] | None:
    """Gather one 1-D integer tensor of rank-dependent length from every rank onto ``dst``.

    Lengths travel first (all_gather of one int64), then payloads padded to the max length.
    Returns the list of per-rank tensors on ``dst`` and ``None`` elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    lens = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(lens, n, group=group)
    lens = [int(x) for x in lens]
    mx = max(lens + [1])
    padded = torch.zeros(mx, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local.reshape(-1)
    if rank == dst:
        bufs = [torch.empty(mx, dtype=local.dtype, device=local.device) for _ in range(world)]
        dist.gather(padded, bufs, dst=dst, group=group)
        return [b[:l] for b, l in zip(bufs, lens)]
    dist.gather(padded, None, dst=dst, group=group)
    return None


def gather_labels(local_labels: torch.Tensor, n_streams: int, dst: int = 0, group=None) -> torch.Tensor | None:
    """local_labels: int32 [n_local, T] for this rank's shard_bounds block.  Returns the full
    [n_streams, T] tensor on ``dst`` (streams in global id order), None elsewhere."""
    T = local_labels.shape[1]
    parts = gather_ragged(local_labels.reshape(-1), dst, group)
    if parts is None:
        return None
    out = torch.cat(parts).reshape(-1, T)
    assert out.shape[0] == n_streams, (out.shape, n_streams)
    return out


def pack_sequences(seqs: List[List[int]]) -> torch.Tensor:
    """Ragged list of int lists -> flat int64 [count, len_0, ..., len_{n-1}, values...] for gather_ragged."""
    lens = [len(s) for s in seqs]
    flat = [len(seqs)] + lens + [v for s in seqs for v in s]
    return torch.tensor(flat, dtype=torch.int64)


def unpack_sequences(flat: torch.Tensor) -> List[List[int]]:
    vals = flat.tolist()
    n = vals[0]
    lens = vals[1:1 + n]
    out, p = [], 1 + n
    for l in lens:
        out.append(vals[p:p + l])
        p += l
    return out


def sharded_average_precision(scores: torch.Tensor, labels: torch.Tensor, ap_fn=None, group=None):
    """Per-class average precision (utils/metrics.py:25-62) over frames that are sharded by stream across the ranks.

    Every class is independent, so the exchange is a re-shard from frames to CLASSES: rank r receives the scores of
    all frames for its contiguous class block (one all_to_all over NVLink, N_total * K * 4 bytes in total), sorts and
    integrates its classes locally (``prego_perframe_ap``), and the K results are all-gathered.  Labels (4 bytes per
    frame) are all-gathered.  scores: [n_local, K] fp32, labels: int [n_local] (one-hot implied).  Returns
    ``(ap float64 [K], num_pos int64 [K])`` as numpy arrays on every rank, identical to the single-device result on
    the concatenated frames (frame order is irrelevant to AP).  ``ap_fn(scores [N, kc], labels [N]) -> (ap, num_pos)``
    defaults to the device kernel; the CPU tests inject the oracle."""
    import numpy as np

    if ap_fn is None:
        from .metrics import average_precision_per_class as ap_fn
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local, K = int(scores.shape[0]), int(scores.shape[1])
    dev = scores.device
    n = torch.tensor([n_local], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c) for c in counts]
    blocks = [shard_bounds(K, r, world) for r in range(world)]
    k0, k1 = blocks[rank]
    kc = k1 - k0
    # frames -> classes: to rank r goes scores[:, block_r] (row-major), from rank q come its n_q rows of my block
    send = torch.cat([scores[:, a:b].reshape(-1) for a, b in blocks]) if n_local else scores.new_empty(0)
    recv = scores.new_empty(sum(counts) * kc)
    dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=[c * kc for c in counts],
                           input_split_sizes=[n_local * (b - a) for a, b in blocks], group=group)
    mine = recv.reshape(-1, kc) if kc else recv.reshape(sum(counts), 0)
    mx = max(counts + [1])
    lab = torch.full((mx,), -1, dtype=torch.int32, device=dev)
    lab[:n_local] = labels.to(torch.int32)
    labs = [torch.empty_like(lab) for _ in range(world)]
    dist.all_gather(labs, lab, group=group)
    all_labels = torch.cat([l[:c] for l, c in zip(labs, counts)])
    if kc:
        ap, pos = ap_fn(mine, all_labels - k0)  # labels outside [0, kc) match no column of the block
        ap, pos = np.asarray(ap, dtype=np.float64), np.asarray(pos, dtype=np.int64)
    else:
        ap, pos = np.zeros(0, np.float64), np.zeros(0, np.int64)
    kmax = max(b - a for a, b in blocks)
    out = torch.zeros(2, max(kmax, 1), dtype=torch.float64, device=dev)
    out[0, :kc] = torch.from_numpy(ap).to(dev)
    out[1, :kc] = torch.from_numpy(pos.astype(np.float64)).to(dev)
    outs = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(outs, out, group=group)
    ap_all = np.concatenate([o[0, :b - a].cpu().numpy() for o, (a, b) in zip(outs, blocks)])
    pos_all = np.concatenate([o[1, :b - a].cpu().numpy() for o, (a, b) in zip(outs, blocks)]).astype(np.int64)
    return ap_all, pos_all
