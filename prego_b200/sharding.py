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
