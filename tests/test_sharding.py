"""Multi-GPU host logic (stream sharding + result gather) on CPU: gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from prego_b200.sharding import gather_labels, gather_ragged, pack_sequences, shard_bounds, unpack_sequences


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 8, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_pack_unpack_roundtrip():
    seqs = [[1, 2, 3], [], [7], list(range(50))]
    assert unpack_sequences(pack_sequences(seqs)) == seqs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, T, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_streams, rank, world)
        # stand-in for the per-rank CUDA result: label of (stream s, frame t) = (s * 31 + t) % 86
        s = torch.arange(lo, hi).unsqueeze(1)
        local = ((s * 31 + torch.arange(T).unsqueeze(0)) % 86).to(torch.int32)
        full = gather_labels(local, n_streams, dst=0)
        seqs = [[int(v) for v in row[: (i % 5) + 1]] for i, row in enumerate(local.tolist())]
        parts = gather_ragged(pack_sequences(seqs), dst=0)
        if rank == 0:
            allseq = [x for p in parts for x in unpack_sequences(p)]
            q.put((full.tolist(), allseq))
    finally:
        dist.destroy_process_group()


def test_gather_world2_matches_single_process():
    n_streams, T, world = 11, 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, allseq = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    s = torch.arange(n_streams).unsqueeze(1)
    expect = ((s * 31 + torch.arange(T).unsqueeze(0)) % 86).to(torch.int32)
    assert full == expect.tolist()
    # ragged sequences arrive in global stream order
    exp_seqs = []
    for r in range(world):
        lo, hi = shard_bounds(n_streams, r, world)
        exp_seqs += [expect[g, : (i % 5) + 1].tolist() for i, g in enumerate(range(lo, hi))]
    assert allseq == exp_seqs
