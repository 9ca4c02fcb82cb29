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


def _oracle_ap_fn(scores, labels):
    """Stand-in for the device kernel in the CPU test: the oracle, class by class, same (ap, num_pos) contract."""
    import numpy as np
    from oracle import metrics_np
    s, l = scores.numpy(), labels.numpy()
    K = s.shape[1]
    ap = np.full(K, np.nan)
    pos = np.zeros(K, np.int64)
    for k in range(K):
        y = l == k
        pos[k] = int(y.sum())
        if pos[k]:
            ap[k] = metrics_np.average_precision(y, s[:, k])
    return ap, pos


def _ap_worker(rank, world, port, sizes, K, q):
    import numpy as np
    from prego_b200.sharding import sharded_average_precision
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(100 + rank)
        n = sizes[rank]
        scores = torch.from_numpy((rs.randint(0, 64, (n, K)) / 64.0).astype(np.float32))  # heavy ties across ranks
        labels = torch.from_numpy(rs.randint(0, K, n).astype(np.int32))
        ap, pos = sharded_average_precision(scores, labels, ap_fn=_oracle_ap_fn)
        q.put((rank, ap.tolist(), pos.tolist(), scores.numpy().tolist(), labels.numpy().tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sizes,K", [((37, 52), 7), ((0, 19), 3), ((25, 25), 1)])
def test_sharded_average_precision_world2_matches_single_process(sizes, K):
    """Frames sharded by stream over two ranks, classes re-sharded by one all_to_all: every rank ends up with the AP
    of all K classes, equal to the oracle on the concatenated frames (uneven shards, an empty shard, K < world)."""
    import numpy as np
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ap_worker, args=(r, 2, port, sizes, K, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    scores = torch.tensor(res[0][3] + res[1][3], dtype=torch.float32).reshape(-1, K)
    labels = torch.tensor(res[0][4] + res[1][4], dtype=torch.int32)
    want_ap, want_pos = _oracle_ap_fn(scores, labels)
    for _, ap, pos, _, _ in res:
        assert pos == want_pos.tolist()
        assert np.allclose(np.array(ap), want_ap, rtol=0, atol=1e-12, equal_nan=True)
