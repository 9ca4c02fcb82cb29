#!/usr/bin/env python
"""BASELINE configs[3]: N streams sharded by stream id across the GPUs of one box, per-frame labels collapsed to step
sequences on each rank (window vote + RLE kernels), the collapsed sequences gathered on rank 0 -- and checked there
against the SAME stream ids computed by rank 0 alone (bit-exact: no result may depend on the sharding).

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/config4_sharded.py --streams 65536 --frames 256

Features are a counter-based hash of (stream id, frame, channel), so any rank can materialise any stream.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prego_b200 import synthetic  # noqa: E402
from prego_b200.aggregate import aggregate_device  # noqa: E402
from prego_b200.sharding import gather_ragged, shard_bounds  # noqa: E402


def hashed_features(ids, t0, T, dev, salt):
    """fp32 [len(ids), T, 2048] in [0, 1): deterministic in (stream id, absolute frame, channel) only."""
    i = ids.to(torch.int64).view(-1, 1, 1)
    t = (t0 + torch.arange(T, device=dev, dtype=torch.int64)).view(1, -1, 1)
    d = torch.arange(2048, device=dev, dtype=torch.int64).view(1, 1, -1)
    h = (i * 1000003 + t * 10007 + d * 101 + salt) * 2654435761
    h = (h ^ (h >> 15)) & 0xFFFFFFFF
    h = (h * 2246822519) & 0xFFFFFFFF
    h = h ^ (h >> 13)
    return (h & 0xFFFFFF).to(torch.float32) / float(1 << 24)


def run_block(model, ids, T, chunk, dev, sub_streams):
    """Labels [len(ids), T] int32 for the given stream ids, in sub-batches of streams and time chunks with carried state."""
    out = torch.empty(len(ids), T, dtype=torch.int32, device=dev)
    for s in range(0, len(ids), sub_streams):
        sid = ids[s:s + sub_streams]
        h = torch.zeros(len(sid), 1024, device=dev)
        for t0 in range(0, T, chunk):
            tc = min(chunk, T - t0)
            rgb = hashed_features(sid, t0, tc, dev, 1)
            flow = hashed_features(sid, t0, tc, dev, 2)
            out[s:s + len(sid), t0:t0 + tc] = model.infer(rgb, flow, h_state=h, want_probs=False, precision="fp16", chunk_T=tc)["labels"]
    return out


def collapse(labels, window):
    """Per-stream step sequences on the device (window vote + RLE); returns (counts int64 [B], compacted values int64)."""
    B, T = labels.shape
    flat = labels.reshape(-1).contiguous()
    r = aggregate_device(flat, [T] * B, flat, [T] * B, window, 86)
    counts = r["pred_counts"].to(torch.int64)
    wl = (T + window - 1) // window  # value slots per stream
    vals = r["pred_vals"][: B * wl].to(torch.int64).reshape(B, wl)
    keep = torch.arange(wl, device=labels.device).view(1, -1) < counts.view(-1, 1)
    return counts, vals[keep]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=65536)
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--sub-streams", type=int, default=4096)
    ap.add_argument("--window", type=int, default=200)
    ap.add_argument("--no-check", action="store_true", help="skip rank 0's single-GPU recomputation of every stream")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = synthetic.seeded_model(dict(synthetic.ASSEMBLY101_O), seed=20, device=dev)
    lo, hi = shard_bounds(args.streams, rank, world)
    ids = torch.arange(lo, hi, device=dev)
    run_block(model, ids[:256], min(args.frames, 64), args.chunk, dev, args.sub_streams)  # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    labels = run_block(model, ids, args.frames, args.chunk, dev, args.sub_streams)
    counts, vals = collapse(labels, args.window)
    flat = torch.cat([torch.tensor([counts.numel()], device=dev, dtype=torch.int64), counts, vals])
    parts = gather_ragged(flat, 0) if world > 1 else [flat]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        seqs = []
        for p in parts:
            n = int(p[0])
            c = p[1:1 + n].tolist()
            v = p[1 + n:].tolist()
            q = 0
            for k in c:
                seqs.append(v[q:q + k])
                q += k
        assert len(seqs) == args.streams
        ok = None
        if not args.no_check:
            all_ids = torch.arange(args.streams, device=dev)
            ref_labels = run_block(model, all_ids, args.frames, args.chunk, dev, args.sub_streams)
            rcnt, rval = collapse(ref_labels, args.window)
            rc, rv = rcnt.tolist(), rval.tolist()
            ref, q = [], 0
            for k in rc:
                ref.append(rv[q:q + k])
                q += k
            ok = ref == seqs
        print(json.dumps({"config": "65,536-stream shape sharded by stream (BASELINE configs[3])", "streams": args.streams, "frames": args.frames,
                          "n_gpus": world, "frames_per_s_incl_feature_synthesis_and_gather": args.streams * args.frames / dt,
                          "mean_steps_per_stream": sum(map(len, seqs)) / len(seqs), "gathered_bytes": int(sum(p.numel() for p in parts) * 8),
                          "sharded_equals_single_gpu": ok}))
        if ok is False:
            sys.exit(1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
