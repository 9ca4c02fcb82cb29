#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MiniROAD online-inference path.

Contract (driver):  python bench.py --gpus N --steps K --warmup W [--impl reference]
One rank per GPU (torchrun env for N > 1).  Prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[2], "Assembly101-O-shaped, 4096 concurrent streams ... on 1xB200"):
a STEP is one pass of the hot path -- feature staging, Linear+LayerNorm+ReLU, GRU input gates,
GRU recurrence with carried state, classifier+softmax+argmax -- over one time chunk of
`--chunk` (64) frames for each of `--streams` (4096) concurrent streams per GPU, i.e. 262 144 frames
(4 GiB of fp32 features, far larger than the 126 MB L2) per GPU per step.  After the K timed steps
the per-frame labels of all K chunks are collapsed to step sequences (200-frame window vote + RLE,
utils/aggregate.py) inside the timed region.  Streams shard by GPU with no data-path collective
(weak scaling: per-GPU work is fixed).  value = frames / s over all GPUs, inputs resident in HBM.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_GEMM1 = 2 * 4096 * 2048          # per frame
FLOP_GEMM2 = 2 * 2048 * 3072
FLOP_REC = 2 * 1024 * 3072
BYTES_FEATURES = 16384                # fp32 rgb + flow per frame


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="concurrent streams per GPU")
    ap.add_argument("--chunk", type=int, default=64, help="frames per stream per step")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--no-latency", action="store_true", help="skip the single-stream latency leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg")
    ap.add_argument("--no-rank4", action="store_true", help="skip the MiniROADA anticipation / device mAP leg (SURVEY 8f rank 4)")
    ap.add_argument("--no-variants", action="store_true", help="skip the feature-format variants (16-bit features, zero flow)")
    ap.add_argument("--e2e-direct", type=int, default=0, help="e2e: send this many leading streams of every batch as plain fp32 next to the host-rounded rest")
    ap.add_argument("--no-library", action="store_true", help="skip the stock-torch (cuBLAS + cuDNN) baseline on the same GPU")
    ap.add_argument("--subchunk", type=int, default=64,
                    help="internal time-chunk of prego_forward inside one step; < --chunk stages the features of chunk c+1 on a side "
                         "stream under chunk c (measured +2 %% at 32, but it blurs the per-kernel roofline timing, so off by default)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock / throttle reasons over the samples that arrived inside the wall-clock window [t_begin, t_end]."""
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for t, r in self.rows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end)]
        sm = [float(r[0]) for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for _, r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


def dist_env(n):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
CPU_SAMPLE_STREAMS = 256  # bounded sample of the 4 096-stream workload: large enough that the CPU GEMMs run at their own batch efficiency


def cpu_baseline_run(streams, chunk, target_seconds=12.0, repeats=1):
    """Time the ATen-based CPU port of the reference path (oracle/miniroad_torch_cpu.py) on a bounded
    sample of the workload, all host threads.  Returns (frames/s, description, cores)."""
    from oracle.miniroad_torch_cpu import CpuMiniROAD
    from oracle import aggregate_np
    from prego_b200 import synthetic

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = synthetic.seeded_model(dict(synthetic.ASSEMBLY101_O), seed=20)
    port = CpuMiniROAD(model.state_dict())
    bs = min(streams, CPU_SAMPLE_STREAMS)
    g = torch.Generator().manual_seed(1)
    rgb = torch.randn(bs, chunk, 2048, generator=g).abs_()
    flow = torch.randn(bs, chunk, 2048, generator=g).abs_()
    t0 = time.perf_counter()
    labels = port.labels(rgb, flow)  # warm-up + calibration
    t1 = time.perf_counter() - t0
    reps = max(1, min(64, int(target_seconds / max(t1, 1e-3))))
    best = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        for _ in range(reps):
            labels = port.labels(rgb, flow)
            for b in range(bs):
                aggregate_np.aggregate_video(labels[b], labels[b])
        best.append((time.perf_counter() - t0) / reps)
    dt = float(np.median(best))
    return bs * chunk / dt, f"{bs} of {streams} streams x {chunk} frames per pass, {reps} passes, torch {torch.__version__} CPU (ATen/oneDNN), + aggregate", cores


def cpu_whole_video_run(lengths=(2011, 9507)):
    """BASELINE.md 4.3(a) / configs[0]: the reference's own evaluation shape on the CPU -- ONE whole video per forward
    (test_batch_size: 1, configs/miniroad_assembly101-O.yaml:17; Assembly101-O median and maximum length), all host
    threads, followed by aggregate().  Returns {T: {...}}."""
    from oracle.miniroad_torch_cpu import CpuMiniROAD
    from oracle import aggregate_np
    from prego_b200 import synthetic

    torch.set_num_threads(os.cpu_count() or 1)
    model = synthetic.seeded_model(dict(synthetic.ASSEMBLY101_O), seed=20)
    port = CpuMiniROAD(model.state_dict())
    out = {}
    for T in lengths:
        rgb, flow = synthetic.feature_batch([7], T, "cpu", False)
        port.labels(rgb[:, :64], flow[:, :64])  # warm-up
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            labels = port.labels(rgb, flow)
            aggregate_np.aggregate_video(labels[0], labels[0])
            ts.append(time.perf_counter() - t0)
        out[f"T{T}"] = {"frames": T, "seconds": min(ts), "frames_per_s": T / min(ts)}
    out["note"] = "B = 1 whole-video sequences (the reference's test_batch_size: 1), ATen CPU port + aggregate, best of 2"
    return out


def run_reference(args, world, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference's Python
    files cannot travel to the GPU box and it has no compiled sources), rank 0 only."""
    if rank != 0:
        return
    from oracle.miniroad_torch_cpu import CpuMiniROAD
    from oracle import aggregate_np
    from prego_b200 import synthetic

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = synthetic.seeded_model(dict(synthetic.ASSEMBLY101_O), seed=20)
    port = CpuMiniROAD(model.state_dict())
    bs = min(args.streams, CPU_SAMPLE_STREAMS)
    g = torch.Generator().manual_seed(1)
    rgb = torch.randn(bs, args.chunk, 2048, generator=g).abs_()
    flow = torch.randn(bs, args.chunk, 2048, generator=g).abs_()

    def step():
        labels = port.labels(rgb, flow)
        for b in range(bs):
            aggregate_np.aggregate_video(labels[b], labels[b])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = bs * args.chunk * args.steps / dt
    sample = f"{bs} of {args.streams} streams x {args.chunk} frames per step (bounded sample of the same workload)"
    line = {"impl": "reference", "metric": "frames/sec MiniROAD online inference", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Assembly101-O-shaped, {args.streams} streams x {args.chunk}-frame chunks (K=86)", "sample": sample},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                             "whole_video": cpu_whole_video_run()},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------- single-stream latency leg
def latency_leg(dev, precision):
    """BASELINE configs[1]: Epic-tent-O shape, one stream, on one B200."""
    from prego_b200 import synthetic
    model = synthetic.seeded_model(dict(synthetic.EPIC_TENT_O), seed=20, device=dev)
    out = {}
    # (a) whole video, T = 12 531 (dataset mean): projections batched over the video, recurrence sequential
    T = 12531
    rgb, flow = synthetic.device_features(1, T, dev, seed=3)
    for _ in range(2):
        model.infer(rgb, flow, want_probs=False, precision=precision)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record()
        model.infer(rgb, flow, want_probs=False, precision=precision)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    out["whole_video"] = {"T": T, "ms_per_video": ms, "ms_per_frame": ms / T, "frames_per_s": T / ms * 1e3}
    model.profile_begin()
    model.infer(rgb, flow, want_probs=False, precision=precision)
    prof = model.profile_end()
    out["whole_video"]["recurrence_us_per_step"] = prof["recurrence"]["ms"] / T * 1e3
    # (b) strict per-frame online stepping: one frame per call, carried h, label read back each frame
    n = 300
    sess = model.online_session(1, dev, precision, host_labels=True)
    frames_r = [rgb[0, t].contiguous() for t in range(n + 20)]
    frames_f = [flow[0, t].contiguous() for t in range(n + 20)]
    for t in range(20):
        sess.step(frames_r[t], frames_f[t])
    torch.cuda.synchronize()
    wall = []
    stream = torch.cuda.current_stream(dev)
    for t in range(20, n + 20):
        t0 = time.perf_counter()
        sess.step_wait(frames_r[t], frames_f[t])
        _ = int(sess.labels_np[0, 0])  # the kernel stored the label into pinned host memory before ringing the doorbell
        wall.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    wall_sync = []
    for t in range(20, n + 20):
        t0 = time.perf_counter()
        sess.step(frames_r[t], frames_f[t])
        stream.synchronize()
        _ = int(sess.labels_np[0, 0])
        wall_sync.append((time.perf_counter() - t0) * 1e3)
    sess_dev = model.online_session(1, dev, precision)  # device-resident labels: the kernel's own cost, no PCIe store
    for t in range(20):
        sess_dev.step(frames_r[t], frames_f[t])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(20, 220):
        sess_dev.step(frames_r[t], frames_f[t])
    e1.record()
    torch.cuda.synchronize()
    gpu_us = e0.elapsed_time(e1) / 200 * 1e3  # back-to-back graph launches: device-side cost of one frame
    out["per_frame_online"] = {"frames": n, "p50_ms": float(np.percentile(wall, 50)), "p99_ms": float(np.percentile(wall, 99)),
                               "p50_ms_stream_sync": float(np.percentile(wall_sync, 50)), "gpu_us_per_frame": gpu_us,
                               "note": "OnlineSession.step per frame: one CUDA-graph launch (one cooperative kernel: all layers, carried state in place), label stored by the kernel into pinned host memory, completion = host spin on a pinned doorbell word per stream {frame number, label} written by the kernel in one 8-byte store (prego_online_wait); p50_ms_stream_sync: the same with cudaStreamSynchronize instead; wall clock"}
    return out


# ----------------------------------------------------------------------------- training-step leg
def training_leg(dev, world):
    """BASELINE configs[4]: MiniROAD training step (forward + BPTT backward + AdamW), per-GPU batch 16 (the
    reference's) and 256 windows of 128 frames, zero flow as the reference loader feeds it, dropout 0.2,
    gradients all-reduced over NCCL when world > 1.  Exact-fp32 CUDA-core kernels (first version)."""
    import torch.distributed as dist
    from prego_b200 import OadLoss, synthetic, train_one_step
    cfg = dict(synthetic.ASSEMBLY101_O)
    model = synthetic.seeded_model(cfg, seed=20, device=dev)
    crit = OadLoss(cfg)
    from prego_b200 import build_optimizer
    opt = build_optimizer({"optimizer": "AdamW", "lr": 1e-4, "weight_decay": 0.05}, model)  # one-launch fused AdamW
    out = {}
    for prec in ("fp32", "tf32x3", "tf32"):
        model.train_precision = prec
        for B in (16, 256):
            T = 128
            rgb, flow = synthetic.device_features(B, T, dev, seed=7, zero_flow=True)
            target = torch.nn.functional.one_hot(torch.randint(0, 86, (B, T), device=dev), 86).float()
            for _ in range(4 if world > 1 else 2):  # NCCL sets up its channels lazily over the first collectives of a size class
                train_one_step(model, crit, opt, rgb, flow, target)
            torch.cuda.synchronize()
            n = 5 if B == 16 else 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            e0.record()
            for _ in range(n):
                loss = train_one_step(model, crit, opt, rgb, flow, target)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            out[f"B{B}_T{T}" + ("" if prec == "fp32" else "_" + prec)] = {"ms_per_step": float(ms), "frames_per_s": world * B * T / float(ms) * 1e3,
                                                                       "loss": float(loss)}
            del rgb, flow, target
    if world > 1:
        # the data-parallel exchange on its own: one all-reduce of the flat gradient buffer (17.9 M fp32) over NCCL / NVLink
        from prego_b200 import allreduce_gradients
        for _ in range(3):
            allreduce_gradients(model)
        torch.cuda.synchronize()
        dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            allreduce_gradients(model)
        a1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a0.elapsed_time(a1) / 10], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        nbytes = sum(p.numel() for p in model.parameters()) * 4
        out["grad_allreduce"] = {"ms": float(ms), "bytes": nbytes, "bus_gbs": 2 * (world - 1) / world * nbytes / (float(ms) * 1e-3) / 1e9,
                                 "note": "the exchange on its own, generic path: flat-buffer gather + NCCL all-reduce (sum) + scatter back; bus GB/s = 2 (N-1)/N x bytes / time. "
                                         "Inside the timed training steps above the gradients are reduced IN the backward instead: two buckets of one flat buffer, the "
                                         "gru / classifier bucket on a side stream under the layer1 backward (prego_b200.training.enable_overlapped_allreduce)"}
    out["note"] = ("fwd + BPTT + fused AdamW (one launch), dropout 0.2, flow = 0; recurrence (forward and BPTT) on the persistent exact-fp32 kernels "
                   "for B <= 64 in every mode; plain keys: every GEMM exact fp32 on CUDA cores (parity mode); *_tf32x3: the projections and their "
                   "gradients on tcgen05 kind::tf32 with every operand split hi + lo, three terms per 1 024-column chunk, small terms first, chunks added in fp32 (fp32-class forward: "
                   "logits 2.3e-6 of ATen's; gradients through the hard ReLU gates within 3e-3 Frobenius, profiles/r02_train_modes.txt); *_tf32: plain TF32 operands; grads all-reduced (NCCL) when n_gpus > 1.  The stock-torch step on the same "
                   "GPU is library_baseline.train_step")
    return out


# ----------------------------------------------------------------------------- SURVEY 8f rank 4 leg
def rank4_leg(dev):
    """MiniROADA (rnn.py:73-137) anticipation inference and the device per-frame mAP (utils/metrics.py:25-62), one GPU.
    Anticipation: 1 024 streams x 64 frames, A = 4 (no shipped config selects MiniROADA, so A is ours), fp16 operands,
    probabilities [B, T, A, K] materialised.  mAP: the 262 144 x 86 probabilities of one bench step."""
    from prego_b200 import MROADA, synthetic
    from prego_b200.metrics import average_precision_per_class
    out = {}
    A, B, T, K = 4, 1024, 64, 86
    cfg = dict(synthetic.ASSEMBLY101_O, model="MiniROADA", anticipation_length=A, actionness=False)
    torch.manual_seed(20)
    model = MROADA(cfg).to(dev).eval()
    rgb, flow = synthetic.device_features(B, T, dev, seed=11)

    def timed(fn, n):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, r

    ms_ant, _ = timed(lambda: model.infer(rgb, flow, want_probs=False, want_anticipation=True), 5)
    ms_plain, _ = timed(lambda: model.infer(rgb, flow, want_probs=False), 5)
    extra_flop = (2 * 1024 * A * 1024 + 2 * 1024 * K * A) * B * T  # anticipation layer + A classifier rows per frame
    out["anticipation"] = {"streams": B, "frames": T, "anticipation_length": A, "ms": ms_ant, "frames_per_s": B * T / ms_ant * 1e3,
                           "ms_trunk_only": ms_plain, "head_tflops": extra_flop / max(ms_ant - ms_plain, 1e-6) / 1e9,
                           "note": "prego_forward_anticipation, fp16 operands; head_tflops = (anticipation layer + A classifier rows) / (ms - ms_trunk_only), "
                                   "includes writing the [B, T, A, K] fp32 probabilities (90 MB)"}
    del rgb, flow, model
    N = 262144
    g = torch.Generator(device=dev).manual_seed(5)
    scores = torch.softmax(torch.randn(N, K, generator=g, device=dev) * 3, -1)
    labels = torch.randint(0, K, (N,), generator=g, device=dev, dtype=torch.int32)
    ms_ap, (ap, _) = timed(lambda: average_precision_per_class(scores, labels), 5)
    from sklearn.metrics import average_precision_score  # what the reference calls per class on the host (metrics.py:43,55)
    s_h, l_h = scores[:, 1].cpu().numpy(), labels.cpu().numpy() == 1
    t0 = time.perf_counter()
    ref = average_precision_score(l_h, s_h)
    host_ms = (time.perf_counter() - t0) * 1e3
    assert abs(ref - ap[1]) <= 1e-12, "device AP differs from sklearn"
    out["perframe_map"] = {"frames": N, "classes": K, "ms": ms_ap, "algorithmic_gbs": N * K * 64 / ms_ap / 1e6,
                           "mean_ap": float(np.nanmean(ap[1:])), "host_sklearn_ms_per_class": host_ms,
                           "note": "prego_perframe_ap incl. the K-double D2H; 64 algorithmic bytes per (frame, class): 4-pass LSD radix sort + scan; "
                                   "the reference runs sklearn once per class on the host (85 classes here); class 1 checked against it to 1e-12"}
    return out


# ----------------------------------------------------------------------------- library baseline (stock torch on the same GPU)
def library_baseline_leg(dev, B, Tc, local):
    """SURVEY 2a / BASELINE.md 4.6: the reference's own modules -- nn.Linear / LayerNorm / nn.GRU (cuBLAS + cuDNN persistent RNN) /
    softmax / argmax, rnn.py:38-71 + eval.py:53 -- on the SAME B200, same step (B streams x Tc frames resident in HBM, K = 86), in
    the three precisions a user of the reference could select.  Library code end to end: none of this repo's kernels run here."""
    import torch.nn as nn

    class Ref(nn.Module):
        def __init__(self):
            super().__init__()
            self.gru = nn.GRU(2048, 1024, 1, batch_first=True)
            self.layer1 = nn.Sequential(nn.Linear(4096, 2048), nn.LayerNorm(2048), nn.ReLU(), nn.Dropout(0.2))
            self.fc = nn.Linear(1024, 86)

        def forward(self, rgb, flow):
            x = self.layer1(torch.cat((rgb, flow), 2))
            ht, _ = self.gru(x, torch.zeros(1, x.shape[0], 1024, device=x.device, dtype=x.dtype))
            return torch.softmax(self.fc(torch.relu(ht)), -1).argmax(-1)

    torch.manual_seed(20)
    m = Ref().to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1)
    rgb = torch.randn(B, Tc, 2048, generator=g, device=dev).abs_()
    flow = torch.randn(B, Tc, 2048, generator=g, device=dev).abs_()
    out = {}
    tf32_before = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    for mode, n in (("bf16_autocast", 8), ("tf32", 8), ("fp32", 3)):
        torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
        torch.backends.cudnn.allow_tf32 = mode != "fp32"
        try:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode.startswith("bf16")):
                for _ in range(3):
                    m(rgb, flow)
                sampler = ClockSampler(local)
                sampler.start()
                time.sleep(0.25)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                t0 = time.time()
                e0.record()
                for _ in range(n):
                    m(rgb, flow)
                e1.record()
                torch.cuda.synchronize()
                t1 = time.time()
            ms = e0.elapsed_time(e1) / n
            out[mode] = {"ms_per_step": ms, "frames_per_s": B * Tc / ms * 1e3, "steps": n, "clocks": sampler.stop(t0, t1)}
        except Exception as e:  # noqa: BLE001
            out[mode] = {"error": repr(e)[:200]}
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32_before
    out["train_step"] = library_train_step(dev, Ref)
    out["note"] = (f"torch {torch.__version__} (cuBLAS + cuDNN GRU), the reference's module structure, {B} streams x {Tc} frames per step resident in HBM, "
                   "eval mode, softmax + device argmax; same GPU, same process, right after this repo's legs")
    del m, rgb, flow
    torch.cuda.empty_cache()
    return out


def library_train_step(dev, Ref):
    """BASELINE configs[4] on the library stack: the reference's training step (trainer/train.py:10-23: forward in train mode,
    last-frame multi-label CE of criterions/loss.py:15-34, zero_grad, backward, torch.optim.AdamW of main.py:62-67) on stock
    torch modules, same shapes and zero flow as training_leg.  `torch_defaults` = what the reference runs as shipped (fp32
    cuBLAS matmuls, cuDNN GRU free to use TF32); `tf32` = both on TF32; `amp_fp16` = main.py --amp (autocast + GradScaler)."""
    out = {}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    for mode in ("torch_defaults", "tf32", "amp_fp16"):
        torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
        torch.backends.cudnn.allow_tf32 = True
        for B in (16, 256):
            T = 128
            try:
                torch.manual_seed(20)
                m = Ref().to(dev).train()
                opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0.05)
                scaler = torch.amp.GradScaler("cuda") if mode == "amp_fp16" else None
                g = torch.Generator(device=dev).manual_seed(7)
                rgb = torch.randn(B, T, 2048, generator=g, device=dev).abs_()
                flow = torch.zeros(B, T, 2048, device=dev)
                target = torch.nn.functional.one_hot(torch.randint(0, 86, (B, T), device=dev), 86).float()

                def step():
                    with torch.autocast("cuda", dtype=torch.float16, enabled=scaler is not None):
                        x = m.layer1(torch.cat((rgb, flow), 2))
                        ht, _ = m.gru(x, torch.zeros(1, B, 1024, device=dev, dtype=x.dtype))
                        logits = m.fc(torch.relu(ht))[:, -1].float()
                        tgt = target[:, -1]
                        loss = (-(tgt / tgt.norm(dim=1, keepdim=True).clamp_min(1e-12)) * torch.log_softmax(logits, -1)).sum(1).mean()
                    opt.zero_grad(set_to_none=True)
                    if scaler is not None:
                        scaler.scale(loss).backward()
                        scaler.step(opt)
                        scaler.update()
                    else:
                        loss.backward()
                        opt.step()
                    return loss

                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                n = 5 if B == 16 else 3
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    loss = step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                out[f"B{B}_T{T}_{mode}"] = {"ms_per_step": ms, "frames_per_s": B * T / ms * 1e3, "loss": float(loss.detach())}
                del m, opt, rgb, flow, target
            except Exception as e:  # noqa: BLE001
                out[f"B{B}_T{T}_{mode}"] = {"error": repr(e)[:200]}
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    torch.cuda.empty_cache()
    return out


def ncu_record(B, Tc, args):
    """`roofline.traffic` cannot be measured inside a timed run (ncu replays kernels): it is read from the committed capture
    profiles/ncu_gemm1.json (written by scripts/ncu_extract.py from one `ncu --set full` of this kernel at this shape) and is
    reported ONLY while the kernel sources still hash to what was captured -- a changed kernel yields null, never a stale number."""
    import hashlib
    rec_path = os.path.join(ROOT, "profiles", "ncu_gemm1.json")
    out = {"traffic": None, "traffic_source": None}
    if not os.path.exists(rec_path):
        return out
    rec = json.load(open(rec_path))
    h = hashlib.sha256()
    for f in rec.get("sources", []):
        h.update(open(os.path.join(ROOT, f), "rb").read())
    same_shape = (rec.get("streams"), rec.get("chunk"), rec.get("precision")) == (B, min(Tc, args.subchunk), args.precision)
    if h.hexdigest() == rec.get("sources_sha256") and same_shape:
        out.update({"traffic": rec["dram_bytes_read"] + rec["dram_bytes_write"], "tensor_pipe_active_pct_ncu": rec.get("tensor_pipe_active_pct"),
                    "traffic_source": f"{rec['capture']} (kernel sources unchanged since: sha256 {rec['sources_sha256'][:12]})"})
    else:
        out["traffic_source"] = f"{rec.get('capture')} is stale for this build / shape: not reported"
    return out


# ----------------------------------------------------------------------------- main arm
def run_ours(args, world, rank, local):
    import torch.distributed as dist
    from prego_b200 import _lib, synthetic
    from prego_b200.aggregate import aggregate_device

    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback)"
    _lib.load()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, Tc, K, W = args.streams, args.chunk, args.steps, args.warmup
    model = synthetic.seeded_model(dict(synthetic.ASSEMBLY101_O), seed=20, device=dev)
    # per-rank shard of the stream population: distinct seeds per rank, features resident in HBM
    rgb, flow = synthetic.device_features(B, Tc, dev, seed=1234 + rank)
    h = torch.zeros(B, 1024, device=dev)
    labels_all = torch.empty(K, B, Tc, dtype=torch.int32, device=dev)

    def step(i):
        out = model.infer(rgb, flow, h_state=h, want_probs=False, want_labels=True, precision=args.precision, chunk_T=min(Tc, args.subchunk))
        labels_all[i % K].copy_(out["labels"])

    from prego_b200.sharding import gather_ragged
    gathered = {"bytes": 0, "sequences": 0}

    def collapse():
        """labels of all K steps -> step sequences per stream (window vote + RLE on the device); under torchrun the collapsed
        sequences of every rank are gathered on rank 0 (BASELINE configs[3]: 'outputs gathered and collapsed'; SURVEY 8e: the ONE
        exchange of the inference path, ~200x smaller than the labels)."""
        seq = labels_all.permute(1, 0, 2).reshape(B, K * Tc).contiguous()
        flat = seq.reshape(-1)
        r = aggregate_device(flat, [K * Tc] * B, flat, [K * Tc] * B, 200, 86)
        total = r["pred_counts"].sum() + r["gt_counts"].sum()
        if world > 1:
            counts = r["pred_counts"].to(torch.int64)
            wl = (K * Tc + 199) // 200
            vals = r["pred_vals"][: B * wl].to(torch.int64).reshape(B, wl)
            keep = torch.arange(wl, device=dev).view(1, -1) < counts.view(-1, 1)
            packed = torch.cat([torch.tensor([B], device=dev, dtype=torch.int64), counts, vals[keep]])
            parts = gather_ragged(packed, 0)
            if parts is not None:
                gathered["bytes"] = int(sum(p.numel() for p in parts) * 8)
                gathered["sequences"] = int(sum(int(p[0]) for p in parts))
        return total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    collapse()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)  # nvidia-smi needs a moment to start
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_timed0 = time.time()
    e0.record()
    for i in range(K):
        step(i)
    total_runs = collapse()
    e1.record()
    barrier()
    t_timed1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    frames = world * B * Tc * K
    value = frames / ms * 1e3
    _ = int(total_runs)

    # phase profile over K more steps (CUDA events on the launching stream inside the library)
    model.profile_begin()
    for i in range(K):
        step(i)
    prof = model.profile_end()
    # the same step held for ~0.5 s: the board settles at its power cap, and the 20 ms clock sampler gets a real window
    n_sus = max(4 * K, 48)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_sus0 = time.time()
    s0.record()
    for i in range(n_sus):
        step(i)
    s1.record()
    barrier()
    t_sus1 = time.time()
    sus_ms = torch.tensor([s0.elapsed_time(s1)], device=dev)
    if world > 1:
        dist.all_reduce(sus_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop(t_timed0, t_timed1) if rank == 0 else None
    sustained = {"value": world * B * Tc * n_sus / float(sus_ms) * 1e3, "unit": "frames/s", "steps": n_sus, "ms_per_step": float(sus_ms) / n_sus,
                 "clocks": sampler.stop(t_sus0, t_sus1) if rank == 0 else None,
                 "note": "the same step repeated back to back right after the timed region (no collapse): the board settles at its power cap"}
    # feature-format variants: the same step with features already in the 16-bit operand format
    # (no staging pass: GEMM1's TMA reads the caller's tensors in place) and / or the flow stream declared all-zero as
    # the shipped configs feed it (its half of the projection is skipped; NOT counted as achieved FLOPs anywhere)
    variants = None
    dt16 = {"fp16": torch.float16, "bf16": torch.bfloat16}.get(args.precision)
    if dt16 is not None and not args.no_variants:
        rgb16, flow16 = rgb.to(dt16), flow.to(dt16)
        hv = torch.zeros(B, 1024, device=dev)

        def timed(r, f, zf):
            def one():
                model.infer(r, f, h_state=hv, want_probs=False, want_labels=True, precision=args.precision,
                            chunk_T=min(Tc, args.subchunk), zero_flow=zf)
            for _ in range(2):
                one()
            barrier()
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record()
            for _ in range(K):
                one()
            v1.record()
            barrier()
            t = torch.tensor([v0.elapsed_time(v1) / K], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)

        variants = {}
        for name, (r, f, zf, flops) in {"feat16": (rgb16, flow16, False, FLOP_GEMM1 + FLOP_GEMM2),
                                        "fp32_zero_flow": (rgb, None, True, FLOP_GEMM1 // 2 + FLOP_GEMM2),
                                        "feat16_zero_flow": (rgb16, None, True, FLOP_GEMM1 // 2 + FLOP_GEMM2)}.items():
            t = timed(r, f, zf)
            variants[name] = {"frames_per_s": world * B * Tc / t * 1e3, "ms_per_step": t, "projection_flops_per_frame": flops}
        # the fp32-class modes on the same step: split-fp16 operands on the tensor cores (1e-4 logit bound, measured ~1e-5, labels identical to
        # the reference in 131 072 / 131 072 frames) and the exact CUDA-core FFMA path it replaces as the fast "exact" mode
        for name, steps in (("fp16x3", max(2, K // 2)), ("fp32", 1)):
            def one_exact():
                model.infer(rgb, flow, h_state=hv, want_probs=False, want_labels=True, precision=name, chunk_T=min(Tc, args.subchunk))
            one_exact()
            barrier()
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record()
            for _ in range(steps):
                one_exact()
            v1.record()
            barrier()
            t = torch.tensor([v0.elapsed_time(v1) / steps], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            variants["precision_" + name] = {"frames_per_s": world * B * Tc / float(t) * 1e3, "ms_per_step": float(t), "dtype": {"fp16x3": "f16 hi+lo (fp32-class)", "fp32": "f32"}[name]}
        model._workspace = None  # the exact modes' workspace is several times larger: give it back before the e2e leg
        torch.cuda.empty_cache()
        variants["note"] = ("inputs resident in HBM, same step as `value` without the collapse; feat16 = rgb/flow stored as "
                            f"{args.precision} (bit-identical results, tests/test_gpu_parity.py); zero_flow = flow declared all-zero "
                            "(dataset.py:63-69), its projection half skipped")

    # end-to-end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream(dev)

        def e2e_measure(hr, hf, zf):
            """hr / hf: pinned host features of one step (hf None with zero_flow).  Returns frames/s over all ranks."""
            # double-buffered device inputs: the H2D copy of step i+1 (copy stream) overlaps the compute of step i
            dbuf = [(torch.empty(hr.shape, dtype=hr.dtype, device=dev), None if hf is None else torch.empty(hf.shape, dtype=hf.dtype, device=dev))
                    for _ in range(2)]
            hl = torch.empty(B, Tc, dtype=torch.int32).pin_memory()
            h2 = torch.zeros(B, 1024, device=dev)
            ready = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]

            def issue_copy(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[i & 1])
                    dbuf[i & 1][0].copy_(hr, non_blocking=True)
                    if hf is not None:
                        dbuf[i & 1][1].copy_(hf, non_blocking=True)
                    ready[i & 1].record(copy_stream)

            def e2e_run(n):
                for ev in consumed:
                    ev.record(main_stream)
                issue_copy(0)
                for i in range(n):
                    if i + 1 < n:
                        issue_copy(i + 1)
                    main_stream.wait_event(ready[i & 1])
                    out = model.infer(dbuf[i & 1][0], dbuf[i & 1][1], h_state=h2, want_probs=False, precision=args.precision,
                                      chunk_T=min(Tc, args.subchunk), zero_flow=zf)
                    consumed[i & 1].record(main_stream)
                    hl.copy_(out["labels"], non_blocking=True)  # every step's labels go back to the host
                torch.cuda.synchronize()

            e2e_run(2)
            barrier()
            n_e2e = max(8, min(3 * K, 24))  # the first step cannot overlap its own staging: enough steps that the fill is ~4 %
            t0 = time.perf_counter()
            e2e_run(n_e2e)
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            h2d = hr.numel() * hr.element_size() + (0 if hf is None else hf.numel() * hf.element_size())
            return B * Tc * n_e2e / float(dt) * world, h2d, hl.numel() * 4, n_e2e

        def e2e_host_round(hr, hf, zf, direct):
            """The same loop with the operand rounding done on the HOST (prego_b200.ingest.HostRoundingStager): the fp32
            host features are rounded to the 16-bit operand format by host threads, slice by slice, and the link carries half
            the bytes; the device reads them in place (PREGO_FEAT_16).  `direct` leading streams travel as plain fp32 so that
            the link and the host's memory system finish together.  Bit-identical labels (tests/test_gpu_parity.py)."""
            from prego_b200.ingest import HostRoundingStager
            st = HostRoundingStager(B, Tc, 2048, 0 if hf is None else 2048, args.precision, dev,
                                    threads=pool_threads, direct_streams=direct)  # this rank's core block (bound below); the pool inherits the mask
            hl = torch.empty(B, Tc, dtype=torch.int32).pin_memory()
            dl = torch.empty(B, Tc, dtype=torch.int32, device=dev)
            h2 = torch.zeros(B, 1024, device=dev)

            def run(n):
                st.submit(0, hr, hf)
                for i in range(n):
                    if i + 1 < n:
                        st.submit(i + 1, hr, hf)
                    st.infer(model, i, h_state=h2, labels=dl, chunk_T=min(Tc, args.subchunk), zero_flow=zf)
                    hl.copy_(dl, non_blocking=True)
                torch.cuda.synchronize()

            run(2)
            barrier()
            n_e2e = max(8, min(3 * K, 24))  # the first step cannot overlap its own staging: enough steps that the fill is ~4 %
            t0 = time.perf_counter()
            run(n_e2e)
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            st.close()
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            n_dir = st.Bd * Tc * (2048 if hf is None else 4096)
            h2d = n_dir * 4 + (hr.numel() + (0 if hf is None else hf.numel()) - n_dir) * 2
            return B * Tc * n_e2e / float(dt) * world, h2d, st.threads

        # one rank = one contiguous block of the host's cores: the rank's Python thread, its rounding pool and (first touch)
        # its pinned buffers stay there instead of all ranks' threads migrating over all cores
        all_cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(all_cores) // world)
        my_cores = all_cores[local * per:(local + 1) * per] or all_cores
        try:
            os.sched_setaffinity(0, my_cores)
        except OSError:
            my_cores = all_cores
        # two cores of the block stay with the Python thread / the CUDA driver / the submit worker: with the pool on ALL cores the
        # launch path timeshares with the rounding threads and the loop loses ~10 % (measured r02: 4.5-4.7 M vs 4.9-5.7 M frames/s)
        pool_threads = len(my_cores) - 2 if len(my_cores) > 4 else len(my_cores)
        hr, hf = rgb.cpu().pin_memory(), flow.cpu().pin_memory()

        def host_bounds():
            """What the HOST side can deliver with every rank active at once: pinned H2D copy rate per GPU and a plain
            memcpy of the fp32 features by this rank's cores (the read the rounding pool has to do)."""
            dst = torch.empty_like(rgb)
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                dst.copy_(hr, non_blocking=True)
            torch.cuda.synchronize()
            h2d = torch.tensor([3 * hr.numel() * 4 / (time.perf_counter() - t0) / 1e9], device=dev)
            del dst
            tot_h2d = h2d.clone()
            if world > 1:
                dist.all_reduce(tot_h2d)
                dist.all_reduce(h2d, op=dist.ReduceOp.MIN)
            out = {"h2d_gbs_per_gpu_min": float(h2d), "h2d_gbs_all_gpus": float(tot_h2d), "cores_per_rank": len(my_cores),
                   "frames_per_s_bound_fp32_over_link": float(tot_h2d) * 1e9 / BYTES_FEATURES}
            if args.precision != "fp32":
                # the rounding pipeline alone (submit + wait, no model call): what this host can stage per second, all ranks at once
                from prego_b200.ingest import HostRoundingStager
                st = HostRoundingStager(B, Tc, 2048, 2048, args.precision, dev, threads=pool_threads)
                n = 6

                def stage_only(k):
                    st.submit(0, hr, hf)
                    for i in range(k):
                        if i + 1 < k:
                            st.submit(i + 1, hr, hf)
                        st.wait(i)
                        st.release(i)
                    torch.cuda.synchronize()

                stage_only(2)
                barrier()
                t0 = time.perf_counter()
                stage_only(n)
                dt = torch.tensor([time.perf_counter() - t0], device=dev)
                st.close()
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                out["frames_per_s_bound_host_rounded"] = world * B * Tc * n / float(dt)
                out["host_rounding_fp32_read_gbs_all_ranks"] = out["frames_per_s_bound_host_rounded"] * BYTES_FEATURES / 1e9
            out["note"] = ("measured with all ranks active at once, staging only (no model call); fp32 host features cost 16 KiB of host-memory reads "
                           "per frame on either path (the DMA's read, or the rounding pool's), so one host feeding N GPUs is bound by its own memory system")
            return out

        bounds = host_bounds()
        v, h2d, d2h, n_e2e = e2e_measure(hr, hf, False)
        e2e = {"value": v, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_e2e,
               "path": "fp32 over the link",
               "note": "pinned host fp32 features -> H2D (copy stream, double-buffered: step i+1 copies while step i computes) -> "
                       "prego_forward -> int32 labels D2H; PCIe-bound (16 KiB/frame); all ranks concurrently, max over ranks"}
        e2e["fp32_over_link"] = {"value": v, "h2d_bytes_per_step": h2d}
        e2e["host_bounds"] = bounds
        if args.precision != "fp32":
            v2, h2d2, nthreads = e2e_host_round(hr, hf, False, 0)
            e2e["host_rounded"] = {"value": v2, "h2d_bytes_per_step": h2d2, "host_threads": nthreads,
                                   "note": "same fp32 host buffers; host threads round them to the 16-bit operand format (the device's own first step, "
                                           "same rule: bit-identical results) through a 24 MiB pinned ring that stays in the CPU's last-level cache "
                                           "(DRAM only sees the fp32 read), pipelined with the H2D copies"}
            if v2 > v:
                e2e.update({"value": v2, "h2d_bytes_per_step": h2d2,
                            "path": "fp32 host buffers, operand rounding on the host, 8 KiB/frame over the link"})
            if args.e2e_direct > 0:
                # split every batch: `direct` leading streams travel as plain fp32 while the pool rounds the rest.  Off by
                # default: both paths read the same 16 KiB per frame out of host memory, which is what bounds this loop (measured
                # r02: 1 152 direct streams 4.12 M frames/s vs 4.72 M without a split on the same lease)
                direct = args.e2e_direct // 128 * 128
                v3, h2d3, _ = e2e_host_round(hr, hf, False, direct)
                e2e["host_rounded_split"] = {"value": v3, "h2d_bytes_per_step": h2d3, "direct_streams": direct,
                                             "note": "the same loop with `direct_streams` leading streams sent as fp32 (bit-identical labels)"}
                if v3 > e2e["value"]:
                    e2e.update({"value": v3, "h2d_bytes_per_step": h2d3,
                                "path": f"fp32 host buffers, {direct} streams as fp32 + {B - direct} rounded on the host"})
        if variants is not None:
            # the same loop fed in the declared ingest formats: the link carries 8 / 8 / 4 KiB per frame instead of 16
            del hf
            v, h2d, _, _ = e2e_measure(hr, None, True)
            variants["fp32_zero_flow"]["e2e_frames_per_s"], variants["fp32_zero_flow"]["h2d_bytes_per_step"] = v, h2d
            del hr
            hr16, hf16 = rgb16.cpu().pin_memory(), flow16.cpu().pin_memory()
            v, h2d, _, _ = e2e_measure(hr16, hf16, False)
            variants["feat16"]["e2e_frames_per_s"], variants["feat16"]["h2d_bytes_per_step"] = v, h2d
            v, h2d, _, _ = e2e_measure(hr16, None, True)
            variants["feat16_zero_flow"]["e2e_frames_per_s"], variants["feat16_zero_flow"]["h2d_bytes_per_step"] = v, h2d
            del hr16, hf16
        else:
            del hr, hf
    if variants is not None:
        del rgb16, flow16

    train = None
    if not args.no_train:
        del rgb, flow
        torch.cuda.empty_cache()
        train = training_leg(dev, world)
        rgb = flow = None
    if rank != 0:
        if world > 1:
            # keep ranks alive until rank 0 finished its extra legs
            dist.barrier()
            dist.destroy_process_group()
        return

    pk = peaks()
    Mc = B * Tc
    rows_per_launch = B * min(Tc, args.subchunk)  # one GEMM1 launch per internal time chunk
    g1_ms = prof["gemm1"]["ms"] / max(prof["gemm1"]["launches"], 1)
    achieved = FLOP_GEMM1 * rows_per_launch / (g1_ms * 1e-3) / 1e12
    phase_share = {p: round(v["ms"] / sum(x["ms"] for x in prof.values()), 4) for p, v in prof.items()}
    phase_tflops = {
        "gemm1": FLOP_GEMM1 * Mc * K / (prof["gemm1"]["ms"] * 1e-3) / 1e12,
        "gemm2": FLOP_GEMM2 * Mc * K / (prof["gemm2"]["ms"] * 1e-3) / 1e12,
        "recurrence": FLOP_REC * Mc * K / (prof["recurrence"]["ms"] * 1e-3) / 1e12,
    }
    stage_gbs = (BYTES_FEATURES + 8192) * Mc * K / (prof["stage"]["ms"] * 1e-3) / 1e9
    roofline = {"bound": "tensor", "kernel": "gemm_tc2_kernel<256,6,...> (Linear 4096->2048, tcgen05 cta_group::2 kind::f16)",
                "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_tflops_sustained"],
                **ncu_record(B, Tc, args),
                "algorithmic_bytes_per_launch": (8192 + 4096) * rows_per_launch + 4096 * 2048 * 2,
                "peak_source": f"{pk['source']} bf16 sustained (kernel timed inside a long step)",
                "per_launch_ms": g1_ms, "flops_per_launch": FLOP_GEMM1 * rows_per_launch,
                "phase_share": phase_share, "phase_tflops": phase_tflops,
                "stage_features_gbs": stage_gbs, "stage_frac_of_hbm": stage_gbs / pk["hbm_gbs"],
                "whole_step_tflops": (FLOP_GEMM1 + FLOP_GEMM2 + FLOP_REC + 2 * 1024 * 86) * value / world / 1e12}
    launches = sum(v["launches"] for v in prof.values()) + 3  # + window_mode, 2 x rle

    lat = None
    if not args.no_latency and world == 1:
        torch.cuda.empty_cache()
        lat = latency_leg(dev, args.precision)

    rank4 = None
    if not args.no_rank4 and world == 1:
        torch.cuda.empty_cache()
        rank4 = rank4_leg(dev)

    library = None
    if not args.no_library and world == 1:
        torch.cuda.empty_cache()
        library = library_baseline_leg(dev, B, Tc, local)

    if library is not None and train is not None and "train_step" in library:
        # same GPU, same process: this repo's training step against the stock-torch step of the same accuracy class
        lt = library["train_step"]
        pairs = {"tf32_vs_library_tf32": ("_tf32", "_tf32"), "tf32_vs_library_amp_fp16": ("_tf32", "_amp_fp16"),
                 "tf32x3_vs_library_as_shipped": ("_tf32x3", "_torch_defaults"), "exact_fp32_vs_library_as_shipped": ("", "_torch_defaults")}
        train["speedup_vs_library"] = {
            f"B{b}_{name}": lt[f"B{b}_T128{lk}"]["ms_per_step"] / train[f"B{b}_T128{ok}"]["ms_per_step"]
            for b in (16, 256) for name, (ok, lk) in pairs.items()
            if f"B{b}_T128{ok}" in train and "ms_per_step" in lt.get(f"B{b}_T128{lk}", {})}

    cpu = None
    if not args.no_cpu and world == 1:
        v, sample, cores = cpu_baseline_run(B, Tc)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample, "whole_video": cpu_whole_video_run()}

    line = {"metric": "frames/sec MiniROAD online inference", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": f"Assembly101-O-shaped MiniROAD eval, {B} concurrent streams/GPU x {Tc}-frame chunks with carried GRU state (K=86), + window-vote/RLE collapse",
                       "streams_per_gpu": B, "chunk_frames": Tc, "internal_subchunk": min(Tc, args.subchunk), "frames_per_step_per_gpu": Mc, "precision": args.precision,
                       "l2_policy": "inputs larger than L2 (4 GiB of features per step vs 126 MB L2)",
                       "gather": (f"collapsed step sequences of all ranks gathered on rank 0 inside the timed region ({gathered['sequences']} sequences, "
                                  f"{gathered['bytes']} bytes over NCCL)") if world > 1 else "single GPU: nothing to gather",
                       "weights": "seed-20 default init (no checkpoint ships with the reference)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "sustained": sustained,
            "single_stream": lat, "train_step": train, "feature_formats": variants, "rank4": rank4, "library_baseline": library}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL prints its version banner) must not pollute the ONE JSON line on stdout: route fd 1 to
    stderr for the duration of the run and keep the real stdout for the final line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    _quiet_stdout()
    args = parse()
    world, rank, local = dist_env(args.gpus)
    if args.impl == "reference":
        run_reference(args, world, rank)
        return
    run_ours(args, world, rank, local)


if __name__ == "__main__":
    main()
