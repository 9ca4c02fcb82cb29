"""GPU-resident batched evaluation: many videos per forward, labels stay on the device into the aggregation
kernels, one compact D2H at the end (SURVEY 8f row 1 -- the caller of the hot path).

The reference evaluates one whole video per forward (``test_batch_size: 1``, trainer/eval.py:36-56), copies
T x K probabilities to the host per video and takes ``np.argmax`` there.  Because the GRU is causal, videos of
different lengths can share a batch: they are bucketed by length, zero-padded at the END (padding frames cannot
influence earlier outputs) and only the first T labels of each row are kept.  Results equal the per-video path.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .aggregate import aggregate_labels


@torch.no_grad()
def predict_labels(model, videos: Sequence[Tuple[str, torch.Tensor, torch.Tensor]], device, batch_streams: int = 64,
                   precision: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """videos: (vid, rgb[T, Dr], flow[T, Df]) fp32 tensors (host or device).  Returns {vid: int32 labels[T]} on
    ``device``.  Videos are sorted by length and processed ``batch_streams`` at a time."""
    model.eval()
    device = torch.device(device)
    order = sorted(range(len(videos)), key=lambda i: -int(videos[i][1].shape[0]))
    out: Dict[str, torch.Tensor] = {}
    for s in range(0, len(order), batch_streams):
        idx = order[s:s + batch_streams]
        tmax = int(videos[idx[0]][1].shape[0])
        dr, df = int(videos[idx[0]][1].shape[1]), int(videos[idx[0]][2].shape[1])
        rgb = torch.zeros(len(idx), tmax, dr, dtype=torch.float32, device=device)
        flow = torch.zeros(len(idx), tmax, df, dtype=torch.float32, device=device)
        for j, i in enumerate(idx):
            _, r, f = videos[i]
            t = int(r.shape[0])
            rgb[j, :t].copy_(r, non_blocking=True)
            flow[j, :t].copy_(f, non_blocking=True)
        labels = model.infer(rgb, flow, want_probs=False, want_labels=True, precision=precision)["labels"]
        for j, i in enumerate(idx):
            out[videos[i][0]] = labels[j, : int(videos[i][1].shape[0])]
    if hasattr(model, "check_device"):
        model.check_device("predict_labels")
    return {v[0]: out[v[0]] for v in videos}  # original order


def recognize_and_aggregate(model, videos, gts: Dict[str, Sequence[int]], device, batch_streams: int = 64,
                            precision: Optional[str] = None, out_dir: Optional[str] = None, window: int = 200):
    """Per-frame recognition + frame->step collapse for a set of videos, entirely on the device.

    Returns (frame_json, aggregated_json) with the reference's layouts; with ``out_dir`` also writes
    ``output_miniRoad/output_miniROAD.json`` (trainer/eval.py:59-65) and ``aggregated_data.json``
    (utils/aggregate.py:81-90) below it."""
    labels = predict_labels(model, videos, device, batch_streams, precision)
    vids = list(labels.keys())
    gt_list = [torch.as_tensor(gts[v]) for v in vids]
    agg = aggregate_labels([labels[v] for v in vids], gt_list, window=window, device=device)
    aggregated = dict(zip(vids, agg))
    frame_json = {v: {"pred": labels[v].cpu().tolist(), "gt": [int(x) for x in gts[v]]} for v in vids}
    if out_dir is not None:
        os.makedirs(os.path.join(out_dir, "output_miniRoad"), exist_ok=True)
        with open(os.path.join(out_dir, "output_miniRoad", "output_miniROAD.json"), "w") as fp:
            json.dump(frame_json, fp)
        with open(os.path.join(out_dir, "aggregated_data.json"), "w") as fp:
            json.dump(aggregated, fp)
    return frame_json, aggregated
