"""Frame -> step aggregation on the GPU; drop-in for the reference ``utils/aggregate.py``.

Same function signature (``aggregate(data, output_path)``, aggregate.py:46-90), same CLI
(``python -m prego_b200.aggregate <input_path> <output_path>``, aggregate.py:93-109) and the
same JSON layout ``{vid: {"pred", "gt", "changes_pred", "changes_gt"}}`` written with
``json.dump`` defaults, so the LLaMA anticipation branch reads it unchanged.

The 200-frame window mode vote (aggregate.py:55,65-72) and the run-length collapse
(aggregate.py:7-43) run as two CUDA kernels over the ragged batch of all videos at once
(``prego_window_mode`` / ``prego_rle`` of the C ABI).  Integer work: bit-exact.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, List, Sequence

import torch

from . import _lib

WINDOW_SIZE = 200  # aggregate.py:55 (hard-coded in the reference)


def _as_device_labels(seqs, device):
    """list of sequences (lists / numpy / torch, host or device) -> (int32 concat, lengths)."""
    lens = [int(len(s)) for s in seqs]
    parts = []
    for s in seqs:
        t = s if isinstance(s, torch.Tensor) else torch.as_tensor(s)
        if t.numel() and (t.dtype.is_floating_point or t.dtype == torch.bool):
            raise TypeError("labels must be integers")
        parts.append(t.to(device=device, dtype=torch.int32).reshape(-1))
    cat = torch.cat(parts) if parts else torch.empty(0, dtype=torch.int32, device=device)
    return cat, lens


def _offsets(lens, device):
    off = torch.zeros(len(lens) + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(torch.tensor(lens, dtype=torch.int64), 0)
    return off.to(device), off


def _rle_call(lib, seq, seg_off_dev, final_len_dev, B, scale, stream):
    n = int(seq.numel())
    vals = torch.empty(max(n, 1), dtype=torch.int32, device=seq.device)
    changes = torch.empty(max(n, 1), dtype=torch.int64, device=seq.device)
    counts = torch.empty(B, dtype=torch.int32, device=seq.device)
    _lib.check(lib.prego_rle(seq.data_ptr(), seg_off_dev.data_ptr(), final_len_dev.data_ptr(), B, scale,
                             vals.data_ptr(), changes.data_ptr(), counts.data_ptr(), stream), "prego_rle")
    return vals, changes, counts


def aggregate_device(pred: torch.Tensor, plens: List[int], gt: torch.Tensor, glens: List[int], window: int,
                     num_labels: int):
    """Device-resident core: concatenated int32 label tensors + per-video lengths in, device result
    tensors out (no host sync except none).  Returns a dict of device tensors and host offsets."""
    device = pred.device
    lib = _lib.load()
    B = len(plens)
    stream = torch.cuda.current_stream(device).cuda_stream
    p_off_dev, _ = _offsets(plens, device)
    g_off_dev, g_off = _offsets(glens, device)
    wlens = [(n + window - 1) // window for n in plens]
    w_off_dev, w_off = _offsets(wlens, device)
    total_windows = int(w_off[-1])
    modes = torch.empty(max(total_windows, 1), dtype=torch.int32, device=device)
    err = torch.zeros(1, dtype=torch.int32, device=device)
    _lib.check(lib.prego_window_mode(pred.data_ptr(), p_off_dev.data_ptr(), w_off_dev.data_ptr(), B,
                                     total_windows, window, num_labels, modes.data_ptr(), err.data_ptr(), stream),
               "prego_window_mode")
    plen_dev = torch.tensor(plens, dtype=torch.int64).to(device, non_blocking=True)
    glen_dev = torch.tensor(glens, dtype=torch.int64).to(device, non_blocking=True)
    pv, pc, pn = _rle_call(lib, modes, w_off_dev, plen_dev, B, window, stream)
    gv, gc, gn = _rle_call(lib, gt, g_off_dev, glen_dev, B, 1, stream)
    return {"pred_vals": pv, "pred_changes": pc, "pred_counts": pn, "pred_offsets": w_off,
            "gt_vals": gv, "gt_changes": gc, "gt_counts": gn, "gt_offsets": g_off, "err": err}


def aggregate_labels(preds, gts, window: int = WINDOW_SIZE, device=None, num_labels=None) -> List[Dict[str, List[int]]]:
    """Collapse a ragged batch of per-frame label sequences.  ``preds`` / ``gts``: lists of sequences
    (lists, numpy, torch; host or device) or 2-D integer tensors [B, T].  Returns one dict per video
    with the reference's four keys.  Raises IndexError on an empty video (aggregate.py:18 does)."""
    if len(preds) != len(gts):
        raise ValueError("preds and gts must have the same number of videos")
    B = len(preds)
    if B == 0:
        return []
    if device is None:
        if isinstance(preds, torch.Tensor) and preds.is_cuda:
            device = preds.device
        elif torch.cuda.is_available():
            device = torch.device("cuda", torch.cuda.current_device())
    if device is None or torch.device(device).type != "cuda":
        raise RuntimeError("prego_b200.aggregate runs on a CUDA (sm_100a) device; there is no CPU fallback")
    device = torch.device(device)
    with torch.cuda.device(device):
        if isinstance(preds, torch.Tensor) and preds.dim() == 2:
            pred, plens = preds.to(device=device, dtype=torch.int32).reshape(-1), [int(preds.shape[1])] * B
        else:
            pred, plens = _as_device_labels(preds, device)
        if isinstance(gts, torch.Tensor) and gts.dim() == 2:
            gt, glens = gts.to(device=device, dtype=torch.int32).reshape(-1), [int(gts.shape[1])] * B
        else:
            gt, glens = _as_device_labels(gts, device)
        if min(plens) == 0 or min(glens) == 0:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")
        if num_labels is None:
            lo, hi = int(pred.min()), int(pred.max())
            if lo < 0:
                raise ValueError("'list' argument must have no negative elements")  # np.bincount's message
            num_labels = hi + 1
        r = aggregate_device(pred, plens, gt, glens, window, num_labels)
        pv, pc, pn, gv, gc, gn, err = (r[k].cpu() for k in ("pred_vals", "pred_changes", "pred_counts", "gt_vals",
                                                             "gt_changes", "gt_counts", "err"))
    if int(err) != 0:
        raise ValueError("label outside [0, num_labels)")
    w_off, g_off = r["pred_offsets"], r["gt_offsets"]
    pv, pc, gv, gc = pv.tolist(), pc.tolist(), gv.tolist(), gc.tolist()
    pn, gn = pn.tolist(), gn.tolist()
    w_off, g_off = w_off.tolist(), g_off.tolist()
    out = []
    for b in range(B):
        ps, pk = w_off[b], pn[b]
        gs, gk = g_off[b], gn[b]
        out.append({"pred": pv[ps:ps + pk], "gt": gv[gs:gs + gk],
                    "changes_pred": pc[ps:ps + pk], "changes_gt": gc[gs:gs + gk]})
    return out


def aggregate_dict(data: Dict[str, Dict[str, Sequence[int]]], window: int = WINDOW_SIZE, device=None):
    keys = list(data.keys())
    res = aggregate_labels([data[k]["pred"] for k in keys], [data[k]["gt"] for k in keys], window, device)
    return dict(zip(keys, res))


def aggregate(data: Dict[str, Dict[str, Sequence[int]]], output_path: str) -> None:
    """aggregate.py:46-90 -- aggregate predictions / ground truth and save the JSON file."""
    aggregated = aggregate_dict(data)
    with open(output_path, "w") as fp:
        json.dump(aggregated, fp)


def main(argv=None):
    import argparse

    parser = argparse.ArgumentParser(description="Aggregate predictions and ground truth data.")
    parser.add_argument("input_path", type=str, help="Path to the input JSON file.")
    parser.add_argument("output_path", type=str, help="Path to save the aggregated JSON file.")
    args = parser.parse_args(argv)
    with open(args.input_path, "r") as fp:
        data = json.load(fp)
    aggregate(data, args.output_path)


if __name__ == "__main__":
    main()
