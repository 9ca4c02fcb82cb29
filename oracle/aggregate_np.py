"""Restatement of the reference frame->step collapse (TEST INFRASTRUCTURE).

Follows ``utils/aggregate.py`` of the reference:

* ``aggregate.py:55,65-72``  every non-overlapping ``window_size = 200`` block
  of predicted labels (the last block may be shorter) is replaced by the
  block's mode, ``np.argmax(np.bincount(block))`` -> lowest label wins ties;
* ``aggregate.py:26-43``   ``find_changes``: indices i >= 1 with
  a[i] != a[i-1], then ``len(a)`` appended;
* ``aggregate.py:7-23``    ``eliminate_consecutive_duplicates``: run-length
  values; raises ``IndexError`` on an empty sequence (``arr[0]``);
* ``aggregate.py:75-78``   ground truth is collapsed WITHOUT the window vote;
* ``aggregate.py:81-90``   JSON: {vid: {"pred","gt","changes_pred","changes_gt"}}
  via ``json.dump`` with default separators.

Pinned against the reference's own golden pair
``output_miniRoad/output_miniROAD.json`` -> ``data/output/aggregated_data.json``
(sha256 47d7c7be...377c) and the known-answer table of SURVEY.md section 4
(see ``tests/test_oracle.py``).
"""
from __future__ import annotations

import json
from typing import Dict, List, Sequence, Tuple

import numpy as np

WINDOW = 200  # aggregate.py:55


def window_mode(labels: Sequence[int], window: int = WINDOW) -> np.ndarray:
    a = np.asarray(labels, dtype=np.int64)
    out = np.zeros_like(a)
    for s in range(0, len(a), window):
        blk = a[s:s + window]
        out[s:s + window] = np.argmax(np.bincount(blk))
    return out


def rle(a: Sequence[int]) -> Tuple[List[int], List[int]]:
    """(values, change indices incl. the final len(a)); IndexError if empty."""
    a = np.asarray(a, dtype=np.int64)
    if a.size == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    brk = np.flatnonzero(a[1:] != a[:-1]) + 1
    vals = a[np.concatenate(([0], brk))]
    return vals.tolist(), brk.tolist() + [int(a.size)]


def aggregate_video(pred: Sequence[int], gt: Sequence[int], window: int = WINDOW) -> Dict[str, List[int]]:
    if len(pred) == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    p_vals, p_chg = rle(window_mode(pred, window))
    g_vals, g_chg = rle(gt)
    return {"pred": p_vals, "gt": g_vals, "changes_pred": p_chg, "changes_gt": g_chg}


def aggregate_dict(data: Dict[str, Dict[str, Sequence[int]]], window: int = WINDOW):
    return {k: aggregate_video(v["pred"], v["gt"], window) for k, v in data.items()}


def aggregate(data, output_path: str) -> None:
    """Same signature and file layout as the reference ``aggregate`` (aggregate.py:46-90)."""
    with open(output_path, "w") as fp:
        json.dump(aggregate_dict(data), fp)
