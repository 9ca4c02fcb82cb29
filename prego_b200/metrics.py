"""Per-frame mean average precision on the device; drop-in for ``perframe_average_precision``
(``step_recognition/utils/metrics.py:25-62``, ``metrics='AP'``) as called by the evaluators
(``trainer/eval.py:67-76`` and ``eval.py:124-141``).

The reference hands [N, K] probabilities and [N, K] one-hot targets to
``sklearn.metrics.average_precision_score`` once per class on the host.  Here the scores stay on
the device: ``prego_perframe_ap`` (``csrc/metrics.cuh``) radix-sorts each class's frames with one CTA per
class and integrates the precision-recall step function in float64; only K doubles come back.
Same result layout (``per_class_AP`` OrderedDict in class order without the background class 0 and
without classes that have no positive frame, ``mean_AP`` = ``np.mean`` of its values).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import _lib


def average_precision_per_class(scores: torch.Tensor, targets: torch.Tensor):
    """scores: CUDA fp32 [N, K] probabilities in [0, 1]; targets: CUDA fp32 [N, K] (non-zero = positive) or an
    integer label vector [N].  Returns (ap float64[K] with NaN where a class has no positives, num_pos int64[K]) as
    numpy arrays."""
    if not isinstance(scores, torch.Tensor) or not scores.is_cuda:
        raise RuntimeError("prego_b200.metrics runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    if scores.dim() != 2 or scores.dtype != torch.float32:
        raise RuntimeError(f"scores must be fp32 [N, K], got {scores.dtype} {tuple(scores.shape)}")
    device = scores.device
    N, K = int(scores.shape[0]), int(scores.shape[1])
    if N == 0:
        raise ValueError("perframe_average_precision needs at least one frame")
    scores = scores.contiguous()
    t_onehot = t_labels = None
    if targets.dim() == 2:
        if tuple(targets.shape) != (N, K):
            raise RuntimeError(f"targets must be [{N}, {K}], got {tuple(targets.shape)}")
        t_onehot = targets.to(device=device, dtype=torch.float32).contiguous()
    else:
        if tuple(targets.shape) != (N,):
            raise RuntimeError(f"target labels must be [{N}], got {tuple(targets.shape)}")
        t_labels = targets.to(device=device, dtype=torch.int32).contiguous()
    lib = _lib.load()
    with torch.cuda.device(device):
        need = lib.prego_ap_workspace_bytes(N, K)
        ws = torch.empty(need + 1024, dtype=torch.uint8, device=device)
        ws_ptr = ws.data_ptr() + (-ws.data_ptr()) % 1024
        ap = torch.empty(K, dtype=torch.float64, device=device)
        num_pos = torch.empty(K, dtype=torch.int64, device=device)
        err = torch.zeros(1, dtype=torch.int32, device=device)
        _lib.check(lib.prego_perframe_ap(scores.data_ptr(), t_onehot.data_ptr() if t_onehot is not None else None,
                                         t_labels.data_ptr() if t_labels is not None else None, N, K, ap.data_ptr(),
                                         num_pos.data_ptr(), ws_ptr, need, err.data_ptr(),
                                         torch.cuda.current_stream(device).cuda_stream), "prego_perframe_ap")
        if int(err.item()) != 0:
            raise ValueError("scores must be probabilities in [0, 1] (found a negative, > 1 or NaN value)")
        return ap.cpu().numpy(), num_pos.cpu().numpy()


def perframe_average_precision(prediction, ground_truth, class_names, postprocessing=None, metrics="AP"):
    """utils/metrics.py:25-62.  prediction / ground_truth: CUDA tensors [N, K] (ground_truth may be int labels [N])."""
    if metrics != "AP":
        raise RuntimeError("Unknown metrics: {}".format(metrics))  # 'cAP' (TVSeries) is outside the hot path
    if postprocessing is not None:
        raise RuntimeError("postprocessing (THUMOS) is outside the hot path")
    ap, num_pos = average_precision_per_class(prediction, ground_truth)
    pred_mass = prediction.sum(dim=0, dtype=torch.float64).cpu().numpy()   # metrics.py:58 prints int(np.sum(prediction[:, idx]))
    result = OrderedDict()
    result["per_class_AP"] = OrderedDict()
    result["num"] = OrderedDict()
    for idx, class_name in enumerate(class_names):
        if idx == 0:  # background class ignored (metrics.py:47,53)
            continue
        if num_pos[idx] > 0:
            result["per_class_AP"][class_name] = float(ap[idx])
            # the reference's log string (metrics.py:58); 'true' = sum of the one-hot column = number of positive frames
            result["num"][class_name] = f"[true: {int(num_pos[idx])}, pred:{int(pred_mass[idx])}, AP:{float(ap[idx]) * 100:.1f}]"
    vals = list(result["per_class_AP"].values())
    result["mean_AP"] = float(np.mean(vals)) if vals else float("nan")
    return result
