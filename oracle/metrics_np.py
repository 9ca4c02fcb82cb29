"""numpy restatement of the reference's per-frame mAP (TEST INFRASTRUCTURE).

Reference: ``step_recognition/utils/metrics.py:25-62`` (``perframe_average_precision`` with
``metrics == 'AP'``), called from ``trainer/eval.py:67-76`` (OAD) and ``eval.py:124-141``
(anticipation, once per anticipation step).  The arithmetic lives in a third-party dependency
that is not vendored: **scikit-learn** (``requirements.txt:9``, un-pinned; this image: 1.9.0),
``sklearn.metrics.average_precision_score`` for a binary target.  Its published algorithm
(``_binary_clf_curve`` -> ``precision_recall_curve`` -> step-function integral):

* sort frames by score, descending (stable);
* keep one threshold per DISTINCT score value (the last index of each run of equal scores);
* ``tps`` = cumulative positives at those indices, ``fps`` = ``1 + index - tps``;
* ``precision = tps / (tps + fps)``, ``recall = tps / tps[-1]``;
* ``AP = sum_n (recall_n - recall_{n-1}) * precision_n`` with ``recall_{-1} = 0``, in float64.

Pinning: ``oracle/gen_golden_rank4.py`` runs the reference's own ``perframe_average_precision``
(imported live from /root/reference, sklearn underneath) on seeded inputs with heavy score ties and
stores the per-class values under ``tests/golden/``; ``tests/test_oracle_rank4.py`` re-checks this
restatement against them.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


def average_precision(y_true, y_score) -> float:
    """sklearn.metrics.average_precision_score for one binary column (see the module docstring)."""
    y_true = np.asarray(y_true) != 0
    y_score = np.asarray(y_score)
    order = np.argsort(-y_score.astype(np.float64), kind="mergesort")
    ys = y_score[order]
    yt = y_true[order]
    ends = np.r_[np.where(np.diff(ys))[0], ys.size - 1]
    tps = np.cumsum(yt, dtype=np.float64)[ends]
    fps = 1.0 + ends - tps
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


def perframe_average_precision(prediction, ground_truth, class_names):
    """metrics.py:25-62: class 0 (background) ignored, classes without positives skipped."""
    ground_truth = np.asarray(ground_truth)
    prediction = np.asarray(prediction)
    per_class = OrderedDict()
    for idx, name in enumerate(class_names):
        if idx == 0:
            continue
        if np.any(ground_truth[:, idx]):
            per_class[name] = average_precision(ground_truth[:, idx], prediction[:, idx])
    return {"per_class_AP": per_class, "mean_AP": float(np.mean(list(per_class.values())))}
