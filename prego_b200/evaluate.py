"""Evaluation loop + JSON writer; drop-in for ``Evaluate`` of the reference
(``step_recognition/trainer/eval.py:15-84``, registered as EVAL["OAD"]).

Same call contract -- ``Evaluate(cfg)(model, dataloader, logger, device) -> mean AP`` with loader
items ``(rgb[1,T,Dr], flow[1,T,Df], target[1,T,K], vid, start, end)`` -- and the same side effect:
``output_miniRoad/output_miniROAD.json`` = ``{vid: {"pred": [...], "gt": [...]}}`` (eval.py:59-65),
written when ``cfg['eval'] is not None``.

Differences that do not change results: the per-frame argmax (eval.py:53) comes fused from the
head kernel (``model.last_labels``) instead of a host ``np.argmax`` over a D2H copy of T x K
probabilities, and the frames/s log line is computed from a real clock (the reference's is
broken).  The mAP itself (utils/metrics.py:25-62, sklearn, class 0 ignored) is host
code outside the hot path and is kept on the host.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.nn as nn

from .registry import EVAL


def perframe_average_precision(prediction, ground_truth, class_names, postprocessing=None, metrics="AP"):
    """utils/metrics.py:25-62: per-class frame-level AP, background class 0 ignored."""
    from sklearn.metrics import average_precision_score

    if metrics != "AP":
        raise RuntimeError(f"Unknown metrics: {metrics}")
    ground_truth = np.asarray(ground_truth)
    prediction = np.asarray(prediction)
    if postprocessing is not None:
        ground_truth, prediction = postprocessing(ground_truth, prediction)
    per_class = {}
    for idx, name in enumerate(class_names):
        if idx == 0:
            continue
        if np.any(ground_truth[:, idx]):
            per_class[name] = average_precision_score(ground_truth[:, idx], prediction[:, idx])
    return {"per_class_AP": per_class, "mean_AP": float(np.mean(list(per_class.values()))) if per_class else float("nan")}


@EVAL.register("OAD")
class Evaluate(nn.Module):
    OUTPUT_DIR = "output_miniRoad"           # eval.py:60 (sic)
    OUTPUT_FILE = "output_miniROAD.json"     # eval.py:64

    def __init__(self, cfg):
        super().__init__()
        self.metric = cfg["metric"]
        self.cfg = cfg
        if cfg.get("class_names") is not None:
            self.all_class_names = list(cfg["class_names"])
        elif cfg.get("video_list_path") and os.path.exists(cfg["video_list_path"]):
            self.all_class_names = json.load(open(cfg["video_list_path"]))[cfg["data_name"].split("_")[0]]["class_index"]
        else:
            self.all_class_names = [str(i) for i in range(cfg["num_classes"])]
        self.last_fps = None

    def eval(self, model, dataloader, logger, device):
        model.eval()
        output = {}
        pred_scores, gt_targets = [], []
        num_frames = 0
        t_start = time.perf_counter()
        with torch.no_grad():
            for rgb_input, flow_input, target, vid, _start, _end in dataloader:
                rgb_input = rgb_input.to(device, non_blocking=True)
                flow_input = flow_input.to(device, non_blocking=True)
                out_dict = model(rgb_input, flow_input)
                prob_val = out_dict["logits"].squeeze(0).cpu().numpy()        # eval.py:46
                target_batch = target.squeeze(0).cpu().numpy()
                pred_scores.append(prob_val)
                gt_targets.append(target_batch)
                num_frames += prob_val.shape[0]
                if self.cfg["eval"] is not None:
                    labels = getattr(model, "last_labels", None)
                    pred = labels.squeeze(0).cpu().numpy() if labels is not None else np.argmax(prob_val, axis=1)
                    gt = np.argmax(target_batch, axis=1)                       # eval.py:54
                    output[vid[0]] = {"pred": pred.tolist(), "gt": gt.tolist()}
        if self.cfg["eval"] is not None:
            os.makedirs(self.OUTPUT_DIR, exist_ok=True)
            with open(os.path.join(self.OUTPUT_DIR, self.OUTPUT_FILE), "w") as fp:
                json.dump(output, fp)
        elapsed = time.perf_counter() - t_start
        self.last_fps = num_frames / max(elapsed, 1e-9)
        result = perframe_average_precision(np.concatenate(pred_scores), np.concatenate(gt_targets),
                                            self.all_class_names, None, self.metric)
        if logger is not None:
            logger.info(f"Processed {num_frames} frames in {elapsed:.1f} seconds ({self.last_fps:.1f} FPS)")
        return result["mean_AP"]

    def forward(self, model, dataloader, logger, device):
        return self.eval(model, dataloader, logger, device)
