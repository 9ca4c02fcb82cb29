"""Evaluation loop + JSON writer; drop-in for ``Evaluate`` of the reference
(``step_recognition/trainer/eval.py:15-84``, registered as EVAL["OAD"]).

Same call contract -- ``Evaluate(cfg)(model, dataloader, logger, device) -> mean AP`` with loader
items ``(rgb[1,T,Dr], flow[1,T,Df], target[1,T,K], vid, start, end)`` -- and the same side effect:
``output_miniRoad/output_miniROAD.json`` = ``{vid: {"pred": [...], "gt": [...]}}`` (eval.py:59-65),
written when ``cfg['eval'] is not None``.

Differences that do not change results: the per-frame argmax (eval.py:53) comes fused from the
head kernel (``model.last_labels``) instead of a host ``np.argmax`` over a D2H copy of T x K
probabilities, the frames/s log line is computed from a real clock (the reference's is
broken), and the probabilities never leave the device: the per-frame mAP
(utils/metrics.py:25-62, sklearn on the host in the reference) is computed by ``prego_perframe_ap``
(``prego_b200/metrics.py``), so the only D2H traffic is the labels for the JSON and K doubles.

``ANT_Evaluate`` (EVAL["ANTICIPATION"], eval.py:85-163) is the same loop for MiniROADA: one OAD mAP plus one
mAP per anticipation step, returning their mean.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.nn as nn

from .metrics import perframe_average_precision  # utils/metrics.py:25-62 on the device
from .registry import EVAL


def _is_zero_flow(model, flow_input) -> bool:
    """The loader's flow tensor is still on the host: one cheap scan tells whether it is the reference's all-zero dummy."""
    return (hasattr(model, "infer") and getattr(model, "use_flow", False) and getattr(model, "use_rgb", False)
            and isinstance(flow_input, torch.Tensor) and not flow_input.is_cuda and not bool(flow_input.any()))


def _class_names(cfg):
    if cfg.get("class_names") is not None:
        return list(cfg["class_names"])
    if cfg.get("video_list_path") and os.path.exists(cfg["video_list_path"]):
        return json.load(open(cfg["video_list_path"]))[cfg["data_name"].split("_")[0]]["class_index"]
    return [str(i) for i in range(cfg["num_classes"])]


@EVAL.register("OAD")
class Evaluate(nn.Module):
    OUTPUT_DIR = "output_miniRoad"           # eval.py:60 (sic)
    OUTPUT_FILE = "output_miniROAD.json"     # eval.py:64

    def __init__(self, cfg):
        super().__init__()
        self.metric = cfg["metric"]
        self.cfg = cfg
        self.all_class_names = _class_names(cfg)
        self.last_fps = None
        self.last_result = None

    def eval(self, model, dataloader, logger, device):
        model.eval()
        output = {}
        pred_scores, gt_targets = [], []
        num_frames = 0
        t_start = time.perf_counter()
        with torch.no_grad():
            for rgb_input, flow_input, target, vid, _start, _end in dataloader:
                rgb_input = rgb_input.to(device, non_blocking=True)
                if _is_zero_flow(model, flow_input):
                    # both shipped configs feed the all-zero flow dummy (datasets/dataset.py:63-69): it is neither copied
                    # nor multiplied (flow_is_zero: bit-identical to passing the zeros, tests/test_gpu_parity.py)
                    out = model.infer(rgb_input, None, want_probs=True, want_labels=True, zero_flow=True)
                    model.last_labels = out["labels"]
                    out_dict = {"logits": out["probs"]}
                else:
                    flow_input = flow_input.to(device, non_blocking=True)
                    out_dict = model(rgb_input, flow_input)
                prob_val = out_dict["logits"].squeeze(0)                       # eval.py:46, kept on the device
                target_dev = target.squeeze(0).to(device, non_blocking=True)
                pred_scores.append(prob_val)
                gt_targets.append(target_dev)
                num_frames += prob_val.shape[0]
                if self.cfg["eval"] is not None:
                    labels = getattr(model, "last_labels", None)
                    pred = labels.squeeze(0) if labels is not None else torch.argmax(prob_val, dim=1)
                    gt = torch.argmax(target_dev, dim=1)                       # eval.py:54 (first max, as np.argmax)
                    output[vid[0]] = {"pred": pred.cpu().tolist(), "gt": gt.cpu().tolist()}
        if self.cfg["eval"] is not None:
            os.makedirs(self.OUTPUT_DIR, exist_ok=True)
            with open(os.path.join(self.OUTPUT_DIR, self.OUTPUT_FILE), "w") as fp:
                json.dump(output, fp)
        elapsed = time.perf_counter() - t_start
        self.last_fps = num_frames / max(elapsed, 1e-9)
        result = perframe_average_precision(torch.cat(pred_scores), torch.cat(gt_targets),
                                            self.all_class_names, None, self.metric)
        self.last_result = result
        if logger is not None:
            logger.info(f"Processed {num_frames} frames in {elapsed:.1f} seconds ({self.last_fps:.1f} FPS)")
        return result["mean_AP"]

    def forward(self, model, dataloader, logger, device):
        return self.eval(model, dataloader, logger, device)


@EVAL.register("ANTICIPATION")
class ANT_Evaluate(nn.Module):
    """eval.py:85-163: loader items ``(rgb[1,T,Dr], flow[1,T,Df], target[1,T,K], ant_target[1,T,A,K])``; logs the OAD
    mAP and one mAP per anticipation step, returns the mean anticipation mAP.  ``device`` defaults to the
    reference's hard-coded "cuda:0" (eval.py:99)."""

    def __init__(self, cfg):
        super().__init__()
        self.metric = cfg["metric"]
        self.cfg = cfg
        self.all_class_names = _class_names(cfg)
        self.last_fps = None
        self.last_result = None

    def eval(self, model, dataloader, logger, device="cuda:0"):
        model.eval()
        pred_scores, gt_targets, ant_pred_scores, ant_gt_targets = [], [], [], []
        t_start = time.perf_counter()
        with torch.no_grad():
            for rgb_input, flow_input, target, ant_target in dataloader:
                rgb_input = rgb_input.to(device, non_blocking=True)
                if _is_zero_flow(model, flow_input) and getattr(model, "anticipation_length", 0) > 0:
                    out = model.infer(rgb_input, None, want_probs=True, want_anticipation=True, zero_flow=True)  # as in Evaluate.eval
                    out_dict = {"logits": out["probs"], "anticipation_logits": out["anticipation_probs"]}
                else:
                    flow_input = flow_input.to(device, non_blocking=True)
                    out_dict = model(rgb_input, flow_input)
                K = out_dict["logits"].shape[-1]
                A = out_dict["anticipation_logits"].shape[-2]
                pred_scores.append(out_dict["logits"].reshape(-1, K))                      # eval.py:119-124
                gt_targets.append(target.to(device, non_blocking=True).reshape(-1, K))
                ant_pred_scores.append(out_dict["anticipation_logits"].reshape(-1, A, K))
                ant_gt_targets.append(ant_target.to(device, non_blocking=True).reshape(-1, A, K))
        elapsed = time.perf_counter() - t_start
        pred, gt = torch.cat(pred_scores), torch.cat(gt_targets)
        ant_pred, ant_gt = torch.cat(ant_pred_scores), torch.cat(ant_gt_targets)
        num_frames = int(pred.shape[0])
        result = perframe_average_precision(pred, gt, self.all_class_names, None, self.metric)
        if logger is not None:
            logger.info(f'OAD mAP: {result["mean_AP"]*100:.2f}')
        anticipation_mAPs = []
        for step in range(ant_gt.shape[1]):                                                # eval.py:141-153
            r = perframe_average_precision(ant_pred[:, step, :].contiguous(), ant_gt[:, step, :].contiguous(),
                                           self.all_class_names, None, self.metric)
            result[f"anticipation_{step+1}"] = r
            anticipation_mAPs.append(r["mean_AP"])
            if logger is not None:
                logger.info(f"Anticipation at step {step+1}: {r['mean_AP']*100:.2f}")
        self.last_fps = num_frames / max(elapsed, 1e-9)
        self.last_result = result
        if logger is not None:
            logger.info(f"Mean Anticipation mAP: {np.mean(anticipation_mAPs)*100:.2f}")
            logger.info(f"Processed {num_frames} frames in {elapsed:.1f} seconds ({self.last_fps:.1f} FPS)")
        return np.mean(anticipation_mAPs)

    def forward(self, model, dataloader, logger, device="cuda:0"):
        return self.eval(model, dataloader, logger, device)
