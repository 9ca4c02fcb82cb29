"""`main.py --eval` compatible entry point (reference: ``step_recognition/main.py:15-57``).

    python -m prego_b200.main --config <yaml> --eval <ckpt.pth> [--no_rgb] [--no_flow] [--synthetic N]

Same flow as the reference's eval branch -- flat YAML merged with argparse (main.py:28-30),
``set_seed(20)`` (main.py:32), ``build_model`` -> ``load_state_dict`` (main.py:44-48), ``build_eval`` ->
``evaluate(model, testloader, logger, device)`` (main.py:49-55), which writes
``output_miniRoad/output_miniROAD.json`` -- without the reference's landmines (hard-coded ``cuda:1``,
``ipdb`` breakpoints).  ``--synthetic N`` replaces the ``.npy`` feature files by N seeded
synthetic videos (no dataset ships with the repo); ``--eval synthetic`` keeps the seeded default weights.
Training (main.py:59-115) is not part of this entry point.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import os.path as osp
import random
import sys

import numpy as np
import torch

from . import synthetic
from .model import FEATURE_SIZES
from .registry import build_eval, build_model


def set_seed(seed):
    """utils/util.py:26-35."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


class EvalFeatures(torch.utils.data.Dataset):
    """Test-mode view of the reference dataset (datasets/dataset.py:30-95,120-132): one item per whole
    video, ``(rgb[T,Dr], flow[T,Df], target[T,K], vid, start, end)``; with ``flow_type ==
    'flow_anet_resnet50'`` the flow stream is all zeros, exactly as dataset.py:63-69 feeds it."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.root = cfg["root_path"]
        vids = json.load(open(cfg["video_list_path"]))[cfg["data_name"]]["test_session_set"]
        self.items = []
        for vid in vids:
            try:
                target = np.load(osp.join(self.root, cfg["annotation_type"], vid + ".npy"))
                rgb = np.load(osp.join(self.root, cfg["rgb_type"], vid + ".npy"))
                if cfg["flow_type"] == "flow_anet_resnet50":
                    flow = np.zeros((rgb.shape[0], FEATURE_SIZES[cfg["flow_type"]]), dtype=np.float32)
                else:
                    flow = np.load(osp.join(self.root, cfg["flow_type"], "assembly_optical_flow_BNInception", vid, "assembling.npy"))
                self.items.append((vid, rgb, flow, target))
            except Exception as e:  # the reference drops unreadable videos too (dataset.py:87-93)
                print("---- Exception in loading video ", e)

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        vid, rgb, flow, target = self.items[i]
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        return f32(rgb), f32(flow), f32(target), vid, 0, target.shape[0]


class SyntheticFeatures(torch.utils.data.Dataset):
    """N seeded synthetic videos of the configured shape (lengths cycle through the Assembly101-O /
    Epic-tent-O quantiles of SURVEY 6)."""
    LENGTHS = {"ASSEMBLY101-O": (538, 1024, 2011, 4096, 9507), "EPIC-TENT-O": (3702, 11734, 12531, 31114)}

    def __init__(self, cfg, n):
        self.cfg, self.n = cfg, n
        self.lengths = self.LENGTHS.get(cfg["data_name"], (1024,))

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        T = self.lengths[i % len(self.lengths)]
        zero_flow = self.cfg["flow_type"] == "flow_anet_resnet50"
        rgb, flow = synthetic.features(i, T, "cpu", zero_flow, FEATURE_SIZES[self.cfg["rgb_type"]], FEATURE_SIZES[self.cfg["flow_type"]])
        return rgb, flow, synthetic.targets(i, T, self.cfg["num_classes"]), f"synthetic_{i}", 0, T


def main(argv=None):
    import yaml

    parser = argparse.ArgumentParser()
    parser.add_argument("--config", type=str, required=True)
    parser.add_argument("--eval", type=str, default=None)
    parser.add_argument("--no_rgb", action="store_true")
    parser.add_argument("--no_flow", action="store_true")
    parser.add_argument("--synthetic", type=int, default=0, help="evaluate on N synthetic videos instead of .npy features")
    parser.add_argument("--device", type=str, default="cuda:0")
    parser.add_argument("--precision", type=str, default=None, choices=["fp16", "bf16", "fp32"])
    args = parser.parse_args(argv)

    cfg = yaml.load(open(args.config), Loader=yaml.FullLoader)
    cfg.update({k: v for k, v in vars(args).items() if k not in ("precision",) or v is not None})
    if args.eval is None:
        parser.error("this entry point implements the --eval branch (main.py:47-57); training is not built yet")
    set_seed(20)
    device = torch.device(args.device)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s")
    logger = logging.getLogger("prego_b200")
    logger.info(cfg)

    dataset = SyntheticFeatures(cfg, args.synthetic) if args.synthetic > 0 else EvalFeatures(cfg)
    testloader = torch.utils.data.DataLoader(dataset, batch_size=cfg.get("test_batch_size", 1), shuffle=False,
                                             num_workers=0, pin_memory=True)
    model = build_model(cfg, device)
    evaluate = build_eval(cfg)
    if args.eval != "synthetic":
        model.load_state_dict(torch.load(args.eval, map_location=device))
    mAP = evaluate(model, testloader, logger, device)
    logger.info(f'{cfg["task"]} result: {mAP * 100:.2f} m{cfg["metric"]}')
    return mAP


if __name__ == "__main__":
    main()
    sys.exit(0)
