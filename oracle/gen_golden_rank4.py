"""Golden fixtures for SURVEY 8f rank 4 (TEST INFRASTRUCTURE): the MiniROADA anticipation head and the per-frame mAP.

Runs ONLY in the build container (reference mounted read-only at /root/reference).  Imports, unmodified,
* the reference's ``MROADA`` (step_recognition/model/rnn/rnn.py:73-137) through its registry/builder, and
* the reference's ``perframe_average_precision`` (step_recognition/utils/metrics.py:25-62; scikit-learn underneath),
runs them on CPU on seeded synthetic inputs and stores their outputs under tests/golden/.

  python oracle/gen_golden_rank4.py       # rewrites tests/golden/anticipation_*.npz, map_cases.npz, meta_rank4.json
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "step_recognition"))

from prego_b200 import synthetic  # noqa: E402
from oracle.map_cases import map_cases, one_hot  # noqa: E402


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


ANT_CASES = [
    # name, base cfg, overrides, stream ids, T
    ("epic_a4_b2_t40", "EPIC_TENT_O", {"anticipation_length": 4, "actionness": False}, [70, 71], 40),
    ("asm_a2_b3_t24_act", "ASSEMBLY101_O", {"anticipation_length": 2, "actionness": True}, [72, 73, 74], 24),
    ("asm_a3_b20_t6", "ASSEMBLY101_O", {"anticipation_length": 3, "actionness": False}, list(range(80, 100)), 6),
]


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    from model import build_model  # reference registry (model/model_builder.py:7-9): "MiniROADA" -> MROADA
    from utils.metrics import perframe_average_precision as ref_pap
    from prego_b200.model import MROADA
    import sklearn
    meta = {"torch": torch.__version__, "sklearn": sklearn.__version__, "seed": 20, "anticipation": {}, "map": {}}
    for name, base, over, sids, T in ANT_CASES:
        cfg = dict(getattr(synthetic, base), model="MiniROADA", **over)
        torch.manual_seed(20)
        ref = build_model(cfg, "cpu").eval()
        torch.manual_seed(20)
        mine = MROADA(cfg)
        shas = {}
        assert list(ref.state_dict().keys()) == list(mine.state_dict().keys()), "state_dict keys / order differ"
        for k, v in ref.state_dict().items():
            assert torch.equal(v, mine.state_dict()[k]), f"seeded init differs from the reference for {k}"
            shas[k] = sha(v)
        rgb, flow = synthetic.feature_batch(sids, T, "cpu", False)
        with torch.no_grad():
            out = ref(rgb, flow)
            ref.train()  # train mode returns raw logits (rnn.py:128-130); dropout off for the capture
            ref.layer1[3].p = 0.0
            raw = ref(rgb, flow)
            ref.eval()
        np.savez_compressed(os.path.join(GOLD, f"anticipation_{name}.npz"),
                            probs=out["logits"].numpy(), ant_probs=out["anticipation_logits"].numpy(),
                            logits=raw["logits"].numpy(), ant_logits=raw["anticipation_logits"].numpy())
        meta["anticipation"][name] = {"cfg": base, "overrides": over, "stream_ids": sids, "T": T, "weights_sha256": shas,
                                      "rgb_sha256": sha(rgb), "flow_sha256": sha(flow)}
        print(name, tuple(out["anticipation_logits"].shape))

    arrays = {}
    for name, scores, labels in map_cases():
        K = scores.shape[1]
        onehot = one_hot(labels, K)
        assert scores.dtype == np.float32 and scores.min() >= 0 and scores.max() <= 1
        names = [str(i) for i in range(K)]
        with contextlib.redirect_stdout(io.StringIO()):  # the reference prints NUM FRAMES (metrics.py:51)
            r = ref_pap(list(scores), list(onehot), names, None, "AP")
        ap = np.full(K, np.nan)
        for k, v in r["per_class_AP"].items():
            ap[int(k)] = v
        arrays[f"{name}.ap"] = ap
        arrays[f"{name}.mean_ap"] = np.array(r["mean_AP"])
        meta["map"][name] = {"N": int(scores.shape[0]), "K": int(K), "mean_AP": float(r["mean_AP"]),
                             "scores_sha256": hashlib.sha256(scores.tobytes()).hexdigest(),
                             "labels_sha256": hashlib.sha256(labels.tobytes()).hexdigest(),
                             "classes_scored": len(r["per_class_AP"])}
        print(name, "mean_AP", r["mean_AP"], "classes", len(r["per_class_AP"]))
    np.savez_compressed(os.path.join(GOLD, "map_cases.npz"), **arrays)
    json.dump(meta, open(os.path.join(GOLD, "meta_rank4.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
