"""Generate the LONG-sequence golden fixtures under tests/golden/ (TEST INFRASTRUCTURE).

The reference evaluates one whole video per forward (``test_batch_size: 1``, configs/*.yaml:17;
datasets/dataset.py:120-123; model/rnn/rnn.py:60-61): up to 31 114 frames for Epic-tent-O and 9 507 for
Assembly101-O.  This script runs the reference's own ``MROAD`` (imported unmodified from
/root/reference, CPU, fp32) on seeded synthetic videos of exactly those lengths and stores

  * ``labels``  int16 [T]   -- np.argmax of the reference probabilities for EVERY frame (trainer/eval.py:53)
  * ``margin``  fp32  [T]   -- top-1 minus top-2 reference logit per frame (the near-tie rule needs it everywhere)
  * ``frames``  int32 [S]   -- a strided subset of frames (every 97th + the last 64)
  * ``logits`` / ``probs`` fp32 [S, K] at those frames, ``h_last`` fp32 [H]

so the fixtures stay small (< 400 KB each) while drift of the 16-bit recurrence over 10^4 steps is still pinned at
the END of the sequence.  Runs only in the build container (the GPU box never sees /root/reference).

  python oracle/gen_golden_long.py        # rewrites tests/golden/long_*.npz and tests/golden/meta_long.json
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from prego_b200 import synthetic  # noqa: E402
from oracle.gen_golden import load_reference_model_pkg, ref_forward_all, sha  # noqa: E402

# name, base cfg, stream id, T  (Epic-tent-O mean / max length, Assembly101-O max length)
CASES = [
    ("epic_b1_t12531", "EPIC_TENT_O", 100, 12531),
    ("epic_b1_t31114", "EPIC_TENT_O", 101, 31114),
    ("asm_b1_t9507", "ASSEMBLY101_O", 102, 9507),
]


def subset(T: int) -> np.ndarray:
    return np.unique(np.concatenate([np.arange(0, T, 97), np.arange(max(T - 64, 0), T)])).astype(np.int32)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    build_model = load_reference_model_pkg()
    meta = {"torch": torch.__version__, "seed": 20, "cases": {}}
    for name, base, sid, T in CASES:
        cfg = dict(getattr(synthetic, base))
        torch.manual_seed(20)
        ref = build_model(cfg, "cpu").eval()
        mine = synthetic.seeded_model(cfg, seed=20)
        for k, v in ref.state_dict().items():
            assert torch.equal(v, mine.state_dict()[k]), f"seeded init differs from the reference for {k}"
        rgb, flow = synthetic.feature_batch([sid], T, "cpu", False)
        probs, logits, hT = ref_forward_all(ref, rgb, flow)
        probs, logits = probs[0], logits[0]
        labels = probs.argmax(-1)
        top2 = np.sort(logits, -1)[:, -2:]
        fr = subset(T)
        np.savez_compressed(os.path.join(GOLD, f"long_{name}.npz"), labels=labels.astype(np.int16),
                            margin=(top2[:, 1] - top2[:, 0]).astype(np.float32), frames=fr,
                            logits=logits[fr].astype(np.float32), probs=probs[fr].astype(np.float32),
                            h_last=hT[0].astype(np.float32), max_abs_logit=np.float32(np.abs(logits).max()))
        meta["cases"][name] = {"cfg": base, "stream_id": sid, "T": T, "rgb_sha256": sha(rgb), "flow_sha256": sha(flow),
                               "labels_sha256": hashlib.sha256(labels.astype(np.int16).tobytes()).hexdigest()}
        print(name, "labels", np.bincount(labels)[:8], "min margin", float((top2[:, 1] - top2[:, 0]).min()))
    json.dump(meta, open(os.path.join(GOLD, "meta_long.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
