"""Seeded inputs of the per-frame mAP golden cases (TEST INFRASTRUCTURE), shared by ``oracle/gen_golden_rank4.py`` and
the tests.  Scores are built with exact float32 arithmetic only (integers times powers of two), so every platform
regenerates them bit-identically; the generator stores their SHA-256 next to the expected values."""
from __future__ import annotations

import numpy as np


def map_cases():
    """-> [(name, scores fp32 [N, K] in [0, 1], labels int32 [N])] with the tie structures the algorithm must get right."""
    out = []
    rs = np.random.RandomState(11)
    # 1. fine-grained scores (20-bit), every class present
    N, K = 3000, 12
    s = (rs.randint(0, 1 << 20, (N, K)).astype(np.float32) / np.float32(1 << 20))
    out.append(("fine_k12", s, rs.randint(0, K, N).astype(np.int32)))
    # 2. heavy ties: scores quantised to 1/16, some classes absent, exact 0.0 and 1.0 present
    N, K = 5000, 20
    s = (rs.randint(0, 17, (N, K)).astype(np.float32) / np.float32(16))
    out.append(("ties_k20", s, rs.choice([0, 1, 2, 3, 5, 8, 13, 19], N).astype(np.int32)))
    # 3. tiny N, not a multiple of anything
    N, K = 37, 5
    s = (rs.randint(0, 1 << 10, (N, K)).astype(np.float32) / np.float32(1 << 10))
    out.append(("tiny_k5", s, np.r_[np.arange(K), rs.randint(0, K, N - K)].astype(np.int32)))
    # 4. all scores equal (one threshold) + a class whose positives all rank last
    N, K = 1025, 4
    s = np.full((N, K), 0.25, np.float32)
    s[:, 3] = np.arange(N, dtype=np.float32) / np.float32(1024)
    y = rs.randint(0, 3, N).astype(np.int32)
    y[:40] = 3
    out.append(("flat_k4", s, y))
    # 5. K = 86, several tiles per class, probabilities spread over 30 binades (softmax-like tails), ties inside
    N, K = 20000, 86
    mant = rs.randint(1, 1 << 8, (N, K)).astype(np.float32)
    expo = rs.randint(-38, -8, (N, K)).astype(np.float32)
    s = np.ldexp(mant, expo.astype(np.int32)).astype(np.float32)
    out.append(("binades_k86", s, rs.randint(0, K, N).astype(np.int32)))
    return out


def one_hot(labels, K):
    return np.eye(K, dtype=np.float32)[labels]
