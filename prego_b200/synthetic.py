"""Synthetic weights and features of the Assembly101-O / Epic-tent-O shapes.

No trained checkpoint ships with the reference, so parity and throughput use
seeded default-initialised weights (``set_seed(20)``, main.py:32) and |N(0,1)| features
(TSN/ResNet-50 pooled features are post-ReLU, hence non-negative).
"""
from __future__ import annotations

import torch

ASSEMBLY101_O = dict(model="MiniROAD", data_name="ASSEMBLY101-O", task="OAD", metric="AP",
                     rgb_type="rgb_anet_resnet50", flow_type="flow_anet_resnet50",
                     window_size=128, dropout=0.20, num_classes=86, embedding_dim=2048,
                     hidden_dim=1024, num_layers=1, no_rgb=False, no_flow=False, eval="synthetic")
EPIC_TENT_O = dict(ASSEMBLY101_O, data_name="EPIC-TENT-O", num_classes=12)


def seeded_model(cfg, seed=20, device=None):
    """Build the registered model under ``torch.manual_seed(seed)`` (CPU RNG, like the reference
    constructs it before ``.to(device)``), so weights are reproducible across machines."""
    from .registry import build_model
    torch.manual_seed(seed)
    m = build_model(cfg, None)
    if device is not None:
        m = m.to(device)
    return m.eval()


def features(stream_id: int, T: int, device="cpu", zero_flow=False, d_rgb=2048, d_flow=2048):
    """(rgb[T, d_rgb], flow[T, d_flow]) fp32, |N(0,1)|, seed 1000 + stream_id (CPU generator:
    identical on every machine with the same torch build)."""
    g = torch.Generator(device="cpu").manual_seed(1000 + stream_id)
    rgb = torch.randn(T, d_rgb, generator=g).abs_()
    flow = torch.zeros(T, d_flow) if zero_flow else torch.randn(T, d_flow, generator=g).abs_()
    return rgb.to(device), flow.to(device)


def feature_batch(stream_ids, T, device="cpu", zero_flow=False):
    pairs = [features(s, T, "cpu", zero_flow) for s in stream_ids]
    rgb = torch.stack([p[0] for p in pairs]).to(device)
    flow = torch.stack([p[1] for p in pairs]).to(device)
    return rgb, flow


def device_features(B, T, device, seed=1234, zero_flow=False):
    """Large synthetic batches generated directly on the device (throughput benches)."""
    g = torch.Generator(device=device).manual_seed(seed)
    rgb = torch.randn(B, T, 2048, generator=g, device=device).abs_()
    flow = torch.zeros(B, T, 2048, device=device) if zero_flow else torch.randn(B, T, 2048, generator=g, device=device).abs_()
    return rgb, flow


def targets(stream_id: int, T: int, K: int):
    """one-hot [T, K] of randint(0, K), seed 2000 + stream_id."""
    g = torch.Generator(device="cpu").manual_seed(2000 + stream_id)
    idx = torch.randint(0, K, (T,), generator=g)
    return torch.nn.functional.one_hot(idx, K).float()
