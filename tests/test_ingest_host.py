"""Host side of the feature ingest (prego_b200/ingest.py): conversion rules, the reference's
on-disk layout and zero-flow rule (datasets/dataset.py:45-95), length bucketing.  No GPU."""
import json
import os

import numpy as np
import pytest
import torch

from prego_b200 import ingest


def test_to_operand_matches_device_rounding_rules():
    x = np.array([0.1, -2.5, 70000.0, -1e9, 3.0e-8, 1.0009765625], dtype=np.float64)
    h = ingest.to_operand(x, torch.float16)
    assert h.dtype == torch.float16 and h.is_contiguous()
    assert float(h[2]) == 65504.0 and float(h[3]) == -65504.0          # saturating, like Op16<0>::pack2
    assert torch.equal(h[:2], torch.tensor([0.1, -2.5], dtype=torch.float32).half())
    b = ingest.to_operand(x, torch.bfloat16)
    assert b.dtype == torch.bfloat16 and torch.isfinite(b.float()).all()
    assert torch.equal(b, torch.from_numpy(x).float().bfloat16())
    f = ingest.to_operand(x, torch.float32)
    assert f.dtype == torch.float32


def test_bucket_order_longest_first_and_complete():
    lens = [5, 900, 17, 900, 33, 1, 64]
    b = ingest.bucket_order(lens, 3)
    assert [len(x) for x in b] == [3, 3, 1]
    flat = [i for x in b for i in x]
    assert sorted(flat) == list(range(len(lens)))
    assert [lens[i] for i in flat] == sorted(lens, reverse=True)
    assert b[0][:2] == [1, 3]  # ties keep input order


def _write_layout(root, cfg, vids, rng, flow_dir=False):
    os.makedirs(os.path.join(root, cfg["rgb_type"]), exist_ok=True)
    os.makedirs(os.path.join(root, cfg["annotation_type"]), exist_ok=True)
    data = {}
    for v, T in vids:
        rgb = np.abs(rng.standard_normal((T, 2048)))          # float64 on disk, as the reference's extractor wrote them
        tgt = np.eye(cfg["num_classes"])[rng.integers(0, cfg["num_classes"], T)]
        np.save(os.path.join(root, cfg["rgb_type"], v + ".npy"), rgb)
        np.save(os.path.join(root, cfg["annotation_type"], v + ".npy"), tgt)
        if flow_dir:
            d = os.path.join(root, cfg["flow_type"], "assembly_optical_flow_BNInception", v)
            os.makedirs(d, exist_ok=True)
            np.save(os.path.join(d, "assembling.npy"), np.abs(rng.standard_normal((T, 2048))))
        data[v] = (rgb, tgt)
    lst = os.path.join(root, "list.json")
    json.dump({cfg["data_name"]: {"test_session_set": [v for v, _ in vids] + ["missing_video"]}}, open(lst, "w"))
    cfg["root_path"], cfg["video_list_path"] = root, lst
    return data


def test_reference_layout_zero_flow_rule(tmp_path):
    cfg = dict(data_name="ASSEMBLY101-O", rgb_type="rgb_anet_resnet50", flow_type="flow_anet_resnet50", annotation_type="target_perframe",
               num_classes=7)
    data = _write_layout(str(tmp_path), cfg, [("a", 11), ("b", 40)], np.random.default_rng(0))
    store = ingest.FeatureStore.from_reference_layout(cfg, "fp16", pin=False)
    assert len(store) == 2 and store.zero_flow                  # the unreadable video is skipped, the dummy flow is not stored
    assert store.frames == 51 and store.host_bytes_per_frame() == 4096.0
    for v in store.videos:
        rgb, tgt = data[v.vid]
        assert v.rgb.dtype == torch.float16 and tuple(v.rgb.shape) == rgb.shape and v.flow is None
        assert torch.equal(v.rgb, torch.from_numpy(rgb).float().half())
        assert np.array_equal(v.gt, tgt.argmax(1))


def test_reference_layout_real_flow_and_mixed_store(tmp_path):
    cfg = dict(data_name="X", rgb_type="rgb_anet_resnet50", flow_type="flow_kinetics_bninception", annotation_type="target_perframe", num_classes=5)
    _write_layout(str(tmp_path), cfg, [("a", 9)], np.random.default_rng(1), flow_dir=True)
    store = ingest.FeatureStore.from_reference_layout(cfg, "bf16", pin=False)
    assert not store.zero_flow and store.videos[0].flow.dtype == torch.bfloat16 and store.host_bytes_per_frame() == 8192.0
    z = ingest.FeatureStore.from_arrays([("z", np.ones((4, 2048)), np.zeros((4, 2048)), None), ("r", np.ones((3, 2048)), np.ones((3, 2048)), None)],
                                        "fp16", pin=False)
    assert z.videos[0].flow is None and z.videos[1].flow is not None and not z.zero_flow


def test_streaming_needs_cuda():
    store = ingest.FeatureStore.from_arrays([("a", np.ones((4, 2048)), None, None)], "fp16", pin=False)
    with pytest.raises(RuntimeError):
        next(ingest.stream_batches(store, "cpu"))


def _adversarial_f32(n, seed=0):
    rs = np.random.RandomState(seed)
    x = rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32).view(np.float32).copy()  # every exponent, NaNs, infs
    x[:12] = [0.0, -0.0, 65504.0, 65520.0, 1e9, -1e9, np.inf, -np.inf, np.nan, 2.0 ** -25, 2.0 ** -24, 5.96e-8]
    return x


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_host_rounding_is_the_device_rule_bit_for_bit(prec):
    """prego_host_round_features == Op16<FMT>::from_float (csrc/gemm_tc.cuh): fp16 clamps to +-65504 (NaN -> -65504) and
    rounds to nearest even, bf16 rounds to nearest even; checked against torch's converters on every kind of fp32 bit
    pattern, on the SIMD path (1 and 5 threads), on the scalar path (short calls) and at unaligned offsets."""
    from prego_b200 import _lib
    lib = _lib.load()
    x = _adversarial_f32(1 << 20)
    t = torch.from_numpy(x)
    dt = ingest.OPERAND_DTYPES[prec]
    if prec == "fp16":
        want = torch.where(torch.isnan(t), torch.full_like(t, -65504.0), t).clamp(-65504.0, 65504.0).to(dt)
    else:
        want = t.to(dt)
    want = want.view(torch.int16).numpy().view(np.uint16)
    keep = ~np.isnan(x) if prec == "bf16" else np.ones(x.size, bool)  # bf16 NaN: any quiet NaN pattern
    for threads in (1, 5):
        got = torch.empty(x.size, dtype=dt)
        ingest.round_features_host(t, got, prec, threads)
        got = got.view(torch.int16).numpy().view(np.uint16)
        assert np.array_equal(got[keep], want[keep])
        assert np.all((got[~keep] & 0x7FFF) > 0x7F80)
    # scalar path: calls shorter than one SIMD vector, at odd offsets
    out = np.zeros(7, np.uint16)
    for off in range(0, 70000, 7):
        assert lib.prego_host_round_features(x[off:].ctypes.data, out.ctypes.data, 7, _lib.PRECISIONS[prec], 1) == 0
        k = keep[off:off + 7]
        assert np.array_equal(out[k], want[off:off + 7][k]), off
    assert lib.prego_host_round_impl() in (0, 1, 2)
    with pytest.raises(RuntimeError):
        ingest.round_features_host(t, torch.empty(x.size, dtype=torch.float32), "fp32")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_ring_stager_fails_cleanly_without_gpu():
    import ctypes as C
    from prego_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.prego_host_stager_create(2, 4, 1 << 20, C.byref(h)) == 2 and not h.value  # PREGO_ERR_CUDA: no pinned memory without a device
    assert lib.prego_host_stager_create(0, 4, 1 << 20, C.byref(h)) == 1                  # PREGO_ERR_INVALID
    assert lib.prego_host_stager_destroy(None) == 0
    with pytest.raises(RuntimeError, match="no CPU path"):
        ingest.HostRoundingStager(4, 4, 2048, 2048, "fp16", "cpu")


def test_host_all_zero_scan():
    from prego_b200 import _lib
    lib = _lib.load()
    n = (1 << 20) + 13
    x = np.zeros(n, np.float32)
    x[5] = -0.0
    for th in (1, 4):
        assert lib.prego_host_all_zero(x.ctypes.data, n, th) == 1          # +-0.0 only
    for pos in (0, 12345, n - 1):
        y = x.copy()
        y[pos] = 1e-45                                                     # smallest subnormal still counts
        for th in (1, 4):
            assert lib.prego_host_all_zero(y.ctypes.data, n, th) == 0, (pos, th)
    assert lib.prego_host_all_zero(x.ctypes.data, 0, 2) == 1 and lib.prego_host_all_zero(None, 4, 1) == -1
