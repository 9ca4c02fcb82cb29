"""CPU-side checks of the training path: the torch gradient oracle against the reference-made golden
gradients, the criterion mirror, and the data-parallel gradient all-reduce (gloo, world_size 2)."""
import pytest
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLD
from oracle.miniroad_torch_cpu import TorchRefMROAD, oad_loss
from prego_b200 import synthetic
from prego_b200.training import OadLoss, allreduce_gradients


def _train_case(golden_meta):
    c = golden_meta["train_case"]
    rgb, flow = synthetic.feature_batch(c["stream_ids"], c["T"], "cpu", False)
    target = torch.stack([synthetic.targets(s, c["T"], 12) for s in c["stream_ids"]])
    return c, rgb, flow, target


def test_gradient_oracle_matches_reference(golden_meta):
    gold = np.load(os.path.join(GOLD, "train_epic_b3_t10.npz"))
    c, rgb, flow, target = _train_case(golden_meta)
    src = synthetic.seeded_model(dict(synthetic.EPIC_TENT_O, dropout=0.0), seed=20)
    port = TorchRefMROAD(4096, 2048, 1024, 12, 0.0).train()
    port.load_state_dict(src.state_dict())
    loss = oad_loss(port(rgb, flow)["logits"], target)
    loss.backward()
    assert abs(float(loss) - float(gold["loss"])) < 1e-5
    for k, p in port.named_parameters():
        g = p.grad.reshape(-1)
        assert np.allclose(g[:64].numpy(), gold[k + ".head"], rtol=1e-4, atol=1e-8), k
        st = np.array([g.sum().item(), g.abs().sum().item(), g.norm().item(), g.abs().max().item()])
        assert np.allclose(st[1:], gold[k + ".stats"][1:], rtol=1e-4), k


def test_criterion_mirror_equals_oracle(golden_meta):
    c, rgb, flow, target = _train_case(golden_meta)
    logits = torch.randn(3, 10, 12, generator=torch.Generator().manual_seed(0))
    a = OadLoss({"num_classes": 12})({"logits": logits}, target)
    assert abs(float(a) - float(oad_loss(logits, target))) < 1e-7


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = torch.nn.Linear(5, 3)
        for i, p in enumerate(m.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        allreduce_gradients(m)
        q.put((rank, [p.grad.flatten()[0].item() for p in m.parameters()]))
    finally:
        dist.destroy_process_group()


def test_allreduce_gradients_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # mean over ranks of (rank+1)*(i+1) = 1.5*(i+1)
    assert got[0] == got[1] == [1.5, 3.0]


def test_window_dataset_matches_reference_sampling():
    """WindowDataset reproduces dataset.py:53-55,77-82,113-119: window_size - 1 zero dummy frames in front of every
    video, windows [seed + k * stride, seed + k * stride + W) in the PADDED numbering while they fit, one random
    offset per video per (re-)initialisation, items in the reference's tuple layout."""
    import numpy as np
    from prego_b200 import WindowDataset

    class FixedRng:
        def __init__(self, vals): self.vals = list(vals)
        def randint(self, n): return self.vals.pop(0) % n

    T, W, S, K = 300, 128, 4, 7
    P = W - 1
    vids = {"a": (np.arange(T * 8, dtype=np.float64).reshape(T, 8), None, np.eye(K)[np.arange(T) % K]),
            "b": (np.ones((130, 8)), np.full((130, 8), 2.0), np.eye(K)[np.zeros(130, dtype=int)]),
            "short": (np.full((5, 8), 3.0), None, np.eye(K)[np.ones(5, dtype=int)])}   # shorter than the window
    ds = WindowDataset(vids, W, S, d_flow=8, rng=FixedRng([3, 1, 2, 0, 2, 1]))
    exp_a = list(zip(range(3, T + P, S), range(3 + W, T + P + 1, S)))
    exp_b = list(zip(range(1, 130 + P, S), range(1 + W, 130 + P + 1, S)))
    exp_s = list(zip(range(2, 5 + P, S), range(2 + W, 5 + P + 1, S)))
    assert [(v, s, e) for v, s, e in ds.inputs] == [("a", s, e) for s, e in exp_a] + [("b", s, e) for s, e in exp_b] + \
        [("short", s, e) for s, e in exp_s]
    assert len(exp_s) == 1 and exp_a[-1][1] <= T + P   # the front pad lets a 5-frame video yield a window
    rgb, flow, tgt, vid, start, end = ds[0]
    assert (vid, start, end) == ("a", 3, 131) and rgb.dtype == torch.float32 and tuple(rgb.shape) == (W, 8)
    assert float(flow.abs().sum()) == 0 and tuple(flow.shape) == (W, 8) and tuple(tgt.shape) == (W, K)
    # padded frames 3..126 are dummies (features and targets zero), padded frame 127 is the video's frame 0
    assert float(rgb[:P - 3].abs().sum()) == 0 and float(tgt[:P - 3].abs().sum()) == 0
    assert float(rgb[P - 3, 1]) == 1.0 and float(rgb[-1, 0]) == 24.0 and int(tgt[-1].argmax()) == 3
    fb = ds[len(exp_a)][1]
    assert float(fb[-1, 0]) == 2.0 and float(fb[0, 0]) == 0.0
    ds._init_features()   # new offsets, as main.py:101 does after every epoch
    assert ds.inputs[0] == ("a", 0, 128) and ds.inputs[-1][0] == "short" and ds.inputs[-1][1] == 1
    plain = WindowDataset(vids, W, S, d_flow=8, rng=FixedRng([3, 1, 0]), front_pad=False)
    assert plain.inputs[0] == ("a", 3, 131) and float(plain[0][0][0, 0]) == 24.0 and all(v != "short" for v, _, _ in plain.inputs)


_REF_DATASET_SCRIPT = r"""
import json, os, sys, types
import numpy as np, torch
sys.modules["ipdb"] = types.SimpleNamespace(set_trace=lambda *a, **k: None)   # the reference leaves breakpoints in (SURVEY 0.5)
sys.path.insert(0, "/root/reference/step_recognition")
sys.path.insert(0, sys.argv[2])
from datasets import build_dataset  # noqa: E402  (datasets/dataset_builder.py:11-13)
from prego_b200 import WindowDataset  # noqa: E402
root = sys.argv[1]
rs = np.random.RandomState(0)
K, W, S = 5, 16, 4
lengths = {"v0": 40, "v1": 9, "v2": 77}
for sub in ("target", "rgb_anet_resnet50", "rgb_as_flow/rgb_anet_resnet50"):
    os.makedirs(os.path.join(root, sub), exist_ok=True)
videos = {}
for vid, T in lengths.items():
    tgt = np.eye(K)[rs.randint(0, K, T)]
    rgb = rs.rand(T, 2048)
    np.save(os.path.join(root, "target", vid + ".npy"), tgt)
    np.save(os.path.join(root, "rgb_anet_resnet50", vid + ".npy"), rgb)
    np.save(os.path.join(root, "rgb_as_flow/rgb_anet_resnet50", vid + ".npy"), rgb)
    videos[vid] = (rgb, None, tgt)
json.dump({"EPIC-TENT-O": {"train_session_set": list(lengths), "test_session_set": []}}, open(os.path.join(root, "vl.json"), "w"))
cfg = dict(root_path=root, window_size=W, stride=S, data_name="EPIC-TENT-O", video_list_path=os.path.join(root, "vl.json"),
           num_classes=K, annotation_type="target", rgb_type="rgb_anet_resnet50", flow_type="flow_anet_resnet50")
np.random.seed(11)
ref = build_dataset(cfg)(cfg, "train")
np.random.seed(11)
mine = WindowDataset(videos, W, S, d_flow=2048)
assert len(ref) == len(mine) > 0, (len(ref), len(mine))
assert [(v, s, e) for v, s, e, _ in ref.inputs] == list(mine.inputs)
for i in range(len(ref)):
    a, b = ref[i], mine[i]
    for x, y in zip(a[:3], b[:3]):
        assert x.dtype == y.dtype and torch.equal(x, y), i
    assert tuple(a[3:]) == tuple(b[3:])
np.random.seed(12); ref._init_features()      # main.py:101
np.random.seed(12); mine._init_features()
assert [(v, s, e) for v, s, e, _ in ref.inputs] == list(mine.inputs)
print("OK", len(ref))
"""


@pytest.mark.skipif(not os.path.isdir("/root/reference/step_recognition"), reason="the reference checkout is only mounted in the build container")
def test_window_dataset_against_the_live_reference_dataset(tmp_path):
    """The reference's own THUMOSDataset (train mode) on .npy files written here vs WindowDataset on the same arrays:
    same (vid, start, end) list under the same numpy seed, bit-identical items, also after _init_features()."""
    import subprocess
    import sys
    script = tmp_path / "ref_ds.py"
    script.write_text(_REF_DATASET_SCRIPT)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, str(script), str(tmp_path / "data"), root], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_trainer_registry_and_optimizer_builder():
    from prego_b200 import TRAINER, build_trainer, train_one_epoch, FusedAdamW, build_optimizer
    assert build_trainer({"task": "OAD"}) is train_one_epoch and "OAD" in TRAINER
    lin = torch.nn.Linear(4, 4)
    opt = build_optimizer({"optimizer": "AdamW", "lr": 1e-4, "weight_decay": 0.05}, lin)
    assert isinstance(opt, FusedAdamW) and opt.param_groups[0]["lr"] == 1e-4 and opt.param_groups[0]["weight_decay"] == 0.05
    assert isinstance(build_optimizer({"optimizer": "Adam", "lr": 1e-4, "weight_decay": 0.0}, lin), torch.optim.Adam)
    lin.weight.grad = torch.zeros_like(lin.weight)
    lin.bias.grad = torch.zeros_like(lin.bias)
    with pytest.raises(RuntimeError):
        opt.step()   # CPU parameters: no fallback


def test_gradient_bucket_layout():
    """The flat gradient buffer of the overlapped all-reduce: bucket A (gru.*, f_classification.*: final first) and bucket B
    (layer1.*) are disjoint contiguous ranges that tile the buffer, every tensor sits 16-byte aligned inside its bucket."""
    from prego_b200 import gradient_buckets
    m = synthetic.seeded_model(dict(synthetic.ASSEMBLY101_O), seed=20)
    shapes = [tuple(t.shape) for t in m._param_tensors()]
    (a0, a1), (b0, b1), = gradient_buckets(shapes)[0]
    _, offsets, total = gradient_buckets(shapes)
    assert a0 == 0 and a1 == b0 and b1 == total
    numel = [int(np.prod(s)) for s in shapes]
    assert total >= sum(numel) and total - sum(numel) < 4 * len(shapes)
    spans = sorted((o, o + n) for o, n in zip(offsets, numel))
    assert all(o % 4 == 0 for o, _ in spans) and all(e <= o2 for (_, e), (o2, _) in zip(spans, spans[1:]))
    for i, (o, n) in enumerate(zip(offsets, numel)):
        lo, hi = (a0, a1) if i >= 4 else (b0, b1)   # _param_tensors(): layer1 first, then gru, then the classifier
        assert lo <= o and o + n <= hi
    assert sum(numel[4:]) == 3072 * 2048 + 3072 * 1024 + 2 * 3072 + 86 * 1024 + 86


def test_weight_repack_tracks_the_operand_formats(monkeypatch):
    """Host logic of MROAD._sync_weights (no GPU: the C call is recorded by a stub): inference packs every operand format
    once, a training step re-packs the fp32 set only after each in-place weight update, and the next inference call packs
    what went stale -- never nothing, never more than needed."""
    from prego_b200 import _lib
    from prego_b200.model import MROAD

    calls = []

    class StubLib:
        def prego_model_load_weights_ex(self, handle, w, formats, stream):
            calls.append(int(formats))
            return 0

    class StubStream:
        cuda_stream = 0

    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: StubStream())
    m = MROAD(dict(synthetic.EPIC_TENT_O))
    m._handle = object()
    lib, cpu = StubLib(), torch.device("cpu")
    m._sync_weights(lib, cpu)                       # first inference call: everything
    m._sync_weights(lib, cpu)                       # nothing changed: no call
    assert calls == [_lib.PACK_ALL] and m._packed_formats == _lib.PACK_ALL
    with torch.no_grad():
        m.layer1[0].weight.add_(1.0)                # optimizer step (in place: bumps the version counter)
    m._sync_weights(lib, cpu, _lib.PACK_F32)        # training forward: the fp32 set only
    m._sync_weights(lib, cpu, _lib.PACK_F32)
    assert calls == [_lib.PACK_ALL, _lib.PACK_F32] and m._packed_formats == _lib.PACK_F32
    m._sync_weights(lib, cpu)                       # inference again: the stale formats are packed (same weights)
    assert calls[-1] == _lib.PACK_ALL and m._packed_formats == _lib.PACK_ALL and len(calls) == 3
    m._sync_weights(lib, cpu, _lib.PACK_F32 | _lib.PACK_16)
    assert len(calls) == 3                          # a subset of what is packed: no call
