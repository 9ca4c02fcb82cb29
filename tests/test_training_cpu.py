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
    """WindowDataset reproduces dataset.py:113-119: windows [seed + k * stride, seed + k * stride + W) while they fit,
    one random offset per video per (re-)initialisation, items in the reference's tuple layout."""
    import numpy as np
    from prego_b200 import WindowDataset

    class FixedRng:
        def __init__(self, vals): self.vals = list(vals)
        def randint(self, n): return self.vals.pop(0) % n

    T, W, S, K = 300, 128, 4, 7
    vids = {"a": (np.arange(T * 8, dtype=np.float64).reshape(T, 8), None, np.eye(K)[np.arange(T) % K]),
            "b": (np.ones((130, 8)), np.full((130, 8), 2.0), np.eye(K)[np.zeros(130, dtype=int)])}
    ds = WindowDataset(vids, W, S, d_flow=8, rng=FixedRng([3, 1, 0, 2]))
    exp_a = list(zip(range(3, T, S), range(3 + W, T + 1, S)))
    exp_b = list(zip(range(1, 130, S), range(1 + W, 131, S)))
    assert [(v, s, e) for v, s, e in ds.inputs] == [("a", s, e) for s, e in exp_a] + [("b", s, e) for s, e in exp_b]
    assert len(exp_b) == 1 and exp_a[-1][1] <= T
    rgb, flow, tgt, vid, start, end = ds[0]
    assert (vid, start, end) == ("a", 3, 131) and rgb.dtype == torch.float32 and tuple(rgb.shape) == (W, 8)
    assert float(flow.abs().sum()) == 0 and tuple(flow.shape) == (W, 8) and tuple(tgt.shape) == (W, K)
    assert float(rgb[0, 0]) == 24.0
    assert float(ds[len(exp_a)][1][0, 0]) == 2.0
    ds._init_features()   # new offsets (0 and 2), as main.py:101 does after every epoch
    assert ds.inputs[0] == ("a", 0, 128) and ds.inputs[-1][0] == "b" and ds.inputs[-1][1] == 2


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
