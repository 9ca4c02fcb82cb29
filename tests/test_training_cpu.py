"""CPU-side checks of the training path: the torch gradient oracle against the reference-made golden
gradients, the criterion mirror, and the data-parallel gradient all-reduce (gloo, world_size 2)."""
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
