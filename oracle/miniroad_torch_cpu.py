"""CPU port of the reference MiniROAD forward on stock ATen ops (TEST INFRASTRUCTURE / CPU BASELINE).

The reference's arithmetic for this path lives in PyTorch ATen: ``nn.Linear``,
``nn.LayerNorm``, ``nn.GRU`` (oneDNN / MKL on CPU) and ``F.softmax``, composed by
``step_recognition/model/rnn/rnn.py:51-71``.  This port calls exactly those ATen kernels through
the functional API on the reference's ten state_dict tensors, so timing it on the host cores is
timing what ``main.py --eval`` executes per video on CPU (``cpu_baseline.kind = "port"``: the
reference's own Python files cannot travel to the GPU box).  It is also cross-checked against the
golden vectors in tests/test_oracle_torch.py.

Never imported by the product package.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


class CpuMiniROAD:
    def __init__(self, state_dict, use_rgb=True, use_flow=True):
        sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in state_dict.items()}
        self.sd = sd
        self.use_rgb, self.use_flow = use_rgb, use_flow
        H = sd["gru.weight_hh_l0"].shape[1]
        E = sd["gru.weight_ih_l0"].shape[1]
        self.gru = torch.nn.GRU(E, H, 1, batch_first=True)   # rnn.py:38
        with torch.no_grad():
            for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
                getattr(self.gru, k).copy_(sd["gru." + k])
        self.gru.eval()
        self.H = H

    @torch.no_grad()
    def forward(self, rgb, flow):
        """rnn.py:51-71 in eval mode -> softmax probabilities [B, T, K]."""
        sd = self.sd
        if self.use_rgb and self.use_flow:
            x = torch.cat((rgb, flow), 2)
        else:
            x = rgb if self.use_rgb else flow
        x = F.linear(x, sd["layer1.0.weight"], sd["layer1.0.bias"])
        x = F.layer_norm(x, (x.shape[-1],), sd["layer1.1.weight"], sd["layer1.1.bias"], 1e-5)
        x = F.relu(x)
        h0 = torch.zeros(1, x.shape[0], self.H)
        ht, _ = self.gru(x, h0)
        logits = F.linear(F.relu(ht), sd["f_classification.0.weight"], sd["f_classification.0.bias"])
        return F.softmax(logits, dim=-1)

    def labels(self, rgb, flow):
        """trainer/eval.py:46-53: .cpu().numpy() then np.argmax(axis=-1)."""
        return np.argmax(self.forward(rgb, flow).numpy(), axis=-1)


class TorchRefMROAD(torch.nn.Module):
    """Train-capable restatement of the reference module on stock torch.nn layers (rnn.py:18-71): the same
    sub-modules, names and creation order, so ``state_dict`` keys and seeded init coincide.  Used as the
    gradient oracle for the CUDA training step (autograd through ATen = what ``loss.backward()`` does in
    trainer/train.py:23).  Checked against the live reference in oracle/gen_golden.py."""

    def __init__(self, input_dim, embedding_dim, hidden_dim, num_classes, dropout):
        super().__init__()
        self.gru = torch.nn.GRU(embedding_dim, hidden_dim, 1, batch_first=True)
        self.layer1 = torch.nn.Sequential(torch.nn.Linear(input_dim, embedding_dim), torch.nn.LayerNorm(embedding_dim),
                                          torch.nn.ReLU(), torch.nn.Dropout(p=dropout))
        self.f_classification = torch.nn.Sequential(torch.nn.Linear(hidden_dim, num_classes))
        self.hidden_dim = hidden_dim

    def forward(self, rgb, flow):
        x = self.layer1(torch.cat((rgb, flow), 2))
        h0 = torch.zeros(1, x.shape[0], self.hidden_dim, device=x.device, dtype=x.dtype)
        ht, _ = self.gru(x, h0)
        logits = self.f_classification(F.relu(ht))
        return {"logits": logits if self.training else F.softmax(logits, dim=-1)}


def oad_loss(logits, target):
    """criterions/loss.py:15-34 (NONUNIFORM): last-frame CE with L2-normalised targets, batch mean."""
    lg, tg = logits[:, -1, :], target[:, -1, :]
    return torch.sum(-F.normalize(tg) * F.log_softmax(lg, dim=-1), dim=1).mean()
