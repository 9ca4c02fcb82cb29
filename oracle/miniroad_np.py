"""numpy restatement of the reference MiniROAD forward (TEST INFRASTRUCTURE).

Follows, equation by equation, the reference module
``step_recognition/model/rnn/rnn.py:51-71`` (``MROAD.forward``) whose
arithmetic lives in PyTorch ATen:

* ``rnn.py:52-57``  concat of rgb / flow features (``no_rgb`` / ``no_flow``)
* ``rnn.py:39-44,58``  Linear(D_in, E) -> LayerNorm(E, eps 1e-5, biased var)
  -> ReLU -> Dropout (identity in eval)
* ``rnn.py:38,60-61``  1-layer GRU, batch_first, h0 = 0, gate row blocks
  ordered r, z, n; ``b_hn`` inside the ``r * (...)`` term; state update in
  ATen's evaluation order ``h' = (h - n) * z + n``
* ``rnn.py:62-64``  ReLU -> Linear(H, K)
* ``rnn.py:65-71``  softmax in eval, raw logits in train
* ``trainer/eval.py:53``  label = ``np.argmax(probs, axis=1)`` (first max)

Pinning: the reference ships no test vectors for this boundary ("parity
unpinned" by the reference itself).  This restatement is pinned against the
reference module imported live from ``/root/reference`` on CPU by
``oracle/gen_golden.py``; the resulting vectors are committed under
``tests/golden/`` and re-checked by ``tests/test_oracle.py`` on every run.

The state-dict keys are the reference's own (``rnn.py:38-47``):
``layer1.0.{weight,bias}``, ``layer1.1.{weight,bias}``,
``gru.{weight_ih_l0,weight_hh_l0,bias_ih_l0,bias_hh_l0}``,
``f_classification.0.{weight,bias}``.
"""
from __future__ import annotations

import numpy as np

LN_EPS = 1e-5  # torch.nn.LayerNorm default, rnn.py:41


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _w(sd, key, dtype):
    v = sd[key]
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.asarray(v, dtype=dtype)


def embed(sd, x, dtype=np.float32):
    """layer1: Linear -> LayerNorm -> ReLU (rnn.py:39-44,58).  x: [M, D_in]."""
    w1 = _w(sd, "layer1.0.weight", dtype)
    b1 = _w(sd, "layer1.0.bias", dtype)
    g = _w(sd, "layer1.1.weight", dtype)
    b = _w(sd, "layer1.1.bias", dtype)
    y = x.astype(dtype) @ w1.T + b1
    mu = y.mean(-1, keepdims=True)
    var = ((y - mu) ** 2).mean(-1, keepdims=True)  # biased, as nn.LayerNorm
    yhat = (y - mu) / np.sqrt(var + dtype(LN_EPS)) * g + b
    return np.maximum(yhat, 0)


def gru_sequence(sd, e, h0=None, dtype=np.float32):
    """nn.GRU (1 layer, batch_first) over e: [B, T, E] -> ht [B, T, H], h_T.

    rnn.py:38,60-61.  Gate order r, z, n; ATen update order (h - n) * z + n.
    """
    w_ih = _w(sd, "gru.weight_ih_l0", dtype)
    w_hh = _w(sd, "gru.weight_hh_l0", dtype)
    b_ih = _w(sd, "gru.bias_ih_l0", dtype)
    b_hh = _w(sd, "gru.bias_hh_l0", dtype)
    B, T, _ = e.shape
    H = w_hh.shape[1]
    gi = e.astype(dtype) @ w_ih.T + b_ih  # [B, T, 3H]
    h = np.zeros((B, H), dtype) if h0 is None else np.asarray(h0, dtype).copy()
    out = np.empty((B, T, H), dtype)
    for t in range(T):
        gh = h @ w_hh.T + b_hh
        r = _sigmoid(gi[:, t, :H] + gh[:, :H])
        z = _sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
        h = ((h - n) * z + n).astype(dtype)
        out[:, t] = h
    return out, h


def head_logits(sd, ht, dtype=np.float32):
    """ReLU -> f_classification (rnn.py:45-47,62-64)."""
    wc = _w(sd, "f_classification.0.weight", dtype)
    bc = _w(sd, "f_classification.0.bias", dtype)
    return np.maximum(ht, 0).astype(dtype) @ wc.T + bc


def softmax(logits):
    m = logits.max(-1, keepdims=True)
    p = np.exp(logits - m)
    return p / p.sum(-1, keepdims=True)


def forward(sd, rgb, flow, *, use_rgb=True, use_flow=True, h0=None,
            training=False, dtype=np.float32, return_all=False):
    """MROAD.forward restated.  rgb/flow: [B, T, 2048] arrays.

    Returns the array the reference puts under ``out['logits']``: softmax
    probabilities in eval, raw logits in train (rnn.py:65-71).  With
    ``return_all`` also returns (logits, h_T).
    """
    rgb = np.asarray(rgb)
    flow = np.asarray(flow)
    if use_rgb and use_flow:
        x = np.concatenate((rgb, flow), axis=2)
    elif use_rgb:
        x = rgb
    else:
        x = flow
    B, T, D = x.shape
    e = embed(sd, x.reshape(B * T, D), dtype).reshape(B, T, -1)
    ht, h_last = gru_sequence(sd, e, h0, dtype)
    logits = head_logits(sd, ht, dtype)
    out = logits if training else softmax(logits)
    if return_all:
        return out, logits, h_last
    return out


def forward_anticipation(sd, rgb, flow, anticipation_length, *, use_rgb=True, use_flow=True, training=False,
                         dtype=np.float32):
    """MROADA.forward restated (``rnn.py:112-137``; registered "MiniROADA", ``rnn.py:73``).

    Same trunk as MROAD (``rnn.py:113-124``), then per frame
    ``ant_h = anticipation_layer.0(relu(h_t)).view(A, H)`` (``rnn.py:108-110,125``) and the SAME classifier on
    ``relu(ant_h)`` (``rnn.py:126``); softmax over classes in eval (``rnn.py:132-135``).
    Returns (out['logits'] [B,T,K], out['anticipation_logits'] [B,T,A,K], raw logits, raw anticipation logits).
    """
    rgb = np.asarray(rgb)
    flow = np.asarray(flow)
    x = np.concatenate((rgb, flow), axis=2) if (use_rgb and use_flow) else (rgb if use_rgb else flow)
    B, T, D = x.shape
    e = embed(sd, x.reshape(B * T, D), dtype).reshape(B, T, -1)
    ht, _ = gru_sequence(sd, e, None, dtype)
    logits = head_logits(sd, ht, dtype)
    wa = _w(sd, "anticipation_layer.0.weight", dtype)
    ba = _w(sd, "anticipation_layer.0.bias", dtype)
    H = ht.shape[-1]
    ant_h = (np.maximum(ht, 0).astype(dtype) @ wa.T + ba).reshape(B, T, anticipation_length, H)
    ant_logits = head_logits(sd, ant_h, dtype)
    if training:
        return logits, ant_logits, logits, ant_logits
    return softmax(logits), softmax(ant_logits), logits, ant_logits


def labels_from_probs(probs):
    """trainer/eval.py:53 -- np.argmax over the class axis, first max wins."""
    return np.argmax(probs, axis=-1)


def top2_margin(logits):
    """Gap between the best and the second-best logit per frame (near-tie metric)."""
    part = np.partition(logits, -2, axis=-1)
    return part[..., -1] - part[..., -2]
