"""CPU ORACLE, second opinion (test infrastructure — NOT part of the product path).

Float64 numpy restatement of the same hot path written as explicit per-destination / per-edge loops straight
from the equations of SURVEY.md Appendix A.1–A.3, sharing no code (and no scatter/segment primitives) with
``oracle/gnn_oracle.py``.  Only small cases: it exists to pin the vectorised torch oracle, which in turn
checks the CUDA kernels.  Parity with the reference itself is unpinned (DGL 0.9.0 not installable; see
``gnn_oracle.py`` header).
"""
from __future__ import annotations

import numpy as np


def gatv2_conv_loops(src, dst, n_dst, x_src, x_dst, Ws, bs, Wd, bd, attn, Wr, br, slope=0.2, relu=True):
    """Reference call sites: ``algos/madrqn/agents/gnn_agents.py:92-97,103-104``.  A.1:
    out_v = relu(sum_e alpha_e * el_u + W_r x_v + b_r),  alpha = softmax_e(attn . leaky_relu(el_u + er_v))."""
    x_src, x_dst = np.asarray(x_src, np.float64), np.asarray(x_dst, np.float64)
    heads, D = attn.shape[-2], attn.shape[-1]
    attn = np.asarray(attn, np.float64).reshape(heads, D)
    out = np.zeros((n_dst, heads, D))
    alpha = np.zeros((len(src), heads))
    for v in range(n_dst):
        er = (Wd @ x_dst[v] + bd).reshape(heads, D)
        edges = [e for e in range(len(src)) if dst[e] == v]
        for k in range(heads):
            scores, els = [], []
            for e in edges:
                el = (Ws @ x_src[src[e]] + bs).reshape(heads, D)[k]
                z = el + er[k]
                y = np.where(z > 0, z, slope * z)
                scores.append(float(attn[k] @ y))
                els.append(el)
            if edges:
                m = max(scores)
                w = [np.exp(s - m) for s in scores]
                tot = sum(w)
                for e, wi, el in zip(edges, w, els):
                    alpha[e, k] = wi / tot
                    out[v, k] += wi / tot * el
        if Wr is not None:
            out[v] += (Wr @ x_dst[v] + (br if br is not None else 0.0)).reshape(heads, D)
    return (np.maximum(out, 0) if relu else out), alpha


def _sigmoid(a):
    return 1.0 / (1.0 + np.exp(-a))


def gru_cell_loops(x, h, w_ih, w_hh, b_ih, b_hh):
    """A.3, one row at a time."""
    H = h.shape[1]
    out = np.zeros_like(h, dtype=np.float64)
    for i in range(x.shape[0]):
        gi, gh = w_ih @ x[i] + b_ih, w_hh @ h[i] + b_hh
        r = _sigmoid(gi[:H] + gh[:H])
        z = _sigmoid(gi[H:2 * H] + gh[H:2 * H])
        n = np.tanh(gi[2 * H:] + r * gh[2 * H:])
        out[i] = (1 - z) * n + z * h[i]
    return out


def tarmac_loops(src, dst, x, h, Wv, bv, Wsg, bsg, Wq, bq, w_ih, w_hh, b_ih, b_hh, key_size, n_rounds=1):
    """A.2 / reference ``gnn_agents.py:248-271``, explicit loops."""
    x, h = np.asarray(x, np.float64), np.asarray(h, np.float64)
    n = x.shape[0]
    for _ in range(n_rounds):
        inp = np.concatenate([x, h], 1)
        v = inp @ Wv.T + bv
        s = inp @ Wsg.T + bsg
        q = inp @ Wq.T + bq
        c = np.zeros((n, v.shape[1]))
        for j in range(n):
            edges = [e for e in range(len(src)) if dst[e] == j]
            if not edges:
                continue
            sc = [float(s[src[e]] @ q[j]) / key_size for e in edges]
            m = max(sc)
            w = [np.exp(a - m) for a in sc]
            tot = sum(w)
            for e, wi in zip(edges, w):
                c[j] += wi / tot * v[src[e]]
        h = gru_cell_loops(np.concatenate([x, c], 1), h, w_ih, w_hh, b_ih, b_hh)
    return h
