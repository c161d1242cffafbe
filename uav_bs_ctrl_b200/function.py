"""Message-passing API for the comm protocols that are not (yet) fused into CUDA kernels.

Mirrors the slice of ``dgl.function`` / ``dgl.nn.functional.edge_softmax`` / UDF ``update_all`` that the
reference's agents use (``algos/madrqn/agents/gnn_agents.py:125-144,168-189,207-228,261-266,287-297``).
Implemented with differentiable torch ops on whatever device the graph lives on; semantics follow
SURVEY.md Appendix A.4: builtin ``sum`` over no edges gives zeros, UDF reduce is degree-bucketed with a
``mailbox`` of shape ``(n, deg, F)`` in edge-id order, nodes that receive no message get zeros.

The fused TarMAC / GATv2 modules do NOT route through this file.
"""
from __future__ import annotations

import torch as th

__all__ = ["u_add_v", "u_dot_v", "u_mul_e", "copy_u", "sum", "mean", "max", "edge_softmax",
           "apply_edges", "update_all"]


class _Msg:
    def __init__(self, kind, a, b, out):
        self.kind, self.a, self.b, self.out = kind, a, b, out


class _Red:
    def __init__(self, kind, msg, out):
        self.kind, self.msg, self.out = kind, msg, out


def u_add_v(u, v, out):
    return _Msg("u_add_v", u, v, out)


def u_dot_v(u, v, out):
    return _Msg("u_dot_v", u, v, out)


def u_mul_e(u, e, out):
    return _Msg("u_mul_e", u, e, out)


def copy_u(u, out):
    return _Msg("copy_u", u, None, out)


def sum(msg, out):  # noqa: A001 - mirrors dgl.function.sum
    return _Red("sum", msg, out)


def mean(msg, out):
    return _Red("mean", msg, out)


def max(msg, out):  # noqa: A001
    return _Red("max", msg, out)


class _EdgeBatch:
    def __init__(self, rel, u, v):
        self.src = {k: t[u] for k, t in rel.srcdata.items()}
        self.dst = {k: t[v] for k, t in rel.dstdata.items()}
        self.data = rel.edata

    def __len__(self):
        return next(iter(self.src.values())).shape[0]


class _NodeBatch:
    def __init__(self, data, mailbox):
        self.data, self.mailbox = data, mailbox


def _bcast_edge(e, like):
    while e.dim() < like.dim():
        e = e.unsqueeze(-1)
    return e


def _eval_msg(rel, m: _Msg):
    u, v = rel.edges()
    if m.kind == "u_add_v":
        return rel.srcdata[m.a][u] + rel.dstdata[m.b][v]
    if m.kind == "u_dot_v":
        return (rel.srcdata[m.a][u] * rel.dstdata[m.b][v]).sum(-1, keepdim=True)
    if m.kind == "u_mul_e":
        x = rel.srcdata[m.a][u]
        return x * _bcast_edge(rel.edata[m.b], x) if rel.edata[m.b].dim() <= x.dim() else x * rel.edata[m.b]
    if m.kind == "copy_u":
        return rel.srcdata[m.a][u]
    raise KeyError(m.kind)


def apply_edges(rel, func):
    if isinstance(func, _Msg):
        rel.edata[func.out] = _eval_msg(rel, func)
        return
    u, v = rel.edges()
    rel.edata.update(func(_EdgeBatch(rel, u, v)))


def _segment_reduce(kind, msg, dst, n_dst):
    shape = (n_dst,) + tuple(msg.shape[1:])
    if kind in ("sum", "mean"):
        out = th.zeros(shape, dtype=msg.dtype, device=msg.device).index_add_(0, dst, msg)
        if kind == "mean":
            deg = th.bincount(dst, minlength=n_dst).clamp_(min=1).to(msg.dtype)
            out = out / deg.view((-1,) + (1,) * (msg.dim() - 1))
        return out
    if kind == "max":
        idx = dst.view((-1,) + (1,) * (msg.dim() - 1)).expand_as(msg)
        out = th.zeros(shape, dtype=msg.dtype, device=msg.device)
        return out.scatter_reduce(0, idx, msg, reduce="amax", include_self=False)
    raise KeyError(kind)


def update_all(rel, message_func, reduce_func):
    u, v = rel.edges()
    n_dst = rel.num_dst_nodes()
    if isinstance(message_func, _Msg):
        msgs = {message_func.out: _eval_msg(rel, message_func)}
    else:
        msgs = message_func(_EdgeBatch(rel, u, v))
    if isinstance(reduce_func, _Red):
        rel.dstdata[reduce_func.out] = _segment_reduce(reduce_func.kind, msgs[reduce_func.msg], v, n_dst)
        return
    # UDF reduce: bucket destinations by in-degree; mailbox rows are in edge-id order.
    deg = th.bincount(v, minlength=n_dst)
    order = th.sort(v, stable=True)[1]
    start = th.cumsum(deg, 0) - deg
    results = {}
    for d in th.unique(deg).tolist():
        if d == 0:
            continue
        nodes = th.nonzero(deg == d).flatten()
        slots = (start[nodes].unsqueeze(1) + th.arange(d, device=v.device)).flatten()
        eids = order[slots]
        mailbox = {k: m[eids].view((nodes.numel(), d) + tuple(m.shape[1:])) for k, m in msgs.items()}
        data = {k: t[nodes] for k, t in rel.dstdata.items()}
        for k, val in reduce_func(_NodeBatch(data, mailbox)).items():
            if k not in results:
                results[k] = th.zeros((n_dst,) + tuple(val.shape[1:]), dtype=val.dtype, device=val.device)
            results[k] = results[k].index_copy(0, nodes, val)
    rel.dstdata.update(results)


def edge_softmax(rel, e: th.Tensor) -> th.Tensor:
    """Softmax of edge scores over the in-edges of every destination (``norm_by='dst'``)."""
    _, v = rel.edges()
    n_dst = rel.num_dst_nodes()
    idx = v.view((-1,) + (1,) * (e.dim() - 1)).expand_as(e)
    mx = th.full((n_dst,) + tuple(e.shape[1:]), float("-inf"), dtype=e.dtype, device=e.device)
    mx = mx.scatter_reduce(0, idx, e.detach(), reduce="amax", include_self=True)
    ex = th.exp(e - mx[v])
    den = th.zeros_like(mx).index_add_(0, v, ex)
    return ex / den[v]
