"""Vectorised observation → graph builder.

The reference builds one DGL star graph per agent per env step, batches them, builds a comm graph with an
O(U²) Python loop and merges the two (``algos/madrqn/utils/env_wrappers.py:65-89,122-154``); the learner then
``dgl.batch``es B of those per timestep (``algos/common.py:40-47``).  The resulting layout is fully determined
by the visibility flags, so for B env instances at once it is just a masked compaction and a cumsum:

* node order of ``gt`` / ``ubs`` = (env, agent, slot) over visible rows  == edge order == CSR slot order
  (star layout: ``src id == edge id``, edges destination-sorted) — identical to what
  ``dgl.batch([dgl.merge([dgl.batch(per_agent_graphs), comm_graph]) for env in envs])`` produces;
* ``talk`` edge order = (env, src i, dst j) over ``adj[env, i, j]`` (reference loop order, src-major), its
  CSR-by-dst slots = (env, dst j, src i) with ``eid`` mapping slots back to edge ids.

Runs on whatever device the dense observations live on (torch ops only; no Python loop over envs/agents).
"""
from __future__ import annotations

from typing import Optional

import torch as th

from .graph import HeteroGraph, RelCSR

OBS_CETS = (("agent", "talk", "agent"), ("gt", "seen", "agent"), ("ubs", "near", "agent"))


def _star(flag: th.Tensor, feats: th.Tensor, n_dst: int):
    """flag (N, S) bool, feats (N, S, F) → packed features (E, F), CSR, per-dst degree."""
    deg = flag.sum(-1).flatten()
    indptr = th.zeros(n_dst + 1, dtype=th.int64, device=flag.device)
    th.cumsum(deg, 0, out=indptr[1:])
    x = feats[flag]
    return x, indptr.to(th.int32), deg


def build_obs_graph_batch(agent_obs: th.Tensor, gt_obs: th.Tensor, ubs_obs: th.Tensor,
                          comm_adj: Optional[th.Tensor] = None) -> HeteroGraph:
    """Dense observations of B envs × U agents → one batched HeteroGraph (batch size B).

    ``agent_obs (B,U,F_ag)``; ``gt_obs (B,U,G,1+F_gt)`` and ``ubs_obs (B,U,U-1,1+F_ubs)`` with column 0 the
    visibility flag (``envs/mubs_cov/mubs_cov.py:215-242``; the wrapper keeps rows with flag == 1 and strips the
    flag, ``env_wrappers.py:71,85-86``); ``comm_adj (B,U,U)`` bool with ``adj[b,i,j]`` ⇔ edge i→j
    (``d_u2u[i,j] <= r_comm``, ``env_wrappers.py:141-144``; self-loops included).  ``comm_adj=None`` builds the
    graph without talk edges (``c=None`` agents)."""
    B, U = agent_obs.shape[:2]
    N = B * U
    dev = agent_obs.device
    Gn, Un = gt_obs.shape[2], ubs_obs.shape[2]
    x_gt, ip_gt, _ = _star(gt_obs[..., 0].reshape(N, Gn) == 1, gt_obs[..., 1:].reshape(N, Gn, gt_obs.shape[3] - 1), N)
    x_ubs, ip_ubs, _ = _star(ubs_obs[..., 0].reshape(N, Un) == 1, ubs_obs[..., 1:].reshape(N, Un, ubs_obs.shape[3] - 1), N)
    E_gt, E_ubs = x_gt.shape[0], x_ubs.shape[0]
    src, dst, csr = {}, {}, {}
    c_seen, c_near, c_talk = OBS_CETS[1], OBS_CETS[2], OBS_CETS[0]
    src[c_seen] = dst[c_seen] = src[c_near] = dst[c_near] = None    # star layout: edge lists derive from the CSR
    csr[c_seen] = RelCSR(ip_gt, None, None, E_gt, N, E_gt)
    csr[c_near] = RelCSR(ip_ubs, None, None, E_ubs, N, E_ubs)
    if comm_adj is None:
        z = th.zeros(0, dtype=th.int64, device=dev)
        src[c_talk], dst[c_talk] = z, z
        csr[c_talk] = RelCSR(th.zeros(N + 1, dtype=th.int32, device=dev), th.zeros(0, dtype=th.int32, device=dev),
                             None, N, N, 0)
        bne_talk = [0] * B
    else:
        adj = comm_adj.to(th.bool)
        src[c_talk] = dst[c_talk] = None                            # (env, src, dst) order is rebuilt from eid
        adj_t = adj.transpose(1, 2)                                   # [b, dst, src]
        bt, jt, it = th.nonzero(adj_t, as_tuple=True)                 # CSR slot order (env, dst, src)
        indeg = adj_t.sum(-1).flatten()
        ip = th.zeros(N + 1, dtype=th.int64, device=dev)
        th.cumsum(indeg, 0, out=ip[1:])
        edge_id = (th.cumsum(adj.flatten().to(th.int64), 0) - 1).view(B, U, U)
        eid = edge_id.transpose(1, 2)[adj_t]
        mask = None
        if U <= 32:
            mask = (adj.to(th.int64) << th.arange(U, device=dev).view(1, U, 1)).sum(1).flatten().to(th.int32)
        E_t = int(bt.numel())
        csr[c_talk] = RelCSR(ip.to(th.int32), (bt * U + it).to(th.int32), eid, N, N, E_t,
                             U if U <= 32 else None, mask)
        bne_talk = adj.sum((1, 2)).tolist()
    per_env_gt = (ip_gt[U::U] - ip_gt[:-1:U]).tolist() if B else []
    per_env_ubs = (ip_ubs[U::U] - ip_ubs[:-1:U]).tolist() if B else []
    nframes = {"agent": {"feat": agent_obs.reshape(N, -1)}, "gt": {"feat": x_gt}, "ubs": {"feat": x_ubs}}
    return HeteroGraph(("agent", "gt", "ubs"), OBS_CETS, {"agent": N, "gt": E_gt, "ubs": E_ubs}, src, dst, nframes,
                       None, {"agent": [U] * B, "gt": per_env_gt, "ubs": per_env_ubs},
                       {c_talk: bne_talk, c_seen: per_env_gt, c_near: per_env_ubs}, csr)


def build_drqn_graph_batch(agent_obs: th.Tensor, gt_obs: th.Tensor) -> HeteroGraph:
    """Single-UBS layout of reference ``algos/drqn/utils/env_wrappers.py:63-77``: every GT is always a node,
    one relation ``('gt','seen-by','agent')``.  ``agent_obs (B,F_ag)``, ``gt_obs (B,G,F_gt)``."""
    B, G = gt_obs.shape[:2]
    dev = gt_obs.device
    cet = ("gt", "seen-by", "agent")
    ip = (th.arange(B + 1, device=dev) * G).to(th.int32)
    src = th.arange(B * G, device=dev)
    dst = th.repeat_interleave(th.arange(B, device=dev), G)
    return HeteroGraph(("agent", "gt"), (cet,), {"agent": B, "gt": B * G}, {cet: src}, {cet: dst},
                       {"agent": {"feat": agent_obs.reshape(B, -1)}, "gt": {"feat": gt_obs.reshape(B * G, -1)}}, None,
                       {"agent": [1] * B, "gt": [G] * B}, {cet: [G] * B},
                       {cet: RelCSR(ip, None, None, B * G, B, B * G)})
