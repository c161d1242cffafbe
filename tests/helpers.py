"""Shared test helpers: reference-style (per-agent) graph construction and argument namespaces."""
from types import SimpleNamespace

import numpy as np
import torch as th

from uav_bs_ctrl_b200 import graph as G


def make_args(**kw):
    """Model-shape fields of reference algos/madrqn/config.py:7-18,39."""
    d = dict(hidden_size=64, n_layers=1, n_heads=4, msg_size=64, key_size=16, n_rounds=1, c="tarmac", o="gnn",
             dueling=False)
    d.update(kw)
    return SimpleNamespace(**d)


def build_obs_graph_ref(obs):
    """Line-by-line use of the container the way reference env_wrappers.py:69-89 uses dgl."""
    gt_ids, ubs_ids = np.equal(obs['gt'][:, 0], 1), np.equal(obs['ubs'][:, 0], 1)
    num_gts, num_ubs = gt_ids.sum(), ubs_ids.sum()
    data_dict = {
        ('gt', 'seen', 'agent'): (th.arange(num_gts), th.zeros(num_gts, dtype=th.long)),
        ('ubs', 'near', 'agent'): (th.arange(num_ubs), th.zeros(num_ubs, dtype=th.long)),
        ('agent', 'talk', 'agent'): ([], []),
    }
    g = G.heterograph(data_dict, num_nodes_dict={'gt': num_gts, 'ubs': num_ubs, 'agent': 1})
    g.ndata['feat'] = {
        'gt': th.as_tensor(obs['gt'][gt_ids, 1:]),
        'ubs': th.as_tensor(obs['ubs'][ubs_ids, 1:]),
        'agent': th.as_tensor(obs['agent']).unsqueeze(0),
    }
    return g


def build_comm_graph_ref(adj):
    """reference env_wrappers.py:139-154 with adj[i, j] = (d_u2u[i, j] <= r_comm)."""
    n = adj.shape[0]
    u, v = [], []
    for i in range(n):
        for j in range(n):
            if adj[i, j]:
                u.append(i), v.append(j)
    return G.heterograph({('gt', 'seen', 'agent'): ([], []), ('ubs', 'near', 'agent'): ([], []),
                          ('agent', 'talk', 'agent'): (u, v)}, num_nodes_dict={'gt': 0, 'ubs': 0, 'agent': n})


def env_graph_ref(agent_obs, gt_obs, ubs_obs, adj):
    """One env: U per-agent star graphs → batch → merge with comm graph (env_wrappers.py:122-137)."""
    U = agent_obs.shape[0]
    per_agent = [build_obs_graph_ref(dict(agent=agent_obs[i].numpy(), gt=gt_obs[i].numpy(), ubs=ubs_obs[i].numpy()))
                 for i in range(U)]
    local = G.batch(per_agent)
    if adj is None:
        return local
    return G.merge([local, build_comm_graph_ref(adj.numpy())])


def batched_graph_ref(agent_obs, gt_obs, ubs_obs, adj):
    """B envs → algos/common.cat (dgl.batch)."""
    return G.batch([env_graph_ref(agent_obs[b], gt_obs[b], ubs_obs[b], None if adj is None else adj[b])
                    for b in range(agent_obs.shape[0])])


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, rtol=1e-5, atol_scale=1e-6, what=""):
    """SURVEY §8(c) tolerance: allclose(rtol=1e-5, atol=1e-6·max|ref|)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    atol = atol_scale * float(b.abs().max()) if b.numel() else 0.0
    bad = (a - b).abs() > atol + rtol * b.abs()
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())}/{a.numel()} outside tol; max abs err "
                                 f"{float((a - b).abs().max()):.3e}, max ref {float(b.abs().max()):.3e}")


def assert_as_accurate(a, ref32, ref64, what="", slack=4.0, floor_scale=2e-6, abs_floor=1e-9):
    """Reduction-heavy results (parameter gradients sum thousands of cancelling terms): the kernel must be as
    accurate against the fp64 oracle as the fp32 oracle itself is (x slack), with a floor of 2e-6 * max|ref|."""
    a = a.detach().double().cpu().reshape(-1)
    r32, r64 = ref32.detach().double().cpu().reshape(-1), ref64.detach().double().cpu().reshape(-1)
    assert a.shape == r64.shape, f"{what}: shape mismatch"
    if a.numel() == 0:
        return
    scale = float(r64.abs().max())
    err_k, err_o = float((a - r64).abs().max()), float((r32 - r64).abs().max())
    bound = max(slack * err_o, floor_scale * scale, abs_floor)
    assert err_k <= bound, f"{what}: kernel err {err_k:.3e} vs fp64 > bound {bound:.3e} (fp32 oracle err {err_o:.3e}, scale {scale:.3e})"
