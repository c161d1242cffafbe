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


def functional_inputs(seed, N, H, T, A):
    """Seeded ``h0 (N,H)``, ``w (T,N,A)``, ``hw (N,H)`` of the scalar functional
    ``(q*w).sum() + (q**2).mean() + (h_T*hw).sum()`` the agent-level parity cases differentiate."""
    gen = th.Generator().manual_seed(5 + seed)
    h0 = th.randn(N, H, generator=gen) * 0.3
    w = th.randn(T, N, A, generator=gen)
    hw = th.randn(N, H, generator=gen)
    return h0, w, hw


def unpack64(r32, packed):
    """Inverse of ``make_ref_golden.pack64``: the float64 reference result from the float32 one + compact correction."""
    return r32.double() + packed["scale"] * packed["diff"].double()


def param_sums(sd):
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()}


def assert_param_sums(sd, sums, what="", rtol=1e-12):
    """A state_dict re-created from its seed must be the one the fixture was generated with."""
    assert set(sd.keys()) == set(sums.keys()), f"{what}: key sets differ: {sorted(set(sd) ^ set(sums))}"
    for k, (s, a) in param_sums(sd).items():
        rs, ra = sums[k]
        assert abs(s - rs) <= rtol * max(ra, 1e-30) and abs(a - ra) <= rtol * max(ra, 1e-30), \
            f"{what}: parameter {k} differs from the fixture's (sum {s} vs {rs})"


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


def assert_grads_close(grads, ref, rtol=1e-5, atol_scale=1e-6, floor_scale=2e-7, what=""):
    """Per-parameter ``assert_close`` with an absolute floor of ``floor_scale * max|grad over ALL parameters|``: some
    gradients are mathematically zero (``f_sign.bias``: softmax is shift invariant) and hold only cancellation noise."""
    assert set(grads) == set(ref), f"{what}: parameter sets differ: {sorted(set(grads) ^ set(ref))}"
    gmax = max(float(v.abs().max()) for v in ref.values())
    for k, g in grads.items():
        a, b = g.detach().float().cpu(), ref[k].detach().float().cpu()
        assert a.shape == b.shape, f"{what} {k}: shape"
        atol = atol_scale * float(b.abs().max()) + floor_scale * gmax
        bad = (a - b).abs() > atol + rtol * b.abs()
        assert not bool(bad.any()), (f"{what} grad {k}: {int(bad.sum())}/{a.numel()} outside tol; max abs err "
                                     f"{float((a - b).abs().max()):.3e}, max ref {float(b.abs().max()):.3e}, all-param max {gmax:.3e}")


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


# ---------------------------------------------------------------------------------------------------------------------
# Fixtures generated by the reference's own source (tests/golden/make_ref_golden.py -> tests/golden/ref_golden_v1.pt)
_REF_GOLDEN = None


def ref_golden():
    global _REF_GOLDEN
    if _REF_GOLDEN is None:
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden_v1.pt")
        _REF_GOLDEN = th.load(path, weights_only=False)
    return _REF_GOLDEN


def ref_agent_obs(case, t):
    """Step-t observations of an ``agent/<tag>`` fixture, regenerated from the stored seed (checksum verified):
    ``(agent, gt, ubs, adj, flat|None)``."""
    from uav_bs_ctrl_b200.synth import synth_dense_obs
    a, gt, ubs, adj = synth_dense_obs(case["B"], case["U"], case["G"], case["profile"], seed=case["obs_seed"] + t,
                                      comm_p=case["comm_p"])
    fl = None
    if case["flat"] is not None:
        gen = th.Generator().manual_seed(case["obs_seed"] + t + 50)
        fl = th.rand(case["B"], case["U"], case["flat"], generator=gen)
    got = float(sum(x.double().sum() for x in (a, gt, ubs, adj, fl) if x is not None))
    assert abs(got - case["obs_checksum"][t]) <= 1e-9 * abs(case["obs_checksum"][t]), "synthetic observations differ from the fixture's"
    return a, gt, ubs, adj, fl


def ref_agent_graph(case, t, device="cpu"):
    """The batched observation graph of step t built by OUR vectorised builder (the fixture side was built by the
    reference's per-agent wrapper + ``dgl.batch`` / ``dgl.merge``)."""
    from uav_bs_ctrl_b200.builder import build_obs_graph_batch
    a, gt, ubs, adj, fl = ref_agent_obs(case, t)
    if fl is not None:
        a, gt, ubs = fl, th.zeros(case["B"], case["U"], 0, 5), th.zeros(case["B"], case["U"], 0, 3)
    g = build_obs_graph_batch(a, gt, ubs, adj if case["args"]["c"] is not None else None)
    return g.to(device) if device != "cpu" else g


def ref_learner_batch(case, idx):
    """The mini-batch reference ``MultiAgentQLearner.update`` assembles (``learner.py:99-116``) from the sampled
    sequence indices ``idx``: per timestep the stacked dense observations of the sampled sequences (sample-major, like
    ``common.cat``), ``h0`` / ``h1`` (``batch['h'][0]``, ``batch['h'][1]``), ``acts (T, n·U, 1)``, ``rews (T, n, U|1)``,
    ``dones (T, n, 1)``, ``states (T+1, n, S)``."""
    T, tr, obs = case["max_seq_len"], case["transitions"], case["obs"]
    share = case["config"].get("share_reward", False)
    steps = []
    for t in range(T + 1):
        ids = [tr[k * T + t]["obs"] if t < T else tr[k * T + T - 1]["next_obs"] for k in idx]
        steps.append({key: th.stack([obs[i][key] for i in ids]) for key in obs[ids[0]].keys()})
    hs = lambda t: th.cat([tr[k * T + t]["h"] if t < T else (1 - float(tr[k * T + T - 1]["done"])) * tr[k * T + T - 1]["next_h"]
                           for k in idx])
    acts = th.stack([th.cat([tr[k * T + t]["act"].reshape(-1, 1) for k in idx]) for t in range(T)]).long()

    def rew(x):
        r = x["rew"].float().reshape(1, -1)
        return r.mean().reshape(1, 1) if share else r
    rews = th.stack([th.cat([rew(tr[k * T + t]) for k in idx]) for t in range(T)])
    dones = th.stack([th.tensor([[(1 - float(tr[k * T + t]["bad"])) * float(tr[k * T + t]["done"])] for k in idx])
                      for t in range(T)])
    states = None
    if "state" in tr[0]:
        states = th.stack([th.cat([tr[k * T + t]["state"] if t < T else tr[k * T + T - 1]["next_state"] for k in idx])
                           for t in range(T + 1)])
    return dict(steps=steps, h0=hs(0), h1=hs(1), acts=acts, rews=rews, dones=dones, states=states)
