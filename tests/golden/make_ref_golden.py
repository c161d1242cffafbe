#!/usr/bin/env python
"""Golden vectors of the hot path produced by the reference's OWN source code (build container only):

    python tests/golden/make_ref_golden.py            # writes tests/golden/ref_golden_v1.pt

What executes is ``/root/reference`` byte for byte — ``algos/madrqn/agents/gnn_agents.py`` (``GnnAgent``,
``GraphObservationEncoder``, ``DenseObservationEncoder``, ``TarMAC``, ``BaseComm``, ``CommNet``, ``DiscreteComm``,
``EdgeConv``), ``agents/dueling.py``, ``agents/mixers.py``, ``algos/madrqn/learner.py`` (``MultiAgentQLearner.act /
cache / update``), ``buffer.py``, ``algos/common.py`` (``cat``), ``algos/madrqn/utils/env_wrappers.py`` (the graphs
are built by the reference's wrapper), ``algos/drqn/*`` (``QLearner``, DRQN ``GnnAgent``, wrapper), ``envs/mubs_cov``
and ``envs/subs_cov`` — with the uninstalled ``dgl`` / ``gym`` / ``matplotlib`` replaced by the stand-ins of
``tests/refshim.py`` (``dgl`` -> CPU graph container + torch message-passing ops, ``dglnn.GATv2Conv`` -> the restated
DGL 0.9.0 module: the one class on this path whose source is not under /root/reference).

Two kinds of fixtures:

* ``agent/<tag>``   the reference agent classes as the oracle: seeded synthetic observations (regenerated from the seed
                    on the checking side; a float64 checksum is stored), random ``h0``, T unrolled steps, a scalar
                    functional of (Q, h_T), its gradient w.r.t. every parameter — in float32 and (same classes,
                    ``.double()``) float64;
* ``learner/<tag>`` the reference learner driven by the reference's own loop (``algos/madrqn/run.py:81-99``,
                    ``algos/drqn/run.py:76-94``) on the reference's own env: every transition handed to ``cache``,
                    the Q values / hidden states ``act`` computed, the sequence indices ``update`` sampled, and after
                    each ``update``: LossQ, QVals, the clipped gradients, the policy / target / mixer parameters.

Gumbel noise of ``DiscreteComm`` (``gnn_agents.py:173``): ``F.gumbel_softmax`` draws ``exponential_()`` from the global
torch RNG; ``GumbelTap`` replaces exactly that draw (nothing else), fills it from a seeded float32 stream
and logs it, so the float64 run and the CUDA side see the same noise.
"""
import copy
import os
import random
import sys
from types import SimpleNamespace as SN

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import refshim  # noqa: E402
from helpers import functional_inputs  # noqa: E402
from uav_bs_ctrl_b200.synth import synth_dense_obs  # noqa: E402

SHAPE = {'agent': 2, 'ubs': 2, 'gt': 4}


def model_args(**kw):
    """Model-shape fields of reference ``algos/madrqn/config.py:7-19,39`` (defaults are the reference's)."""
    d = dict(hidden_size=64, n_layers=1, n_heads=4, msg_size=64, key_size=16, n_rounds=1, embed_dim=32, c="tarmac",
             o="gnn", dueling=False)
    d.update(kw)
    return d


class GumbelTap:
    """Context manager that replaces ``Tensor.exponential_`` (the only RNG draw of ``F.gumbel_softmax``) by a seeded
    float32 stream (or a replay of a logged one) and logs every draw.  Neither the reference nor
    ``torch.nn.functional.gumbel_softmax`` is touched: only where the random numbers come from."""

    def __init__(self, seed=None, replay=None):
        self.gen = th.Generator().manual_seed(seed) if seed is not None else None
        self.replay, self.log, self.i = replay, [], 0

    def __enter__(self):
        tap = self

        def exponential_(t, *a, **k):
            if tap.replay is not None:
                e = tap.replay[tap.i]
            else:
                e = th.empty(t.shape, dtype=th.float32)
                th._C.TensorBase.exponential_(e, generator=tap.gen)
            tap.i += 1
            tap.log.append(e)
            with th.no_grad():
                t.copy_(e)
            return t

        th.Tensor.exponential_ = exponential_          # shadows TensorBase.exponential_ on the Python subclass
        return self

    def __exit__(self, *exc):
        del th.Tensor.exponential_
        return False


# ---------------------------------------------------------------------------------------------------------------------
def ref_graph(M, agent_obs, gt_obs, ubs_obs, adj, comm, flat=None):
    """Observation object of ONE env built by the reference's wrapper code (``env_wrappers.py:65-89,122-154``)."""
    W = M.wrappers
    U = agent_obs.shape[0]
    if flat is not None:
        local = th.as_tensor(flat, dtype=th.float32)                          # FlattenedObservation.local_observation
    else:
        obs = [dict(agent=agent_obs[i].numpy(), gt=gt_obs[i].numpy(), ubs=ubs_obs[i].numpy()) for i in range(U)]
        local = W.GraphObservation(None).local_observation(obs)
    if comm is None:
        return local
    d = np.where(adj.numpy(), 0.0, 2.0)
    fake = SN(n_agents=U, d_u2u=d, r_comm=1.0)
    comm_graph = W.MultiUbsCoverageWrapper.build_comm_graph(fake)
    if flat is not None:
        comm_graph.nodes['agent'].data['feat'] = local
        return comm_graph
    import dgl
    return dgl.merge([local, comm_graph])


def g64(g):
    return g._map(lambda t: t.double() if t.is_floating_point() else t, lambda r: r)


def checksum(obs):
    return float(sum(t.double().sum() for t in obs))


def param_sums(sd):
    """Per-parameter (sum, abs-sum) in float64: pins a state_dict that the checking side re-creates from its seed."""
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()}


def pack64(r32, r64):
    """float64 result stored as a compact correction of the float32 one: ``r64 = r32 + scale * diff16`` (the correction
    is ~1e-7 of the values, so 3 significant digits of it reproduce r64 to ~1e-10 relative)."""
    d = r64.double() - r32.double()
    scale = float(d.abs().max())
    return dict(scale=scale, diff=(d / scale).to(th.float16) if scale > 0 else th.zeros_like(d, dtype=th.float16))


def synth_step(B, U, G, profile, seed, comm_p, flat):
    a, gt, ubs, adj = synth_dense_obs(B, U, G, profile, seed=seed, comm_p=comm_p)
    fl = None
    if flat is not None:
        gen = th.Generator().manual_seed(seed + 50)
        fl = th.rand(B, U, flat, generator=gen)
    return a, gt, ubs, adj, fl


def agent_case(M, tag, c="tarmac", H=64, U=8, G=80, B=6, T=4, profile="full", comm_p=1.0, flat=None, seed=1,
               obs_seed=300, n_actions=9, keep_sd=False, quiet=False, **extra):
    args = model_args(c=c, hidden_size=H, o="mlp" if flat else "gnn", n_layers=2 if flat else 1, **extra)
    th.manual_seed(seed)
    agent = M.agents.GnnAgent(flat if flat else SHAPE, n_actions, SN(**args))
    sd = {k: v.clone() for k, v in agent.state_dict().items()}
    N = B * U
    steps = [synth_step(B, U, G, profile, obs_seed + t, comm_p, flat) for t in range(T)]
    graphs = []
    for a, gt, ubs, adj, fl in steps:
        per_env = [ref_graph(M, a[b], gt[b], ubs[b], adj[b], c, None if fl is None else fl[b]) for b in range(B)]
        graphs.append(M.common.cat(per_env))
    h0, w, hw = functional_inputs(seed, N, H, T, n_actions)       # regenerated from the seed on the checking side
    out = dict(args=args, obs_shape=flat if flat else dict(SHAPE), n_actions=n_actions, B=B, U=U, G=G, T=T,
               profile=profile, comm_p=comm_p, flat=flat, obs_seed=obs_seed,
               obs_checksum=[checksum([t for t in s if t is not None]) for s in steps], seed=seed,
               param_sums=param_sums(sd), fn_checksum=checksum([h0, w, hw]))
    if keep_sd:
        out["state_dict"] = sd
    noise = None
    for name, dt in (("r32", th.float32), ("r64", th.float64)):
        net = copy.deepcopy(agent).to(dt)
        tap = GumbelTap(seed=77, replay=noise)
        h, qs = h0.to(dt), []
        with tap:
            for t in range(T):
                q, h = net(graphs[t] if dt == th.float32 else g64(graphs[t]), h)
                qs.append(q)
        if c == "disc" and noise is None:
            noise = tap.log
        qs = th.stack(qs)
        ((qs * w.to(dt)).sum() + (qs ** 2).mean() + (h * hw.to(dt)).sum()).backward()
        out[name] = dict(q=qs.detach(), h=h.detach(), grads={k: p.grad.clone() for k, p in net.named_parameters()})
    r32, r64 = out["r32"], out.pop("r64")
    out["err64"] = float((r32["q"].double() - r64["q"]).abs().max())
    out["r64"] = dict(q=pack64(r32["q"], r64["q"]), h=pack64(r32["h"], r64["h"]),
                      grads={k: pack64(r32["grads"][k], r64["grads"][k]) for k in r64["grads"]})
    if noise is not None:
        out["exponential"] = noise                      # gumbel = -log(exponential); (E, msg, 2) per step, edge-id order
    if not quiet:
        print(f"agent/{tag}: N={N} T={T} |q|max={float(out['r32']['q'].abs().max()):.3f} fp32-vs-fp64 {out['err64']:.2e}")
    return out


AGENT_CASES = [
    # tag,                 kwargs
    ("exp3", dict(c="tarmac", H=64, U=8, G=80, B=6, profile="full", comm_p=1.0)),
    ("exp3-thin", dict(c="tarmac", H=64, U=8, G=20, B=5, profile="realistic", comm_p=0.5)),
    ("u3", dict(c="tarmac", H=64, U=3, G=10, B=7, profile="realistic", comm_p=0.6)),
    ("u16", dict(c="tarmac", H=64, U=16, G=12, B=3, profile="realistic", comm_p=0.7)),
    ("h32", dict(c="tarmac", H=32, U=4, G=10, B=9, profile="random", comm_p=1.0)),
    ("h128", dict(c="tarmac", H=128, U=8, G=10, B=4, profile="realistic", comm_p=0.8)),
    ("indep", dict(c=None, H=64, U=8, G=30, B=5, profile="realistic", comm_p=1.0)),
    ("exp2-mlp", dict(c="tarmac", H=64, U=8, G=0, B=6, profile="full", comm_p=0.7, flat=423)),
    ("tc-rows", dict(c="tarmac", H=64, U=8, G=12, B=16, profile="realistic", comm_p=0.8)),
    ("rounds2", dict(c="tarmac", H=64, U=8, G=20, B=5, profile="realistic", comm_p=0.6, n_rounds=2)),
    ("dueling", dict(c="tarmac", H=64, U=8, G=20, B=4, profile="realistic", comm_p=0.8, dueling=True)),
    ("base", dict(c="base", H=64, U=8, G=20, B=5, profile="realistic", comm_p=0.6)),
    ("commnet", dict(c="commnet", H=64, U=8, G=20, B=5, profile="realistic", comm_p=0.6)),
    ("commnet-r2", dict(c="commnet", H=64, U=8, G=20, B=5, profile="realistic", comm_p=0.6, n_rounds=2)),
    ("econv", dict(c="econv", H=64, U=8, G=20, B=5, profile="realistic", comm_p=0.6)),
    ("econv-r2", dict(c="econv", H=64, U=6, G=10, B=4, profile="realistic", comm_p=0.5, n_rounds=2)),
    ("disc", dict(c="disc", H=64, U=8, G=20, B=5, profile="realistic", comm_p=0.6)),
    ("disc-self", dict(c="disc", H=64, U=8, G=20, B=5, profile="realistic", comm_p=0.0)),   # in-degree 1: no max ties
    # BASELINE.json configs at (or near) full size
    ("exp3-full", dict(c="tarmac", H=64, U=8, G=80, B=256, T=2, profile="full", comm_p=1.0, obs_seed=900)),
    ("exp3-real", dict(c="tarmac", H=64, U=8, G=80, B=256, T=2, profile="realistic", comm_p=0.7, obs_seed=910)),
    ("scaled-full", dict(c="tarmac", H=128, U=16, G=320, B=4, T=2, profile="full", comm_p=1.0, obs_seed=920)),
    ("scaled-real", dict(c="tarmac", H=128, U=16, G=320, B=8, T=3, profile="realistic", comm_p=0.6, obs_seed=930)),
    ("exp2-full", dict(c="tarmac", H=64, U=8, G=0, B=256, T=2, profile="full", comm_p=0.8, flat=423, obs_seed=940)),
]


# shape-coverage cases that are only run LIVE (CPU test-suite in the build container: reference vs oracle restatement);
# the committed fixture keeps the configs BASELINE.json names and one case per module of the reference
LIVE_ONLY = {"u3", "u16", "h32", "h128", "tc-rows", "commnet", "scaled-full"}


# ---------------------------------------------------------------------------------------------------------------------
def dense_of(obs_list):
    """The env's per-agent observation dicts -> stacked arrays (what our builder consumes)."""
    return dict(agent=th.as_tensor(np.stack([o['agent'] for o in obs_list]), dtype=th.float32),
                gt=th.as_tensor(np.stack([o['gt'] for o in obs_list]), dtype=th.float32),
                ubs=th.as_tensor(np.stack([o['ubs'] for o in obs_list]), dtype=th.float32))


def _full_update(learner, out, first):
    upd = dict(LossQ=out["LossQ"], QVals=th.as_tensor(out["QVals"]))
    pol = {n: v.clone() for n, v in learner.policy_net.state_dict().items()}
    upd["policy_sums"] = param_sums(pol)
    if first:                                            # later updates: loss / Q values / parameter checksums only
        upd["grads"] = {n: p.grad.clone() for n, p in learner.policy_net.named_parameters()}
        upd["policy"] = pol
        upd["target_sums"] = param_sums(learner.target_net.state_dict())
        if getattr(learner, "mixer", None) is not None:
            upd["mixer_grads"] = {n: p.grad.clone() for n, p in learner.mixer.named_parameters()}
            upd["mixer_sums"] = param_sums(learner.mixer.state_dict())
            upd["target_mixer_sums"] = param_sums(learner.target_mixer.state_dict())
    return upd


def learner_case(M, tag, map_id, n_steps, seed, hover_p=0.0, n_updates=2, keep_sd=False, **cfg):
    """Reference loop ``algos/madrqn/run.py:81-99`` on the reference env, then ``n_updates`` reference updates."""
    Env, MAPS = refshim.mubs_env()
    config = copy.deepcopy(M.config.DEFAULT_CONFIG)
    config.update(cfg)
    args = M.common.check_args_sanity(SN(**config))
    M.common.set_rand_seed(seed)
    raw_env = Env(map_id=map_id, record=False)
    env = M.wrappers.MultiUbsCoverageWrapper(raw_env, args)
    env_info = env.get_env_info()
    learner = M.learner.MultiAgentQLearner(env_info, args)
    sd0 = {k: v.clone() for k, v in learner.policy_net.state_dict().items()}
    mixer0 = None if learner.mixer is None else {k: v.clone() for k, v in learner.mixer.state_dict().items()}
    logits_log = []
    hook = learner.policy_net.register_forward_hook(lambda m, i, o: logits_log.append(o[0].detach().clone()))

    def pack_obs():
        d = dense_of(raw_env.get_obs())
        d["adj"] = th.as_tensor(raw_env.d_u2u <= raw_env.r_comm)
        if args.o == "mlp":
            d["flat"] = env.local_obs_wrapper.local_observation(raw_env.get_obs())
        return d

    obs_store, trans = [], []
    (o, s), h = env.reset(), learner.init_hidden()
    obs_store.append(pack_obs())
    cur = 0
    for t in range(n_steps):
        eps = 0.3
        a, h2 = learner.act(o, h, eps)
        o2, s2, r, d, info = env.step(a)
        obs_store.append(pack_obs())
        nxt = len(obs_store) - 1
        learner.cache(o, h, s, a, r, o2, h2, s2, d, info.get("BadMask"))
        trans.append(dict(obs=cur, next_obs=nxt, h=h.clone(), next_h=h2.clone(), state=s.clone(), next_state=s2.clone(),
                          act=th.tensor(a), rew=th.as_tensor(np.asarray(r, dtype=np.float64)), done=bool(d),
                          bad=bool(info.get("BadMask")), logits=logits_log[-1], eps=eps))
        o, s, h, cur = o2, s2, h2, nxt
        if d:
            (o, s), h = env.reset(), learner.init_hidden()
            obs_store.append(pack_obs())
            cur = len(obs_store) - 1
    hook.remove()
    n_seq = len(learner.buffer)
    updates, sample_idx = [], []
    for k in range(n_updates):
        random.seed(1000 + seed + k)
        idx = random.sample(range(n_seq), learner.batch_size)
        random.seed(1000 + seed + k)                     # update() draws random.sample(self.memory, batch_size)
        out = learner.update()
        sample_idx.append(idx)
        upd = _full_update(learner, out, k == 0)
        updates.append(upd)
    # check that the index replay is what update() really sampled: the sampled sequences' first actions
    keep = {k: config[k] for k in ("o", "c", "share_reward", "hidden_size", "n_layers", "n_heads", "msg_size", "key_size",
                                   "n_rounds", "embed_dim", "lr", "gamma", "polyak", "batch_size", "replay_size",
                                   "max_seq_len", "double_q", "dueling", "mixer", "anneal_lr")}
    print(f"learner/{tag}: {len(trans)} transitions, {n_seq} sequences of {learner.max_seq_len}, "
          f"LossQ {[round(u['LossQ'], 5) for u in updates]}, dones at "
          f"{[i for i, x in enumerate(trans) if x['done']]}, mean seen degree "
          f"{float(np.mean([(o_['gt'][..., 0] == 1).sum(-1).float().mean() for o_ in obs_store])):.2f}")
    res = dict(config=keep, env_info=env_info, map_id=map_id, n_gts=raw_env.n_gts, obs=obs_store, transitions=trans,
               seed=seed, policy0_sums=param_sums(sd0), mixer0=mixer0, sample_idx=sample_idx, updates=updates,
               max_seq_len=learner.max_seq_len)
    if keep_sd:
        res["policy0"] = sd0
    return res


def drqn_case(D, tag, n_steps, seed, n_updates=2, **cfg):
    """Reference loop ``algos/drqn/run.py:76-94`` with the reference single-UBS env + DRQN learner (exp1)."""
    refshim.install()
    from envs.subs_cov.subs_cov import SingleUbsCoverageEnv
    config = copy.deepcopy(D.config.DEFAULT_CONFIG)
    config.update(cfg)
    args = SN(**config)
    random.seed(seed), np.random.seed(seed), th.manual_seed(seed)
    raw_env = SingleUbsCoverageEnv(n_grps=2, gts_per_grp=5, episode_limit=12, record=False)     # 1 UBS x 10 GT (exp1)
    env = D.wrappers.Wrapper(raw_env, args)
    env_info = env.get_env_info()
    learner = D.learner.QLearner(env_info, args)
    sd0 = {k: v.clone() for k, v in learner.policy_net.state_dict().items()}
    logits_log = []
    hook = learner.policy_net.register_forward_hook(lambda m, i, o: logits_log.append(o[0].detach().clone()))
    feats = lambda g: dict(agent=g.nodes['agent'].data['feat'].clone().float(), gt=g.nodes['gt'].data['feat'].clone().float())
    obs_store, trans = [], []
    o, h = env.reset(), learner.init_hidden()
    obs_store.append(feats(o))
    cur = 0
    for t in range(n_steps):
        a, h2 = learner.act(o, h, 0.3)
        o2, r, d, info = env.step(a)
        obs_store.append(feats(o2))
        nxt = len(obs_store) - 1
        learner.cache(o, h, a, r, o2, h2, d, info.get('BadMask'))
        trans.append(dict(obs=cur, next_obs=nxt, h=h.clone(), next_h=h2.clone(), act=int(a), rew=float(r), done=bool(d),
                          bad=bool(info.get('BadMask')), logits=logits_log[-1]))
        o, h, cur = o2, h2, nxt
        if d:
            o, h = env.reset(), learner.init_hidden()
            obs_store.append(feats(o))
            cur = len(obs_store) - 1
    hook.remove()
    n_seq = len(learner.buffer)
    updates, sample_idx = [], []
    for k in range(n_updates):
        random.seed(2000 + seed + k)
        idx = random.sample(range(n_seq), learner.batch_size)
        random.seed(2000 + seed + k)
        out = learner.update()
        sample_idx.append(idx)
        updates.append(_full_update(learner, out, k == 0))
    keep = {k: config[k] for k in ("agent", "hidden_size", "n_layers", "n_heads", "lr", "gamma", "polyak", "batch_size",
                                   "replay_size", "max_seq_len", "anneal_lr")}
    print(f"learner/{tag}: {len(trans)} transitions, {n_seq} sequences, LossQ {[round(u['LossQ'], 5) for u in updates]}")
    return dict(config=keep, env_info=env_info, obs=obs_store, transitions=trans, seed=seed, policy0=sd0,
                policy0_sums=param_sums(sd0), sample_idx=sample_idx, updates=updates, max_seq_len=learner.max_seq_len)


def make():
    M, D = refshim.madrqn(), refshim.drqn()
    _, MAPS = refshim.mubs_env()
    from envs.mubs_cov.maps import DenseHotSpot, HotSpot
    # the reference's own map classes with a small arena (dense visibility under a random walk) and short episodes
    MAPS["ref8x80"] = DenseHotSpot(range_pos=1600, episode_limit=6, n_ubs=8, n_grps=16)
    MAPS["ref8x80r"] = DenseHotSpot(range_pos=1600, episode_limit=6, n_ubs=8, n_grps=16, r_comm=700.)
    MAPS["ref4x4"] = HotSpot(range_pos=1200, episode_limit=5, r_comm=500.)
    out = {"meta": dict(torch=th.__version__, numpy=np.__version__,
                        note="generated by tests/golden/make_ref_golden.py from the unmodified reference sources")}
    for tag, kw in AGENT_CASES:
        if tag not in LIVE_ONLY:
            out[f"agent/{tag}"] = agent_case(M, tag, keep_sd=(tag == "exp3"), **kw)
    common = dict(device="cpu", hidden_size=64, batch_size=4, replay_size=64, anneal_lr=False, n_layers=1)
    out["learner/exp3"] = learner_case(M, "exp3", "ref8x80", 48, 3, o="gnn", c="tarmac", double_q=True, keep_sd=True,
                                       **common)
    out["learner/exp3-split"] = learner_case(M, "exp3-split", "ref8x80r", 40, 4, o="gnn", c="tarmac", double_q=False,
                                             max_seq_len=4, **common)
    out["learner/qmix"] = learner_case(M, "qmix", "ref8x80", 36, 5, o="gnn", c="tarmac", double_q=True, mixer=True,
                                       share_reward=True, **common)
    out["learner/exp2"] = learner_case(M, "exp2", "ref4x4", 40, 6, o="mlp", c="tarmac", double_q=True, **common)
    out["learner/indep"] = learner_case(M, "indep", "ref8x80", 36, 7, o="gnn", c=None, double_q=True, dueling=True,
                                        **common)
    out["learner/exp1-drqn"] = drqn_case(D, "exp1-drqn", 96, 8, agent="gnn", hidden_size=32, batch_size=4,
                                         replay_size=64, max_seq_len=6, anneal_lr=False)
    return out


if __name__ == "__main__":
    data = make()
    path = os.path.join(HERE, "ref_golden_v1.pt")
    th.save(data, path)
    print(path, os.path.getsize(path), "bytes;", len(data), "entries")
