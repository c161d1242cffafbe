"""Generates the committed golden vectors under tests/golden/ from the CPU oracle (oracle/gnn_oracle.py, float64
then rounded to float32).  The reference itself (DGL 0.9.0) cannot be run here, so these pin OUR restatement:
`python tests/golden/make_golden.py` must reproduce the committed files bit for bit (tests/test_golden.py)."""
import os
import sys
from types import SimpleNamespace

import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import gnn_oracle as O  # noqa: E402
from uav_bs_ctrl_b200.builder import build_obs_graph_batch  # noqa: E402
from uav_bs_ctrl_b200.synth import synth_dense_obs  # noqa: E402


def args_ns(**kw):
    d = dict(hidden_size=64, n_layers=1, n_heads=4, msg_size=64, key_size=16, n_rounds=1, c="tarmac", o="gnn", dueling=False)
    d.update(kw)
    return SimpleNamespace(**d)


def make():
    out = {}
    # 1) one GATv2 relation, exp3 'seen' shape: 6 destinations, ragged degrees incl. 0, F_s=4, F_d=2, heads=4, D=16
    g = th.Generator().manual_seed(11)
    deg = th.tensor([80, 0, 3, 1, 17, 42])
    E = int(deg.sum())
    dst = th.repeat_interleave(th.arange(6), deg)
    src = th.arange(E)
    r = lambda *s: th.randn(*s, generator=g, dtype=th.float64)
    p = dict(fc_src_w=r(64, 4) * .7, fc_src_b=r(64) * .2, fc_dst_w=r(64, 2) * .7, fc_dst_b=r(64) * .2,
             attn=r(1, 4, 16) * .7, res_w=r(64, 2) * .7, res_b=r(64) * .2)
    xs, xd = th.rand(E, 4, generator=g, dtype=th.float64) * 2 - 1, th.rand(6, 2, generator=g, dtype=th.float64)
    y = O.gatv2_conv(src, dst, 6, xs, xd, **p)
    out["gatv2"] = dict(deg=deg, x_src=xs.float(), x_dst=xd.float(), out=y.float().view(6, 64),
                        **{k: v.float() for k, v in p.items()})
    # 2) exp3 agent (TarMAC), 2 envs x 8 UBS x 80 GT, 3 unrolled steps from h = 0
    th.manual_seed(21)
    agent = O.GnnAgent({'agent': 2, 'ubs': 2, 'gt': 4}, 9, args_ns()).double()
    obs = [synth_dense_obs(2, 8, 80, "realistic", seed=31 + t, comm_p=0.6) for t in range(3)]
    h = agent.init_hidden().expand(16, -1).double()
    qs = []
    with th.no_grad():
        for o in obs:
            gph = build_obs_graph_batch(*o)
            gph = gph._map(lambda t: t.double() if t.is_floating_point() else t, lambda c: c)
            q, h = agent(gph, h)
            qs.append(q.float())
    out["agent_exp3"] = dict(state_dict={k: v.float() for k, v in agent.state_dict().items()}, obs=obs,
                             q=th.stack(qs), h=h.float())
    # 3) exp1 DRQN agent: 1 UBS x 10 GT, hidden 32, 4 steps
    th.manual_seed(41)
    d = O.DrqnGnnAgent({'agent': 2, 'gt': 4}, 5, args_ns(hidden_size=32)).double()
    gen = th.Generator().manual_seed(43)
    ag, gt = th.rand(4, 1, 2, generator=gen), th.rand(4, 1, 10, 4, generator=gen)
    from uav_bs_ctrl_b200.builder import build_drqn_graph_batch
    h = d.init_hidden().double()
    qs = []
    with th.no_grad():
        for t in range(4):
            q, h = d(build_drqn_graph_batch(ag[t].double(), gt[t].double()), h)
            qs.append(q.float())
    out["agent_exp1"] = dict(state_dict={k: v.float() for k, v in d.state_dict().items()}, agent_obs=ag, gt_obs=gt,
                             q=th.stack(qs), h=h.float())
    return out


if __name__ == "__main__":
    th.save(make(), os.path.join(HERE, "golden_v1.pt"))
    print("wrote", os.path.join(HERE, "golden_v1.pt"))
