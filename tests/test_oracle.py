"""Pins for the CPU oracle (parity with the reference itself is unpinned — DGL 0.9.0 is not installable):
independent fp64 loop restatement, torch.nn.GRUCell as a live reference, gradcheck, hand-computed micro cases."""
import math

import numpy as np
import torch as th
import torch.nn as nn
import torch.nn.functional as F
import pytest

from oracle import gnn_oracle as O
from oracle import loops as L
from uav_bs_ctrl_b200.builder import build_obs_graph_batch
from uav_bs_ctrl_b200.synth import synth_dense_obs
from helpers import make_args, batched_graph_ref


def _rand_gat(F_s, F_d, heads, D, dtype=th.float64, seed=0):
    g = th.Generator().manual_seed(seed)
    r = lambda *s: th.randn(*s, generator=g, dtype=dtype)
    H = heads * D
    return dict(fc_src_w=r(H, F_s), fc_src_b=r(H), fc_dst_w=r(H, F_d), fc_dst_b=r(H), attn=r(1, heads, D),
                res_w=r(H, F_d), res_b=r(H))


def _rand_graph(n_src, n_dst, E, seed=0):
    g = th.Generator().manual_seed(seed)
    return th.randint(0, n_src, (E,), generator=g), th.randint(0, n_dst, (E,), generator=g)


def test_gatv2_matches_loops():
    p = _rand_gat(4, 2, 3, 5)
    src, dst = _rand_graph(9, 6, 20, seed=1)
    dst[dst == 4] = 3                                      # guarantee a zero-in-degree destination
    xs, xd = th.randn(9, 4, dtype=th.float64), th.randn(6, 2, dtype=th.float64)
    out = O.gatv2_conv(src, dst, 6, xs, xd, **p)
    ref, _ = L.gatv2_conv_loops(src.numpy(), dst.numpy(), 6, xs.numpy(), xd.numpy(), p['fc_src_w'].numpy(),
                                p['fc_src_b'].numpy(), p['fc_dst_w'].numpy(), p['fc_dst_b'].numpy(),
                                p['attn'].numpy(), p['res_w'].numpy(), p['res_b'].numpy())
    assert np.allclose(out.numpy(), ref, rtol=1e-12, atol=1e-12)
    # zero in-degree ⇒ relu(res_fc(x_dst))   
    z = F.relu(F.linear(xd[4], p['res_w'], p['res_b'])).view(3, 5)
    assert th.allclose(out[4], z)


def test_gatv2_hand_case_single_edge_and_two_edges():
    """1 head, D=1, identity-ish weights: checkable by hand."""
    Ws, bs = th.tensor([[2.0]]), th.tensor([0.5])
    Wd, bd = th.tensor([[1.0]]), th.tensor([0.0])
    attn = th.tensor([[[1.0]]])
    Wr, br = th.tensor([[0.0]]), th.tensor([0.0])
    xs, xd = th.tensor([[1.0], [-3.0]]), th.tensor([[1.0]])
    # one edge: alpha = 1 → out = relu(el) = 2.5
    o1 = O.gatv2_conv(th.tensor([0]), th.tensor([0]), 1, xs, xd, Ws, bs, Wd, bd, attn, Wr, br)
    assert math.isclose(o1.item(), 2.5, rel_tol=1e-6)
    # two edges: el = [2.5, -5.5]; z = el + 1 = [3.5, -4.5]; e = [3.5, -0.9]; alpha = softmax
    a = math.exp(3.5) / (math.exp(3.5) + math.exp(-0.9))
    want = max(a * 2.5 + (1 - a) * -5.5, 0.0)
    o2 = O.gatv2_conv(th.tensor([0, 1]), th.tensor([0, 0]), 1, xs, xd, Ws, bs, Wd, bd, attn, Wr, br)
    assert math.isclose(o2.item(), want, rel_tol=1e-6)


def test_gatv2_permutation_invariant_over_in_edges():
    p = _rand_gat(4, 2, 2, 4)
    src, dst = _rand_graph(12, 4, 30, seed=2)
    xs, xd = th.randn(12, 4, dtype=th.float64), th.randn(4, 2, dtype=th.float64)
    perm = th.randperm(30, generator=th.Generator().manual_seed(0))
    a = O.gatv2_conv(src, dst, 4, xs, xd, **p)
    b = O.gatv2_conv(src[perm], dst[perm], 4, xs, xd, **p)
    assert th.allclose(a, b, rtol=1e-12, atol=1e-12)


def test_edge_softmax_rows_sum_to_one():
    dst = th.tensor([0, 0, 2, 2, 2, 3])
    a = O.edge_softmax(dst, th.randn(6, 4, 1, dtype=th.float64), 5)
    s = th.zeros(5, 4, 1, dtype=th.float64).index_add_(0, dst, a)
    assert th.allclose(s[[0, 2, 3]], th.ones(3, 4, 1, dtype=th.float64)) and float(s[[1, 4]].abs().max()) == 0


def test_gatv2_gradcheck():
    p = {k: v.requires_grad_() for k, v in _rand_gat(3, 2, 2, 3, seed=4).items()}
    src, dst = _rand_graph(5, 4, 9, seed=3)
    xs = th.randn(5, 3, dtype=th.float64, requires_grad=True)
    xd = th.randn(4, 2, dtype=th.float64, requires_grad=True)
    f = lambda *a: O.gatv2_conv(src, dst, 4, *a, activation=None)   # relu kink excluded from fd check
    assert th.autograd.gradcheck(f, (xs, xd, *p.values()), eps=1e-6, atol=1e-5)


def test_gru_cell_matches_torch():
    th.manual_seed(0)
    cell = nn.GRUCell(7, 5).double()
    x, h = th.randn(4, 7, dtype=th.float64), th.randn(4, 5, dtype=th.float64)
    ours = O.gru_cell(x, h, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)
    assert th.allclose(ours, cell(x, h), rtol=1e-12, atol=1e-12)
    lo = L.gru_cell_loops(x.numpy(), h.numpy(), *(t.detach().numpy() for t in
                                                  (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)))
    assert np.allclose(lo, cell(x, h).detach().numpy(), rtol=1e-12, atol=1e-12)


def test_tarmac_matches_loops_and_detach():
    th.manual_seed(1)
    args = make_args(hidden_size=6, msg_size=5, key_size=3, n_rounds=2)
    m = O.TarMAC(args).double()
    src, dst = _rand_graph(4, 4, 7, seed=5)
    src, dst = th.cat([src, th.arange(4)]), th.cat([dst, th.arange(4)])   # self loops as in the env
    x = th.randn(4, 6, dtype=th.float64)
    h = th.randn(4, 6, dtype=th.float64, requires_grad=True)

    class R:
        def edges(self):
            return src, dst
    out = m(R(), x, h)
    P = lambda t: t.detach().numpy()
    u = m.f_udt
    ref = L.tarmac_loops(src.numpy(), dst.numpy(), x.numpy(), P(h), P(m.f_val.weight), P(m.f_val.bias),
                         P(m.f_sign.weight), P(m.f_sign.bias), P(m.f_que.weight), P(m.f_que.bias),
                         P(u.weight_ih), P(u.weight_hh), P(u.bias_ih), P(u.bias_hh), 3, 2)
    assert np.allclose(P(out), ref, rtol=1e-11, atol=1e-12)
    # scores are divided by key_size itself, not sqrt (gnn_agents.py:262): a sqrt version must differ
    inp = th.cat((x, h.detach()), 1)
    s, q = m.f_sign(inp), m.f_que(inp)
    e = (s[src] * q[dst]).sum(-1) / 3
    assert not th.allclose(e, (s[src] * q[dst]).sum(-1) / math.sqrt(3))


def test_agent_shapes_state_dict_keys_and_param_count():
    """SURVEY §8(b) state_dict contract and §8 table footnote 2 (59 881 params at H=64 / TarMAC)."""
    args = make_args()
    agent = O.GnnAgent({'agent': 2, 'ubs': 2, 'gt': 4}, 9, args)
    keys = set(agent.state_dict().keys())
    for rel in ('seen', 'near'):
        for k in ('fc_src.weight', 'fc_src.bias', 'fc_dst.weight', 'fc_dst.bias', 'attn', 'res_fc.weight', 'res_fc.bias'):
            assert f'enc.f_conv.{rel}.{k}' in keys
    for k in ('enc.f_aggr.0.weight', 'enc.f_aggr.0.bias', 'f_comm.f_val.weight', 'f_comm.f_sign.bias',
              'f_comm.f_que.weight', 'f_comm.f_udt.weight_ih', 'f_comm.f_udt.weight_hh', 'f_comm.f_udt.bias_ih',
              'f_comm.f_udt.bias_hh', 'f_out.weight', 'f_out.bias'):
        assert k in keys
    assert sum(p.numel() for p in agent.parameters()) == 59881
    a, gt, ubs, adj = synth_dense_obs(2, 8, 10, "realistic", seed=2)
    g = build_obs_graph_batch(a, gt, ubs, adj)
    q, h = agent(g, agent.init_hidden().expand(16, -1))
    assert q.shape == (16, 9) and h.shape == (16, 64)
    # same answer through the reference-style per-agent graph construction
    q2, h2 = agent(batched_graph_ref(a, gt, ubs, adj), agent.init_hidden().expand(16, -1))
    assert th.equal(q, q2) and th.equal(h, h2)


def test_drqn_agent_exp1_config():
    """BASELINE configs[0]: exp1 single-UBS DRQN, 1 UBS × 10 GT, hidden=32, 1 env on CPU."""
    from uav_bs_ctrl_b200.builder import build_drqn_graph_batch
    th.manual_seed(0)
    args = make_args(hidden_size=32, n_heads=4)
    agent = O.DrqnGnnAgent({'agent': 2, 'gt': 4}, 5, args)
    assert sum(p.numel() for p in agent.parameters()) == 6885
    assert set(agent.state_dict()) >= {'enc.fc_src.weight', 'enc.attn', 'enc.res_fc.bias', 'rnn.weight_ih', 'f_out.bias'}
    g = build_drqn_graph_batch(th.rand(1, 2), th.rand(1, 10, 4))
    h = agent.init_hidden()
    for _ in range(10):
        q, h = agent(g, h)
    assert q.shape == (1, 5) and h.shape == (1, 32) and bool(th.isfinite(q).all())


def test_dgl_init_statistics():
    """reset_parameters: xavier_normal_(gain=sqrt 2) on weights/attn, zero biases."""
    th.manual_seed(0)
    conv = O.GATv2Conv((4, 2), 64, 8, residual=True, allow_zero_in_degree=True)
    assert float(conv.fc_src.bias.abs().max()) == 0 and float(conv.res_fc.bias.abs().max()) == 0
    std = math.sqrt(2) * math.sqrt(2.0 / (4 + 512))
    assert abs(float(conv.fc_src.weight.std()) - std) / std < 0.1
