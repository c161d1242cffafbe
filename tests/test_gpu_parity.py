"""GPU parity: CUDA kernels (through the C ABI / modules) vs the CPU oracle on identical seeded inputs.
Tolerance: allclose(rtol=1e-5, atol=1e-6 * max|ref|) against the fp32 oracle; the fp64 oracle gives
the noise floor."""
import copy

import pytest
import torch as th
import torch.nn as nn
import torch.nn.functional as F

from oracle import gnn_oracle as O
from uav_bs_ctrl_b200 import ops, agents as A, graph as G
from uav_bs_ctrl_b200.builder import build_obs_graph_batch, build_drqn_graph_batch
from uav_bs_ctrl_b200.synth import synth_dense_obs
from helpers import make_args, assert_close, assert_as_accurate, batched_graph_ref

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _gat_params(F_s, F_d, heads, D, seed):
    g = th.Generator().manual_seed(seed)
    r = lambda *s: (th.randn(*s, generator=g) * 0.7)
    H = heads * D
    return dict(fc_src_w=r(H, F_s), fc_src_b=r(H) * 0.3, fc_dst_w=r(H, F_d), fc_dst_b=r(H) * 0.3, attn=r(1, heads, D),
                res_w=r(H, F_d), res_b=r(H) * 0.3)


def _run_gat(p, xs, xd, indptr, src_idx, heads, D, need_x_grad):
    P = {k: v.clone().to(DEV).requires_grad_() for k, v in p.items()}
    xs_d, xd_d = xs.to(DEV).requires_grad_(need_x_grad), xd.to(DEV).requires_grad_(need_x_grad)
    out = ops.GATv2Fused.apply(xs_d, xd_d, indptr.to(DEV), None if src_idx is None else src_idx.to(DEV),
                               P['fc_src_w'], P['fc_src_b'], P['fc_dst_w'], P['fc_dst_b'], P['attn'], P['res_w'],
                               P['res_b'], heads, D, 0.2, ops.GAT_RESIDUAL | ops.GAT_RELU)
    return out, P, xs_d, xd_d


@pytest.mark.parametrize("F_s,heads,D,n_dst,deg_kind,star", [
    (4, 4, 16, 64, "full80", True),        # exp3 'seen'
    (2, 4, 16, 64, "seven", True),         # exp3 'near'
    (4, 4, 8, 7, "ten", True),             # exp1 (H=32)
    (4, 4, 32, 40, "ragged", True),        # scaled (H=128), degrees 0..150
    (3, 2, 16, 33, "ragged", True),        # fair_service=False (F_gt=3)
    (4, 8, 8, 50, "ragged", False),        # general CSR with src_idx, shared sources
    (1, 1, 32, 19, "ragged", False),
])
def test_gatv2_fused_matches_oracle(F_s, heads, D, n_dst, deg_kind, star):
    g = th.Generator().manual_seed(7)
    if deg_kind == "full80":
        deg = th.full((n_dst,), 80)
    elif deg_kind == "seven":
        deg = th.full((n_dst,), 7)
    elif deg_kind == "ten":
        deg = th.full((n_dst,), 10)
    else:
        deg = th.randint(0, 151, (n_dst,), generator=g)
        deg[::5] = 0
        deg[1] = 1
    E = int(deg.sum())
    indptr = th.zeros(n_dst + 1, dtype=th.int32)
    indptr[1:] = th.cumsum(deg, 0)
    dst = th.repeat_interleave(th.arange(n_dst), deg)
    if star:
        n_src, src, src_idx = E, th.arange(E), None
    else:
        n_src = max(E // 3, 1)
        src = th.randint(0, n_src, (E,), generator=g)
        src_idx = src.to(th.int32)
    xs = th.rand(n_src, F_s, generator=g) * 2 - 1
    xd = th.rand(n_dst, 2, generator=g)
    p = _gat_params(F_s, 2, heads, D, seed=F_s * 10 + heads)
    out, P, xs_d, xd_d = _run_gat(p, xs, xd, indptr, src_idx, heads, D, need_x_grad=True)
    go = th.randn(n_dst, heads * D, generator=g)
    out.backward(go.to(DEV))

    refs = {}
    for dt in (th.float32, th.float64):
        Pr = {k: v.clone().to(dt).requires_grad_() for k, v in p.items()}
        xs_r, xd_r = xs.clone().to(dt).requires_grad_(), xd.clone().to(dt).requires_grad_()
        ref = O.gatv2_conv(src, dst, n_dst, xs_r, xd_r, **Pr).view(n_dst, -1)
        ref.backward(go.to(dt))
        refs[dt] = (ref, Pr, xs_r, xd_r)
    r32, r64 = refs[th.float32], refs[th.float64]
    assert_close(out, r32[0], what="out")
    assert float((out.detach().cpu().double() - r64[0].detach()).abs().max() / r64[0].detach().abs().max()) < 2e-6
    for k in p:
        assert_as_accurate(P[k].grad, r32[1][k].grad, r64[1][k].grad, what=f"grad {k}")
    assert_as_accurate(xs_d.grad, r32[2].grad, r64[2].grad, what="grad x_src")
    assert_as_accurate(xd_d.grad, r32[3].grad, r64[3].grad, what="grad x_dst")
    # zero in-degree rows equal relu(res_fc(x_dst))
    z = (deg == 0).nonzero().flatten()
    if z.numel():
        want = F.relu(F.linear(xd[z], p['res_w'], p['res_b']))
        assert_close(out[z.to(DEV)], want, what="zero-degree rows")


def test_gatv2_inference_needs_no_stats_and_is_deterministic():
    p = _gat_params(4, 2, 4, 16, 1)
    deg = th.full((128,), 80)
    indptr = th.zeros(129, dtype=th.int32)
    indptr[1:] = th.cumsum(deg, 0)
    xs, xd = th.rand(128 * 80, 4), th.rand(128, 2)
    with th.no_grad():
        a, *_ = _run_gat(p, xs, xd, indptr, None, 4, 16, False)
        b, *_ = _run_gat(p, xs, xd, indptr, None, 4, 16, False)
    assert th.equal(a, b)
    o1, P1, *_ = _run_gat(p, xs, xd, indptr, None, 4, 16, False)
    o2, P2, *_ = _run_gat(p, xs, xd, indptr, None, 4, 16, False)
    o1.sum().backward(), o2.sum().backward()
    for k in p:
        assert th.equal(P1[k].grad, P2[k].grad), f"{k}: parameter gradients must be run-to-run deterministic"


@pytest.mark.parametrize("U,K,M,B,comm_p", [(8, 16, 64, 32, 1.0), (8, 16, 64, 17, 0.5), (4, 16, 64, 5, 0.3),
                                             (16, 16, 64, 9, 0.7), (3, 5, 7, 4, 0.5), (32, 8, 40, 2, 0.5)])
def test_block_attention_matches_oracle(U, K, M, B, comm_p):
    g = th.Generator().manual_seed(3)
    N = B * U
    adj = (th.rand(B, U, U, generator=g) < comm_p) | th.eye(U, dtype=th.bool)
    b, i, j = th.nonzero(adj, as_tuple=True)
    src, dst = b * U + i, b * U + j
    mask = (adj.long() << th.arange(U).view(1, U, 1)).sum(1).flatten().to(th.int32)
    vsq = th.randn(N, M + 2 * K, generator=g)
    gc = th.randn(N, M, generator=g)
    v_d = vsq.to(DEV).requires_grad_()
    c = ops.BlockAttention.apply(v_d, mask.to(DEV), U, K, M, 1.0 / K)
    c.backward(gc.to(DEV))
    refs = {}
    for dt in (th.float32, th.float64):
        v_r = vsq.clone().to(dt).requires_grad_()
        val, s, q = v_r[:, :M], v_r[:, M:M + K], v_r[:, M + K:]
        e = (s[src] * q[dst]).sum(-1, keepdim=True) / K
        a = O.edge_softmax(dst, e, N)
        ref = th.zeros(N, M, dtype=dt).index_add_(0, dst, val[src] * a)
        ref.backward(gc.to(dt))
        refs[dt] = (ref, v_r.grad)
    assert_close(c, refs[th.float32][0], what="c")
    assert_as_accurate(v_d.grad, refs[th.float32][1], refs[th.float64][1], what="grad vsq")


@pytest.mark.parametrize("N,I,H", [(2048, 128, 64), (5, 32, 32), (300, 256, 128)])
def test_gru_cell_matches_torch(N, I, H):
    th.manual_seed(0)
    ref = nn.GRUCell(I, H)
    ref64 = copy.deepcopy(ref).double()
    mine = A.GRUCell(I, H).to(DEV)
    mine.load_state_dict(ref.state_dict())
    x, h = th.randn(N, I), th.randn(N, H)
    xr, hr = x.clone().requires_grad_(), h.clone().requires_grad_()
    x6, h6 = x.double().requires_grad_(), h.double().requires_grad_()
    xd, hd = x.to(DEV).requires_grad_(), h.to(DEV).requires_grad_()
    go = th.randn(N, H)
    o_r = ref(xr, hr)
    o_r.backward(go)
    ref64(x6, h6).backward(go.double())
    o_d = mine(xd, hd)
    o_d.backward(go.to(DEV))
    assert_close(o_d, o_r, what="h'")
    assert_as_accurate(xd.grad, xr.grad, x6.grad, what="grad x")
    assert_as_accurate(hd.grad, hr.grad, h6.grad, what="grad h")
    for (k, a), (_, b), (_, b6) in zip(mine.named_parameters(), ref.named_parameters(), ref64.named_parameters()):
        assert_as_accurate(a.grad, b.grad, b6.grad, what=f"grad {k}")


def _agent_pair(args, obs_shape, n_actions, seed=0, cls=("GnnAgent", "GnnAgent")):
    th.manual_seed(seed)
    ref = getattr(O, cls[1])(obs_shape, n_actions, args)
    mine = getattr(A, cls[0])(obs_shape, n_actions, args).to(DEV)
    mine.load_state_dict(ref.state_dict())
    return mine, ref


def _g64(g):
    return g._map(lambda t: t.double() if t.is_floating_point() else t, lambda r: r)


def _check_param_grads(mine, ref, ref64):
    for (k, a), (_, b), (_, b6) in zip(mine.named_parameters(), ref.named_parameters(), ref64.named_parameters()):
        assert_as_accurate(a.grad, b.grad, b6.grad, what=f"grad {k}", slack=6.0, floor_scale=5e-6)


@pytest.mark.parametrize("c,profile,comm_p,dueling", [("tarmac", "full", 1.0, False), ("tarmac", "realistic", 0.5, True),
                                                      (None, "random", 1.0, False), ("base", "realistic", 0.6, False),
                                                      ("commnet", "realistic", 0.6, False)])
def test_madrqn_agent_bptt_matches_oracle(c, profile, comm_p, dueling):
    """exp3 shapes (8 UBS x 80 GT, H=64), B=12 envs, 3 unrolled steps, loss = sum of squares of Q."""
    args = make_args(c=c, dueling=dueling)
    mine, ref = _agent_pair(args, {'agent': 2, 'ubs': 2, 'gt': 4}, 9)
    B, U, Gn, T = 12, 8, 80, 3
    graphs = [build_obs_graph_batch(*synth_dense_obs(B, U, Gn, profile, seed=100 + t, comm_p=comm_p)) for t in range(T)]
    ref64 = copy.deepcopy(ref).double()
    h_r = ref.init_hidden().expand(B * U, -1)
    h_6 = h_r.double()
    h_d = mine.init_hidden().expand(B * U, -1).to(DEV)
    loss_r = loss_d = loss_6 = 0
    for t in range(T):
        q_r, h_r = ref(graphs[t], h_r)
        q_6, h_6 = ref64(_g64(graphs[t]), h_6)
        q_d, h_d = mine(graphs[t].to(DEV), h_d)
        assert_close(q_d, q_r, rtol=2e-5, atol_scale=2e-6, what=f"q[{t}]")
        assert_close(h_d, h_r, rtol=2e-5, atol_scale=2e-6, what=f"h[{t}]")
        loss_r = loss_r + (q_r ** 2).mean()
        loss_6 = loss_6 + (q_6 ** 2).mean()
        loss_d = loss_d + (q_d ** 2).mean()
    loss_r.backward(), loss_d.backward(), loss_6.backward()
    _check_param_grads(mine, ref, ref64)


def test_reference_style_graph_gives_same_answer_as_builder():
    args = make_args()
    mine, ref = _agent_pair(args, {'agent': 2, 'ubs': 2, 'gt': 4}, 9, seed=3)
    a, gt, ubs, adj = synth_dense_obs(3, 4, 9, "realistic", seed=5, comm_p=0.5)
    g1 = batched_graph_ref(a, gt, ubs, adj).to(DEV)
    g2 = build_obs_graph_batch(a, gt, ubs, adj).to(DEV)
    h = mine.init_hidden().expand(12, -1).to(DEV)
    with th.no_grad():
        q1, h1 = mine(g1, h)
        q2, h2 = mine(g2, h)
        qr, hr = ref(batched_graph_ref(a, gt, ubs, adj), ref.init_hidden().expand(12, -1))
    assert th.equal(q1, q2) and th.equal(h1, h2)
    assert_close(q1, qr, rtol=2e-5, atol_scale=2e-6, what="q")


def test_drqn_agent_exp1_matches_oracle():
    """BASELINE configs[0]: 1 UBS x 10 GT, hidden=32, 10-step sequence."""
    args = make_args(hidden_size=32)
    mine, ref = _agent_pair(args, {'agent': 2, 'gt': 4}, 5, cls=("DrqnGnnAgent", "DrqnGnnAgent"))
    ref64 = copy.deepcopy(ref).double()
    h_r, h_d = ref.init_hidden(), mine.init_hidden().to(DEV)
    h_6 = h_r.double()
    loss_r = loss_d = loss_6 = 0
    for t in range(10):
        th.manual_seed(t)
        g = build_drqn_graph_batch(th.rand(1, 2), th.rand(1, 10, 4))
        q_r, h_r = ref(g, h_r)
        q_6, h_6 = ref64(_g64(g), h_6)
        q_d, h_d = mine(g.to(DEV), h_d)
        loss_r, loss_d, loss_6 = loss_r + q_r.pow(2).sum(), loss_d + q_d.pow(2).sum(), loss_6 + q_6.pow(2).sum()
    assert_close(q_d, q_r, rtol=2e-5, atol_scale=2e-6, what="q")
    loss_r.backward(), loss_d.backward(), loss_6.backward()
    _check_param_grads(mine, ref, ref64)


def test_full_size_properties_exp3():
    """BASELINE full size (B=256, U=8, G=80): size-independent properties instead of a slow CPU oracle run."""
    args = make_args()
    th.manual_seed(0)
    mine = A.GnnAgent({'agent': 2, 'ubs': 2, 'gt': 4}, 9, args).to(DEV)
    B, U, Gn = 256, 8, 80
    a, gt, ubs, adj = synth_dense_obs(B, U, Gn, "full", seed=1)
    g = build_obs_graph_batch(a, gt, ubs, adj).to(DEV)
    h = mine.init_hidden().expand(B * U, -1).to(DEV)
    with th.no_grad():
        q, h1 = mine(g, h)
        # (1) batching independence: the first 16 envs alone give the same rows
        gs = build_obs_graph_batch(a[:16], gt[:16], ubs[:16], adj[:16]).to(DEV)
        qs, hs = mine(gs, h[:16 * U])
        assert th.equal(q[:16 * U], qs) and th.equal(h1[:16 * U], hs)
        # (2) permutation invariance over the in-edges (GT order) of every agent
        perm = th.randperm(Gn)
        gp = build_obs_graph_batch(a, gt[:, :, perm], ubs, adj).to(DEV)
        qp, _ = mine(gp, h)
        assert_close(qp, q, rtol=1e-4, atol_scale=1e-5, what="GT-order permutation")
        # (3) an invisible GT changes nothing
        gt2 = gt.clone()
        gt2[:, :, 0, 0] = 0
        gt3 = gt2.clone()
        gt3[:, :, 0, 1:] = 123.0
        q2, _ = mine(build_obs_graph_batch(a, gt2, ubs, adj).to(DEV), h)
        q3, _ = mine(build_obs_graph_batch(a, gt3, ubs, adj).to(DEV), h)
        assert th.equal(q2, q3)
    assert bool(th.isfinite(q).all())
