"""GPU parity of the fused agent step / persistent sequence kernels (ubs_agent_seq_fwd / _bwd) against the CPU oracle."""
import copy

import pytest
import torch as th

from oracle import gnn_oracle as O
from uav_bs_ctrl_b200 import agents as A, ops
from uav_bs_ctrl_b200.builder import build_obs_graph_batch
from uav_bs_ctrl_b200.synth import synth_dense_obs
from helpers import make_args, assert_close, assert_as_accurate

pytestmark = pytest.mark.gpu
DEV = "cuda"
SHAPE = {'agent': 2, 'ubs': 2, 'gt': 4}


def _pair(args, obs_shape=SHAPE, n_actions=9, seed=0):
    th.manual_seed(seed)
    ref = O.GnnAgent(obs_shape, n_actions, args)
    mine = A.GnnAgent(obs_shape, n_actions, args).to(DEV)
    mine.load_state_dict(ref.state_dict())
    return mine, ref


def _g64(g):
    return g._map(lambda t: t.double() if t.is_floating_point() else t, lambda r: r)


def _graphs(B, U, Gn, T, profile, comm_p, flat=None, seed=100):
    out = []
    for t in range(T):
        a, gt, ubs, adj = synth_dense_obs(B, U, Gn, profile, seed=seed + t, comm_p=comm_p)
        if flat is not None:        # o='mlp': flat local observations, comm graph only (exp2)
            gen = th.Generator().manual_seed(seed + 50 + t)
            a = th.rand(B, U, flat, generator=gen)
            gt, ubs = th.zeros(B, U, 0, 5), th.zeros(B, U, 0, 3)
        out.append(build_obs_graph_batch(a, gt, ubs, adj))
    return out


CASES = [
    # c,       H,  U, G,  B,  profile,     comm_p, flat
    ("tarmac", 64, 8, 80, 6, "full", 1.0, None),          # exp3
    ("tarmac", 64, 8, 20, 5, "realistic", 0.5, None),     # thinned talk graph, ragged degrees, B*U not a tile multiple
    ("tarmac", 64, 3, 10, 7, "realistic", 0.6, None),     # 3 agents: 5 envs per 15-row tile
    ("tarmac", 64, 16, 12, 3, "realistic", 0.7, None),    # scaled config's 16 UBS
    ("tarmac", 32, 4, 10, 9, "random", 1.0, None),        # H=32
    ("tarmac", 128, 8, 10, 4, "realistic", 0.8, None),    # H=128 (scaled)
    (None, 64, 8, 30, 5, "realistic", 1.0, None),         # independent agents: plain GRU
    ("tarmac", 64, 8, 0, 6, "full", 0.7, 423),            # exp2: MLP observation encoder + TarMAC
    ("tarmac", 64, 8, 12, 16, "realistic", 0.8, None),    # T*N = 512 rows: window projections run on tcgen05 (3xTF32)
]


@pytest.mark.parametrize("c,H,U,Gn,B,profile,comm_p,flat", CASES)
def test_fused_inference_step_matches_oracle(c, H, U, Gn, B, profile, comm_p, flat):
    args = make_args(c=c, hidden_size=H, o="mlp" if flat else "gnn", n_layers=2)
    mine, ref = _pair(args, obs_shape=flat if flat else SHAPE)
    graphs = _graphs(B, U, Gn, 3, profile, comm_p, flat)
    assert mine.can_fuse(graphs[0].to(DEV))
    ref64 = copy.deepcopy(ref).double()
    h_r = ref.init_hidden().expand(B * U, -1)
    h_6 = h_r.double()
    h_d = mine.init_hidden().expand(B * U, -1).to(DEV)
    ops.TIMER = ops.KernelTimer()
    with th.no_grad():
        for t in range(3):
            q_r, h_r = ref(graphs[t], h_r)
            q_6, h_6 = ref64(_g64(graphs[t]), h_6)
            q_d, h_d = mine(graphs[t].to(DEV), h_d)
            # 1e-5-class agreement with the fp32 oracle, and as close to the fp64 oracle as the fp32 oracle is
            assert_close(q_d, q_r, rtol=3e-5, atol_scale=1e-5, what=f"q[{t}]")
            assert_close(h_d, h_r, rtol=3e-5, atol_scale=1e-5, what=f"h[{t}]")
            assert_as_accurate(q_d, q_r, q_6, what=f"q[{t}] vs fp64", slack=4.0, floor_scale=3e-6)
            assert_as_accurate(h_d, h_r, h_6, what=f"h[{t}] vs fp64", slack=4.0, floor_scale=3e-6)
    used = ops.TIMER.summary()
    ops.TIMER = None
    assert used.get("agent_seq_fwd", {}).get("count") == 3, "the fused kernel must be the path that ran"


@pytest.mark.parametrize("use_seq2", [True, False], ids=["resident", "streaming"])
@pytest.mark.parametrize("c,H,U,Gn,B,profile,comm_p,flat", CASES)
def test_forward_sequence_bptt_matches_oracle(c, H, U, Gn, B, profile, comm_p, flat, use_seq2):
    T = 4
    args = make_args(c=c, hidden_size=H, o="mlp" if flat else "gnn", n_layers=2)
    mine, ref = _pair(args, obs_shape=flat if flat else SHAPE, seed=1)
    mine.use_seq2 = use_seq2
    ref64 = copy.deepcopy(ref).double()
    graphs = _graphs(B, U, Gn, T, profile, comm_p, flat, seed=300)
    gen = th.Generator().manual_seed(5)
    h0 = th.randn(B * U, H, generator=gen) * 0.3
    w = th.randn(T, B * U, 9, generator=gen)                     # arbitrary linear functional of the Q values
    hw = th.randn(B * U, H, generator=gen)                       # ... and of the final hidden state
    outs = {}
    for name, net, dt in (("r32", ref, th.float32), ("r64", ref64, th.float64)):
        h, qs = h0.to(dt), []
        for t in range(T):
            q, h = net(graphs[t] if dt == th.float32 else _g64(graphs[t]), h)
            qs.append(q)
        qs = th.stack(qs)
        ((qs * w.to(dt)).sum() + (qs ** 2).mean() + (h * hw.to(dt)).sum()).backward()
        outs[name] = (qs, h)
    ops.TIMER = ops.KernelTimer()
    q_d, h_d = mine.forward_sequence([g.to(DEV) for g in graphs], h0.to(DEV))
    ((q_d * w.to(DEV)).sum() + (q_d ** 2).mean() + (h_d * hw.to(DEV)).sum()).backward()
    used = ops.TIMER.summary()
    ops.TIMER = None
    if use_seq2 and H <= 64:
        assert "agent_seq2_fwd" in used and "agent_seq2_bwd" in used, "resident-weight kernels must be the path that ran"
        assert ("tf32x3_gemm" in used) == (T * B * U >= 512), "tensor-core projections for windows of >= 512 rows"
    else:
        assert "agent_seq_fwd" in used and "agent_seq_bwd" in used
    assert_close(q_d, outs["r32"][0], rtol=3e-5, atol_scale=1e-5, what="q sequence")
    assert_close(h_d, outs["r32"][1], rtol=3e-5, atol_scale=1e-5, what="h last")
    assert_as_accurate(q_d, outs["r32"][0], outs["r64"][0], what="q sequence vs fp64", slack=4.0, floor_scale=3e-6)
    assert_as_accurate(h_d, outs["r32"][1], outs["r64"][1], what="h last vs fp64", slack=4.0, floor_scale=3e-6)
    for (k, a), (_, b), (_, b6) in zip(mine.named_parameters(), ref.named_parameters(), ref64.named_parameters()):
        assert_as_accurate(a.grad, b.grad, b6.grad, what=f"grad {k}", slack=6.0, floor_scale=5e-6)


def test_forward_sequence_no_grad_equals_stepwise_calls():
    args = make_args()
    mine, _ = _pair(args, seed=2)
    B, U, T = 10, 8, 5
    graphs = [g.to(DEV) for g in _graphs(B, U, 40, T, "realistic", 0.6)]
    h = mine.init_hidden().expand(B * U, -1).to(DEV)
    with th.no_grad():
        q_seq, h_last = mine.forward_sequence(graphs, h)
        qs = []
        for t in range(T):
            q, h = mine(graphs[t], h)
            qs.append(q)
    # the window kernel splits every projection into an observation part (batched GEMM) and a recurrent part, so
    # the sums associate differently from the per-step kernel: tight tolerance instead of bitwise equality
    assert_close(q_seq, th.stack(qs), rtol=1e-5, atol_scale=2e-6, what="q")
    assert_close(h_last, h, rtol=1e-5, atol_scale=2e-6, what="h")
    mine.use_seq2 = False
    with th.no_grad():
        q_str, h_str = mine.forward_sequence(graphs, mine.init_hidden().expand(B * U, -1).to(DEV))
    assert th.equal(q_str, th.stack(qs)) and th.equal(h_str, h)       # same kernel per step or per window


def test_pack_cache_follows_parameter_updates():
    args = make_args()
    mine, ref = _pair(args, seed=3)
    g = _graphs(4, 8, 10, 1, "full", 1.0)[0]
    h_d = mine.init_hidden().expand(32, -1).to(DEV)
    with th.no_grad():
        q0, _ = mine(g.to(DEV), h_d)
        for p in mine.parameters():
            p.mul_(1.05)
        for p in ref.parameters():
            p.mul_(1.05)
        q1, _ = mine(g.to(DEV), h_d)
        q_r, _ = ref(g, ref.init_hidden().expand(32, -1))
    assert not th.equal(q0, q1)
    assert_close(q1, q_r, rtol=2e-5, atol_scale=3e-6, what="q after in-place parameter update")


def test_learner_fused_update_matches_stepwise_update():
    """MultiAgentQLearner.update with the sequence-fused path vs the per-step module path (same weights, same batch)."""
    from types import SimpleNamespace
    from uav_bs_ctrl_b200.learner import MultiAgentQLearner
    B, U, T = 6, 8, 5

    def make(fused):
        th.manual_seed(0)
        a = SimpleNamespace(device=DEV, o="gnn", c="tarmac", share_reward=False, hidden_size=64, n_layers=1, n_heads=4,
                            msg_size=64, key_size=16, n_rounds=1, lr=1e-3, gamma=0.99, polyak=0.9, batch_size=1,
                            replay_size=2, max_seq_len=T, anneal_lr=False, double_q=True, dueling=False, mixer=False,
                            n_envs=B, fused=fused)
        return MultiAgentQLearner(dict(obs_shape=SHAPE, state_shape=None, n_actions=9, n_agents=U, episode_limit=T), a)

    graphs = [g.to(DEV) for g in _graphs(B, U, 30, T + 1, "realistic", 0.7, seed=900)]
    gen = th.Generator().manual_seed(1)
    rews = th.rand(T, B, U, generator=gen).to(DEV)
    dones = th.zeros(T, B, device=DEV)
    results = []
    for fused in (True, False):
        L = make(fused)
        th.manual_seed(7)
        h = L.init_hidden(B).to(DEV)
        for t in range(T):
            acts, h2 = L.act(graphs[t], h, 0.3)
            L.cache(graphs[t], h, None, acts, rews[t], graphs[t + 1], h2, None, dones[t], dones[t])
            h = h2
        out = L.update(samples=[L.buffer.memory[-1]])
        # compare the (clipped) gradients the optimiser saw: AdamW's first step normalises even pure rounding noise
        # to +-lr, so post-step parameters are not a meaningful comparison for mathematically-zero gradients
        results.append((out["LossQ"], L.grad_bucket.flat.detach().clone()))
    assert abs(results[0][0] - results[1][0]) <= 1e-5 * abs(results[1][0])
    assert_close(results[0][1], results[1][1], rtol=1e-4, atol_scale=1e-5, what="policy gradients seen by AdamW")


@pytest.mark.parametrize("B,U,F_,p", [(5, 8, 64, 0.6), (3, 4, 7, 0.3), (2, 16, 130, 0.9), (4, 1, 16, 1.0)])
def test_block_mean_matches_the_edge_list_path(B, U, F_, p):
    """ubs_block_mean_fwd / bwd (BaseComm / CommNet reduce) vs gather + index_add over the explicit edge list, including
    destinations without in-edges; deterministic."""
    from uav_bs_ctrl_b200 import ops
    g = th.Generator().manual_seed(B * 100 + U)
    adj = th.rand(B, U, U, generator=g) < p                       # adj[b, i, j]: edge i -> j
    adj[0, :, 0] = False                                          # a destination nobody talks to
    N = B * U
    mask = (adj.to(th.int64) << th.arange(U).view(1, U, 1)).sum(1).flatten().to(th.int32)
    b, i, j = th.nonzero(adj, as_tuple=True)
    src, dst = b * U + i, b * U + j
    msg = th.randn(N, F_, generator=g)
    go = th.randn(N, F_, generator=g)
    m_ref = msg.double().requires_grad_(True)
    out_ref = th.zeros(N, F_, dtype=th.float64).index_add_(0, dst, m_ref.index_select(0, src))
    out_ref = out_ref / th.bincount(dst, minlength=N).clamp_(min=1).double().unsqueeze(1)
    (g_ref,) = th.autograd.grad(out_ref, m_ref, go.double())
    m_dev = msg.cuda().requires_grad_(True)
    out = ops.BlockMean.apply(m_dev, mask.cuda(), U)
    (g_dev,) = th.autograd.grad(out, m_dev, go.cuda())
    assert float((out.cpu().double() - out_ref).abs().max()) <= 2e-6 * max(1.0, float(out_ref.abs().max()))
    assert float((g_dev.cpu().double() - g_ref).abs().max()) <= 2e-6 * max(1.0, float(g_ref.abs().max()))
    assert float(out[0].abs().max()) == 0.0
    out2 = ops.BlockMean.apply(m_dev, mask.cuda(), U)
    assert th.equal(out, out2)


@pytest.mark.parametrize("U,B,M,comm_p", [(8, 6, 64, 0.6), (3, 5, 16, 0.5), (16, 2, 32, 0.8), (8, 4, 64, 0.0), (1, 7, 8, 1.0)])
def test_block_bitmax_matches_the_edge_list_path(U, B, M, comm_p):
    """DiscreteComm's fused kernel (per-edge hard Gumbel-softmax + OR over the destination's bit mask, one-winner
    gradient) against the edge-list formulation with torch ops (mailbox max in edge-id order)."""
    from uav_bs_ctrl_b200.agents.gnn_agents import _first_max_by_dst
    a, gt, ubs, adj = synth_dense_obs(B, U, 4, "full", seed=5, comm_p=comm_p)
    g = build_obs_graph_batch(a, gt, ubs, adj).to(DEV)["talk"]
    N, E = B * U, g.number_of_edges()
    gen = th.Generator(device=DEV).manual_seed(3)
    logits = th.randn(N, 2 * M, device=DEV, generator=gen, requires_grad=True)
    expo = th.empty(E, M, 2, device=DEV).exponential_(generator=gen)
    w = th.randn(N, 2 * M, device=DEV, generator=gen)
    block, mask = g.block_mask()
    c1 = ops.BlockBitMax.apply(logits, expo, mask, block, 0.5)
    (c1 * w).sum().backward()
    g1, logits.grad = logits.grad.clone(), None
    src, _ = g.edges()
    lg = logits.index_select(0, src).view(-1, M, 2)
    y = ((lg - expo.log()) / 0.5).softmax(-1)
    hard = th.zeros_like(y).scatter_(-1, y.argmax(-1, keepdim=True), 1.0)
    c2 = _first_max_by_dst(g, (hard - y.detach() + y).flatten(1), N)
    (c2 * w).sum().backward()
    assert th.equal(c1, c2), "messages are exactly 0 / 1: the OR must be bit-identical"
    assert_close(g1, logits.grad, rtol=1e-5, atol_scale=1e-6, what="grad of the encoder logits")
    assert float(c1.sum()) > 0
