"""Host logic of the packed observation packets / sequence arena (CPU) and their GPU parity with the graph path."""
from types import SimpleNamespace

import pytest
import torch as th

from uav_bs_ctrl_b200.arena import PacketLayout, ObsPacket, SequenceArena, packet_graph
from uav_bs_ctrl_b200.builder import build_obs_graph_batch
from uav_bs_ctrl_b200.synth import synth_dense_obs
from helpers import assert_close

SHAPE = {'agent': 2, 'ubs': 2, 'gt': 4}


def test_layout_sections_are_aligned_and_disjoint():
    L = PacketLayout(5, 3, 7)
    ends = 0
    for name in ("x_agent", "ip_seen", "ip_near", "mask", "rew", "done", "bad", "x_ubs", "x_gt"):
        assert L.off[name] % 32 == 0 and L.off[name] >= ends          # 128-byte lines
        ends = L.off[name] + L.size[name]
    assert L.words % 32 == 0 and L.words >= ends
    assert L.size["x_gt"] == 5 * 3 * 7 * 4 and L.size["ip_seen"] == 16
    # the `seen` rows are the last section: a packet's live words are one prefix
    assert L.off["x_gt"] + L.size["x_gt"] == ends and L.used_words(0) <= L.off["x_gt"] + 3
    assert L.used_words(5 * 3 * 7) == ends <= L.words and L.used_words(10) == (L.off["x_gt"] + 40 + 3) // 4 * 4


def test_compact_load_ships_the_live_prefix_only():
    B, U, G, T = 4, 5, 6, 2
    L = PacketLayout(B, U, G)
    ar = SequenceArena(L, T, 8, "cpu")
    ar.buf.fill_(-7)
    pk = ObsPacket(L).fill_from_dense(*synth_dense_obs(B, U, G, "realistic", seed=3, comm_p=0.5, near_p=0.6))
    n_seen = int(pk.sec("ip_seen")[-1])
    assert 0 < n_seen < B * U * G
    nbytes = ar.load(1, pk)
    assert nbytes == 4 * L.used_words(n_seen) < 4 * L.words
    assert th.equal(ar.buf[1, :nbytes // 4], pk.buf[:nbytes // 4]) and bool((ar.buf[1, nbytes // 4:] == -7).all())
    g, ref = ar.graph(1), pk.to_graph()                         # the slot reads back as the same graph
    assert th.equal(g.nodes["gt"].data["feat"], ref.nodes["gt"].data["feat"])
    assert th.equal(g.nodes["ubs"].data["feat"], ref.nodes["ubs"].data["feat"])
    # refilling the packet with a denser observation re-sizes the prefix
    pk.fill_from_dense(*synth_dense_obs(B, U, G, "full", seed=3))
    assert ar.load(0, pk) == 4 * L.used_words(B * U * G) and ar.load(0, pk, compact=False) == 4 * L.words


@pytest.mark.parametrize("profile,comm_p", [("full", 1.0), ("realistic", 0.5), ("random", 0.3)])
def test_packet_graph_equals_builder_graph(profile, comm_p):
    B, U, G = 4, 5, 6
    a, gt, ubs, adj = synth_dense_obs(B, U, G, profile, seed=2, comm_p=comm_p, near_p=0.6)
    ref = build_obs_graph_batch(a, gt, ubs, adj)
    pk = ObsPacket(PacketLayout(B, U, G)).fill_from_dense(a, gt, ubs, adj, rew=th.ones(B, U), done=th.zeros(B))
    g = pk.to_graph()
    for nt in ("agent", "gt", "ubs"):
        assert th.equal(g.nodes[nt].data["feat"], ref.nodes[nt].data["feat"])
    for et in ("seen", "near", "talk"):
        assert th.equal(g[et].csr().indptr, ref[et].csr().indptr)
        assert th.equal(g[et].csr().src_of_slot(), ref[et].csr().src_of_slot())
        (u1, v1), (u2, v2) = g.edges(et), ref.edges(et)
        assert th.equal(u1, u2) and th.equal(v1, v2), et
    assert th.equal(g["talk"].block_mask()[1], ref["talk"].block_mask()[1])
    assert g.batch_size == B and g.uniform_block("agent") == U


def test_arena_load_and_reward_bookkeeping():
    B, U, G, T = 3, 2, 4, 3
    L = PacketLayout(B, U, G)
    ar = SequenceArena(L, T + 1, hidden=8, device="cpu")
    for t in range(T + 1):
        a, gt, ubs, adj = synth_dense_obs(B, U, G, "realistic", seed=t)
        done = th.tensor([0., 1., 0.]) if t == 2 else th.zeros(B)
        bad = th.tensor([0., 1., 0.]) if t == 2 else th.zeros(B)
        ar.load(t, ObsPacket(L).fill_from_dense(a, gt, ubs, adj, rew=th.full((B, U), float(t)), done=done, bad=bad))
    assert ar.rewards(T).shape == (T, B, U) and float(ar.rewards(T)[1].mean()) == 2.0     # reward of transition 1 sits in slot 2
    assert ar.dones(T).shape == (T, B, 1) and float(ar.dones(T).sum()) == 0.0                # done muted by bad_mask
    g = ar.graph(1)
    a, gt, ubs, adj = synth_dense_obs(B, U, G, "realistic", seed=1)
    assert th.equal(g.nodes["gt"].data["feat"], build_obs_graph_batch(a, gt, ubs, adj).nodes["gt"].data["feat"])


# ---------------------------------------------------------------------------------------------------------------- GPU
def _learner(B, U, T, fused=True, graphs=False, seed=0):
    from uav_bs_ctrl_b200.learner import MultiAgentQLearner
    th.manual_seed(seed)
    a = SimpleNamespace(device="cuda", o="gnn", c="tarmac", share_reward=False, hidden_size=64, n_layers=1, n_heads=4,
                        msg_size=64, key_size=16, n_rounds=1, lr=1e-3, gamma=0.99, polyak=0.9, batch_size=1,
                        replay_size=2, max_seq_len=T, anneal_lr=False, double_q=True, dueling=False, mixer=False,
                        n_envs=B, fused=fused, cuda_graphs=graphs)
    return MultiAgentQLearner(dict(obs_shape=SHAPE, state_shape=None, n_actions=9, n_agents=U, episode_limit=T), a)


def _episode(B, U, G, T, seed=50):
    L = PacketLayout(B, U, G)
    gen = th.Generator().manual_seed(seed)
    pk = []
    for t in range(T + 1):
        a, gt, ubs, adj = synth_dense_obs(B, U, G, "realistic", seed=seed + t, comm_p=0.7)
        done = th.zeros(B)
        if t == T:
            done[:] = 1.0
        pk.append(ObsPacket(L).fill_from_dense(a, gt, ubs, adj, rew=th.rand(B, U, generator=gen), done=done, bad=done))
    return L, pk


@pytest.mark.gpu
def test_arena_step_and_sequence_match_graph_path():
    B, U, G, T = 6, 8, 30, 4
    Lr = _learner(B, U, T)
    layout, pk = _episode(B, U, G, T)
    ar = Lr.new_arena(G)
    for t in range(T + 1):
        ar.load(t, pk[t])
    net = Lr.policy_net
    graphs = [p.to_graph().to("cuda") for p in pk]
    h = th.randn(B * U, 64, device="cuda") * 0.2
    ar.h[0].copy_(h)
    with th.no_grad():
        for t in range(T + 1):
            q_a = net.arena_step(ar, t)
            q_g, h = net(graphs[t], h)
            # same kernels, but the lane-group width of the relation kernel is picked from the (capacity vs actual)
            # edge-count hint, so sums may associate differently: tight tolerance instead of bitwise equality
            assert_close(q_a, q_g, rtol=1e-5, atol_scale=1e-6, what=f"q[{t}]")
            assert_close(ar.h[t + 1], h, rtol=1e-5, atol_scale=1e-6, what=f"h[{t}]")
            assert th.equal(ar.acts[t], q_a.argmax(1))
            h = ar.h[t + 1].clone()
    # training: arena sequence vs list-of-graphs sequence
    h0 = ar.h[0].clone()
    q1, _ = net.arena_sequence(ar, 0, T + 1, h0)
    (q1 ** 2).mean().backward()
    g1 = [p.grad.clone() for p in net.parameters()]
    for p in net.parameters():
        p.grad = None
    q2, _ = net.forward_sequence(graphs, h0)
    (q2 ** 2).mean().backward()
    assert_close(q1, q2, rtol=1e-5, atol_scale=1e-6, what="q")
    gmax = max(float(p.grad.abs().max()) for p in net.parameters())
    for a, p in zip(g1, net.parameters()):
        assert float((a - p.grad).abs().max()) <= 1e-5 * float(p.grad.abs().max()) + 1e-7 * gmax, "param grad"


@pytest.mark.gpu
@pytest.mark.parametrize("use_graphs", [False, True])
def test_learner_arena_cycle_matches_graph_object_cycle(use_graphs):
    B, U, G, T = 5, 8, 20, 4
    layout, pk = _episode(B, U, G, T, seed=70)
    # (1) graph-object API (reference-shaped act / cache / update)
    L1 = _learner(B, U, T, graphs=False, seed=3)
    graphs = [p.to_graph().to("cuda") for p in pk]
    h = L1.init_hidden(B).to("cuda")
    acts_1 = []
    for t in range(T):
        acts, h2 = L1.act(graphs[t], h, 0.0)                          # greedy: no RNG in the comparison
        rew, done = pk[t + 1].sec("rew").view(B, U).cuda(), pk[t + 1].sec("done").cuda()
        L1.cache(graphs[t], h, None, acts, rew, graphs[t + 1], h2, None, done, pk[t + 1].sec("bad").cuda())
        acts_1.append(acts)
        h = h2
    out1 = L1.update(samples=[L1.buffer.memory[-1]])
    g1 = L1.grad_bucket.flat.clone()
    # (2) arena API, two consecutive cycles (the second one replays captured graphs when enabled)
    L2 = _learner(B, U, T, graphs=use_graphs, seed=3)
    ar = L2.new_arena(G)
    for cycle in range(2):
        L2.begin_sequence(ar)
        for t in range(T + 1):
            ar.load(t, pk[t])
        for t in range(T):
            acts = L2.act_arena(ar, t, 0.0)
            if cycle == 0:
                assert float((acts != acts_1[t]).float().mean()) < 0.01      # argmax ties / 1e-6 differences only
        if cycle == 0:
            out2 = L2.update_arena(ar)
            g2 = L2.grad_bucket.flat.clone()
    assert abs(out1["LossQ"] - out2["LossQ"]) <= 1e-5 * abs(out1["LossQ"])
    assert_close(g2, g1, rtol=1e-4, atol_scale=1e-5, what="policy gradients")
    # after the update the act path must see the new weights (packed buffer refreshed, graphs still valid)
    with th.no_grad():
        q_ref, _ = L2.policy_net(graphs[0], th.zeros(B * U, 64, device="cuda"))
    assert float((ar.acts[0] != q_ref.argmax(1)).float().mean()) < 0.01


def test_replay_ring_keeps_the_order_of_the_reference_deque():
    """``ArenaReplay`` mirrors ``ReplayBuffer`` (reference ``algos/madrqn/buffer.py:7-42``): capacity-bounded, oldest window
    overwritten first, ``sample`` draws distinct stored windows."""
    import random
    from uav_bs_ctrl_b200.arena import ArenaReplay
    L = PacketLayout(2, 2, 3)
    made = []

    def make():
        made.append(SequenceArena(L, 3, 4, "cpu"))
        return made[-1]
    rp = ArenaReplay(make, capacity=3)
    assert len(rp) == 0 and rp.windows() == []
    for k in range(5):
        ar = rp.next_arena()
        ar.acts.fill_(k)                                   # tag the window
        rp.commit()
        assert len(rp) == min(k + 1, 3)
    assert len(made) == 3, "arenas are allocated once and recycled"
    assert [int(a.acts[0, 0]) for a in rp.windows()] == [2, 3, 4], "oldest first, like deque(maxlen=capacity)"
    random.seed(0)
    picked = rp.sample(2)
    assert len(picked) == 2 and picked[0] is not picked[1] and all(p in rp.windows() for p in picked)
    with pytest.raises(ValueError):
        rp.sample(4)
    assert rp.nbytes() == 3 * (made[0].buf.numel() * 4 + made[0].h.numel() * 4 + made[0].acts.numel() * 8)


@pytest.mark.gpu
def test_stream_overlap_and_programmatic_launch_do_not_change_results():
    """The update's second stream (target window), the forked relation streams and the programmatic launch of the act
    steps inside the rollout graph only change WHEN kernels run: losses, gradients, hidden states and actions are
    bit-identical to the in-order execution."""
    from uav_bs_ctrl_b200 import ops
    B, U, G, T = 256, 8, 40, 16                       # 2048 rows x 17 slots: above the size gates of both overlaps
    layout, pk = _episode(B, U, G, T, seed=90)
    results = []
    for overlap, fork, pdl in ((False, False, False), (True, True, True)):
        L = _learner(B, U, T, graphs=True, seed=5)
        L.args.overlap_target, L.args.act_pdl = overlap, pdl
        old_fork, ops.FORK_RELATIONS = ops.FORK_RELATIONS, fork
        try:
            ar = L.new_arena(G)
            outs = []
            for cycle in range(2):                        # the second cycle replays the captured rollout graph
                L.begin_sequence(ar)
                for t in range(T + 1):
                    ar.load(t, pk[t])
                L.rollout_arena(None, ar, 0.0)
                out = L.update_arena(ar, sync=False)
                outs.append((ar.h.clone(), ar.acts.clone(), out["LossQ"].clone(), L.grad_bucket.flat.clone()))
            th.cuda.synchronize()
            results.append(outs)
        finally:
            ops.FORK_RELATIONS = old_fork
    for a, b in zip(*results):
        for x, y, what in zip(a, b, ("hidden states", "actions", "loss", "gradients")):
            assert th.equal(x, y), what


@pytest.mark.gpu
def test_graphed_update_follows_the_parameters_over_several_optimizer_steps():
    """The update graph rebuilds every parameter-derived buffer (window weight layouts, packed weights) INSIDE the graph:
    three consecutive updates through the replayed graph give the losses and parameters of three eager updates."""
    B, U, G, T = 64, 8, 80, 6
    layout, pk = _episode(B, U, G, T, seed=95)
    runs = []
    for graphed in (False, True):
        L = _learner(B, U, T, graphs=True, seed=7)
        L.args.update_graph = graphed
        ar = L.new_arena(G)
        L.begin_sequence(ar)
        for t in range(T + 1):
            ar.load(t, pk[t])
        L.rollout_arena(None, ar, 0.0)
        losses = [float(L.update_arena(ar, sync=True)["LossQ"]) for _ in range(3)]
        runs.append((losses, [p.detach().clone() for p in L.policy_net.parameters()],
                     [p.detach().clone() for p in L.target_net.parameters()]))
    (l0, p0, t0), (l1, p1, t1) = runs
    assert l0[0] != l0[1] != l0[2], "the parameters must move between the updates"
    assert l0 == l1, (l0, l1)
    for a, b in zip(p0 + t0, p1 + t1):
        assert th.equal(a, b)
