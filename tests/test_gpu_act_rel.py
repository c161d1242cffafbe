"""The act vector-step as ONE kernel (``ubs_agent_act_rel_fwd``: both GATv2 observation relations staged by bulk async
copies + aggregator + TarMAC / GRU + Q head + epsilon-greedy) against the CPU oracle and against the three-launch path."""
from types import SimpleNamespace as SN

import pytest
import torch as th

from oracle import gnn_oracle as O
from uav_bs_ctrl_b200 import agents as A, ops
from uav_bs_ctrl_b200.arena import PacketLayout, ObsPacket, SequenceArena
from uav_bs_ctrl_b200.synth import synth_dense_obs
from helpers import make_args, assert_close, assert_as_accurate

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    # c,       H,  heads, U, G,  B,   profile,     comm_p, near_p, F_gt
    ("tarmac", 64, 4, 8, 80, 256, "full", 1.0, 1.0, 4),          # exp3 at full size (128 CTAs)
    ("tarmac", 64, 4, 8, 80, 37, "realistic", 0.6, 0.7, 4),      # ragged degrees incl. 0, odd near offsets, partial last tile
    ("tarmac", 64, 4, 8, 50, 9, "random", 1.0, 0.3, 4),          # map '8ubs' (50 GTs), mean degree ~1
    ("tarmac", 64, 4, 3, 10, 7, "realistic", 0.6, 0.5, 4),       # 3 agents: 5 envs per 15-row tile
    ("tarmac", 64, 4, 4, 20, 11, "realistic", 1.0, 1.0, 3),      # fair_service=False: 12-byte source rows (unaligned starts)
    ("tarmac", 32, 4, 4, 10, 9, "realistic", 1.0, 1.0, 4),       # H = 32
    ("tarmac", 64, 8, 8, 30, 5, "realistic", 0.8, 1.0, 4),       # 8 heads
    ("tarmac", 64, 2, 8, 30, 5, "full", 0.8, 1.0, 4),            # 2 heads
    (None, 64, 4, 8, 30, 5, "realistic", 1.0, 1.0, 4),           # independent agents (plain GRU)
    ("tarmac", 64, 4, 16, 40, 3, "realistic", 0.7, 0.8, 4),      # 16 agents per env: one env per tile
]


@pytest.mark.parametrize("c,H,heads,U,G,B,profile,comm_p,near_p,F_gt", CASES)
def test_one_kernel_act_step_matches_oracle_and_three_launch_path(c, H, heads, U, G, B, profile, comm_p, near_p, F_gt):
    shape = {'agent': 2, 'ubs': 2, 'gt': F_gt}
    args = make_args(c=c, hidden_size=H, n_heads=heads)
    th.manual_seed(2)
    ref = O.GnnAgent(shape, 9, args)
    net = A.GnnAgent(shape, 9, args).to(DEV)
    net.load_state_dict(ref.state_dict())
    ref64 = O.GnnAgent(shape, 9, args).double()
    ref64.load_state_dict(ref.state_dict())
    T = 3
    L = PacketLayout(B, U, G, F_gt=F_gt)
    ar = SequenceArena(L, T, H, DEV)
    pk = []
    for t in range(T):
        a, gt, ubs, adj = synth_dense_obs(B, U, G, profile, seed=40 + t, comm_p=comm_p, near_p=near_p, F_gt=F_gt)
        pk.append(ObsPacket(L).fill_from_dense(a, gt, ubs, adj if c else None))
        ar.load(t, pk[-1])
    assert net.rel_act_supported(ar), "this configuration must take the one-kernel path"
    h0 = th.randn(B * U, H, generator=th.Generator().manual_seed(3)) * 0.3
    # oracle (fp32 and fp64)
    g64 = lambda g: g._map(lambda x: x.double() if x.is_floating_point() else x, lambda r: r)
    h_r, h_6, q_r, q_6 = h0, h0.double(), [], []
    with th.no_grad():
        for t in range(T):
            q, h_r = ref(pk[t].to_graph(), h_r)
            q_r.append(q)
            q, h_6 = ref64(g64(pk[t].to_graph()), h_6)
            q_6.append(q)
    outs = {}
    for fused in (True, False):
        net.use_rel_act = fused
        ar.h[0].copy_(h0)
        ops.TIMER = ops.KernelTimer()
        qs = [net.arena_step(ar, t).clone() for t in range(T)]
        used = ops.TIMER.summary()
        ops.TIMER = None
        if fused:
            assert used.get("agent_act_rel", {}).get("count") == T and "gatv2_fwd" not in used, "ONE kernel per step"
        else:
            assert used.get("gatv2_fwd", {}).get("count") == 2 * T
        outs[fused] = (th.stack(qs), ar.h[1:T + 1].clone(), ar.acts[:T].clone())
    q_f, h_f, a_f = outs[True]
    assert_close(q_f, th.stack(q_r), rtol=1e-5, atol_scale=1e-6, what="q vs oracle fp32")
    assert_close(h_f[-1], h_r, rtol=1e-5, atol_scale=1e-6, what="h vs oracle fp32")
    assert_as_accurate(q_f, th.stack(q_r), th.stack(q_6), what="q vs oracle fp64", slack=4.0, floor_scale=2e-6)
    assert_as_accurate(h_f[-1], h_r, h_6, what="h vs oracle fp64", slack=4.0, floor_scale=2e-6)
    assert_close(q_f, outs[False][0], rtol=1e-5, atol_scale=1e-6, what="one kernel vs three launches: q")
    assert_close(h_f, outs[False][1], rtol=1e-5, atol_scale=1e-6, what="one kernel vs three launches: h")
    assert th.equal(a_f, q_f.argmax(2)), "greedy actions are the argmax of the Q values the kernel wrote"


def test_epsilon_greedy_and_weight_refresh_in_the_one_kernel_step():
    """Fused epsilon-greedy (one uniform per env) and the in-place refresh of the packed relation tables after a
    parameter update (captured graphs read them by address)."""
    B, U, G, H = 6, 8, 20, 64
    args = make_args()
    th.manual_seed(5)
    net = A.GnnAgent({'agent': 2, 'ubs': 2, 'gt': 4}, 9, args).to(DEV)
    L = PacketLayout(B, U, G)
    ar = SequenceArena(L, 1, H, DEV)
    a, gt, ubs, adj = synth_dense_obs(B, U, G, "realistic", seed=9, comm_p=0.7)
    ar.load(0, ObsPacket(L).fill_from_dense(a, gt, ubs, adj))
    u = th.rand(B, 1, device=DEV).expand(B, U).reshape(-1).contiguous()
    rnd = th.randint(0, 9, (B * U,), device=DEV)
    eps = th.tensor(0.5, device=DEV)
    q = net.arena_step(ar, 0, explore=(u, rnd, eps))
    want = th.where(u <= eps, rnd, q.argmax(1))
    assert th.equal(ar.acts[0], want)
    tab0, ptr0 = net._relpack_cache[1].clone(), net._relpack_cache[1].data_ptr()
    with th.no_grad():
        for p in net.parameters():
            p.mul_(1.1)
    net.refresh_packed()
    assert net._relpack_cache[1].data_ptr() == ptr0, "tables are rebuilt in place"
    assert not th.equal(net._relpack_cache[1], tab0)
    q2 = net.arena_step(ar, 0)
    net.use_rel_act = False
    ar_h = ar.h[1].clone()
    q3 = net.arena_step(ar, 0)
    assert_close(q2, q3, rtol=1e-5, atol_scale=1e-6, what="after a parameter update")
    assert_close(ar_h, ar.h[1], rtol=1e-5, atol_scale=1e-6, what="h after a parameter update")
