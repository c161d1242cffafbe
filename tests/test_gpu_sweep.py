"""GPU parity of the general-CSR path (BASELINE configs[4], the synthetic hetero-graph sweep): wide inputs go through
library GEMM projections + the gather / online-softmax / aggregate kernel (ubs_gat_aggr_fwd / _bwd)."""
import pytest
import torch as th
import torch.nn as nn

from oracle import gnn_oracle as O
from uav_bs_ctrl_b200 import agents as A, graph as G, ops
from helpers import assert_close, assert_as_accurate

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _graph(n_src, n_dst, deg, seed, sorted_dst=False):
    g = th.Generator().manual_seed(seed)
    E = n_dst * deg
    src = th.randint(0, n_src, (E,), generator=g)
    dst = th.randint(0, n_dst, (E,), generator=g)             # ragged in-degrees (Poisson-like), some zero
    if sorted_dst:
        dst = dst.sort()[0]
    return G.heterograph({("s", "e", "d"): (src, dst)}, num_nodes_dict={"s": n_src, "d": n_dst}), src, dst


@pytest.mark.parametrize("n,deg,H,heads", [(1000, 4, 32, 4), (1000, 16, 64, 4), (4000, 8, 128, 4), (1500, 64, 64, 8),
                                            (700, 16, 256, 4), (512, 4, 256, 1)])
def test_wide_gatv2_matches_oracle(n, deg, H, heads):
    D = H // heads
    th.manual_seed(0)
    ref = O.GATv2Conv((H, H), D, heads, residual=True, allow_zero_in_degree=True, activation=nn.ReLU())
    ref64 = O.GATv2Conv((H, H), D, heads, residual=True, allow_zero_in_degree=True, activation=nn.ReLU()).double()
    ref64.load_state_dict(ref.state_dict())
    mine = A.GATv2Conv((H, H), D, heads, residual=True, allow_zero_in_degree=True, activation=nn.ReLU()).to(DEV)
    mine.load_state_dict(ref.state_dict())
    g, src, dst = _graph(n, n, deg, seed=n + deg)
    gen = th.Generator().manual_seed(1)
    xs, xd = th.randn(n, H, generator=gen) * 0.5, th.randn(n, H, generator=gen) * 0.5
    go = th.randn(n, heads, D, generator=gen)
    outs = {}
    for name, net, dt in (("r32", ref, th.float32), ("r64", ref64, th.float64)):
        a, b = xs.clone().to(dt).requires_grad_(), xd.clone().to(dt).requires_grad_()
        o = net(g["e"], (a, b))
        o.backward(go.to(dt))
        outs[name] = (o, a.grad, b.grad, [p.grad for p in net.parameters()])
    a, b = xs.clone().to(DEV).requires_grad_(), xd.clone().to(DEV).requires_grad_()
    ops.TIMER = ops.KernelTimer()
    o = mine(g.to(DEV)["e"], (a, b))
    o.backward(go.to(DEV))
    used = ops.TIMER.summary()
    ops.TIMER = None
    assert "gat_aggr_fwd" in used and "gat_aggr_bwd" in used
    assert_close(o, outs["r32"][0], rtol=3e-5, atol_scale=1e-5, what="out")
    assert_as_accurate(o, outs["r32"][0], outs["r64"][0], what="out vs fp64", slack=4.0, floor_scale=3e-6)
    assert_as_accurate(a.grad, outs["r32"][1], outs["r64"][1], what="grad x_src", slack=6.0, floor_scale=5e-6)
    assert_as_accurate(b.grad, outs["r32"][2], outs["r64"][2], what="grad x_dst", slack=6.0, floor_scale=5e-6)
    for (k, p), g32, g64 in zip(mine.named_parameters(), outs["r32"][3], outs["r64"][3]):
        assert_as_accurate(p.grad, g32, g64, what=f"grad {k}", slack=6.0, floor_scale=5e-6)


def test_wide_gatv2_star_layout_and_no_residual():
    H, heads = 64, 4
    th.manual_seed(3)
    ref = O.GATv2Conv((H, H), H // heads, heads, residual=False, allow_zero_in_degree=True)
    mine = A.GATv2Conv((H, H), H // heads, heads, residual=False, allow_zero_in_degree=True).to(DEV)
    mine.load_state_dict(ref.state_dict())
    deg = th.tensor([3, 0, 40, 1, 7, 33])
    E = int(deg.sum())
    g = G.heterograph({("s", "e", "d"): (th.arange(E), th.repeat_interleave(th.arange(6), deg))},
                      num_nodes_dict={"s": E, "d": 6})
    assert g["e"].csr().is_star
    xs, xd = th.randn(E, H), th.randn(6, H)
    a, ar = xs.to(DEV).requires_grad_(), xs.clone().requires_grad_()
    o = mine(g.to(DEV)["e"], (a, xd.to(DEV)))
    orf = ref(g["e"], (ar, xd))
    (o ** 2).sum().backward()
    (orf ** 2).sum().backward()
    assert_close(o, orf, rtol=3e-5, atol_scale=1e-5, what="out")
    assert float(o[1].abs().max()) == 0.0                      # zero in-degree, no residual, no activation => zeros
    assert_close(a.grad, ar.grad, rtol=1e-4, atol_scale=1e-5, what="grad x_src (star: stores, no atomics)")
