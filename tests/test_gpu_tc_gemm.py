"""GPU parity of the 3xTF32 tcgen05 projection kernel (ubs_tf32x3_gemm) against fp64 / fp32 matmul."""
import pytest
import torch as th

from uav_bs_ctrl_b200 import ops
from helpers import assert_as_accurate

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("M,N,K,relu,bias", [
    (128, 64, 32, False, False),          # one tile, one K chunk
    (300, 64, 128, True, True),           # aggregator: ragged last tile
    (2048, 144, 64, False, True),         # half of the fused [pv | pg] projection (N = 288 = 2 x 144)
    (5000, 64, 288, False, False),        # dx = [dgi | dv] W_dx
    (1000, 128, 64, False, False),        # d_xin = dpre W_aggr
    (104448, 64, 128, True, True),        # exp3 window: T*N rows
    (20000, 256, 64, False, True),        # widest N, single-buffered accumulators
])
def test_tf32x3_gemm_is_fp32_accurate(M, N, K, relu, bias):
    g = th.Generator().manual_seed(M + N + K)
    x = th.randn(M, K, generator=g)
    w = th.randn(N, K, generator=g) / K ** 0.5
    b = th.randn(N, generator=g) if bias else None
    ref64 = x.double() @ w.double().t() + (b.double() if bias else 0)
    ref32 = x @ w.t() + (b if bias else 0)
    if relu:
        ref64, ref32 = ref64.relu(), ref32.relu()
    out = ops.tc_linear(x.to(DEV), w.to(DEV), None if b is None else b.to(DEV), relu=relu)
    th.cuda.synchronize()
    # as accurate against fp64 as an fp32 matmul (x4), i.e. far inside the 1e-5 bar; single-pass TF32 would be ~1e-3
    assert_as_accurate(out, ref32, ref64, what="3xTF32 gemm", slack=4.0, floor_scale=2e-6)


def test_tf32x3_gemm_strided_views():
    g = th.Generator().manual_seed(0)
    big = th.randn(700, 480, generator=g).to(DEV)
    x = big[:, 192:480]                      # row-strided view, K = 288
    w = th.randn(64, 288, generator=g).to(DEV)
    out = th.zeros(700, 352, device=DEV)
    ops.tc_linear(x, w, out=out[:, 64:128])
    ref = x.double() @ w.double().t()
    assert float((out[:, 64:128].double() - ref).abs().max() / ref.abs().max()) < 2e-6
    assert float(out[:, :64].abs().max()) == 0 and float(out[:, 128:].abs().max()) == 0


@pytest.mark.parametrize("R,Mo,No", [
    (256, 128, 64),            # one slice, one tile
    (300, 64, 16),             # ragged slice, Mo < one tile, narrowest No
    (5000, 288, 64),           # [dW_ih_x; dW_vsq_x]: three Mo tiles, the last one ragged (32 rows)
    (4099, 9, 64),             # dW_out: Mo = 9 (unaligned rows -> scalar loads)
    (104448, 192, 64),         # exp3 window: dW_ih[:, H:]
    (104448, 64, 128),         # exp3 window: dW_aggr
    (20000, 130, 128),         # widest No, ragged Mo
])
def test_tf32x3_gemm_tn_is_fp32_accurate(R, Mo, No):
    g = th.Generator().manual_seed(R + Mo + No)
    a = th.randn(R, Mo, generator=g)
    b = th.randn(R, No, generator=g)
    ref64 = a.double().t() @ b.double()
    ref32 = a.t() @ b
    out = ops.tc_matmul_tn(a.to(DEV), b.to(DEV))
    th.cuda.synchronize()
    assert_as_accurate(out, ref32, ref64, what="3xTF32 gemm_tn", slack=4.0, floor_scale=2e-6)
    out2 = ops.tc_matmul_tn(a.to(DEV), b.to(DEV))
    assert th.equal(out, out2)                                   # fixed-order reduction of the slice partials


def test_tf32x3_gemm_tn_strided_views():
    g = th.Generator().manual_seed(1)
    S = th.randn(6000, 480, generator=g).to(DEV)
    x = th.randn(6000, 64, generator=g).to(DEV)
    a = S[:, 192:480]                        # [dvsq | dgh] column block of the stash
    out = ops.tc_matmul_tn(a, x)
    ref = a.double().t() @ x.double()
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 2e-6


def test_matmul_tn_pads_rows_that_are_not_16_byte_granular():
    """dW_out: the Q head has n_actions = 9 columns; ``matmul_tn`` pads them to 12 so that the long reduction stays on the
    kernel's vector producer path, and cuts the result back."""
    g = th.Generator().manual_seed(5)
    a = th.randn(8192, 9, generator=g)
    b = th.randn(8192, 64, generator=g)
    ops.TIMER = ops.KernelTimer()
    out = ops.matmul_tn(a.to(DEV), b.to(DEV))
    used = ops.TIMER.summary()
    ops.TIMER = None
    assert "tf32x3_gemm_tn" in used and out.shape == (9, 64)
    assert_as_accurate(out, a.t() @ b, a.double().t() @ b.double(), what="padded gemm_tn", slack=4.0, floor_scale=2e-6)


@pytest.mark.parametrize("R,C,ld", [(1, 4, 4), (300, 64, 64), (5000, 480, 480), (104448, 480, 480), (4097, 96, 352), (70000, 1024, 1024)])
def test_colsum_matches_fp64(R, C, ld):
    g = th.Generator().manual_seed(R + C)
    big = th.randn(R, ld, generator=g).to(DEV)
    x = big[:, ld - C:]                                   # row-strided view
    ops.TIMER = ops.KernelTimer()
    out = ops.colsum(x)
    used = ops.TIMER.summary()
    ops.TIMER = None
    assert "colsum" in used
    ref64 = x.double().sum(0)
    assert_as_accurate(out, x.sum(0), ref64, what="colsum", slack=4.0, floor_scale=2e-6)
    assert th.equal(out, ops.colsum(x))                   # fixed summation order


def test_colsum_falls_back_for_rows_that_are_not_16_byte_granular():
    x = th.randn(1000, 9, device=DEV)
    assert th.allclose(ops.colsum(x), x.sum(0), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("R,C", [(7, 64), (104448, 64), (5001, 128)])
def test_relu_bwd_colsum_matches_torch(R, C):
    g = th.Generator().manual_seed(R)
    dy, y = th.randn(R, C, generator=g).to(DEV), th.randn(R, C, generator=g).relu().to(DEV)
    ref = dy * (y > 0)
    ref64 = ref.double().sum(0)
    dx, cs = ops.relu_bwd_colsum_(dy.clone(), y)
    assert th.equal(dx, ref)
    assert_as_accurate(cs, ref.sum(0), ref64, what="relu_bwd colsum", slack=4.0, floor_scale=2e-6)
