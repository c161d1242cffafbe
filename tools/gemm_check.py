"""Quick GPU check of the two tcgen05 GEMM kernels on a handful of shapes (progress is flushed line by line, so a hang
is attributable): python tools/gemm_check.py [nn|tn]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from uav_bs_ctrl_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "both"
dev = "cuda"


def say(*a):
    print(*a, flush=True)


def timeit(fn, n=20):
    fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


if which in ("nn", "both"):
    for M, N, K in [(128, 64, 32), (300, 64, 128), (2048, 144, 64), (5000, 64, 288), (104448, 64, 128), (104448, 144, 64),
                    (104448, 64, 288), (104448, 128, 64), (20000, 256, 64)]:
        say("nn", M, N, K, "...")
        x, w = th.randn(M, K, device=dev), th.randn(N, K, device=dev) / K ** 0.5
        out = ops.tc_linear(x, w)
        th.cuda.synchronize()
        ref = x.double() @ w.double().t()
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        us = timeit(lambda: ops.tc_linear(x, w))
        say(f"   rel err {err:.2e}  {us:.1f} us  {(M * K + M * N) * 4 / us / 1e3:.0f} GB/s")
        assert err < 2e-6
if which in ("tn", "both"):
    for R, Mo, No in [(256, 128, 64), (300, 64, 16), (5000, 288, 64), (4099, 9, 64), (20000, 132, 128), (104448, 288, 64),
                      (104448, 192, 64), (104448, 64, 128), (104448, 12, 64)]:
        say("tn", R, Mo, No, "...")
        a, b = th.randn(R, Mo, device=dev), th.randn(R, No, device=dev)
        out = ops.tc_matmul_tn(a, b)
        th.cuda.synchronize()
        ref = a.double().t() @ b.double()
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        us = timeit(lambda: ops.tc_matmul_tn(a, b))
        say(f"   rel err {err:.2e}  {us:.1f} us  {(R * Mo + R * No) * 4 / us / 1e3:.0f} GB/s")
        assert err < 2e-6
if which == "tnbig":                      # one window-sized launch sequence, for ncu
    a, b = th.randn(104448, 288, device=dev), th.randn(104448, 64, device=dev)
    for _ in range(3):
        ops.tc_matmul_tn(a, b)
    th.cuda.synchronize()
say("ok")
