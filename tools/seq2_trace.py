#!/usr/bin/env python
"""Per-phase clock64() trace of the tensor-core window kernel (debug build: make EXTRA=-DUBS_SEQ2_TRACE).
Prints, for block 0 and steps 1..6, when every warp passed each trace point (cycles since the step's first stamp)."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th
from types import SimpleNamespace
from uav_bs_ctrl_b200 import agents as A, ops, _lib

dev = "cuda"
args = SimpleNamespace(hidden_size=64, n_layers=1, n_heads=4, msg_size=64, key_size=16, n_rounds=1, c="tarmac", o="gnn", dueling=False)
th.manual_seed(0)
net = A.GnnAgent({"agent": 2, "ubs": 2, "gt": 4}, 9, args).to(dev)
T, N = 51, 2048
dims = net.fused_dims(8)
xin = th.randn(T, N, 128, device=dev)
h0 = th.zeros(N, 64, device=dev)
mask = th.full((T, N), 255, dtype=th.int32, device=dev)
for training in (False, True):
    for _ in range(2):
        if training:
            q, h_last, _ = ops.AgentSequence2.apply(xin, h0, mask, dims, *[net._fused_params()[k] for k in ops.PARAM_ORDER])
        else:
            with th.no_grad():
                ops.agent_seq2_infer(dims, net._fused_params(), xin, h0, mask)
    th.cuda.synchronize()
    buf = (C.c_longlong * (8 * 16 * 8))()
    lib = _lib.load()
    lib.ubs_seq2_trace_read.argtypes = [C.c_void_p]
    assert lib.ubs_seq2_trace_read(buf) == 0
    t = th.tensor(list(buf)).view(8, 16, 8)
    print("training" if training else "inference")
    for step in (2, 3):
        base = int(t[step, :, 0].min())
        print(" step", step, "(cycles since first warp entered the step; rows = warps 0..15, cols = trace points 0..7)")
        for w in range(16):
            print("  w%02d" % w, " ".join("%6d" % (int(v) - base) if v else "     -" for v in t[step, w]))
        print("  step length:", int(t[step + 1, :, 0].min()) - base)
