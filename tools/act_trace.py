#!/usr/bin/env python
"""clock64() phase trace of one CTA of the one-kernel act step (debug build: make EXTRA=-DUBS_ACT_TRACE)."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th
from types import SimpleNamespace
from uav_bs_ctrl_b200 import agents as A, _lib
from uav_bs_ctrl_b200.arena import PacketLayout, ObsPacket, SequenceArena
from uav_bs_ctrl_b200.synth import synth_dense_obs

args = SimpleNamespace(hidden_size=64, n_layers=1, n_heads=4, msg_size=64, key_size=16, n_rounds=1, c="tarmac", o="gnn", dueling=False)
th.manual_seed(0)
net = A.GnnAgent({"agent": 2, "ubs": 2, "gt": 4}, 9, args).to("cuda")
B, U, G = 256, 8, 80
L = PacketLayout(B, U, G)
ar = SequenceArena(L, 2, 64, "cuda")
ar.load(0, ObsPacket(L).fill_from_dense(*synth_dense_obs(B, U, G, "full", seed=1)))
for _ in range(5):
    net.arena_step(ar, 0)
th.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * 16)()
lib.ubs_act_trace_read.argtypes = [C.c_void_p]
assert lib.ubs_act_trace_read(buf) == 0
t = list(buf)
names = ["start", "relations", "aggr", "vsq", "attention", "gi", "gh", "gates", "q head", "stores"]
for i in range(1, 10):
    print(f"{names[i]:10s} {t[i] - t[i - 1]:6d} cycles   (cum {t[i] - t[0]})")
