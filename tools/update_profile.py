#!/usr/bin/env python
"""torch.profiler table of ONE update_arena (exp3 shapes): which kernels — ours and the library's — the update spends its time in."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th
import bench as Bn
from uav_bs_ctrl_b200.learner import MultiAgentQLearner

dev = th.device("cuda:0"); B, T = 256, 50
th.manual_seed(0)
learner = MultiAgentQLearner(dict(obs_shape=Bn.OBS_SHAPE, state_shape=None, n_actions=9, n_agents=8, episode_limit=T), Bn.model_args(dev, T, B))
L, packets = Bn.make_packets(B, T, "full", 1234, pin=False)
arena = learner.new_arena(80)
for t in range(T + 1):
    arena.load(t, packets[t])
learner.begin_sequence(arena)
for t in range(T):
    learner.act_arena(arena, t, 0.05)
for _ in range(3):
    learner.update_arena(arena, sync=False)
th.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        learner.update_arena(arena, sync=False)
    th.cuda.synchronize()
rows = []
for e in prof.key_averages():
    if e.device_time_total > 0 and e.device_type.name == "CUDA":
        rows.append((e.device_time_total / 3, e.count // 3, e.key[:90]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time per update: {tot:.0f} us")
for t_, c_, k_ in rows[:45]:
    print(f"{t_:9.1f} us {c_:4d}x  {k_}")
