#!/usr/bin/env python
"""Times the device env step (ubs_env_step = step kernel + pack kernel) at the BASELINE shape (256 envs, 8 UBS x 80 GT)
in two regimes: `spread` (RNG-matched resets: UBSs far from the hot spot, almost nothing to schedule) and `hover`
(every UBS on top of the hot spot: full RB contention, the scheduler's worst case).

    python tools/env_profile.py                      # JSON with per-step microseconds (CUDA events, eager + graph)
    ncu --set full -k regex:env_ ... python tools/env_profile.py --ncu
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch as th  # noqa: E402

from uav_bs_ctrl_b200 import envs as E  # noqa: E402
from uav_bs_ctrl_b200.arena import SequenceArena  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--map", default="8ubs80")
    a = ap.parse_args()
    B, T = a.envs, 50
    env = E.MultiUbsCoverageVecEnv(a.map, B)
    m = env.map
    arena = SequenceArena(env.new_layout(), T + 1, 64, env.device)
    pu, pg, pr = E.sample_layouts(m, range(B))
    rng = np.random.RandomState(0)
    idx = rng.randint(0, m.n_gts, size=(B, m.n_ubs))
    near = np.take_along_axis(pg.astype(np.float64), idx[..., None].repeat(2, -1), 1) + rng.uniform(-90, 90, (B, m.n_ubs, 2))
    out = {}
    for regime, pos in (("spread", pu), ("hover", np.clip(near, 0, m.range_pos))):
        env.reset(arena, 0, layouts=(pos, pg, pr))
        arena.acts.zero_()                                   # action 0 = stay: the regime persists over the steps
        th.cuda.synchronize()
        if a.ncu:
            for t in range(3):
                env.step(arena, t)
            th.cuda.synchronize()
            continue
        for t in range(5):
            env.step(arena, t)
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(5, T):
            env.step(arena, t)
        e1.record()
        th.cuda.synchronize()
        eager = 1e3 * e0.elapsed_time(e1) / (T - 5)
        g = th.cuda.CUDAGraph()
        with th.cuda.graph(g):
            for t in range(T):
                env.step(arena, t)
        g.replay()
        th.cuda.synchronize()
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        th.cuda.synchronize()
        # per-phase cycles of CTA 0 (UBS_ENV_PROFILE=1 makes the step kernel stamp clock64() after every barrier)
        import ctypes
        os.environ["UBS_ENV_PROFILE"] = "1"
        env.step(arena, 0)
        th.cuda.synchronize()
        os.environ.pop("UBS_ENV_PROFILE")
        clk = (ctypes.c_int64 * 32)()
        env._lib.ubs_env_phase_clocks(clk)
        stamps = [int(c) for c in clk]
        n = max(i for i, c in enumerate(stamps) if c) + 1
        phases = [stamps[i + 1] - stamps[i] for i in range(n - 1)]
        sched = int((env.buf.sched[..., 0] >= 0).sum())
        out[regime] = {"eager_us_per_step": round(eager, 2), "graph_us_per_step": round(1e3 * e0.elapsed_time(e1) / (5 * T), 2),
                       "scheduled_gts_per_env": sched / B, "phase_cycles_cta0": phases,
                       "mean_seen_degree": float(arena.sec("ip_seen")[T, -1]) / (B * m.n_ubs)}
    if not a.ncu:
        print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
