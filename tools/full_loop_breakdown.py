import sys, time; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import bench as Bn
from uav_bs_ctrl_b200 import envs as E, ops
from uav_bs_ctrl_b200.learner import MultiAgentQLearner
dev = th.device("cuda:0"); B, T = 256, 50
th.manual_seed(0)
learner = MultiAgentQLearner(dict(obs_shape=Bn.OBS_SHAPE, state_shape=None, n_actions=9, n_agents=8, episode_limit=T), Bn.model_args(dev, T, B))
learner.args.cuda_graphs = True
arena = learner.new_arena(80)
m = E.DenseHotSpot(n_ubs=8, n_grps=16, episode_limit=T)
env = E.MultiUbsCoverageVecEnv(n_envs=B, device=dev, map=m)
pool = env.make_layout_pool(2, seed0=5)
def cyc(mode):
    ev = [th.cuda.Event(enable_timing=True) for _ in range(4)]
    learner.begin_sequence(arena)
    ev[0].record()
    if mode == "dev": env.reset(arena, 0)
    else: env.reset(arena, 0, layouts=pool[0])
    ev[1].record()
    learner.rollout_arena(env, arena, 0.05)
    ev[2].record()
    learner.update_arena(arena, sync=False)
    ev[3].record()
    return ev
for mode in ("pool", "dev", "pool", "dev"):
    for _ in range(4): cyc(mode)
    th.cuda.synchronize()
    t0 = time.perf_counter()
    evs = [cyc(mode) for _ in range(10)]
    th.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 10 * 1e3
    r = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) / 10 for i in range(3)]
    print(mode, "wall ms/cycle %.2f" % wall, "reset %.3f rollout %.3f update %.3f" % tuple(r), "deg", float(arena.sec("ip_seen")[:, -1].float().mean()) / (B * 8))
ops.SAVE_SCORES = False
for _ in range(4): cyc("dev")
th.cuda.synchronize()
evs = [cyc("dev") for _ in range(10)]
th.cuda.synchronize()
print("no saved scores: update %.3f" % (sum(e[2].elapsed_time(e[3]) for e in evs) / 10))
