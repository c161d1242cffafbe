#!/usr/bin/env python
"""End-to-end training on the device: MADRQN (graph observation encoder + TarMAC) on B parallel instances of the
device-resident MultiUbsCoverageEnv, reference loop cadence (algos/madrqn/run.py:81-99: act / env.step every step, one
BPTT update per episode window), epsilon annealed 1 -> 0.05, fresh layouts every episode sampled on the device
(ubs_env_sample_layouts).  Prints one JSON object with the mean episode return (info['EpRet'], mubs_cov.py:113-119) per block
of cycles — the check that env + encoder + comm + learner work together, not a benchmark.

    python tools/train_demo.py --cycles 1500 > gpurun_out/train_demo.json
"""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th  # noqa: E402

from uav_bs_ctrl_b200 import envs as E  # noqa: E402
from uav_bs_ctrl_b200.learner import MultiAgentQLearner  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cycles", type=int, default=1500)
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--map", default="4ubs")
    ap.add_argument("--block", type=int, default=100)
    ap.add_argument("--lr", type=float, default=5e-4)
    a = ap.parse_args()
    dev = th.device("cuda:0")
    B = a.envs
    env = E.MultiUbsCoverageVecEnv(a.map, B, dev)
    T = env.episode_limit
    args = SimpleNamespace(device="cuda", o="gnn", c="tarmac", share_reward=False, hidden_size=64, n_layers=2, n_heads=4,
                           msg_size=64, key_size=16, n_rounds=1, lr=a.lr, gamma=0.99, polyak=0.995, batch_size=1,
                           replay_size=2, max_seq_len=T, anneal_lr=False, double_q=True, dueling=False, mixer=False,
                           n_envs=B, cuda_graphs=True)
    th.manual_seed(0)
    learner = MultiAgentQLearner(env.get_env_info(), args)
    arena = learner.new_arena(env.cfg.n_gts)
    env.seed = 1
    curve, acc = [], []
    decay = int(0.6 * a.cycles)
    th.cuda.synchronize()
    t0 = time.perf_counter()
    for c in range(a.cycles):
        eps = max(0.05, 1.0 - 0.95 * c / max(decay, 1))
        learner.begin_sequence(arena)
        env.reset(arena, 0)                                        # fresh layouts, sampled on the device
        learner.rollout_arena(env, arena, eps)
        acc.append(env.buf.info[:, 0].mean())                      # EpRet of the finished episodes (device scalar)
        out = learner.update_arena(arena, sync=False)
        if (c + 1) % a.block == 0:
            curve.append({"cycle": c + 1, "eps": round(eps, 3), "mean_ep_ret": float(th.stack(acc).mean()),
                          "loss": float(out["LossQ"]), "mean_seen_degree": float(arena.sec("ip_seen")[:, -1].float().mean()) / (B * env.n_agents),
                          "fair_idx": float(env.buf.info[:, 4].mean()), "total_throughput": float(env.buf.info[:, 1].mean())})
            acc = []
    th.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"map": a.map, "envs": B, "T": T, "cycles": a.cycles, "env_steps": a.cycles * B * T,
                      "seconds": round(dt, 2), "env_steps_per_sec": a.cycles * B * T / dt,
                      "curve": curve}, indent=1))


if __name__ == "__main__":
    main()
