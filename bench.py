#!/usr/bin/env python
"""Headline benchmark: env-steps/s of the MADRQN exp3 hot path (BASELINE.json metric).

One *step* of this bench = one training cycle of the reference's loop cadence (``algos/madrqn/run.py:81-99``) on
B parallel env instances per GPU: T vector-steps of ``learner.act`` (graph encoder + TarMAC + GRU + Q head +
ε-greedy) on replayed synthetic observations, the T ``learner.cache`` calls, and one ``learner.update`` (BPTT over
the B sequences × T just collected: T+1 policy forwards, T target forwards, double-Q loss, backward, gradient
all-reduce, value clip, AdamW, polyak).  env-steps per step = B·T per GPU (one env-step = one transition of one
env instance, all U agents acting).  Workload = BASELINE configs[1]: exp3, 8 UBS × 80 GT, hidden 64, 256 envs per
GPU, every GT visible (full degree: E_seen = N·G, the heaviest degree profile), T = episode_limit = 50.

    python bench.py --gpus N --steps K --warmup W                  # this framework
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU oracle (DGL-equivalent restatement)

Prints ONE JSON line (rank 0).  Legs of that line:
  value      observations resident in the sequence arena; the T act steps of a window are ONE CUDA graph
             (`--per-step-graphs`: one replay per vector-step), then `update_arena`
  e2e        the same cycle fed from pinned HOST packets: one H2D copy + actions D2H + stream sync per vector-step
  full_loop  the device-resident MultiUbsCoverageEnv (ubs_env_step) produces every observation: reset, ONE graph
             with T x (relations + act + env step + pack), update — nothing returns to the host inside a cycle
  roofline   dominant (kernel, shape) group of one extra eager cycle, CUDA events around each C-ABI call; `groups`
             lists every kernel group with its algorithmic GB/s and FP32 TFLOP/s
  cpu_baseline / --impl reference   the DGL-equivalent CPU oracle on the box's host cores (bounded sample)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch as th  # noqa: E402

METRIC, UNIT = "env_steps_per_sec", "env-steps/s"
U, G, H, HEADS, M, K, N_ACT = 8, 80, 64, 4, 64, 16, 9
OBS_SHAPE = {"agent": 2, "ubs": 2, "gt": 4}
FLAT, N_LAYERS, CONFIG = 0, 2, "exp3"

# BASELINE.json configs[1..3].  `exp3` is the headline (the default; what the driver runs); the others are extra runs
# (`--config`) whose lines carry their own `config.workload`.
CONFIGS = {
    "exp3": dict(U=8, G=80, H=64, envs=256, T=50, flat=False,
                 name="exp3 MADRQN gnn obs + TarMAC comm"),
    "exp2": dict(U=8, G=80, H=64, envs=256, T=50, flat=True,
                 name="exp2 MADRQN mlp obs (flattened, 423-d) + TarMAC comm"),
    "scaled": dict(U=16, G=320, H=128, envs=128, T=50, flat=False,
                   name="exp3 scaled MADRQN gnn obs + TarMAC comm (1024 envs over 8 GPUs = 128 per GPU)"),
}


def apply_config(name, a):
    """Sets the module-level workload shape from ``--config`` (explicit --envs / --T still win)."""
    global U, G, H, OBS_SHAPE, FLAT, CONFIG, N_ACT
    c = CONFIGS[name]
    U, G, H, CONFIG = c["U"], c["G"], c["H"], name
    FLAT = 2 + 5 * G + 3 * (U - 1) if c["flat"] else 0          # FlattenedObservation size (env_wrappers.py:41-54)
    OBS_SHAPE = FLAT if FLAT else {"agent": 2, "ubs": 2, "gt": 4}
    if a.envs is None:
        a.envs = c["envs"]
    if a.T is None:
        a.T = c["T"]
    return c["name"]


def model_args(device, T, n_envs):
    """exp3 model / optimiser settings: ``algos/madrqn/config.py`` defaults + ``run_exp3.py:30-54`` overrides, with
    hidden_size=64 as BASELINE.json names it."""
    return SimpleNamespace(device=str(device), o="mlp" if FLAT else "gnn", c="tarmac", share_reward=False, hidden_size=H,
                           n_layers=N_LAYERS,
                           n_heads=HEADS, msg_size=M, key_size=K, n_rounds=1, lr=2.5e-4, gamma=0.99, polyak=0.999,
                           batch_size=1, replay_size=2, max_seq_len=T, anneal_lr=False, double_q=True, dueling=False,
                           mixer=False, n_envs=n_envs)


def make_episode(B, T, profile, seed):
    """T+1 host-resident batched observation graphs + rewards / dones, seeded (SURVEY §8(d) distributions)."""
    from uav_bs_ctrl_b200.builder import build_obs_graph_batch
    from uav_bs_ctrl_b200.synth import synth_dense_obs
    graphs = []
    for t in range(T + 1):
        a, gt, ubs, adj = synth_dense_obs(B, U, G, profile, seed=seed + t)
        if FLAT:                                    # exp2: flattened local observations on the comm graph
            flat = th.cat((a.reshape(B, U, -1), gt.reshape(B, U, -1), ubs.reshape(B, U, -1)), 2)
            graphs.append(build_obs_graph_batch(flat, th.zeros(B, U, 0, 5), th.zeros(B, U, 0, 3), adj))
        else:
            graphs.append(build_obs_graph_batch(a, gt, ubs, adj))
    gen = th.Generator().manual_seed(seed + 7919)
    rews = th.rand(T, B, U, generator=gen)
    dones = th.zeros(T, B)
    dones[T - 1] = 1.0                              # episode_limit reached on the last step (bad_mask mutes it)
    bad = dones.clone()
    return graphs, rews, dones, bad


def graph_bytes(g):
    n = 0
    for f in g._nframes.values():
        for v in f.values():
            n += v.numel() * v.element_size()
    for r in g._csr.values():
        for t in (r.indptr, r.src_idx, r.eid, r.mask):
            if t is not None:
                n += t.numel() * t.element_size()
    return n


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.index = None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, val in zip(names, p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ our arm
def train_cycle(learner, obs, rews, dones, bad, T, B, eps, host_io, acts_host=None):
    """One step of the bench.  ``host_io``: observations come from pinned host memory and actions / loss go back."""
    dev = learner.device
    h = learner.init_hidden(B).to(dev)
    o = learner.stage(obs[0]) if host_io else obs[0]
    for t in range(T):
        acts, h2 = learner.act(o, h, eps)
        if host_io:                                     # the env needs the actions on the host before it can step
            acts_host[t].copy_(acts, non_blocking=True)
            th.cuda.current_stream().synchronize()
        o2 = learner.stage(obs[t + 1]) if host_io else obs[t + 1]
        learner.cache(o, h, None, acts, rews[t], o2, h2, None, dones[t], bad[t])
        o, h = o2, h2
    return learner.update(samples=[learner.buffer.memory[-1]], sync=host_io)


def train_cycle_arena(learner, arena, packets, T, eps, host_io, acts_host=None):
    """Same step on the packed path: observations are packets staged into the sequence arena (ONE copy each), act is
    three kernels per vector-step (optionally a replayed CUDA graph), update reads the arena in place."""
    learner.begin_sequence(arena)
    if host_io:
        arena.load(0, packets[0])
    elif getattr(learner.args, "cuda_graphs", False) and not getattr(learner.args, "per_step_graphs", False):
        # observations already resident: nothing on the host between the T act steps -> ONE graph for the window
        learner.rollout_arena(None, arena, eps)
        return learner.update_arena(arena, sync=False)
    for t in range(T):
        acts = learner.act_arena(arena, t, eps)
        if host_io:
            acts_host[t].copy_(acts, non_blocking=True)
            th.cuda.current_stream().synchronize()           # the env needs the actions before it can step
            arena.load(t + 1, packets[t + 1])                 # next observation (+ reward / done) arrives
    return learner.update_arena(arena, sync=host_io)


def make_packets(B, T, profile, seed, pin):
    from uav_bs_ctrl_b200.arena import PacketLayout, ObsPacket
    from uav_bs_ctrl_b200.synth import synth_dense_obs
    L = PacketLayout(B, U, G, flat_dim=FLAT)
    gen = th.Generator().manual_seed(seed + 7919)
    out = []
    for t in range(T + 1):
        a, gt, ubs, adj = synth_dense_obs(B, U, G, profile, seed=seed + t)
        done = th.ones(B) if t == T else th.zeros(B)       # episode_limit reached on the last step (bad_mask mutes it)
        pk = ObsPacket(L, pin=pin).fill_from_dense(a, gt, ubs, adj, rew=th.rand(B, U, generator=gen), done=done, bad=done)
        if FLAT:                                           # exp2: the MLP encoder reads flattened local observations
            flat = th.cat((a.reshape(B * U, -1), gt.reshape(B * U, -1), ubs.reshape(B * U, -1)), 1)   # agent | gt | ubs
            pk.sec("x_flat").view(B * U, L.flat_ld)[:, :FLAT] = flat
        out.append(pk)
    return L, out


def algorithmic_bytes_gat(meta):
    """star-layout GATv2 relation.  fwd 4(E·F_s + N·F_d + N+1 + N·H) + 4P (+8·N·heads when stats are
    saved); bwd 4(E·F_s + N·F_d + N+1 + 2N·H + 2N·heads) + 8P."""
    n, e, fs, fd, heads, d, train = meta
    Hh = heads * d
    P = Hh * (fs + 1) + 2 * Hh * (fd + 1) + Hh
    fwd = 4 * (e * fs + n * fd + (n + 1) + n * Hh) + 4 * P + (8 * n * heads if train else 0)
    bwd = 4 * (e * fs + n * fd + (n + 1) + 2 * n * Hh + 2 * n * heads) + 8 * P
    return fwd, bwd


def algorithmic_cost(name, meta):
    """(bytes, flops) one launch of a library kernel must move / perform (SURVEY §8(d) formulas)."""
    if name.startswith("gatv2"):
        n, e, fs, fd, heads, d, train = meta
        Hh = heads * d
        fwd, bwd = algorithmic_bytes_gat(meta)
        fl = e * (2 * fs * Hh + 6 * Hh + 8 * heads) + n * (4 * fd * Hh + 2 * Hh)
        return (fwd, fl) if name.endswith("fwd") else (bwd, 2 * fl)
    if name == "env_step":                              # state in + out, staged rows + packet rows, per env instance
        B_, U_, G_, Fg = meta
        words = U_ * 4 * 2 + G_ * 2 + 4 * G_ * 2 + 2 * (U_ * G_ * Fg + U_ * (U_ - 1) * 2) + 6 * U_ + 2 * G_
        return 4 * B_ * words, B_ * U_ * G_ * 60
    if name.startswith("tf32x3"):                       # C (M,N) = A (M,K) B (K,N): fp32 in / out, 3 TF32 MMAs per product
        M_, N_, K_ = meta
        return 4 * (M_ * K_ + K_ * N_ + M_ * N_), 2 * M_ * N_ * K_
    if name == "colsum":                                # read once, C sums out
        R_, C_ = meta
        return 4 * (R_ * C_ + C_), R_ * C_
    if name == "relu_bwd_colsum":                       # read dy and y, write dx, C sums out
        R_, C_ = meta
        return 4 * (3 * R_ * C_ + C_), 2 * R_ * C_
    if name == "agent_act_rel":
        # one-kernel act step: both star relations at capacity degree (cap rows per destination) + the agent step; the
        # relation outputs never leave the SM, so no xin traffic
        _, N_, ints, (F_gt, cap_gt, F_ubs, cap_ubs, heads) = meta
        Hh, M_, K_, A_, U_, Fin, flags = ints
        b_seen, f_seen = algorithmic_cost("gatv2_fwd", (N_, N_ * cap_gt, F_gt, 2, heads, Hh // heads, False))
        b_near, f_near = algorithmic_cost("gatv2_fwd", (N_, N_ * cap_ubs, F_ubs, 2, heads, Hh // heads, False))
        b_step, f_step = algorithmic_cost("agent_seq_fwd", (1, N_, ints, False))
        return b_seen + b_near + b_step - 4 * N_ * (2 * Hh + Fin), f_seen + f_near + f_step
    T_, N_, ints, train = meta
    Hh, M_, K_, A_, U_, Fin, flags = ints
    tm = bool(flags & 2)
    Vp = ((M_ + 2 * K_ + 3) // 4 * 4) if tm else 0
    if name.startswith("agent_seq2"):
        w = Hh * Vp + M_ * 3 * Hh + Hh * 3 * Hh + 3 * Hh
        macs = Hh * Vp + M_ * 3 * Hh + Hh * 3 * Hh + (U_ * (K_ + M_) if tm else 0)
        if name.endswith("fwd"):
            per_row = Vp + 3 * Hh + 1 + Hh + ((Vp + U_ + M_ + 4 * Hh) if train else 0)
        else:
            per_row = 4 * Hh + 2 * Hh + Vp + U_ + 6 * Hh + Vp
            w = 3 * Hh * Hh + 3 * Hh * M_
        return 4 * (T_ * N_ * per_row + w), 2 * macs * T_ * N_
    # streaming step kernel (act / H > 64 windows)
    I = Hh + M_ if tm else Hh
    w = (Fin * Hh if flags & 1 else 0) + 2 * Hh * Vp + I * 3 * Hh + Hh * 3 * Hh + Hh * A_
    per_row = Fin + 2 * Hh + A_ + 2 + ((I + Vp + U_ + 4 * Hh) if train else 0)
    return 4 * (T_ * N_ * per_row + w), 2 * (w + (U_ * (K_ + M_) if tm else 0)) * T_ * N_


def run_ours(a):
    from uav_bs_ctrl_b200 import _lib, dist, ops
    from uav_bs_ctrl_b200.learner import MultiAgentQLearner
    local = dist.init_from_env("nccl")
    world, rank = dist.world_size(), dist.rank()
    assert th.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    dev = th.device("cuda", local)
    th.cuda.set_device(dev)
    _lib.load()
    B, T = a.envs, a.T
    th.manual_seed(0)
    learner = MultiAgentQLearner(dict(obs_shape=OBS_SHAPE, state_shape=None, n_actions=N_ACT, n_agents=U,
                                      episode_limit=T), model_args(dev, T, B))
    eps = 0.05
    d2h = T * B * U * 8 + 4 + T * B * U * 4
    use_arena = a.path == "arena"
    if use_arena:
        learner.args.cuda_graphs = not a.no_graphs
        learner.args.per_step_graphs = a.per_step_graphs
        learner.args.act_pdl = not a.no_pdl
        learner.args.overlap_target = not a.no_overlap
        learner.args.update_graph = not a.no_update_graph
        learner.policy_net.use_seq2_act = a.act_seq2
        layout, packets = make_packets(B, T, a.profile, seed=1234 + 100 * rank, pin=True)
        h2d = sum(p.used_words() for p in packets) * 4        # what arena.load ships: header + the CSR rows in use
        arena = learner.new_arena(G)
        for t in range(T + 1):
            arena.load(t, packets[t])
        value_step = lambda: train_cycle_arena(learner, arena, None, T, eps, False)
    else:
        graphs, rews, dones, bad = make_episode(B, T, a.profile, seed=1234 + 100 * rank)
        h2d = sum(graph_bytes(g) for g in graphs) + rews.numel() * 4 + dones.numel() * 8
        dev_graphs = [g.to(dev) for g in graphs]
        dev_rews, dev_dones, dev_bad = rews.to(dev), dones.to(dev), bad.to(dev)
        value_step = lambda: train_cycle(learner, dev_graphs, dev_rews, dev_dones, dev_bad, T, B, eps, False)

    def barrier():
        if world > 1:
            th.distributed.barrier()
        th.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks=False):
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()                     # nvidia-smi needs ~0.5 s to start: launch it before the warm-up
        for _ in range(warmup):
            fn()
        if sampler:
            for _ in range(40):                 # ~0.5 s under the same load so that nvidia-smi is sampling by now
                fn()                            # (a FIXED count: every rank must issue the same number of collectives)
        barrier()
        _lib.reset_launch_count()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        launches = _lib.launch_count()
        clocks = sampler.stop() if sampler else None
        ms = dist.all_reduce_max_scalar(e0.elapsed_time(e1), dev)
        return ms, launches, clocks

    # ---- value: inputs resident in HBM
    mem0 = th.cuda.memory_allocated()
    th.cuda.reset_peak_memory_stats()
    ms, launches, clocks = timed(value_step, a.steps, a.warmup, sample_clocks=True)
    value = world * B * T * a.steps / (ms * 1e-3)
    ws_mib = (th.cuda.max_memory_allocated() - mem0) / 2**20      # activations one update writes and reads back

    # ---- strong scaling (extra leg, N > 1): the SAME total number of envs split over the ranks (256 ->
    # 256 / N per GPU); `value` above stays the weak-scaling number the driver computes its efficiency from
    strong = None
    if world > 1 and use_arena and not a.no_strong and B % world == 0 and B // world >= 2:
        Bs = B // world
        th.manual_seed(0)
        learner_s = MultiAgentQLearner(dict(obs_shape=OBS_SHAPE, state_shape=None, n_actions=N_ACT, n_agents=U,
                                            episode_limit=T), model_args(dev, T, Bs))
        learner_s.args.cuda_graphs = not a.no_graphs
        learner_s.args.per_step_graphs = a.per_step_graphs
        _, packets_s = make_packets(Bs, T, a.profile, seed=4321 + 100 * rank, pin=False)
        arena_s = learner_s.new_arena(G)
        for t in range(T + 1):
            arena_s.load(t, packets_s[t])
        ms_s, _, _ = timed(lambda: train_cycle_arena(learner_s, arena_s, None, T, eps, False), a.steps, a.warmup)
        strong = {"value": B * T * a.steps / (ms_s * 1e-3), "unit": UNIT, "scaling": "strong", "total_envs": B,
                  "envs_per_gpu": Bs, "ms_per_step": ms_s / a.steps}
        del learner_s, arena_s, packets_s

    # ---- e2e: pinned host observations in, actions / loss / q-values out, every step
    e2e = None
    if not a.no_e2e:
        acts_host = th.empty(T, B * U, dtype=th.int64).pin_memory()
        if use_arena:
            e2e_step = lambda: train_cycle_arena(learner, arena, packets, T, eps, True, acts_host)
        else:
            pin_graphs = [g.pin_memory() for g in graphs]
            pin_rews, pin_dones, pin_bad = rews.pin_memory(), dones.pin_memory(), bad.pin_memory()
            e2e_step = lambda: train_cycle(learner, pin_graphs, pin_rews, pin_dones, pin_bad, T, B, eps, True, acts_host)
        ms_e, _, _ = timed(e2e_step, a.steps, max(1, a.warmup // 2 + 1))
        e2e = {"value": world * B * T * a.steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e / a.steps}

    # ---- full loop: the device-resident env (ubs_env_step) closes the loop — reset, T x (act -> env.step), update, with
    # nothing returning to the host; initial layouts come from the RNG-matched host sampler, drawn ahead of time
    full = None
    if use_arena and not a.no_full and not FLAT:
        from uav_bs_ctrl_b200 import envs as E
        m = E.DenseHotSpot(n_ubs=U, n_grps=G // 5, gts_per_grp=5, episode_limit=T)
        env = E.MultiUbsCoverageVecEnv(n_envs=B, device=dev, map=m)
        env.seed = 10_000 * (rank + 1)

        def full_step():
            learner.begin_sequence(arena)
            env.reset(arena, 0)                       # fresh layouts every episode: sampled on the device (Philox)
            learner.rollout_arena(env, arena, eps)
            return learner.update_arena(arena, sync=False)

        ms_f, launches_f, _ = timed(full_step, a.steps, a.warmup)
        deg = float(arena.sec("ip_seen")[:, -1].float().mean()) / (B * U)
        ev0, ev1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        er0, er1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        er0.record()
        for _ in range(10):
            env.reset(arena, 0)
        er1.record()
        ev0.record()
        for t in range(T):                                # env alone: 2 kernels per step (step + pack), eager launches
            env.step(arena, t)
        ev1.record()
        th.cuda.synchronize()
        env_us = 1e3 * ev0.elapsed_time(ev1) / T
        reset_us = 1e3 * er0.elapsed_time(er1) / 10
        full = {"value": world * B * T * a.steps / (ms_f * 1e-3), "unit": UNIT, "ms_per_step": ms_f / a.steps,
                "gpu_launches": int(launches_f), "mean_seen_degree": deg, "env_step_us": env_us,
                "env_steps_per_sec_env_alone": B / (env_us * 1e-6), "reset_us": reset_us,
                "env": f"device-resident MultiUbsCoverageEnv, DenseHotSpot {U} UBS x {G} GT (maps.py:83-113), "
                       f"episode_limit {T}, eps-greedy {eps}; every episode starts from fresh layouts sampled on the device "
                       f"(ubs_env_sample_layouts, Philox); "
                       "one CUDA graph per rollout" + ("" if learner.args.cuda_graphs else " (graphs off)")}
        if rank == 0 and world == 1 and not a.no_cpu:
            full["cpu_env_port"] = cpu_env_port(m, B)
        for t in range(T + 1):                            # the value / roofline passes below replay the synthetic episode
            arena.load(t, packets[t])

    # ---- where the step goes: act loop (T replayed graphs) vs update, 3 extra cycles (every rank: all-reduce inside)
    phases = None
    if use_arena:
        ev = [[th.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(3)]
        # park the stream behind a ~50 ms spin kernel: the host enqueues the three cycles meanwhile, so the events bracket
        # back-to-back device work (as in the timed region, where the host runs ahead of the device) and not host gaps
        th.cuda._sleep(int(0.05 * 1.9e9))
        for i in range(3):
            learner.begin_sequence(arena)
            ev[i][0].record()
            if learner.args.cuda_graphs and not a.per_step_graphs:
                learner.rollout_arena(None, arena, eps)
            else:
                for t in range(T):
                    learner.act_arena(arena, t, eps)
            ev[i][1].record()
            learner.update_arena(arena, sync=False)
            ev[i][2].record()
        th.cuda.synchronize()
        phases = {"act_ms": sum(e[0].elapsed_time(e[1]) for e in ev) / 3,
                  "update_ms": sum(e[1].elapsed_time(e[2]) for e in ev) / 3}

    # ---- roofline of the dominant kernel: CUDA events around every C-ABI call during one extra step
    roofline = None
    if use_arena:
        learner.args.cuda_graphs = False                 # replayed graphs bypass the Python-side event hooks
    if rank == 0:
        ops.TIMER = ops.KernelTimer()
    # The eager step is host-bound (Python between launches): park the stream behind a ~150 ms spin kernel so the host
    # runs ahead and every event pair brackets back-to-back device work instead of host gaps.
    th.cuda._sleep(int(0.15 * 1.9e9))
    value_step()                                         # every rank runs it: the update contains the all-reduce
    th.cuda.synchronize()
    if rank == 0:
        recs = [(n, m, s_.elapsed_time(e_)) for n, m, s_, e_ in ops.TIMER.records]
        ops.TIMER = None
        groups, by_name = {}, {}
        for n, m, ms_ in recs:                         # one group per (kernel, shape): act launches vs window launches
            g_ = groups.setdefault((n, m), [0, 0.0])
            g_[0] += 1
            g_[1] += ms_
            b_ = by_name.setdefault(n, [0, 0.0])
            b_[0] += 1
            b_[1] += ms_
        (dom, meta), (cnt, tot_ms) = max(groups.items(), key=lambda kv: kv[1][1])

        def group_rate(k, v):
            b_, f_ = algorithmic_cost(k[0], k[1])
            s_ = v[1] * 1e-3 / v[0]
            return {"launches": v[0], "avg_us": round(1e6 * s_, 2), "GBps": round(b_ / s_ / 1e9, 1),
                    "fp32_tflops": round(f_ / s_ / 1e12, 2)}

        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else \
            (6650.0, "fallback (B200_PROFILING.md)")
        nbytes, nflops = algorithmic_cost(dom, meta)
        # DRAM bytes per launch of the dominant kernel: NOT measured in this run (that needs a profiler) — the value of
        # the committed `ncu --set full` capture of the same kernel and shape (profiles/traffic.json), labelled as such
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(f"{dom}{list(meta[:2])}")
            traffic_src = f"static: {tj.get('_source', 'profiles/traffic.json')}" if traffic is not None else None
        except (OSError, ValueError):
            pass
        avg_s = tot_ms * 1e-3 / cnt
        achieved = nbytes / avg_s / 1e9
        fp32_peak = 148 * 128 * 2 * (clocks["sm_mhz"] or 1965.0) * 1e6 / 1e12 if clocks else 74.5
        roofline = {"bound": "hbm", "kernel": dom, "shape": str(meta), "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "launches_timed": cnt, "avg_launch_us": 1e6 * avg_s, "algorithmic_bytes_per_launch": nbytes,
                    "algorithmic_flops_per_launch": nflops,
                    "fp32": {"achieved_tflops": nflops / avg_s / 1e12, "peak_tflops": fp32_peak,
                             "frac": nflops / avg_s / 1e12 / fp32_peak,
                             "note": "these kernels are FP32-issue bound (arithmetic intensity far right of the ridge); "
                                     "the HBM fraction is reported because the contract asks for it"},
                    "timing": "CUDA events around each C-ABI call of one extra EAGER step (launches queued behind a spin "
                              "kernel so the device runs them back to back) after the timed region; inside the timed "
                              "region the act kernels are CUDA-graph nodes: their in-graph time is phases.act_ms / T",
                    "kernel_ms_per_step": {k: round(v[1], 4) for k, v in by_name.items()},
                    "kernel_calls_per_step": {k: v[0] for k, v in by_name.items()},
                    "groups_ms": {f"{k[0]}{list(k[1][:2])}": round(v[1], 4) for k, v in groups.items()},
                    # per (kernel, shape) group: launches, achieved algorithmic GB/s and FP32 TFLOP/s of one launch
                    "groups": {f"{k[0]}{list(k[1][:2])}": group_rate(k, v) for k, v in groups.items()
                               }}

    # ---- CPU baseline (oracle, rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu = cpu_reference(B, min(a.cpu_T, T), a.profile, steps=2, warmup=1)     # ~20 s of host work at T = 50

    if world > 1:
        th.distributed.barrier()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{a.workload_name}, {U} UBS x {G} GT, hidden={H}, "
                                       f"{B} envs/GPU, T={T}, degree profile '{a.profile}'", "name": CONFIG,
                           "env_steps_per_step": world * B * T, "update_batch": f"{B} sequences x {T} per GPU",
                           "parallelism": f"dp{world}", "path": a.path + ("" if a.no_graphs or a.path != "arena" else
                                             "+cudagraphs(per step)" if a.per_step_graphs else "+cudagraph(act window)"),
                           "l2_policy": "inputs exceed L2: every step streams "
                           f"{h2d / 2**20:.0f} MiB of observation packets and ~{ws_mib:.0f} MiB of window activations "
                           "(> 126 MB L2)"},
                "roofline": roofline, "phases": phases, "cpu_baseline": cpu, "e2e": e2e, "full_loop": full,
                "strong_scaling": strong,
                "clocks": clocks,
                "gpu_launches": int(launches)}
        print(json.dumps(line), flush=True)
    if world > 1:
        th.distributed.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference(B, T, profile, steps, warmup, seed=1234):
    """Times the CPU oracle (pure-PyTorch restatement of the DGL path — DGL 0.9.0 itself is not installable) on the
    same cycle with all host threads.  Returns the ``cpu_baseline`` object."""
    from oracle import gnn_oracle as O
    cores = os.cpu_count() or 1
    th.set_num_threads(cores)
    args = model_args("cpu", T, B)
    th.manual_seed(0)
    policy, target = O.GnnAgent(OBS_SHAPE, N_ACT, args), O.GnnAgent(OBS_SHAPE, N_ACT, args)
    target.load_state_dict(policy.state_dict())
    opt = th.optim.AdamW(policy.parameters(), lr=args.lr)
    graphs, rews, dones, bad = make_episode(B, T, profile, seed)
    dones_m = ((1 - bad) * dones).view(T, B, 1)

    def cycle():
        h = policy.init_hidden().expand(B * U, -1)
        hs, acts = [h], []
        for t in range(T):
            with th.no_grad():
                q, h = policy(graphs[t], h)
            acts.append(q.argmax(1, keepdim=True))
            hs.append(h)
        loss, _ = O.bptt_loss(policy, target, graphs, hs[0], hs[1], th.stack(acts), rews, dones_m, args.gamma,
                              args.double_q, U)
        opt.zero_grad()
        loss.backward()
        th.nn.utils.clip_grad_value_(policy.parameters(), clip_value=1)
        opt.step()
        with th.no_grad():
            for p, pt in zip(policy.parameters(), target.parameters()):
                pt.mul_(args.polyak).add_((1 - args.polyak) * p)
        return float(loss.detach())

    for _ in range(warmup):
        cycle()
    t0 = time.perf_counter()
    for _ in range(steps):
        cycle()
    dt = time.perf_counter() - t0
    return {"value": B * T * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} cycle(s) of {B} envs x T={T} steps (act x T + one BPTT update), torch CPU oracle, "
                      f"{cores} threads, {dt:.1f} s", "seconds": dt}


def cpu_env_port(m, B, steps=20):
    """The device env's CPU counterpart, timed on one host core: the serial host build of the SAME env core
    (oracle/env_host.cpp — test infrastructure; the reference's numpy env itself cannot travel to the GPU box, its
    measured rate in the build container is 150 env-steps/s per core at 8 x 80)."""
    import ctypes as C
    import numpy as np
    from uav_bs_ctrl_b200 import envs as E
    from uav_bs_ctrl_b200.arena import PacketLayout
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libubs_env_host.so"))
    lib.ubs_env_host_step.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_int]
    lib.ubs_env_host_scratch_words.restype = C.c_int64
    lib.ubs_env_host_scratch_words.argtypes = [C.c_void_p, C.c_int64]
    cfg = E.make_cfg(m)
    buf = E.EnvBuffers(cfg, B, "cpu", lib.ubs_env_host_scratch_words(C.byref(cfg), B))
    buf.set_layout(*E.sample_layouts(m, range(B)))
    L = PacketLayout(B, cfg.n_ubs, cfg.n_gts)
    pkt = th.zeros(L.words, dtype=th.int32)
    st, pk = buf.state_struct(), E.packet_struct(L, pkt)
    lib.ubs_env_host_step(C.byref(cfg), C.byref(st), None, C.byref(pk), buf.scratch.data_ptr(), B, 1)
    acts = th.as_tensor(np.random.RandomState(0).randint(0, cfg.n_actions, size=(steps, B * cfg.n_ubs)), dtype=th.int64)
    t0 = time.perf_counter()
    for k in range(steps):
        lib.ubs_env_host_step(C.byref(cfg), C.byref(st), acts[k].data_ptr(), C.byref(pk), buf.scratch.data_ptr(), B, 0)
    dt = time.perf_counter() - t0
    return {"value": B * steps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{steps} steps of {B} env instances, serial C++ build of the env core, {dt:.2f} s"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(a.gpus)))
    B, T = a.envs, a.T                       # the SAME cycle as the GPU arm (T = 50): ~6 s of host work per step
    cpu = cpu_reference(B, T, a.profile, a.steps, a.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * cpu["seconds"] / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{a.workload_name}, {U} UBS x {G} GT, hidden={H}, {B} envs, "
                                   f"T={T}, degree profile '{a.profile}'",
                       "env_steps_per_step": B * T,
                       "note": "reference's DGL-on-CPU path restated op-for-op in PyTorch (DGL 0.9.0 not installable)"},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="exp3", choices=sorted(CONFIGS), help="BASELINE.json workload (default: the headline)")
    ap.add_argument("--envs", type=int, default=None, help="parallel env instances per GPU (default: the config's)")
    ap.add_argument("--T", type=int, default=None, help="sequence length = episode_limit (default: the config's, 50)")
    ap.add_argument("--cpu-T", dest="cpu_T", type=int, default=50,
                    help="sequence length of the cpu_baseline leg (default: the full T; its sample is bounded by running 1 + 2 cycles)")
    ap.add_argument("--profile", default="full", choices=["full", "realistic", "random"])
    ap.add_argument("--path", default="arena", choices=["arena", "graph"],
                    help="arena: packed packets + sequence arena (+ CUDA graphs); graph: reference-shaped graph objects")
    ap.add_argument("--no-graphs", action="store_true", help="arena path without CUDA-graph replay of the act step")
    ap.add_argument("--per-step-graphs", action="store_true",
                    help="value leg: one graph replay per vector-step (as the e2e leg must) instead of one per window")
    ap.add_argument("--act-seq2", action="store_true", help="act step through the resident-weight kernel (T=1) + small GEMMs")
    ap.add_argument("--no-update-graph", action="store_true", help="update: eager launches instead of one CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="update: target window on the main stream (no second stream)")
    ap.add_argument("--no-pdl", action="store_true", help="rollout graph without programmatic dependent launch of the act steps")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-full", action="store_true", help="skip the full-loop leg (device env in the loop)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling extra leg of multi-GPU runs")
    a = ap.parse_args()
    a.workload_name = apply_config(a.config, a)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
