#!/usr/bin/env python
"""Where one act vector-step goes when replayed from a CUDA graph (exp3 shape, full degree): graphs of 50 x
{relation encoders only | fused act kernel only | both} on a sequence arena, CUDA events around 5 replays each."""
import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th  # noqa: E402

import bench  # noqa: E402
from uav_bs_ctrl_b200 import ops  # noqa: E402
from uav_bs_ctrl_b200.learner import MultiAgentQLearner  # noqa: E402


def main():
    dev = th.device("cuda:0")
    B, T = 256, 50
    learner = MultiAgentQLearner(dict(obs_shape=bench.OBS_SHAPE, state_shape=None, n_actions=bench.N_ACT, n_agents=bench.U,
                                      episode_limit=T), bench.model_args(dev, T, B))
    layout, packets = bench.make_packets(B, T, "full", seed=1234, pin=False)
    arena = learner.new_arena(bench.G)
    for t in range(T + 1):
        arena.load(t, packets[t])
    learner.begin_sequence(arena)
    net = learner.policy_net
    dims = net.arena_dims(arena)
    packed = net._packed(dims, net._fused_params())
    xin_fixed = net._arena_xin(arena, 0, 1).clone()
    eps = th.zeros((), device=dev)

    def enc_only():
        for t in range(T):
            net._arena_xin(arena, t, 1)

    L, W = arena.layout, arena.layout.words
    convs = (net.enc.f_conv["seen"], net.enc.f_conv["near"])

    def one_relation(r):
        c = convs[r]
        params = [c.fc_src.weight, c.fc_src.bias, c.fc_dst.weight, c.fc_dst.bias, c.attn, c.res_fc.weight, c.res_fc.bias]

        def run():
            for t in range(T):
                spec = ops.RelSpec(arena.ptr("x_gt", t), W, L.F_gt, arena.ptr("ip_seen", t), W, L.cap_gt) if r == 0 else \
                    ops.RelSpec(arena.ptr("x_ubs", t), W, L.F_ubs, arena.ptr("ip_near", t), W, L.cap_ubs)
                ops.SegmentEncode.apply(arena.buf, [spec], arena.ptr("x_agent", t), W, L.F_ag, 1, L.N, c._num_heads,
                                        c._out_feats, c._negative_slope, ops.GAT_RESIDUAL | ops.GAT_RELU, *params)
        return run

    def act_only():
        for t in range(T):
            ops.agent_seq_infer(dims, packed, xin_fixed, arena.h[t], arena.sec("mask", t), h_out=arena.h[t + 1].unsqueeze(0),
                                acts=arena.acts[t].unsqueeze(0), explore=(arena.explore_u[t], arena.explore_a[t], eps))

    def both():
        for t in range(T):
            net.arena_step(arena, t, explore=(arena.explore_u[t], arena.explore_a[t], eps))

    out = {}
    with th.no_grad():
        for name, fn in (("seen_only", one_relation(0)), ("near_only", one_relation(1)), ("encoders_only", enc_only),
                         ("act_kernel_only", act_only), ("both", both)):
            s = th.cuda.Stream()
            s.wait_stream(th.cuda.current_stream())
            with th.cuda.stream(s):
                fn()
            th.cuda.current_stream().wait_stream(s)
            g = th.cuda.CUDAGraph()
            with th.cuda.graph(g):
                fn()
            g.replay()
            th.cuda.synchronize()
            e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            th.cuda.synchronize()
            out[name + "_us_per_step"] = round(1e3 * e0.elapsed_time(e1) / (5 * T), 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
