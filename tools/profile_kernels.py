#!/usr/bin/env python
"""Times every kernel of the hot path in isolation at the exp3 shapes (CUDA events, L2 flushed between timed launches)
and prints one JSON object; `--ncu` runs each kernel once after a short warm-up so that an outer
`ncu --set full -k regex:...` capture sees a small, known launch list.

    python tools/profile_kernels.py                       # table for profiles/
    ncu --set full ... python tools/profile_kernels.py --ncu
"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch as th  # noqa: E402

from uav_bs_ctrl_b200 import agents as A, ops, _lib  # noqa: E402
from uav_bs_ctrl_b200.builder import build_obs_graph_batch  # noqa: E402
from uav_bs_ctrl_b200.graph import batch as graph_batch  # noqa: E402
from uav_bs_ctrl_b200.synth import synth_dense_obs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--T", type=int, default=51)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="", help="'agent' = recurrent kernels only, 'gat' = relation kernels only")
    a = ap.parse_args()
    dev = th.device("cuda:0")
    B, U, G, H, T = a.envs, 8, 80, 64, a.T
    N = B * U
    args = SimpleNamespace(hidden_size=H, n_layers=1, n_heads=4, msg_size=64, key_size=16, n_rounds=1, c="tarmac",
                           o="gnn", dueling=False)
    th.manual_seed(0)
    agent = A.GnnAgent({"agent": 2, "ubs": 2, "gt": 4}, 9, args).to(dev)
    g1 = build_obs_graph_batch(*synth_dense_obs(B, U, G, "full", seed=1)).to(dev)
    gT = graph_batch([build_obs_graph_batch(*synth_dense_obs(B, U, G, "full", seed=10 + t)) for t in range(T)]).to(dev)
    flush = th.empty(256 * 1024 * 1024 // 4, dtype=th.float32, device=dev)          # 256 MB > 126 MB L2
    results = {}

    def timeit(name, fn, bytes_=None, flops=None, reps=a.reps):
        if a.ncu:
            fn()
            th.cuda.synchronize()
            return
        for _ in range(2):
            fn()
        th.cuda.synchronize()
        tot = 0.0
        best = 1e9
        for _ in range(reps):
            flush.zero_()
            s, e = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            th.cuda.synchronize()
            ms = s.elapsed_time(e)
            tot += ms
            best = min(best, ms)
        r = {"avg_us": 1e3 * tot / reps, "min_us": 1e3 * best}
        if bytes_:
            r["alg_MB"] = bytes_ / 1e6
            r["GBps_avg"] = bytes_ / (tot / reps * 1e-3) / 1e9
        if flops:
            r["TFLOPs_avg"] = flops / (tot / reps * 1e-3) / 1e12
        results[name] = r

    def gat_bytes(n, e, fs, train, bwd=False):
        P = H * (fs + 1) + 2 * H * 3 + H
        if bwd:
            return 4 * (e * fs + n * 2 + n + 1 + 2 * n * H + 2 * n * 4) + 8 * P
        return 4 * (e * fs + n * 2 + n + 1 + n * H) + 4 * P + (8 * n * 4 if train else 0)

    def gat_flops(n, e, fs):
        return e * (2 * fs * H + 6 * H + 8 * 4) + n * (4 * 2 * H + 2 * H)

    for tag, g, nn_ in (() if a.only == "agent" else (("seq%d" % T, gT, N * T),) if a.ncu else (("step", g1, N), ("seq%d" % T, gT, N * T))):
        x = g.ndata["feat"]
        for rel, src, fs in (("seen", "gt", 4), ("near", "ubs", 2)):
            conv = agent.enc.f_conv[rel]
            relg = g[rel]
            e = relg.num_edges()
            with th.no_grad():
                timeit(f"gatv2_fwd[{rel},{tag},infer]", lambda: conv(relg, (x[src], x["agent"])),
                       gat_bytes(nn_, e, fs, False), gat_flops(nn_, e, fs))
            out = conv(relg, (x[src], x["agent"]))
            go = th.randn_like(out)
            timeit(f"gatv2_fwd[{rel},{tag},train]", lambda: conv(relg, (x[src], x["agent"])),
                   gat_bytes(nn_, e, fs, True), gat_flops(nn_, e, fs))
            timeit(f"gatv2_bwd[{rel},{tag}]", lambda: th.autograd.grad(out, list(conv.parameters()), go, retain_graph=True),
                   gat_bytes(nn_, e, fs, True, bwd=True), 2 * gat_flops(nn_, e, fs))

    # recurrent part
    if a.only == "gat":
        print(json.dumps({"kernels": results}, indent=1))
        return
    block, mask1 = g1["talk"].block_mask()
    dims = agent.fused_dims(block)
    params = agent._fused_params()
    packed = agent._packed(dims, params)
    macs_row = 2 * H * H + 2 * H * 96 + (H + 64) * 3 * H + H * 3 * H + H * 9
    xin1 = th.randn(1, N, 2 * H, device=dev)
    h0 = th.randn(N, H, device=dev) * 0.1
    timeit("agent_seq_fwd[1 step,infer]", lambda: ops.agent_seq_infer(dims, packed, xin1, h0, mask1), flops=2 * macs_row * N)
    xinT = th.randn(T, N, 2 * H, device=dev, requires_grad=True)
    maskT = mask1.repeat(T).view(T, N)
    timeit(f"agent_seq_fwd[{T} steps,infer]", lambda: ops.agent_seq_infer(dims, packed, xinT.detach(), h0, maskT),
           flops=2 * macs_row * N * T)
    plist = [params[k] for k in ops.PARAM_ORDER]
    timeit(f"agent_seq_fwd[{T} steps,train]", lambda: ops.AgentSequence.apply(xinT, h0, maskT, dims, packed, *plist),
           flops=2 * macs_row * N * T)
    q, hl, _ = ops.AgentSequence.apply(xinT, h0, maskT, dims, packed, *plist)
    gq = th.randn_like(q)
    ops.TIMER = ops.KernelTimer()
    for _ in range(1 if a.ncu else 3):
        th.autograd.grad(q, [xinT] + [p for p in plist if p is not None], gq, retain_graph=True)
    summ = ops.TIMER.summary()
    ops.TIMER = None
    if "agent_seq_bwd" in summ:
        results[f"agent_seq_bwd[{T} steps] (kernel only)"] = {"avg_us": 1e3 * summ["agent_seq_bwd"]["ms"] / summ["agent_seq_bwd"]["count"]}
    timeit(f"agent sequence backward incl. weight-grad GEMMs[{T} steps]",
           lambda: th.autograd.grad(q, [xinT] + [p for p in plist if p is not None], gq, retain_graph=True))
    timeit("agent_pack", lambda: ops.agent_pack(dims, params))
    # resident-weight sequence kernels (+ their batched GEMMs)
    if ops.seq2_supported(dims):
        ops.TIMER = ops.KernelTimer()
        timeit(f"seq2 forward incl. batched GEMMs[{T} steps,train]", lambda: ops.AgentSequence2.apply(xinT, h0, maskT, dims, *plist))
        q2, _, _ = ops.AgentSequence2.apply(xinT, h0, maskT, dims, *plist)
        gq2 = th.randn_like(q2)
        timeit(f"seq2 backward incl. batched GEMMs[{T} steps]",
               lambda: th.autograd.grad(q2, [xinT] + [p for p in plist if p is not None], gq2, retain_graph=True))
        summ = ops.TIMER.summary()
        ops.TIMER = None
        for k in ("agent_seq2_fwd", "agent_seq2_bwd"):
            if k in summ:
                results[f"{k}[{T} steps] (kernel only, L2 flushed before the enclosing op)"] = {
                    "avg_us": 1e3 * summ[k]["ms"] / summ[k]["count"]}
    if not a.ncu:
        print(json.dumps({"shape": {"B": B, "U": U, "G": G, "H": H, "T": T}, "kernels": results}, indent=1))


if __name__ == "__main__":
    main()
