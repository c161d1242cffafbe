#!/usr/bin/env python
"""Synthetic hetero-graph sweep (BASELINE configs[4], SURVEY §8(d) "sweep"): general CSR, N_src = N_dst = n,
src ids uniform random, dst-sorted, F_in = H, heads = 4, forward + backward including grad_x_src.

Per point: CUDA-event time of the aggregate kernels (L2 flushed between launches), algorithmic bytes
  fwd 4 (E·H + E + (N+1) + 2N·H [er,res] + N·H [out])      bwd 4 (2E·H + E + 3N·H + 2N·heads)
achieved GB/s against the measured HBM peak, the whole-module fwd+bwd time (projection GEMMs included), and — with
--cpu — the CPU oracle (DGL-equivalent PyTorch restatement) on the same graph.

    python tools/sweep.py [--cpu] [--quick] > profiles/r01_sweep.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch as th  # noqa: E402
import torch.nn as nn  # noqa: E402

from uav_bs_ctrl_b200 import agents as A, graph as G, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--big", action="store_true", help="only points whose gathered rows exceed the 126 MB L2 (n = 1 M)")
    ap.add_argument("--once", action="store_true", help="one forward + backward per point (for an outer ncu capture)")
    a = ap.parse_args()
    dev = th.device("cuda:0")
    peak = 6541.1
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError):
        pass
    ns = [1000, 16000] if a.quick else [1000, 4000, 16000, 64000, 1000000]
    degs = [4, 64] if a.quick else [4, 16, 64]
    Hs = [64, 256] if a.quick else [32, 64, 128, 256]
    if a.big:
        ns, degs, Hs = [1000000], [4, 16], [64, 128]
    heads = 4
    flush = th.empty(256 * 1024 * 1024 // 4, dtype=th.float32, device=dev)
    rows = []
    for n in ns:
        for deg in degs:
            gen = th.Generator().manual_seed(0)
            E = n * deg
            src = th.randint(0, n, (E,), generator=gen)
            dst = th.randint(0, n, (E,), generator=gen).sort()[0]
            g = G.heterograph({("s", "e", "d"): (src, dst)}, num_nodes_dict={"s": n, "d": n})
            gd = g.to(dev)
            for H in Hs:
                if n * deg * H * 4 > 24e9:                 # keep the gathered volume of a point under 24 GB
                    continue
                th.manual_seed(0)
                conv = A.GATv2Conv((H, H), H // heads, heads, residual=True, allow_zero_in_degree=True,
                                   activation=nn.ReLU()).to(dev)
                xs = th.randn(n, H, device=dev, requires_grad=True)
                xd = th.randn(n, H, device=dev, requires_grad=True)
                go = th.randn(n, heads, H // heads, device=dev)

                def step():
                    o = conv(gd["e"], (xs, xd))
                    th.autograd.grad(o, [xs, xd] + list(conv.parameters()), go)

                if a.once:
                    step()
                    th.cuda.synchronize()
                    continue
                for _ in range(2):
                    step()
                th.cuda.synchronize()
                ops.TIMER = ops.KernelTimer()
                tot = 0.0
                reps = 5
                for _ in range(reps):
                    flush.zero_()
                    s, e = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
                    s.record()
                    step()
                    e.record()
                    th.cuda.synchronize()
                    tot += s.elapsed_time(e)
                summ = ops.TIMER.summary()
                ops.TIMER = None
                f_us = 1e3 * summ["gat_aggr_fwd"]["ms"] / summ["gat_aggr_fwd"]["count"]
                b_us = 1e3 * summ["gat_aggr_bwd"]["ms"] / summ["gat_aggr_bwd"]["count"]
                fb = 4 * (E * H + E + (n + 1) + 3 * n * H)
                bb = 4 * (2 * E * H + E + 3 * n * H + 2 * n * heads)
                row = {"n": n, "deg": deg, "H": H, "E": E, "aggr_fwd_us": round(f_us, 2), "aggr_bwd_us": round(b_us, 2),
                       "fwd_GBps": round(fb / f_us / 1e3, 1), "bwd_GBps": round(bb / b_us / 1e3, 1),
                       "fwd_frac_of_measured_hbm": round(fb / f_us / 1e3 / peak, 4),
                       "bwd_frac_of_measured_hbm": round(bb / b_us / 1e3 / peak, 4),
                       "module_fwd_bwd_us": round(1e3 * tot / reps, 1),
                       "el_MB": round(n * H * 4 / 1e6, 1), "beyond_L2": n * H * 4 > 126e6,
                       "projection_us": {k: round(1e3 * v["ms"] / reps, 1) for k, v in summ.items() if k.startswith("tf32x3")}}
                if a.cpu and n <= 16000:
                    from oracle import gnn_oracle as O
                    ref = O.GATv2Conv((H, H), H // heads, heads, residual=True, allow_zero_in_degree=True,
                                      activation=nn.ReLU())
                    ref.load_state_dict({k: v.cpu() for k, v in conv.state_dict().items()})
                    xc, xdc = xs.detach().cpu().requires_grad_(), xd.detach().cpu().requires_grad_()
                    goc = go.cpu()
                    t0 = time.perf_counter()
                    o = ref(g["e"], (xc, xdc))
                    th.autograd.grad(o, [xc, xdc] + list(ref.parameters()), goc)
                    row["cpu_oracle_fwd_bwd_us"] = round(1e6 * (time.perf_counter() - t0), 1)
                    row["cpu_threads"] = th.get_num_threads()
                rows.append(row)
    print(json.dumps({"peak_hbm_GBps": peak, "heads": heads, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
