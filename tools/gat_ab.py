#!/usr/bin/env python
"""A/B of the fused GATv2 forward launch shape: lanes per destination (UBS_GAT_GS) x edges per lane per pass
(UBS_GAT_EPL), at the exp3 act-step and window shapes (CUDA events, L2 flushed between timed launches).
The library reads both overrides at every launch, so one process sweeps the whole matrix.

    python tools/gat_ab.py > gpurun_out/gat_ab.json
"""
import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch as th  # noqa: E402

from uav_bs_ctrl_b200 import agents as A  # noqa: E402
from uav_bs_ctrl_b200.builder import build_obs_graph_batch  # noqa: E402
from uav_bs_ctrl_b200.graph import batch as graph_batch  # noqa: E402
from uav_bs_ctrl_b200.synth import synth_dense_obs  # noqa: E402


def main():
    dev = th.device("cuda:0")
    B, U, G, H, T = 256, 8, 80, 64, 51
    args = SimpleNamespace(hidden_size=H, n_layers=1, n_heads=4, msg_size=64, key_size=16, n_rounds=1, c="tarmac",
                           o="gnn", dueling=False)
    th.manual_seed(0)
    agent = A.GnnAgent({"agent": 2, "ubs": 2, "gt": 4}, 9, args).to(dev)
    flush = th.empty(256 * 1024 * 1024 // 4, dtype=th.float32, device=dev)
    out = {}
    for profile in ("full", "realistic"):
        g1 = build_obs_graph_batch(*synth_dense_obs(B, U, G, profile, seed=1)).to(dev)
        gT = graph_batch([build_obs_graph_batch(*synth_dense_obs(B, U, G, profile, seed=10 + t))
                          for t in range(T)]).to(dev)
        for tag, g in (("step", g1), ("window", gT)):
            x = g.ndata["feat"]
            for rel, src in (("seen", "gt"), ("near", "ubs")):
                conv, relg = agent.enc.f_conv[rel], g[rel]
                ref = {}
                for gs in ("auto", "8", "16", "32"):
                    for epl in ("auto", "1", "2"):
                        if (gs == "auto") != (epl == "auto"):
                            continue
                        for k, v in (("UBS_GAT_GS", gs), ("UBS_GAT_EPL", epl)):
                            if v == "auto":
                                os.environ.pop(k, None)
                            else:
                                os.environ[k] = v
                        with th.no_grad():
                            y = conv(relg, (x[src], x["agent"]))
                            same = bool(th.equal(y, ref.setdefault(gs, y.clone())))     # EPL must not change a bit
                            th.cuda.synchronize()
                            ts = []
                            for _ in range(8):
                                flush.zero_()
                                s, e = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
                                s.record()
                                conv(relg, (x[src], x["agent"]))
                                e.record()
                                th.cuda.synchronize()
                                ts.append(s.elapsed_time(e) * 1e3)
                        ts.sort()
                        out[f"{profile}/{tag}/{rel}/gs={gs},epl={epl}"] = {
                            "median_us": round(ts[len(ts) // 2], 2), "min_us": round(ts[0], 2),
                            "edges": relg.num_edges(), "bit_identical_across_epl": same}
    os.environ.pop("UBS_GAT_GS", None)
    os.environ.pop("UBS_GAT_EPL", None)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
