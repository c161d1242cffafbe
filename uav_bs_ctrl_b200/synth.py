"""Seeded synthetic observations in the reference env's dense format.

Feature distributions follow ``envs/mubs_cov/mubs_cov.py:215-242``: agent = normalised position U(0,1)²;
gt = [visible, Δx, Δy ~ U(-1,1), rate ~ U(0,1)·1[p<0.3], avg-rate ~ U(0,1)]; ubs = [visible, Δx, Δy ~ U(-1,1)].
Degree profiles for the ``seen`` relation: ``full`` (every GT visible), ``realistic`` (per-env visibility
probability drawn from {0.01, 0.1, 0.5, 1.0}, includes degree 0), ``random`` (mean degree ≈ 1).
"""
from __future__ import annotations

import torch as th


def synth_dense_obs(B: int, U: int, G: int, profile: str = "full", seed: int = 1234, comm_p: float = 1.0,
                    near_p: float = 1.0, F_gt: int = 4, device="cpu"):
    """Returns ``(agent_obs (B,U,2), gt_obs (B,U,G,1+F_gt), ubs_obs (B,U,U-1,3), comm_adj (B,U,U) bool)``."""
    gen = th.Generator().manual_seed(seed)
    r = lambda *s: th.rand(*s, generator=gen)
    agent = r(B, U, 2)
    gt = th.empty(B, U, G, 1 + F_gt)
    gt[..., 1:3] = r(B, U, G, 2) * 2 - 1
    gt[..., 3] = r(B, U, G) * (r(B, U, G) < 0.3)
    if F_gt > 3:
        gt[..., 4:] = r(B, U, G, F_gt - 3)
    if profile == "full":
        vis = th.ones(B, U, G, dtype=th.bool)
    elif profile == "realistic":
        p = th.tensor([0.01, 0.1, 0.5, 1.0])[th.randint(0, 4, (B,), generator=gen)]
        vis = r(B, U, G) < p.view(B, 1, 1)
    elif profile == "random":
        vis = r(B, U, G) < (1.0 / max(G, 1))
    else:
        raise KeyError(profile)
    gt[..., 0] = vis.float()
    ubs = th.empty(B, U, max(U - 1, 0), 3)
    ubs[..., 1:] = r(B, U, max(U - 1, 0), 2) * 2 - 1
    ubs[..., 0] = (r(B, U, max(U - 1, 0)) < near_p).float()
    adj = r(B, U, U) < comm_p
    adj = adj | th.eye(U, dtype=th.bool).unsqueeze(0)            # d_u2u[i,i] = 0 <= r_comm: self-loops always
    return agent.to(device), gt.to(device), ubs.to(device), adj.to(device)
