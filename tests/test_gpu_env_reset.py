"""On-device episode reset (``ubs_env_sample_layouts``): the layouts must follow the distribution of the reference's
``DenseHotSpot / HotSpot / Map.set_positions`` + ``np.random.permutation`` (``envs/mubs_cov/maps.py:30-34,64-77,97-113``,
``envs/common.py:13-16``, ``mubs_cov.py:96-98``).  The RNG-matched host sampler ``envs.sample_layouts`` (pinned bit for
bit against the reference in tests/test_env.py) is the yardstick: structural invariants must hold exactly, marginal
statistics must agree within sampling error."""
import numpy as np
import pytest
import torch as th

from uav_bs_ctrl_b200 import envs as E

pytestmark = pytest.mark.gpu


def _device_layouts(m, B, episodes, seed=7):
    env = E.MultiUbsCoverageVecEnv(n_envs=B, device="cuda", map=m)
    env.seed = seed
    arena_layout = env.new_layout()
    from uav_bs_ctrl_b200.arena import SequenceArena
    ar = SequenceArena(arena_layout, 2, 8, "cuda")
    import ctypes as C
    from uav_bs_ctrl_b200 import _lib
    pu, pg, pr = [], [], []
    for ep in range(episodes):
        # the sampler alone (env.reset = this + ubs_env_reset, whose t = 0 transmit re-sorts the priorities)
        _lib.check(env._lib.ubs_env_sample_layouts(C.byref(env.layout_cfg), C.byref(env._state), int(env.seed), ep, B,
                                                   _lib.stream()), "ubs_env_sample_layouts")
        pu.append(env.buf.pos_ubs.cpu().numpy().copy())
        pg.append(env.buf.pos_gts.cpu().numpy().copy())
        pr.append(env.buf.prior.cpu().numpy().copy())
    return np.concatenate(pu), np.concatenate(pg), np.concatenate(pr), env, ar


def _chi2_uniform(counts):
    exp = counts.sum() / counts.size
    return float(((counts - exp) ** 2 / exp).sum()), counts.size - 1


def test_dense_hotspot_layouts_follow_the_reference_distribution():
    m = E.DenseHotSpot(n_ubs=8, n_grps=16)                 # BASELINE exp3: 8 UBS x 80 GT
    B, EP = 256, 8
    pu, pg, pr, env, ar = _device_layouts(m, B, EP)
    n = B * EP
    # --- UBS: distinct points of the 200 m grid (select_from_cube(n_ubs, 0, range_pos // 200, 2) * 200)
    assert np.all(pu % 200 == 0) and pu.min() >= 0 and pu.max() <= 5800
    cells = (pu[..., 0] // 200 * 30 + pu[..., 1] // 200).astype(int)
    assert all(len(set(c)) == 8 for c in cells), "UBS positions must be distinct within an instance"
    chi, dof = _chi2_uniform(np.bincount((pu[..., 0] // 200).astype(int).ravel(), minlength=30).astype(float))
    assert chi < dof + 6 * np.sqrt(2 * dof), f"UBS x-coordinate not uniform over the grid (chi2 {chi:.1f}, dof {dof})"
    # --- GTs: 16 groups of 5 around distinct cells of ONE 4 x 4 hotspot block aligned to 800 m, jitter < r_cov / 2
    centre = np.round(pg / 200.0) * 200.0
    assert np.abs(pg - centre).max() <= 50.0 + 1e-3, "GT = group centre + r_cov * (rand - 0.5)"
    for b in range(0, n, 37):
        c = {tuple(x) for x in centre[b]}
        assert len(c) == 16, "16 distinct group centres"
        cnt = np.unique(centre[b], axis=0, return_counts=True)[1]
        assert np.all(cnt == 5), "5 GTs per group"
        lo = centre[b].min(0)
        assert np.all(lo % 800 == 0) and np.all(centre[b].max(0) - lo == 600), "groups fill one 4 x 4 block"
    # --- statistics vs the RNG-matched host sampler (the reference's own draws)
    hu, hg, hp = E.sample_layouts(m, range(1000, 1000 + 1024))
    for name, dev_, host in (("ubs", pu, hu), ("gt", pg, hg)):
        d, h = dev_.reshape(-1, 2).astype(np.float64), host.reshape(-1, 2).astype(np.float64)
        se = h.std(0) / np.sqrt(min(len(d), len(h)) / (16 if name == "gt" else 1))      # GTs of an instance are correlated
        assert np.all(np.abs(d.mean(0) - h.mean(0)) < 6 * se + 1.0), f"{name}: mean {d.mean(0)} vs reference {h.mean(0)}"
        assert np.all(np.abs(d.std(0) / h.std(0) - 1) < 0.08), f"{name}: std {d.std(0)} vs reference {h.std(0)}"
    jit_d, jit_h = (pg - centre).ravel(), (hg - np.round(hg / 200.0) * 200.0).ravel()
    assert abs(jit_d.std() - jit_h.std()) < 0.5 and abs(jit_d.mean()) < 0.5          # uniform(-50, 50): std 28.87
    # --- np.random.shuffle(pos_gts): the group of slot 0 is uniform over the groups; the rows are not left in group order
    same = (np.abs(centre[:, 0] - centre[:, 1]).sum(-1) == 0).mean()
    assert abs(same - 4 / 79) < 0.02, f"P(slots 0 and 1 share a group) = {same:.3f}, expected 4/79"
    # --- prior_gts = permutation(n_gts)
    assert np.all(np.sort(pr, 1) == np.arange(80))
    chi, dof = _chi2_uniform(np.bincount(pr[:, 0], minlength=80).astype(float))
    assert chi < dof + 6 * np.sqrt(2 * dof), "first priority not uniform"
    # --- determinism / episode counter
    pu2, pg2, pr2, _, _ = _device_layouts(m, B, 2)
    assert np.array_equal(pu2, pu[:2 * B]) and np.array_equal(pg2, pg[:2 * B]) and np.array_equal(pr2, pr[:2 * B])
    assert not np.array_equal(pg[:B], pg[B:2 * B]), "a new episode draws a new layout"


def test_hotspot_and_grid_maps():
    m = E.HotSpot(r_comm=400.)                             # exp2 map 'r400': 4 UBS, 4 GTs on distinct cells of a 2 x 2 block
    pu, pg, pr, _, _ = _device_layouts(m, 128, 4)
    assert np.all(pu % 200 == 0) and pu.max() <= 1800 and np.all(pg % 200 == 0)
    for b in range(len(pg)):
        assert len({tuple(x) for x in pg[b]}) == 4 and len({tuple(x) for x in pu[b]}) == 4
        lo = pg[b].min(0)
        assert np.all(lo % 400 == 0) and np.all(pg[b].max(0) - lo == 200)
    assert np.all(np.sort(pr, 1) == np.arange(4))
    m0 = E.Map(n_ubs=3, n_gts=6)                           # base Map: uniform distinct integer grid points
    pu, pg, pr, _, _ = _device_layouts(m0, 64, 2)
    assert np.all(pu == np.round(pu)) and pu.max() <= 499 and np.all(pg == np.round(pg))
    assert all(len({tuple(x) for x in pg[b]}) == 6 for b in range(len(pg)))


def test_device_reset_is_two_launches_and_fast():
    m = E.DenseHotSpot(n_ubs=8, n_grps=16)
    _, _, _, env, ar = _device_layouts(m, 256, 1)
    for _ in range(3):
        env.reset(ar, 0)
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    from uav_bs_ctrl_b200 import _lib
    n0 = _lib.launch_count()
    e0.record()
    for _ in range(20):
        env.reset(ar, 0)
    e1.record()
    th.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 20
    per_reset = (_lib.launch_count() - n0) / 20
    print(f"device reset of 256 envs: {us:.1f} us ({per_reset:.0f} launches)")
    assert per_reset <= 3 and us < 150.0


def test_debug_map_has_no_device_sampler():
    env = E.MultiUbsCoverageVecEnv(n_envs=2, device="cuda", map=E.Debug())
    from uav_bs_ctrl_b200.arena import SequenceArena
    ar = SequenceArena(env.new_layout(), 2, 8, "cuda")
    env.reset(ar, 0)                                        # falls back to the (deterministic) host layout
    with pytest.raises(ValueError):
        env.reset(ar, 0, device=True)
