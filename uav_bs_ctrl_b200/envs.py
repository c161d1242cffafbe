"""Device-resident, vectorised ``MultiUbsCoverageEnv`` behind the C ABI of
``include/ubs_env.h``.

Reference: ``envs/mubs_cov/mubs_cov.py`` (env), ``envs/mubs_cov/maps.py`` (maps), ``envs/common.py`` (channel model),
``algos/madrqn/utils/env_wrappers.py`` (observation graphs).  ``B`` env instances live on the GPU; ``step`` reads the
actions the fused act kernel wrote into a sequence arena and writes the next observation *packet* (compacted star
graphs + talk mask + reward / done / bad-mask) into the next arena slot — no host round trip, no graph objects.

Host side (this file): the static parameters of a map and every derived constant, computed with the reference's own
expressions (bit-identical doubles), and the RNG-matched initial layouts of ``Map.set_positions`` (python ``random`` +
``numpy.random`` legacy streams, consumed in the reference's order).  There is no CPU fallback: the device entry
points raise if ``libubs_gnn.so`` is missing or the tensors are not CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import random as _pyrandom
from typing import Optional, Sequence

import numpy as np
import torch as th

MAX_ACTIONS, MAX_UBS, INFO = 33, 32, 8
INFO_KEYS = ("EpRet", "TotalThroughput", "NColls", "AvgGlobalUtility", "FairIdx", "GlobalUtil", "EpLen")


class UbsEnvCfg(C.Structure):
    """``ubs_env_cfg`` of ``include/ubs_env.h``."""
    _fields_ = [(n, C.c_int32) for n in ("n_ubs", "n_gts", "n_rbs", "n_actions", "episode_limit", "fair_service",
                                         "avoid_collision", "gts_f64")] + \
               [(n, C.c_double) for n in ("range_pos", "r_cov", "r_sns", "r_comm", "dt", "rew_scale", "h_ubs", "p_tx",
                                          "n0", "bw", "c_fspl", "chan_a", "chan_b", "k_los", "k_nlos", "max_rate",
                                          "safe_dist", "penalty")] + \
               [("moves", (C.c_double * 2) * MAX_ACTIONS)]


class UbsEnvState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("pos_ubs", "pos_gts", "avg_rate", "rate", "prior", "t", "info", "sched")]


class UbsEnvPacket(C.Structure):
    _fields_ = [("packet", C.c_void_p)] + [(n, C.c_int64) for n in ("off_x_gt", "off_x_ubs", "off_x_agent",
                                                                   "off_ip_seen", "off_ip_near", "off_mask", "off_rew",
                                                                   "off_done", "off_bad", "off_state", "off_flat",
                                                                   "ld_flat")]


class UbsEnvLayoutCfg(C.Structure):
    """``ubs_env_layout_cfg`` of ``include/ubs_env.h``: the map's reset distribution for the device sampler."""
    _fields_ = [(n, C.c_int32) for n in ("kind", "n_ubs", "n_gts", "n_grps", "gts_per_grp", "ubs_cells", "range_spot",
                                         "spot_cells")] + [(n, C.c_double) for n in ("min_dist", "range_pos", "r_cov")]


# ------------------------------------------------------------------------------------------------------ maps
def _select_from_cube(rng: _pyrandom.Random, n_els, min_val, max_val, n_dims=2):
    """``envs/common.py:13-16``: ``random.sample`` over ``product(arange(min, max), ...)``.  ``random.sample`` only looks
    at ``len(population)``, so sampling indices consumes the RNG identically; the product is row-major."""
    axis = np.arange(min_val, max_val)
    side = len(axis)
    idx = rng.sample(range(side ** n_dims), n_els)
    pts = np.empty((n_els, n_dims), dtype=axis.dtype)
    for k, i in enumerate(idx):
        for d in range(n_dims - 1, -1, -1):
            pts[k, d] = axis[i % side]
            i //= side
    return pts


class Map:
    """Parameters + initial layout of a scenario (``envs/mubs_cov/maps.py:4-34``)."""

    def __init__(self, range_pos=500, episode_limit=20, dt=10, n_ubs=1, n_gts=1, r_cov=100., n_rbs=1, r_sns=np.inf,
                 r_comm=np.inf, vels=10, n_dirs=4, rew_scale=1.):
        self.range_pos, self.episode_limit, self.dt = range_pos, episode_limit, dt
        self.n_ubs, self.n_gts, self.r_cov, self.n_rbs = n_ubs, n_gts, r_cov, n_rbs
        self.r_sns, self.r_comm, self.vels, self.n_dirs = r_sns, r_comm, vels, n_dirs
        self.reward_scale_rate = rew_scale

    def set_positions(self, py_rng, np_rng):
        pos_ubs = _select_from_cube(py_rng, self.n_ubs, 0, self.range_pos)
        pos_gts = _select_from_cube(py_rng, self.n_gts, 0, self.range_pos)
        return dict(ubs=pos_ubs, gt=pos_gts)


class Debug(Map):
    """``maps.py:37-50``."""

    def __init__(self, range_pos=1000, episode_limit=10, dt=10, n_ubs=3, n_gts=4, r_cov=100., n_rbs=1, r_sns=300.,
                 r_comm=np.inf, vels=10., n_dirs=4, rew_scale=1.):
        super().__init__(range_pos, episode_limit, dt, n_ubs, n_gts, r_cov, n_rbs, r_sns, r_comm, vels, n_dirs, rew_scale)

    def set_positions(self, py_rng, np_rng):
        return dict(ubs=100 * np.array([[3, 3], [8, 2], [8, 9]], dtype=np.float32),
                    gt=100 * np.array([[3, 4], [4, 2], [3, 1], [6, 9]], dtype=np.float32))


class HotSpot(Map):
    """``maps.py:56-76``."""

    def __init__(self, range_pos=2000, episode_limit=40, dt=20, n_ubs=4, n_gts=4, r_cov=100., n_rbs=1, r_sns=200.,
                 r_comm=np.inf, vels=(5, 10), n_dirs=4, rew_scale=10.):
        super().__init__(range_pos, episode_limit, dt, n_ubs, n_gts, r_cov, n_rbs, r_sns, r_comm, list(vels), n_dirs,
                         rew_scale)

    def set_positions(self, py_rng, np_rng):
        min_dist = 200.
        pos_ubs = min_dist * _select_from_cube(py_rng, self.n_ubs, 0, self.range_pos // min_dist)
        range_spot = 1
        while np.square(range_spot) < self.n_gts:
            range_spot += 1
        pos_spot = min_dist * range_spot * _select_from_cube(py_rng, 1, 0, self.range_pos // min_dist // range_spot)
        pos_gts = pos_spot + min_dist * _select_from_cube(py_rng, self.n_gts, 0, range_spot)
        pos_gts = np.clip(pos_gts, 0, self.range_pos)
        np_rng.shuffle(pos_gts)
        return dict(ubs=pos_ubs, gt=pos_gts)


class DenseHotSpot(Map):
    """``maps.py:83-113`` (experiment 3)."""

    def __init__(self, range_pos=6000, episode_limit=50, dt=40, n_ubs=4, n_grps=10, gts_per_grp=5, r_cov=100., n_rbs=5,
                 r_sns=400., r_comm=np.inf, vels=(5, 10), n_dirs=4, rew_scale=10):
        super().__init__(range_pos, episode_limit, dt, n_ubs, n_grps * gts_per_grp, r_cov, n_rbs, r_sns, r_comm,
                         list(vels), n_dirs, rew_scale)
        self.n_grps, self.gts_per_grp = n_grps, gts_per_grp

    def set_positions(self, py_rng, np_rng):
        min_dist = 200.
        pos_ubs = min_dist * _select_from_cube(py_rng, self.n_ubs, 0, self.range_pos // min_dist)
        range_spot = 1
        while np.square(range_spot) < self.n_grps:
            range_spot += 1
        pos_spot = min_dist * range_spot * _select_from_cube(py_rng, 1, 0, self.range_pos // min_dist // range_spot)
        pos_grps = pos_spot + min_dist * _select_from_cube(py_rng, self.n_grps, 0, range_spot)
        pos_gts = np.empty((self.n_gts, 2), dtype=np.float32)
        for g in range(self.n_grps):
            sl = slice(g * self.gts_per_grp, (g + 1) * self.gts_per_grp)
            pos_gts[sl] = pos_grps[g] + self.r_cov * (np_rng.rand(self.gts_per_grp, 2) - 0.5)
        pos_gts = np.clip(pos_gts, 0, self.range_pos)
        np_rng.shuffle(pos_gts)
        return dict(ubs=pos_ubs, gt=pos_gts)


class DenseHotSpotV2(Map):
    """``maps.py:117-133``."""

    def __init__(self, range_pos=6000., episode_limit=100, dt=10, n_ubs=4, n_gts=100, r_cov=100., n_rbs=10, r_sns=400,
                 r_comm=np.inf, vels=(5., 10.), n_dirs=4, rew_scale=10):
        super().__init__(range_pos, episode_limit, dt, n_ubs, n_gts, r_cov, n_rbs, r_sns, r_comm, list(vels), n_dirs,
                         rew_scale)

    def set_positions(self, py_rng, np_rng):
        pos_ubs = 100 * _select_from_cube(py_rng, self.n_ubs, 0, self.range_pos // 100)
        radius_spot = 400
        pos_spot = radius_spot * _select_from_cube(py_rng, 1, 1, self.range_pos // radius_spot)
        pos_gts = pos_spot + radius_spot * 2 * (np_rng.rand(self.n_gts, 2) - 0.5)
        pos_gts = np.clip(pos_gts, 0, self.range_pos)
        np_rng.shuffle(pos_gts)
        return dict(ubs=pos_ubs, gt=pos_gts)


def make_maps():
    """The reference registry (``maps.py:139-153``) + the BASELINE.json exp3 shape (8 UBS x 80 GT = 16 groups of 5)."""
    return {"test": Map(), "debug": Debug(), "inf": HotSpot(), "r400": HotSpot(r_comm=400.), "r800": HotSpot(r_comm=800.),
            "4ubs": DenseHotSpot(n_ubs=4), "6ubs": DenseHotSpot(n_ubs=6), "8ubs": DenseHotSpot(n_ubs=8),
            "8ubs80": DenseHotSpot(n_ubs=8, n_grps=16), "16ubs320": DenseHotSpot(n_ubs=16, n_grps=64)}


# physical constants of MultiUbsCoverageEnv (mubs_cov.py:14-21) and the channel table (common.py:33-38)
H_UBS, BW, FC, SCENE, SAFE_DIST, PENALTY = 100., 180e3, 2.4e9, "dense-urban", 10., 5
P_TX = 1e-3 * np.power(10, 10 / 10)
N0 = 1e-3 * np.power(10, -170 / 10)
CHAN_PARAMS = {"suburban": (4.88, 0.43, 0.1, 21), "urban": (9.61, 0.16, 1, 20), "dense-urban": (12.08, 0.11, 1.6, 23),
               "high-rise-urban": (27.23, 0.08, 2.3, 34)}


def _chan_gain(d_level, h_ubs, a, b, eta_los, eta_nlos, fc):
    """``AirToGroundChannel.estimate_chan_gain`` (``common.py:45-55``), same expression."""
    p_los = 1 / (1 + a * np.exp(-b * (np.arctan(h_ubs / (d_level + 1e-5)) - a)))
    d = np.sqrt(np.square(d_level) + np.square(h_ubs))
    fspl = (4 * np.pi * fc * d / 3e8) ** 2
    pl = p_los * fspl * 10 ** (eta_los / 20) + (1 - p_los) * fspl * 10 ** (eta_nlos / 20)
    return 1 / pl


def avail_moves(m: Map) -> np.ndarray:
    """``mubs_cov.py:61-65``."""
    move_amounts = m.dt * np.array(m.vels).reshape(-1, 1)
    ang = 2 * np.pi * np.arange(m.n_dirs) / m.n_dirs
    move_dirs = np.stack([np.cos(ang), np.sin(ang)]).T
    return np.concatenate((np.zeros((1, 2)), np.kron(move_amounts, move_dirs)))


def make_cfg(m: Map, fair_service: bool = True, avoid_collision: bool = True) -> UbsEnvCfg:
    a, b, eta_los, eta_nlos = CHAN_PARAMS[SCENE]
    g_max = _chan_gain(0, H_UBS, a, b, eta_los, eta_nlos, FC)
    snr_max = P_TX * g_max / (N0 * BW)
    max_rate = BW * np.log2(1 + snr_max) * 1e-6                                    # mubs_cov.py:36-39
    mv = avail_moves(m)
    if mv.shape[0] > MAX_ACTIONS or m.n_ubs > MAX_UBS:
        raise ValueError("map outside the device env's limits (<= 33 actions, <= 32 UBSs)")
    c = UbsEnvCfg()
    c.n_ubs, c.n_gts, c.n_rbs, c.n_actions = int(m.n_ubs), int(m.n_gts), int(m.n_rbs), int(mv.shape[0])
    c.episode_limit, c.fair_service, c.avoid_collision = int(m.episode_limit), int(fair_service), int(avoid_collision)
    c.gts_f64 = int(np.asarray(m.set_positions(_pyrandom.Random(0), np.random.RandomState(0))["gt"]).dtype == np.float64)
    c.range_pos, c.r_cov, c.r_sns, c.r_comm = float(m.range_pos), float(m.r_cov), float(m.r_sns), float(m.r_comm)
    c.dt, c.rew_scale = float(m.dt), float(m.reward_scale_rate)
    c.h_ubs, c.p_tx, c.n0, c.bw = H_UBS, float(P_TX), float(N0), BW
    c.c_fspl = float(4 * np.pi * FC)
    c.chan_a, c.chan_b = a, b
    c.k_los, c.k_nlos = float(10 ** (eta_los / 20)), float(10 ** (eta_nlos / 20))
    c.max_rate = float(max_rate)
    c.safe_dist, c.penalty = SAFE_DIST, float(PENALTY)
    for i in range(mv.shape[0]):
        c.moves[i][0], c.moves[i][1] = float(mv[i, 0]), float(mv[i, 1])
    return c


def make_layout_cfg(m: "Map") -> Optional[UbsEnvLayoutCfg]:
    """Parameters of ``m.set_positions()`` for ``ubs_env_sample_layouts`` (``maps.py:30-34,64-77,97-113``), or ``None``
    for maps without a device sampler (``Debug``: fixed positions; ``DenseHotSpotV2``: not in the reference registry)."""
    c = UbsEnvLayoutCfg()
    c.n_ubs, c.n_gts, c.range_pos, c.r_cov, c.min_dist = int(m.n_ubs), int(m.n_gts), float(m.range_pos), float(m.r_cov), 200.0
    if type(m) is Map:
        c.kind, c.ubs_cells = 0, int(m.range_pos)
        return c
    if type(m) in (HotSpot, DenseHotSpot):
        dense = type(m) is DenseHotSpot
        n = m.n_grps if dense else m.n_gts
        range_spot = 1
        while range_spot * range_spot < n:                   # "ensure sufficient area to hold GTs" (maps.py:69-70,102-103)
            range_spot += 1
        c.kind = 2 if dense else 1
        c.n_grps, c.gts_per_grp = (int(m.n_grps), int(m.gts_per_grp)) if dense else (0, 0)
        c.ubs_cells = int(m.range_pos // c.min_dist)
        c.range_spot, c.spot_cells = range_spot, int(m.range_pos // c.min_dist // range_spot)
        return c
    return None


def flat_obs_dim(cfg: UbsEnvCfg) -> int:
    """2 own features + G rows of (flag, dx, dy, rate[, avg rate]) + (U-1) rows of (flag, dx, dy)  (``mubs_cov.py:247-278``)."""
    return 2 + cfg.n_gts * (5 if cfg.fair_service else 4) + (cfg.n_ubs - 1) * 3


def sample_layouts(m: Map, seeds: Sequence[int]):
    """RNG-matched ``reset`` draws for ``len(seeds)`` env instances: what the reference produces after
    ``random.seed(s); np.random.seed(s)`` — ``map.set_positions()`` then ``np.random.permutation(n_gts)``
    (``mubs_cov.py:96-98``).  Returns ``pos_ubs (B,U,2) f64``, ``pos_gts (B,G,2) f32``, ``prior (B,G) i32``."""
    pu, pg, pr = [], [], []
    for s in seeds:
        py_rng, np_rng = _pyrandom.Random(int(s)), np.random.RandomState(int(s))
        pos = m.set_positions(py_rng, np_rng)
        pu.append(np.asarray(pos["ubs"], dtype=np.float64))
        pg.append(np.asarray(pos["gt"], dtype=np.float32))
        pr.append(np_rng.permutation(m.n_gts).astype(np.int32))
    return np.stack(pu), np.stack(pg), np.stack(pr)


# ------------------------------------------------------------------------------------------------------ buffers
class EnvBuffers:
    """State tensors of ``B`` env instances + the scratch area, on ``device``, and the ctypes views of them."""

    def __init__(self, cfg: UbsEnvCfg, B: int, device, scratch_words: int):
        self.cfg, self.B, self.device = cfg, B, th.device(device)
        U, G = cfg.n_ubs, cfg.n_gts
        z = lambda *s, dt=th.float32: th.zeros(*s, dtype=dt, device=self.device)
        self.pos_ubs, self.pos_gts = z(B, U, 2, dt=th.float64), z(B, G, 2)
        self.avg_rate, self.rate = z(B, G), z(B, G)
        self.prior = th.arange(G, dtype=th.int32, device=self.device).repeat(B, 1).contiguous()
        self.t = z(B, dt=th.int32)
        self.info = z(B, INFO, dt=th.float64)
        self.sched = th.full((B, G, 2), -1, dtype=th.int32, device=self.device)
        self.scratch = z(max(int(scratch_words), 4), dt=th.int32)

    def state_struct(self) -> UbsEnvState:
        s = UbsEnvState()
        for n in ("pos_ubs", "pos_gts", "avg_rate", "rate", "prior", "t", "info", "sched"):
            setattr(s, n, getattr(self, n).data_ptr())
        return s

    def set_layout(self, pos_ubs, pos_gts, prior):
        def as_t(a, dt):
            if not isinstance(a, th.Tensor):
                a = th.as_tensor(np.ascontiguousarray(a))
            return a.to(device=self.device, dtype=dt, non_blocking=True)
        self.pos_ubs.copy_(as_t(pos_ubs, th.float64).view_as(self.pos_ubs))
        self.pos_gts.copy_(as_t(pos_gts, th.float32).view_as(self.pos_gts))
        self.prior.copy_(as_t(prior, th.int32).view_as(self.prior))

    _STATE = ("pos_ubs", "pos_gts", "avg_rate", "rate", "prior", "t", "info", "sched")

    def snapshot(self):
        return {n: getattr(self, n).clone() for n in self._STATE}

    def restore(self, snap):
        for n in self._STATE:
            getattr(self, n).copy_(snap[n])

    def info_dict(self, b: int = 0) -> dict:
        v = self.info[b].tolist()
        d = dict(zip(INFO_KEYS, v))
        d["ProbCollision"] = d["NColls"] / max(d["EpLen"], 1)
        return d


def packet_struct(layout, buf: th.Tensor) -> UbsEnvPacket:
    """``ubs_env_packet`` for one packet ``buf (words,) int32`` of ``arena.PacketLayout`` ``layout``."""
    p = UbsEnvPacket()
    p.packet = buf.data_ptr()
    o = layout.off
    p.off_x_gt, p.off_x_ubs, p.off_x_agent = o["x_gt"], o["x_ubs"], o["x_agent"]
    p.off_ip_seen, p.off_ip_near, p.off_mask = o["ip_seen"], o["ip_near"], o["mask"]
    p.off_rew, p.off_done, p.off_bad = o["rew"], o["done"], o["bad"]
    p.off_state = o.get("state", -1)
    p.off_flat, p.ld_flat = o.get("x_flat", -1), layout.flat_ld
    return p


class MultiUbsCoverageVecEnv:
    """``B`` instances of the reference's ``MultiUbsCoverageEnv(map_id, fair_service, avoid_collision)`` on one GPU.

    ``reset(arena, slot)`` / ``step(arena, t)`` mirror ``env.reset()`` / ``env.step(actions)`` for all instances at
    once; observations, rewards and termination flags go to arena packets (``reset`` -> slot, ``step`` reads
    ``arena.acts[t]`` and fills slot ``t+1``) instead of being returned as Python objects."""

    def __init__(self, map_id="8ubs", n_envs: int = 1, device="cuda", fair_service=True, avoid_collision=True,
                 map: Optional[Map] = None):
        from . import _lib
        self.map = map if map is not None else make_maps()[map_id]
        self.cfg = make_cfg(self.map, fair_service, avoid_collision)
        self.n_envs, self.device = n_envs, th.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("MultiUbsCoverageVecEnv runs on a CUDA device only (no CPU fallback)")
        if self.device.index is None:
            self.device = th.device("cuda", th.cuda.current_device())
        self._lib = _lib.load()
        words = int(self._lib.ubs_env_scratch_words(C.byref(self.cfg), n_envs))
        self.buf = EnvBuffers(self.cfg, n_envs, self.device, words)
        self._state = self.buf.state_struct()
        self.n_agents, self.n_actions = self.cfg.n_ubs, self.cfg.n_actions
        self.episode_limit = self.cfg.episode_limit
        self._episode = 0
        self.layout_cfg = make_layout_cfg(self.map)           # None: no device sampler for this map
        self.seed = 0                                          # key of the device sampler's Philox stream

    # reference-shaped metadata (env_wrappers.py:117-120, :62-63)
    def get_env_info(self, o="gnn"):
        """``o='gnn'``: graph observations (``GraphObservation.get_obs_size``); ``o='mlp'``: flattened observations."""
        if o == "mlp":
            return dict(obs_shape=self.flat_obs_dim, state_shape=self.state_dim, n_actions=self.n_actions,
                        n_agents=self.n_agents, episode_limit=self.episode_limit)
        return dict(obs_shape=dict(agent=2, ubs=2, gt=4 if self.cfg.fair_service else 3), state_shape=self.state_dim,
                    n_actions=self.n_actions, n_agents=self.n_agents, episode_limit=self.episode_limit)

    @property
    def state_dim(self) -> int:
        """``get_state_size()`` (``mubs_cov.py:264-265``)."""
        return 2 * self.cfg.n_ubs + (4 if self.cfg.fair_service else 3) * self.cfg.n_gts

    @property
    def flat_obs_dim(self) -> int:
        """Size of a flattened local observation (``FlattenedObservation.get_obs_size``, ``env_wrappers.py:48-49``)."""
        return flat_obs_dim(self.cfg)

    def new_layout(self, F_gt=None, with_state=False, with_flat=False):
        from .arena import PacketLayout
        return PacketLayout(self.n_envs, self.cfg.n_ubs, self.cfg.n_gts, 2, F_gt or (4 if self.cfg.fair_service else 3), 2,
                            state_dim=self.state_dim if with_state else 0, flat_dim=self.flat_obs_dim if with_flat else 0)

    def make_layout_pool(self, n_batches: int, seed0: int = 0):
        """``n_batches`` RNG-matched reset batches sampled ahead of time and parked on the device (the host sampler is
        python + numpy RNG code, ~0.3 ms per instance: far slower than the device loop it feeds)."""
        pool = []
        for k in range(n_batches):
            pu, pg, pr = sample_layouts(self.map, [seed0 + k * self.n_envs + b for b in range(self.n_envs)])
            pool.append(tuple(th.as_tensor(a).to(self.device) for a in (pu, pg, pr)))
        return pool

    def _check(self, arena):
        L = arena.layout
        if (L.B, L.U, L.G, L.F_gt) != (self.n_envs, self.cfg.n_ubs, self.cfg.n_gts, 4 if self.cfg.fair_service else 3):
            raise ValueError("arena layout does not match the env (B, U, G, F_gt)")
        if arena.buf.device != self.device:
            raise RuntimeError("arena and env must live on the same CUDA device")

    def reset(self, arena, slot: int = 0, seeds: Optional[Sequence[int]] = None, layouts=None, device: Optional[bool] = None):
        """New episode in every instance.  Where the initial layouts come from:

        * default (``device=None`` -> True when the map has a device sampler): ``ubs_env_sample_layouts`` draws them on
          the GPU from the map's reset distribution (Philox stream keyed by ``self.seed``, the instance and a running
          episode counter) — nothing touches the host, the call is two kernel launches;
        * ``seeds=[...]`` (or ``device=False``): the RNG-matched host sampler — ``seeds[b]`` plays the role of the
          reference's global seed for instance b (``random.seed(s); np.random.seed(s)``), bit-identical layouts to the
          reference's ``reset()`` but ~0.3 ms of python per instance;
        * ``layouts=(pos_ubs, pos_gts, prior)``: caller-provided."""
        from . import _lib
        self._check(arena)
        if device is None:
            device = layouts is None and seeds is None and self.layout_cfg is not None
        if device:
            if self.layout_cfg is None:
                raise ValueError(f"map {type(self.map).__name__} has no device sampler: pass seeds= or layouts=")
            _lib.check(self._lib.ubs_env_sample_layouts(C.byref(self.layout_cfg), C.byref(self._state), int(self.seed),
                                                        int(self._episode) & 0xFFFFFFFF, self.n_envs, _lib.stream()),
                       "ubs_env_sample_layouts")
        else:
            if layouts is None:
                if seeds is None:
                    seeds = [self._episode * self.n_envs + b for b in range(self.n_envs)]
                layouts = sample_layouts(self.map, seeds)
            self.buf.set_layout(*layouts)
        self._episode += 1
        pk = packet_struct(arena.layout, arena.buf[slot])
        _lib.check(self._lib.ubs_env_reset(C.byref(self.cfg), C.byref(self._state), C.byref(pk),
                                           self.buf.scratch.data_ptr(), self.n_envs, _lib.stream()), "ubs_env_reset")

    def step(self, arena, t: int, actions: Optional[th.Tensor] = None):
        """``env.step(arena.acts[t])`` for every instance -> observation / reward / done / bad-mask in slot ``t+1``."""
        from . import _lib
        self._check(arena)
        acts = arena.acts[t] if actions is None else actions
        if not acts.is_cuda or acts.dtype != th.int64 or acts.numel() != self.n_envs * self.n_agents:
            raise ValueError("actions must be a CUDA int64 tensor of B*U elements")
        from . import ops
        pk = packet_struct(arena.layout, arena.buf[t + 1])
        with ops._timed("env_step", (self.n_envs, self.cfg.n_ubs, self.cfg.n_gts, 4 if self.cfg.fair_service else 3)):
            _lib.check(self._lib.ubs_env_step(C.byref(self.cfg), C.byref(self._state), acts.data_ptr(), C.byref(pk),
                                              self.buf.scratch.data_ptr(), self.n_envs, _lib.stream()), "ubs_env_step")
