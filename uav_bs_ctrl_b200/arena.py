"""Packed observation packets and the device-resident sequence arena.

The reference keeps its replay as Python lists of per-step DGL graph objects on the host and re-uploads /
re-batches them at every update (``algos/madrqn/buffer.py:7-42``, ``learner.py:99-116``).  Here one timestep of B
env instances is ONE fixed-layout buffer of 4-byte words (a *packet*)

    [ x_agent (N, F_ag) | indptr_seen (N+1) | indptr_near (N+1) | talk mask (N) | reward (N) | done (B) | bad_mask (B) |
      (state) | (flattened observations) | x_ubs (N·(U-1), F_ubs) | x_gt (N·G, F_gt) ]    every section on a 128-byte line

so that (i) host → device staging of an observation is a single ``cudaMemcpyAsync`` from pinned memory of the packet's
live prefix (fixed-size sections first, the ``seen`` rows last: rows past ``indptr_seen[-1]`` are never shipped), (ii) a
*sequence arena* — ``(T+1)`` packets at a fixed stride in HBM plus the hidden states and actions — IS the replay
entry, and (iii) the strided-segment kernels (``ubs_gatv2_seg_*``) encode all T+1 timesteps in one launch straight
from the arena: no graph objects, no re-batching, no copies between ``act`` and ``update``.

Star layout inside a packet: visible GT / UBS rows are compacted in (env, agent, slot) order, ``indptr`` is the
cumulative degree — exactly the node / edge order of the reference's ``dgl.batch`` of per-agent graphs
(``env_wrappers.py:65-89``); ``reward`` / ``done`` / ``bad_mask`` belong to the transition that LED to the observation.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch as th

from .graph import HeteroGraph, RelCSR
from .builder import OBS_CETS


def _al4(n: int) -> int:
    return (n + 3) & ~3


class PacketLayout:
    """Word offsets of the sections of one observation packet for ``B`` envs × ``U`` agents × ``G`` ground terminals."""

    FLOAT = ("x_gt", "x_ubs", "x_agent", "rew", "done", "bad", "state", "x_flat")

    def __init__(self, B: int, U: int, G: int, F_ag: int = 2, F_gt: int = 4, F_ubs: int = 2, state_dim: int = 0,
                 flat_dim: int = 0):
        self.B, self.U, self.G, self.N = B, U, G, B * U
        self.F_ag, self.F_gt, self.F_ubs = F_ag, F_gt, F_ubs
        N = self.N
        self.cap_gt, self.cap_ubs = N * G, N * max(U - 1, 0)
        # fixed-size sections first, the two CSR row stores last: a packet's live words are one prefix
        # (``used_words``) — everything up to the last `seen` row — so staging copies the rows an observation has,
        # not the N*G-row capacity
        sizes = [("x_agent", N * F_ag), ("ip_seen", N + 1), ("ip_near", N + 1), ("mask", N), ("rew", N), ("done", B),
                 ("bad", B)]
        self.state_dim = state_dim                                   # global env state (QMIX), optional
        if state_dim:
            sizes.append(("state", B * state_dim))
        # flattened local observations for the MLP encoder (optional): rows padded to a multiple of 32 floats so that
        # the first encoder layer can run on the tcgen05 projection kernel (K % 32 == 0); the pad columns are zero
        self.flat_dim, self.flat_ld = flat_dim, (flat_dim + 31) // 32 * 32
        if flat_dim:
            sizes.append(("x_flat", N * self.flat_ld))
        sizes += [("x_ubs", self.cap_ubs * F_ubs), ("x_gt", self.cap_gt * F_gt)]
        self.off: Dict[str, int] = {}
        self.size: Dict[str, int] = {}
        o = 0
        for name, n in sizes:
            self.off[name], self.size[name] = o, n
            o += (n + 31) // 32 * 32                  # every section starts on a 128-byte line (slots too: words % 32 == 0)
        self.words = o

    def section(self, buf: th.Tensor, name: str) -> th.Tensor:
        """Typed view of a section of ``buf`` (``(..., words)`` int32): float32 for feature / reward sections."""
        v = buf[..., self.off[name]:self.off[name] + self.size[name]]
        return v.view(th.float32) if name in self.FLOAT else v

    def used_words(self, n_seen: int) -> int:
        """Length of the live prefix of a packet whose `seen` relation has ``n_seen`` edges (16-byte granular)."""
        return min(self.words, _al4(self.off["x_gt"] + n_seen * self.F_gt))

    def key(self):
        return (self.B, self.U, self.G, self.F_ag, self.F_gt, self.F_ubs, self.state_dim, self.flat_dim)


class ObsPacket:
    """One packet (host — optionally pinned — or device)."""

    def __init__(self, layout: PacketLayout, device="cpu", pin: bool = False, buf: Optional[th.Tensor] = None):
        self.layout = layout
        if buf is None:
            buf = th.zeros(layout.words, dtype=th.int32, device=device)
            if pin and buf.device.type == "cpu":
                buf = buf.pin_memory()
        self.buf = buf
        self._used = None

    def sec(self, name):
        return self.layout.section(self.buf, name)

    def used_words(self) -> int:
        """Words of the live prefix (header + CSR rows in use); cached — call ``touch()`` after rewriting the packet."""
        if self._used is None:
            if self.buf.device.type != "cpu":
                return self.layout.words                      # a device packet: no host read-back just to size a copy
            self._used = self.layout.used_words(int(self.sec("ip_seen")[self.layout.N]))
        return self._used

    def touch(self):
        self._used = None
        return self

    def fill_from_dense(self, agent_obs, gt_obs, ubs_obs, comm_adj=None, rew=None, done=None, bad=None, state=None):
        """Dense env observations (``envs/mubs_cov/mubs_cov.py:215-242`` format, see ``builder.py``) → packet."""
        L = self.layout
        B, U, N = L.B, L.U, L.N
        gflag = gt_obs[..., 0].reshape(N, L.G) == 1
        x_gt = gt_obs[..., 1:].reshape(N, L.G, L.F_gt)[gflag]
        deg = gflag.sum(1)
        self.sec("x_gt")[:x_gt.numel()] = x_gt.reshape(-1)
        ip = self.sec("ip_seen")
        ip[0] = 0
        ip[1:] = th.cumsum(deg, 0).to(th.int32)
        if U > 1:
            uflag = ubs_obs[..., 0].reshape(N, U - 1) == 1
            x_ubs = ubs_obs[..., 1:].reshape(N, U - 1, L.F_ubs)[uflag]
            self.sec("x_ubs")[:x_ubs.numel()] = x_ubs.reshape(-1)
            udeg = uflag.sum(1)
        else:
            udeg = th.zeros(N, dtype=th.int64)
        ipn = self.sec("ip_near")
        ipn[0] = 0
        ipn[1:] = th.cumsum(udeg, 0).to(th.int32)
        self.sec("x_agent")[:] = agent_obs.reshape(-1)
        if comm_adj is not None:
            m = (comm_adj.to(th.int64) << th.arange(U, device=comm_adj.device).view(1, U, 1)).sum(1).flatten()
            self.sec("mask")[:] = m.to(th.int32)
        else:
            self.sec("mask").zero_()
        self.sec("rew")[:] = 0 if rew is None else rew.reshape(-1).float()
        self.sec("done")[:] = 0 if done is None else done.reshape(-1).float()
        self.sec("bad")[:] = 0 if bad is None else bad.reshape(-1).float()
        if L.state_dim:
            self.sec("state")[:] = 0 if state is None else state.reshape(-1).float()
        self._used = None
        return self

    def to_graph(self) -> HeteroGraph:
        return packet_graph(self.layout, self.buf)


def packet_graph(L: PacketLayout, buf: th.Tensor) -> HeteroGraph:
    """HeteroGraph view (no copies of the big sections) of one packet — the same object ``builder.build_obs_graph_batch``
    would have produced; used for API compatibility and tests (the fast path never builds graph objects)."""
    N, U, B = L.N, L.U, L.B
    ip_s, ip_n = L.section(buf, "ip_seen"), L.section(buf, "ip_near")
    E_gt, E_ubs = int(ip_s[-1]), int(ip_n[-1])
    x_gt = L.section(buf, "x_gt")[:E_gt * L.F_gt].view(E_gt, L.F_gt)
    x_ubs = L.section(buf, "x_ubs")[:E_ubs * L.F_ubs].view(E_ubs, L.F_ubs)
    x_ag = L.section(buf, "x_agent").view(N, L.F_ag)
    mask = L.section(buf, "mask")
    dev = buf.device
    c_talk, c_seen, c_near = OBS_CETS
    bits = (mask.to(th.int64).view(N, 1) >> th.arange(U, device=dev).view(1, U)) & 1            # [dst, local src]
    dstv, srcl = th.nonzero(bits, as_tuple=True)
    src_idx = ((dstv // U) * U + srcl).to(th.int32)
    ip_t = th.zeros(N + 1, dtype=th.int64, device=dev)
    th.cumsum(bits.sum(1), 0, out=ip_t[1:])
    E_t = int(src_idx.numel())
    # edge ids in the reference's (env, src, dst) order
    key = (src_idx.to(th.int64) * N + dstv)
    eid = th.argsort(th.argsort(key)) if E_t else th.zeros(0, dtype=th.int64, device=dev)
    csr = {c_seen: RelCSR(ip_s, None, None, E_gt, N, E_gt), c_near: RelCSR(ip_n, None, None, E_ubs, N, E_ubs),
           c_talk: RelCSR(ip_t.to(th.int32), src_idx, eid, N, N, E_t, U if U <= 32 else None, mask if U <= 32 else None)}
    per_env = lambda ip: (ip[U::U] - ip[:-1:U]).tolist()
    bne_t = bits.view(B, U * U).sum(1).tolist()
    return HeteroGraph(("agent", "gt", "ubs"), OBS_CETS, {"agent": N, "gt": E_gt, "ubs": E_ubs},
                       {c: None for c in OBS_CETS}, {c: None for c in OBS_CETS},
                       {"agent": {"feat": x_ag}, "gt": {"feat": x_gt}, "ubs": {"feat": x_ubs}}, None,
                       {"agent": [U] * B, "gt": per_env(ip_s), "ubs": per_env(ip_n)},
                       {c_talk: bne_t, c_seen: per_env(ip_s), c_near: per_env(ip_n)}, csr)


class SequenceArena:
    """``S`` packets at a fixed stride on the device + per-step hidden states, actions and Q values.

    ``buf (S, words) int32``; ``h (S+1, N, H)``: ``h[t]`` is the hidden state ENTERING step t (``h[t+1]`` is written by
    the act step on slot t); ``acts (S, N) int64``."""

    def __init__(self, layout: PacketLayout, n_slots: int, hidden: int, device):
        self.layout, self.S, self.H = layout, n_slots, hidden
        self.device = th.device(device)
        self.buf = th.zeros(n_slots, layout.words, dtype=th.int32, device=self.device)
        self.h = th.zeros(n_slots + 1, layout.N, hidden, dtype=th.float32, device=self.device)
        self.acts = th.zeros(n_slots, layout.N, dtype=th.int64, device=self.device)

    def sec(self, name, t=None):
        return self.layout.section(self.buf if t is None else self.buf[t], name)

    def ptr(self, name, t=0) -> int:
        return self.buf.data_ptr() + 4 * (t * self.layout.words + self.layout.off[name])

    def load(self, t: int, packet: ObsPacket, compact: bool = True) -> int:
        """Stages a packet into slot t: ONE asynchronous copy (H2D when the packet is pinned host memory) of the
        packet's live prefix — the rows past ``ip_seen[-1]`` are never read (every kernel walks the CSR row pointers),
        so they are not shipped.  Returns the bytes copied."""
        n = packet.used_words() if compact else self.layout.words
        self.buf[t, :n].copy_(packet.buf[:n], non_blocking=True)
        return n * 4

    def graph(self, t: int) -> HeteroGraph:
        return packet_graph(self.layout, self.buf[t])

    def rewards(self, T: int) -> th.Tensor:
        """``(T, B, U)`` rewards of transitions 0..T-1 (stored with observations 1..T)."""
        return self.sec("rew")[1:T + 1].reshape(T, self.layout.B, self.layout.U)

    def flat_obs(self, t0: int, T: int) -> th.Tensor:
        """``(T, N, flat_dim)`` flattened local observations of slots ``t0 .. t0+T-1`` (a strided view, pad cut off)."""
        L = self.layout
        if not L.flat_dim:
            raise ValueError("this arena was created without flattened observations (PacketLayout(flat_dim=...))")
        return self.sec("x_flat")[t0:t0 + T].view(T, L.N, L.flat_ld)[..., :L.flat_dim]

    def states(self, T: int) -> th.Tensor:
        """``(T, B, state_dim)`` global states of slots 0..T-1 (QMIX)."""
        if not self.layout.state_dim:
            raise ValueError("this arena was created without a state section (PacketLayout(state_dim=...))")
        return self.sec("state")[:T].reshape(T, self.layout.B, self.layout.state_dim)

    def dones(self, T: int) -> th.Tensor:
        """``(T, B, 1)``: ``(1 - bad_mask) * done`` (reference ``learner.py:91``)."""
        d, b = self.sec("done")[1:T + 1], self.sec("bad")[1:T + 1]
        return ((1 - b) * d).reshape(T, self.layout.B, 1)


class ArenaReplay:
    """Device-resident sequence replay: a ring of ``capacity`` sequence arenas (reference ``ReplayBuffer``,
    ``algos/madrqn/buffer.py:7-42``, whose entries are Python lists of host graph objects).

    One arena = one finished window of ``max_seq_len`` transitions of the learner's ``n_envs`` env instances, i.e.
    ``n_envs`` of the reference's sequences; it never leaves HBM.  ``sample(batch_size)`` draws ``batch_size`` stored
    windows uniformly without replacement (``random.sample``, ``buffer.py:37-39``) and returns them as a list — pointer
    selection, no copies; ``learner.update_arena(list)`` encodes every selected window in place with the
    strided-segment kernels and runs the recurrent window kernels over all their agent rows at once.
    HBM budget: ``capacity`` x arena bytes (142 MB per exp3 window of 256 envs: ~1 000 windows fit 180 GB)."""

    def __init__(self, make_arena, capacity: int):
        self._make, self.capacity = make_arena, int(capacity)
        self._ring, self._filled, self._next = [], 0, 0

    def next_arena(self) -> "SequenceArena":
        """The arena the next window is collected into (the oldest stored window once the ring is full)."""
        if len(self._ring) < self.capacity and self._next == len(self._ring):
            self._ring.append(self._make())
        return self._ring[self._next]

    def commit(self):
        """The window in ``next_arena()`` is complete (``buffer.py:30-35``: the sequence is appended to memory)."""
        self._next = (self._next + 1) % self.capacity
        self._filled = min(self._filled + 1, self.capacity)

    def __len__(self):
        return self._filled

    def windows(self):
        """Stored windows, oldest first (the order of the reference's ``deque``)."""
        if self._filled < self.capacity:
            return self._ring[:self._filled]
        return self._ring[self._next:] + self._ring[:self._next]

    def sample(self, batch_size: int):
        import random
        return random.sample(self.windows(), batch_size)

    def nbytes(self) -> int:
        return sum(a.buf.numel() * 4 + a.h.numel() * 4 + a.acts.numel() * 8 for a in self._ring)
