"""DGL-free heterogeneous graph container for the UBS hot path.

The reference builds its observations as ``dgl`` heterographs (reference
``algos/madrqn/utils/env_wrappers.py:65-89,139-154``, ``algos/drqn/utils/env_wrappers.py:63-77``)
and batches them with ``dgl.batch`` (``algos/common.py:40-47``).  DGL is not part of this
framework; this module provides exactly the graph-object surface those callers and the agent
modules touch: ``heterograph``, ``batch``, ``merge``, ``g[etype]``,
``g.ndata / nodes[nt].data / srcdata / dstdata / edata``, ``num_nodes``, ``number_of_edges``,
``local_scope``, ``to``, plus the small message-passing API used by the non-fused comm protocols
(``apply_edges``, ``update_all``, ``edge_softmax``; see ``function.py``).

What is different from DGL: every relation carries a **CSR-by-destination** view
(``indptr`` int32, optional ``src_idx`` int32, ``is_star`` flag) that the CUDA kernels consume
directly.  For the observation relations the CSR is free: the reference's per-agent star graphs
have ``src id == edge id`` and destination-sorted edges, so ``indptr = cumsum(degree)`` and no
gather index is needed.  ``batch`` and ``merge`` compose CSRs by
pointer arithmetic instead of re-sorting.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict, List, Optional, Sequence, Tuple

import torch as th

__all__ = ["HeteroGraph", "DGLGraph", "RelGraph", "RelCSR", "heterograph", "batch", "merge"]


def _as_index(x, device=None) -> th.Tensor:
    if isinstance(x, th.Tensor):
        t = x.to(th.int64)
    else:
        t = th.as_tensor(list(x) if not hasattr(x, "dtype") else x, dtype=th.int64)
    t = t.reshape(-1)
    return t if device is None else t.to(device)


class RelCSR:
    """CSR-by-destination of one relation.

    ``indptr``  int32 (n_dst+1,)  in-edge segment of destination v is ``[indptr[v], indptr[v+1])``.
    ``src_idx`` int32 (E,) source node of every CSR slot, or ``None`` for the *star layout*
                (slot j holds source node j; every source has exactly one out-edge).
    ``eid``     int64 (E,) original edge id of every CSR slot, or ``None`` when the edge list was
                already destination-sorted (slot j == edge j).
    """

    __slots__ = ("indptr", "src_idx", "eid", "n_src", "n_dst", "n_edges", "block", "mask")

    def __init__(self, indptr, src_idx, eid, n_src, n_dst, n_edges, block=None, mask=None):
        self.indptr, self.src_idx, self.eid = indptr, src_idx, eid
        self.n_src, self.n_dst, self.n_edges = int(n_src), int(n_dst), int(n_edges)
        # block-diagonal form (comm graphs): nodes come in consecutive blocks of `block` and bit i of mask[v]
        # says whether the i-th node of v's block sends to v.  Filled by RelGraph.block_mask().
        self.block, self.mask = block, mask

    @property
    def is_star(self) -> bool:
        return self.src_idx is None

    def to(self, device, non_blocking=False) -> "RelCSR":
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return RelCSR(mv(self.indptr), mv(self.src_idx), mv(self.eid), self.n_src, self.n_dst, self.n_edges,
                      self.block, mv(self.mask))

    def pin_memory(self) -> "RelCSR":
        mv = lambda t: None if t is None else t.pin_memory()
        return RelCSR(mv(self.indptr), mv(self.src_idx), mv(self.eid), self.n_src, self.n_dst, self.n_edges,
                      self.block, mv(self.mask))

    def dst_of_slot(self) -> th.Tensor:
        """int64 destination id of every CSR slot (expands indptr)."""
        deg = (self.indptr[1:] - self.indptr[:-1]).to(th.int64)
        return th.repeat_interleave(th.arange(self.n_dst, device=self.indptr.device), deg)

    def src_of_slot(self) -> th.Tensor:
        if self.src_idx is None:
            return th.arange(self.n_edges, device=self.indptr.device)
        return self.src_idx.to(th.int64)

    @staticmethod
    def from_edges(src: th.Tensor, dst: th.Tensor, n_src: int, n_dst: int) -> "RelCSR":
        E = int(dst.numel())
        dev = dst.device
        if E == 0:
            return RelCSR(th.zeros(n_dst + 1, dtype=th.int32, device=dev), None if n_src == 0 else
                          th.zeros(0, dtype=th.int32, device=dev), None, n_src, n_dst, 0)
        is_sorted = bool((dst[1:] >= dst[:-1]).all()) if E > 1 else True
        if is_sorted:
            eid, s = None, src
        else:
            _, eid = th.sort(dst, stable=True)
            s = src[eid]
        counts = th.bincount(dst, minlength=n_dst)
        indptr = th.zeros(n_dst + 1, dtype=th.int64, device=dev)
        th.cumsum(counts, 0, out=indptr[1:])
        star = is_sorted and E == n_src and bool((s == th.arange(E, device=dev)).all())
        return RelCSR(indptr.to(th.int32), None if star else s.to(th.int32), eid, n_src, n_dst, E)


class _NodeSpace:
    """``g.nodes[ntype].data`` accessor."""

    def __init__(self, g):
        self._g = g

    def __getitem__(self, ntype):
        g = self._g
        if ntype not in g._nframes:
            raise KeyError(f"unknown node type {ntype!r}")

        class _V:
            data = g._nframes[ntype]

        return _V


class _TypedDataView:
    """``g.ndata`` on a graph with several node types: values are ``{ntype: tensor}`` dicts
    restricted to the node types that hold the key (DGL semantics relied on by reference
    ``algos/madrqn/agents/gnn_agents.py:53``).  With one node type it degenerates to the frame."""

    def __init__(self, frames: Dict[str, dict], types: Sequence[str]):
        self._frames, self._types = frames, list(types)

    def _single(self):
        return len(self._types) == 1

    def __getitem__(self, key):
        if self._single():
            return self._frames[self._types[0]][key]
        out = {t: self._frames[t][key] for t in self._types if key in self._frames[t]}
        if not out:
            raise KeyError(key)
        return out

    def __setitem__(self, key, val):
        if self._single():
            if isinstance(val, dict):
                val = val[self._types[0]]
            self._frames[self._types[0]][key] = val
            return
        if not isinstance(val, dict):
            raise ValueError("graph has several node types: assign a {ntype: tensor} dict")
        for t, v in val.items():
            self._frames[t][key] = v

    def __contains__(self, key):
        return any(key in self._frames[t] for t in self._types)

    def update(self, d):
        for k, v in d.items():
            self[k] = v

    def pop(self, key, *default):
        if self._single():
            return self._frames[self._types[0]].pop(key, *default)
        out = {t: self._frames[t].pop(key) for t in self._types if key in self._frames[t]}
        if not out and default:
            return default[0]
        if not out:
            raise KeyError(key)
        return out

    def keys(self):
        ks = []
        for t in self._types:
            ks += [k for k in self._frames[t] if k not in ks]
        return ks


class HeteroGraph:
    """Heterogeneous graph with per-node-type feature frames and per-relation CSR-by-dst."""

    def __init__(self, ntypes, cets, num_nodes, src, dst, nframes=None, eframes=None,
                 bnn=None, bne=None, csr=None):
        self._ntypes: Tuple[str, ...] = tuple(ntypes)
        self._cets: Tuple[Tuple[str, str, str], ...] = tuple(cets)
        self._nn: Dict[str, int] = {k: int(v) for k, v in num_nodes.items()}
        self._src: Dict[tuple, th.Tensor] = src
        self._dst: Dict[tuple, th.Tensor] = dst
        self._nframes: Dict[str, dict] = nframes if nframes is not None else {t: {} for t in self._ntypes}
        self._eframes: Dict[tuple, dict] = eframes if eframes is not None else {c: {} for c in self._cets}
        self._bnn: Dict[str, List[int]] = bnn if bnn is not None else {t: [self._nn[t]] for t in self._ntypes}
        self._bne: Dict[tuple, List[int]] = bne if bne is not None else {
            c: [self._ne(c)] for c in self._cets}
        self._csr: Dict[tuple, RelCSR] = csr if csr is not None else {}

    # ------------------------------------------------------------------ structure
    @property
    def ntypes(self):
        return list(self._ntypes)

    @property
    def etypes(self):
        return [c[1] for c in self._cets]

    @property
    def canonical_etypes(self):
        return list(self._cets)

    def to_canonical_etype(self, etype):
        if isinstance(etype, tuple):
            if etype not in self._cets:
                raise KeyError(f"unknown relation {etype!r}")
            return etype
        hits = [c for c in self._cets if c[1] == etype]
        if len(hits) != 1:
            raise KeyError(f"relation {etype!r} is {'ambiguous' if hits else 'unknown'}")
        return hits[0]

    def _only_etype(self):
        if len(self._cets) != 1:
            raise KeyError("graph has several relations: name one")
        return self._cets[0]

    def num_nodes(self, ntype=None) -> int:
        if ntype is None:
            return sum(self._nn.values())
        return self._nn[ntype]

    number_of_nodes = num_nodes

    def num_edges(self, etype=None) -> int:
        if etype is None:
            return sum(self._ne(c) for c in self._cets)
        return self._ne(self.to_canonical_etype(etype))

    number_of_edges = num_edges

    def edges(self, etype=None):
        c = self._only_etype() if etype is None else self.to_canonical_etype(etype)
        return self._edge_list(c)

    def _ne(self, c) -> int:
        d = self._dst.get(c)
        return int(d.numel()) if d is not None else self._csr[c].n_edges

    def _edge_list(self, c):
        """(src, dst) int64 in edge-id order; rebuilt from the CSR (and its slot -> edge-id map) when the graph
        was moved across devices or built CSR-first (edge lists are not shipped to the GPU)."""
        if self._dst.get(c) is None:
            r = self._csr[c]
            u, v = r.src_of_slot(), r.dst_of_slot()
            if r.eid is not None:
                uu, vv = th.empty_like(u), th.empty_like(v)
                uu[r.eid], vv[r.eid] = u, v
                u, v = uu, vv
            self._src[c], self._dst[c] = u, v
        return self._src[c], self._dst[c]

    @property
    def device(self):
        for r in self._csr.values():
            return r.indptr.device
        for d in self._dst.values():
            if d is not None:
                return d.device
        return th.device("cpu")

    @property
    def batch_size(self) -> int:
        return len(next(iter(self._bnn.values())))

    def batch_num_nodes(self, ntype=None):
        if ntype is None:
            if len(self._ntypes) != 1:
                raise KeyError("graph has several node types: name one")
            ntype = self._ntypes[0]
        return th.tensor(self._bnn[ntype], dtype=th.int64)

    def batch_num_edges(self, etype=None):
        c = self._only_etype() if etype is None else self.to_canonical_etype(etype)
        return th.tensor(self._bne[c], dtype=th.int64)

    def csr(self, etype=None) -> RelCSR:
        """CSR-by-destination of one relation (built once, cached, carried through batch/merge/to)."""
        c = self._only_etype() if etype is None else self.to_canonical_etype(etype)
        r = self._csr.get(c)
        if r is None:
            r = RelCSR.from_edges(self._src[c], self._dst[c], self._nn[c[0]], self._nn[c[2]])  # lists exist here
            self._csr[c] = r
        return r

    def uniform_block(self, ntype) -> Optional[int]:
        """Node count per batched graph when all parts have the same count, else None.  The fused
        comm kernels use it: a batch of per-env graphs is block-diagonal with blocks of U agents."""
        b = self._bnn[ntype]
        return b[0] if b and all(x == b[0] for x in b) else None

    # ------------------------------------------------------------------ features
    @property
    def nodes(self):
        return _NodeSpace(self)

    @property
    def ndata(self):
        return _TypedDataView(self._nframes, self._ntypes)

    @property
    def srcdata(self):
        return _TypedDataView(self._nframes, sorted({c[0] for c in self._cets}, key=self._ntypes.index))

    @property
    def dstdata(self):
        return _TypedDataView(self._nframes, sorted({c[2] for c in self._cets}, key=self._ntypes.index))

    @property
    def edata(self):
        if len(self._cets) == 1:
            return self._eframes[self._cets[0]]
        return _TypedDataView(self._eframes, self._cets)

    def __getitem__(self, etype) -> "RelGraph":
        return RelGraph(self, self.to_canonical_etype(etype))

    @contextmanager
    def local_scope(self):
        """Feature writes inside the scope are dropped on exit (reference ``gnn_agents.py:136,249``)."""
        nsave = {t: dict(f) for t, f in self._nframes.items()}
        esave = {c: dict(f) for c, f in self._eframes.items()}
        try:
            yield self
        finally:
            for t in self._nframes:
                self._nframes[t].clear()
                self._nframes[t].update(nsave[t])
            for c in self._eframes:
                self._eframes[c].clear()
                self._eframes[c].update(esave[c])

    # ------------------------------------------------------------------ movement
    def _map(self, fn_t, fn_csr, build_csr=True) -> "HeteroGraph":
        if build_csr:
            for c in self._cets:
                self.csr(c)
                if c[0] == c[2]:
                    RelGraph(self, c).block_mask()
        return HeteroGraph(
            self._ntypes, self._cets, self._nn,
            {c: None for c in self._cets}, {c: None for c in self._cets},     # edge lists stay behind (lazy)
            {t: {k: fn_t(v) for k, v in f.items()} for t, f in self._nframes.items()},
            {c: {k: fn_t(v) for k, v in f.items()} for c, f in self._eframes.items()},
            {t: list(v) for t, v in self._bnn.items()}, {c: list(v) for c, v in self._bne.items()},
            {c: fn_csr(r) for c, r in self._csr.items()})

    def to(self, device, non_blocking: bool = False) -> "HeteroGraph":
        """Moves structure, CSR and features.  CSRs are finalised on the source device first so that
        no sort / host sync happens on the GPU (reference call sites ``learner.py:71,116``)."""
        device = th.device(device)
        if device == self.device and all(v.device == device for f in self._nframes.values() for v in f.values()):
            return self
        return self._map(lambda t: t.to(device, non_blocking=non_blocking),
                         lambda r: r.to(device, non_blocking=non_blocking))

    def pin_memory(self) -> "HeteroGraph":
        return self._map(lambda t: t.pin_memory(), lambda r: r.pin_memory())

    def cpu(self):
        return self.to("cpu")

    def __repr__(self):
        nn_ = {t: self._nn[t] for t in self._ntypes}
        ne = {c: self._ne(c) for c in self._cets}
        return f"HeteroGraph(num_nodes={nn_}, num_edges={ne}, batch_size={self.batch_size})"


DGLGraph = HeteroGraph  # ``isinstance(x, dgl.DGLGraph)`` in reference algos/common.py:44


class RelGraph:
    """View of one relation ``(src_type, etype, dst_type)`` of a HeteroGraph (``g['talk']``)."""

    def __init__(self, parent: HeteroGraph, cet):
        self._g, self._c = parent, cet

    @property
    def parent(self):
        return self._g

    @property
    def canonical_etype(self):
        return self._c

    @property
    def device(self):
        return self._g.device

    def csr(self) -> RelCSR:
        return self._g.csr(self._c)

    def edges(self):
        return self._g._edge_list(self._c)

    def num_src_nodes(self):
        return self._g._nn[self._c[0]]

    def num_dst_nodes(self):
        return self._g._nn[self._c[2]]

    def num_nodes(self, ntype=None):
        if ntype is not None:
            return self._g._nn[ntype]
        st, _, dt = self._c
        return self._g._nn[st] if st == dt else self._g._nn[st] + self._g._nn[dt]

    number_of_nodes = num_nodes

    def num_edges(self):
        return self._g._ne(self._c)

    number_of_edges = num_edges

    def uniform_block(self) -> Optional[int]:
        return self._g.uniform_block(self._c[2])

    def block_mask(self):
        """``(block, mask)`` of a block-diagonal homogeneous relation, or ``None`` when the relation is not a
        batch of equally sized graphs with at most 32 nodes each.  ``mask`` is int32 (uint32 bit pattern)."""
        st, _, dt = self._c
        if st != dt:
            return None
        c = self.csr()
        if c.mask is None:
            U = self.uniform_block()
            if U is None or U < 1 or U > 32:
                return None
            dst, src = c.dst_of_slot(), c.src_of_slot()
            local = src - (dst // U) * U
            bits = th.zeros(c.n_dst, dtype=th.int64, device=dst.device)
            if c.n_edges:
                bits.index_add_(0, dst, th.ones_like(local) << local)   # edges are unique => add == or
            c.block, c.mask = U, bits.to(th.int32)
        return c.block, c.mask

    def in_degrees(self):
        ip = self.csr().indptr
        return (ip[1:] - ip[:-1]).to(th.int64)

    @property
    def srcdata(self):
        return self._g._nframes[self._c[0]]

    @property
    def dstdata(self):
        return self._g._nframes[self._c[2]]

    @property
    def ndata(self):
        st, _, dt = self._c
        if st == dt:
            return self._g._nframes[st]
        return _TypedDataView(self._g._nframes, [st, dt])

    @property
    def edata(self):
        return self._g._eframes[self._c]

    def local_scope(self):
        return self._g.local_scope()

    # message passing (generic torch path; the fused CUDA modules never go through here)
    def apply_edges(self, func):
        from . import function as _fn
        _fn.apply_edges(self, func)

    def update_all(self, message_func, reduce_func):
        from . import function as _fn
        _fn.update_all(self, message_func, reduce_func)


# ---------------------------------------------------------------------- constructors
def heterograph(data_dict, num_nodes_dict=None, device=None) -> HeteroGraph:
    """``dgl.heterograph`` equivalent (reference ``env_wrappers.py:81,146``; drqn ``:70``): edge ids follow
    list order; node counts come from ``num_nodes_dict`` (or max id + 1)."""
    cets, src, dst = [], {}, {}
    ntypes: List[str] = []
    for cet, (u, v) in data_dict.items():
        cet = tuple(cet)
        cets.append(cet)
        src[cet], dst[cet] = _as_index(u, device), _as_index(v, device)
        if src[cet].numel() != dst[cet].numel():
            raise ValueError(f"relation {cet}: src/dst length mismatch")
        for t in (cet[0], cet[2]):
            if t not in ntypes:
                ntypes.append(t)
    nn_ = {}
    for t in sorted(ntypes):
        if num_nodes_dict is not None and t in num_nodes_dict:
            nn_[t] = int(num_nodes_dict[t])
        else:
            m = 0
            for c in cets:
                if c[0] == t and src[c].numel():
                    m = max(m, int(src[c].max()) + 1)
                if c[2] == t and dst[c].numel():
                    m = max(m, int(dst[c].max()) + 1)
            nn_[t] = m
    for c in cets:
        if src[c].numel() and (int(src[c].max()) >= nn_[c[0]] or int(dst[c].max()) >= nn_[c[2]]):
            raise ValueError(f"relation {c}: node id out of range")
    return HeteroGraph(sorted(ntypes), sorted(cets), nn_, src, dst)


def _cat_frames(frames: List[dict], counts: List[int], what: str) -> dict:
    keys = None
    for f, n in zip(frames, counts):
        if keys is None:
            keys = list(f.keys())
        elif set(f.keys()) != set(keys):
            if n == 0 and not f:
                continue
            raise ValueError(f"batch: {what} feature keys differ across graphs")
    out = {}
    for k in keys or []:
        out[k] = th.cat([f[k] for f in frames if k in f], 0)
    return out


def batch(graphs: Sequence[HeteroGraph]) -> HeteroGraph:
    """``dgl.batch`` equivalent (reference ``env_wrappers.py:67``, ``algos/common.py:45``): node ids per type
    and edge ids per relation are concatenated in list order with offsets; features are
    concatenated; batch bookkeeping of already-batched inputs is flattened."""
    graphs = list(graphs)
    if not graphs:
        raise ValueError("batch of zero graphs")
    g0 = graphs[0]
    if len(graphs) == 1:
        return g0
    for g in graphs[1:]:
        if g._ntypes != g0._ntypes or g._cets != g0._cets:
            raise ValueError("batch: graphs must share node and edge types")
    nn_, bnn, noff = {}, {}, {}
    for t in g0._ntypes:
        counts = [g._nn[t] for g in graphs]
        nn_[t] = sum(counts)
        offs, acc = [], 0
        for cnt in counts:
            offs.append(acc)
            acc += cnt
        noff[t] = offs
        bnn[t] = [x for g in graphs for x in g._bnn[t]]
    src, dst, bne, csr = {}, {}, {}, {}
    for c in g0._cets:
        st, _, dt = c
        if all(g._dst.get(c) is not None for g in graphs):
            src[c] = th.cat([g._src[c] + noff[st][i] for i, g in enumerate(graphs)])
            dst[c] = th.cat([g._dst[c] + noff[dt][i] for i, g in enumerate(graphs)])
        else:
            src[c] = dst[c] = None                      # rebuilt lazily from the composed CSR
        bne[c] = [x for g in graphs for x in g._bne[c]]
        parts = [g.csr(c) for g in graphs]
        eoff, acc = [], 0
        for p in parts:
            eoff.append(acc)
            acc += p.n_edges
        dev = parts[0].indptr.device
        indptr = th.cat([p.indptr[:-1] + eoff[i] for i, p in enumerate(parts)]
                        + [th.tensor([acc], dtype=th.int32, device=dev)])
        all_star = all(p.src_idx is None for p in parts)
        if all_star:
            src_idx = None
        else:
            src_idx = th.cat([(p.src_idx if p.src_idx is not None
                               else th.arange(p.n_edges, dtype=th.int32, device=dev)) + noff[st][i]
                              for i, p in enumerate(parts)])
        if all(p.eid is None for p in parts):
            eid = None
        else:
            eid = th.cat([(p.eid if p.eid is not None else th.arange(p.n_edges, device=dev)) + eoff[i]
                          for i, p in enumerate(parts)])
        blk, msk = None, None
        if st == dt and all(p.mask is not None and p.block == parts[0].block for p in parts):
            blk, msk = parts[0].block, th.cat([p.mask for p in parts])   # block-diagonal form composes by concat
        csr[c] = RelCSR(indptr, src_idx, eid, nn_[st], nn_[dt], acc, blk, msk)
    nframes = {t: _cat_frames([g._nframes[t] for g in graphs], [g._nn[t] for g in graphs], f"node[{t}]")
               for t in g0._ntypes}
    eframes = {c: _cat_frames([g._eframes[c] for g in graphs], [g._ne(c) for g in graphs],
                              f"edge[{c}]") for c in g0._cets}
    return HeteroGraph(g0._ntypes, g0._cets, nn_, src, dst, nframes, eframes, bnn, bne, csr)


def merge(graphs: Sequence[HeteroGraph]) -> HeteroGraph:
    """``dgl.merge`` equivalent (reference ``env_wrappers.py:137``): same node/edge types; per node type
    the count is the max over inputs; per relation the edge lists are concatenated in input order; node
    features come from the inputs that hold them.  The result is one (unbatched) graph."""
    graphs = list(graphs)
    g0 = graphs[0]
    for g in graphs[1:]:
        if g._ntypes != g0._ntypes or g._cets != g0._cets:
            raise ValueError("merge: graphs must share node and edge types")
    nn_ = {t: max(g._nn[t] for g in graphs) for t in g0._ntypes}
    src, dst, csr = {}, {}, {}
    for c in g0._cets:
        src[c] = th.cat([g._edge_list(c)[0] for g in graphs])
        dst[c] = th.cat([g._edge_list(c)[1] for g in graphs])
        holders = [g for g in graphs if g._ne(c) > 0]
        if len(holders) == 1:  # CSR can be reused; only the destination count may grow
            p = holders[0].csr(c)
            indptr = p.indptr
            if p.n_dst < nn_[c[2]]:
                indptr = th.cat([indptr, indptr[-1:].expand(nn_[c[2]] - p.n_dst)])
            src_idx = p.src_idx
            if src_idx is None and p.n_edges != nn_[c[0]]:
                src_idx = th.arange(p.n_edges, dtype=th.int32, device=indptr.device)
            csr[c] = RelCSR(indptr, src_idx, p.eid, nn_[c[0]], nn_[c[2]], p.n_edges)
    nframes = {t: {} for t in g0._ntypes}
    for t in g0._ntypes:
        for g in graphs:
            for k, v in g._nframes[t].items():
                if k not in nframes[t] and v.shape[0] == nn_[t]:
                    nframes[t][k] = v
    eframes = {c: {} for c in g0._cets}
    for c in g0._cets:
        keys = [k for g in graphs for k in g._eframes[c]]
        for k in dict.fromkeys(keys):
            if all(k in g._eframes[c] or g._ne(c) == 0 for g in graphs):
                eframes[c][k] = th.cat([g._eframes[c][k] for g in graphs if k in g._eframes[c]])
    return HeteroGraph(g0._ntypes, g0._cets, nn_, src, dst, nframes, eframes, None, None, csr)
