"""``torch.autograd.Function`` wrappers around the C ABI (``include/ubs_gnn.h``).

Each wrapper only checks dtype / contiguity / device, allocates outputs and workspaces through PyTorch's
caching allocator, and enqueues the kernels on the current CUDA stream.  CUDA tensors only.
"""
from __future__ import annotations

import os

import torch as th

from . import _lib

GAT_RESIDUAL, GAT_RELU = 1, 2
ACT_PDL = 0x100            # ubs_agent_act_rel_fwd: programmatic dependent launch (include/ubs_gnn.h UBS_ACT_PDL)


class KernelTimer:
    """CUDA-event timing of individual C-ABI calls on the launching stream (used by bench.py's roofline leg)."""

    def __init__(self):
        self.records = []          # (name, meta, start_event, end_event)

    def summary(self):
        th.cuda.synchronize()
        out = {}
        for name, meta, s, e in self.records:
            d = out.setdefault(name, dict(count=0, ms=0.0, metas=[]))
            d["count"] += 1
            d["ms"] += s.elapsed_time(e)
            d["metas"].append(meta)
        return out


TIMER = None       # set to a KernelTimer to time every call below


class _timed:
    def __init__(self, name, meta):
        self.name, self.meta = name, meta

    def __enter__(self):
        if TIMER is not None:
            self.s = th.cuda.Event(enable_timing=True)
            self.e = th.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if TIMER is not None:
            self.e.record()
            TIMER.records.append((self.name, self.meta, self.s, self.e))
        return False


def _f32c(t):
    if t is None:
        return None
    if t.dtype != th.float32:
        raise TypeError(f"expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def gatv2_fused_supported(F_s: int, F_d: int, heads: int, D: int, slope: float) -> bool:
    return (1 <= F_s <= 4 and 1 <= F_d <= 2 and heads in (1, 2, 4, 8) and heads * D in (32, 64, 128)
            and 0.0 <= slope <= 1.0)


class GATv2Fused(th.autograd.Function):
    """Fused GATv2 relation (``ubs_gatv2_fwd`` / ``ubs_gatv2_bwd``).  Returns ``(n_dst, heads*D)``."""

    @staticmethod
    def forward(ctx, x_src, x_dst, indptr, src_idx, W_src, b_src, W_dst, b_dst, attn, W_res, b_res,
                heads, D, slope, flags):
        _lib.require_cuda(x_src, x_dst, indptr, W_src)
        lib = _lib.load()
        x_src, x_dst = _f32c(x_src), _f32c(x_dst)
        W_src, b_src, W_dst, b_dst = _f32c(W_src), _f32c(b_src), _f32c(W_dst), _f32c(b_dst)
        attn_c, W_res, b_res = _f32c(attn), _f32c(W_res), _f32c(b_res)
        if indptr.dtype != th.int32 or (src_idx is not None and src_idx.dtype != th.int32):
            raise TypeError("indptr / src_idx must be int32")
        n_dst, H = x_dst.shape[0], heads * D
        n_edges = x_src.shape[0] if src_idx is None else src_idx.shape[0]
        F_s, F_d = x_src.shape[1], x_dst.shape[1]
        if indptr.numel() != n_dst + 1:
            raise ValueError("indptr must have n_dst + 1 entries")
        need_grad = any(t is not None and t.requires_grad for t in
                        (x_src, x_dst, W_src, b_src, W_dst, b_dst, attn, W_res, b_res))
        out = th.empty(n_dst, H, dtype=th.float32, device=x_dst.device)
        smax = ssum = None
        if need_grad:
            smax = th.empty(n_dst, heads, dtype=th.float32, device=x_dst.device)
            ssum = th.empty_like(smax)
        P = _lib.ptr
        with _timed("gatv2_fwd", (n_dst, n_edges, F_s, F_d, heads, D, need_grad)):
            _lib.check(lib.ubs_gatv2_fwd(P(x_src), P(x_dst), P(indptr), P(src_idx), P(W_src), P(b_src), P(W_dst),
                                         P(b_dst), P(attn_c), P(W_res), P(b_res), P(out), P(smax), P(ssum), n_dst,
                                         n_edges, F_s, F_d, heads, D, float(slope), int(flags), _lib.stream()),
                       "ubs_gatv2_fwd")
        if need_grad:
            ctx.save_for_backward(x_src, x_dst, indptr, src_idx, W_src, b_src, W_dst, b_dst, attn_c, W_res, b_res,
                                  out, smax, ssum)
            ctx.cfg = (heads, D, float(slope), int(flags), n_edges, attn.shape)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (x_src, x_dst, indptr, src_idx, W_src, b_src, W_dst, b_dst, attn, W_res, b_res, out, smax, ssum) = ctx.saved_tensors
        heads, D, slope, flags, n_edges, attn_shape = ctx.cfg
        lib = _lib.load()
        H, F_s, F_d, n_dst, n_src = heads * D, x_src.shape[1], x_dst.shape[1], x_dst.shape[0], x_src.shape[0]
        dev = x_dst.device
        grad_out = _f32c(grad_out)
        Pn = H * (F_s + 2 * F_d + 4)
        gparams = th.empty(Pn, dtype=th.float32, device=dev)
        ws = th.empty(int(lib.ubs_gatv2_bwd_workspace(n_dst, F_s, F_d, heads, D)), dtype=th.float32, device=dev)
        gxs = gxd = None
        if ctx.needs_input_grad[0]:
            gxs = (th.zeros if src_idx is not None else th.empty)(n_src, F_s, dtype=th.float32, device=dev)
            if src_idx is None and n_src != n_edges:
                gxs.zero_()
        if ctx.needs_input_grad[1]:
            gxd = th.empty(n_dst, F_d, dtype=th.float32, device=dev)
        P = _lib.ptr
        with _timed("gatv2_bwd", (n_dst, n_edges, F_s, F_d, heads, D, True)):
            _lib.check(lib.ubs_gatv2_bwd(P(x_src), P(x_dst), P(indptr), P(src_idx), P(W_src), P(b_src), P(W_dst),
                                         P(b_dst), P(attn), P(W_res), P(b_res), P(out), P(grad_out), P(smax), P(ssum),
                                         P(gparams), P(gxs), P(gxd), P(ws), n_dst, n_edges, n_src, F_s, F_d, heads, D,
                                         slope, flags, _lib.stream()), "ubs_gatv2_bwd")
        o = 0

        def take(n, shape):
            nonlocal o
            t = gparams[o:o + n].view(shape)
            o += n
            return t
        gWs, gbs = take(H * F_s, (H, F_s)), take(H, (H,))
        gWd, gbd = take(H * F_d, (H, F_d)), take(H, (H,))
        gat = take(H, attn_shape)
        gWr, gbr = take(H * F_d, (H, F_d)), take(H, (H,))
        return (gxs, gxd, None, None, gWs, gbs if b_src is not None else None, gWd,
                gbd if b_dst is not None else None, gat, gWr if W_res is not None else None,
                gbr if b_res is not None else None, None, None, None, None)


def gat_aggregate_supported(heads: int, D: int, slope: float) -> bool:
    H = heads * D
    return H in (32, 64, 128, 256) and heads in (1, 2, 4, 8) and D % (H // 32) == 0 and 0.0 <= slope <= 1.0


class GATAggregate(th.autograd.Function):
    """GATv2 attention + aggregation on pre-projected features over a general CSR (``ubs_gat_aggr_fwd/bwd``).
    ``forward(el (n_src,H), er (n_dst,H), res (n_dst,H)|None, indptr, src_idx|None, attn, heads, D, slope, flags)``."""

    @staticmethod
    def forward(ctx, el, er, res, indptr, src_idx, attn, heads, D, slope, flags):
        _lib.require_cuda(el, er, indptr, attn)
        lib = _lib.load()
        el, er, res, attn_c = _f32c(el), _f32c(er), _f32c(res), _f32c(attn)
        if indptr.dtype != th.int32 or (src_idx is not None and src_idx.dtype != th.int32):
            raise TypeError("indptr / src_idx must be int32")
        n_dst, H = er.shape[0], heads * D
        n_edges = el.shape[0] if src_idx is None else src_idx.shape[0]
        need = any(t is not None and t.requires_grad for t in (el, er, res, attn))
        out = th.empty(n_dst, H, dtype=th.float32, device=er.device)
        smax = th.empty(n_dst, heads, dtype=th.float32, device=er.device) if need else None
        ssum = th.empty_like(smax) if need else None
        P = _lib.ptr
        with _timed("gat_aggr_fwd", (n_dst, n_edges, H, heads, need)):
            _lib.check(lib.ubs_gat_aggr_fwd(P(el), P(er), P(res), P(indptr), P(src_idx), P(attn_c), P(out), P(smax),
                                            P(ssum), n_dst, n_edges, heads, D, float(slope), int(flags), _lib.stream()),
                       "ubs_gat_aggr_fwd")
        if need:
            ctx.save_for_backward(el, er, res, indptr, src_idx, attn_c, out, smax, ssum)
            ctx.cfg = (heads, D, float(slope), int(flags), n_edges, attn.shape)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        el, er, res, indptr, src_idx, attn, out, smax, ssum = ctx.saved_tensors
        heads, D, slope, flags, n_edges, attn_shape = ctx.cfg
        lib = _lib.load()
        n_dst, H = er.shape[0], heads * D
        grad_out = _f32c(grad_out)
        g_el = (th.zeros_like if src_idx is not None else th.empty_like)(el)
        if src_idx is None and el.shape[0] != n_edges:
            g_el.zero_()
        g_er = th.empty_like(er)
        g_res = th.empty_like(er) if res is not None else None
        g_attn = th.empty(H, dtype=th.float32, device=er.device)
        ws = th.empty(int(lib.ubs_gat_aggr_bwd_workspace(n_dst, heads, D)), dtype=th.float32, device=er.device)
        P = _lib.ptr
        with _timed("gat_aggr_bwd", (n_dst, n_edges, H, heads, True)):
            _lib.check(lib.ubs_gat_aggr_bwd(P(el), P(er), P(res), P(indptr), P(src_idx), P(attn), P(out), P(grad_out),
                                            P(smax), P(ssum), P(g_el), P(g_er), P(g_res), P(g_attn), P(ws), n_dst,
                                            n_edges, heads, D, slope, flags, _lib.stream()), "ubs_gat_aggr_bwd")
        return g_el, g_er, g_res, None, None, g_attn.view(attn_shape), None, None, None, None


class BlockAttention(th.autograd.Function):
    """TarMAC attention over a block-diagonal comm graph.  ``vsq`` is ``(N, M + 2K)`` laid out ``[v | s | q]``."""

    @staticmethod
    def forward(ctx, vsq, mask, block, K, M, scale):
        _lib.require_cuda(vsq, mask)
        lib = _lib.load()
        vsq = _f32c(vsq)
        N, ld = vsq.shape
        if ld != M + 2 * K:
            raise ValueError("vsq must be (N, M + 2K)")
        if mask.dtype not in (th.int32, th.uint32) or mask.numel() != N:
            raise TypeError("mask must be int32 (bit pattern of uint32) with one entry per node")
        c = th.empty(N, M, dtype=th.float32, device=vsq.device)
        alpha = th.empty(N, block, dtype=th.float32, device=vsq.device)
        base, es = vsq.data_ptr(), 4
        with _timed("block_attn_fwd", (N, block, K, M)):
            _lib.check(lib.ubs_block_attn_fwd(base + es * M, ld, base + es * (M + K), ld, base, ld, _lib.ptr(mask),
                                              _lib.ptr(c), _lib.ptr(alpha), N, block, K, M, float(scale),
                                              _lib.stream()), "ubs_block_attn_fwd")
        ctx.save_for_backward(vsq, mask, alpha)
        ctx.cfg = (block, K, M, float(scale))
        return c

    @staticmethod
    def backward(ctx, grad_c):
        vsq, mask, alpha = ctx.saved_tensors
        block, K, M, scale = ctx.cfg
        lib = _lib.load()
        grad_c = _f32c(grad_c)
        N, ld = vsq.shape
        g = th.empty_like(vsq)
        ds = th.empty(N, block, dtype=th.float32, device=vsq.device)
        base, gb, es = vsq.data_ptr(), g.data_ptr(), 4
        with _timed("block_attn_bwd", (N, block, K, M)):
            _lib.check(lib.ubs_block_attn_bwd(base + es * M, ld, base + es * (M + K), ld, base, ld, _lib.ptr(mask),
                                              _lib.ptr(alpha), _lib.ptr(grad_c), gb + es * M, ld, gb + es * (M + K),
                                              ld, gb, ld, _lib.ptr(ds), N, block, K, M, scale, _lib.stream()),
                       "ubs_block_attn_bwd")
        return g, None, None, None, None, None


class BlockMean(th.autograd.Function):
    """``c_v = mean_{u -> v} msg[u]`` over a block-diagonal comm graph (``ubs_block_mean_fwd / bwd``): the reduce step of
    BaseComm / CommNet.  ``msg (N, F)``, ``mask (N,) int32`` bit patterns, ``block`` = agents per env."""

    @staticmethod
    def forward(ctx, msg, mask, block):
        _lib.require_cuda(msg, mask)
        lib = _lib.load()
        msg = _f32c(msg)
        N, F_ = msg.shape
        if mask.dtype not in (th.int32, th.uint32) or mask.numel() != N:
            raise TypeError("mask must be int32 (bit pattern of uint32) with one entry per node")
        out = th.empty(N, F_, dtype=th.float32, device=msg.device)
        with _timed("block_mean_fwd", (N, block, F_)):
            _lib.check(lib.ubs_block_mean_fwd(msg.data_ptr(), F_, _lib.ptr(mask), out.data_ptr(), F_, N, block, F_,
                                              _lib.stream()), "ubs_block_mean_fwd")
        ctx.save_for_backward(mask)
        ctx.block = block
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (mask,) = ctx.saved_tensors
        lib = _lib.load()
        grad_out = _f32c(grad_out)
        N, F_ = grad_out.shape
        g = th.empty_like(grad_out)
        with _timed("block_mean_bwd", (N, ctx.block, F_)):
            _lib.check(lib.ubs_block_mean_bwd(grad_out.data_ptr(), F_, _lib.ptr(mask), g.data_ptr(), F_, N, ctx.block, F_,
                                              _lib.stream()), "ubs_block_mean_bwd")
        return g, None, None


class BlockBitMax(th.autograd.Function):
    """DiscreteComm's per-edge hard Gumbel-softmax + element-wise max over in-edges on a block-diagonal comm graph
    (``ubs_block_bitmax_fwd / bwd``).  ``logits (N, 2M)`` = ``f_enc`` per node, ``expo (E, M, 2)`` Exponential(1) noise in
    edge-id order, ``mask (N,)`` int32 bit patterns, ``block`` agents per env -> ``c (N, 2M)``."""

    @staticmethod
    def forward(ctx, logits, expo, mask, block, tau):
        _lib.require_cuda(logits, expo, mask)
        lib = _lib.load()
        logits, expo = _f32c(logits), _f32c(expo)
        N, F2 = logits.shape
        M = F2 // 2
        if mask.dtype not in (th.int32, th.uint32) or mask.numel() != N:
            raise TypeError("mask must be int32 (bit pattern of uint32) with one entry per node")
        # first edge id of every env: exclusive prefix sum of the envs' edge counts (popcount of their masks)
        m64 = mask.to(th.int64) & 0xFFFFFFFF
        cnt = th.zeros_like(m64)
        for i in range(block):
            cnt += (m64 >> i) & 1
        per_env = cnt.view(-1, block).sum(1)
        eoff = (th.cumsum(per_env, 0) - per_env).contiguous()
        out = th.empty(N, F2, dtype=th.float32, device=logits.device)
        need = logits.requires_grad
        winner = th.empty(N, F2, dtype=th.uint8, device=logits.device) if need else None
        with _timed("block_bitmax_fwd", (N, block, M)):
            _lib.check(lib.ubs_block_bitmax_fwd(logits.data_ptr(), F2, expo.data_ptr(), _lib.ptr(mask), eoff.data_ptr(),
                                                out.data_ptr(), F2, _lib.ptr(winner), N, block, M, float(tau), _lib.stream()),
                       "ubs_block_bitmax_fwd")
        if need:
            ctx.save_for_backward(logits, expo, mask, eoff, winner)
            ctx.cfg = (block, M, float(tau))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        logits, expo, mask, eoff, winner = ctx.saved_tensors
        block, M, tau = ctx.cfg
        lib = _lib.load()
        grad_out = _f32c(grad_out)
        N, F2 = logits.shape
        g = th.empty_like(logits)
        with _timed("block_bitmax_bwd", (N, block, M)):
            _lib.check(lib.ubs_block_bitmax_bwd(logits.data_ptr(), F2, expo.data_ptr(), _lib.ptr(mask), eoff.data_ptr(),
                                                winner.data_ptr(), grad_out.data_ptr(), F2, g.data_ptr(), F2, N, block, M,
                                                tau, _lib.stream()), "ubs_block_bitmax_bwd")
        return g, None, None, None, None


class GRUGates(th.autograd.Function):
    """``nn.GRUCell`` gate math on precomputed projections ``gi (N,3H)``, ``gh (N,3H)`` and ``h (N,H)``."""

    @staticmethod
    def forward(ctx, gi, gh, h):
        _lib.require_cuda(gi, gh, h)
        lib = _lib.load()
        gi, gh, h = _f32c(gi), _f32c(gh), _f32c(h)
        N, H = h.shape
        out = th.empty_like(h)
        with _timed("gru_gates_fwd", (N, H)):
            _lib.check(lib.ubs_gru_gates_fwd(_lib.ptr(gi), _lib.ptr(gh), _lib.ptr(h), _lib.ptr(out), N, H,
                                             _lib.stream()), "ubs_gru_gates_fwd")
        ctx.save_for_backward(gi, gh, h)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        gi, gh, h = ctx.saved_tensors
        lib = _lib.load()
        grad_out = _f32c(grad_out)
        N, H = h.shape
        ggi, ggh, gh_d = th.empty_like(gi), th.empty_like(gh), th.empty_like(h)
        with _timed("gru_gates_bwd", (N, H)):
            _lib.check(lib.ubs_gru_gates_bwd(_lib.ptr(gi), _lib.ptr(gh), _lib.ptr(h), _lib.ptr(grad_out), _lib.ptr(ggi),
                                             _lib.ptr(ggh), _lib.ptr(gh_d), N, H, _lib.stream()), "ubs_gru_gates_bwd")
        return ggi, ggh, gh_d


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """GRUCell = two library GEMMs (cuBLAS fp32 through ``torch.addmm``) + the fused gate kernel."""
    gi = th.addmm(b_ih, x, w_ih.t()) if b_ih is not None else x @ w_ih.t()
    gh = th.addmm(b_hh, h, w_hh.t()) if b_hh is not None else h @ w_hh.t()
    return GRUGates.apply(gi, gh, h)


# ====================================================================================================================
# Fused recurrent agent sequence (ubs_agent_pack / ubs_agent_seq_fwd / ubs_agent_seq_bwd)
STEP_AGGR, STEP_TARMAC = 1, 2
PARAM_ORDER = ("W_aggr", "b_aggr", "W_val", "b_val", "W_sign", "b_sign", "W_que", "b_que", "W_ih", "b_ih", "W_hh",
               "b_hh", "W_out", "b_out")


class AgentDims:
    """Shape of the fused step: hidden H, message M, key K, actions A, agents per env U, input width Fin, flags."""

    def __init__(self, H, M, K, A, U, Fin, flags):
        self.H, self.M, self.K, self.A, self.U, self.Fin, self.flags = H, M, K, A, U, Fin, flags
        self.tarmac, self.aggr = bool(flags & STEP_TARMAC), bool(flags & STEP_AGGR)
        self.V = (M + 2 * K) if self.tarmac else 0
        self.Vp = (self.V + 3) // 4 * 4
        self.I = H + M if self.tarmac else H

    def ints(self):
        return (self.H, self.M, self.K, self.A, self.U, self.Fin, self.flags)

    def supported(self):
        ok = self.H % 4 == 0 and 16 <= self.H <= 256 and 1 <= self.A <= 64
        if self.tarmac:
            ok = ok and 1 <= self.U <= 16 and self.M % 4 == 0 and self.M >= 4 and self.K >= 1
        ok = ok and ((self.Fin % 4 == 0 and self.Fin >= 4) if self.aggr else self.Fin == self.H)
        return ok


def agent_pack(dims: AgentDims, params: dict, out=None):
    """Builds the packed weight buffer (transposed copies for the forward GEMMs + originals for the backward)."""
    lib = _lib.load()
    n = int(lib.ubs_agent_pack_size(*dims.ints()))
    dev = params["W_ih"].device
    if out is None or out.numel() != n:
        out = th.empty(n, dtype=th.float32, device=dev)
    ptrs = [_lib.ptr(_f32c(params.get(k).detach()) if params.get(k) is not None else None) for k in PARAM_ORDER]
    _lib.check(lib.ubs_agent_pack(*dims.ints(), *ptrs, _lib.ptr(out), _lib.stream()), "ubs_agent_pack")
    return out


def agent_seq_infer(dims: AgentDims, packed, xin, h0, mask, want_actions=False, h_out=None, q=None, acts=None,
                    explore=None):
    """Inference: ``xin (T,N,Fin)``, ``h0 (N,H)`` -> ``q (T,N,A)``, ``h_out (T,N,H)`` [, greedy actions (T,N) int64].
    ``h_out`` / ``q`` / ``acts`` may be preallocated contiguous destinations (e.g. slices of a sequence arena)."""
    lib = _lib.load()
    _lib.require_cuda(xin, h0, packed)
    xin, h0 = _f32c(xin), _f32c(h0)
    T, N = xin.shape[0], xin.shape[1]
    if h_out is None:
        h_out = th.empty(T, N, dims.H, dtype=th.float32, device=xin.device)
    if q is None:
        q = th.empty(T, N, dims.A, dtype=th.float32, device=xin.device)
    if acts is None and want_actions:
        acts = th.empty(T, N, dtype=th.int64, device=xin.device)
    for t_, n_ in ((h_out, T * N * dims.H), (q, T * N * dims.A), (acts, T * N)):
        if t_ is not None and (not t_.is_contiguous() or t_.numel() != n_):
            raise ValueError("agent_seq_infer: output buffers must be contiguous and exactly sized")
    want_actions = acts is not None
    eg_u, eg_a, eg_eps = explore if explore is not None else (None, None, None)      # fused epsilon-greedy
    with _timed("agent_seq_fwd", (T, N, dims.ints(), False)):
        _lib.check(lib.ubs_agent_act_fwd(*dims.ints(), _lib.ptr(packed), _lib.ptr(xin), _lib.ptr(h0), _lib.ptr(mask),
                                         _lib.ptr(h_out), _lib.ptr(q), _lib.ptr(acts), _lib.ptr(eg_u), _lib.ptr(eg_a),
                                         _lib.ptr(eg_eps), None, None, None, None, N, T, _lib.stream()),
                   "ubs_agent_act_fwd")
    return (q, h_out, acts) if want_actions else (q, h_out)


def gatv2_rel_pack(conv_params, F_s, F_d, heads, D, slope, flags, out=None):
    """Per-relation constant tables of the fused act step (``ubs_gatv2_rel_pack``) for the relations in
    ``conv_params`` = ``[(W_src, b_src, W_dst, b_dst, attn, W_res, b_res), ...]`` (``F_s`` one width per relation),
    concatenated into ONE buffer (rebuilt in place when ``out`` is given: captured CUDA graphs keep its address)."""
    lib = _lib.load()
    n = int(lib.ubs_gatv2_rel_pack_size(heads, D))
    dev = conv_params[0][0].device
    if out is None or out.numel() != n * len(conv_params):
        out = th.empty(n * len(conv_params), dtype=th.float32, device=dev)
    for r, (ps, fs) in enumerate(zip(conv_params, F_s)):
        ptrs = [_lib.ptr(_f32c(p.detach()) if p is not None else None) for p in ps]
        _lib.check(lib.ubs_gatv2_rel_pack(*ptrs, fs, F_d, heads, D, float(slope), int(flags), out.data_ptr() + 4 * n * r,
                                          _lib.stream()), "ubs_gatv2_rel_pack")
    return out


def agent_act_rel_supported(dims: "AgentDims", heads, F_gt, cap_gt, F_ubs, cap_ubs, F_d) -> bool:
    if not (dims.supported() and dims.aggr and dims.Fin == 2 * dims.H):
        return False
    return bool(_lib.load().ubs_agent_act_rel_supported(dims.H, dims.M, dims.K, dims.A, dims.U, dims.flags, heads, F_gt, cap_gt,
                                                        F_ubs, cap_ubs, F_d))


def agent_act_rel(dims: "AgentDims", packed, relpack, x_gt_ptr, ip_seen_ptr, F_gt, cap_gt, x_ubs_ptr, ip_near_ptr, F_ubs,
                  cap_ubs, x_agent_ptr, F_d, heads, gat_flags, h0, mask, h_out, q, acts=None, explore=None):
    """ONE launch per act vector-step (``ubs_agent_act_rel_fwd``): both GATv2 observation relations read from raw packet
    addresses + aggregator + TarMAC / GRU + Q head + argmax + epsilon-greedy.  ``h0 (N,H)`` -> ``h_out (N,H)``,
    ``q (N,A)``, ``acts (N,) int64``; every output is a caller-provided contiguous buffer."""
    lib = _lib.load()
    _lib.require_cuda(packed, relpack, h0, h_out, q)
    N = h0.shape[0]
    eg_u, eg_a, eg_eps = explore if explore is not None else (None, None, None)
    with _timed("agent_act_rel", (1, N, dims.ints(), (F_gt, cap_gt, F_ubs, cap_ubs, heads))):
        _lib.check(lib.ubs_agent_act_rel_fwd(dims.H, dims.M, dims.K, dims.A, dims.U, dims.flags, _lib.ptr(packed),
                                             _lib.ptr(relpack), x_gt_ptr, ip_seen_ptr, F_gt, cap_gt, x_ubs_ptr, ip_near_ptr,
                                             F_ubs, cap_ubs, x_agent_ptr, F_d, heads, int(gat_flags), _lib.ptr(h0),
                                             _lib.ptr(mask), _lib.ptr(h_out), _lib.ptr(q), _lib.ptr(acts), _lib.ptr(eg_u),
                                             _lib.ptr(eg_a), _lib.ptr(eg_eps), N, _lib.stream()), "ubs_agent_act_rel_fwd")
    return q


class AgentSequence(th.autograd.Function):
    """Whole-sequence forward / backward of the recurrent part of the agent.

    ``forward(xin (T,N,Fin), h0 (N,H), mask (T,N) int32|None, dims, packed, *params[PARAM_ORDER])``
    returns ``(q (T,N,A), h_last (N,H), h_all (T,N,H) [non-differentiable])``.  The backward walks the sequence in
    reverse inside one kernel and forms every parameter gradient with batched library GEMMs over all T·N rows."""

    @staticmethod
    def forward(ctx, xin, h0, mask, dims, packed, *params):
        lib = _lib.load()
        _lib.require_cuda(xin, h0, packed)
        xin, h0 = _f32c(xin), _f32c(h0)
        T, N = xin.shape[0], xin.shape[1]
        dev = xin.device
        f32 = dict(dtype=th.float32, device=dev)
        h_out = th.empty(T, N, dims.H, **f32)
        q = th.empty(T, N, dims.A, **f32)
        sv_xc = th.empty(T, N, dims.I, **f32)
        sv_gate = th.empty(T, N, 4 * dims.H, **f32)
        sv_vsq = th.empty(T, N, dims.Vp, **f32) if dims.tarmac else None
        sv_alpha = th.empty(T, N, dims.U, **f32) if dims.tarmac else None
        with _timed("agent_seq_fwd", (T, N, dims.ints(), True)):
            _lib.check(lib.ubs_agent_seq_fwd(*dims.ints(), _lib.ptr(packed), _lib.ptr(xin), _lib.ptr(h0), _lib.ptr(mask),
                                             _lib.ptr(h_out), _lib.ptr(q), None, _lib.ptr(sv_xc), _lib.ptr(sv_vsq),
                                             _lib.ptr(sv_alpha), _lib.ptr(sv_gate), N, T, _lib.stream()),
                       "ubs_agent_seq_fwd")
        ctx.dims = dims
        ctx.has = [p is not None for p in params]
        ctx.save_for_backward(xin, h0, packed, h_out, sv_xc, sv_vsq, sv_alpha, sv_gate)
        ctx.mark_non_differentiable(h_out)
        return q, h_out[T - 1].clone(), h_out

    @staticmethod
    def backward(ctx, dq, dh_last, _dh_all):
        xin, h0, packed, h_out, sv_xc, sv_vsq, sv_alpha, sv_gate = ctx.saved_tensors
        dims = ctx.dims
        lib = _lib.load()
        T, N = xin.shape[0], xin.shape[1]
        H, M, K, A, I, V, Vp = dims.H, dims.M, dims.K, dims.A, dims.I, dims.V, dims.Vp
        f32 = dict(dtype=th.float32, device=xin.device)
        dq = _f32c(dq) if dq is not None else th.zeros(T, N, A, **f32)
        dh_last = _f32c(dh_last) if dh_last is not None else None
        d_xin = th.empty_like(xin)
        d_h0 = th.empty_like(h0) if ctx.needs_input_grad[1] else None
        st_dgi, st_dgh = th.empty(T, N, 3 * H, **f32), th.empty(T, N, 3 * H, **f32)
        st_dvsq = th.empty(T, N, Vp, **f32) if dims.tarmac else None
        st_dpre = th.empty(T, N, H, **f32) if dims.aggr else None
        with _timed("agent_seq_bwd", (T, N, dims.ints(), True)):
            _lib.check(lib.ubs_agent_seq_bwd(*dims.ints(), _lib.ptr(packed), _lib.ptr(h0), _lib.ptr(h_out),
                                             _lib.ptr(sv_xc), _lib.ptr(sv_vsq), _lib.ptr(sv_alpha), _lib.ptr(sv_gate),
                                             _lib.ptr(dq), _lib.ptr(dh_last), _lib.ptr(d_xin), _lib.ptr(d_h0),
                                             _lib.ptr(st_dgi), _lib.ptr(st_dgh), _lib.ptr(st_dvsq), _lib.ptr(st_dpre),
                                             N, T, _lib.stream()), "ubs_agent_seq_bwd")
        # parameter gradients: batched fp32 library GEMMs over all T*N rows
        TN = T * N
        xc, dgi, dgh = sv_xc.view(TN, I), st_dgi.view(TN, 3 * H), st_dgh.view(TN, 3 * H)
        hprev = th.cat((h0.unsqueeze(0), h_out[:-1]), 0).view(TN, H)
        g = {}
        g["W_ih"], g["b_ih"] = dgi.t() @ xc, dgi.sum(0)
        g["W_hh"], g["b_hh"] = dgh.t() @ hprev, dgh.sum(0)
        dq2 = dq.view(TN, A)
        g["W_out"], g["b_out"] = dq2.t() @ h_out.view(TN, H), dq2.sum(0)
        if dims.tarmac:
            dv = st_dvsq.view(TN, Vp)
            gw = th.cat((dv.t() @ xc[:, :H], dv.t() @ hprev), 1)           # (Vp, 2H) for inputs [x ‖ h]
            gb = dv.sum(0)
            g["W_val"], g["b_val"] = gw[:M], gb[:M]
            g["W_sign"], g["b_sign"] = gw[M:M + K], gb[M:M + K]
            g["W_que"], g["b_que"] = gw[M + K:V], gb[M + K:V]
        if dims.aggr:
            dp = st_dpre.view(TN, H)
            g["W_aggr"], g["b_aggr"] = dp.t() @ xin.view(TN, dims.Fin), dp.sum(0)
        grads = tuple(g.get(k) if has else None for k, has in zip(PARAM_ORDER, ctx.has))
        return (d_xin if ctx.needs_input_grad[0] else None, d_h0, None, None, None) + grads


# ====================================================================================================================
# Strided-segment relation encoder (ubs_gatv2_seg_fwd / ubs_gatv2_seg_bwd): all timesteps of an arena in one launch per
# relation, every relation writing its own column block of ONE (rows, R*H) output buffer.
_SIDE_STREAMS = {}
SAVE_SCORES = True      # training forward keeps the per-(edge, head) attention scores for the backward (False: recompute)
# SegmentEncode: the relations of one encode run on forked streams, forward and backward (UBS_FORK_RELATIONS=0: in order)
FORK_RELATIONS = os.environ.get("UBS_FORK_RELATIONS", "1") != "0"


def _side_stream(dev, i):
    key = (dev.index, i)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = th.cuda.Stream(device=dev, priority=-1)     # high priority: the short relation goes first
    return _SIDE_STREAMS[key]


class _nullctx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class RelSpec:
    """One relation of a strided-segment encode: raw device addresses + strides (elements) + feature widths."""

    def __init__(self, x_src_ptr, st_xsrc, F_s, indptr_ptr, st_ip, n_edges_hint, src_idx_ptr=None, st_sidx=0):
        self.x_src_ptr, self.st_xsrc, self.F_s = x_src_ptr, st_xsrc, F_s
        self.indptr_ptr, self.st_ip, self.n_edges_hint = indptr_ptr, st_ip, n_edges_hint
        self.src_idx_ptr, self.st_sidx = src_idx_ptr, st_sidx


def _fork_pays(rows):
    """Act-step-sized launches cannot fill the chip and full windows are device-bound: forking helps both.  In between
    (a 32-env window) the update is bound by the host enqueueing launches and the extra stream waits only cost."""
    return rows <= 8192 or rows >= 32768


class SegmentEncode(th.autograd.Function):
    """``forward(keepalive, specs, x_dst_ptr, st_xdst, F_d, n_seg, n_dst_seg, heads, D, slope, flags, *params)``
    with ``params`` = 7 tensors per relation ``(W_src, b_src, W_dst, b_dst, attn, W_res, b_res)``.
    Returns ``(n_seg * n_dst_seg, len(specs) * H)``.  Observations are leaves: only parameter gradients are produced.
    ``keepalive`` is the tensor that owns the addressed memory (the arena buffer)."""

    @staticmethod
    def forward(ctx, keepalive, specs, x_dst_ptr, st_xdst, F_d, n_seg, n_dst_seg, heads, D, slope, flags, *params):
        lib = _lib.load()
        dev = keepalive.device
        H, R = heads * D, len(specs)
        rows = n_seg * n_dst_seg
        need_grad = any(p is not None and p.requires_grad for p in params)      # callers detach under no_grad
        out = th.empty(rows, R * H, dtype=th.float32, device=dev)
        stats = th.empty(R, 2, rows, heads, dtype=th.float32, device=dev) if need_grad else None
        # training: the raw attention score of every (edge slot, head) is kept for the backward (16 B per edge at 4
        # heads — as much as the source row itself) instead of being recomputed from the H-channel projection
        caps = [max(int(sp.n_edges_hint) // max(int(n_seg), 1), 1) for sp in specs]
        scores = [th.empty(n_seg * cap * heads, dtype=th.float32, device=dev) if (need_grad and SAVE_SCORES) else None
                  for cap in caps]
        ps = [_f32c(p.detach()) if p is not None else None for p in params]
        # relations run side by side on a forked stream: small launches (the act step) cannot fill the chip, and in a
        # window the short relation (`near`: 7 edges per destination, most lanes idle) fills issue slots the long one leaves
        fork = R > 1 and FORK_RELATIONS and TIMER is None and _fork_pays(rows)
        cur = th.cuda.current_stream()
        # forked: the side-stream relations (short: `near`) are enqueued first on a high-priority stream so that their
        # CTAs are resident next to the long relation's instead of queueing behind them
        order = (list(range(1, R)) + [0]) if fork else list(range(R))
        sides = []
        for r in order:
            sp = specs[r]
            W = ps[7 * r:7 * r + 7]
            side = _side_stream(dev, r) if (fork and r > 0) else None
            if side is not None:
                side.wait_stream(cur)
            with _timed("gatv2_fwd", (rows, sp.n_edges_hint, sp.F_s, F_d, heads, D, need_grad)), \
                    (th.cuda.stream(side) if side is not None else _nullctx()):
                _lib.check(lib.ubs_gatv2_seg_fwd_scores(
                    sp.x_src_ptr, x_dst_ptr, sp.indptr_ptr, sp.src_idx_ptr, *[_lib.ptr(w) for w in W],
                    out.data_ptr() + 4 * r * H, _lib.ptr(stats[r, 0]) if need_grad else None,
                    _lib.ptr(stats[r, 1]) if need_grad else None, _lib.ptr(scores[r]), caps[r], n_seg, n_dst_seg,
                    sp.n_edges_hint, sp.st_xsrc, st_xdst, sp.st_ip, sp.st_sidx, R * H, sp.F_s, F_d, heads, D, float(slope),
                    int(flags), _lib.stream()), "ubs_gatv2_seg_fwd")
            if side is not None:
                sides.append(side)
        for side in sides:
            cur.wait_stream(side)
        if need_grad:
            ctx.save_for_backward(keepalive, out, stats, *[p for p in ps if p is not None])
            ctx.scores, ctx.caps = scores, caps                 # plain buffers (not autograd inputs / outputs)
            ctx.cfg = (specs, x_dst_ptr, st_xdst, F_d, n_seg, n_dst_seg, heads, D, float(slope), int(flags),
                       [p is not None for p in params], [None if p is None else tuple(p.shape) for p in params])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        keepalive, out, stats, *saved = ctx.saved_tensors
        specs, x_dst_ptr, st_xdst, F_d, n_seg, n_dst_seg, heads, D, slope, flags, has, shapes = ctx.cfg
        lib = _lib.load()
        dev = out.device
        H, R = heads * D, len(specs)
        rows = n_seg * n_dst_seg
        grad_out = _f32c(grad_out)
        it = iter(saved)
        ps = [next(it) if h else None for h in has]
        fork = R > 1 and FORK_RELATIONS and TIMER is None and _fork_pays(rows)
        cur = th.cuda.current_stream()
        bufs = []                                           # allocated on THIS stream, whichever stream fills them
        for sp in specs:
            bufs.append((th.empty(H * (sp.F_s + 2 * F_d + 4), dtype=th.float32, device=dev),
                         th.empty(int(lib.ubs_gatv2_bwd_workspace(rows, sp.F_s, F_d, heads, D)), dtype=th.float32, device=dev)))
        sides = []
        for r in ((list(range(1, R)) + [0]) if fork else range(R)):
            sp = specs[r]
            W = ps[7 * r:7 * r + 7]
            gparams, ws = bufs[r]
            side = _side_stream(dev, r) if (fork and r > 0) else None
            if side is not None:
                side.wait_stream(cur)
                sides.append(side)
            with _timed("gatv2_bwd", (rows, sp.n_edges_hint, sp.F_s, F_d, heads, D, True)), \
                    (th.cuda.stream(side) if side is not None else _nullctx()):
                _lib.check(lib.ubs_gatv2_seg_bwd_scores(
                    sp.x_src_ptr, x_dst_ptr, sp.indptr_ptr, sp.src_idx_ptr, *[_lib.ptr(w) for w in W],
                    out.data_ptr() + 4 * r * H, grad_out.data_ptr() + 4 * r * H, _lib.ptr(stats[r, 0]),
                    _lib.ptr(stats[r, 1]), _lib.ptr(ctx.scores[r]), ctx.caps[r], _lib.ptr(gparams), None, None, _lib.ptr(ws),
                    n_seg, n_dst_seg,
                    sp.n_edges_hint, sp.st_xsrc, st_xdst, sp.st_ip, sp.st_sidx, R * H, R * H, sp.F_s, F_d, heads, D,
                    slope, flags, _lib.stream()), "ubs_gatv2_seg_bwd")
        for side in sides:
            cur.wait_stream(side)
        grads = []
        for r, sp in enumerate(specs):
            gparams, o = bufs[r][0], 0
            for n, idx in ((H * sp.F_s, 0), (H, 1), (H * F_d, 2), (H, 3), (H, 4), (H * F_d, 5), (H, 6)):
                k = 7 * r + idx
                grads.append(gparams[o:o + n].view(shapes[k]) if has[k] else None)
                o += n
        return (None,) * 11 + tuple(grads)


# ====================================================================================================================
# Sequence path v2: batched GEMMs for everything that depends on the observation only + a persistent kernel whose
# recurrent weights stay in shared memory (ubs_agent_seq2_fwd / ubs_agent_seq2_bwd).
def seq2_supported(dims: AgentDims) -> bool:
    lib = _lib.load()
    if not dims.supported():
        return False
    return max(int(lib.ubs_agent_seq2_smem_bytes(dims.H, dims.M, dims.K, dims.U, dims.flags, b)) for b in (0, 1)) <= 227 * 1024


def tc_linear_supported(K: int, N: int) -> bool:
    return K % 32 == 0 and K >= 32 and N % 16 == 0 and 16 <= N <= 256 and 2 * (K // 32) * N * 128 + 2 * 32768 + 2048 <= 227 * 1024


def tc_linear(x, w, bias=None, relu=False, out=None):
    """``act(x @ w.T + bias)`` on the tensor cores with 3xTF32 (``ubs_tf32x3_gemm``): fp32-accurate.  ``x (M,K)`` and
    ``w (N,K)`` may be row-strided views (last dim contiguous, 16-byte aligned rows); no autograd."""
    lib = _lib.load()
    _lib.require_cuda(x, w)
    M, K = x.shape
    N = w.shape[0]
    if x.stride(1) != 1:
        x = x.contiguous()
    if w.stride(1) != 1:
        w = w.contiguous()
    if out is None:
        out = th.empty(M, N, dtype=th.float32, device=x.device)
    with _timed("tf32x3_gemm", (M, N, K)):
        _lib.check(lib.ubs_tf32x3_gemm(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _lib.ptr(bias),
                                       out.data_ptr(), out.stride(0), M, N, K, int(relu), _lib.stream()),
                   "ubs_tf32x3_gemm")
    return out


USE_TC = True       # batched observation-side projections on the tensor cores (3xTF32); False: cuBLAS fp32
USE_TC_TN = True    # weight-gradient products (A^T B over the T*N rows) on the tensor cores (3xTF32)
_TN_WS = {}


def tc_matmul_tn(a, b, out=None):
    """``a.T @ b`` for ``a (R, Mo)``, ``b (R, No)`` (row-strided views allowed) with the reduction over the R rows on
    the tensor cores (``ubs_tf32x3_gemm_tn``, 3xTF32: fp32-accurate, deterministic).  No autograd."""
    lib = _lib.load()
    _lib.require_cuda(a, b)
    R, Mo = a.shape
    No = b.shape[1]
    if a.stride(1) != 1:
        a = a.contiguous()
    if b.stride(1) != 1:
        b = b.contiguous()
    if out is None:
        out = th.empty(Mo, No, dtype=th.float32, device=a.device)
    n = int(lib.ubs_tf32x3_gemm_tn_workspace(R, Mo, No))
    key = (a.device, th.cuda.current_stream().cuda_stream)
    ws = _TN_WS.get(key)
    if ws is None or ws.numel() < n:
        ws = th.empty(n, dtype=th.float32, device=a.device)
        _TN_WS[key] = ws
    with _timed("tf32x3_gemm_tn", (R, Mo, No)):
        _lib.check(lib.ubs_tf32x3_gemm_tn(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(),
                                          out.stride(0), ws.data_ptr(), R, Mo, No, _lib.stream()), "ubs_tf32x3_gemm_tn")
    return out


_COLSUM_WS = {}


def _colsum_ok(x):
    return (x.is_cuda and x.dtype == th.float32 and x.dim() == 2 and x.stride(1) == 1 and x.shape[1] % 4 == 0
            and 4 <= x.shape[1] <= 1024 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0)


def _colsum_ws(x):
    key = (x.device, th.cuda.current_stream().cuda_stream)
    n = int(_lib.load().ubs_colsum_workspace(x.shape[1]))
    ws = _COLSUM_WS.get(key)
    if ws is None or ws.numel() < n:
        ws = th.empty(n, dtype=th.float32, device=x.device)
        _COLSUM_WS[key] = ws
    return ws


def colsum(x):
    """``x.sum(0)`` of a (row-strided) ``(R, C)`` fp32 matrix: the bias gradients of a window (``ubs_colsum``, fixed
    summation order).  Shapes outside the kernel (C % 4 != 0, unaligned rows) go to ``Tensor.sum``.  No autograd."""
    if not _colsum_ok(x):
        return x.sum(0)
    lib = _lib.load()
    out = th.empty(x.shape[1], dtype=th.float32, device=x.device)
    with _timed("colsum", (x.shape[0], x.shape[1])):
        _lib.check(lib.ubs_colsum(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], out.data_ptr(), _colsum_ws(x).data_ptr(),
                                  _lib.stream()), "ubs_colsum")
    return out


def relu_bwd_colsum_(dy, y):
    """In place ``dy *= (y > 0)`` and the column sums of the result, one pass (``ubs_relu_bwd_colsum``): ReLU backward
    of the aggregator layer + its bias gradient.  Returns ``(dy, colsum)``.  No autograd."""
    if not (_colsum_ok(dy) and _colsum_ok(y) and dy.shape == y.shape):
        dy.mul_(y > 0)
        return dy, dy.sum(0)
    lib = _lib.load()
    out = th.empty(dy.shape[1], dtype=th.float32, device=dy.device)
    with _timed("relu_bwd_colsum", (dy.shape[0], dy.shape[1])):
        _lib.check(lib.ubs_relu_bwd_colsum(dy.data_ptr(), dy.stride(0), y.data_ptr(), y.stride(0), dy.data_ptr(), dy.stride(0),
                                           dy.shape[0], dy.shape[1], out.data_ptr(), _colsum_ws(dy).data_ptr(), _lib.stream()),
                   "ubs_relu_bwd_colsum")
    return dy, out


def matmul_tn(a, b):
    """``a.T @ b``: the tcgen05 kernel for the long-reduction weight-gradient shapes, else the fp32 library GEMM."""
    if USE_TC_TN and a.is_cuda and a.shape[0] >= 4096 and b.shape[1] % 16 == 0 and 16 <= b.shape[1] <= 128 \
            and a.dtype == th.float32 and b.dtype == th.float32:
        Mo = a.shape[1]
        if Mo % 4 or a.stride(0) % 4 or a.stride(1) != 1:
            # rows that are not 16-byte granular (the Q head: n_actions = 9) would take the kernel's scalar producer
            # path: one small padded copy keeps the long reduction on the vector path
            a = th.nn.functional.pad(a, (0, (-Mo) % 4))
        return tc_matmul_tn(a, b)[:Mo]
    return a.t() @ b


def dense(x, w, bias=None, relu=False):
    """``act(x @ w.T + bias)`` for the batched (T*N-row) projections: the tcgen05 3xTF32 kernel when the shape fits
    (wide outputs are split into column blocks), else the fp32 library GEMM.  No autograd."""
    M, K = x.shape
    N = w.shape[0]
    if USE_TC and x.is_cuda and M >= 512:
        blocks = None
        if tc_linear_supported(K, N):
            blocks = [(0, N)]
        elif N % 32 == 0 and tc_linear_supported(K, N // 2):
            blocks = [(0, N // 2), (N // 2, N)]
        if blocks is not None and x.data_ptr() % 16 == 0 and x.stride(0) % 4 == 0 and x.stride(1) == 1 and w.is_contiguous():
            out = th.empty(M, N, dtype=th.float32, device=x.device)
            for lo, hi in blocks:
                tc_linear(x, w[lo:hi], None if bias is None else bias[lo:hi], relu=relu, out=out[:, lo:hi])
            return out
    out = th.addmm(bias, x, w.t()) if bias is not None else x @ w.t()
    return th.relu_(out) if relu else out


class TCLinear(th.autograd.Function):
    """``x @ w.T + b`` with autograd, every product on the tensor cores when the shape fits (3xTF32, fp32-accurate):
    forward and ``grad_x`` through ``ubs_tf32x3_gemm``, ``grad_w = grad.T @ x`` through ``ubs_tf32x3_gemm_tn`` — the dense
    per-relation feature projection of wide-input GATv2 relations (the synthetic sweep).  Falls back to the fp32 library
    GEMM per product where a shape is outside the kernels (``dense`` / ``matmul_tn`` decide)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = _f32c(x), _f32c(w)
        ctx.save_for_backward(x, w)
        ctx.has_b = b is not None
        return dense(x, w, b)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = _f32c(g)
        gx = dense(g, w.t().contiguous()) if ctx.needs_input_grad[0] else None
        gw = matmul_tn(g, x) if ctx.needs_input_grad[1] else None
        gb = g.sum(0) if ctx.has_b and ctx.needs_input_grad[2] else None
        return gx, gw, gb


def linear(x, w, b=None):
    """``F.linear`` for the wide projections: tensor cores (3xTF32) for CUDA inputs with >= 512 rows, else the library."""
    if x.is_cuda and x.dim() == 2 and x.shape[0] >= 512 and x.dtype == th.float32:
        return TCLinear.apply(x, w, b)
    return th.nn.functional.linear(x, w, b)


class Seq2Weights:
    """Derived weight tensors of the resident-weight sequence path, rebuilt only when a parameter changes:
    ``Wx (Vp+3H, H)`` = ``[W_vsq[:, :H]; W_ih[:, :H]]`` and its bias (ONE observation-side GEMM gives ``[pv | pg]``),
    the K-major recurrent blocks the forward kernel keeps in shared memory, and the original-layout blocks of the
    backward kernel."""

    def __init__(self, dims, params):
        H, M, K = dims.H, dims.M, dims.K
        p = {k: (None if v is None else v.detach()) for k, v in params.items()}
        W_ih, W_hh = p["W_ih"], p["W_hh"]
        self.b_hh = p["b_hh"].contiguous()
        self.wt_hh = W_hh.t().contiguous()                       # (H, 3H)
        self.w_hh = W_hh.contiguous()                            # (3H, H)
        if dims.tarmac:
            Wv = th.cat((p["W_val"], p["W_sign"], p["W_que"]), 0)
            bv = th.cat((p["b_val"], p["b_sign"], p["b_que"]), 0)
            if dims.Vp != dims.V:
                Wv = th.cat((Wv, Wv.new_zeros(dims.Vp - dims.V, Wv.shape[1])), 0)
                bv = th.cat((bv, bv.new_zeros(dims.Vp - dims.V)), 0)
            self.Wx = th.cat((Wv[:, :H], W_ih[:, :H]), 0).contiguous()          # (Vp + 3H, H)
            self.bx = th.cat((bv, p["b_ih"]), 0)
            self.wt_vsq_h = Wv[:, H:].t().contiguous()           # (H, Vp)
            self.wt_ih_c = W_ih[:, H:].t().contiguous()          # (M, 3H)
            self.w_ih_c = W_ih[:, H:].contiguous()               # (3H, M)
            # dx = [dgi | dv] @ [W_ih[:, :H]; W_vsq[:, :H]]
            self.Wdx = th.cat((W_ih[:, :H], Wv[:, :H]), 0).contiguous()         # (3H + Vp, H)
        else:
            self.Wx, self.bx = W_ih[:, :H].contiguous(), p["b_ih"]
            self.wt_vsq_h = self.wt_ih_c = self.w_ih_c = None
            self.Wdx = self.Wx
        self.Wdx_t = self.Wdx.t().contiguous()                   # (H, 3H + Vp): nn.Linear layout of the dx projection
        self.W_aggr, self.b_aggr = p["W_aggr"], p["b_aggr"]
        self.W_aggr_t = None if p["W_aggr"] is None else p["W_aggr"].t().contiguous()      # (Fin, H)
        self.W_out, self.b_out = p["W_out"], p["b_out"]


_SEQ2_CACHE = {}


def seq2_weights(dims, params) -> Seq2Weights:
    key = tuple((t.data_ptr(), t._version) for t in params.values() if t is not None)
    slot = (dims.ints(), key[0][0])
    hit = _SEQ2_CACHE.get(slot)
    if hit is None or hit[0] != key:
        if len(_SEQ2_CACHE) > 16:
            _SEQ2_CACHE.clear()
        hit = (key, Seq2Weights(dims, params))
        _SEQ2_CACHE[slot] = hit
    return hit[1]


def _seq2_forward(dims, W: Seq2Weights, xg, h0, mask, training):
    lib = _lib.load()
    T, N = xg.shape[0], xg.shape[1]
    TN, H, M, K, U, Vp = T * N, dims.H, dims.M, dims.K, dims.U, dims.Vp
    f32 = dict(dtype=th.float32, device=xg.device)
    xg2 = xg.reshape(TN, dims.Fin)
    x = dense(xg2, W.W_aggr, W.b_aggr, relu=True) if dims.aggr else xg2
    pvg = dense(x, W.Wx, W.bx)                                     # (TN, Vp + 3H) = [pv | pg], one projection
    ld = pvg.shape[1]
    h_out = th.empty(T, N, H, **f32)
    sv_gate = th.empty(T, N, 4 * H, **f32) if training else None
    sv_vsq = th.empty(T, N, Vp, **f32) if training and dims.tarmac else None
    sv_alpha = th.empty(T, N, U, **f32) if training and dims.tarmac else None
    sv_c = th.empty(T, N, M, **f32) if training and dims.tarmac else None
    P = _lib.ptr
    pv_ptr = pvg.data_ptr() if dims.tarmac else None
    pg_ptr = pvg.data_ptr() + 4 * (Vp if dims.tarmac else 0)
    with _timed("agent_seq2_fwd", (T, N, dims.ints(), training)):
        _lib.check(lib.ubs_agent_seq2_fwd(H, M, K, U, dims.flags, P(W.wt_vsq_h), P(W.wt_ih_c), P(W.wt_hh), P(W.b_hh),
                                          pv_ptr, pg_ptr, P(h0), P(mask), P(h_out), P(sv_vsq), P(sv_alpha), P(sv_c),
                                          P(sv_gate), ld, ld, N, T, _lib.stream()), "ubs_agent_seq2_fwd")
    q = th.addmm(W.b_out, h_out.view(TN, H), W.W_out.t()).view(T, N, dims.A)
    return q, h_out, (x, sv_vsq, sv_alpha, sv_c, sv_gate)


def agent_seq2_infer(dims: AgentDims, params: dict, xg, h0, mask):
    """Inference through the resident-weight sequence kernel: returns ``q (T,N,A)``, ``h_out (T,N,H)``."""
    _lib.require_cuda(xg, h0)
    q, h_out, _ = _seq2_forward(dims, seq2_weights(dims, params), _f32c(xg), _f32c(h0), mask, False)
    return q, h_out


class AgentSequence2(th.autograd.Function):
    """Same contract as :class:`AgentSequence` (``forward(xg, h0, mask, dims, *params[PARAM_ORDER])`` ->
    ``(q, h_last, h_all)``) on the resident-weight kernels.  All parameter gradients come from a handful of batched
    GEMMs over the T*N rows: the backward kernel writes ONE stash ``[dgi | dvsq | dgh]`` whose column blocks feed
    ``dx``, ``[dW_ih_x; dW_vsq_x]``, ``[dW_vsq_h; dW_hh]`` and all bias gradients without copies."""

    @staticmethod
    def forward(ctx, xg, h0, mask, dims, *params):
        _lib.require_cuda(xg, h0)
        xg, h0 = _f32c(xg), _f32c(h0)
        W = seq2_weights(dims, dict(zip(PARAM_ORDER, params)))
        q, h_out, (x, sv_vsq, sv_alpha, sv_c, sv_gate) = _seq2_forward(dims, W, xg, h0, mask, True)
        ctx.dims, ctx.W = dims, W
        ctx.has = [v is not None for v in params]
        ctx.save_for_backward(xg, h0, h_out, x, sv_vsq, sv_alpha, sv_c, sv_gate)
        ctx.mark_non_differentiable(h_out)
        return q, h_out[-1].clone(), h_out

    @staticmethod
    def backward(ctx, dq, dh_last, _dh_all):
        xg, h0, h_out, x, sv_vsq, sv_alpha, sv_c, sv_gate = ctx.saved_tensors
        dims, W = ctx.dims, ctx.W
        lib = _lib.load()
        T, N = xg.shape[0], xg.shape[1]
        TN, H, M, K, U, A, V, Vp = T * N, dims.H, dims.M, dims.K, dims.U, dims.A, dims.V, dims.Vp
        H3 = 3 * H
        f32 = dict(dtype=th.float32, device=xg.device)
        dq2 = (_f32c(dq) if dq is not None else th.zeros(T, N, A, **f32)).view(TN, A)
        dhq = (dq2 @ W.W_out).view(T, N, H)
        if dh_last is not None:
            dhq[T - 1] += dh_last
        ldS = H3 + Vp + H3                                         # [dgi | dvsq | dgh]
        S = th.empty(TN, ldS, **f32)
        d_h0 = th.empty_like(h0) if ctx.needs_input_grad[1] else None
        P = _lib.ptr
        sp = S.data_ptr()
        with _timed("agent_seq2_bwd", (T, N, dims.ints(), True)):
            _lib.check(lib.ubs_agent_seq2_bwd(H, M, K, U, dims.flags, P(W.w_hh), P(W.w_ih_c), P(h0), P(h_out), P(sv_vsq),
                                              P(sv_alpha), P(sv_gate), P(dhq), sp, sp + 4 * (H3 + Vp),
                                              sp + 4 * H3 if dims.tarmac else None, P(d_h0), ldS, N, T, _lib.stream()),
                       "ubs_agent_seq2_bwd")
        hprev = th.cat((h0.unsqueeze(0), h_out[:-1]), 0).view(TN, H)
        Sx, Sh = S[:, :H3 + Vp], S[:, H3:]                          # [dgi | dvsq] and [dvsq | dgh]
        gb = colsum(S)
        g = {"b_ih": gb[:H3], "b_hh": gb[H3 + Vp:]}
        g["W_out"], g["b_out"] = matmul_tn(dq2, h_out.view(TN, H)), dq2.sum(0)
        Gx = matmul_tn(Sx, x)                                       # (3H + Vp, H): [dW_ih[:, :H]; dW_vsq[:, :H]]
        Gh = matmul_tn(Sh, hprev)                                   # (Vp + 3H, H): [dW_vsq[:, H:]; dW_hh]
        g["W_hh"] = Gh[Vp:]
        dx = dense(Sx, W.Wdx_t)                                     # (TN, H) = [dgi | dvsq] @ [W_ih[:, :H]; W_vsq[:, :H]]
        if dims.tarmac:
            g["W_ih"] = th.cat((Gx[:H3], matmul_tn(S[:, :H3], sv_c.view(TN, M))), 1)
            gw = th.cat((Gx[H3:], Gh[:Vp]), 1)                      # (Vp, 2H)
            gbv = gb[H3:H3 + Vp]
            g["W_val"], g["b_val"] = gw[:M], gbv[:M]
            g["W_sign"], g["b_sign"] = gw[M:M + K], gbv[M:M + K]
            g["W_que"], g["b_que"] = gw[M + K:V], gbv[M + K:V]
        else:
            g["W_ih"] = Gx[:H3]
        if dims.aggr:
            dpre, g["b_aggr"] = relu_bwd_colsum_(dx, x)
            xg2 = xg.view(TN, dims.Fin)
            g["W_aggr"] = matmul_tn(dpre, xg2)
            d_xg = dense(dpre, W.W_aggr_t).view(T, N, dims.Fin) if ctx.needs_input_grad[0] else None
        else:
            d_xg = dx.view(T, N, dims.Fin) if ctx.needs_input_grad[0] else None
        grads = tuple(g.get(k) if has else None for k, has in zip(PARAM_ORDER, ctx.has))
        return (d_xg, d_h0, None, None) + grads
