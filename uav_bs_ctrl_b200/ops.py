"""``torch.autograd.Function`` wrappers around the C ABI (``include/ubs_gnn.h``).

Each wrapper only checks dtype / contiguity / device, allocates outputs and workspaces through PyTorch's
caching allocator, and enqueues the kernels on the current CUDA stream.  CUDA tensors only.
"""
from __future__ import annotations

import torch as th

from . import _lib

GAT_RESIDUAL, GAT_RELU = 1, 2


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
