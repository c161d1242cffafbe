"""Data-parallel plumbing: one process per GPU, flat-bucket gradient all-reduce.

API shape of reference ``utils/mpi_pytorch.py:19-35`` (``mpi_avg_grads`` / ``sync_params``; there implemented with
mpi4py over per-parameter CPU numpy buffers and never called by ``algos/*``).  Here: ``torch.distributed`` with the
NCCL backend over NVLink/NVSwitch (``gloo`` on CPU for tests).  The path shards by env instance / sampled sequence
— graphs are block-diagonal per env, no edge crosses envs — so the only exchange is ONE all-reduce per update over a
single flat fp32 bucket holding every parameter gradient (≈240 KB at H=64), issued between ``loss.backward()`` and
``clip_grad_value_`` (reference ``algos/madrqn/learner.py:158-159``): clipping is element-wise, so it has to see the
averaged gradient to match single-process semantics.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch as th
import torch.distributed as td


def is_dist() -> bool:
    return td.is_available() and td.is_initialized()


def world_size() -> int:
    return td.get_world_size() if is_dist() else 1


def rank() -> int:
    return td.get_rank() if is_dist() else 0


def init_from_env(backend: Optional[str] = None) -> int:
    """Initialises the default process group from torchrun's RANK / WORLD_SIZE / MASTER_* (no-op for 1 process).
    Returns the local rank."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1 and not is_dist():
        if backend is None:
            backend = "nccl" if th.cuda.is_available() else "gloo"
        if backend == "nccl":
            th.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        td.init_process_group(backend=backend)
    return local


class FlatGradBucket:
    """All parameter gradients of a module as views into one contiguous fp32 buffer.

    ``p.grad`` of every parameter aliases a slice of ``self.flat``, so autograd accumulates straight into the
    bucket and the all-reduce needs no flatten / unflatten copies."""

    def __init__(self, params: Iterable[th.nn.Parameter]):
        self.params: List[th.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        n = sum(p.numel() for p in self.params)
        self.flat = th.zeros(n, dtype=dt, device=dev)
        self._zeros = None
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero_(self):
        """Replaces ``optimizer.zero_grad()`` (which would drop the views when ``set_to_none=True``)."""
        self.flat.zero_()

    def rebind(self):
        o = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + o * self.flat.element_size():
                g = self.flat[o:o + p.numel()].view_as(p)
                if p.grad is not None:
                    g.copy_(p.grad)
                p.grad = g
            o += p.numel()

    def release(self):
        """Before ``backward()``: drop the views, so that autograd hands every parameter its gradient tensor instead of
        launching one ``grad += g`` kernel per parameter into a zeroed bucket."""
        for p in self.params:
            p.grad = None

    def gather(self):
        """After ``backward()``: ONE concatenation of the produced gradients into the flat buffer (zeros where a
        parameter got none), then the parameters' ``.grad`` are the bucket views again."""
        if self._zeros is None:
            self._zeros = th.zeros_like(self.flat)
        parts, o = [], 0
        for p in self.params:
            n = p.numel()
            parts.append(self._zeros[o:o + n] if p.grad is None else p.grad.reshape(-1))
            o += n
        th.cat(parts, out=self.flat)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def all_reduce_mean(self):
        if world_size() > 1:
            td.all_reduce(self.flat, op=td.ReduceOp.SUM)
            self.flat.mul_(1.0 / world_size())


def avg_grads(module_or_bucket):
    """Average gradients across processes (``mpi_avg_grads``).  Accepts a module (per-call flatten) or a bucket."""
    if world_size() == 1:
        return
    if isinstance(module_or_bucket, FlatGradBucket):
        module_or_bucket.all_reduce_mean()
        return
    grads = [p.grad for p in module_or_bucket.parameters() if p.grad is not None]
    flat = th.cat([g.reshape(-1) for g in grads])
    td.all_reduce(flat, op=td.ReduceOp.SUM)
    flat.mul_(1.0 / world_size())
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def sync_params(module):
    """Broadcast rank 0's parameters and buffers (``sync_params``), one flat message."""
    if world_size() == 1:
        return
    ts = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    flat = th.cat([t.reshape(-1).float() for t in ts])
    td.broadcast(flat, src=0)
    o = 0
    for t in ts:
        t.copy_(flat[o:o + t.numel()].view_as(t).to(t.dtype))
        o += t.numel()


def all_reduce_max_scalar(x: float, device) -> float:
    if world_size() == 1:
        return x
    t = th.tensor([x], dtype=th.float64, device=device)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())
