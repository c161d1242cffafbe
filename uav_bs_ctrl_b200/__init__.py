"""B200-native hetero-graph message-passing hot path of zhangxiaochen95/uav_bs_ctrl.

Public surface (mirrors the reference's module/graph API for this path):

* ``graph``     DGL-free ``heterograph`` / ``batch`` / ``merge`` / ``DGLGraph`` container with CSR-by-dst.
* ``builder``   vectorised dense-observation → graph builder.
* ``agents``    ``GATv2Conv``, ``GraphObservationEncoder``, ``TarMAC``, ``GnnAgent`` (MADRQN and DRQN) — same
                constructor / forward signatures and ``state_dict`` keys as the reference, CUDA kernels inside.
* ``learner``   ``MultiAgentQLearner`` / ``QLearner`` mirrors (act / cache / update).
* ``dist``      flat-bucket NCCL gradient all-reduce (``avg_grads`` / ``sync_params``).

The CUDA kernels live in ``csrc/`` behind the C ABI declared in ``include/ubs_gnn.h`` and are loaded from the
in-tree ``libubs_gnn.so``; there is no CPU fallback — calling a kernel without the library raises.
"""
from . import graph, function  # noqa: F401
from .graph import HeteroGraph, DGLGraph, heterograph, batch, merge  # noqa: F401

__version__ = "0.1.0"
