"""ctypes binding of the C ABI in ``include/ubs_gnn.h`` (in-tree ``libubs_gnn.so``, sm_100a).

There is no fallback: if the library is missing or a call fails, a ``RuntimeError`` is raised.  PyTorch is used
only for device memory and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch as th

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libubs_gnn.so")
_lib = None

_F, _I, _U = C.c_void_p, C.c_void_p, C.c_void_p      # device pointers travel as void*
_i64, _int, _flt, _ptr = C.c_int64, C.c_int, C.c_float, C.c_void_p

_PROTOS = {
    "ubs_version": (C.c_int, []),
    "ubs_last_error": (C.c_char_p, []),
    "ubs_launch_count": (_i64, []),
    "ubs_reset_launch_count": (None, []),
    "ubs_add_launch_count": (None, [_i64]),
    "ubs_gatv2_fwd": (C.c_int, [_F, _F, _I, _I, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _i64, _i64,
                                _int, _int, _int, _int, _flt, _int, _ptr]),
    "ubs_gatv2_bwd_workspace": (_i64, [_i64, _int, _int, _int, _int]),
    "ubs_gatv2_bwd": (C.c_int, [_F, _F, _I, _I, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F,
                                _i64, _i64, _i64, _int, _int, _int, _int, _flt, _int, _ptr]),
    "ubs_block_attn_fwd": (C.c_int, [_F, _i64, _F, _i64, _F, _i64, _U, _F, _F, _i64, _int, _int, _int, _flt, _ptr]),
    "ubs_block_attn_bwd": (C.c_int, [_F, _i64, _F, _i64, _F, _i64, _U, _F, _F, _F, _i64, _F, _i64, _F, _i64, _F,
                                     _i64, _int, _int, _int, _flt, _ptr]),
    "ubs_block_mean_fwd": (C.c_int, [_F, _i64, _U, _F, _i64, _i64, _int, _int, _ptr]),
    "ubs_block_mean_bwd": (C.c_int, [_F, _i64, _U, _F, _i64, _i64, _int, _int, _ptr]),
    "ubs_block_bitmax_fwd": (C.c_int, [_F, _i64, _F, _U, _I, _F, _i64, _ptr, _i64, _int, _int, _flt, _ptr]),
    "ubs_block_bitmax_bwd": (C.c_int, [_F, _i64, _F, _U, _I, _ptr, _F, _i64, _F, _i64, _i64, _int, _int, _flt, _ptr]),
    "ubs_gru_gates_fwd": (C.c_int, [_F, _F, _F, _F, _i64, _int, _ptr]),
    "ubs_gru_gates_bwd": (C.c_int, [_F, _F, _F, _F, _F, _F, _F, _i64, _int, _ptr]),
    "ubs_gatv2_seg_fwd": (C.c_int, [_F] * 14 + [_i64] * 8 + [_int] * 4 + [_flt, _int, _ptr]),
    "ubs_gatv2_seg_bwd": (C.c_int, [_F] * 19 + [_i64] * 9 + [_int] * 4 + [_flt, _int, _ptr]),
    "ubs_gatv2_seg_fwd_scores": (C.c_int, [_F] * 15 + [_i64] * 9 + [_int] * 4 + [_flt, _int, _ptr]),
    "ubs_gatv2_seg_bwd_scores": (C.c_int, [_F] * 16 + [_i64] + [_F] * 4 + [_i64] * 9 + [_int] * 4 + [_flt, _int, _ptr]),
    "ubs_gat_aggr_fwd": (C.c_int, [_F] * 9 + [_i64, _i64, _int, _int, _flt, _int, _ptr]),
    "ubs_gat_aggr_bwd_workspace": (_i64, [_i64, _int, _int]),
    "ubs_gat_aggr_bwd": (C.c_int, [_F] * 15 + [_i64, _i64, _int, _int, _flt, _int, _ptr]),
    "ubs_tf32x3_gemm": (C.c_int, [_F, _i64, _F, _i64, _F, _F, _i64, _i64, _int, _int, _int, _ptr]),
    "ubs_tf32x3_gemm_tn_workspace": (_i64, [_i64, _int, _int]),
    "ubs_tf32x3_gemm_tn": (C.c_int, [_F, _i64, _F, _i64, _F, _i64, _F, _i64, _int, _int, _ptr]),
    "ubs_colsum_workspace": (_i64, [_int]),
    "ubs_colsum": (C.c_int, [_F, _i64, _i64, _int, _F, _F, _ptr]),
    "ubs_relu_bwd_colsum": (C.c_int, [_F, _i64, _F, _i64, _F, _i64, _i64, _int, _F, _F, _ptr]),
    "ubs_agent_seq2_smem_bytes": (_i64, [_int] * 6),
    "ubs_agent_seq2_fwd": (C.c_int, [_int] * 5 + [_F] * 13 + [_i64, _i64, _i64, _int, _ptr]),
    "ubs_agent_seq2_bwd": (C.c_int, [_int] * 5 + [_F] * 12 + [_i64, _i64, _int, _ptr]),
    "ubs_agent_pack_size": (_i64, [_int] * 7),
    "ubs_agent_act_uses_tma": (C.c_int, [_int] * 7),
    "ubs_agent_pack": (C.c_int, [_int] * 7 + [_F] * 15 + [_ptr]),
    "ubs_agent_seq_fwd": (C.c_int, [_int] * 7 + [_F] * 11 + [_i64, _int, _ptr]),
    "ubs_agent_act_fwd": (C.c_int, [_int] * 7 + [_F] * 14 + [_i64, _int, _ptr]),
    "ubs_agent_seq_bwd": (C.c_int, [_int] * 7 + [_F] * 15 + [_i64, _int, _ptr]),
    "ubs_gatv2_rel_pack_size": (_i64, [_int, _int]),
    "ubs_gatv2_rel_pack": (C.c_int, [_F] * 7 + [_int] * 4 + [_flt, _int, _F, _ptr]),
    "ubs_agent_act_rel_supported": (C.c_int, [_int] * 12),
    "ubs_agent_act_rel_fwd": (C.c_int, [_int] * 6 + [_F, _F, _F, _I, _int, _int, _F, _I, _int, _int, _F, _int, _int, _int]
                              + [_F] * 8 + [_i64, _ptr]),
    # include/ubs_env.h (config / state / packet structs travel by host pointer)
    "ubs_env_scratch_words": (_i64, [_ptr, _i64]),
    "ubs_env_phase_clocks": (C.c_int, [_ptr]),
    "ubs_env_sample_layouts": (C.c_int, [_ptr, _ptr, C.c_uint64, C.c_uint32, _i64, _ptr]),
    "ubs_env_reset": (C.c_int, [_ptr, _ptr, _ptr, _I, _i64, _ptr]),
    "ubs_env_step": (C.c_int, [_ptr, _ptr, _I, _ptr, _I, _i64, _ptr]),
}


def exported_symbols():
    """Names every build of the library must export (checked by the CPU test-suite against the header)."""
    return list(_PROTOS)


def build(verbose: bool = False) -> str:
    """Compiles ``csrc/*.cu`` for sm_100a into the in-tree shared library (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"building libubs_gnn.so failed:\n{res.stdout[-4000:]}\n{res.stderr[-4000:]}")
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           f"(there is no CPU / eager fallback for the CUDA hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.ubs_version() != 100:
        raise RuntimeError(f"libubs_gnn.so version mismatch: {lib.ubs_version()}")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {load().ubs_last_error().decode()}")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return th.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().ubs_launch_count())


def reset_launch_count():
    load().ubs_reset_launch_count()


def add_launches(n: int):
    load().ubs_add_launch_count(int(n))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("uav_bs_ctrl_b200 kernels run on CUDA tensors only (no CPU fallback); "
                               "use oracle/ for CPU checks")
