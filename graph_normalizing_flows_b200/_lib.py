"""ctypes binding of libgnf_b200.so (C ABI in include/gnf_b200.h).

There is no CPU fallback: if the shared library is missing, or a compute entry point is
reached without a CUDA device, this module raises.  PyTorch is used only for device
memory and streams; every tensor crosses the boundary as a raw device pointer.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgnf_b200.so")

GNF_OK = 0
ABI_VERSION = 5
GNF_EINVAL, GNF_ECUDA, GNF_EUNSUPPORTED, GNF_EWORKSPACE = -1, -2, -3, -4
AGG = {"sum": 0, "mean": 1}
BLOCK = {"concat": 0, "agg_then": 1, "dm_attn": 2}
ATTN_CONCAT, ATTN_RESIDUAL, ATTN_KQ_DIV, ATTN_LAYER_NORM = 1, 2, 4, 8
ACT = {"leaky_relu": 0, "relu": 1}
MATH = {"fp32": 0, "tc3x": 1, "bf16": 2, "tc3x_bf16": 3, "tc2x": 4}


class FlowDesc(C.Structure):
    _fields_ = [
        ("num_timesteps", C.c_int32), ("node_embedding_dim", C.c_int32),
        ("latent_dim", C.c_int32), ("num_layers", C.c_int32),
        ("agg", C.c_int32), ("block", C.c_int32), ("act", C.c_int32),
        ("weight_sharing", C.c_int32), ("eps", C.c_float),
        ("attn_num_heads", C.c_int32), ("attn_kq_dim", C.c_int32), ("attn_v_dim", C.c_int32),
        ("attn_out_dim", C.c_int32), ("attn_flags", C.c_int32),
    ]


_p = C.c_void_p
_i32, _i64, _sz = C.c_int32, C.c_int64, C.c_size_t

# name -> (restype, argtypes); every symbol include/gnf_b200.h declares
SIGNATURES = {
    "gnf_abi_version": (C.c_int, []),
    "gnf_last_error": (C.c_char_p, []),
    "gnf_launch_count": (_i64, [C.c_int]),
    "gnf_debug_set_trace": (C.c_int, [_p]),
    "gnf_debug_kernel_timing": (C.c_int, [_i32]),
    "gnf_debug_kernel_time": (C.c_int, [C.POINTER(C.c_double), C.POINTER(_i64)]),
    "gnf_build_csr_workspace": (_sz, [_i64, _i64]),
    "gnf_build_csr": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "gnf_validate_indices": (C.c_int, [_p, _p, _i64, _i64, _p, _p]),
    "gnf_gather_rows": (C.c_int, [_p, _i32, _p, _i64, _p, _p]),
    "gnf_segment_sum": (C.c_int, [_p, _i32, _p, _p, _i64, _i32, _p, _p]),
    "gnf_gather_segment_sum": (C.c_int, [_p, _i32, _p, _p, _i64, _i32, _p, _p]),
    "gnf_flow_param_count": (_i64, [C.POINTER(FlowDesc)]),
    "gnf_flow_create": (C.c_int, [C.POINTER(_p), C.POINTER(FlowDesc)]),
    "gnf_flow_set_params": (C.c_int, [_p, _p, _p]),
    "gnf_flow_range_flag": (C.c_int, [_p, _p, _i32, _p]),
    "gnf_flow_destroy": (C.c_int, [_p]),
    "gnf_flow_supports": (C.c_int, [_p, _i32]),
    "gnf_flow_supports_backward": (C.c_int, [_p, _i32]),
    "gnf_grevnet_workspace": (_sz, [_p, _i64, _i32]),
    "gnf_grevnet_forward": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _i32, _p, _sz, _p]),
    "gnf_grevnet_inverse": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _i32, _p, _sz, _p]),
    "gnf_padded_half": (_i32, [_i32]),
    "gnf_coupling_step": (C.c_int, [_p, _i32, _i32, _p, _p, _i64, _i64, _p, _p, _p, _i32, _p, _sz, _p]),
    "gnf_coupling_half": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _i64, _i64, _p, _p, _p, _i32, _p, _sz, _p]),
    "gnf_split_halves": (C.c_int, [_p, _i64, _i32, _p, _p, _p]),
    "gnf_merge_halves": (C.c_int, [_p, _p, _i64, _i32, _p, _p]),
    "gnf_bn_moments_workspace": (_sz, [_i32]),
    "gnf_bn_moments": (C.c_int, [_p, _i64, _i32, _p, _p, _sz, _p]),
    "gnf_affine_rows": (C.c_int, [_p, _i64, _i32, _p, _p, _p]),
    "gnf_bn_finalize": (C.c_int, [_p, _i32, _p, _p, C.c_double, C.c_double, _p, _p, _p, _p, _p, C.c_float, _p]),
    "gnf_grevnet_bn_workspace": (_sz, [_p]),
    "gnf_grevnet_forward_bn": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _p, _p, C.c_double, C.c_float, _p, _p, _p,
                                         _i32, _p, _sz, _p, _sz, _p]),
    "gnf_grevnet_inverse_bn": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _p, _p, C.c_double, _p, _i32, _p, _sz, _p,
                                         _sz, _p]),
    "gnf_bn_backward_coef": (C.c_int, [_p, _p, _i32, _p, _p, C.c_double, C.c_double, _p, _p, _p, _p, _p]),
    "gnf_bn_backward_sums": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _p, _p, _sz, _p]),
    "gnf_bn_backward_apply": (C.c_int, [_p, _p, _i64, _i32, _p, _p]),
    "gnf_coupling_half_backward": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p, _i64, _i64, _p, _p, _p, _p, C.c_double, _p,
                                             _i32, _p, _sz, _p]),
    "gnf_gnn_forward": (C.c_int, [_p, _i32, _i32, _i32, _p, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "gnf_grevnet_backward_workspace": (_sz, [_p, _i64, _i32]),
    "gnf_grevnet_backward": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, C.c_double, _p, _p, _i32, _p, _sz, _p]),
    "gnf_debug_bwd_layout": (C.c_int, [_p, _i64, _p]),
    "gnf_debug_dw_gemm": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _i32, _p, _p, _sz, _p]),
    "gnf_debug_linear_tc_workspace": (_sz, [_i32, _i32]),
    "gnf_debug_linear_tc": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _i32, _p, _p, _sz, _p]),
    "gnf_pred_adj": (C.c_int, [_p, _i32, _p, _p, _i64, C.c_float, C.c_float, _p, _p]),
    "gnf_log_prob_workspace": (_sz, [_i64, _i32]),
    "gnf_log_prob": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _sz, _p]),
    "gnf_peer_create": (C.c_int, [C.POINTER(_p), _i32, _i32, _p]),
    "gnf_peer_connect": (C.c_int, [_p, _p]),
    "gnf_peer_destroy": (C.c_int, [_p]),
    "gnf_peer_allreduce4": (C.c_int, [_p, _p, _p]),
    "gnf_log_prob_allreduce": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _sz, _p, _p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (or `make -C "
                "graph_normalizing_flows_b200/csrc`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.gnf_abi_version() != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH} has ABI version {lib.gnf_abi_version()}, this package binds version "
                               f"{ABI_VERSION}: rebuild it (make -C graph_normalizing_flows_b200/csrc)")
        _lib = lib
    return _lib


def last_error() -> str:
    return load().gnf_last_error().decode()


def check(rc: int, what: str = ""):
    if rc == GNF_OK:
        return
    msg = f"{what}: {last_error()}" if what else last_error()
    if rc in (GNF_EINVAL, GNF_EUNSUPPORTED):
        raise ValueError(msg)
    raise RuntimeError(msg)


def require_cuda(t: torch.Tensor, name: str, dtype=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (no CPU fallback); got {t.device}")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def workspace(nbytes: int, device) -> torch.Tensor:
    # torch's caching allocator returns >=512-byte aligned blocks.  GNF_POISON_WORKSPACE=1 (set by the GPU tests) fills
    # the block with 0xFF bytes -- NaN as fp32/fp64, -1 as int32 -- so a kernel that reads a workspace word nobody wrote
    # fails its test every time instead of whenever the caching allocator hands back a dirty block
    if os.environ.get("GNF_POISON_WORKSPACE") == "1":
        return torch.full((max(int(nbytes), 256),), 255, dtype=torch.uint8, device=device)
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
