"""ctypes binding of libglass_b200.so (C ABI declared in include/glass_b200.h).

The library is built in-tree by glass_b200/build.py (nvcc, sm_100a).  There is no CPU or
torch.sparse fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libglass_b200.so")

# enums of include/glass_b200.h
AGGR = {"mean": 0, "sum": 1, "gcn": 2}
ACT_NONE, ACT_RELU, ACT_ELU = 0, 1, 2
POOL = {"sum": 0, "mean": 1, "max": 2, "size": 3}
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2

_vp, _i64, _i32, _f32, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); kept in the order of the header so tests can diff it against the header
PROTOTYPES = {
    "glass_abi_version": (_i32, []),
    "glass_last_error": (C.c_char_p, []),
    "glass_sm_count": (_i32, []),
    "glass_csr_build_workspace_bytes": (_sz, [_i64, _i64]),
    "glass_csr_build": (_i32, [_vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                               C.POINTER(C.c_int64), _vp, _sz, _vp]),
    "glass_to_undirected_workspace_bytes": (_sz, [_i64]),
    "glass_to_undirected": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, C.POINTER(C.c_int64), C.POINTER(C.c_int), _vp, _sz, _vp]),
    "glass_spmm_stats_ld": (_i32, []),
    "glass_tune": (_i32, [C.c_char_p, _i32]),
    "glass_spmm_csr": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i32, _vp, _i32, C.POINTER(C.c_int), _vp]),
    "glass_spmm_plan_size": (_i32, [_vp, _i64, _i32, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int64), _vp]),
    "glass_spmm_plan_build": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "glass_spmm_csr_planned": (_i32, [_vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _i64,
                                      _vp, _vp, _vp, _i64, _vp, _vp, _i32, C.POINTER(C.c_int), _vp]),
    "glass_spmm_delta": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _i64,
                                _vp, _vp, _vp, _i64, _vp, _vp, _i32, C.POINTER(C.c_int), _vp]),
    "glass_pair_linear_mix_fwd": (_i32, [_vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _f32, _i32,
                                         _vp, _i64, _vp, _i64, _i32, _i32, _vp]),
    "glass_pair_norm_operand_supported": (_i32, [_i32, _i32, _i32]),
    "glass_pair_linear_mix_fwd_ex": (_i32, [_vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _f32, _i32,
                                            _vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "glass_pair_linear_mix_bwd_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "glass_pair_linear_mix_bwd_ex": (_i32, [_vp, _i64, _vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _f32,
                                            _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _sz,
                                            _i32, _vp, _vp, _i32, _i32, _vp]),
    "glass_pair_linear_mix_bwd": (_i32, [_vp, _i64, _vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _f32,
                                         _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _sz,
                                         _i32, _vp]),
    "glass_graphnorm_workspace_bytes": (_sz, [_i64, _i32]),
    "glass_dropout_bits_bytes": (_sz, [_i64, _i32]),
    "glass_graphnorm_launches": (_i32, [_i64, _i32]),
    "glass_graphnorm_fwd": (_i32, [_vp, _i64, _vp, _vp, _vp, _f32, _i32, _vp, _f32, _vp, _vp, _vp, _i64, _vp, _i64,
                                   _i32, _vp, _sz, _vp]),
    "glass_graphnorm_bwd": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _vp, _f32, _vp, _vp, _vp, _i64, _vp,
                                   _vp, _vp, _i64, _i32, _vp, _sz, _vp]),
    "glass_graphnorm_stats": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _f32, _vp, _f32, _vp, _vp, _vp, _i64, _i32, _vp]),
    "glass_graphnorm_apply": (_i32, [_vp, _i64, _vp, _i32, _vp, _f32, _vp, _vp, _i64, _i64, _i32, _vp]),
    "glass_graphnorm_bwd_from_sums": (_i32, [_vp, _i32, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp,
                                             _vp, _vp, _i64, _i32, _vp, _sz, _vp]),
    "glass_graphnorm_partials_ld": (_i32, []),
    "glass_graphnorm_partials": (_i32, [_vp, _i64, _i64, _i32, _vp, _i32, C.POINTER(C.c_int), _vp]),
    "glass_graphnorm_bwd_partials": (_i32, [_vp, _i64, _vp, _i64, _vp, _i32, _vp, _f32, _vp, _i64, _i32, _vp, _i32,
                                            C.POINTER(C.c_int), _vp]),
    "glass_graphnorm_bwd_finish": (_i32, [_vp, _i32, _i32, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _vp, _f32,
                                          _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _vp, _sz, _vp]),
    "glass_spmm_csr_acc": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _vp,
                                  _vp, _i64, _vp, _i32, _vp]),
    "glass_l2_gather_probe": (_i32, [_vp, _i64, _i64, _i32, _i64, _vp, _i64, _vp]),
    "glass_l2_gather_probe256": (_i32, [_vp, _i64, _i64, _i32, _i64, _vp, _i64, _i32, _vp]),
    "glass_l2_gather_probe24": (_i32, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, _i64, _i32, _vp]),
    "glass_embedding_fwd": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp]),
    "glass_embedding_bwd": (_i32, [_vp, _i64, _vp, _vp, _i64, _i64, _i32, _vp]),
    "glass_embedding_bwd_ordered": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _i32, _vp]),
    "glass_segment_pool_fwd": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _vp, _i32, _i64, _vp]),
    "glass_segment_pool_bwd_scratch_bytes": (_sz, [_i64, _i64]),
    "glass_segment_pool_bwd": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _i32, _i64, _vp, _sz, _vp]),
    "glass_norm_pool_fwd": (_i32, [_vp, _i64, _vp, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _vp, _i32, _i64, _vp]),
    "glass_norm_pool_bwd": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp,
                                   _vp, _i32, _i64, _vp, _sz, _vp]),
    "glass_segment_pool_batch_fwd": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _vp, _i32, _vp]),
    "glass_segment_pool_batch_bwd": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _i32, _vp]),
    "glass_adam_chunk": (_i32, []),
    "glass_adam_step": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _f32, _f32, _f32, _f32, _vp]),
    "glass_dp_flags_bytes": (_i32, []),
    "glass_dp_adam_step": (_i32, [_i32, _i32, _vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _i64, _i64, _i64, _i64,
                                  _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "glass_maxzoz": (_i32, [_vp, _i64, _vp, _vp, _i64, _vp]),
    "glass_label_mask": (_i32, [_vp, _vp, _i64, _vp]),
    "glass_pad2batch": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp]),
}



class NormOperand(C.Structure):
    """glass_norm_operand of include/glass_b200.h."""
    _fields_ = [("stats", C.c_void_p), ("bits", C.c_void_p), ("drop_p", C.c_float), ("act", C.c_int)]


_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m glass_b200.build` "
            "(glass_b200 has no CPU / torch.sparse fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.glass_abi_version() != 2:
        raise RuntimeError("libglass_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().glass_last_error().decode(errors="replace")
        if rc == -4:
            raise NotImplementedError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")
