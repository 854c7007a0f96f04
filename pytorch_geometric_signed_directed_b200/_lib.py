"""ctypes binding of libpgsd_b200.so (the C ABI declared in include/pgsd_b200.h).

There is deliberately NO fallback: if the CUDA library has not been built, or a layer is fed
CPU tensors, the call raises.  (The CPU restatement under oracle/ is test infrastructure and
is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libpgsd_b200.so")

PGSD_F32, PGSD_BF16 = 0, 1
DENSE_MAX_TERMS = 16

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class SpmmArgs(C.Structure):
    _fields_ = [
        ("n_rows", _i64), ("feat", _i32), ("n_ops", _i32), ("dtype", _i32), ("mean", _i32),
        ("row_ptr", _vp), ("col", _vp),
        ("val", _vp * 2), ("diag", _vp * 2), ("diag_const", _f32 * 2),
        ("x", _vp * 2), ("ldx", _i64 * 2),
        ("alpha", _f32), ("beta", _f32),
        ("z", _vp * 2), ("ldz", _i64 * 2),
        ("y", _vp * 2), ("ldy", _i64 * 2),
        ("bias", _vp), ("variant", _i32), ("diag_row_offset", _i32),
        ("op_scale", _f32 * 2),
        ("long_rows", _vp), ("long_chunk_ptr", _vp), ("n_long_rows", _i32),
        ("long_row_threshold", _i32), ("long_chunk", _i32), ("grid_reserve", _i32),
    ]


class DenseArgs(C.Structure):
    _fields_ = [
        ("n_rows", _i64), ("n_out", _i32), ("n_terms", _i32), ("dtype", _i32), ("combine", _i32),
        ("x", _vp * DENSE_MAX_TERMS), ("ldx", _i64 * DENSE_MAX_TERMS),
        ("k", _i32 * DENSE_MAX_TERMS), ("group", _i32 * DENSE_MAX_TERMS),
        ("w", _vp * DENSE_MAX_TERMS), ("ldw_k", _i64 * DENSE_MAX_TERMS),
        ("ldw_n", _i64 * DENSE_MAX_TERMS),
        ("bias", _vp), ("y", _vp * 2), ("ldy", _i64 * 2),
        ("relu_mode", _i32), ("variant", _i32),
    ]


class MagnetFusedArgs(C.Structure):
    _fields_ = [
        ("n_rows", _i64), ("feat_in", _i32), ("feat_out", _i32),
        ("row_ptr", _vp), ("col", _vp),
        ("val", _vp * 2), ("diag", _vp * 2), ("diag_const", _f32 * 2),
        ("x", _vp * 2), ("ldx", _i64 * 2),
        ("w", _vp * 2), ("ldw_k", _i64 * 2), ("ldw_n", _i64 * 2),
        ("bias", _vp), ("y", _vp * 2), ("ldy", _i64 * 2),
        ("relu_mode", _i32), ("variant", _i32),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [
        ("n_rows", _i64), ("n_types", _i32), ("act", _i32), ("slope", _f32), ("reserved", _i32),
        ("row_ptr", _vp * 2), ("col", _vp * 2), ("s_src", _vp * 2), ("s_dst", _vp * 2),
        ("dalpha", _vp * 2), ("row_coef", _vp * 2), ("g_s_src", _vp * 2), ("g_s_dst", _vp * 2),
        ("type_sum", _vp * 2),
    ]


MAX_RANKS, MAX_SLICES = 16, 16


class PushArgs(C.Structure):
    _fields_ = [
        ("world", _i32), ("rank", _i32), ("n_tensors", _i32), ("row_bytes", _i32), ("n_rows", _i64),
        ("src", _vp * 2), ("ld_src_bytes", _i64 * 2),
        ("dst", (_vp * MAX_RANKS) * 2), ("ld_dst_bytes", _i64 * 2), ("mc_dst", _vp * 2),
        ("n_slices", _i32), ("n_ctas", _i32), ("slice_row", _i64 * (MAX_SLICES + 1)),
        ("flag", _vp * MAX_RANKS), ("counters", _vp), ("seq", C.c_uint32), ("include_self", _i32),
        ("engine", _i32), ("chunk_bytes", _i32), ("stages", _i32), ("reserved", _i32),
        ("started", _vp),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("n_rows", _i64), ("feat", _i32), ("n_types", _i32), ("act", _i32), ("slope", _f32),
        ("row_ptr", _vp * 2), ("col", _vp * 2), ("s_src", _vp * 2), ("s_dst", _vp * 2),
        ("xd", _vp * 2), ("ldxd", _i64 * 2), ("y", _vp), ("ldy", _i64), ("alpha_out", _vp * 2),
    ]


_PROTOTYPES = {
    "pgsd_abi_version": (C.c_int, []),
    "pgsd_last_error": (C.c_char_p, []),
    "pgsd_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "pgsd_sizeof_args": (C.c_int, [C.POINTER(C.c_size_t)] * 2),
    "pgsd_plan_workspace_bytes": (C.c_int, [_i64, _i64, C.POINTER(C.c_size_t)]),
    "pgsd_build_csr": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp,
                                 C.c_size_t, _vp]),
    "pgsd_build_csr_rw_norm": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _f32, C.c_int, _vp, _vp, _vp,
                                         _vp, C.POINTER(_i64), _vp, C.c_size_t, _vp]),
    "pgsd_build_csr_sym_norm": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _f32, C.c_int, _vp, _vp, _vp,
                                          _vp, C.POINTER(_i64), _vp, C.c_size_t, _vp]),
    "pgsd_build_magnetic_laplacian": (C.c_int, [_vp, _vp, _vp, _i64, _i64, C.c_double, C.c_int,
                                                _f32, C.c_int, _vp, _vp, _vp, _vp, _vp,
                                                C.POINTER(_i64), _vp, C.c_size_t, _vp]),
    "pgsd_build_magnetic_laplacian_theta": (C.c_int, [_vp, _vp, _vp, _i64, _i64, C.c_double, C.c_int,
                                                      _f32, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp,
                                                      C.POINTER(_i64), _vp, C.c_size_t, _vp]),
    "pgsd_build_magnetic_rows_begin": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, C.c_int, _vp, _vp, _vp,
                                                 _vp, _vp, C.POINTER(_i64), _vp, C.c_size_t, _vp]),
    "pgsd_build_magnetic_rows_finish": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, C.c_double,
                                                  C.c_int, _f32, _vp, _vp, _vp, _vp]),
    "pgsd_signed_triangle_counts": (C.c_int, [C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _i64, _i64, _vp, _vp]),
    "pgsd_gat_aggregate": (C.c_int, [_vp, _vp, _vp, _vp, _f32, _vp, _i64, _i32, _i64, _vp, _vp, _i64, _f32, _vp, _i64, _vp]),
    "pgsd_magnetic_q_grad": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i64, _vp, _i64,
                                       _vp, _i64, _vp, _i64, C.c_double, _vp, _vp]),
    "pgsd_spmm_csr": (C.c_int, [C.POINTER(SpmmArgs), _vp]),
    "pgsd_last_spmm_kernel": (C.c_char_p, []),
    "pgsd_dense_transform": (C.c_int, [C.POINTER(DenseArgs), _vp]),
    "pgsd_magnet_fused_supported": (C.c_int, [_i32, _i32, _i32]),
    "pgsd_sizeof_magnet_fused_args": (C.c_size_t, []),
    "pgsd_magnet_layer_fused": (C.c_int, [C.POINTER(MagnetFusedArgs), _vp]),
    "pgsd_edge_softmax": (C.c_int, [C.POINTER(AttnArgs), _vp]),
    "pgsd_sizeof_attn_bwd_args": (C.c_size_t, []),
    "pgsd_edge_softmax_backward": (C.c_int, [C.POINTER(AttnBwdArgs), _vp]),
    "pgsd_sddmm_rows": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp]),
    "pgsd_xtg_accumulate": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _vp, _vp]),
    "pgsd_coalesce_workspace_bytes": (C.c_int, [_i64, C.POINTER(C.c_size_t)]),
    "pgsd_coo_coalesce": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _vp, C.POINTER(_i64), _vp, C.c_size_t, _vp]),
    "pgsd_gram_expand": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "pgsd_ppr_stationary": (C.c_int, [_vp, _vp, _vp, _i64, C.c_double, _i32, _vp, _vp, _vp]),
    "pgsd_gather_rows": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _i64, _vp]),
    "pgsd_sizeof_push_args": (C.c_size_t, []),
    "pgsd_shard_push": (C.c_int, [C.POINTER(PushArgs), _vp]),
    "pgsd_fingerprint": (C.c_int, [_vp, _i64, _vp, _vp]),
    "pgsd_peer_copy": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "pgsd_signal_flag": (C.c_int, [_vp, C.c_uint32, _vp]),
    "pgsd_wait_flags": (C.c_int, [_vp, C.POINTER(_i32), _i32, C.c_uint32, C.c_uint64, _vp, _vp]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib: Optional[C.CDLL] = None


class PgsdError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PgsdError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python -m pytorch_geometric_signed_directed_b200.build); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.pgsd_abi_version() != 1:
            raise PgsdError("libpgsd_b200.so ABI version mismatch; rebuild")
        a, b = C.c_size_t(0), C.c_size_t(0)
        lib.pgsd_sizeof_args(C.byref(a), C.byref(b))
        if (a.value, b.value) != (C.sizeof(SpmmArgs), C.sizeof(DenseArgs)):
            raise PgsdError("ctypes struct mirror out of sync with include/pgsd_b200.h; rebuild")
        if lib.pgsd_sizeof_magnet_fused_args() != C.sizeof(MagnetFusedArgs):
            raise PgsdError("ctypes mirror of pgsd_magnet_fused_args out of sync with include/pgsd_b200.h; rebuild")
        if lib.pgsd_sizeof_attn_bwd_args() != C.sizeof(AttnBwdArgs):
            raise PgsdError("ctypes mirror of pgsd_attn_bwd_args out of sync with include/pgsd_b200.h; rebuild")
        if lib.pgsd_sizeof_push_args() != C.sizeof(PushArgs):
            raise PgsdError("ctypes mirror of pgsd_push_args out of sync with include/pgsd_b200.h; rebuild")
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().pgsd_last_error().decode(errors="replace")
        raise PgsdError(f"{what or 'pgsd'} failed (code {rc}): {msg}")
