"""ctypes binding of the C ABI in ``include/edgpu.h`` (libedgpu.so).

This is the same surface the Fortran ``iso_c_binding`` shim binds.  There
is no CPU fallback anywhere: if the shared library is missing or no B200 is visible, calls
raise :class:`EdgpuError`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

MAXORB, MAXBATH, UID_BYTES = 5, 32, 128
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libedgpu.so")

# every symbol include/edgpu.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "edgpu_init", "edgpu_finalize", "edgpu_comm_unique_id", "edgpu_comm_init", "edgpu_comm_rank",
    "edgpu_comm_size", "edgpu_sector_open_normal", "edgpu_sector_close", "edgpu_sector_vecdim",
    "edgpu_sector_dim", "edgpu_sector_dims", "edgpu_sector_get_map", "edgpu_sector_hop_count",
    "edgpu_sector_get_hops", "edgpu_hxv_d", "edgpu_status", "edgpu_hxv_dev", "edgpu_vec_padded_len",
    "edgpu_vec_upload", "edgpu_vec_download", "edgpu_lanczos_gs", "edgpu_lanczos_tridiag",
    "edgpu_state_store", "edgpu_state_free", "edgpu_apply_op", "edgpu_state_observables",
    "edgpu_last_error", "edgpu_launch_count", "edgpu_last_hxv_stage_ms", "edgpu_set_kernel_variant",
    "edgpu_stream", "edgpu_profile_begin", "edgpu_profile_end", "edgpu_csr_open_d",
    "edgpu_csr_open_z", "edgpu_hxv_z", "edgpu_eigh", "edgpu_eigh_state_store",
    "edgpu_sector_open_nonsu2", "edgpu_csr_nnz", "edgpu_csr_get", "edgpu_lanczos_last_info",
    "edgpu_release_cache", "edgpu_sector_open_superc", "edgpu_apply_ops_packed", "edgpu_seed_norm2",
    "edgpu_set_coulomb_sundry", "edgpu_set_phonons", "edgpu_set_hbath_packed", "edgpu_state_twin", "edgpu_state_download", "edgpu_sector_open_normal_orbs", "edgpu_apply_ops_normal",
    "edgpu_sector_comm_info", "edgpu_set_sparse_h", "edgpu_host_register", "edgpu_host_unregister", "edgpu_halo_plan",
]


class EdgpuError(RuntimeError):
    pass


class NormalParams(C.Structure):
    """``edgpu_normal_params`` (include/edgpu.h)."""

    _fields_ = [
        ("Ns", C.c_int32), ("Norb", C.c_int32), ("Nbath", C.c_int32), ("bath_type", C.c_int32),
        ("hfmode", C.c_int32), ("Nfoo", C.c_int32), ("pad0", C.c_int32), ("pad1", C.c_int32),
        ("xmu", C.c_double),
        ("eloc", C.c_double * (2 * MAXORB * MAXORB)),
        ("spin_field_z", C.c_double * MAXORB),
        ("exc_field", C.c_double * 4),
        ("Uloc", C.c_double * MAXORB),
        ("Ust", C.c_double * (MAXORB * MAXORB)),
        ("Jh", C.c_double * (MAXORB * MAXORB)),
        ("Jx", C.c_double * (MAXORB * MAXORB)),
        ("Jp", C.c_double * (MAXORB * MAXORB)),
        ("diag_hybr", C.c_double * (2 * MAXORB * MAXBATH)),
        ("bath_diag", C.c_double * (2 * MAXORB * MAXBATH)),
        ("hbath", C.c_double * (2 * MAXORB * MAXORB * MAXBATH)),
        ("stride", C.c_int32 * (MAXORB * MAXBATH)),
    ]


class SundryTerm(C.Structure):
    """``edgpu_sundry_term`` = coulomb_matrix_element (ED_VARS_GLOBAL.f90:24-31): every operator
    as (orbital 1-based, spin 1 up / 2 dw)."""

    _fields_ = [("cd_i", C.c_int32 * 2), ("cd_j", C.c_int32 * 2), ("c_k", C.c_int32 * 2),
                ("c_l", C.c_int32 * 2), ("U", C.c_double)]


class Nonsu2Params(C.Structure):
    """``edgpu_nonsu2_params`` (include/edgpu.h)."""

    _fields_ = [
        ("Ns", C.c_int32), ("Norb", C.c_int32), ("Nbath", C.c_int32), ("bath_type", C.c_int32),
        ("hfmode", C.c_int32), ("Nfoo", C.c_int32), ("pad0", C.c_int32), ("pad1", C.c_int32),
        ("xmu", C.c_double),
        ("hloc", C.c_double * (2 * 2 * MAXORB * MAXORB * 2)),
        ("spin_field", C.c_double * (MAXORB * 3)),
        ("Uloc", C.c_double * MAXORB),
        ("Ust", C.c_double * (MAXORB * MAXORB)),
        ("Jh", C.c_double * (MAXORB * MAXORB)),
        ("Jx", C.c_double * (MAXORB * MAXORB)),
        ("Jp", C.c_double * (MAXORB * MAXORB)),
        ("bath_e", C.c_double * (2 * MAXORB * MAXBATH)),
        ("bath_v", C.c_double * (2 * MAXORB * MAXBATH)),
        ("bath_u", C.c_double * (2 * MAXORB * MAXBATH)),
        ("stride", C.c_int32 * (MAXORB * MAXBATH)),
    ]


class SupercParams(C.Structure):
    """``edgpu_superc_params`` (include/edgpu.h)."""

    _fields_ = [
        ("Ns", C.c_int32), ("Norb", C.c_int32), ("Nbath", C.c_int32), ("bath_type", C.c_int32),
        ("hfmode", C.c_int32), ("Nfoo", C.c_int32), ("pad0", C.c_int32), ("pad1", C.c_int32),
        ("xmu", C.c_double),
        ("hloc", C.c_double * (2 * MAXORB * MAXORB * 2)),
        ("hloc_anomalous", C.c_double * (MAXORB * MAXORB * 2)),
        ("pair_field", C.c_double * MAXORB),
        ("Uloc", C.c_double * MAXORB),
        ("Ust", C.c_double * (MAXORB * MAXORB)),
        ("Jh", C.c_double * (MAXORB * MAXORB)),
        ("Jx", C.c_double * (MAXORB * MAXORB)),
        ("Jp", C.c_double * (MAXORB * MAXORB)),
        ("bath_e", C.c_double * (2 * MAXORB * MAXBATH)),
        ("bath_d", C.c_double * (MAXORB * MAXBATH)),
        ("bath_v", C.c_double * (2 * MAXORB * MAXBATH)),
        ("stride", C.c_int32 * (MAXORB * MAXBATH)),
    ]


_lib = None


def load():
    """Load libedgpu.so and declare the prototypes; raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EdgpuError(
            f"{LIB_PATH} not found: build it with `python -m edipack_b200.build` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    i64 = C.c_int64
    L.edgpu_init.argtypes = [C.c_int]
    L.edgpu_comm_unique_id.argtypes = [C.c_void_p]
    L.edgpu_comm_init.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.edgpu_sector_open_normal.argtypes = [C.POINTER(NormalParams), C.c_int, C.c_int]
    L.edgpu_set_coulomb_sundry.argtypes = [C.c_int, C.c_void_p]
    L.edgpu_set_hbath_packed.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.edgpu_set_phonons.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_int]
    L.edgpu_sector_open_normal_orbs.argtypes = [C.POINTER(NormalParams), C.c_void_p, C.c_void_p]
    L.edgpu_sector_vecdim.restype = i64
    L.edgpu_sector_dim.restype = i64
    L.edgpu_sector_dims.argtypes = [C.POINTER(i64)] * 4
    L.edgpu_halo_plan.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.POINTER(C.c_int64)]
    L.edgpu_host_register.argtypes = [C.c_void_p, C.c_int64]
    L.edgpu_host_unregister.argtypes = [C.c_void_p]
    L.edgpu_sector_comm_info.argtypes = [C.POINTER(C.c_int), C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_int)]
    L.edgpu_sector_get_map.argtypes = [C.c_int, C.c_void_p]
    L.edgpu_sector_hop_count.restype = i64
    L.edgpu_sector_hop_count.argtypes = [C.c_int]
    L.edgpu_sector_get_hops.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.edgpu_hxv_d.restype = None
    L.edgpu_hxv_d.argtypes = [C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]
    L.edgpu_hxv_z.restype = None
    L.edgpu_hxv_z.argtypes = [C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]
    for f in (L.edgpu_csr_open_d, L.edgpu_csr_open_z):
        f.argtypes = [i64, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.edgpu_sector_open_nonsu2.argtypes = [C.POINTER(Nonsu2Params), C.c_int]
    L.edgpu_sector_open_superc.argtypes = [C.POINTER(SupercParams), C.c_int]
    L.edgpu_csr_nnz.restype = i64
    L.edgpu_csr_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.edgpu_hxv_dev.argtypes = [C.c_void_p, C.c_void_p]
    L.edgpu_vec_padded_len.restype = i64
    L.edgpu_vec_upload.argtypes = [C.c_void_p, C.c_void_p]
    L.edgpu_vec_download.argtypes = [C.c_void_p, C.c_void_p]
    L.edgpu_lanczos_gs.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_uint64, dp,
                                   C.c_void_p, C.POINTER(C.c_int)]
    L.edgpu_lanczos_tridiag.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_int), dp]
    L.edgpu_eigh.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_void_p,
                             C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.edgpu_eigh_state_store.argtypes = [C.c_int, C.c_int]
    L.edgpu_lanczos_last_info.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.edgpu_state_store.argtypes = [C.c_int]
    L.edgpu_state_free.argtypes = [C.c_int]
    L.edgpu_apply_op.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    L.edgpu_apply_ops_packed.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.edgpu_apply_ops_normal.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.edgpu_seed_norm2.argtypes = [dp]
    L.edgpu_state_twin.argtypes = [C.c_int, C.c_int]
    L.edgpu_state_download.argtypes = [C.c_int, C.c_void_p]
    L.edgpu_state_observables.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    L.edgpu_last_error.restype = C.c_char_p
    L.edgpu_launch_count.restype = i64
    L.edgpu_launch_count.argtypes = [C.c_int]
    L.edgpu_last_hxv_stage_ms.argtypes = [C.c_void_p]
    L.edgpu_set_kernel_variant.argtypes = [C.c_int]
    L.edgpu_stream.restype = C.c_void_p
    L.edgpu_profile_begin.argtypes = [C.c_int]
    L.edgpu_profile_end.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise EdgpuError(load().edgpu_last_error().decode())


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)
