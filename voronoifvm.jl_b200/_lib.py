"""ctypes binding of libvfvmb200.so -- the same C ABI (include/vfvm_b200.h) a Julia `ccall` shim binds.

There is no CPU fallback: if the shared library is missing, or there is no CUDA device, the functions here raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvfvmb200.so")

OK = 0
ERR_ARG, ERR_STATE, ERR_CUDA, ERR_NAN, ERR_LINSOLVE, ERR_UNREGISTERED, ERR_COMM, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6, -7, -8
HOST, DEVICE = 0, 1
VEC_SOLUTION, VEC_OLDSOL, VEC_RESIDUAL, VEC_UPDATE = 0, 1, 2, 3
KRYLOV_BICGSTAB, KRYLOV_CG, KRYLOV_GMRES = 0, 1, 2
PRECON_NONE, PRECON_JACOBI, PRECON_BLOCKJACOBI, PRECON_ILU0, PRECON_ILU0_MC, PRECON_AMG = 0, 1, 2, 3, 4, 5
TIME_ASSEMBLE, TIME_LINSOLVE_SETUP, TIME_LINSOLVE_SOLVE, TIME_EDGE_KERNEL = 0, 1, 2, 3
NUM_TIMES = 8


class BCEntry(C.Structure):
    _fields_ = [("kind", C.c_int32), ("species", C.c_int32), ("region", C.c_int32), ("has_ramp", C.c_int32), ("value", C.c_double),
                ("factor", C.c_double), ("t0", C.c_double), ("t1", C.c_double), ("v0", C.c_double), ("v1", C.c_double)]


_p = C.POINTER
_H = C.c_void_p
_D, _I32, _I64, _U8 = _p(C.c_double), _p(C.c_int32), _p(C.c_int64), _p(C.c_uint8)

# name -> argtypes (every entry point of include/vfvm_b200.h; all return int unless noted)
SIGNATURES = {
    "vfvm_create": [C.c_int, _p(_H)],
    "vfvm_set_grid": [_H, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, _D, _I32, _I32, _I32, _I32],
    "vfvm_set_owned_nodes": [_H, C.c_int64],
    "vfvm_build_geometry": [_H],
    "vfvm_num_edges": [_H, _I64],
    "vfvm_get_edgenodes": [_H, _I32],
    "vfvm_get_celledges": [_H, _I32],
    "vfvm_num_factors": [_H, _I64, _I64],
    "vfvm_get_nodefactors": [_H, _I64, _I32, _D],
    "vfvm_get_edgefactors": [_H, _I64, _I32, _D],
    "vfvm_get_bfacefactors": [_H, _D],
    "vfvm_set_system": [_H, C.c_int, _U8],
    "vfvm_set_boundary_species": [_H, C.c_int, _U8],
    "vfvm_set_physics": [_H, C.c_int, C.c_int, _D, C.c_int],
    "vfvm_set_nodal_source": [_H, _D],
    "vfvm_set_legacy_bc": [_H, C.c_int, _D, _D],
    "vfvm_set_bc_entries": [_H, C.c_int, _p(BCEntry)],
    "vfvm_build_pattern": [_H],
    "vfvm_pattern_size": [_H, _I64, _I64],
    "vfvm_get_pattern_csr": [_H, _I64, _I64],
    "vfvm_get_pattern_csc": [_H, _I64, _I64],
    "vfvm_set_vector": [_H, C.c_int, C.c_void_p, C.c_int],
    "vfvm_get_vector": [_H, C.c_int, C.c_void_p, C.c_int],
    "vfvm_copy_vector": [_H, C.c_int, C.c_int],
    "vfvm_init_dirichlet": [_H, C.c_double, C.c_double],
    "vfvm_assemble": [_H, C.c_double, C.c_double, C.c_double],
    "vfvm_mass_matrix": [_H, C.c_void_p],
    "vfvm_edgeflux": [_H, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "vfvm_amg_set_options": [_H, C.c_void_p, C.c_int],
    "vfvm_integrate": [_H, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "vfvm_integrate_boundary": [_H, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "vfvm_edgeintegrate": [_H, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "vfvm_peer_export": [_H, C.c_char_p],
    "vfvm_peer_connect": [_H, C.c_char_p],
    "vfvm_peer_active": [_H],
    "vfvm_assemble_async": [_H, C.c_double, C.c_double, C.c_double],
    "vfvm_sync": [_H],
    "vfvm_eval_res_jac": [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double],
    "vfvm_get_nzval_csr": [_H, C.c_void_p, C.c_int],
    "vfvm_get_nzval_csc": [_H, C.c_void_p, C.c_int],
    "vfvm_get_rows_csr": [_H, C.c_int64, C.c_int64, _I64, _I64, _I64, _D],
    "vfvm_linsolve_setup": [_H, C.c_int, C.c_int, C.c_int],
    "vfvm_linsolve": [_H, C.c_double, C.c_double, C.c_int, C.c_int, _p(C.c_int), _D],
    "vfvm_linsolve_status": [_H, _p(C.c_int), _D],
    "vfvm_spmv": [_H, C.c_void_p, C.c_void_p, C.c_int],
    "vfvm_newton_update": [_H, C.c_double, _D, _D],
    "vfvm_vector_norms": [_H, C.c_int, _D, _D],
    "vfvm_vector_diffnorm": [_H, C.c_int, C.c_int, _D],
    "vfvm_comm_unique_id": [C.c_char_p],
    "vfvm_comm_init": [_H, C.c_int, C.c_int, C.c_char_p],
    "vfvm_set_halo": [_H, C.c_int, _I32, _I64, _I32, _I64],
    "vfvm_halo_exchange": [_H, C.c_int],
    "vfvm_timings": [_H, _D],
    "vfvm_launch_count": [_H, _I64],
    "vfvm_stream": [_H, _p(C.c_void_p)],
    "vfvm_device_bytes": [_H, _I64],
    "vfvm_plane_counts": [_H, _p(C.c_int), _p(C.c_int)],
    "vfvm_block_counts": [_H, _I64, _I64],
    "vfvm_probe_inplace_linsolve": [_H, C.c_int, C.c_int, C.c_int, _D, _D, _D],
    "vfvm_probe_bernoulli": [_H, C.c_int, _D, _D, _D, _D],
}
OTHER = {"vfvm_destroy": ([_H], None), "vfvm_last_error": ([_H], C.c_char_p), "vfvm_abi_version": ([], C.c_int)}

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with voronoifvm.jl_b200/csrc/build.sh (__graft_entry__.build()). "
                               "The B200 path has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        for name, (args, res) in OTHER.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _LIB = L
    return _LIB


class VfvmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libvfvmb200 error {code}: {msg}")
        self.code = code


def check(h, rc):
    if rc != OK:
        msg = lib().vfvm_last_error(h)
        raise VfvmError(rc, msg.decode() if msg else "")
    return rc


def dptr(a: np.ndarray):
    assert a.dtype == np.float64
    return a.ctypes.data_as(_D)


def i32ptr(a: np.ndarray):
    assert a.dtype == np.int32
    return a.ctypes.data_as(_I32)


def i64ptr(a: np.ndarray):
    assert a.dtype == np.int64
    return a.ctypes.data_as(_I64)
