"""ctypes binding of libfeast_cuda.so (the C ABI declared in include/feast_cuda.h).

This is the Python twin of the Julia `ccall` shim in julia/FEASTSolverB200.jl.
There is no CPU fallback: if the shared library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libfeast_cuda.so")

FEAST_OK = 0
FEAST_ERR_CUDA, FEAST_ERR_NCCL, FEAST_ERR_OOM, FEAST_ERR_STATE, FEAST_ERR_SINGULAR = 1000, 1001, 1002, 1003, 1004
FEAST_WARN_INNER_MAXIT = 2000
SOLVER_AUTO, SOLVER_DENSE_LU, SOLVER_KRYLOV, SOLVER_BANDED_LU = 0, 1, 2, 3
KRYLOV_AUTO, KRYLOV_COCG, KRYLOV_BICGSTAB, KRYLOV_GMRES = 0, 1, 2, 3
PROBLEM_STANDARD, PROBLEM_GENERALIZED, PROBLEM_POLYNOMIAL, PROBLEM_SAMPLED = 0, 1, 2, 3
PRECOND_NONE, PRECOND_AMG, PRECOND_AUTO = 0, 1, 2
SHARD_AUTO, SHARD_NODES, SHARD_COLUMNS = 0, 1, 2
MAX_SLOTS = 8
MAX_MOMENTS = 8


class c128(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


class FeastStats(C.Structure):
    _fields_ = [("nodes_local", C.c_int), ("inner_iters_total", C.c_int), ("inner_iters_max", C.c_int),
                ("info", C.c_int), ("inner_relres_max", C.c_double), ("t_factor_ms", C.c_double),
                ("t_solve_ms", C.c_double), ("t_reduce_ms", C.c_double), ("t_total_ms", C.c_double),
                ("t_spmm_ms", C.c_double), ("spmm_launches", C.c_int64), ("precond_levels", C.c_int), ("col_sharded", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class FeastError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfeast_cuda error {code}: {msg}")
        self.code = code


_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double

# name -> (restype, argtypes); must list EVERY symbol include/feast_cuda.h declares
SIGNATURES = {
    "feast_version": (_i, []),
    "feast_device_count": (_i, [C.POINTER(_i)]),
    "feast_ctx_create": (_i, [C.POINTER(_vp), _i]),
    "feast_ctx_destroy": (_i, [_vp]),
    "feast_last_error": (C.c_char_p, [_vp]),
    "feast_contour_circular_trapezoidal": (_i, [c128, _d, _i, _vp, _vp]),
    "feast_contour_circular_gauss": (_i, [c128, _d, _i, _vp, _vp]),
    "feast_contour_rectangular_gauss": (_i, [c128, c128, _i, _vp, _vp]),
    "feast_contour_rectangular_trapezoidal": (_i, [c128, c128, _i, _vp, _vp]),
    "feast_gauss_legendre": (_i, [_i, _vp, _vp]),
    "feast_set_dense": (_i, [_vp, _i, _i64, _vp, _i64, _i]),
    "feast_set_csc": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _i, _i]),
    "feast_set_identity": (_i, [_vp, _i, _i64]),
    "feast_set_problem": (_i, [_vp, _i, _i]),
    "feast_set_contour": (_i, [_vp, _i, _vp, _vp]),
    "feast_set_solver": (_i, [_vp, _i, _i, _d, _i, _i]),
    "feast_comm_unique_id": (_i, [_vp]),
    "feast_comm_init": (_i, [_vp, _i, _i, _vp]),
    "feast_set_node_owners": (_i, [_vp, _i, _vp]),
    "feast_set_subspace": (_i, [_vp, _i64, _i, _vp, _i64]),
    "feast_set_X": (_i, [_vp, _vp, _i64]),
    "feast_get_X": (_i, [_vp, _vp, _i64]),
    "feast_get_Q": (_i, [_vp, _vp, _i64]),
    "feast_get_R": (_i, [_vp, _vp, _i64]),
    "feast_project": (_i, [_vp, _vp, _vp]),
    "feast_recover_residual": (_i, [_vp, _vp, _vp, _vp]),
    "feast_contour_apply": (_i, [_vp, _vp, _i, C.POINTER(FeastStats)]),
    "feast_set_sample_dense": (_i, [_vp, _i64, _vp, _i64, _i]),
    "feast_set_sample_csc": (_i, [_vp, _i64, _vp, _vp, _vp, _i, _i]),
    "feast_contour_node": (_i, [_vp, _i, _vp, _i, _i, C.POINTER(FeastStats)]),
    "feast_node_needs_sample": (_i, [_vp, _i]),
    "feast_sampled_residual": (_i, [_vp, _i, _d, C.POINTER(C.c_double)]),
    "feast_set_sharding": (_i, [_vp, _i]),
    "feast_set_moments": (_i, [_vp, _i]),
    "feast_block_gram": (_i, [_vp, _i, _i, _vp]),
    "feast_moment_combine": (_i, [_vp, _i, _vp, _i64]),
    "feast_last_fro": (_i, [_vp, _vp]),
    "feast_beyn_reduce": (_i, [_vp, _vp, _vp]),
    "feast_orthonormalize_X": (_i, [_vp]),
    "feast_dual_set_subspace": (_i, [_vp, _i64, _i, _vp, _i64, _vp, _i64]),
    "feast_dual_project": (_i, [_vp, _vp]),
    "feast_dual_rotate": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "feast_dual_recover_residual": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "feast_dual_contour_apply": (_i, [_vp, _vp, C.POINTER(FeastStats)]),
    "feast_dual_get": (_i, [_vp, _vp, _i64, _vp, _i64]),
    "feast_estimate_count": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(FeastStats)]),
    "feast_factorize": (_i, [_vp, _vp, _i, C.POINTER(_vp)]),
    "feast_solve": (_i, [_vp, _vp, _i64, _i, _vp, _i64, _vp, _i64, _i]),
    "feast_factor_free": (_i, [_vp, _vp]),
    "feast_apply_operator": (_i, [_vp, _i, _i, _vp, _i64, _i, C.POINTER(C.c_float)]),
    "feast_kernel_bench": (_i, [_vp, _i, _i, C.POINTER(C.c_float)]),
    "feast_sync": (_i, [_vp]),
    "feast_timer_start": (_i, [_vp]),
    "feast_timer_stop": (_i, [_vp, C.POINTER(C.c_float)]),
    "feast_launch_count": (_i64, [_vp]),
    "feast_phase_times": (_i, [_vp, _vp, _i]),
    "feast_set_mixed_precision": (_i, [_vp, _i]),
    "feast_set_preconditioner": (_i, [_vp, _i]),
    "feast_set_preconditioner_shift": (_i, [_vp, _d]),
    "feast_preconditioner_info": (_i, [_vp, C.POINTER(_i), _vp, _i, C.POINTER(C.c_double)]),
    "feast_layout_info": (_i, [_vp, _vp, C.POINTER(C.c_double)]),
    "feast_debug_amg_build": (_vp, [_i64, _vp, _vp, _i, _vp, _i, C.POINTER(_i), C.POINTER(C.c_double)]),
    "feast_debug_amg_level_info": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(C.c_double)]),
    "feast_debug_amg_level_get": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "feast_debug_amg_free": (None, [_vp]),
    "feast_debug_pick_groups": (_i, [_i, _vp, _i, _i, _vp]),
    "feast_debug_cholqr": (_i, [_i64, _i, _vp, _i64, _vp, C.POINTER(_i)]),
    "feast_debug_tile_plan": (_i, [_i64, _vp, _vp, _i, _i, _i, _i, _i, _vp, C.POINTER(_i), C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """dlopen libfeast_cuda.so (built by feastsolver_jl_b200.build).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FeastError(-1, f"{LIB_PATH} not found: run `python -m feastsolver_jl_b200.build` "
                             "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def cplx(z):
    z = complex(z)
    return c128(z.real, z.imag)


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def check(rc, ctx=None, allow=(FEAST_OK,)):
    if rc in allow:
        return rc
    msg = load().feast_last_error(ctx)
    raise FeastError(rc, msg.decode() if msg else "unknown error")


def as_f_c128(a):
    """Column-major complex128 view/copy (what Julia's Matrix{ComplexF64} is)."""
    return np.asfortranarray(a, dtype=np.complex128)
