"""Host-side mirror of the reference drivers over the C ABI.

`feast`, `gen_feast`, `nlfeast` keep the names, argument order, keyword names,
defaults, in-place mutation of X and return tuples of `feast!`, `gen_feast!`,
`nlfeast!` (src/feast.jl:3-156, src/nlfeast.jl:2-84).  The outer loop below is
the reference's loop with each inner block replaced by ONE call into
libfeast_cuda.so; the m0 x m0 reduced eigenproblem / SVD stays on host LAPACK
(scipy here, LinearAlgebra.eigen!/svd! in the Julia shim).

There is no CPU fallback: every compute call goes through the CUDA library and
raises FeastError when it is missing or no device is present.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
from scipy.sparse.linalg import norm as spla_norm

from . import _lib
from ._lib import FeastError, FeastStats
from .contour import (CircularContour, Contour, circular_contour_trapezoidal, in_contour)
from .partition import node_owners

I = "I"  # stand-in for Julia's UniformScaling `I` as the B argument


def _eig_sorted(Aq, Bq=None):
    """eigen!(Aq) / eigen!(Aq, Bq) with Julia's (real, imag) ordering of the values."""
    if Bq is None:
        w, v = sla.eig(Aq, check_finite=False)
    else:
        w, v = sla.eig(Aq, Bq, check_finite=False)
    p = np.lexsort((w.imag, w.real))
    return np.ascontiguousarray(w[p]), np.asfortranarray(v[:, p])


class FeastContext:
    """Owns a feast_ctx*: device-resident operators, subspace blocks, stored factors."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        if device is None:
            device = _default_device()
        h = C.c_void_p()
        _lib.check(self.lib.feast_ctx_create(C.byref(h), int(device)))
        self.h = h
        self.device = int(device)
        self.n = 0
        self.m0 = 0
        self.nranks, self.rank = 1, 0
        self.last_stats = None

    # -- lifetime
    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.feast_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, allow=(0,)):
        return _lib.check(rc, self.h, allow)

    # -- operators
    def set_operator(self, slot, M, n=None):
        if M is None or (isinstance(M, str) and M == I):
            self._ck(self.lib.feast_set_identity(self.h, slot, int(n)))
            return
        if sp.issparse(M):
            M = sp.csc_matrix(M)
            M.sort_indices()
            if M.shape[0] != M.shape[1]:
                raise ValueError("Incorrect dimensions of A, must be square")
            is_c = np.iscomplexobj(M.data)
            data = np.ascontiguousarray(M.data, dtype=np.complex128 if is_c else np.float64)
            indptr = np.ascontiguousarray(M.indptr, dtype=np.int64)
            indices = np.ascontiguousarray(M.indices, dtype=np.int64)
            self._ck(self.lib.feast_set_csc(self.h, slot, M.shape[0], _lib.ptr(indptr), _lib.ptr(indices),
                                            _lib.ptr(data), int(is_c), 0))
            return
        M = np.asarray(M)
        if M.ndim != 2 or M.shape[0] != M.shape[1]:
            raise ValueError("Incorrect dimensions of A, must be square")
        is_c = np.iscomplexobj(M)
        Mf = np.asfortranarray(M, dtype=np.complex128 if is_c else np.float64)  # also converts integer A
        self._ck(self.lib.feast_set_dense(self.h, slot, Mf.shape[0], _lib.ptr(Mf), Mf.shape[0], int(is_c)))

    def set_problem(self, kind, nslots, n):
        self._ck(self.lib.feast_set_problem(self.h, kind, nslots))
        self.n = n

    def set_contour(self, nodes, weights):
        z = np.ascontiguousarray(nodes, dtype=np.complex128)
        w = np.ascontiguousarray(weights, dtype=np.complex128)
        self._ck(self.lib.feast_set_contour(self.h, len(z), _lib.ptr(z), _lib.ptr(w)))

    def set_solver(self, kind=_lib.SOLVER_AUTO, krylov=_lib.KRYLOV_AUTO, inner_tol=1e-8, max_inner=4000, store=False, precond=None,
                   shard=None, precond_shift=None):
        if precond is not None:   # before set_solver: one layout rebuild at most
            self.set_preconditioner(precond)
        if precond_shift is not None:
            self._ck(self.lib.feast_set_preconditioner_shift(self.h, float(precond_shift)))
        if shard is not None:
            self.set_sharding(shard)
        self._ck(self.lib.feast_set_solver(self.h, kind, krylov, float(inner_tol), int(max_inner), int(bool(store))))

    def set_preconditioner(self, kind=_lib.PRECOND_AUTO):
        """Krylov preconditioner: PRECOND_NONE / PRECOND_AMG (smoothed-aggregation V-cycle) / PRECOND_AUTO."""
        self._ck(self.lib.feast_set_preconditioner(self.h, int(kind)))

    def preconditioner_info(self):
        nl, secs = C.c_int(0), C.c_double(0.0)
        sizes = np.zeros(16, dtype=np.int32)
        self._ck(self.lib.feast_preconditioner_info(self.h, C.byref(nl), _lib.ptr(sizes), 16, C.byref(secs)))
        return {"levels": int(nl.value), "sizes": [int(v) for v in sizes[:nl.value]], "setup_s": float(secs.value)}

    def layout_info(self):
        """Internal layout of the sparse path: renumbered?, tiles, natural bandwidth, tiled SpMM in use, halo rows per row."""
        info = np.zeros(4, dtype=np.int32)
        halo = C.c_double(0.0)
        self._ck(self.lib.feast_layout_info(self.h, _lib.ptr(info), C.byref(halo)))
        return {"reordered": bool(info[0]), "ntiles": int(info[1]), "bandwidth": int(info[2]), "tiled_spmm": bool(info[3]),
                "halo_rows_per_row": halo.value}

    def set_mixed_precision(self, on=True):
        self._ck(self.lib.feast_set_mixed_precision(self.h, int(bool(on))))

    def set_node_owners(self, owners):
        o = np.ascontiguousarray(owners, dtype=np.int32)
        self._ck(self.lib.feast_set_node_owners(self.h, len(o), _lib.ptr(o)))

    def comm_init(self, nranks, rank, uid: bytes | None):
        buf = (C.c_char * 128).from_buffer_copy(uid) if uid is not None else None
        self._ck(self.lib.feast_comm_init(self.h, nranks, rank, buf))
        self.nranks, self.rank = nranks, rank

    # -- subspace
    def set_subspace(self, X):
        Xf = _lib.as_f_c128(X)
        n, m0 = Xf.shape
        self._ck(self.lib.feast_set_subspace(self.h, n, m0, _lib.ptr(Xf), n))
        self.m0 = m0

    def set_X(self, X):
        """Overwrite the device block X only (Q / moment accumulators untouched)."""
        Xf = _lib.as_f_c128(X)
        self._ck(self.lib.feast_set_X(self.h, _lib.ptr(Xf), Xf.shape[0]))

    def _get(self, fn):
        out = np.empty((self.n, self.m0), dtype=np.complex128, order="F")
        self._ck(fn(self.h, _lib.ptr(out), self.n))
        return out

    def get_X(self):
        return self._get(self.lib.feast_get_X)

    def get_Q(self):
        return self._get(self.lib.feast_get_Q)

    def get_R(self):
        return self._get(self.lib.feast_get_R)

    # -- phases
    def project(self, generalized):
        m = self.m0
        Aq = np.empty((m, m), np.complex128, order="F")
        Bq = np.empty((m, m), np.complex128, order="F") if generalized else None
        self._ck(self.lib.feast_project(self.h, _lib.ptr(Aq), _lib.ptr(Bq)))
        return Aq, Bq

    def recover_residual(self, Xq, lam):
        """X = Q Xq (Xq None: X as it is), x_j /= ||x_j||, residual vectors and norms."""
        Xq_p = None
        if Xq is not None:
            Xq = np.asfortranarray(Xq, dtype=np.complex128)
            Xq_p = _lib.ptr(Xq)
        lam = np.ascontiguousarray(lam, dtype=np.complex128)
        res = np.empty(self.m0, np.float64)
        self._ck(self.lib.feast_recover_residual(self.h, Xq_p, _lib.ptr(lam), _lib.ptr(res)))
        return res

    def last_fro(self):
        fro = np.empty(self.m0, np.float64)
        self._ck(self.lib.feast_last_fro(self.h, _lib.ptr(fro)))
        return fro

    def set_sharding(self, mode=_lib.SHARD_AUTO):
        self._ck(self.lib.feast_set_sharding(self.h, int(mode)))

    # -- moments (beyn / block_SS! / nlfeast_moments!)
    def set_moments(self, nmom):
        self._ck(self.lib.feast_set_moments(self.h, int(nmom)))

    def block_gram(self, a, b):
        """block(a)^H block(b); ids: p >= 0 moment S_p, -1 X, -2 R."""
        G = np.empty((self.m0, self.m0), np.complex128, order="F")
        self._ck(self.lib.feast_block_gram(self.h, int(a), int(b), _lib.ptr(G)))
        return G

    def moment_combine(self, W):
        """X = sum_p S_p W[p*m0:(p+1)*m0, :]."""
        W = np.asfortranarray(W, dtype=np.complex128)
        nblk = W.shape[0] // self.m0
        assert W.shape == (nblk * self.m0, self.m0)
        self._ck(self.lib.feast_moment_combine(self.h, nblk, _lib.ptr(W), W.shape[0]))

    def contour_apply(self, lam, first_pass=False):
        st = FeastStats()
        lam_p = None
        if lam is not None:
            lam = np.ascontiguousarray(lam, dtype=np.complex128)
            lam_p = _lib.ptr(lam)
        rc = self.lib.feast_contour_apply(self.h, lam_p, int(first_pass), C.byref(st))
        self._ck(rc, allow=(0, _lib.FEAST_WARN_INNER_MAXIT))
        self.last_stats = st.as_dict()
        self.last_stats["warn_inner_maxit"] = rc == _lib.FEAST_WARN_INNER_MAXIT
        return self.last_stats

    # -- sampled operators (opaque T(z) closures)
    def set_sample(self, M):
        """Replace the sample T(point) held in slot 0 of a PROBLEM_SAMPLED problem (same shape / storage class)."""
        if sp.issparse(M):
            M = sp.csc_matrix(M)
            M.sort_indices()
            is_c = np.iscomplexobj(M.data)
            data = np.ascontiguousarray(M.data, dtype=np.complex128 if is_c else np.float64)
            indptr = np.ascontiguousarray(M.indptr, dtype=np.int64)
            indices = np.ascontiguousarray(M.indices, dtype=np.int64)
            self._ck(self.lib.feast_set_sample_csc(self.h, M.shape[0], _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data),
                                                   int(is_c), 0))
            return
        M = np.asarray(M)
        is_c = np.iscomplexobj(M)
        Mf = np.asfortranarray(M, dtype=np.complex128 if is_c else np.float64)
        self._ck(self.lib.feast_set_sample_dense(self.h, Mf.shape[0], _lib.ptr(Mf), Mf.shape[0], int(is_c)))

    def contour_node(self, k, lam, first_pass, phase):
        st = FeastStats()
        lam_p = None
        if lam is not None:
            lam = np.ascontiguousarray(lam, dtype=np.complex128)
            lam_p = _lib.ptr(lam)
        rc = self.lib.feast_contour_node(self.h, int(k), lam_p, int(first_pass), int(phase), C.byref(st))
        self._ck(rc, allow=(0, _lib.FEAST_WARN_INNER_MAXIT))
        self.last_stats = st.as_dict()
        self.last_stats["warn_inner_maxit"] = rc == _lib.FEAST_WARN_INNER_MAXIT
        return self.last_stats

    def node_needs_sample(self, k):
        return self.lib.feast_node_needs_sample(self.h, int(k)) != 0

    def sampled_residual(self, j, fro):
        res = C.c_double(0.0)
        self._ck(self.lib.feast_sampled_residual(self.h, int(j), float(fro), C.byref(res)))
        return float(res.value)

    def beyn_reduce(self):
        m = self.m0
        Rf = np.empty((m, m), np.complex128, order="F")
        G1 = np.empty((m, m), np.complex128, order="F")
        self._ck(self.lib.feast_beyn_reduce(self.h, _lib.ptr(Rf), _lib.ptr(G1)))
        self._ck(self.lib.feast_sync(self.h))
        return Rf, G1

    def estimate_count(self):
        est = C.c_double(0.0)
        st = FeastStats()
        rc = self.lib.feast_estimate_count(self.h, C.byref(est), C.byref(st))
        self._ck(rc, allow=(0, _lib.FEAST_WARN_INNER_MAXIT))
        self.last_stats = st.as_dict()
        return float(est.value)

    # -- two-sided driver
    def dual_set_subspace(self, Xr, Xl):
        Xr, Xl = _lib.as_f_c128(Xr), _lib.as_f_c128(Xl)
        n, m0 = Xr.shape
        self._ck(self.lib.feast_dual_set_subspace(self.h, n, m0, _lib.ptr(Xr), n, _lib.ptr(Xl), n))
        self.m0 = m0

    def dual_project(self):
        G = np.empty((self.m0, self.m0), np.complex128, order="F")
        self._ck(self.lib.feast_dual_project(self.h, _lib.ptr(G)))
        return G

    def dual_rotate(self, Mr, Ml):
        m = self.m0
        Mr, Ml = np.asfortranarray(Mr, dtype=np.complex128), np.asfortranarray(Ml, dtype=np.complex128)
        Aq = np.empty((m, m), np.complex128, order="F")
        Bq = np.empty((m, m), np.complex128, order="F")
        self._ck(self.lib.feast_dual_rotate(self.h, _lib.ptr(Mr), _lib.ptr(Ml), _lib.ptr(Aq), _lib.ptr(Bq)))
        return Aq, Bq

    def dual_recover_residual(self, Xqr, Xql, lam):
        Xqr, Xql = np.asfortranarray(Xqr, dtype=np.complex128), np.asfortranarray(Xql, dtype=np.complex128)
        lam = np.ascontiguousarray(lam, dtype=np.complex128)
        res = np.empty(self.m0, np.float64)
        self._ck(self.lib.feast_dual_recover_residual(self.h, _lib.ptr(Xqr), _lib.ptr(Xql), _lib.ptr(lam), _lib.ptr(res)))
        return res

    def dual_contour_apply(self, lam):
        st = FeastStats()
        lam = np.ascontiguousarray(lam, dtype=np.complex128)
        rc = self.lib.feast_dual_contour_apply(self.h, _lib.ptr(lam), C.byref(st))
        self._ck(rc, allow=(0, _lib.FEAST_WARN_INNER_MAXIT))
        self.last_stats = st.as_dict()
        return self.last_stats

    def dual_get(self):
        Xr = np.empty((self.n, self.m0), np.complex128, order="F")
        Xl = np.empty((self.n, self.m0), np.complex128, order="F")
        self._ck(self.lib.feast_dual_get(self.h, _lib.ptr(Xr), self.n, _lib.ptr(Xl), self.n))
        return Xr, Xl

    def orthonormalize_X(self):
        self._ck(self.lib.feast_orthonormalize_X(self.h))

    # -- fine grained plugin path (factorizer / left_divider / finalize!)
    def factorize(self, coefs):
        cf = np.ascontiguousarray(coefs, dtype=np.complex128)
        F = C.c_void_p()
        self._ck(self.lib.feast_factorize(self.h, _lib.ptr(cf), len(cf), C.byref(F)))
        return F

    def solve(self, F, B, conj_transpose=False):
        Bf = _lib.as_f_c128(B)
        n, m = Bf.shape
        Y = np.empty((n, m), np.complex128, order="F")
        self._ck(self.lib.feast_solve(self.h, F, n, m, _lib.ptr(Bf), n, _lib.ptr(Y), n, int(conj_transpose)),
                 allow=(0, _lib.FEAST_WARN_INNER_MAXIT))
        self.m0 = m
        return Y

    def factor_free(self, F):
        self.lib.feast_factor_free(self.h, F)

    # -- kernel level
    def apply_operator(self, slot, which=0, download=True, reps=0):
        Y = np.empty((self.n, self.m0), np.complex128, order="F") if download else None
        ms = C.c_float(0.0)
        self._ck(self.lib.feast_apply_operator(self.h, slot, which, _lib.ptr(Y), self.n, int(reps),
                                               C.byref(ms) if reps > 0 else None))
        return Y, float(ms.value)

    def kernel_time(self, which="cocg_direction", reps=20):
        """Mean device ms of one stand-alone launch of a Krylov vector kernel on the resident blocks (measurement only)."""
        ms = C.c_float(0.0)
        self._ck(self.lib.feast_kernel_bench(self.h, {"cocg_direction": 0, "cocg_update": 1}[which], int(reps), C.byref(ms)))
        return float(ms.value)

    def sync(self):
        self._ck(self.lib.feast_sync(self.h))

    def timer_start(self):
        self._ck(self.lib.feast_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0.0)
        self._ck(self.lib.feast_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self):
        return int(self.lib.feast_launch_count(self.h))

    def phase_times(self, reset=False):
        t = np.zeros(3)
        self._ck(self.lib.feast_phase_times(self.h, _lib.ptr(t), int(reset)))
        return {"project_ms": t[0], "recover_ms": t[1], "contour_apply_ms": t[2]}


def _default_device():
    import os
    return int(os.environ.get("LOCAL_RANK", "0"))


def _check_plugins(factorizer, left_divider, mixed_prec, solver_opts=None):
    if factorizer is not None or left_divider is not None:
        raise FeastError(-1, "custom factorizer/left_divider callbacks cannot run inside libfeast_cuda "
                             "(no CPU fallback); use set_solver options instead")
    if mixed_prec and (solver_opts or {}).get("kind") != _lib.SOLVER_KRYLOV:
        raise FeastError(-1, "mixed_prec=true exists (experimental, complex64 Krylov blocks) for Krylov inner solves only: "
                             "pass solver_opts={'kind': SOLVER_KRYLOV}; the dense LU path has no single-precision variant")


def _densify_if_mixed(A, B):
    """The library needs all operators dense or all sparse (feast_set_problem)."""
    if B is None or (isinstance(B, str)):
        return A, B
    if sp.issparse(A) != sp.issparse(B):
        A = A.toarray() if sp.issparse(A) else A
        B = B.toarray() if sp.issparse(B) else B
    return A, B


def iter_debug_print(nit, Lam, res, contour, spurious=1e-5):
    """src/utils.jl:23-42"""
    ins = in_contour(Lam, contour)
    in_res = res[ins]
    line = f"{nit}:\t{int(ins.sum())} ({int((in_res < spurious).sum())})\t"
    if ins.sum() > 0:
        line += f"{in_res.max()}"
        conv = in_res[in_res < spurious]
        if conv.size:
            line += f"\t({conv.max()})"
    print(line)


def _linear_driver(X, A, B, contour, iter, eps, debug, store, ctx, solver_opts, generalized, stats_out, comm,
                   mixed_prec=False):
    import time as _time
    N, m0 = X.shape
    if A.shape[0] != A.shape[1]:
        raise ValueError("Incorrect dimensions of A, must be square")  # feast.jl:13
    if A.shape[0] != N:
        raise ValueError("Incorrect dimensions of X, must match A")  # feast.jl:15
    ph = {}          # wall-clock phase breakdown of the call (seconds), returned in stats["phases"]
    t_last = [_time.perf_counter()]

    def tick(name):
        now = _time.perf_counter()
        ph[name] = ph.get(name, 0.0) + (now - t_last[0])
        t_last[0] = now

    own_ctx = ctx is None
    if own_ctx:
        ctx = FeastContext()
    tick("ctx_create_s")
    try:
        A, B = _densify_if_mixed(A, B)
        comm_thread, comm_err = None, []
        if comm is not None:   # NCCL bootstrap (id broadcast + ncclCommInitRank, ~1 s) overlaps the host-side layout build
            import threading

            def _comm():
                try:
                    comm(ctx)
                except BaseException as e:   # re-raised on the calling thread below
                    comm_err.append(e)
            comm_thread = threading.Thread(target=_comm)
            comm_thread.start()
        # inner-tolerance schedule (the reference's own precedent: 1e-3 in the first pass, tight afterwards, nlfeast.jl:106,139):
        # the first contour pass starts from a random subspace and cannot gain more than the filter's contraction anyway
        solver_opts = dict(solver_opts)
        first_tol = solver_opts.pop("first_pass_tol", None)
        later_tol = solver_opts.get("inner_tol", 1e-8)
        if first_tol is not None:
            solver_opts["inner_tol"] = max(float(first_tol), float(later_tol))
        ctx.set_solver(store=store, **solver_opts)   # before the operators: the device layout is then built once
        ctx.set_operator(0, A)
        if generalized:
            ctx.set_operator(1, B, n=N)
        tick("upload_operators_s")                   # host CSC -> CSR (+ symmetry check), dense uploads
        ctx.set_problem(_lib.PROBLEM_GENERALIZED if generalized else _lib.PROBLEM_STANDARD, 2 if generalized else 1, N)
        tick("build_layout_s")                       # union pattern, tile plan, multigrid hierarchy, uploads
        if comm_thread is not None:
            comm_thread.join()
            if comm_err:
                raise comm_err[0]
        tick("comm_init_wait_s")
        ctx.set_contour(contour.nodes, contour.weights)
        if ctx.nranks > 1:  # balanced node -> rank map (near-axis nodes cost more Krylov iterations)
            ctx.set_node_owners(node_owners(contour.nodes, ctx.nranks))
        if mixed_prec:
            ctx.set_mixed_precision(True)   # feast.jl:19-25 (complex64 COCG blocks on the device)
        ctx.set_subspace(X)
        tick("upload_subspace_s")
        Lam = np.zeros(m0, complex)
        res = np.zeros(m0)
        hist = []
        for nit in range(iter + 1):  # for nit=0:iter
            Aq, Bq = ctx.project(generalized)  # feast.jl:41-43 / 117-121
            Lam, Xq = _eig_sorted(Aq, Bq)  # feast.jl:45-47 / 122-124 (host LAPACK)
            res = ctx.recover_residual(Xq, Lam)  # feast.jl:48-50 / 125-127
            tick("rayleigh_ritz_s")
            inside = in_contour(Lam, contour)
            if debug:
                iter_debug_print(nit, Lam, res, contour, 1e-5)
            rec = {"nit": nit, "inside": int(inside.sum()),
                   "max_res_inside": float(res[inside].max()) if inside.any() else float("nan")}
            if inside.any() and res[inside].max() < eps:  # feast.jl:53
                hist.append(rec)
                if debug:
                    print(f"converged in {nit} iteration")
                break
            if nit < iter:  # feast.jl:57
                st = ctx.contour_apply(Lam)
                rec.update(st)
                rec["inner_tol"] = solver_opts["inner_tol"]
                if first_tol is not None and nit == 0:
                    solver_opts["inner_tol"] = later_tol
                    ctx.set_solver(store=store, **solver_opts)   # same solver kind: the device layout is kept
                tick("contour_passes_s")
                ph.setdefault("contour_pass_s", []).append(st["t_total_ms"] / 1e3)
                ph.setdefault("contour_pass_allreduce_s", []).append(st["t_reduce_ms"] / 1e3)
                ph.setdefault("contour_pass_col_slices", []).append(st.get("col_sharded", 0))
            hist.append(rec)
        X[:, :] = ctx.get_X()
        tick("download_s")
        if stats_out is not None:
            stats_out["history"] = hist
            stats_out["Lam_all"] = Lam
            stats_out["res_all"] = res
            stats_out["launches"] = ctx.launch_count()
            stats_out["phase_ms"] = ctx.phase_times()
            stats_out["phases"] = ph
            stats_out["preconditioner"] = ctx.preconditioner_info()
    finally:
        if own_ctx:
            ctx.close()
    inside = in_contour(Lam, contour)
    if not inside.any():
        print("no eigenvalues found in contour!")  # feast.jl:78
    return Lam[inside], X[:, inside], res[inside]


def feast(X, A, contour: Contour | None = None, *, nodes=8, iter=10, c=complex(0.0, 0.0), r=1.0, eps=1e-12,
          debug=False, store=False, mixed_prec=False, factorizer=None, left_divider=None,
          ctx=None, solver_opts=None, stats=None, comm=None):
    """feast!(X, A; ...) and feast!(X, A, contour; ...)  (src/feast.jl:3-80).

    X (N x m0 complex128) is mutated in place and ends up holding all m0 unit-norm
    Ritz vectors; returns (L[in], X[:, in], res[in]) filtered by in_contour.  `eps` is
    the reference's keyword ϵ.
    """
    _check_plugins(factorizer, left_divider, mixed_prec, solver_opts)
    if contour is None:
        contour = circular_contour_trapezoidal(c, r, nodes)  # feast.jl:6
    return _linear_driver(X, A, None, contour, iter, eps, debug, store, ctx, solver_opts or {}, False, stats, comm,
                          mixed_prec=bool(mixed_prec))


def gen_feast(X, A, B, contour: Contour | None = None, *, nodes=8, iter=10, c=complex(0.0, 0.0), r=1.0,
              debug=False, store=False, eps=1e-12, factorizer=None, left_divider=None,
              ctx=None, solver_opts=None, stats=None, comm=None):
    """gen_feast!(X, A, B; ...) and gen_feast!(X, A, B, contour; ...)  (src/feast.jl:82-156)."""
    _check_plugins(factorizer, left_divider, False)
    if contour is None:
        contour = circular_contour_trapezoidal(c, r, nodes)  # feast.jl:85
        store = False  # the reference's convenience wrapper drops `store` (feast.jl:86)
    return _linear_driver(X, A, B, contour, iter, eps, debug, store, ctx, solver_opts or {}, True, stats, comm)


def nlfeast(T, X, nodes, iter, *, c=complex(0.0, 0.0), r=1.0, debug=False, eps=10e-12, store=True,
            spurious=1e-5, factorizer=None, left_divider=None, ctx=None, solver_opts=None, stats=None, comm=None,
            _stop_rule="nlfeast"):
    """nlfeast!(T, X, nodes, iter; ...)  (src/nlfeast.jl:2-84) -> (L, X, res), all m0, unfiltered.

    `T` is either the reference's callable z -> N x N matrix (dense ndarray or scipy sparse), or -- ADDED METHOD
    (SURVEY 8b), the fast path -- the list of polynomial coefficient matrices [A_0, ..., A_d] with
    T(z) = sum z^i A_i, so that assembly, solves and residuals all run on the device.  A callable is evaluated on the
    HOST: once per contour node (and only in the first pass when store=true keeps the factorisations) and once per
    Ritz value for the residuals (src/utils.jl:107,154 do the same m0 evaluations), each sample being uploaded.
    """
    _check_plugins(factorizer, left_divider, False)
    sampled = callable(T)
    coeffs = None if sampled else list(T)
    N, m0 = X.shape
    if not sampled and any(sp.issparse(a) for a in coeffs) and not all(sp.issparse(a) for a in coeffs):
        coeffs = [a.toarray() if sp.issparse(a) else np.asarray(a) for a in coeffs]
    own_ctx = ctx is None
    if own_ctx:
        ctx = FeastContext()
    try:
        contour = circular_contour_trapezoidal(c, r, nodes)  # nlfeast.jl:8 hard-wires circle + trapezoid
        if sampled:
            T0 = T(contour.nodes[0])
            sparse_T = sp.issparse(T0)
            if T0.shape != (N, N):
                raise ValueError("Incorrect dimensions of X, must match T")
            ctx.set_operator(0, T0, n=N)
            ctx.set_problem(_lib.PROBLEM_SAMPLED, 1, N)
        else:
            for i, Ai in enumerate(coeffs):
                ctx.set_operator(i, Ai, n=N)
            ctx.set_problem(_lib.PROBLEM_POLYNOMIAL, len(coeffs), N)
        if comm is not None:
            comm(ctx)
        ctx.set_contour(contour.nodes, contour.weights)
        owners = node_owners(contour.nodes, ctx.nranks) if ctx.nranks > 1 else [0] * nodes
        if ctx.nranks > 1:
            ctx.set_node_owners(owners)
        sched = getattr(solver_opts, "tol_schedule", None)    # (first pass, later passes): nlfeast_it
        base_opts = dict(solver_opts or {})
        if sched is not None:
            base_opts["inner_tol"] = sched[0]
        ctx.set_solver(store=store, **base_opts)
        ctx.set_subspace(X)
        ctx.orthonormalize_X()  # nlfeast.jl:12-13
        Lam = np.zeros(m0, complex)
        res = np.zeros(m0)
        hist = []
        for nit in range(iter + 1):
            if sched is not None and nit == 1:
                base_opts["inner_tol"] = sched[1]
                ctx.set_solver(store=store, **base_opts)   # same solver kind: the device layout is kept
            if sampled:   # nlfeast.jl:36-61 with T(z_k) evaluated here and uploaded, node by node
                mine = [k for k in range(nodes) if owners[k] == ctx.rank] or [-1]
                for i, k in enumerate(mine):
                    if k >= 0 and ctx.node_needs_sample(k):
                        ctx.set_sample(T(contour.nodes[k]))
                    st = ctx.contour_node(k, Lam if nit > 0 else None, nit == 0, (1 if i == 0 else 0) | (2 if i == len(mine) - 1 else 0))
            else:
                st = ctx.contour_apply(Lam if nit > 0 else None, first_pass=(nit == 0))  # nlfeast.jl:36-61
            Rf, G1 = ctx.beyn_reduce()  # utils.jl:70-71 (tall part)
            U, S, Vh = sla.svd(Rf, check_finite=False)  # m0 x m0 (host)
            Am = (U.conj().T @ G1) @ Vh.conj().T * (1.0 / S)[None, :]  # utils.jl:71-73
            w, v = sla.eig(Am, check_finite=False)  # utils.jl:74
            p = np.lexsort((w.imag, w.real))
            Lam = np.ascontiguousarray(w[p])
            Xq = U @ v[:, p]  # utils.jl:75  X = U * vectors
            res = ctx.recover_residual(Xq, Lam)  # nlfeast.jl:66-67
            if sampled:   # update_R! / residuals with T(l_j) evaluated on the host (utils.jl:104-109, 151-157)
                for j in range(m0):
                    Tj = T(Lam[j])
                    ctx.set_sample(Tj)
                    fro = spla_norm(Tj) if sp.issparse(Tj) else float(np.linalg.norm(Tj))
                    res[j] = ctx.sampled_residual(j, fro)
            inside = in_contour(Lam, c, r)
            res_inside = res[inside]
            rec = {"nit": nit, "inside": int(inside.sum()),
                   "max_res_inside": float(res_inside.max()) if inside.any() else float("nan")}
            rec.update(st)
            hist.append(rec)
            if debug:
                iter_debug_print(nit, Lam, res, CircularContour(c, r, contour.nodes, contour.weights), spurious)
            if _stop_rule == "it":
                if nit >= 1 and res_inside.size > 0 and res_inside.max() < eps:  # nlfeast.jl:164 (checked from the 2nd pass on)
                    break
                continue
            if res_inside.size > 0 and res_inside.max() < eps:  # nlfeast.jl:73
                break
            good = res_inside[res_inside < spurious]
            if nit > 1 and good.size > 0 and good.max() < eps:  # nlfeast.jl:76
                break
        X[:, :] = ctx.get_X()  # already unit-norm columns (normalize!, nlfeast.jl:82)
        if stats is not None:
            stats["history"] = hist
            stats["launches"] = ctx.launch_count()
            stats["phase_ms"] = ctx.phase_times()
    finally:
        if own_ctx:
            ctx.close()
    return Lam, X, res


class _NepSession:
    """Shared plumbing of the moment-based nonlinear drivers: a context holding T (polynomial coefficients on the
    device, or a host-evaluated closure through the sampled-operator entries), one contour pass with `nmom` moment
    accumulators, and the residuals of the current block X."""

    def __init__(self, T, X, nodes, c, r, store, solver_opts, comm, ctx=None):
        self.sampled = callable(T)
        self.T = T
        N, m0 = X.shape
        self.N, self.m0, self.nodes = N, m0, nodes
        self.own = ctx is None
        self.ctx = FeastContext() if ctx is None else ctx
        self.contour = circular_contour_trapezoidal(c, r, nodes)      # theta = LinRange(pi/nodes, 2pi - pi/nodes, nodes)
        ctx = self.ctx
        if self.sampled:
            T0 = T(self.contour.nodes[0])
            if T0.shape != (N, N):
                raise ValueError("Incorrect dimensions of X, must match T")
            ctx.set_operator(0, T0, n=N)
            ctx.set_problem(_lib.PROBLEM_SAMPLED, 1, N)
        else:
            coeffs = list(T)
            if any(sp.issparse(a) for a in coeffs) and not all(sp.issparse(a) for a in coeffs):
                coeffs = [a.toarray() if sp.issparse(a) else np.asarray(a) for a in coeffs]
            for i, Ai in enumerate(coeffs):
                ctx.set_operator(i, Ai, n=N)
            ctx.set_problem(_lib.PROBLEM_POLYNOMIAL, len(coeffs), N)
        if comm is not None:
            comm(ctx)
        ctx.set_contour(self.contour.nodes, self.contour.weights)
        self.owners = node_owners(self.contour.nodes, ctx.nranks) if ctx.nranks > 1 else [0] * nodes
        if ctx.nranks > 1:
            ctx.set_node_owners(self.owners)
        ctx.set_solver(store=store, **(solver_opts or {}))
        ctx.set_subspace(X)

    def contour_pass(self, nmom, Lam=None):
        ctx = self.ctx
        ctx.set_moments(nmom)
        first = Lam is None
        if not self.sampled:
            return ctx.contour_apply(Lam, first_pass=first)
        mine = [k for k in range(self.nodes) if self.owners[k] == ctx.rank] or [-1]
        st = None
        for i, k in enumerate(mine):
            if k >= 0 and ctx.node_needs_sample(k):
                ctx.set_sample(self.T(self.contour.nodes[k]))
            st = ctx.contour_node(k, Lam, first, (1 if i == 0 else 0) | (2 if i == len(mine) - 1 else 0))
        return st

    def residuals(self, Lam, Xq=None):
        """x_j /= ||x_j||, R_j = T(l_j) x_j; returns (relative residuals, ||T(l_j)||_F)."""
        ctx = self.ctx
        res = ctx.recover_residual(Xq, Lam)
        if not self.sampled:
            return res, ctx.last_fro()
        fro = np.empty(self.m0)
        for j in range(self.m0):
            Tj = self.T(Lam[j])
            ctx.set_sample(Tj)
            fro[j] = spla_norm(Tj) if sp.issparse(Tj) else float(np.linalg.norm(Tj))
            res[j] = ctx.sampled_residual(j, fro[j])
        return res, fro

    def close(self):
        self.ctx.set_moments(0)
        if self.own:
            self.ctx.close()


def _beyn_small(Rf, G1):
    """m0 x m0 part of beyn_svd_step! (src/utils.jl:70-75) given Q0 = U Rf and G1 = U' Q1."""
    U, S, Vh = sla.svd(Rf, check_finite=False)
    Am = (U.conj().T @ G1) @ Vh.conj().T * (1.0 / S)[None, :]
    w, v = sla.eig(Am, check_finite=False)
    p = np.lexsort((w.imag, w.real))
    return np.ascontiguousarray(w[p]), U @ v[:, p]


def beyn(T, A, X, nodes, *, c=complex(0.0, 0.0), r=1.0, ctx=None, solver_opts=None, comm=None):
    """beyn(T, A, X, nodes; c, r)  (src/beyn.jl:2-34): Beyn's integral method -- ONE contour pass with the two moments
    Q0, Q1, the Beyn reduction, ABSOLUTE residuals ||T(l) x||, everything sorted by residual.  `A` only supplies the
    dimensions, X is not modified.  T: polynomial coefficient list (device assembly) or a callable (host samples).
    (Upstream weights Q0/Q1 by exp(i theta)/nodes without the factor r, beyn.jl:19-20; r cancels in U' Q1 V S^-1.)"""
    N, m0 = X.shape
    if A.shape[0] != A.shape[1]:
        raise ValueError("Incorrect dimensions of A, must be square")          # beyn.jl:5-6
    if A.shape[0] != N:
        raise ValueError("Incorrect dimensions of X0, must match A")           # beyn.jl:7-8
    ses = _NepSession(T, np.asarray(X, dtype=np.complex128), nodes, c, r, False, solver_opts, comm, ctx)
    try:
        ses.contour_pass(2)                                                    # beyn.jl:16-21
        Rf, G1 = ses.ctx.beyn_reduce()                                         # tall part of svd!(Q0), U' Q1
        Lam, Xq = _beyn_small(Rf, G1)                                          # beyn.jl:22-25
        res, fro = ses.residuals(Lam, Xq)
        Xn = ses.ctx.get_X()
    finally:
        ses.close()
    res = res * fro                                                            # beyn.jl:28: no normalisation by ||T||
    p = np.argsort(res, kind="stable")
    return Lam[p], Xn[:, p], res[p]


def block_SS(T, X, nodes=16, moments=2, *, c=complex(0.0, 0.0), r=1.0, debug=False, Y=None, seed=0, ctx=None,
             solver_opts=None, comm=None):
    """block_SS!(T, X, nodes=2^4, moments=2; c, r)  (src/beyn.jl:36-94): block Sakurai-Sugiura.  2*moments+1 moment
    accumulators S_p in one contour pass (the p-loop of the accumulate kernel), Hankel blocks Y' S_p as m0 x m0 Grams on
    the device, the small SVD / generalized eigenproblem on the host, X = [S_0 .. S_{moments-1}] V Xq back on the device.
    Upstream draws the probe Y = rand(ComplexF64, N, m0) unseeded; pass `Y` (or `seed`) for reproducibility."""
    N, m0 = X.shape
    K = moments * m0
    if 2 * moments + 1 > _lib.MAX_MOMENTS:
        raise ValueError(f"moments <= {(_lib.MAX_MOMENTS - 1) // 2}")
    if Y is None:
        rng = np.random.default_rng(seed)
        Y = rng.random((N, m0)) + 1j * rng.random((N, m0))
    ses = _NepSession(T, np.asarray(X, dtype=np.complex128), nodes, c, r, False, solver_opts, comm, ctx)
    try:
        cx = ses.ctx
        cx.orthonormalize_X()                                                  # X = Matrix(qr(X).Q)      beyn.jl:41
        ses.contour_pass(2 * moments + 1)                                      # beyn.jl:50-56
        cx.set_X(Y)                                                            # the probe takes the X block: Grams Y' S_p
        G = [cx.block_gram(-1, p) for p in range(2 * moments + 1)]
        Q0 = np.zeros((m0 * moments, K), complex)
        Q1 = np.zeros((m0 * moments, K), complex)
        for i in range(1, moments + 1):
            for j in range(1, moments + 1):
                Q0[(i - 1) * m0:i * m0, (j - 1) * m0:j * m0] = G[i + j - 1]     # beyn.jl:66
                Q1[(i - 1) * m0:i * m0, (j - 1) * m0:j * m0] = G[i + j]         # beyn.jl:67
        U, sv, Vh = sla.svd(Q0, check_finite=False)
        n = min(int(np.count_nonzero(sv / sv[0] > 1e-13)), K)                   # beyn.jl:78
        V = Vh.conj().T
        H1 = U[:, :n].conj().T @ Q1 @ V[:, :n]
        H0 = U[:, :n].conj().T @ Q0 @ V[:, :n]
        Lam, Xq = sla.eig(H1, H0, check_finite=False)                           # beyn.jl:83
        W = V[:, :n] @ Xq                                                       # X = S[:, 1:K] * V * Xq    beyn.jl:87
        Xn = np.empty((N, n), np.complex128, order="F")
        res = np.empty(n)
        for c0 in range(0, n, m0):                                              # the device block is m0 wide
            c1 = min(n, c0 + m0)
            Wc = np.zeros((K, m0), complex)
            Wc[:, :c1 - c0] = W[:, c0:c1]
            lam_c = np.full(m0, Lam[c0], complex)
            lam_c[:c1 - c0] = Lam[c0:c1]
            if c1 - c0 < m0:
                Wc[:, c1 - c0:] = W[:, [c0]]                                    # pad with a copy (no zero columns to normalise)
            cx.moment_combine(Wc)
            rc, _ = ses.residuals(lam_c)                                        # normalise, relative residuals beyn.jl:90-93
            Xn[:, c0:c1] = cx.get_X()[:, :c1 - c0]
            res[c0:c1] = rc[:c1 - c0]
    finally:
        ses.close()
    return Lam, Xn, res


def nlfeast_moments(T, X, nodes, iter, *, c=complex(0.0, 0.0), r=1.0, debug=False, eps=10e-12, moments=2, store=True,
                    spurious=1e-5, ctx=None, solver_opts=None, comm=None, stats=None):
    """nlfeast_moments!(T, X, nodes, iter; ...)  (src/nlfeast.jl:173-318): nlfeast with 2*moments moment accumulators and a
    moments*m0-dimensional Beyn reduction of the block-Hankel matrices per pass.  Returns (L, Y, res) with moments*m0
    entries sorted by residual; X is overwritten by the m0 best unit-norm vectors.

    Device form of the tall SVD of Q0 (moments*N x moments*m0): its Gram matrix and Q0' Q1 are sums of the m0 x m0 Grams
    S_a' S_b of the moment blocks, Q0 = U S V' follows from the Hermitian eigendecomposition of the Gram matrix, and
    Y = U[1:N, :] vecs = [S_0 .. S_{moments-1}] V S^-1 vecs is one combination of moment blocks.  Directions whose
    singular value is below 1e-7 of the largest carry no information in this form (the Gram matrix squares the
    condition number) and are dropped; upstream divides by them (nlfeast.jl:224), which only produces spurious pairs."""
    N, m0 = X.shape
    M = moments
    if 2 * M > _lib.MAX_MOMENTS:
        raise ValueError(f"moments <= {_lib.MAX_MOMENTS // 2}")
    ses = _NepSession(T, X, nodes, c, r, store, solver_opts, comm, ctx)
    hist = []
    try:
        cx = ses.ctx
        K = M * m0

        def reduce_and_residuals():
            G = {(a, b): cx.block_gram(a, b) for a in range(2 * M - 1) for b in range(a, 2 * M)}

            def gram(a, b):
                return G[(a, b)] if a <= b else G[(b, a)].conj().T
            G0 = np.zeros((K, K), complex)
            G01 = np.zeros((K, K), complex)
            for j in range(M):
                for jp in range(M):
                    G0[j * m0:(j + 1) * m0, jp * m0:(jp + 1) * m0] = sum(gram(i + j, i + jp) for i in range(M))
                    G01[j * m0:(j + 1) * m0, jp * m0:(jp + 1) * m0] = sum(gram(i + j, i + jp + 1) for i in range(M))
            ev, V = np.linalg.eigh((G0 + G0.conj().T) / 2)
            ev, V = ev[::-1], V[:, ::-1]
            keep = ev > (1e-7 ** 2) * ev[0]
            sv = np.sqrt(ev[keep])
            V = V[:, keep]
            Am = (V.conj().T @ G01 @ V) / sv[:, None] / sv[None, :]               # S^-1 V' Q0' Q1 V S^-1    nlfeast.jl:222-224
            w, v = sla.eig(Am, check_finite=False)
            p = np.lexsort((w.imag, w.real))
            w, v = w[p], v[:, p]
            W = (V / sv[None, :]) @ v                                             # Y = [S_0 .. S_{M-1}] W       nlfeast.jl:226
            nk = w.size
            Lam_all, res_all = np.empty(nk, complex), np.empty(nk)
            for c0 in range(0, nk, m0):
                c1 = min(nk, c0 + m0)
                Wc = np.zeros((K, m0), complex)
                Wc[:, :c1 - c0] = W[:, c0:c1]
                lam_c = np.full(m0, w[c0], complex)
                lam_c[:c1 - c0] = w[c0:c1]
                if c1 - c0 < m0:
                    Wc[:, c1 - c0:] = W[:, [c0]]
                cx.moment_combine(Wc)
                rc, _ = ses.residuals(lam_c)                                      # update_R_moments!, utils.jl:118-123
                Lam_all[c0:c1], res_all[c0:c1] = w[c0:c1], rc[:c1 - c0]
            p = np.argsort(res_all, kind="stable")                                # utils.jl:125-133
            Lam_all, res_all, W = Lam_all[p], res_all[p], W[:, p]
            # X = Y[:, 1:m0] with its residual vectors R (the right-hand side of the next pass)
            Wc = np.zeros((K, m0), complex)
            nb = min(m0, nk)
            Wc[:, :nb] = W[:, :nb]
            lam_c = np.full(m0, Lam_all[0], complex)
            lam_c[:nb] = Lam_all[:nb]
            if nb < m0:
                Wc[:, nb:] = W[:, [0]]
            cx.moment_combine(Wc)
            ses.residuals(lam_c)
            return Lam_all, res_all, W, lam_c

        ses.contour_pass(2 * M)                                                   # nlfeast.jl:195-213
        Lam, res, W, lam_x = reduce_and_residuals()
        hist.append({"nit": 0})
        for nit in range(1, iter + 1):
            ses.contour_pass(2 * M, Lam=lam_x)                                    # nlfeast.jl:255-275
            Lam, res, W, lam_x = reduce_and_residuals()
            nb = min(m0, Lam.size)
            inside = in_contour(Lam[:nb], c, r)
            res_inside = res[:nb][inside]
            hist.append({"nit": nit, "inside": int(inside.sum()),
                         "max_res_inside": float(res_inside.max()) if inside.any() else float("nan")})
            if debug:
                iter_debug_print(nit, Lam[:nb], res[:nb], CircularContour(c, r, None, None), spurious)
            if res_inside.size > 0 and res_inside.max() < eps:                    # nlfeast.jl:297
                break
            good = res_inside[res_inside < spurious]
            if nit > 1 and good.size > 0 and good.max() < eps:                    # nlfeast.jl:300
                break
        X[:, :] = cx.get_X()
        # all moments*m0 vectors, in residual order
        Yall = np.empty((N, Lam.size), np.complex128, order="F")
        for c0 in range(0, Lam.size, m0):
            c1 = min(Lam.size, c0 + m0)
            Wc = np.zeros((K, m0), complex)
            Wc[:, :c1 - c0] = W[:, c0:c1]
            if c1 - c0 < m0:
                Wc[:, c1 - c0:] = W[:, [c0]]
            cx.moment_combine(Wc)
            lam_c = np.full(m0, Lam[c0], complex)
            lam_c[:c1 - c0] = Lam[c0:c1]
            ses.residuals(lam_c)
            Yall[:, c0:c1] = cx.get_X()[:, :c1 - c0]
        if stats is not None:
            stats["history"] = hist
    finally:
        ses.close()
    return Lam, Yall, res


def ifeast(A, X0, nodes, iter, *, c=complex(0.0, 0.0), r=1.0, debug=False, eps=0.05, ctx=None, solver_opts=None, stats=None):
    """ifeast!(A, X0, nodes, iter; c, r, debug, eps)  (src/feast_experimental.jl:1-60): FEAST with INEXACT inner solves,
    exactly `iter` contour passes, all m0 Ritz pairs returned unfiltered with absolute residuals; X0 is not mutated.

    The reference applies (zI - A)^-1 to X column by column with bicgstabl and solves the generalized reduced problem
    (Q'AQ, Q'Q).  On the device the same filter is applied in residual-inverse-iteration form -- (zI - A)^-1 x_j =
    (x_j - (A - zI)^-1 R_j) / (z - l_j) exactly -- by the pseudo-block Krylov solver over the tiled SpMM (one SpMM per
    iteration for all m0 columns), and Q is orthonormalised before the projection (same Ritz pairs).  A must be sparse
    (the Krylov path has no dense operator); the inner tolerance defaults to IterativeSolvers' sqrt(eps).
    """
    N, m0 = X0.shape
    if A.shape[0] != A.shape[1]:
        raise ValueError("Incorrect dimensions of A, must be square")       # feast_experimental.jl:4
    if A.shape[0] != N:
        raise ValueError("Incorrect dimensions of X0, must match A")        # :6
    if not sp.issparse(A):
        raise TypeError("ifeast on the B200 path needs a sparse A (Krylov inner solves); use feast for dense operators")
    opts = {"kind": _lib.SOLVER_KRYLOV, "inner_tol": float(np.sqrt(np.finfo(float).eps)), "max_inner": 4000}
    opts.update(solver_opts or {})
    X = np.array(X0, dtype=np.complex128, order="F")                        # X = deepcopy(X0), :9
    st = {} if stats is None else stats
    contour = circular_contour_trapezoidal(c, r, nodes)                     # theta of :15
    _linear_driver(X, A, None, contour, iter, -1.0, debug, False, ctx, opts, False, st, None)   # eps < 0: never stops early
    return st["Lam_all"], X, st["res_all"]


def nlfeast_it(T, X, nodes, iter, *, c=complex(0.0, 0.0), r=1.0, debug=False, eps=0.05, ctx=None, solver_opts=None,
               stats=None):
    """nlfeast_it!(T, X, nodes, iter; c, r, debug, eps)  (src/nlfeast.jl:87-171): nlfeast with inexact inner solves --
    relative tolerance 1e-3 in the first contour pass (:106), 1e-8 afterwards (:139) -- stopping when
    max(res[inside]) < eps (:164).  `T` is the list of polynomial coefficients (see nlfeast).  The warm start from the
    previous solution (`Tinv`, nodes x N x m0 of storage upstream) is not kept on the device: in residual-inverse-
    iteration form the right-hand side R shrinks with the outer iteration, which plays the same role.
    """
    class _Sched(dict):      # inner tolerance schedule read by nlfeast before every contour pass
        pass
    opts = _Sched({"kind": _lib.SOLVER_KRYLOV, "max_inner": 4000})
    opts.update(solver_opts or {})
    opts.tol_schedule = (1e-3, 1e-8)
    st = {} if stats is None else stats
    lam, X, res = nlfeast(T, X, nodes, iter, c=c, r=r, debug=debug, eps=eps, store=False, ctx=ctx, solver_opts=opts, stats=st,
                          _stop_rule="it")
    return lam, X, res


def contour_estimate_eig(A, contour, B=I, *, samples=None, eps=1e-12, debug=False, mixed_prec=False,
                         factorizer=None, left_divider=None, X=None, seed=0, ctx=None, solver_opts=None, comm=None):
    """contour_estimate_eig(A, contour, B=I; samples=min(100, N), ...)  (src/stochastic.jl:2-33):
    Hutchinson estimate of the number of eigenvalues inside the contour,
    real(sum_k w_k tr(X' (z_k B - A)^-1 X) / samples) with X = randn(ComplexF64, N, samples).
    `X` may be passed explicitly (parity tests ship the probe block as data)."""
    _check_plugins(factorizer, left_divider, mixed_prec)
    N = A.shape[0]
    m0 = min(100, N) if samples is None else int(samples)
    if X is None:
        rng = np.random.default_rng(seed)
        X = (rng.standard_normal((N, m0)) + 1j * rng.standard_normal((N, m0))) / np.sqrt(2.0)
    own_ctx = ctx is None
    if own_ctx:
        ctx = FeastContext()
    try:
        generalized = not (B is None or isinstance(B, str))
        if generalized:
            A, B = _densify_if_mixed(A, B)
        ctx.set_operator(0, A)
        if generalized:
            ctx.set_operator(1, B, n=N)
            ctx.set_problem(_lib.PROBLEM_GENERALIZED, 2, N)
        else:
            ctx.set_problem(_lib.PROBLEM_STANDARD, 1, N)
        if comm is not None:
            comm(ctx)
        ctx.set_contour(contour.nodes, contour.weights)
        if ctx.nranks > 1:
            ctx.set_node_owners(node_owners(contour.nodes, ctx.nranks))
        ctx.set_solver(**(solver_opts or {}))
        ctx.set_subspace(X)
        return ctx.estimate_count()
    finally:
        if own_ctx:
            ctx.close()


def dual_gen_feast(Xr, Xl, A, B, contour: Contour | None = None, *, nodes=8, iter=10, c=complex(0.0, 0.0), r=1.0,
                   debug=False, store=False, eps=1e-12, factorizer=None, left_divider=None,
                   ctx=None, solver_opts=None, stats=None, comm=None):
    """dual_gen_feast!(Xr, Xl, A, B[, contour]; ...)  (src/feast.jl:158-257): two-sided FEAST for
    non-Hermitian pencils; returns (L[in], Xr[:, in], Xl[:, in], resr[in]).  `B` may be `I`.

    Restates the INTENDED semantics where upstream has defects (both documented in the oracle):
    `Diagonal(1.0/S.S)` (feast.jl:200-201) is taken as the elementwise Sigma^-1/2 on both sides (so that
    Ql' B Qr = I, the intent stated at feast.jl:205), and the left residual uses conj(lambda) (feast.jl:214).  The m0 x m0 SVD and the two reduced eigenproblems run on host
    LAPACK; everything n-sized runs in libfeast_cuda.so with ONE LU per node serving A - zB and its
    adjoint."""
    _check_plugins(factorizer, left_divider, False)
    if contour is None:
        contour = circular_contour_trapezoidal(c, r, nodes)  # feast.jl:163
        store = False                                         # the wrapper drops `store` (feast.jl:164)
    N, m0 = Xl.shape
    if A.shape[0] != A.shape[1]:
        raise ValueError("Incorrect dimensions of A, must be square")
    if A.shape[0] != N:
        raise ValueError("Incorrect dimensions of X, must match A")
    own_ctx = ctx is None
    if own_ctx:
        ctx = FeastContext()
    try:
        if B is None or isinstance(B, str):
            Bop = None
        else:
            A, Bop = _densify_if_mixed(A, B)
        ctx.set_operator(0, A)
        ctx.set_operator(1, Bop, n=N)
        ctx.set_problem(_lib.PROBLEM_GENERALIZED, 2, N)
        if comm is not None:
            comm(ctx)
        ctx.set_contour(contour.nodes, contour.weights)
        if ctx.nranks > 1:
            ctx.set_node_owners(node_owners(contour.nodes, ctx.nranks))
        ctx.set_solver(store=store, **(solver_opts or {}))
        ctx.dual_set_subspace(Xr, Xl)
        Lam = np.zeros(m0, complex)
        resr = np.zeros(m0)
        hist = []
        for nit in range(iter + 1):
            G = ctx.dual_project()                                   # Ql' B Qr                feast.jl:199
            U, S, Vh = sla.svd(G, check_finite=False)                # host m0 x m0 SVD
            sc = 1.0 / np.sqrt(np.maximum(S, S[0] * 1e-280))         # Sigma^-1/2 on both sides: Ql' B Qr = I
            Aq, Bq = ctx.dual_rotate(Vh.conj().T * sc[None, :], U * sc[None, :])   # feast.jl:200-205
            Lam, Xq = _eig_sorted(Aq, Bq)                            # feast.jl:206-208
            wl, vl = sla.eig(Aq.conj().T, Bq.conj().T, check_finite=False)   # feast.jl:210-211
            order = [int(np.argmin(np.abs(wl - np.conj(l)))) for l in Lam]   # pair left vectors with Lam
            resr = ctx.dual_recover_residual(Xq, vl[:, order], Lam)  # feast.jl:209-215
            inside = in_contour(Lam, contour)
            if debug:
                iter_debug_print(nit, Lam, resr, contour, 1e-5)
            rec = {"nit": nit, "inside": int(inside.sum()),
                   "max_res_inside": float(resr[inside].max()) if inside.any() else float("nan")}
            if inside.any() and resr[inside].max() < eps:            # feast.jl:218
                hist.append(rec)
                break
            if nit < iter:
                rec.update(ctx.dual_contour_apply(Lam))              # feast.jl:222-248
            hist.append(rec)
        xr, xl = ctx.dual_get()
        Xr[:, :] = xr
        Xl[:, :] = xl
        if stats is not None:
            stats["history"] = hist
            stats["launches"] = ctx.launch_count()
    finally:
        if own_ctx:
            ctx.close()
    inside = in_contour(Lam, contour)
    if not inside.any():
        print("no eigenvalues found in contour!")
    return Lam[inside], Xr[:, inside], Xl[:, inside], resr[inside]
