"""Kernel-level bench of the CSR SpMM at the C2 shape (n = grid^3, m0 columns): real-valued
operator (RR projection A*Q) and the complex assembled shifted operator (inside COCG)."""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import feastsolver_jl_b200 as fs
from feastsolver_jl_b200 import _lib, workloads as wl

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=100)
ap.add_argument("--m0", type=int, default=64)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--krylov-iters", type=int, default=24)
a = ap.parse_args()
A, B = wl.laplacian3d_pencil(a.grid)
n = a.grid ** 3
X0 = wl.rand_subspace(n, a.m0, seed=0)
ctx = fs.FeastContext()
ctx.set_operator(0, A); ctx.set_operator(1, B); ctx.set_problem(1, 2, n)
ctx.set_subspace(X0)
print(json.dumps({"layout": ctx.layout_info()}))
_, ms = ctx.apply_operator(0, which=0, download=False, reps=a.reps)
bytes_real = A.nnz * 12 + 4 * (n + 1) + 32 * n * a.m0
print(json.dumps({"kernel": "spmm real", "ms": ms, "GBs": bytes_real / ms / 1e6}))
# complex shifted operator inside COCG: a short capped solve
c, r, cnt = wl.c2_slice(a.grid, target=36)
ct = fs.circular_contour_gauss(c, r, 16)
ctx.set_contour(ct.nodes[:1], ct.weights[:1])
ctx.set_solver(kind=_lib.SOLVER_KRYLOV, inner_tol=1e-6, max_inner=a.krylov_iters)
Aq, Bq = ctx.project(True)
from feastsolver_jl_b200.feast import _eig_sorted
Lam, Xq = _eig_sorted(Aq, Bq)
ctx.recover_residual(Xq, Lam)
st = ctx.contour_apply(Lam)
bytes_c = A.nnz * 20 + 4 * (n + 1) + 32 * n * a.m0
msc = st["t_spmm_ms"] / max(1, st["spmm_launches"])
print(json.dumps({"kernel": "spmm complex (COCG)", "ms": msc, "GBs": bytes_c / msc / 1e6, "solve_ms_per_iter": st["t_solve_ms"] / st["inner_iters_total"]}))
ctx.close()
