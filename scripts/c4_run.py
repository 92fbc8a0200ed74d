"""Config C4 (BASELINE.json): nlfeast! on the quartic butterfly-structured polynomial NEP scaled to
mb x mb one-dimensional blocks (n = mb^2, sparse 5-point patterns), circle + trapezoid nodes.

    python scripts/c4_run.py --mb 100 --m0 48 --nodes 24 --r 0.05
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import feastsolver_jl_b200 as fs
from feastsolver_jl_b200 import _lib, workloads as wl

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=100)
ap.add_argument("--m0", type=int, default=48)
ap.add_argument("--nodes", type=int, default=24)
ap.add_argument("--r", type=float, default=0.05)
ap.add_argument("--iter", type=int, default=12)
ap.add_argument("--tol", type=float, default=1e-8)
ap.add_argument("--maxit", type=int, default=2000)
ap.add_argument("--solver", default="auto", choices=["auto", "dense", "krylov"])
ap.add_argument("--eps", type=float, default=10e-12, help="outer tolerance (nlfeast! default 10e-12)")
ap.add_argument("--golden", default=None, help="oracle golden (.npz from tests/golden/make_c4_full_oracle.py) to compare with")
ap.add_argument("--no-store", action="store_true", help="refactor at every outer iteration (store=false)")
a = ap.parse_args()
coeffs = wl.butterfly_coeffs(a.mb)
n = a.mb ** 2
X0 = wl.rand_subspace(n, a.m0, seed=0)
kind = {"auto": _lib.SOLVER_AUTO, "dense": _lib.SOLVER_DENSE_LU, "krylov": _lib.SOLVER_KRYLOV}[a.solver]
st = {}
t0 = time.perf_counter()
lam, X, res = fs.nlfeast(coeffs, X0, a.nodes, a.iter, c=1 + 1j, r=a.r, eps=a.eps, stats=st, store=not a.no_store,
                         solver_opts={"kind": kind, "inner_tol": a.tol, "max_inner": a.maxit})
tts = time.perf_counter() - t0
inside = np.abs(lam - (1 + 1j)) <= a.r
good = inside & (res < 1e-8)
hist = st["history"]
cmp = {}
if a.golden:
    g = np.load(a.golden)
    gl, gr = g["lam"], g["res"]
    gin = np.abs(gl - (1 + 1j)) <= a.r
    cmp = {"oracle_inside": int(gin.sum()), "oracle_max_res_inside": float(gr[gin].max()),
           # distance of every converged device eigenvalue to the nearest oracle eigenvalue inside (and vice versa)
           "max_dist_gpu_to_oracle": float(max(np.abs(gl[gin] - l).min() for l in lam[good])) if good.any() else None,
           "max_dist_oracle_to_gpu": float(max(np.abs(lam[good] - l).min() for l in gl[gin])) if good.any() else None}
print(json.dumps({"config": f"C4 butterfly quartic n={n} nnz={coeffs[0].nnz} m0={a.m0} nodes={a.nodes} r={a.r} solver={a.solver}",
                  "inside": int(inside.sum()), "converged_inside": int(good.sum()),
                  "max_res_converged": float(res[good].max()) if good.any() else None, "outer_iterations": len(hist),
                  "tts_s": tts, "inner_iters": [h.get("inner_iters_total") for h in hist],
                  "inner_relres_max": [h.get("inner_relres_max") for h in hist],
                  "res_history": [h.get("max_res_inside") for h in hist], "vs_oracle": cmp,
                  "lam_inside": [[float(l.real), float(l.imag)] for l in lam[inside]]}))
