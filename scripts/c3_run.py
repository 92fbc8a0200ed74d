"""Config C3 (BASELINE.json): dense non-Hermitian complex n x n standard problem, circular contour,
m0 columns, `nodes` trapezoid nodes, feast! with store=true (one LU per node, reused every iteration).

    python scripts/c3_run.py --n 16384 --m0 128 --nodes 32 --r 7.0
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import feastsolver_jl_b200 as fs
from feastsolver_jl_b200 import workloads as wl

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--m0", type=int, default=128)
ap.add_argument("--nodes", type=int, default=32)
ap.add_argument("--r", type=float, default=7.0)
ap.add_argument("--iter", type=int, default=10)
ap.add_argument("--check", action="store_true", help="compare with numpy eigvals (small n only)")
a = ap.parse_args()
t0 = time.perf_counter()
A = wl.dense_nonhermitian(a.n, seed=1551)
X0 = wl.rand_subspace(a.n, a.m0, seed=0)
t_gen = time.perf_counter() - t0
st = {}
t0 = time.perf_counter()
e, v, res = fs.feast(X0, A, nodes=a.nodes, iter=a.iter, c=0.0, r=a.r, eps=1e-12, store=True, stats=st)
tts = time.perf_counter() - t0
hist = st["history"]
solves = sum(h.get("nodes_local", 0) for h in hist)
t_factor = sum(h.get("t_factor_ms", 0.0) for h in hist)
t_solve = sum(h.get("t_solve_ms", 0.0) for h in hist)
out = {"config": f"C3 dense non-Hermitian n={a.n} m0={a.m0} nodes={a.nodes} r={a.r} store=true", "found": int(e.size),
       "max_res": float(res.max()) if res.size else None, "outer_iterations": len(hist), "tts_s": tts,
       "node_solves": solves, "node_solves_per_s": solves / tts, "factor_ms_total": t_factor, "solve_ms_total": t_solve,
       "lu_tflops": (8 / 3) * a.n ** 3 * a.nodes / (t_factor * 1e-3) / 1e12 if t_factor else None,
       "getrs_tflops": 8 * a.n ** 2 * a.m0 * solves / (t_solve * 1e-3) / 1e12 if t_solve else None,
       "phase_ms": {k: float(x) for k, x in st["phase_ms"].items()}, "gen_s": t_gen}
if a.check:
    ex = np.linalg.eigvals(A)
    ex = ex[np.abs(ex) <= a.r]
    good = e[res < 1e-8]
    out["exact_inside"] = int(ex.size)
    out["max_eig_err"] = float(max(np.abs(ex - l).min() for l in good)) if good.size else None
print(json.dumps(out))
