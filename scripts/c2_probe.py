"""Development probe: run the C2 workload (3-D Laplacian + mass pencil) at grid m with
given inner tolerance and print per-outer-iteration statistics.

    python scripts/c2_probe.py --m 100 --m0 64 --nodes 16 --tol 1e-6
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feastsolver_jl_b200 as fs  # noqa: E402
from feastsolver_jl_b200 import _lib, workloads as wl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=100)
    ap.add_argument("--m0", type=int, default=64)
    ap.add_argument("--nodes", type=int, default=16)
    ap.add_argument("--target", type=int, default=36)
    ap.add_argument("--tol", type=float, default=1e-6)
    ap.add_argument("--maxit", type=int, default=4000)
    ap.add_argument("--iter", type=int, default=10)
    ap.add_argument("--eps", type=float, default=1e-12)
    a = ap.parse_args()
    t0 = time.time()
    A, B = wl.laplacian3d_pencil(a.m)
    n = a.m ** 3
    c, r, cnt = wl.c2_slice(a.m, target=a.target)
    X0 = wl.rand_subspace(n, a.m0, seed=0)
    print(f"n={n} nnz={A.nnz} slice c={c:.6f} r={r:.6f} inside={cnt} setup {time.time()-t0:.1f}s", flush=True)
    ct = fs.circular_contour_gauss(c, r, a.nodes)
    st = {}
    t0 = time.time()
    e, v, res = fs.gen_feast(X0, A, B, ct, eps=a.eps, iter=a.iter, stats=st,
                             solver_opts={"kind": _lib.SOLVER_KRYLOV, "inner_tol": a.tol, "max_inner": a.maxit})
    tts = time.time() - t0
    for h in st["history"]:
        print(json.dumps({k: (round(v, 6) if isinstance(v, float) else v) for k, v in h.items()}), flush=True)
    exact = wl.laplacian3d_spectrum(a.m, count=cnt + 8)
    exact = exact[np.abs(exact - c) <= r]
    err = np.abs(np.sort(e.real) - exact).max() / np.abs(exact).max() if e.size == exact.size else float("nan")
    print(f"found {e.size} (exact {exact.size}) max res {res.max() if res.size else None} rel eig err {err:.2e} "
          f"tts {tts:.2f}s launches {st['launches']} phases {st['phase_ms']}")


if __name__ == "__main__":
    main()
