"""CPU oracle (oracle/feast_oracle.py nlfeast, SuperLU factorizations) on the C4 workload at a given block size:

    python tests/golden/make_c4_full_oracle.py 500 0.006 64 6 tests/golden/c4_full_oracle.npz      (~5 min, 16 cores)

writes the m0 Ritz values / residuals and the in-contour mask.  Test infrastructure only."""
import os, sys, time, json, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import feast_oracle as fo
from feastsolver_jl_b200 import workloads as wl
mb, r, m0, nodes, iters = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), 24, int(sys.argv[4])
coeffs = wl.butterfly_coeffs(mb)
n = mb*mb
def T(z):
    M = coeffs[0].astype(complex)
    p = 1.0+0j
    out = coeffs[0].astype(complex)
    zp = z
    for A in coeffs[1:]:
        out = out + zp*A
        zp *= z
    return out.tocsc()
X0 = wl.rand_subspace(n, m0, seed=0)
hist = []
t0 = time.time()
lam, X, res = fo.nlfeast(T, X0.copy(), nodes, iters, c=1+1j, r=r, eps=1e-10, history=hist)
inside = np.abs(lam-(1+1j)) <= r
if len(sys.argv) > 5: np.savez(sys.argv[5], lam=lam, res=res, inside=inside, mb=mb, r=r, m0=m0, nodes=nodes, iters=len(hist), secs=time.time()-t0)
print(json.dumps({"mb": mb, "r": r, "inside": int(inside.sum()), "res_inside_max": float(res[inside].max()) if inside.any() else None,
   "res_inside_min": float(res[inside].min()) if inside.any() else None, "secs": time.time()-t0, "hist": [h.get("max_res_inside") if isinstance(h, dict) else None for h in hist]}))
