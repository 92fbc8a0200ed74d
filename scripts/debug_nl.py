import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, scipy.linalg as sla
import feastsolver_jl_b200 as fs
from feastsolver_jl_b200 import _lib
from conftest import load_golden, csc_unpack
nep = load_golden("nep_fixtures.npz"); g = load_golden("nlfeast_golden.npz")
coeffs = [csc_unpack(nep, f"butterfly{i}").toarray() for i in range(5)]
X = g["butterfly_X0"].copy()
N, m0 = X.shape
ctx = fs.FeastContext()
for i, a in enumerate(coeffs): ctx.set_operator(i, a, n=N)
ctx.set_problem(2, 5, N)
ct = fs.circular_contour_trapezoidal(1+1j, 0.5, 16)
ctx.set_contour(ct.nodes, ct.weights); ctx.set_solver(store=True); ctx.set_subspace(X)
ctx.orthonormalize_X()
Xo = ctx.get_X(); print("orth X dev", np.abs(Xo.conj().T@Xo - np.eye(m0)).max())
st = ctx.contour_apply(None, first_pass=True); print(st)
Q0 = ctx.get_Q(); print("Q0 finite", np.isfinite(Q0).all(), np.abs(Q0).max())
T = lambda z: sum(coeffs[i]*z**i for i in range(5))
ref = sum(w*np.linalg.solve(T(z), Xo) for z, w in zip(ct.nodes, ct.weights))
print("Q0 err", np.abs(Q0-ref).max()/np.abs(ref).max())
Rf, G1 = ctx.beyn_reduce(); print("Rf finite", np.isfinite(Rf).all(), "G1 finite", np.isfinite(G1).all())
U = ctx.get_Q(); print("U orth dev", np.abs(U.conj().T@U-np.eye(m0)).max(), "Q0=U Rf err", np.abs(U@Rf-Q0).max())
