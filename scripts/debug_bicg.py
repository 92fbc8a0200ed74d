import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, scipy.sparse as sp
import feastsolver_jl_b200 as fs
from feastsolver_jl_b200 import _lib
from conftest import x0
n = 400
d = np.linspace(1.0, 40.0, n)
A = sp.diags([d, 0.3 * np.ones(n - 1), -0.2 * np.ones(n - 1)], [0, 1, -1], format="csc")
B = sp.diags([np.full(n, 2.0), 0.1 * np.ones(n - 1), 0.1 * np.ones(n - 1)], [0, 1, -1], format="csc")
ct = fs.circular_contour_trapezoidal(3.0, 0.3, 16)
for kry in (_lib.KRYLOV_GMRES, _lib.KRYLOV_BICGSTAB):
    ctx = fs.FeastContext()
    ctx.set_operator(0, A); ctx.set_operator(1, B); ctx.set_problem(1, 2, n)
    ctx.set_solver(kind=_lib.SOLVER_KRYLOV, krylov=kry, inner_tol=1e-10, max_inner=2000)
    Bm = x0(n, 24, 9)
    for z in ct.nodes[:4]:
        F = ctx.factorize([1.0, -z])
        Y = ctx.solve(F, Bm)
        Z = (A - z * B).toarray()
        ref = np.linalg.solve(Z, Bm)
        print("z", z, "relerr", np.abs(Y - ref).max() / np.abs(ref).max(), "resid", np.abs(Z @ Y - Bm).max(), flush=True)
        ctx.factor_free(F)
    ctx.close()
