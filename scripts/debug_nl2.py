import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, scipy.linalg as sla
import feastsolver_jl_b200 as fs
from conftest import load_golden, csc_unpack, x0
nep = load_golden("nep_fixtures.npz"); g = load_golden("nlfeast_golden.npz")

def run(coeffs, X, nodes, iters, c, r, label):
    dense = [a.toarray() if hasattr(a, "toarray") else np.asarray(a, dtype=complex) for a in coeffs]
    T = lambda z: sum(dense[i] * z**i for i in range(len(dense)))
    N, m0 = X.shape
    ctx = fs.FeastContext()
    for i, a in enumerate(coeffs): ctx.set_operator(i, a, n=N)
    ctx.set_problem(2, len(coeffs), N)
    ct = fs.circular_contour_trapezoidal(c, r, nodes)
    ctx.set_contour(ct.nodes, ct.weights); ctx.set_solver(store=True); ctx.set_subspace(X)
    ctx.orthonormalize_X()
    Lam = None
    for nit in range(iters):
        Xc = ctx.get_X()
        Rc = ctx.get_R() if nit > 0 else None
        st = ctx.contour_apply(Lam, first_pass=(nit == 0))
        Q0 = ctx.get_Q()
        if nit == 0:
            ref = sum(w * np.linalg.solve(T(z), Xc) for z, w in zip(ct.nodes, ct.weights))
        else:
            ref = sum((Xc - np.linalg.solve(T(z), Rc)) * (w / (z - Lam))[None, :] for z, w in zip(ct.nodes, ct.weights))
        e_q0 = np.abs(Q0 - ref).max() / np.abs(ref).max()
        Rf, G1 = ctx.beyn_reduce()
        U = ctx.get_Q()
        e_orth = np.abs(U.conj().T @ U - np.eye(m0)).max()
        e_fact = np.abs(U @ Rf - Q0).max() / np.abs(Q0).max()
        Us, S, Vh = sla.svd(Rf)
        Am = (Us.conj().T @ G1) @ Vh.conj().T * (1.0 / S)[None, :]
        w_, v_ = sla.eig(Am); p = np.lexsort((w_.imag, w_.real)); Lam = w_[p]; Xq = Us @ v_[:, p]
        res = ctx.recover_residual(Xq, Lam)
        Xn = ctx.get_X(); Rn = ctx.get_R()
        Xref = U @ Xq; Xref /= np.linalg.norm(Xref, axis=0)
        Rref = np.stack([T(Lam[j]) @ Xref[:, j] for j in range(m0)], axis=1)
        rref = np.array([np.linalg.norm(Rref[:, j]) / np.linalg.norm(T(Lam[j])) for j in range(m0)])
        ins = np.abs(Lam - c) <= r
        print(f"{label} nit={nit} Q0err={e_q0:.1e} orth={e_orth:.1e} fact={e_fact:.1e} Xerr={np.abs(Xn-Xref).max():.1e} "
              f"Rerr={np.abs(Rn-Rref).max():.1e} reserr={np.abs(res-rref).max():.1e} inside={ins.sum()} maxres_in={res[ins].max() if ins.any() else None} S[-1]/S[0]={S[-1]/S[0]:.1e}", flush=True)
    ctx.close()

bf_d = [csc_unpack(nep, f"butterfly{i}").toarray() for i in range(5)]
bf_s = [csc_unpack(nep, f"butterfly{i}") for i in range(5)]
run(bf_d, g["butterfly_X0"].copy(), 16, 6, 1+1j, 0.5, "bf-dense")
run(bf_s, g["butterfly_X0"].copy(), 16, 6, 1+1j, 0.5, "bf-sparse")
n = 100
A = np.diag(np.full(n, 2.0)) + np.diag(np.full(n - 1, -1.0), 1) + np.diag(np.full(n - 1, -1.0), -1)
run([-A, np.eye(n)], x0(n, 10, 5), 8, 5, 0.02, 0.02, "linpencil")
s5 = [csc_unpack(nep, f"system5_{i}") for i in range(3)]
run(s5, x0(1000, 80, 4), 32, 4, -1.55, 0.05, "system5")
