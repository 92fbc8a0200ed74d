"""Generate tests/golden/*.npz from the reference's shipped fixtures.

Run ONCE in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Nothing under tests/ or bench.py reads /root/reference at run time; the GPU box
only sees the committed .npz files.  What is stored:

* the coefficient matrices of data/butterflyM0-M4.mtx, data/quadraticM0-M1.mtx,
  data/system5A0-A2.mtx as CSC triplets (fixtures, small);
* exact answers from the reference's own `companion()` construction
  (src/companion.jl:1-28, restated in oracle/feast_oracle.py) inside the contours
  used by test/butterfly.jl:67-72, test/polynomial.jl:13-20, test/deficient.jl:209-219;
* seeded initial subspaces X0 and the oracle's converged eigenvalues for the
  reference's asserted cases test/runtests.jl:16-23,33-49 (T1, T2, T3a-c), so that
  GPU parity tests can be checked against stored vectors as well as a live oracle.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import feast_oracle as fo  # noqa: E402

REF = "/root/reference/data"


def csc_pack(prefix, M, out):
    M = sp.csc_matrix(M)
    M.sort_indices()
    out[prefix + "_indptr"] = M.indptr.astype(np.int64)
    out[prefix + "_indices"] = M.indices.astype(np.int64)
    out[prefix + "_data"] = M.data
    out[prefix + "_shape"] = np.array(M.shape, np.int64)


def x0(n, m0, seed):
    rng = np.random.default_rng(seed)
    return rng.random((n, m0)) + 1j * rng.random((n, m0))  # rand(ComplexF64, n, m0)


def main():
    # ---------------- polynomial NEP fixtures + companion answers ----------
    out = {}
    bf = [fo.mmread(f"{REF}/butterflyM{i}.mtx") for i in range(5)]
    for i, M in enumerate(bf):
        csc_pack(f"butterfly{i}", M, out)
    w, _, res = fo.companion(bf)
    inside = np.abs(w - (1 + 1j)) <= 0.5
    out["butterfly_companion_inside"] = w[inside]
    out["butterfly_companion_res"] = res[inside]
    print("butterfly: total", w.size, "inside", inside.sum(), "max res", res[inside].max())

    qd = [fo.mmread(f"{REF}/quadraticM{i}.mtx") for i in range(2)]
    for i, M in enumerate(qd):
        csc_pack(f"quadratic{i}", M, out)
    A0, A1 = qd  # test/deficient.jl: T(z) = (z+0.2)(z-0.1) A1 + A0 = z^2 A1 + 0.1 z A1 + (A0 - 0.02 A1)
    coeffs = [A0 - 0.02 * A1, 0.1 * A1, A1]
    w, _, res = fo.companion(coeffs)
    fin = np.isfinite(w)
    inside = fin & (np.abs(w) <= 0.25)
    out["quadratic_companion_inside"] = w[inside]
    print("deficient quadratic inside:", np.sort_complex(w[inside]))

    s5 = [fo.mmread(f"{REF}/system5A{i}.mtx") for i in range(3)]
    for i, M in enumerate(s5):
        csc_pack(f"system5_{i}", M, out)
    w, _, res = fo.companion(s5)
    inside = np.isfinite(w) & (np.abs(w + 1.55) <= 0.05)
    out["system5_companion_inside"] = w[inside]
    print("system5 inside:", inside.sum(), "max res", res[inside].max())
    np.savez_compressed(os.path.join(HERE, "nep_fixtures.npz"), **out)

    # ---------------- linear known answers (runtests.jl) --------------------
    lin = {}
    A = np.diag(np.arange(1.0, 26.0))
    X = x0(25, 5, 101)
    lin["T1_X0"] = X.copy()
    e, v, r = fo.feast(X, A, nodes=8, iter=10, c=1.5, r=2.0)
    lin["T1_e"], lin["T1_res"] = e, r
    X = x0(25, 5, 102)
    lin["T2_X0"] = X.copy()
    e, v, r = fo.gen_feast(X, A, np.eye(25), nodes=8, iter=100, c=1.5, r=2.0)
    lin["T2_e"], lin["T2_res"] = e, r
    L = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(100, 100), format="csc")
    contours = {
        "T3a": fo.circular_contour_trapezoidal(0.05 + 0j, 0.05, 8),
        "T3b": fo.rectangular_contour_trapezoidal(0.0 - 0.05j, 0.1 + 0.05j, 8),
        "T3c": fo.rectangular_contour_gauss(0.0 - 0.05j, 0.1 + 0.05j, 8),
        "T3d": fo.circular_contour_gauss(0.05 + 0j, 0.05, 8),
    }
    for k, (name, ct) in enumerate(contours.items()):
        X = x0(100, 20, 200 + k)
        lin[name + "_X0"] = X.copy()
        lin[name + "_nodes"], lin[name + "_weights"] = ct.nodes, ct.weights
        e, v, r = fo.feast(X, L, ct, eps=10e-15)
        lin[name + "_e"], lin[name + "_res"] = e, r
        print(name, len(e), r.max())
    lin["lap1d_exact"] = 2 - 2 * np.cos(np.arange(1, 11) * np.pi / 101)
    np.savez_compressed(os.path.join(HERE, "linear_golden.npz"), **lin)

    # ---------------- nlfeast on the shipped butterfly ----------------------
    nl = {}
    T = fo.polynomial(bf)
    X = x0(64, 20, 300)
    nl["butterfly_X0"] = X.copy()
    lam, Xo, res = fo.nlfeast(T, X, 16, 30, c=1 + 1j, r=0.5, eps=1e-13)
    ins = np.abs(lam - (1 + 1j)) <= 0.5
    nl["butterfly_nlfeast_lam"], nl["butterfly_nlfeast_res"] = lam, res
    print("nlfeast butterfly inside", ins.sum(), "max res inside", res[ins].max())
    np.savez_compressed(os.path.join(HERE, "nlfeast_golden.npz"), **nl)


if __name__ == "__main__":
    main()
