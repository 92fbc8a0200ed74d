"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol
include/feast_cuda.h declares, its host-only entries (contour constructors) match
the oracle, and compute entries fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import feast_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from feastsolver_jl_b200 import _lib, build
    build.build()
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    from feastsolver_jl_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "feast_cuda.h")).read()
    declared = set(re.findall(r"FEAST_API[^;]*?\b(feast_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.feast_version() == 100


@pytest.mark.parametrize("N", [4, 8, 16, 32, 64])
def test_contour_constructors_match_oracle(lib, N):
    import feastsolver_jl_b200 as fs
    pairs = [
        (fs.circular_contour_trapezoidal(0.3 + 0.1j, 2.0, N), fo.circular_contour_trapezoidal(0.3 + 0.1j, 2.0, N)),
        (fs.circular_contour_gauss(-1.0, 0.5, N), fo.circular_contour_gauss(-1.0, 0.5, N)),
        (fs.rectangular_contour_gauss(-1 - 2j, 2 + 1j, N), fo.rectangular_contour_gauss(-1 - 2j, 2 + 1j, N)),
        (fs.rectangular_contour_trapezoidal(-1 - 2j, 2 + 1j, N), fo.rectangular_contour_trapezoidal(-1 - 2j, 2 + 1j, N)),
    ]
    for a, b in pairs:
        assert np.abs(a.nodes - b.nodes).max() <= 4e-15
        assert np.abs(a.weights - b.weights).max() <= 1e-15


def test_contour_errors(lib):
    import feastsolver_jl_b200 as fs
    with pytest.raises(ValueError, match="multiple of 2"):
        fs.circular_contour_gauss(0, 1, 7)
    with pytest.raises(ValueError, match="multiple of 4"):
        fs.rectangular_contour_gauss(-1 - 1j, 1 + 1j, 6)
    with pytest.raises(ValueError, match="multiple of 4"):
        fs.rectangular_contour_trapezoidal(-1 - 1j, 1 + 1j, 10)
    with pytest.raises(ValueError, match="Invalid corners"):
        fs.rectangular_contour_trapezoidal(1 + 1j, -1 - 1j, 8)
    ct = fs.circular_contour_trapezoidal(0.0, 1.0, 8)
    assert fs.in_contour(np.array([1.0 + 0j]), ct)[0]
    rc = fs.rectangular_contour_trapezoidal(-1 - 1j, 1 + 1j, 8)
    assert not fs.in_contour(np.array([1.0 + 0j]), rc)[0]
    assert abs(fs.rational_func(0.2, fs.circular_contour_trapezoidal(0.0, 1.0, 16)) - 1) < 1e-6
    assert len(ct) == 1  # length(::Contour) = 1


def test_gauss_legendre_matches_numpy(lib):
    for n in (1, 2, 3, 8, 16, 33):
        x = np.empty(n)
        w = np.empty(n)
        assert lib.feast_gauss_legendre(n, x.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p)) == 0
        xr, wr = np.polynomial.legendre.leggauss(n)
        assert np.abs(x - xr).max() < 2e-16 * 8 and np.abs(w - wr).max() < 1e-15


def test_no_cpu_fallback(lib):
    """Without a CUDA device the compute path must fail loudly, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import feastsolver_jl_b200 as fs
    with pytest.raises(fs.FeastError, match="no CPU fallback"):
        fs.FeastContext()
    with pytest.raises(fs.FeastError):
        fs.feast(np.ones((4, 2), complex), np.eye(4))


def test_bad_arguments_return_negative_codes(lib):
    assert lib.feast_ctx_create(None, 0) == -1
    assert lib.feast_set_contour(None, 4, None, None) == -1
    z = np.zeros(8, complex)
    from feastsolver_jl_b200._lib import cplx
    assert lib.feast_contour_circular_gauss(cplx(0), 1.0, 7, z.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p)) == -3
    assert lib.feast_contour_rectangular_gauss(cplx(1 + 1j), cplx(0), 8, z.ctypes.data_as(C.c_void_p),
                                               z.ctypes.data_as(C.c_void_p)) == -1
    assert b"invalid" in lib.feast_last_error(None)


def test_node_owner_partition():
    import feastsolver_jl_b200 as fs
    ct = fo.circular_contour_gauss(1.0, 0.5, 16)
    for nr in (1, 2, 4, 8):
        own = fs.node_owners(ct.nodes, nr)
        counts = np.bincount(own, minlength=nr)
        assert counts.max() - counts.min() == 0
        from feastsolver_jl_b200.partition import node_cost
        cost = node_cost(ct.nodes)
        loads = np.array([cost[own == r].sum() for r in range(nr)])
        assert loads.max() / loads.mean() < 1.35  # near/far nodes are paired


def test_workload_generators():
    from feastsolver_jl_b200 import workloads as wl
    A, B = wl.laplacian3d_pencil(6)
    assert A.shape == (216, 216) and A.nnz == 7 * 216 - 6 * 36
    import scipy.linalg as sla
    ex = np.sort(sla.eigh(A.toarray(), B.toarray(), eigvals_only=True))
    an = wl.laplacian3d_spectrum(6)
    assert np.abs(ex - an).max() < 1e-11
    c, r, cnt = wl.c2_slice(6, target=10)
    assert np.sum(np.abs(an - c) <= r) == cnt


def test_tile_plan_is_a_permutation_and_cuts_the_halo(lib):
    """Host side of the tiled SpMM (csrc/reorder.cpp): the greedy graph-growing order is a permutation, respects
    the shared-memory capacities, and references far fewer out-of-tile rows than the natural order of a 3-D grid."""
    import scipy.sparse as sp
    from feastsolver_jl_b200 import _lib, workloads as wl
    A, _ = wl.laplacian3d_pencil(24)
    A = sp.csr_matrix(A)
    n = A.shape[0]
    rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
    col = np.ascontiguousarray(A.indices, dtype=np.int32)
    res = {}
    for reorder in (0, 1):
        order = np.zeros(n, dtype=np.int32)
        nt, halo = C.c_int(0), C.c_double(0.0)
        rc = lib.feast_debug_tile_plan(n, _lib.ptr(rowptr), _lib.ptr(col), reorder, 192, 704, 160, 4096 if reorder else 0, _lib.ptr(order),
                                       C.byref(nt), C.byref(halo))
        assert rc == 0
        assert np.array_equal(np.sort(order), np.arange(n))
        res[reorder] = (nt.value, halo.value, order)
    assert np.array_equal(res[0][2], np.arange(n))          # natural order kept when not renumbering
    assert res[1][1] < 0.45 * res[0][1]                       # halo rows per row: ~1.2 vs ~4
    assert res[1][1] < 1.6
    # capacity: a row with more distinct columns than the shared-memory row capacity cannot be tiled
    D = sp.csr_matrix(np.ones((1, 300)))
    Dfull = sp.vstack([D, sp.csr_matrix((299, 300))]).tocsr()
    rc = lib.feast_debug_tile_plan(300, _lib.ptr(np.ascontiguousarray(Dfull.indptr, dtype=np.int64)),
                                   _lib.ptr(np.ascontiguousarray(Dfull.indices, dtype=np.int32)), 1, 192, 704, 160, 0, None, None, None)
    assert rc == 1


@pytest.mark.parametrize("kind", ["lap2d_nonsym", "random_sparse", "dense_band"])
def test_tile_plan_self_check_on_irregular_patterns(lib, kind):
    """The plan builder's self-check (capacities respected, every 16-bit tile-local column maps back to the permuted
    global column) on patterns that are not grid-like: non-symmetric, random, and a wide band."""
    import ctypes as C
    import scipy.sparse as sp
    from feastsolver_jl_b200 import _lib
    rng = np.random.default_rng(3)
    if kind == "lap2d_nonsym":
        m = 40
        K = sp.diags([-1.0, 2.0, -0.5], [-1, 0, 1], shape=(m, m))
        A = (sp.kron(K, sp.identity(m)) + sp.kron(sp.identity(m), K)).tocsr()
    elif kind == "random_sparse":
        A = (sp.random(3000, 3000, density=0.002, random_state=5) + sp.identity(3000)).tocsr()
    else:
        n = 1500
        A = sp.diags([rng.standard_normal(n - abs(o)) for o in range(-20, 21)], list(range(-20, 21))).tocsr()
    A.sort_indices()
    n = A.shape[0]
    rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
    col = np.ascontiguousarray(A.indices, dtype=np.int32)
    for reorder in (0, 1):
        for caps in ((192, 768, 160, 0), (96, 384, 80, 512)):
            order = np.zeros(n, dtype=np.int32)
            nt, halo = C.c_int(0), C.c_double(0.0)
            rc = lib.feast_debug_tile_plan(n, _lib.ptr(rowptr), _lib.ptr(col), reorder, caps[0], caps[1], caps[2], caps[3],
                                           _lib.ptr(order), C.byref(nt), C.byref(halo))
            assert rc == 0, (kind, reorder, caps, rc)
            assert np.array_equal(np.sort(order), np.arange(n))
            assert nt.value >= -(-n // caps[2])


def _cholqr(lib, V):
    import ctypes as C
    from feastsolver_jl_b200 import _lib
    n, m = V.shape
    Vf = np.asfortranarray(V, dtype=np.complex128).copy(order="F")
    R = np.zeros((m, m), dtype=np.complex128, order="F")
    passes = C.c_int(0)
    rc = lib.feast_debug_cholqr(n, m, _lib.ptr(Vf), n, _lib.ptr(R), C.byref(passes))
    assert rc == 0
    return Vf, R, passes.value


def test_cholqr_regression_filtered_feast_block(lib):
    """Host side of the orthonormalisation (host_small.h cholqr_pass, shared with the device path): the block that broke
    the first (clamped-pivot) version -- the accumulator Q0 of an emulated nlfeast pass, singular values 3.7 ... 1e-9 --
    must come back orthonormal with V_in = V_out * Rtot to rounding (the old version: reconstruction error 5e17)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "cholqr_failing_block.npz"))
    R0 = d["R0"]
    rng = np.random.default_rng(11)
    n, m = 1500, R0.shape[0]
    Q, _ = np.linalg.qr(rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m)))
    V = Q @ R0
    U, R, passes = _cholqr(lib, V)
    assert passes <= 6
    assert np.abs(U.conj().T @ U - np.eye(m)).max() < 5e-15
    assert np.abs(U @ R - V).max() <= 1e-13 * np.abs(V).max()
    # the small singular values survive in Rtot (Beyn's step divides by them, src/utils.jl:73)
    s_true = np.linalg.svd(R0, compute_uv=False)
    s_got = np.linalg.svd(R, compute_uv=False)
    assert np.abs(s_got - s_true).max() <= 1e-13 * s_true[0]


@pytest.mark.parametrize("cond", [1e3, 1e9, 1e15, 0.0])
def test_cholqr_ill_conditioned_and_rank_deficient(lib, cond):
    rng = np.random.default_rng(5)
    n, m = 1200, 24
    U0, _ = np.linalg.qr(rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m)))
    W0, _ = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    sv = np.logspace(0, -np.log10(cond), m) if cond > 0 else np.concatenate([np.ones(m - 5), np.zeros(5)])
    V = (U0 * sv[None, :]) @ W0.conj().T
    U, R, passes = _cholqr(lib, V)
    assert np.abs(U.conj().T @ U - np.eye(m)).max() < 5e-15
    assert np.abs(U @ R - V).max() <= 1e-13 * np.abs(V).max()


def test_amg_host_hierarchy_galerkin_and_convergence():
    """Host setup of the Krylov preconditioner (csrc/amg_setup.cpp): every coarse slot is the Galerkin product P^T S P of the
    level above, and the numpy restatement of the device V-cycle on that hierarchy cuts the COCG iteration count of a
    shifted C2-type system several times (the device path is checked against the same restatement in the GPU suite)."""
    import scipy.sparse as sp
    from feastsolver_jl_b200 import workloads as wl
    import amg_ref   # tests/amg_ref.py (the tests directory is on sys.path under pytest rootdir conftest)
    m = 22
    A, B = wl.laplacian3d_pencil(m)
    A, B = sp.csr_matrix(A), sp.csr_matrix(B)
    levels, secs = amg_ref.cpp_hierarchy([A, B], max_coarse=2000)
    assert len(levels) >= 2 and levels[0]["n"] == m ** 3 and levels[-1]["n"] <= 2000
    for l in range(len(levels) - 1):
        P = levels[l]["P"]
        assert P.shape == (levels[l]["n"], levels[l + 1]["n"])
        for k in range(2):
            G = (P.T @ levels[l]["slots"][k] @ P).tocsr()
            assert abs(G - levels[l + 1]["slots"][k]).max() < 1e-13 * abs(G).max()
        assert np.allclose(np.asarray(P.sum(axis=1)).ravel()[:5] != 0, True)
    c, r, cnt = wl.c2_slice(m, target=12)
    z = c + r * np.exp(0.35j * np.pi)
    Z = (A - z * B).tocsr()
    rng = np.random.default_rng(0)
    b = rng.standard_normal((m ** 3, 3)) + 1j * rng.standard_normal((m ** 3, 3))
    x0_, it_plain, _ = amg_ref.pcocg(Z, b, 1e-8, 3000)
    x1_, it_amg, rel = amg_ref.pcocg(Z, b, 1e-8, 300, amg_ref.VCycle(levels, [1.0, -z]))
    assert rel < 1e-8 and it_amg * 4 < it_plain, (it_amg, it_plain)
    assert np.abs(Z @ x1_ - b).max() < 1e-6 * np.abs(b).max()


def test_column_slices_partition_the_columns():
    from feastsolver_jl_b200.partition import column_slice
    for m0 in (1, 7, 20, 64, 100):
        for nr in (1, 2, 3, 8):
            sl = [column_slice(m0, nr, r) for r in range(nr)]
            assert sl[0][0] == 0 and sl[-1][1] == m0
            assert all(sl[i][1] == sl[i + 1][0] for i in range(nr - 1))
            assert max(b - a for a, b in sl) - min(b - a for a, b in sl) <= 1


def test_group_choice_of_the_column_sharded_loop(lib):
    """csrc/api.cu pick_groups: G rank groups (nodes -> groups by LPT on the measured costs) x nranks/G column slices,
    scored by makespan x (columns per rank + 12).  With the iteration counts measured on the C2 pencil (near-axis nodes
    ~12x the others) 8 ranks take 4 groups x 2 slices, 2 ranks take 2 groups (pure node sharding), equal costs keep
    one group (pure column split)."""
    from feastsolver_jl_b200 import _lib
    cost = np.array([126, 52, 25, 16, 12, 11, 10, 10] * 2, dtype=np.float64)
    own = np.zeros(16, dtype=np.int32)
    G = lib.feast_debug_pick_groups(16, _lib.ptr(cost), 8, 64, _lib.ptr(own))
    assert G == 4
    loads = np.bincount(own, weights=cost, minlength=G)
    assert loads.max() <= 1.1 * cost.sum() / G and set(own) == set(range(G))
    assert lib.feast_debug_pick_groups(16, _lib.ptr(cost), 2, 64, _lib.ptr(own)) == 2
    flat = np.ones(16)
    assert lib.feast_debug_pick_groups(16, _lib.ptr(flat), 8, 64, _lib.ptr(own)) in (4, 8)   # balanced either way: widest slices win
    one = np.array([100.0] + [1.0] * 15)
    assert lib.feast_debug_pick_groups(16, _lib.ptr(one), 8, 64, _lib.ptr(own)) == 1           # one dominant node: split its columns


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm: oracle port on the host cores) on a tiny grid: one JSON line with the
    contract keys, impl = reference, no GPU and no libfeast_cuda involved."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "10", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "contour_node_solves_per_sec" and d["unit"] == "node_solves/s"
    assert d["steps"] == 2 and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
