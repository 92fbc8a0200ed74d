"""Dense path bench: complex128 LU (factorize through the plugin ABI) + m0-rhs solve, against the
cuBLAS/cuSOLVER denominators measured through torch (ZGEMM 8192^3, DGEMM 8192^3, torch.linalg.lu_factor)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import feastsolver_jl_b200 as fs
from feastsolver_jl_b200 import workloads as wl

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[2048, 4096, 8192])
ap.add_argument("--m0", type=int, default=128)
ap.add_argument("--peaks", action="store_true")
a = ap.parse_args()
out = {}
if a.peaks:
    import torch
    def timeit(f, reps=5):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(reps):
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    N = 8192
    A = torch.randn(N, N, dtype=torch.float64, device="cuda"); B = torch.randn(N, N, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: A @ B); out["cublas_dgemm_8192_tflops"] = 2 * N**3 / ms / 1e9
    Ac = torch.randn(N, N, dtype=torch.complex128, device="cuda"); Bc = torch.randn(N, N, dtype=torch.complex128, device="cuda")
    ms = timeit(lambda: Ac @ Bc, 3); out["cublas_zgemm_8192_tflops"] = 8 * N**3 / ms / 1e9
    ms = timeit(lambda: torch.linalg.lu_factor(Ac), 2); out["cusolver_zgetrf_8192_tflops"] = (8 / 3) * N**3 / ms / 1e9
    del A, B, Ac, Bc
    torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)
for n in a.n:
    A = wl.dense_nonhermitian(n, seed=1551)
    with fs.FeastContext() as ctx:
        ctx.set_operator(0, A)
        ctx.set_problem(0, 1, n)
        ctx.sync()
        z = 0.3 + 0.7j
        F = ctx.factorize([1.0, -z]); ctx.factor_free(F)       # warm-up
        t0 = time.perf_counter(); F = ctx.factorize([1.0, -z]); t1 = time.perf_counter()
        B = wl.rand_subspace(n, a.m0, seed=1)
        Y = ctx.solve(F, B)                                     # warm-up (allocations)
        t2 = time.perf_counter(); Y = ctx.solve(F, B); t3 = time.perf_counter()
        ctx.factor_free(F)
    res = np.abs((A - z * np.eye(n)) @ Y - B).max() / np.abs(B).max() if n <= 8192 else None
    print(json.dumps({"n": n, "lu_s": t1 - t0, "lu_tflops": (8 / 3) * n**3 / (t1 - t0) / 1e12,
                      "solve_s_incl_copies": t3 - t2, "relres": res}), flush=True)
