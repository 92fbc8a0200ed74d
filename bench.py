#!/usr/bin/env python
"""Benchmark of the FEAST contour-quadrature hot path on B200 (BASELINE.json metric:
time-to-solution and contour-node solves/sec).

Workload (config C2 of BASELINE.json / SURVEY.md 8d): sparse generalized Hermitian pencil,
3-D Laplacian A + Kronecker-sum mass matrix B on a 100^3 grid (n = 1e6, nnz = 6.94e6 each),
lowest spectral slice (~38 eigenvalues), m0 = 64, 16 Gauss-Legendre nodes
(circular_contour_gauss), contour nodes sharded over the N GPUs.

A "step" is ONE outer FEAST iteration over the whole contour: Rayleigh-Ritz projection,
reduced eigenproblem (host LAPACK), Ritz recovery + residual, and the 16 shifted
m0-right-hand-side node solves with fused accumulation (+ NCCL all-reduce of Q for N > 1).
`value` = node solves per second with everything resident in HBM; `e2e` = the same metric
through the public `gen_feast` call with HOST buffers run to convergence (uploads,
downloads and all outer iterations inside the timed region; its wall time is the
time-to-solution).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 100
M0 = 64
NODES = 16
TARGET = 36
INNER_TOL = 1e-5
MAX_INNER = 6000
EPS = 1e-12
CPU_GRID = 24          # reduced grid for the CPU arm (sparse LU of the 100^3 pencil does not fit)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(grid):
    from feastsolver_jl_b200 import workloads as wl
    A, B = wl.laplacian3d_pencil(grid)
    c, r, cnt = wl.c2_slice(grid, target=TARGET)
    X0 = wl.rand_subspace(grid ** 3, M0, seed=0)
    return A, B, c, r, cnt, X0


def spmm_bytes(n, nnz, m0):
    """ALGORITHMIC bytes of one complex-valued CSR SpMM launch: values 16 B +
    4 B column index per nonzero, row pointers, X read once, Y written once."""
    return nnz * (16 + 4) + 4 * (n + 1) + 2 * 16 * n * m0


# ------------------------------------------------------------------------------------ CPU arm
def cpu_sample(nsolves, grid=CPU_GRID):
    """Oracle (numpy/scipy restatement of the reference, SuperLU in place of UMFPACK) on a bounded
    sample: `nsolves` node solves (factor A - zB + solve m0 right-hand sides) at a reduced grid."""
    from oracle import feast_oracle as fo
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        threads = os.cpu_count() or 1
    A, B, c, r, cnt, X0 = build_workload(grid)
    ct = fo.circular_contour_gauss(c, r, NODES)
    R = X0
    t0 = time.perf_counter()
    for k in range(nsolves):
        F = fo.lu_factorizer(A - B * ct.nodes[k % NODES])
        fo.left_divide(F, R)
    dt = time.perf_counter() - t0
    return {"value": nsolves / dt, "unit": "node_solves/s", "cores": int(threads), "kind": "port",
            "sample": f"{nsolves} node solves (sparse LU factor + {M0}-rhs solve) of the same pencil on a "
                      f"{grid}^3 grid (n={grid**3}); the 100^3 sparse LU does not fit host memory/time; "
                      f"numpy/scipy restatement of the Julia reference (SuperLU for UMFPACK)",
            "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import feast_oracle as fo
    A, B, c, r, cnt, X0 = build_workload(CPU_GRID)
    ct = fo.circular_contour_gauss(c, r, NODES)
    # one step = one outer iteration of the oracle driver (16 factor+solve node solves + RR) on the sample
    steps, warm = max(1, args.steps), max(0, min(args.warmup, 1))
    tms = fo.Timers()
    X = X0.copy()
    t_all = []
    for it in range(warm + steps):
        t0 = time.perf_counter()
        fo.gen_feast(X, A, B, ct, iter=1, eps=0.0, timers=tms)  # iter=1: exactly one pass with solves
        t_all.append(time.perf_counter() - t0)
    dt = sum(t_all[warm:])
    val = NODES * steps / dt
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        threads = os.cpu_count() or 1
    sample = (f"oracle gen_feast outer iterations on a {CPU_GRID}^3 grid (n={CPU_GRID**3}), m0={M0}, {NODES} nodes; "
              f"numpy/scipy restatement of the Julia reference (SuperLU for UMFPACK); 100^3 sparse LU infeasible on host")
    out = {"impl": "reference", "metric": "contour_node_solves_per_sec", "value": val, "unit": "node_solves/s",
           "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
           "config": {"workload": f"C2-sample: 3-D Laplacian+mass pencil grid {CPU_GRID}^3, m0={M0}, {NODES} Gauss nodes (CPU arm)"},
           "cpu_baseline": {"value": val, "unit": "node_solves/s", "cores": int(threads), "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "node_solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def outer_iteration(fs, ctx, contour, generalized=True):
    from feastsolver_jl_b200.feast import _eig_sorted
    Aq, Bq = ctx.project(generalized)
    Lam, Xq = _eig_sorted(Aq, Bq)
    res = ctx.recover_residual(Xq, Lam)
    st = ctx.contour_apply(Lam)
    return Lam, res, st


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import feastsolver_jl_b200 as fs
    from feastsolver_jl_b200 import _lib
    from feastsolver_jl_b200.distributed import make_comm_hook
    from feastsolver_jl_b200.partition import node_owners

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    grid = args.grid
    A, B, c, r, cnt, X0 = build_workload(grid)
    n = grid ** 3
    contour = fs.circular_contour_gauss(c, r, NODES)
    hook = make_comm_hook()
    owners = node_owners(contour.nodes, world)
    solver_opts = {"kind": _lib.SOLVER_KRYLOV, "inner_tol": INNER_TOL, "max_inner": MAX_INNER,
                   "precond": {"auto": _lib.PRECOND_AUTO, "none": _lib.PRECOND_NONE, "amg": _lib.PRECOND_AMG}[args.precond]}

    ctx = fs.FeastContext(device=local)
    ctx.set_operator(0, A)
    ctx.set_operator(1, B)
    ctx.set_problem(_lib.PROBLEM_GENERALIZED, 2, n)
    if hook is not None:
        hook(ctx)
    ctx.set_contour(contour.nodes, contour.weights)
    ctx.set_node_owners(owners)
    ctx.set_solver(**solver_opts)
    if args.mixed_prec:
        ctx.set_mixed_precision(True)     # experimental: complex64 COCG blocks (the reference's mixed_prec=true)
    ctx.set_subspace(X0)
    layout = ctx.layout_info()
    pinfo = ctx.preconditioner_info()

    for _ in range(args.warmup):
        outer_iteration(fs, ctx, contour)
    ctx.set_subspace(X0)          # timed steps are the first K iterations of the real solve

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    agg = {"inner_iters_total": 0, "t_spmm_ms": 0.0, "spmm_launches": 0, "t_solve_ms": 0.0, "t_reduce_ms": 0.0}
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        Lam, res, st = outer_iteration(fs, ctx, contour)
        for k in agg:
            agg[k] += st[k]
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    launches = int(sum_over_ranks(ctx.launch_count() - launches0))
    spmm_ms_rank = agg["t_spmm_ms"] / max(1, agg["spmm_launches"])
    spmm_ms = max_over_ranks(spmm_ms_rank)
    inner_total = int(sum_over_ranks(agg["inner_iters_total"]))
    reduce_ms = max_over_ranks(agg["t_reduce_ms"])
    value = NODES * args.steps / (ms / 1e3)
    ctx.close()

    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"metric": "contour_node_solves_per_sec", "value": value, "ms_per_step": ms / args.steps,
                              "spmm_ms_per_launch": spmm_ms, "gpu_launches": launches, "inner_iters_per_step": inner_total / args.steps,
                              "preconditioner": pinfo, "note": "profiling run (--no-e2e)"}))
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- e2e: public API with host buffers, run to convergence (time-to-solution) ----
    st_e2e = {}
    Xh = X0.copy()
    barrier()
    t0 = time.perf_counter()

    ctx2 = fs.FeastContext(device=local)
    e, v, rs = fs.gen_feast(Xh, A, B, contour, eps=EPS, iter=10, ctx=ctx2, solver_opts=solver_opts, stats=st_e2e,
                            comm=hook)
    barrier()
    tts = max_over_ranks(time.perf_counter() - t0)
    ctx2.close()
    iters_with_solves = sum(1 for h in st_e2e["history"] if "nodes_local" in h)
    e2e_val = NODES * iters_with_solves / tts
    h2d = (A.data.nbytes + A.indices.nbytes * 2 + A.indptr.nbytes * 2) * 2 + X0.nbytes  # int64 indices cross the ABI
    d2h = X0.nbytes
    from feastsolver_jl_b200 import workloads as wl
    exact = wl.laplacian3d_spectrum(grid, count=cnt + 8)
    exact = exact[np.abs(exact - c) <= r]
    eig_err = float(np.abs(np.sort(e.real) - exact).max() / np.abs(exact).max()) if e.size == exact.size else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    bytes_per_launch = spmm_bytes(n, A.nnz, M0)
    achieved = bytes_per_launch / (spmm_ms * 1e-3) / 1e9 if spmm_ms > 0 else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "spmm_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    out = {
        "metric": "contour_node_solves_per_sec", "value": value, "unit": "node_solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": f"C2: sparse generalized Hermitian 3-D Laplacian+mass pencil, grid {grid}^3 (n={n}, "
                               f"nnz={A.nnz}), lowest slice ({cnt} eigenvalues), m0={M0}, {NODES} Gauss-Legendre nodes "
                               f"sharded over {world} GPU(s); step = one outer FEAST iteration ({NODES} node solves + RR)",
                   "inner_solver": f"pseudo-block COCG, rel tol {INNER_TOL}" + (", complex64 blocks (mixed_prec, steady-state leg only)" if args.mixed_prec else ""), "l2_policy": "inputs (5 GB of Krylov blocks) exceed the 126 MB L2",
                   "node_owners": [int(o) for o in owners],
                   "layout": {"rows_renumbered": layout["reordered"], "spmm_tiles": layout["ntiles"],
                              "halo_rows_per_row": round(layout["halo_rows_per_row"], 3)}},
        "time_to_solution_s": tts, "outer_iterations": len(st_e2e["history"]), "eigenvalues_found": int(e.size),
        "eigenvalues_exact": int(exact.size), "max_residual": float(rs.max()) if rs.size else None,
        "eig_rel_err_vs_analytic": eig_err, "inner_iters_per_step": inner_total / args.steps,
        "allreduce_ms_per_step": reduce_ms / args.steps,
        "e2e": {"value": e2e_val, "unit": "node_solves/s", "h2d_bytes_per_step": int(h2d / max(1, iters_with_solves)),
                "d2h_bytes_per_step": int(d2h / max(1, iters_with_solves)), "time_to_solution_s": tts,
                "api": "feastsolver_jl_b200.gen_feast(X, A, B, contour) with host numpy/scipy buffers, to convergence"},
        "gpu_launches": launches,
        "roofline": {"kernel": "spmm_tiled_kernel<c128, DOT> (COCG q = (A - zB) p, fused <p,q>)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "bytes_per_launch": bytes_per_launch, "ms_per_launch": spmm_ms,
                     "share_of_step": (agg["t_spmm_ms"] / ms) if world == 1 else None},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_sample(8)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=GRID, help="grid points per dimension (default: the C2 size 100)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs only)")
    ap.add_argument("--precond", default="auto", choices=["auto", "none", "amg"],
                    help="Krylov preconditioner (auto: smoothed-aggregation V-cycle when applicable)")
    ap.add_argument("--mixed-prec", action="store_true",
                    help="EXPERIMENTAL: complex64 storage of the COCG blocks in the steady-state leg (not the headline configuration)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
