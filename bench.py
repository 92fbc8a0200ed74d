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
SAME_GRID = 48         # largest CPU-feasible grid of the same pencil: both arms run it for the like-for-like ratio
CPU_BUDGET_S = 240.0   # cap on the CPU work of one bench invocation (the sparse LU of the 100^3 pencil does not fit at all)


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
    # column-major like Julia's Matrix{ComplexF64} (the layout that crosses the ABI): no host-side transposes in the timed region
    X0 = np.asfortranarray(wl.rand_subspace(grid ** 3, M0, seed=0))
    return A, B, c, r, cnt, X0


def spmm_bytes(n, nnz, m0):
    """ALGORITHMIC bytes of one complex-valued CSR SpMM launch: values 16 B +
    4 B column index per nonzero, row pointers, X read once, Y written once."""
    return nnz * (16 + 4) + 4 * (n + 1) + 2 * 16 * n * m0


# ------------------------------------------------------------------------------------ CPU arm
def _blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return int(max([p.get("num_threads", 1) for p in threadpool_info()] + [1]))
    except Exception:
        return int(os.cpu_count() or 1)


def cpu_node_solves(grid, nsolves):
    """Oracle (numpy/scipy restatement of the reference, SuperLU in place of UMFPACK): `nsolves` contour-node solves
    (factor A - z_k B, solve M0 right-hand sides: src/feast.jl:141-143) of the C2 pencil at `grid`^3.  Returns seconds."""
    from oracle import feast_oracle as fo
    A, B, c, r, cnt, X0 = build_workload(grid)
    ct = fo.circular_contour_gauss(c, r, NODES)
    t0 = time.perf_counter()
    for k in range(nsolves):
        F = fo.lu_factorizer(A - B * ct.nodes[k % NODES])
        fo.left_divide(F, X0)
    return time.perf_counter() - t0


def pick_cpu_grid(nsolves, budget_s, candidates=(48, 40, 32, 24)):
    """Largest grid whose `nsolves` node solves fit the budget, from one timed solve at 20^3 and the n^2 growth of a
    3-D nested-dissection LU (measured here: 12.6 s / 40.3 s per factorisation at 32^3 / 40^3, ratio 3.2 ~ (40/32)^6 = 3.8)."""
    t20 = cpu_node_solves(20, 1)
    for g in candidates:
        if nsolves * t20 * (g / 20.0) ** 6 <= budget_s:
            return g, t20
    return candidates[-1], t20


def cpu_sample(grid, nsolves):
    dt = cpu_node_solves(grid, nsolves)
    return {"value": nsolves / dt, "unit": "node_solves/s", "cores": _blas_threads(), "kind": "port", "grid": grid,
            "sample": f"{nsolves} contour-node solve(s) (sparse LU of A - zB + {M0}-rhs solve) of the same pencil on a "
                      f"{grid}^3 grid (n={grid**3}): the largest grid whose factorisation fits {CPU_BUDGET_S:.0f} s of host time "
                      f"(the 100^3 sparse LU does not fit host memory/time); numpy/scipy restatement of the Julia reference "
                      f"(SuperLU for UMFPACK)",
            "seconds": dt}


def run_reference(args):
    """Reference arm: the reference's CPU path (oracle port) on the host cores.  One step = ONE contour-node solve of the
    C2 pencil at the largest grid for which steps + warmup solves fit CPU_BUDGET_S."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, min(args.warmup, 1))
    grid = args.grid if args.grid_given else pick_cpu_grid(steps + warm, CPU_BUDGET_S)[0]
    if warm:
        cpu_node_solves(grid, warm)
    dt = cpu_node_solves(grid, steps)
    val = steps / dt
    sample = (f"{steps} contour-node solves (sparse LU + {M0}-rhs solve; one per step) on a {grid}^3 grid (n={grid**3}), "
              f"the largest of 48/40/32/24 whose {steps + warm} solves fit {CPU_BUDGET_S:.0f} s; numpy/scipy restatement of the "
              f"Julia reference (SuperLU for UMFPACK); the 100^3 sparse LU is infeasible on the host")
    out = {"impl": "reference", "metric": "contour_node_solves_per_sec", "value": val, "unit": "node_solves/s",
           "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
           "config": {"workload": f"C2 pencil (3-D Laplacian + mass), grid {grid}^3, m0={M0}, {NODES} Gauss nodes; CPU arm: "
                                  f"one node solve per step"},
           "cpu_baseline": {"value": val, "unit": "node_solves/s", "cores": _blas_threads(), "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "node_solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def rr_phase(ctx, generalized=True):
    from feastsolver_jl_b200.feast import _eig_sorted
    Aq, Bq = ctx.project(generalized)
    Lam, Xq = _eig_sorted(Aq, Bq)
    res = ctx.recover_residual(Xq, Lam)
    return Lam, res


FIRST_PASS_TOL = 1e-3   # e2e leg: inner tolerance of the first contour pass (random start: the 16-node filter contracts by ~5e-5 per pass anyway)


def e2e_solve(fs, A, B, contour, X0, solver_opts, device, hook):
    """The public call with HOST buffers, run to convergence (context creation, operator upload, layout build, all
    outer iterations and the download of X inside the timed region)."""
    st = {}
    Xh = X0.copy(order="F")
    t0 = time.perf_counter()
    ctx = fs.FeastContext(device=device)
    t1 = time.perf_counter()
    e, v, rs = fs.gen_feast(Xh, A, B, contour, eps=EPS, iter=10, ctx=ctx, solver_opts=dict(solver_opts, first_pass_tol=FIRST_PASS_TOL),
                            stats=st, comm=hook)
    t2 = time.perf_counter()
    ctx.close()
    t3 = time.perf_counter()
    st["phases"]["ctx_create_s"] = t1 - t0                     # outside the driver: context + stream + events
    st["phases"]["ctx_destroy_s"] = t3 - t2                    # frees every device allocation (blocks, hierarchy, cached inverses)
    st["phases"]["driver_total_s"] = t2 - t1                   # the gen_feast call itself (sum of the phases above + result slicing)
    return e, rs, st, t3 - t0


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
    from feastsolver_jl_b200 import _lib, workloads as wl
    from feastsolver_jl_b200.contour import in_contour
    from feastsolver_jl_b200.distributed import make_comm_hook
    from feastsolver_jl_b200.partition import node_owners

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MAX) if world > 1 else x

    def sum_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.SUM) if world > 1 else x

    grid = args.grid
    A, B, c, r, cnt, X0 = build_workload(grid)
    n = grid ** 3
    contour = fs.circular_contour_gauss(c, r, NODES)
    hook = make_comm_hook()
    owners = node_owners(contour.nodes, world)
    solver_opts = {"kind": _lib.SOLVER_KRYLOV, "inner_tol": INNER_TOL, "max_inner": MAX_INNER,
                   "precond": {"auto": _lib.PRECOND_AUTO, "none": _lib.PRECOND_NONE, "amg": _lib.PRECOND_AMG}[args.precond]}

    ctx = fs.FeastContext(device=local)
    ctx.set_solver(**solver_opts)
    ctx.set_operator(0, A)
    ctx.set_operator(1, B)
    ctx.set_problem(_lib.PROBLEM_GENERALIZED, 2, n)
    if hook is not None:
        hook(ctx)
    ctx.set_contour(contour.nodes, contour.weights)
    ctx.set_node_owners(owners)
    if args.mixed_prec:
        ctx.set_mixed_precision(True)     # complex64 COCG blocks (the reference's mixed_prec=true); unpreconditioned path
    ctx.set_subspace(X0)
    layout = ctx.layout_info()
    pinfo = ctx.preconditioner_info()

    def converged(Lam, res):
        ins = in_contour(Lam, contour)
        return bool(ins.any() and res[ins].max() < EPS)

    for _ in range(args.warmup):
        Lam, res = rr_phase(ctx)
        ctx.contour_apply(Lam)
    ctx.set_subspace(X0)

    # ---- timed region: K steps, every one a REAL (pre-convergence) outer iteration of the solve from X0.  When the
    # solve converges inside the region it is restarted from X0 (the re-upload and its Rayleigh-Ritz phase stay inside the
    # bracket), so no post-convergence pass -- which would need fewer inner iterations -- is ever timed.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    agg = {"inner_iters_total": 0, "t_spmm_ms": 0.0, "spmm_launches": 0, "t_solve_ms": 0.0, "t_factor_ms": 0.0, "t_reduce_ms": 0.0}
    restarts = 0
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        Lam, res = rr_phase(ctx)
        if converged(Lam, res):
            ctx.set_subspace(X0)
            restarts += 1
            Lam, res = rr_phase(ctx)
        st = ctx.contour_apply(Lam)
        for k in agg:
            agg[k] += st[k]
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    launches = int(sum_over_ranks(ctx.launch_count() - launches0))
    spmm_ms = max_over_ranks(agg["t_spmm_ms"] / max(1, agg["spmm_launches"]))
    inner_total = int(sum_over_ranks(agg["inner_iters_total"]))
    col_sharded = int(st.get("col_sharded", 0))
    if col_sharded:
        inner_total //= col_sharded     # every node ran on `col_sharded` ranks, each on its own slice of the right-hand-side columns
    reduce_ms = max_over_ranks(agg["t_reduce_ms"])
    solve_ms_max, solve_ms_sum = max_over_ranks(agg["t_solve_ms"] + agg["t_factor_ms"]), sum_over_ranks(agg["t_solve_ms"] + agg["t_factor_ms"])
    value = NODES * args.steps / (ms / 1e3)
    # ---- dominant vector kernel of the Krylov iteration, timed alone on the resident blocks (CUDA events, library stream)
    vec_ms = ctx.kernel_time("cocg_direction", reps=20) if rank == 0 else None
    ctx.close()

    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"metric": "contour_node_solves_per_sec", "value": value, "ms_per_step": ms / args.steps,
                              "spmm_ms_per_launch": spmm_ms, "gpu_launches": launches, "inner_iters_per_step": inner_total / args.steps,
                              "preconditioner": pinfo, "vector_kernel_ms": vec_ms, "col_sharded": col_sharded,
                              "allreduce_ms_per_step": reduce_ms / args.steps, "note": "profiling run (--no-e2e)"}))
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- e2e: public API with host buffers, run to convergence (time-to-solution) ----
    barrier()
    t0 = time.perf_counter()
    e, rs, st_e2e, _ = e2e_solve(fs, A, B, contour, X0, solver_opts, local, hook)
    barrier()
    tts = max_over_ranks(time.perf_counter() - t0)
    iters_with_solves = sum(1 for h in st_e2e["history"] if "nodes_local" in h)
    e2e_val = NODES * iters_with_solves / tts
    h2d = (A.data.nbytes + A.indices.nbytes * 2 + A.indptr.nbytes * 2) * 2 + X0.nbytes  # int64 indices cross the ABI
    d2h = X0.nbytes
    exact = wl.laplacian3d_spectrum(grid, count=cnt + 8)
    exact = exact[np.abs(exact - c) <= r]
    eig_err = float(np.abs(np.sort(e.real) - exact).max() / np.abs(exact).max()) if e.size == exact.size else None
    phases = {k: (round(v, 4) if not isinstance(v, list) else [round(x, 4) for x in v]) for k, v in st_e2e["phases"].items()}
    phases_max = {k: max_over_ranks(v) for k, v in st_e2e["phases"].items() if not isinstance(v, list)} if world > 1 else None

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
    vec_bytes = 5 * 16 * n * M0   # x, p read + written, z read: the direction kernel of the (preconditioned) COCG iteration
    out = {
        "metric": "contour_node_solves_per_sec", "value": value, "unit": "node_solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": f"C2: sparse generalized Hermitian 3-D Laplacian+mass pencil, grid {grid}^3 (n={n}, "
                               f"nnz={A.nnz}), lowest slice ({cnt} eigenvalues), m0={M0}, {NODES} Gauss-Legendre nodes "
                               f"sharded over {world} GPU(s); step = one outer FEAST iteration ({NODES} node solves + RR)",
                   "inner_solver": ("pseudo-block COCG" + (" preconditioned by a smoothed-aggregation V(1,1) cycle, levels "
                                    + "/".join(str(v) for v in pinfo["sizes"]) if pinfo["levels"] else "")
                                    + f", rel tol {INNER_TOL}" + (", complex64 blocks (mixed_prec)" if args.mixed_prec else "")),
                   "timed_steps": "real pre-convergence passes only (solve restarted from X0 when it converges inside the "
                                  f"timed region; {restarts} restart(s)); per-node coarse multigrid inverses are cached after the first pass "
                                  "of a contour, as in passes 2+ of a solve (the e2e figure includes computing them)",
                   "l2_policy": "inputs (5 GB of Krylov blocks) exceed the 126 MB L2",
                   "sharding": (f"{world // col_sharded} node group(s) x {col_sharded} column slice(s) of the right-hand sides "
                                "(groups chosen from the measured node costs)" if col_sharded else "contour nodes") if world > 1 else "single GPU",
                   "node_owners": [int(o) for o in owners],
                   "layout": {"rows_renumbered": layout["reordered"], "spmm_tiles": layout["ntiles"],
                              "halo_rows_per_row": round(layout["halo_rows_per_row"], 3)}},
        "value_real_passes": value,
        "time_to_solution_s": tts, "outer_iterations": len(st_e2e["history"]), "eigenvalues_found": int(e.size),
        "eigenvalues_exact": int(exact.size), "max_residual": float(rs.max()) if rs.size else None,
        "eig_rel_err_vs_analytic": eig_err, "inner_iters_per_step": inner_total / args.steps,
        "ms_per_inner_iteration": (solve_ms_sum / max(1, inner_total)),
        "allreduce_ms_per_step": reduce_ms / args.steps,
        "node_solve_ms_per_step": {"max_over_ranks": solve_ms_max / args.steps, "mean_over_ranks": solve_ms_sum / world / args.steps},
        "e2e": {"value": e2e_val, "unit": "node_solves/s", "h2d_bytes_per_step": int(h2d / max(1, iters_with_solves)),
                "d2h_bytes_per_step": int(d2h / max(1, iters_with_solves)), "time_to_solution_s": tts,
                "api": "feastsolver_jl_b200.gen_feast(X, A, B, contour) with host numpy/scipy buffers, to convergence",
                "inner_tol_schedule": [FIRST_PASS_TOL, INNER_TOL],
                "phases_rank0_s": phases, "phases_max_over_ranks_s": phases_max,
                "preconditioner_setup_s": st_e2e["preconditioner"]["setup_s"]},
        "gpu_launches": launches, "host_cpus": os.cpu_count(),
        "roofline": {"kernel": "spmm_tiled_kernel<c128, DOT> (COCG q = (A - zB) p, fused <p,q>)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "bytes_per_launch": bytes_per_launch, "ms_per_launch": spmm_ms,
                     "share_of_step": (agg["t_spmm_ms"] / ms) if world == 1 else None,
                     "note": "timed inside the solves (events around every q = Zp launch); the V-cycle launches the same kernel twice more per iteration"},
        "roofline_vector_kernel": {"kernel": "cocg_p_kernel (x += alpha p ; p = z + beta p)", "bound": "hbm",
                                   "achieved": vec_bytes / (vec_ms * 1e-3) / 1e9 if vec_ms else None, "peak": peak, "unit": "GB/s",
                                   "frac": (vec_bytes / (vec_ms * 1e-3) / 1e9 / peak) if vec_ms else None,
                                   "bytes_per_launch": vec_bytes, "ms_per_launch": vec_ms,
                                   "note": "timed alone on the resident blocks after the timed region (20 launches, CUDA events)"},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        # like-for-like leg: BOTH arms on the same pencil at the largest CPU-feasible grid (cap: CPU_BUDGET_S of host time)
        g, t20 = pick_cpu_grid(1, CPU_BUDGET_S, candidates=(SAME_GRID, 40, 32, 24))
        cpu = cpu_sample(g, 1)
        A2, B2, c2, r2, cnt2, X2 = build_workload(g)
        ct2 = fs.circular_contour_gauss(c2, r2, NODES)
        e2, rs2, st2, tts2 = e2e_solve(fs, A2, B2, ct2, X2, solver_opts, local, None)
        it2 = sum(1 for h in st2["history"] if "nodes_local" in h)
        out["cpu_baseline"] = cpu
        out["same_config"] = {
            "workload": f"C2 pencil on a {g}^3 grid (n={g**3}), m0={M0}, {NODES} Gauss nodes, lowest slice ({cnt2} eigenvalues)",
            "same_config": True, "grid_cap": f"largest of {SAME_GRID}/40/32/24 whose sparse LU fits {CPU_BUDGET_S:.0f} s on this host "
                                             f"(one 20^3 node solve took {t20:.2f} s)",
            "gpu_e2e_node_solves_per_s": NODES * it2 / tts2, "gpu_time_to_solution_s": tts2, "gpu_eigenvalues_found": int(e2.size),
            "gpu_max_residual": float(rs2.max()) if rs2.size else None,
            "cpu_node_solves_per_s": cpu["value"], "cpu_cores": cpu["cores"], "cpu_sample": "1 node solve (factor + 64-rhs solve)",
            "e2e_ratio_gpu_over_cpu": (NODES * it2 / tts2) / cpu["value"]}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------ C3 (dense) line
def fp64_tensor_peak(torch, N=8192):
    """FP64 tensor-pipe denominator measured in THIS run: cuBLAS ZGEMM N^3 through torch (8 N^3 real flops), best of 3."""
    a = torch.randn(N, N, dtype=torch.complex128, device="cuda")
    b = torch.randn(N, N, dtype=torch.complex128, device="cuda")
    (a @ b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        (a @ b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 8.0 * N ** 3 / (best * 1e-3) / 1e12


def run_c3(args):
    """BASELINE config 3: dense non-Hermitian complex n = 16384, circular contour, m0 = 128, 32 trapezoid nodes, feast!
    with store=true (one LU per node, reused by every outer iteration).  python bench.py --config C3 [--gpus N]"""
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
    from feastsolver_jl_b200 import workloads as wl
    from feastsolver_jl_b200.distributed import make_comm_hook
    n, m0, nodes, rad = (args.grid if args.grid_given else 16384), 128, 32, 7.0
    peak = fp64_tensor_peak(torch) if rank == 0 else None
    A = wl.dense_nonhermitian(n, seed=1551)
    X0 = wl.rand_subspace(n, m0, seed=0)
    hook = make_comm_hook()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    st = {}
    t0 = time.perf_counter()
    e, v, res = fs.feast(X0, A, nodes=nodes, iter=10, c=0.0, r=rad, eps=1e-12, store=True, stats=st, comm=hook)
    torch.cuda.synchronize()
    tts = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    hist = [h for h in st["history"] if "nodes_local" in h]

    def red(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())
    tts = red(tts, dist.ReduceOp.MAX) if world > 1 else tts
    t_factor = red(sum(h["t_factor_ms"] for h in hist), dist.ReduceOp.MAX) if world > 1 else sum(h["t_factor_ms"] for h in hist)
    t_solve = red(sum(h["t_solve_ms"] for h in hist), dist.ReduceOp.MAX) if world > 1 else sum(h["t_solve_ms"] for h in hist)
    t_pass = red(sum(h["t_total_ms"] for h in hist), dist.ReduceOp.MAX) if world > 1 else sum(h["t_total_ms"] for h in hist)
    launches = int(red(st["launches"], dist.ReduceOp.SUM)) if world > 1 else int(st["launches"])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    steps = len(hist)
    local_nodes = nodes // world
    lu_tf = (8.0 / 3.0) * n ** 3 * local_nodes / (t_factor * 1e-3) / 1e12 if t_factor else None
    getrs_tf = 8.0 * n ** 2 * m0 * local_nodes * steps / (t_solve * 1e-3) / 1e12 if t_solve else None
    out = {"metric": "contour_node_solves_per_sec", "value": nodes * steps / (t_pass * 1e-3), "unit": "node_solves/s", "n_gpus": world,
           "steps": steps, "warmup": 0, "ms_per_step": t_pass / max(1, steps), "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "c128", "data": "synthetic",
           "config": {"workload": f"C3: dense non-Hermitian complex n={n} standard problem, circular contour r={rad}, m0={m0}, "
                                  f"{nodes} trapezoid nodes sharded over {world} GPU(s), feast! with store=true; step = one outer "
                                  f"iteration ({nodes} node solves; the {nodes} LU factorisations happen in the first one)",
                      "l2_policy": "operands (4.3 GB per matrix) exceed the 126 MB L2"},
           # feast! returns every Ritz value inside the contour, spurious ones included (src/feast.jl:77-79): with m0 = 128 > #eigenvalues
           # one or two of them wander through the contour, which is also why the run uses all 10 iterations (feast.jl:53 looks at the
           # maximum over everything inside).  The converged pairs are the ones with a small residual.
           "time_to_solution_s": tts, "outer_iterations": len(st["history"]), "ritz_values_inside": int(e.size),
           "eigenvalues_found": int((res < 1e-9).sum()), "max_residual": float(res[res < 1e-9].max()) if (res < 1e-9).any() else None,
           "max_residual_all_inside": float(res.max()) if res.size else None,
           "e2e": {"value": nodes * steps / tts, "unit": "node_solves/s", "h2d_bytes_per_step": int((A.nbytes + X0.nbytes) / max(1, steps)),
                   "d2h_bytes_per_step": int(X0.nbytes / max(1, steps)), "time_to_solution_s": tts,
                   "api": "feastsolver_jl_b200.feast(X, A; nodes, c, r, store=true) with host numpy buffers, to convergence",
                   "phases_rank0_s": {k: (round(x, 3) if not isinstance(x, list) else [round(y, 3) for y in x]) for k, x in st["phases"].items()}},
           "gpu_launches": launches,
           "roofline": {"kernel": "dense LU (lu_panel_kernel + laswp + trsm + zgemm_dmma_kernel trailing updates)", "bound": "tensor",
                        "achieved": lu_tf, "peak": peak, "peak_source": "cuBLAS ZGEMM 8192^3 through torch, measured in this run (FP64 tensor pipe)",
                        "unit": "TFLOP/s", "frac": (lu_tf / peak) if (lu_tf and peak) else None, "traffic": None,
                        "flops_per_launch": (8.0 / 3.0) * n ** 3, "getrs_tflops": getrs_tf,
                        "getrs_frac": (getrs_tf / peak) if (getrs_tf and peak) else None},
           "clocks": clocks}
    if world == 1 and not args.no_cpu:
        import scipy.linalg as sla
        nc = min(n, 8192)    # bounded sample: LU + m0-rhs solve at n = 8192 (1/8 of the flops of one C3 node solve)
        Z = A[:nc, :nc] - (0.3 + 0.7j) * np.eye(nc)
        t0 = time.perf_counter()
        lu = sla.lu_factor(Z, check_finite=False)
        sla.lu_solve(lu, X0[:nc], check_finite=False)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "node_solves/s", "cores": _blas_threads(), "kind": "port",
                               "sample": f"1 node solve (zgetrf + {m0}-rhs zgetrs through scipy/OpenBLAS, the reference's LAPACK path) of the leading "
                                         f"{nc} x {nc} block: {(nc / n) ** 3:.3f} of the flops of a C3 node solve", "seconds": dt,
                               "tflops": ((8.0 / 3.0) * nc ** 3 + 8.0 * nc ** 2 * m0) / dt / 1e12}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=None, help="grid points per dimension (default: the C2 size 100)")
    ap.add_argument("--config", default="C2", choices=["C2", "C3"], help="C2 (default, the headline) or C3 (dense n=16384 line)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs only)")
    ap.add_argument("--precond", default="auto", choices=["auto", "none", "amg"],
                    help="Krylov preconditioner (auto: smoothed-aggregation V-cycle when applicable)")
    ap.add_argument("--mixed-prec", action="store_true",
                    help="mixed_prec: complex64 storage of the COCG blocks, unpreconditioned recurrence (not the headline configuration)")
    args = ap.parse_args()
    args.grid_given = args.grid is not None
    if args.grid is None:
        args.grid = GRID
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "C3":
        run_c3(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
