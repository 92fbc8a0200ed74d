"""Config C5 (BASELINE.json): multi-interval spectral-slicing sweep -- S disjoint adjacent contours on
the C2 pencil, one slice per GPU ("replicas only": independent contexts, NO data-path collective).
Reports aggregate eigenpairs / second = (sum of eigenpairs found) / (slowest slice's time).

    torchrun --nproc-per-node 8 scripts/c5_sweep.py --grid 100      # one slice per rank
    python scripts/c5_sweep.py --grid 40 --slices 4                  # single process, slices in sequence
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import feastsolver_jl_b200 as fs
from feastsolver_jl_b200 import _lib, workloads as wl

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=100)
ap.add_argument("--slices", type=int, default=None)
ap.add_argument("--target", type=int, default=36)
ap.add_argument("--m0", type=int, default=64)
ap.add_argument("--tol", type=float, default=1e-5)
ap.add_argument("--precond-shift", type=float, default=0.5,
                help="beta of the complex-shifted multigrid preconditioner for the INTERIOR slices (slice 0 uses 0)")
ap.add_argument("--only", type=int, nargs="*", default=None, help="run these slice indices only")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
nslices = a.slices or max(world, 1)
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
A, B = wl.laplacian3d_pencil(a.grid)
n = a.grid ** 3
# adjacent slices: slice s starts where slice s-1 ended
bounds, first = [], 0
for s in range(nslices):
    c, r, cnt = wl.c2_slice(a.grid, target=a.target, first=first)
    bounds.append((c, r, cnt, first))
    first += cnt
mine = [s for s in range(nslices) if s % world == rank and (a.only is None or s in a.only)]
found, t_my, res_max, ok = 0, 0.0, 0.0, True
lam_all = wl.laplacian3d_spectrum(a.grid, count=first + 16)
for s in mine:
    c, r, cnt, f0 = bounds[s]
    ct = fs.circular_contour_gauss(c, r, 16)
    X0 = wl.rand_subspace(n, a.m0, seed=100 + s)
    t0 = time.perf_counter()
    e, v, res = fs.gen_feast(X0, A, B, ct, eps=1e-12, iter=10,
                             solver_opts={"kind": _lib.SOLVER_KRYLOV, "inner_tol": a.tol, "max_inner": 8000,
                                          "precond_shift": a.precond_shift if s > 0 else 0.0},
                             ctx=fs.FeastContext(device=local))
    t_my += time.perf_counter() - t0
    exact = lam_all[f0:f0 + cnt]
    conv = res < 1e-9               # Ritz values inside with a large residual are spurious (feast.jl:77-79 keeps them)
    ec = np.sort(e[conv].real)
    this_ok = ec.size == cnt and np.abs(ec - exact).max() <= 1e-10 * exact.max()
    ok = ok and this_ok
    found += int(ec.size)
    res_max = max(res_max, float(res[conv].max()) if conv.any() else 0.0)
    print(f"[rank {rank}] slice {s}: c={c:.5f} r={r:.5f} expected {cnt} returned {e.size} converged {ec.size} "
          f"ok={this_ok} time {time.perf_counter() - t0:.1f}s precond_shift={a.precond_shift if s > 0 else 0.0}", file=sys.stderr, flush=True)
if world > 1:
    t = torch.tensor([float(found), t_my, res_max, 0.0 if ok else 1.0], dtype=torch.float64, device="cuda")
    tsum, tmax = t.clone(), t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    found, t_my, res_max, ok = int(tsum[0].item()), float(tmax[1].item()), float(tmax[2].item()), tsum[3].item() == 0.0
if rank == 0:
    print(json.dumps({"config": f"C5 multi-interval sweep: {nslices} adjacent slices of the grid {a.grid}^3 pencil over {world} GPU(s), "
                                f"m0={a.m0}, 16 Gauss nodes per slice, replicas only (no collective)",
                      "eigenpairs_found": found, "eigenpairs_expected": int(sum(b[2] for b in bounds)), "matches_analytic": bool(ok),
                      "max_residual": res_max, "slowest_rank_s": t_my, "eigenpairs_per_s": found / t_my}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
