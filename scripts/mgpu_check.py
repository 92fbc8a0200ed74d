"""torchrun --nproc-per-node N scripts/mgpu_check.py : gen_feast (Krylov solves, sharded by right-hand-side columns and,
forced, by contour nodes) and nlfeast (dense LU, node-sharded) on N GPUs compared with a single-GPU run on rank 0 and
with the analytic spectrum."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import feastsolver_jl_b200 as fs
from feastsolver_jl_b200 import _lib, workloads as wl
from feastsolver_jl_b200.distributed import make_comm_hook

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
m = int(os.environ.get("MGPU_GRID", "24"))
A, B = wl.laplacian3d_pencil(m)
c, r, cnt = wl.c2_slice(m, target=20)
X0 = wl.rand_subspace(m ** 3, 32, seed=0)
ct = fs.circular_contour_gauss(c, r, 16)
opts = {"kind": _lib.SOLVER_KRYLOV, "inner_tol": 1e-8}
st = {}
e, v, res = fs.gen_feast(X0.copy(), A, B, ct, eps=1e-12, iter=10, solver_opts=opts, stats=st, comm=make_comm_hook())
exact = wl.laplacian3d_spectrum(m); exact = exact[np.abs(exact - c) <= r]
ok = e.size == exact.size and np.abs(np.sort(e.real) - exact).max() < 1e-10 * exact.max() and res.max() < 1e-11
nodes_local = st["history"][0].get("nodes_local")
col_sharded = st["history"][0].get("col_sharded")
ok = ok and col_sharded >= 1 and 1 <= nodes_local <= 16       # AUTO (Krylov): rank groups x column slices of the right-hand sides
stn = {}
en, vn, resn = fs.gen_feast(X0.copy(), A, B, ct, eps=1e-12, iter=10, solver_opts=dict(opts, shard=_lib.SHARD_NODES), stats=stn,
                            comm=make_comm_hook())
ok = ok and en.size == exact.size and np.abs(np.sort(en.real) - exact).max() < 1e-10 * exact.max() and resn.max() < 1e-11
ok = ok and stn["history"][0].get("col_sharded") == 0 and stn["history"][0].get("nodes_local") == 16 // world
# polynomial problem, dense LU solves, two accumulators reduced
coeffs = [a.toarray() for a in wl.butterfly_coeffs(8)]
lam, X, rs = fs.nlfeast(coeffs, wl.rand_subspace(64, 20, seed=300), 16, 30, c=1 + 1j, r=0.5, eps=1e-12, comm=make_comm_hook())
good = (np.abs(lam - (1 + 1j)) <= 0.5) & (rs < 1e-8)
ok = ok and int(good.sum()) == 13
# every rank must hold the same answer
t = torch.tensor([float(np.sort(e.real).sum()), float(lam[good].real.sum())], dtype=torch.float64, device="cuda")
tmax, tmin = t.clone(), t.clone()
dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
ok = ok and bool((tmax == tmin).all().item())
if rank == 0:
    e1, v1, r1 = fs.gen_feast(X0.copy(), A, B, ct, eps=1e-12, iter=10, solver_opts=opts)
    ok = ok and e1.size == e.size and np.abs(np.sort(e1.real) - np.sort(e.real)).max() < 1e-11 * exact.max()
    print(json.dumps({"ok": bool(ok), "world": world, "found": int(e.size), "exact": int(exact.size),
                      "max_res": float(res.max()), "nodes_local_rank0": nodes_local, "col_sharded": col_sharded, "nl_good": int(good.sum()),
                      "allreduce_ms": st["history"][0].get("t_reduce_ms")}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
