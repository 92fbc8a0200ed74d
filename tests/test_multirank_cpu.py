"""World-size-2 gloo test (CPU) of the node-sharded path: each rank accumulates the contour
nodes that feastsolver_jl_b200.partition assigns to it, the partial Q blocks are summed with a
torch.distributed all-reduce (the role ncclAllReduce plays inside libfeast_cuda.so), and the
result must equal the single-rank oracle run.  The per-node compute here is the oracle (this is a
test of the host-side sharding logic, not of the CUDA kernels)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import feast_oracle as fo
    from feastsolver_jl_b200 import workloads as wl
    from feastsolver_jl_b200.partition import column_slice, local_nodes, node_owners
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allreduce(Q):
        t = torch.view_as_real(torch.from_numpy(np.ascontiguousarray(Q)))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return torch.view_as_complex(t).numpy()

    m = 8
    A, B = wl.laplacian3d_pencil(m)
    c, r, cnt = wl.c2_slice(m, target=10)
    ct = fo.circular_contour_gauss(c, r, 16)
    owners = node_owners(ct.nodes, world)
    mine = local_nodes(owners, rank)
    X0 = wl.rand_subspace(m ** 3, 16, seed=0)
    e, v, res = fo.gen_feast(X0.copy(), A, B, ct, iter=8, node_subset=mine, reduce_fn=allreduce)
    # column sharding (the device Krylov path): every rank runs all nodes on its slice of the right-hand-side columns
    ec, vc, resc = fo.gen_feast(X0.copy(), A, B, ct, iter=8, col_slice=column_slice(16, world, rank), reduce_fn=allreduce)
    assert np.abs(np.sort_complex(ec) - np.sort_complex(e)).max() < 1e-11 and resc.max() < 1e-11
    # nlfeast sharded the same way (Q0 and Q1 are reduced)
    coeffs = wl.butterfly_coeffs(8)
    T = fo.polynomial([a.toarray() for a in coeffs])
    zs = fo.circular_contour_trapezoidal(1 + 1j, 0.5, 16).nodes
    own2 = node_owners(zs, world)
    X1 = wl.rand_subspace(64, 20, seed=300)
    lam, X, rs = fo.nlfeast(T, X1.copy(), 16, 12, c=1 + 1j, r=0.5, eps=1e-12,
                            node_subset=local_nodes(own2, rank), reduce_fn=allreduce)
    out_q.put((rank, np.sort_complex(e), res.max(), len(mine), np.sort_complex(lam[np.abs(lam - (1 + 1j)) <= 0.5])))
    dist.barrier()
    dist.destroy_process_group()


def test_node_sharding_world2_gloo():
    sys.path.insert(0, ROOT)
    from oracle import feast_oracle as fo
    from feastsolver_jl_b200 import workloads as wl
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = 8
    A, B = wl.laplacian3d_pencil(m)
    c, r, cnt = wl.c2_slice(m, target=10)
    ct = fo.circular_contour_gauss(c, r, 16)
    e1, _, r1 = fo.gen_feast(wl.rand_subspace(m ** 3, 16, seed=0), A, B, ct, iter=8)
    T = fo.polynomial([a.toarray() for a in wl.butterfly_coeffs(8)])
    lam1, _, _ = fo.nlfeast(T, wl.rand_subspace(64, 20, seed=300), 16, 12, c=1 + 1j, r=0.5, eps=1e-12)
    lam1 = np.sort_complex(lam1[np.abs(lam1 - (1 + 1j)) <= 0.5])
    outs.sort(key=lambda t: t[0])
    assert outs[0][3] + outs[1][3] == 16 and outs[0][3] == 8
    for rank, e, rmax, nloc, lam in outs:
        assert e.size == cnt == e1.size
        assert np.abs(e - np.sort_complex(e1)).max() < 1e-11
        assert rmax < 1e-11
        assert lam.size == lam1.size == 13
        assert np.abs(lam - lam1).max() < 1e-9
    # both ranks hold identical results (replicated reduced problem)
    assert np.array_equal(outs[0][1], outs[1][1])
