"""Contour-node -> rank assignment (the axis the reference threads over:
src/feast.jl:34, src/nlfeast.jl:19,36).

Krylov iteration counts grow as a node approaches the real axis (SURVEY.md
section 7, hard part 4), so a balanced map pairs cheap and expensive nodes on
each rank instead of handing out contiguous runs.
"""
from __future__ import annotations

import numpy as np


def node_cost(nodes, centre=None):
    """Relative cost model: inverse distance of the node from the real axis through
    the contour centre (near-axis nodes make A - zB nearly singular)."""
    z = np.asarray(nodes, dtype=complex)
    c = np.mean(z) if centre is None else complex(centre)
    r = np.max(np.abs(z - c))
    return 1.0 / (np.abs(z.imag - c.imag) / r + 0.15)


def node_owners(nodes, nranks, costs=None):
    """Greedy longest-processing-time assignment with equal node counts per rank when
    nranks divides the node count.  Returns an int32 array owner[k]."""
    nn = len(nodes)
    costs = node_cost(nodes) if costs is None else np.asarray(costs, dtype=float)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(nranks)
    count = np.zeros(nranks, dtype=int)
    cap = -(-nn // nranks)
    owner = np.zeros(nn, dtype=np.int32)
    for k in order:
        cand = [r for r in range(nranks) if count[r] < cap]
        r = min(cand, key=lambda q: (load[q], q))
        owner[k] = r
        load[r] += costs[k]
        count[r] += 1
    return owner


def local_nodes(owner, rank):
    return [k for k, o in enumerate(owner) if o == rank]


def column_slice(m0, nranks, rank):
    """Columns [j0, j1) of every node's right-hand side that `rank` solves when the contour loop shards COLUMNS
    (Krylov inner solves; csrc/api.cu contour_node): j0 = floor(rank m0 / nranks), j1 = floor((rank + 1) m0 / nranks)."""
    return (rank * m0) // nranks, ((rank + 1) * m0) // nranks
