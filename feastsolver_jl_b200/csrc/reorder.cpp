// Tile plan of the sparse operators (host, runs once per feast_set_problem).
//
// The tiled CSR SpMM of spmm.cu computes entirely out of shared memory: a CTA pass brings the rows of
// the input block that one TILE of consecutive matrix rows references -- the tile's own rows plus its
// "halo" (referenced rows outside the tile) -- into shared memory with bulk async copies and indexes them
// with 16-bit tile-local column numbers.  L2->SM traffic per SpMM is therefore (1 + halo/rows) block reads.
// In the natural ordering of a 3-D discretisation a run of consecutive rows references ~4 halo rows per
// row (the +-nx and +-nx*ny neighbours; ncu round 1: 5.5 row reads per row, L2->SM bound).  Renumbering the
// rows so that each tile is a compact ball of the matrix graph cuts the halo to ~1.1 rows per row on the
// 100^3 7-point pencil.
//
// Greedy graph growing: a tile is grown breadth-first from a seed until (rows + halo) would exceed the
// shared-memory row capacity, or its nonzeros the staged-CSR capacity; the vertices left in its queue
// seed the following tiles (FIFO), which makes consecutive tiles spatial neighbours -- concurrently running
// CTAs then share their halos in L2.  O(nnz).  Without renumbering the same capacity rule cuts the natural
// ordering into tiles.  Eigenvalues are invariant under the symmetric permutation and blocks are permuted
// back on download, so none of this is visible at the C ABI (the reference hands the matrix to UMFPACK,
// which reorders internally for the same reason: src/feast.jl:36,65 `lu`).
#include "reorder.h"

#include <limits.h>

#include <algorithm>

#include "../../include/feast_cuda.h"
#include "host_small.h"

// Greedy growth of tiles.  dom (optional): vertex -> domain; tiles never cross a domain boundary and the domains
// are swept one after the other (dom_list holds the vertices grouped by domain, dom_ptr the group offsets).
static void grow_tiles(int64_t n, const int64_t* rowptr, const int* col, bool reorder, const TileCaps& caps, const int* dom,
                       const std::vector<int>* dom_list, const std::vector<int>* dom_ptr, TilePlan& plan) {
    plan.order.clear();
    plan.order.reserve((size_t)n);
    plan.tile_ptr.assign(1, 0);
    plan.ok = true;
    std::vector<unsigned char> assigned((size_t)n, 0);
    std::vector<int> mark((size_t)n, -1);  // tile that referenced (or queued) the vertex last
    std::vector<int> seeds;                // FIFO of candidate seeds
    std::vector<int> q;
    int tile = 0;
    const int ndom = dom ? (int)dom_ptr->size() - 1 : 1;
    for (int d = 0; d < ndom; ++d) {
        seeds.clear();
        size_t seed_head = 0;
        int64_t scan = dom ? (*dom_ptr)[d] : 0;
        const int64_t scan_end = dom ? (*dom_ptr)[d + 1] : n;
        int64_t todo = scan_end - scan;
        auto next_seed = [&](int tile_id) -> int {
            if (reorder) {
                while (seed_head < seeds.size()) {
                    const int c = seeds[seed_head++];
                    if (!assigned[c] && mark[c] != tile_id) return c;
                }
            }
            while (scan < scan_end && assigned[dom ? (*dom_list)[scan] : scan]) ++scan;
            return scan < scan_end ? (dom ? (*dom_list)[scan] : (int)scan) : -1;
        };
        while (todo > 0) {
            q.clear();
            size_t head = 0;
            int rows = 0, refs = 0;   // refs = distinct vertices that are members of, or referenced by, the tile
            int64_t nnzt = 0;
            while (todo > 0) {
                if (head == q.size()) {  // tile (still) empty, natural order, or the component is exhausted: new seed
                    const int s = next_seed(tile);
                    if (s < 0) break;
                    if (mark[s] != tile) { mark[s] = tile; ++refs; }
                    q.push_back(s);
                }
                const int v = q[head];
                if (assigned[v]) { ++head; continue; }
                int newrefs = 0;
                for (int64_t e = rowptr[v]; e < rowptr[v + 1]; ++e) newrefs += mark[col[e]] != tile;
                const int64_t deg = (rowptr[v + 1] - rowptr[v] + 7) & ~(int64_t)7;   // rows are padded to 8 entries on the device
                if (rows > 0 && (refs + newrefs > caps.rows_cap || nnzt + deg > caps.nnz_cap || rows >= caps.tile_max)) break;
                if (refs + newrefs > caps.rows_cap || deg > caps.nnz_cap) plan.ok = false;   // a single row does not fit
                ++head;
                assigned[v] = 1;
                plan.order.push_back(v);
                ++rows;
                --todo;
                nnzt += deg;
                for (int64_t e = rowptr[v]; e < rowptr[v + 1]; ++e) {
                    const int u = col[e];
                    if (mark[u] != tile) {
                        mark[u] = tile;
                        ++refs;
                        if (reorder && !assigned[u] && (!dom || dom[u] == d)) q.push_back(u);
                    }
                }
            }
            if (reorder)
                for (size_t t = head; t < q.size(); ++t)
                    if (!assigned[q[t]]) seeds.push_back(q[t]);
            if (seed_head > ((size_t)1 << 22)) {  // compact the FIFO
                seeds.erase(seeds.begin(), seeds.begin() + (long)seed_head);
                seed_head = 0;
            }
            plan.tile_ptr.push_back((int)plan.order.size());
            ++tile;
        }
    }
}

void build_tile_order(int64_t n, const int64_t* rowptr, const int* col, bool reorder, const TileCaps& caps, TilePlan& plan) {
    if (!reorder || caps.domain_rows <= 0 || n <= 2 * (int64_t)caps.domain_rows) {
        grow_tiles(n, rowptr, col, reorder, caps, nullptr, nullptr, nullptr, plan);
        return;
    }
    // Two levels: the same greedy growth first cuts the graph into compact DOMAINS of ~domain_rows vertices, then the
    // tiles are grown inside one domain after the other.  A one-level sweep of a 3-D grid has a wavefront of
    // ~(n/tile)^(2/3) tiles, so a halo row comes back after ~160 MB of traffic at n = 1e6 -- beyond the 126 MB L2
    // (ncu: +36 % DRAM reads); inside a domain the wavefront is a few tens of MB and only domain surfaces miss.
    TileCaps big{INT32_MAX, INT64_MAX, caps.domain_rows, 0};
    TilePlan level1;
    grow_tiles(n, rowptr, col, true, big, nullptr, nullptr, nullptr, level1);
    const int ndom = (int)level1.tile_ptr.size() - 1;
    std::vector<int> dom((size_t)n), dom_ptr(level1.tile_ptr);
    for (int d = 0; d < ndom; ++d)
        for (int i = level1.tile_ptr[d]; i < level1.tile_ptr[d + 1]; ++i) dom[level1.order[i]] = d;
    grow_tiles(n, rowptr, col, true, caps, dom.data(), &level1.order, &dom_ptr, plan);
}

int build_tile_halo(int64_t n, const int64_t* rowptr, const int* col, TilePlan& plan, std::vector<uint16_t>& lcol) {
    const int ntiles = (int)plan.tile_ptr.size() - 1;
    plan.halo_ptr.assign(1, 0);
    plan.halo_idx.clear();
    lcol.resize((size_t)rowptr[n]);
    std::vector<int> pos((size_t)n, -1);   // position in the current tile's halo list
    std::vector<int> stamp((size_t)n, -1);
    std::vector<int> halo;
    int64_t total_halo = 0;
    for (int t = 0; t < ntiles; ++t) {
        const int r0 = plan.tile_ptr[t], r1 = plan.tile_ptr[t + 1];
        halo.clear();
        for (int i = r0; i < r1; ++i)
            for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
                const int c = col[e];
                if (c < 0 || (c >= r0 && c < r1)) continue;   // c < 0: padding entry
                if (stamp[c] != t) { stamp[c] = t; halo.push_back(c); }
            }
        std::sort(halo.begin(), halo.end());   // neighbouring halo rows are fetched from neighbouring addresses
        const int rows = r1 - r0;
        if (rows + (int)halo.size() >= 65535) return -1;   // 0xFFFF marks padding
        for (size_t h = 0; h < halo.size(); ++h) pos[halo[h]] = rows + (int)h;
        for (int i = r0; i < r1; ++i)
            for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
                const int c = col[e];
                lcol[(size_t)e] = c < 0 ? (uint16_t)0xFFFF : (uint16_t)((c >= r0 && c < r1) ? c - r0 : pos[c]);
            }
        plan.halo_idx.insert(plan.halo_idx.end(), halo.begin(), halo.end());
        plan.halo_ptr.push_back((int)plan.halo_idx.size());
        total_halo += (int64_t)halo.size();
    }
    plan.halo_ratio = n ? (double)total_halo / (double)n : 0.0;
    return 0;
}

// Host-only diagnostic entry (no device needed): the plan the library would build for a 0-based CSR pattern.
extern "C" FEAST_API int feast_debug_tile_plan(int64_t n, const int64_t* rowptr, const int* col, int reorder, int rows_cap,
                                               int nnz_cap, int tile_max, int domain_rows, int* order, int* ntiles,
                                               double* halo_ratio) {
    if (n < 1) return -1;
    if (!rowptr) return -2;
    if (!col) return -3;
    if (rows_cap < 2) return -5;
    if (nnz_cap < 1) return -6;
    if (tile_max < 1) return -7;
    TileCaps caps{rows_cap, nnz_cap, tile_max, domain_rows};
    TilePlan plan;
    build_tile_order(n, rowptr, col, reorder != 0, caps, plan);
    if ((int64_t)plan.order.size() != n) return -3;
    // permute the pattern and measure the halo in the new ordering
    std::vector<int> inv((size_t)n);
    for (int64_t i = 0; i < n; ++i) inv[plan.order[i]] = (int)i;
    std::vector<int64_t> rp((size_t)n + 1, 0);
    std::vector<int> cp((size_t)rowptr[n]);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t old = plan.order[i];
        int64_t d = rp[i];
        for (int64_t e = rowptr[old]; e < rowptr[old + 1]; ++e) cp[d++] = inv[col[e]];
        rp[i + 1] = d;
    }
    std::vector<uint16_t> lcol;
    if (build_tile_halo(n, rp.data(), cp.data(), plan, lcol)) return -3;
    if (order) for (int64_t i = 0; i < n; ++i) order[i] = plan.order[i];
    if (ntiles) *ntiles = (int)plan.tile_ptr.size() - 1;
    if (halo_ratio) *halo_ratio = plan.halo_ratio;
    // self-check of the layout the tiled SpMM consumes: every tile respects the capacities and every tile-local
    // column number maps back (own rows, then the halo list) to the column of the permuted pattern
    const int nt = (int)plan.tile_ptr.size() - 1;
    for (int t = 0; t < nt && plan.ok; ++t) {
        const int r0 = plan.tile_ptr[t], r1 = plan.tile_ptr[t + 1], rows = r1 - r0;
        const int h0 = plan.halo_ptr[t], nh = plan.halo_ptr[t + 1] - h0;
        if (rows < 1 || rows > tile_max || rows + nh > rows_cap) return 2;
        int64_t nnzp = 0;
        for (int i = r0; i < r1; ++i) {
            nnzp += (rp[i + 1] - rp[i] + 7) & ~(int64_t)7;
            for (int64_t e = rp[i]; e < rp[i + 1]; ++e) {
                const int lc = lcol[(size_t)e];
                const int back = lc < rows ? r0 + lc : (lc - rows < nh ? plan.halo_idx[h0 + lc - rows] : -1);
                if (back != cp[e]) return 3;
            }
        }
        if (nnzp > nnz_cap) return 2;
    }
    return plan.ok ? 0 : 1;
}

// Host-only restatement of orthonormalize() (api.cu) for CPU regression tests: the same pass logic (host_small.h
// cholqr_pass) with the Gram matrix and the block update computed on the host.  V: n x m column-major, overwritten
// with the orthonormal factor; Rtot (m x m column-major): V_in = V_out * Rtot.
extern "C" FEAST_API int feast_debug_cholqr(int64_t n, int m, feast_c128* V, int64_t ldv, feast_c128* Rtot_out, int* passes) {
    if (n < 1) return -1;
    if (m < 1 || m > n) return -2;
    if (!V) return -3;
    if (ldv < n) return -4;
    hc128* Vc = reinterpret_cast<hc128*>(V);
    std::vector<hc128> G((size_t)m * m), Ri, RD, Rtot((size_t)m * m, hc128(0, 0)), tmp, row(m);
    for (int j = 0; j < m; ++j) Rtot[(size_t)j * m + j] = 1.0;
    int pass = 0;
    double prev_err = -1.0;
    for (; pass < 10; ++pass) {
        for (int j = 0; j < m; ++j)
            for (int i = 0; i < m; ++i) {
                hc128 s(0, 0);
                for (int64_t r = 0; r < n; ++r) s += std::conj(Vc[(size_t)i * ldv + r]) * Vc[(size_t)j * ldv + r];
                G[(size_t)j * m + i] = s;
            }
        if (cholqr_pass(m, (double)n, G, Ri, RD, prev_err)) break;
        for (int64_t r = 0; r < n; ++r) {       // V_new = V * Ri, row by row
            for (int j = 0; j < m; ++j) {
                hc128 s(0, 0);
                for (int k = 0; k <= j; ++k) s += Vc[(size_t)k * ldv + r] * Ri[(size_t)j * m + k];
                row[j] = s;
            }
            for (int j = 0; j < m; ++j) Vc[(size_t)j * ldv + r] = row[j];
        }
        matmul_small(m, RD, Rtot, tmp);
        Rtot.swap(tmp);
    }
    if (Rtot_out) for (size_t t = 0; t < (size_t)m * m; ++t) { Rtot_out[t].re = Rtot[t].real(); Rtot_out[t].im = Rtot[t].imag(); }
    if (passes) *passes = pass;
    return 0;
}
