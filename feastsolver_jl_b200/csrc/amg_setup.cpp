// Host setup of the smoothed-aggregation hierarchy (see amg.h).  Plain C++ with std::thread row parallelism; runs
// once per feast_set_problem.  Algorithm (Vanek, Mandel, Brezina 1996, unfiltered strength graph):
//   aggregates = a root and its whole neighbourhood while all of them are free, leftovers join a neighbouring
//   aggregate; tentative prolongator T = normalised aggregate indicators; P = (I - 4/(3 rho) D^-1 A) T with
//   rho ~ rho(D^-1 A) from a power iteration; coarse slots = P^T slot P (Gustavson SpGEMM, one pattern for all slots).
#include "amg.h"

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <numeric>
#include <thread>

int host_threads() {
    static const int t = [] {
        const char* e = getenv("FEAST_HOST_THREADS");
        int v = e ? atoi(e) : 0;
        if (v <= 0) {
            v = (int)std::thread::hardware_concurrency();
            const char* lw = getenv("LOCAL_WORLD_SIZE");      // one process per GPU: share the host cores between the ranks
            const int ranks = lw ? atoi(lw) : 1;
            if (ranks > 1) v = v / ranks;
            if (v > 16) v = 16;
        }
        return v < 1 ? 1 : v;
    }();
    return t;
}

namespace {

// fn(thread, row_begin, row_end) over contiguous row chunks
void parallel_rows(int nrows, const std::function<void(int, int, int)>& fn) {
    int T = host_threads();
    if (nrows < 20000) T = 1;
    if (T == 1) { fn(0, 0, nrows); return; }
    std::vector<std::thread> th;
    const int chunk = (nrows + T - 1) / T;
    for (int t = 0; t < T; ++t) {
        const int r0 = t * chunk, r1 = std::min(nrows, r0 + chunk);
        if (r0 >= r1) break;
        th.emplace_back(fn, t, r0, r1);
    }
    for (auto& x : th) x.join();
}

struct MultiCSR {                       // one pattern, ns value arrays
    int nrows = 0, ncols = 0;
    std::vector<int> rp, ci;
    std::vector<std::vector<double>> v;
};

// C = L * Rt.  Exactly one side carries the ns value arrays: left_multi ? L : Rt; the other uses its v[0].
void spgemm(const MultiCSR& L, const MultiCSR& Rt, bool left_multi, int ns, MultiCSR& C) {
    const int nrows = L.nrows, ncols = Rt.ncols;
    const int T = host_threads();
    struct Part { std::vector<int> len, ci; std::vector<std::vector<double>> v; int r0 = 0, r1 = 0; };
    std::vector<Part> parts(T);
    parallel_rows(nrows, [&](int t, int r0, int r1) {
        Part& P = parts[t];
        P.r0 = r0; P.r1 = r1;
        P.len.assign(r1 - r0, 0);
        P.v.assign(ns, {});
        std::vector<int> marker(ncols, -1), rowcols;
        std::vector<double> acc((size_t)ns * ncols, 0.0);
        std::vector<int> perm;
        for (int i = r0; i < r1; ++i) {
            rowcols.clear();
            for (int e = L.rp[i]; e < L.rp[i + 1]; ++e) {
                const int k = L.ci[e];
                for (int f = Rt.rp[k]; f < Rt.rp[k + 1]; ++f) {
                    const int c = Rt.ci[f];
                    if (marker[c] != i) {
                        marker[c] = i;
                        rowcols.push_back(c);
                        for (int s = 0; s < ns; ++s) acc[(size_t)s * ncols + c] = 0.0;
                    }
                    if (left_multi) {
                        const double rv = Rt.v[0][f];
                        for (int s = 0; s < ns; ++s) acc[(size_t)s * ncols + c] += L.v[s][e] * rv;
                    } else {
                        const double lv = L.v[0][e];
                        for (int s = 0; s < ns; ++s) acc[(size_t)s * ncols + c] += lv * Rt.v[s][f];
                    }
                }
            }
            std::sort(rowcols.begin(), rowcols.end());
            P.len[i - r0] = (int)rowcols.size();
            for (int c : rowcols) {
                P.ci.push_back(c);
                for (int s = 0; s < ns; ++s) P.v[s].push_back(acc[(size_t)s * ncols + c]);
            }
        }
    });
    C.nrows = nrows; C.ncols = ncols;
    C.rp.assign(nrows + 1, 0);
    for (auto& P : parts)
        for (int i = P.r0; i < P.r1; ++i) C.rp[i + 1] = P.len[i - P.r0];
    for (int i = 0; i < nrows; ++i) C.rp[i + 1] += C.rp[i];
    C.ci.resize(C.rp[nrows]);
    C.v.assign(ns, std::vector<double>(C.rp[nrows]));
    for (auto& P : parts) {
        if (P.r1 <= P.r0) continue;
        const int off = C.rp[P.r0];
        std::copy(P.ci.begin(), P.ci.end(), C.ci.begin() + off);
        for (int s = 0; s < ns; ++s) std::copy(P.v[s].begin(), P.v[s].end(), C.v[s].begin() + off);
    }
}

// greedy aggregation on the pattern graph; returns the number of aggregates
int aggregate(int n, const std::vector<int>& rp, const std::vector<int>& ci, std::vector<int>& agg) {
    agg.assign(n, -1);
    int na = 0;
    for (int i = 0; i < n; ++i) {                       // pass 1: a root whose whole neighbourhood is free
        if (agg[i] >= 0) continue;
        bool free_nb = true;
        for (int e = rp[i]; e < rp[i + 1] && free_nb; ++e) free_nb = agg[ci[e]] < 0;
        if (!free_nb) continue;
        for (int e = rp[i]; e < rp[i + 1]; ++e) agg[ci[e]] = na;
        agg[i] = na++;
    }
    std::vector<int> snap(agg);                         // pass 2: leftovers join an aggregate of pass 1
    for (int i = 0; i < n; ++i) {
        if (snap[i] >= 0) continue;
        for (int e = rp[i]; e < rp[i + 1]; ++e)
            if (snap[ci[e]] >= 0) { agg[i] = snap[ci[e]]; break; }
    }
    for (int i = 0; i < n; ++i) {                       // pass 3: isolated remainders
        if (agg[i] >= 0) continue;
        agg[i] = na;
        for (int e = rp[i]; e < rp[i + 1]; ++e)
            if (agg[ci[e]] < 0) agg[ci[e]] = na;
        ++na;
    }
    return na;
}

double power_rho(int n, const std::vector<int>& rp, const std::vector<int>& ci, const std::vector<double>& a,
                 const std::vector<double>& dinv) {
    std::vector<double> v(n), w(n);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < n; ++i) { s = s * 6364136223846793005ull + 1442695040888963407ull; v[i] = (double)(s >> 11) / 9007199254740992.0 - 0.5; }
    double rho = 0.0;
    for (int it = 0; it < 25; ++it) {
        double vn = 0.0;
        for (int i = 0; i < n; ++i) vn += v[i] * v[i];
        vn = std::sqrt(vn);
        if (vn == 0.0) return 0.0;
        parallel_rows(n, [&](int, int r0, int r1) {
            for (int i = r0; i < r1; ++i) {
                double acc = 0.0;
                for (int e = rp[i]; e < rp[i + 1]; ++e) acc += a[e] * v[ci[e]];
                w[i] = acc * dinv[i] / vn;
            }
        });
        double wn = 0.0;
        for (int i = 0; i < n; ++i) wn += w[i] * w[i];
        rho = std::sqrt(wn);
        v.swap(w);
    }
    return rho;
}

}  // namespace

void amg_transpose(int nrows, int ncols, const std::vector<int>& rp, const std::vector<int>& ci, const std::vector<double>& v,
                   std::vector<int>& trp, std::vector<int>& tci, std::vector<double>& tv) {
    trp.assign(ncols + 1, 0);
    for (int e = 0; e < rp[nrows]; ++e) trp[ci[e] + 1]++;
    for (int c = 0; c < ncols; ++c) trp[c + 1] += trp[c];
    tci.resize(rp[nrows]);
    tv.resize(rp[nrows]);
    std::vector<int> cur(trp.begin(), trp.end() - 1);
    for (int i = 0; i < nrows; ++i)
        for (int e = rp[i]; e < rp[i + 1]; ++e) {
            const int d = cur[ci[e]]++;
            tci[d] = i;
            tv[d] = v[e];
        }
}

void amg_setup_host(int64_t n64, const int64_t* rowptr, const int* col, int nslots, const double* const* vals, int max_coarse,
                    AmgHost& out) {
    const auto t0 = std::chrono::steady_clock::now();
    out = AmgHost();
    MultiCSR cur;
    cur.nrows = cur.ncols = (int)n64;
    cur.rp.resize(n64 + 1);
    for (int64_t i = 0; i <= n64; ++i) cur.rp[i] = (int)rowptr[i];
    cur.ci.assign(col, col + rowptr[n64]);
    cur.v.resize(nslots);
    for (int s = 0; s < nslots; ++s) cur.v[s].assign(vals[s], vals[s] + rowptr[n64]);

    static const bool verbose = getenv("FEAST_AMG_VERBOSE") != nullptr;
    auto tick = [&](const char* what, int lev) {
        static auto last = std::chrono::steady_clock::now();
        const auto now = std::chrono::steady_clock::now();
        if (verbose) fprintf(stderr, "[amg] level %d %-12s %.3f s\n", lev, what, std::chrono::duration<double>(now - last).count());
        last = now;
    };
    tick("start", 0);
    for (int lev = 0; lev < 12; ++lev) {
        const int n = cur.nrows;
        AmgHostLevel L;
        L.n = n;
        L.dpos.assign(n, -1);
        std::vector<double> dinv(n, 0.0);
        for (int i = 0; i < n; ++i) {
            for (int e = cur.rp[i]; e < cur.rp[i + 1]; ++e)
                if (cur.ci[e] == i) L.dpos[i] = e;
            if (L.dpos[i] < 0 || !(cur.v[0][L.dpos[i]] > 0.0)) {
                out.why = "slot 0 has a missing or non-positive diagonal entry (level " + std::to_string(lev) + ")";
                return;
            }
            dinv[i] = 1.0 / cur.v[0][L.dpos[i]];
        }
        tick("diag", lev);
        L.rho = power_rho(n, cur.rp, cur.ci, cur.v[0], dinv);
        tick("power", lev);
        if (!(L.rho > 0.0) || !std::isfinite(L.rho)) { out.why = "power iteration failed"; return; }
        const bool coarsest = n <= max_coarse;
        MultiCSR next;
        if (!coarsest) {
            std::vector<int> agg;
            const int na = aggregate(n, cur.rp, cur.ci, agg);
            if (na > (int)(0.7 * n)) { out.why = "aggregation stalled (coarsening factor < 1.4) above the dense-solve size"; return; }
            tick("aggregate", lev);
            std::vector<int> cnt(na, 0);
            for (int i = 0; i < n; ++i) cnt[agg[i]]++;
            // P = (I - w D^-1 A) T, T(j, agg j) = cnt^-1/2
            const double w = 4.0 / (3.0 * L.rho);
            MultiCSR P;
            P.nrows = n; P.ncols = na;
            P.rp.assign(n + 1, 0);
            P.v.resize(1);
            {
                const int T = host_threads();
                struct Part { std::vector<int> len, ci; std::vector<double> v; int r0 = 0, r1 = 0; };
                std::vector<Part> parts(T);
                parallel_rows(n, [&](int t, int r0, int r1) {
                    Part& Q = parts[t];
                    Q.r0 = r0; Q.r1 = r1;
                    Q.len.assign(r1 - r0, 0);
                    std::vector<std::pair<int, double>> row;
                    for (int i = r0; i < r1; ++i) {
                        row.clear();
                        row.emplace_back(agg[i], 1.0 / std::sqrt((double)cnt[agg[i]]));
                        for (int e = cur.rp[i]; e < cur.rp[i + 1]; ++e) {
                            const int j = cur.ci[e], c = agg[j];
                            const double val = -w * dinv[i] * cur.v[0][e] / std::sqrt((double)cnt[c]);
                            bool found = false;
                            for (auto& pr : row) if (pr.first == c) { pr.second += val; found = true; break; }
                            if (!found) row.emplace_back(c, val);
                        }
                        std::sort(row.begin(), row.end());
                        Q.len[i - r0] = (int)row.size();
                        for (auto& pr : row) { Q.ci.push_back(pr.first); Q.v.push_back(pr.second); }
                    }
                });
                for (auto& Q : parts) for (int i = Q.r0; i < Q.r1; ++i) P.rp[i + 1] = Q.len[i - Q.r0];
                for (int i = 0; i < n; ++i) P.rp[i + 1] += P.rp[i];
                P.ci.resize(P.rp[n]);
                P.v[0].resize(P.rp[n]);
                for (auto& Q : parts) {
                    if (Q.r1 <= Q.r0) continue;
                    std::copy(Q.ci.begin(), Q.ci.end(), P.ci.begin() + P.rp[Q.r0]);
                    std::copy(Q.v.begin(), Q.v.end(), P.v[0].begin() + P.rp[Q.r0]);
                }
            }
            tick("prolongator", lev);
            MultiCSR R;
            R.nrows = na; R.ncols = n;
            R.v.resize(1);
            amg_transpose(n, na, P.rp, P.ci, P.v[0], R.rp, R.ci, R.v[0]);
            MultiCSR AP;
            tick("transpose", lev);
            spgemm(cur, P, true, nslots, AP);        // slot * P
            tick("A*P", lev);
            spgemm(R, AP, false, nslots, next);      // P^T (slot P)
            tick("R*(AP)", lev);
            L.nc = na;
            L.p_rowptr = std::move(P.rp); L.p_col = std::move(P.ci); L.p_val = std::move(P.v[0]);
            L.r_rowptr = std::move(R.rp); L.r_col = std::move(R.ci); L.r_val = std::move(R.v[0]);
        }
        L.rowptr = std::move(cur.rp);
        L.col = std::move(cur.ci);
        L.vals = std::move(cur.v);
        out.levels.push_back(std::move(L));
        if (coarsest) break;
        cur = std::move(next);
    }
    if (out.levels.empty() || out.levels.back().nc != 0) { out.why = "hierarchy deeper than 12 levels"; out.levels.clear(); return; }
    if (out.levels.size() < 2) { out.why = "problem already at dense-solve size"; out.levels.clear(); return; }
    out.ok = true;
    out.setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ------------------------------------------------------------------ host-only debug entries (CPU tests of the setup)
#include "../../include/feast_cuda.h"
extern "C" {

void* feast_debug_amg_build(int64_t n, const int64_t* rowptr, const int* col, int nslots, const double* vals_flat, int max_coarse,
                            int* nlevels, double* seconds) {
    if (!rowptr || !col || !vals_flat || nslots < 1 || nslots > FEAST_MAX_SLOTS || n < 1) return nullptr;
    const double* vp[FEAST_MAX_SLOTS];
    for (int s = 0; s < nslots; ++s) vp[s] = vals_flat + (size_t)s * rowptr[n];
    AmgHost* h = new AmgHost();
    amg_setup_host(n, rowptr, col, nslots, vp, max_coarse, *h);
    if (nlevels) *nlevels = h->ok ? (int)h->levels.size() : 0;
    if (seconds) *seconds = h->setup_seconds;
    return h;
}

int feast_debug_amg_level_info(const void* handle, int lev, int* n, int* nnz, int* nc, int* pnnz, double* rho) {
    const AmgHost* h = (const AmgHost*)handle;
    if (!h || lev < 0 || lev >= (int)h->levels.size()) return -2;
    const AmgHostLevel& L = h->levels[lev];
    if (n) *n = L.n;
    if (nnz) *nnz = (int)L.col.size();
    if (nc) *nc = L.nc;
    if (pnnz) *pnnz = (int)L.p_col.size();
    if (rho) *rho = L.rho;
    return 0;
}

int feast_debug_amg_level_get(const void* handle, int lev, int* rowptr, int* col, double* vals_flat, int* p_rowptr, int* p_col,
                              double* p_val) {
    const AmgHost* h = (const AmgHost*)handle;
    if (!h || lev < 0 || lev >= (int)h->levels.size()) return -2;
    const AmgHostLevel& L = h->levels[lev];
    if (rowptr) std::copy(L.rowptr.begin(), L.rowptr.end(), rowptr);
    if (col) std::copy(L.col.begin(), L.col.end(), col);
    if (vals_flat)
        for (size_t s = 0; s < L.vals.size(); ++s) std::copy(L.vals[s].begin(), L.vals[s].end(), vals_flat + s * L.col.size());
    if (L.nc) {
        if (p_rowptr) std::copy(L.p_rowptr.begin(), L.p_rowptr.end(), p_rowptr);
        if (p_col) std::copy(L.p_col.begin(), L.p_col.end(), p_col);
        if (p_val) std::copy(L.p_val.begin(), L.p_val.end(), p_val);
    }
    return 0;
}

void feast_debug_amg_free(void* handle) { delete (AmgHost*)handle; }

}  // extern "C"
