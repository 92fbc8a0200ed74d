// Device side of the smoothed-aggregation preconditioner of the Krylov inner solves (host setup: amg_setup.cpp).
// One V(1,1) cycle per preconditioned-COCG iteration, applied to all m0 columns at once:
//     y = w D^-1 r ; rc = P^T (r - Z y) ; yc = cycle(rc) ; y += P yc ; y += w D^-1 (r - Z y)
// with damped Jacobi (w = 2 / (1.1 rho + rho/30), the one-step Chebyshev weight for [rho/30, 1.1 rho]), P^T as the
// restriction and an explicit inverse of the coarsest shifted operator (one DMMA GEMM per cycle).  Pre- and
// post-smoother are the same diagonal scaling, so the cycle is a complex SYMMETRIC operator and COCG stays valid.
// Every level's shifted operator is sum_i c_i slot_i on the level's own pattern, assembled per contour node.
#include <algorithm>

#include "amg.h"
#include "kernels.cuh"

struct AmgDevLevel {
    int n = 0, nnz = 0, nc = 0, p_nnz = 0;
    int *rowptr = nullptr, *col = nullptr, *dpos = nullptr;      // level 0: aliases of the context's union pattern
    double* vals[FEAST_MAX_SLOTS] = {};                            // level >= 1: Galerkin slots
    c128* z = nullptr;                                             // level >= 1: assembled shifted operator
    c128* dinv = nullptr;                                          // 1 / diag(z)
    double omega = 0.0;
    int *p_rowptr = nullptr, *p_col = nullptr, *r_rowptr = nullptr, *r_col = nullptr;
    double *p_val = nullptr, *r_val = nullptr;
    c128 *r = nullptr, *y = nullptr, *t = nullptr;                 // level >= 1: n x m0 work blocks
};

struct AmgDev {
    std::vector<AmgDevLevel> lev;
    int nslots = 0, m_alloc = 0;
    int ncoarse = 0;
    c128 *zd = nullptr, *zinv = nullptr, *ident = nullptr, *cwork = nullptr, *dinvb = nullptr;
    int *ipiv = nullptr, *perm = nullptr;
    // The contour nodes do not change between outer iterations, so the explicit inverse of a node's coarsest operator
    // (LU + nc right-hand sides: the dominant per-node setup cost, ~50 ms at nc = 2788) is kept per node and reused by every
    // later pass.  key = the node's coefficients; dropped with the contour (feast_set_contour) or the problem.
    struct CachedInverse { c128* zinv = nullptr; hc128 coef[FEAST_MAX_SLOTS]; };
    std::vector<CachedInverse> cache;
    c128* zinv_scratch = nullptr;    // inverse of an uncached solve (node index < 0 or cache budget exhausted)
    int64_t cache_bytes = 0;
    double setup_seconds = 0.0;
    int64_t bytes = 0;
};

namespace {

template <typename T>
int up(feast_ctx* ctx, T** dst, const std::vector<T>& src, AmgDev* A) {
    const size_t bytes = sizeof(T) * (src.empty() ? 1 : src.size());
    CUDA_TRY(ctx, cudaMalloc((void**)dst, bytes));
    A->bytes += (int64_t)bytes;
    if (!src.empty()) CUDA_TRY(ctx, cudaMemcpyAsync(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
template <typename T>
int alloc(feast_ctx* ctx, T** dst, size_t count, AmgDev* A) {
    CUDA_TRY(ctx, cudaMalloc((void**)dst, sizeof(T) * (count ? count : 1)));
    A->bytes += (int64_t)(sizeof(T) * count);
    return 0;
}
template <typename T>
void fr(T*& p) { if (p) cudaFree(p); p = nullptr; }

struct ShiftArgs { const double* v[FEAST_MAX_SLOTS]; c128 c[FEAST_MAX_SLOTS]; int ns; };

__global__ void amg_shift_kernel(int nnz, ShiftArgs a, c128* __restrict__ z) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += gridDim.x * blockDim.x) {
        c128 acc = cmake(0.0, 0.0);
#pragma unroll 1
        for (int s = 0; s < a.ns; ++s) rfma(acc, __ldg(a.v[s] + e), a.c[s]);
        z[e] = acc;
    }
}
__global__ void amg_dinv_kernel(int n, const c128* __restrict__ z, const int* __restrict__ dpos, c128* __restrict__ dinv) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dinv[i] = cdiv(cmake(1.0, 0.0), z[dpos[i]]);
}
// y = w dinv r
__global__ void amg_jacobi0_kernel(int64_t total, int m, double w, const c128* __restrict__ dinv, const c128* __restrict__ r,
                                   c128* __restrict__ y) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
        y[t] = cmul(cscale(w, __ldg(dinv + t / m)), __ldg(r + t));
}
// y += w dinv (r - t)
__global__ void amg_jacobi_kernel(int64_t total, int m, double w, const c128* __restrict__ dinv, const c128* __restrict__ r,
                                  const c128* __restrict__ tz, c128* __restrict__ y) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        c128 v = y[t];
        cfma(v, cscale(w, __ldg(dinv + t / m)), csub(__ldg(r + t), __ldg(tz + t)));
        y[t] = v;
    }
}
// Y(nrows x m) (+)= S * (X1 - X2), S real CSR (rectangular); X2 may be nullptr.  One warp per row, lanes over columns.
template <bool ACC, bool DIFF>
__global__ void __launch_bounds__(256)
amg_spmm_real_kernel(int nrows, int m, const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                     const c128* __restrict__ X1, const c128* __restrict__ X2, c128* __restrict__ Y) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int row = warp; row < nrows; row += nwarps) {
        const int e0 = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
        for (int j0 = 0; j0 < m; j0 += 64) {
            const int ja = j0 + lane, jb = j0 + 32 + lane;
            c128 a0 = cmake(0.0, 0.0), a1 = cmake(0.0, 0.0);
            int e = e0;
            for (; e + 4 <= e1; e += 4) {   // four entries in flight: the gathers are L2 latency bound
                int64_t c[4];
                double v[4];
                c128 xa[4], xb[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { c[q] = __ldg(col + e + q); v[q] = __ldg(val + e + q); }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    xa[q] = ja < m ? __ldg(X1 + c[q] * m + ja) : cmake(0.0, 0.0);
                    xb[q] = jb < m ? __ldg(X1 + c[q] * m + jb) : cmake(0.0, 0.0);
                    if (DIFF) {
                        if (ja < m) xa[q] = csub(xa[q], __ldg(X2 + c[q] * m + ja));
                        if (jb < m) xb[q] = csub(xb[q], __ldg(X2 + c[q] * m + jb));
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) { rfma(a0, v[q], xa[q]); rfma(a1, v[q], xb[q]); }
            }
            for (; e < e1; ++e) {
                const int64_t c = __ldg(col + e);
                const double v = __ldg(val + e);
                if (ja < m) {
                    c128 x = __ldg(X1 + c * m + ja);
                    if (DIFF) x = csub(x, __ldg(X2 + c * m + ja));
                    rfma(a0, v, x);
                }
                if (jb < m) {
                    c128 x = __ldg(X1 + c * m + jb);
                    if (DIFF) x = csub(x, __ldg(X2 + c * m + jb));
                    rfma(a1, v, x);
                }
            }
            if (ja < m) { c128* o = Y + (int64_t)row * m + ja; *o = ACC ? cadd(*o, a0) : a0; }
            if (jb < m) { c128* o = Y + (int64_t)row * m + jb; *o = ACC ? cadd(*o, a1) : a1; }
        }
    }
}
__global__ void amg_identity_kernel(int n, c128* __restrict__ I) {
    const int64_t total = (int64_t)n * n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
        I[t] = (t / n == t % n) ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
}

int ew_grid(int64_t total) {
    int64_t g = (total + 255) / 256, cap = (int64_t)kNumSMs * 16;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}
int row_grid(int nrows) {   // one warp per row, 8 warps per CTA
    int64_t g = ((int64_t)nrows + 7) / 8, cap = (int64_t)kNumSMs * 8;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

// One V(1,1) cycle on level l.  The result is left in *out (y or t: the fused post-smoothing epilogue of the tiled
// SpMM writes y_new into the other buffer because neighbouring tiles still read y as halo).  dot_rz (level 0 only,
// optional): per-column <r, result> fused into the last kernel; *dot_done tells whether it was produced.
int cycle(feast_ctx* ctx, AmgDev* A, int l, const c128* zl, const c128* r, c128* y, c128* t, c128** out, c128* dot_rz, bool* dot_done) {
    AmgDevLevel& L = A->lev[l];
    const int m = ctx->m0;
    cudaStream_t st = ctx->stream;
    *out = y;
    if (dot_done) *dot_done = false;
    if (L.nc == 0) {   // coarsest: y = Z^-1 r with the explicit inverse (row-major nc x nc)
        return launch_zgemm(ctx, L.n, m, L.n, hc128(1, 0), A->zinv, L.n, 1, false, r, m, 1, hc128(0, 0), y, m, 1);
    }
    const int64_t total = (int64_t)L.n * m;
    AmgDevLevel& C = A->lev[l + 1];
    amg_jacobi0_kernel<<<ew_grid(total), 256, 0, st>>>(total, m, L.omega, L.dinv, r, y);
    KLAUNCH_CHECK(ctx);
    // residual t = r - Z y, restricted: fused into the tiled SpMM's epilogue on level 0
    int fused = l == 0 ? launch_spmm_epi(ctx, m, zl, y, t, 1, r, nullptr, 0.0, nullptr) : 1;
    if (fused > 1) return fused;
    if (fused == 0) {
        amg_spmm_real_kernel<false, false><<<row_grid(L.nc), 256, 0, st>>>(L.nc, m, L.r_rowptr, L.r_col, L.r_val, t, nullptr, C.r);
    } else {
        FEAST_TRY(launch_spmm(ctx, L.n, m, L.rowptr, L.col, nullptr, zl, y, m, t, m, nullptr));
        amg_spmm_real_kernel<false, true><<<row_grid(L.nc), 256, 0, st>>>(L.nc, m, L.r_rowptr, L.r_col, L.r_val, r, t, C.r);
    }
    KLAUNCH_CHECK(ctx);
    c128* yc = nullptr;
    FEAST_TRY(cycle(ctx, A, l + 1, C.z, C.r, C.y, C.t, &yc, nullptr, nullptr));
    amg_spmm_real_kernel<true, false><<<row_grid(L.n), 256, 0, st>>>(L.n, m, L.p_rowptr, L.p_col, L.p_val, yc, nullptr, y);
    KLAUNCH_CHECK(ctx);
    // post-smoothing y <- y + w D^-1 (r - Z y)
    fused = l == 0 ? launch_spmm_epi(ctx, m, zl, y, t, 2, r, L.dinv, L.omega, dot_rz) : 1;
    if (fused > 1) return fused;
    if (fused == 0) {
        *out = t;
        if (dot_done) *dot_done = dot_rz != nullptr;
        return 0;
    }
    FEAST_TRY(launch_spmm(ctx, L.n, m, L.rowptr, L.col, nullptr, zl, y, m, t, m, nullptr));
    amg_jacobi_kernel<<<ew_grid(total), 256, 0, st>>>(total, m, L.omega, L.dinv, r, t, y);
    KLAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace

void amg_free(feast_ctx* ctx) {
    AmgDev* A = ctx->amg;
    if (!A) return;
    for (size_t l = 0; l < A->lev.size(); ++l) {
        AmgDevLevel& L = A->lev[l];
        if (l > 0) { fr(L.rowptr); fr(L.col); fr(L.z); for (auto& v : L.vals) fr(v); }
        fr(L.dpos); fr(L.dinv);
        fr(L.p_rowptr); fr(L.p_col); fr(L.p_val); fr(L.r_rowptr); fr(L.r_col); fr(L.r_val);
        fr(L.r); fr(L.y); fr(L.t);
    }
    fr(A->zd); fr(A->zinv_scratch); fr(A->ident); fr(A->cwork); fr(A->dinvb); fr(A->ipiv); fr(A->perm);
    for (auto& c : A->cache) fr(c.zinv);
    delete A;
    ctx->amg = nullptr;
}

int amg_max_coarse() {
    static const int v = getenv("FEAST_AMG_COARSE") ? atoi(getenv("FEAST_AMG_COARSE")) : 4096;
    return v;
}

// Upload a hierarchy built by amg_setup_host from the natural-order union pattern.  `order` (new -> old, may be empty) is
// the row renumbering of the device layout of level 0; dpos0 = diagonal positions in the (padded) device layout.
int amg_build(feast_ctx* ctx, AmgHost& H, int nslots, const std::vector<int>& order, const std::vector<int>& dpos0, std::string* why) {
    amg_free(ctx);
    if (!H.ok) { if (why) *why = H.why; return 0; }
    AmgDev* A = new AmgDev();
    ctx->amg = A;
    A->nslots = nslots;
    A->setup_seconds = H.setup_seconds;
    A->lev.resize(H.levels.size());
    for (size_t l = 0; l < H.levels.size(); ++l) {
        AmgHostLevel& h = H.levels[l];
        AmgDevLevel& L = A->lev[l];
        L.n = h.n; L.nc = h.nc;
        L.nnz = (int)h.col.size();
        L.omega = 2.0 / (1.1 * h.rho + h.rho / 30.0);
        if (l == 0) {
            L.rowptr = ctx->u_rowptr; L.col = ctx->u_col;       // aliases (padded device layout, renumbered rows)
            FEAST_TRY(up(ctx, &L.dpos, dpos0, A));
        } else {
            FEAST_TRY(up(ctx, &L.rowptr, h.rowptr, A));
            FEAST_TRY(up(ctx, &L.col, h.col, A));
            FEAST_TRY(up(ctx, &L.dpos, h.dpos, A));
            for (int s = 0; s < nslots; ++s) FEAST_TRY(up(ctx, &L.vals[s], h.vals[s], A));
            FEAST_TRY(alloc(ctx, &L.z, (size_t)L.nnz, A));
        }
        FEAST_TRY(alloc(ctx, &L.dinv, (size_t)L.n, A));
        if (h.nc) {
            if (l == 0 && !order.empty()) {   // rows of P follow the renumbering of the device layout; R = P^T afterwards
                std::vector<int> prp(h.n + 1, 0), pci(h.p_col.size());
                std::vector<double> pv(h.p_val.size());
                for (int i = 0; i < h.n; ++i) prp[i + 1] = prp[i] + (h.p_rowptr[order[i] + 1] - h.p_rowptr[order[i]]);
                for (int i = 0; i < h.n; ++i) {
                    const int s0 = h.p_rowptr[order[i]], len = h.p_rowptr[order[i] + 1] - s0;
                    std::copy(h.p_col.begin() + s0, h.p_col.begin() + s0 + len, pci.begin() + prp[i]);
                    std::copy(h.p_val.begin() + s0, h.p_val.begin() + s0 + len, pv.begin() + prp[i]);
                }
                h.p_rowptr.swap(prp); h.p_col.swap(pci); h.p_val.swap(pv);
                amg_transpose(h.n, h.nc, h.p_rowptr, h.p_col, h.p_val, h.r_rowptr, h.r_col, h.r_val);
            }
            L.p_nnz = (int)h.p_col.size();
            FEAST_TRY(up(ctx, &L.p_rowptr, h.p_rowptr, A));
            FEAST_TRY(up(ctx, &L.p_col, h.p_col, A));
            FEAST_TRY(up(ctx, &L.p_val, h.p_val, A));
            FEAST_TRY(up(ctx, &L.r_rowptr, h.r_rowptr, A));
            FEAST_TRY(up(ctx, &L.r_col, h.r_col, A));
            FEAST_TRY(up(ctx, &L.r_val, h.r_val, A));
        }
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // the host vectors of this level go out of scope below
    }
    const int nc = A->lev.back().n;
    A->ncoarse = nc;
    FEAST_TRY(alloc(ctx, &A->zd, (size_t)nc * nc, A));
    FEAST_TRY(alloc(ctx, &A->zinv_scratch, (size_t)nc * nc, A));
    A->zinv = A->zinv_scratch;
    FEAST_TRY(alloc(ctx, &A->ident, (size_t)nc * nc, A));
    FEAST_TRY(alloc(ctx, &A->cwork, (size_t)nc * nc, A));
    FEAST_TRY(alloc(ctx, &A->dinvb, (size_t)2 * nc * kDiagNB, A));
    FEAST_TRY(alloc(ctx, &A->ipiv, (size_t)nc, A));
    FEAST_TRY(alloc(ctx, &A->perm, (size_t)nc, A));
    amg_identity_kernel<<<ew_grid((int64_t)nc * nc), 256, 0, ctx->stream>>>(nc, A->ident);
    KLAUNCH_CHECK(ctx);
    return 0;
}

// per-level work blocks for the current m0 (the V-cycle needs r, y, t on every coarse level)
int amg_ensure_blocks(feast_ctx* ctx) {
    AmgDev* A = ctx->amg;
    if (!A || A->m_alloc >= ctx->m0) return 0;
    for (size_t l = 1; l < A->lev.size(); ++l) {
        AmgDevLevel& L = A->lev[l];
        fr(L.r); fr(L.y); fr(L.t);
        const size_t cnt = (size_t)L.n * ctx->m0;
        FEAST_TRY(alloc(ctx, &L.r, cnt, A));
        FEAST_TRY(alloc(ctx, &L.y, cnt, A));
        FEAST_TRY(alloc(ctx, &L.t, cnt, A));
    }
    A->m_alloc = ctx->m0;
    return 0;
}

// Shifted operators of every level for one contour node: z_l = sum_i coef[i] slot_i, Jacobi diagonals, and the
// explicit inverse of the coarsest one (dense LU + solve against the identity).  Level 0 uses ctx->zvals (assembled
// by the caller).
int amg_assemble(feast_ctx* ctx, const hc128* coef, const c128* zvals0, int node, int* info) {
    AmgDev* A = ctx->amg;
    cudaStream_t st = ctx->stream;
    if (info) *info = 0;
    for (size_t l = 0; l < A->lev.size(); ++l) {
        AmgDevLevel& L = A->lev[l];
        const c128* z = zvals0;
        if (l > 0) {
            ShiftArgs a;
            a.ns = A->nslots;
            for (int s = 0; s < FEAST_MAX_SLOTS; ++s) {
                a.v[s] = s < A->nslots ? L.vals[s] : nullptr;
                a.c[s] = s < A->nslots ? cmake(coef[s].real(), coef[s].imag()) : cmake(0, 0);
            }
            amg_shift_kernel<<<ew_grid(L.nnz), 256, 0, st>>>(L.nnz, a, L.z);
            KLAUNCH_CHECK(ctx);
            z = L.z;
        }
        amg_dinv_kernel<<<ew_grid(L.n), 256, 0, st>>>(L.n, z, L.dpos, L.dinv);
        KLAUNCH_CHECK(ctx);
    }
    AmgDevLevel& C = A->lev.back();
    const int nc = C.n;
    // cached inverse of this node?
    static const int64_t budget = (getenv("FEAST_AMG_CACHE_GB") ? atoll(getenv("FEAST_AMG_CACHE_GB")) : 16) << 30;
    AmgDev::CachedInverse* slot = nullptr;
    if (node >= 0) {
        if ((int)A->cache.size() <= node) A->cache.resize(node + 1);
        slot = &A->cache[node];
        bool same = slot->zinv != nullptr;
        for (int s = 0; s < A->nslots && same; ++s) same = slot->coef[s] == coef[s];
        if (same) { A->zinv = slot->zinv; return 0; }
        if (!slot->zinv) {
            const int64_t bytes = (int64_t)sizeof(c128) * nc * nc;
            if (A->cache_bytes + bytes <= budget && cudaMalloc((void**)&slot->zinv, bytes) == cudaSuccess) A->cache_bytes += bytes;
            else { cudaGetLastError(); slot->zinv = nullptr; slot = nullptr; }
        }
    }
    c128* dst = slot ? slot->zinv : A->zinv_scratch;
    FEAST_TRY(launch_scatter_dense(ctx, nc, C.rowptr, C.col, C.z, A->zd));
    int inf = 0;
    FEAST_TRY(dense_getrf(ctx, nc, A->zd, A->ipiv, &inf));
    if (info) *info = inf;
    FEAST_TRY(dense_build_perm(ctx, nc, A->ipiv, A->perm));
    FEAST_TRY(dense_build_diag_inverses(ctx, nc, A->zd, A->dinvb));
    FEAST_TRY(dense_getrs(ctx, nc, A->zd, A->perm, A->dinvb, nc, A->ident, dst, false, A->cwork));
    if (slot) for (int s = 0; s < FEAST_MAX_SLOTS; ++s) slot->coef[s] = s < A->nslots ? coef[s] : hc128(0, 0);
    A->zinv = dst;
    return 0;
}

// the contour changed: the cached coarse inverses belong to the old nodes
void amg_drop_cache(feast_ctx* ctx) {
    AmgDev* A = ctx->amg;
    if (!A) return;
    for (auto& c : A->cache) fr(c.zinv);
    A->cache.clear();
    A->cache_bytes = 0;
    A->zinv = A->zinv_scratch;
}

// M^-1 r (one V-cycle on all m0 columns) with work blocks y, t (n x m0); *out = the one holding the result.  r is not
// modified.  dot_rz (optional): <r, M^-1 r> per column when the fused epilogue produced it (*dot_done).
int amg_apply(feast_ctx* ctx, const c128* zvals0, const c128* r, c128* y, c128* t, c128** out, c128* dot_rz, bool* dot_done) {
    return cycle(ctx, ctx->amg, 0, zvals0, r, y, t, out, dot_rz, dot_done);
}

int amg_info(const feast_ctx* ctx, int* nlevels, int* sizes, int cap, double* setup_seconds) {
    const AmgDev* A = ctx->amg;
    if (nlevels) *nlevels = A ? (int)A->lev.size() : 0;
    if (A && sizes) for (int l = 0; l < cap && l < (int)A->lev.size(); ++l) sizes[l] = A->lev[l].n;
    if (setup_seconds) *setup_seconds = A ? A->setup_seconds : 0.0;
    return 0;
}
