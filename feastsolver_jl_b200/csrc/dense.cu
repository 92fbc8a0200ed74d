// K2 / K3: dense complex128 LU with partial pivoting (recursive, GEMM-rich) and the
// [storage: the factored matrix is ROW-MAJOR, Z(i,j) = Z[i*ld + j], so that row interchanges move
//  contiguous segments (the column-major laswp was 30 % of the LU time in the first ncu launch list)]
// multi-right-hand-side triangular solves.  Replaces `factorizer(C)` = lu and
// `left_divider(temp, F, R)` = ldiv! of src/utils.jl:175-179 / src/feast.jl:30,36,62,65
// (LAPACK zgetrf / zgetrs upstream).  Same pivoting rule as zgetf2: the pivot is the
// first row maximising |re| + |im| (izamax).
//
//   rec(j0, w): factor the panel A[j0:n, j0:j0+w]
//       w <= 32 : cooperative panel kernel (slab of the panel resident in shared memory,
//                 ONE grid-wide sync per column: every CTA publishes its pivot candidate
//                 row together with the candidate value)
//       else    : rec(left half); laswp; recursive TRSM; ZGEMM trailing update; rec(right half); laswp
// so that all O(n^3) work is in ZGEMM calls whose K grows with the recursion level.
#include <cooperative_groups.h>

#include "kernels.cuh"
namespace cg = cooperative_groups;

namespace {

constexpr int PW = 32;  // max panel width handled by the cooperative kernel

struct Cand { double val; int row; int pad; };

struct PanelArgs {
    c128* A;        // top-left of the panel (row j0, col j0)
    int64_t lda;
    int rows;       // n - j0
    int w;          // panel width (<= PW)
    int R;          // rows per CTA
    int j0;
    int* ipiv;      // global pivot array (absolute, 0-based)
    int* info;      // first zero pivot column (1-based) or 0
    Cand* cand;     // [2][grid]
    c128* candrow;  // [2][grid][PW]
    c128* diagrow;  // [2][PW]
};

__global__ void __launch_bounds__(256) lu_panel_kernel(PanelArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ c128 slab[];  // [w][R]
    __shared__ c128 urow[PW];
    __shared__ double s_val[8];
    __shared__ int s_row[8];
    __shared__ int s_best, s_piv, s_pblk;
    __shared__ double s_pval;
    const int tid = threadIdx.x, G = gridDim.x, R = a.R, w = a.w;
    const int r0 = blockIdx.x * R;
    int nr = a.rows - r0;
    if (nr > R) nr = R;
    if (nr < 0) nr = 0;
    for (int idx = tid; idx < nr * w; idx += 256) {
        const int k = idx % w, r = idx / w;     // consecutive threads read consecutive columns of a row
        slab[k * R + r] = a.A[(int64_t)(r0 + r) * a.lda + k];
    }
    __syncthreads();
    for (int jj = 0; jj < w; ++jj) {
        const int buf = jj & 1;
        // ---- local pivot candidate among rows >= jj ----
        double bv = -1.0;
        int br = 0x7fffffff;
        for (int r = tid; r < nr; r += 256) {
            const int gr = r0 + r;
            if (gr >= jj) {
                const double v = cabs1(slab[jj * R + r]);
                if (v > bv || (v == bv && gr < br)) { bv = v; br = gr; }
            }
        }
        for (int off = 16; off; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int orow = __shfl_xor_sync(0xffffffffu, br, off);
            if (ov > bv || (ov == bv && orow < br)) { bv = ov; br = orow; }
        }
        if ((tid & 31) == 0) { s_val[tid >> 5] = bv; s_row[tid >> 5] = br; }
        __syncthreads();
        if (tid == 0) {
            for (int q = 1; q < 8; ++q)
                if (s_val[q] > bv || (s_val[q] == bv && s_row[q] < br)) { bv = s_val[q]; br = s_row[q]; }
            Cand c; c.val = bv; c.row = br; c.pad = 0;
            a.cand[buf * G + blockIdx.x] = c;
            s_best = br;
        }
        __syncthreads();
        if (tid < w) {
            const int best = s_best;
            if (best != 0x7fffffff) a.candrow[((int64_t)buf * G + blockIdx.x) * PW + tid] = slab[tid * R + (best - r0)];
            if (jj >= r0 && jj < r0 + nr) a.diagrow[buf * PW + tid] = slab[tid * R + (jj - r0)];
        }
        __threadfence();
        grid.sync();
        // ---- global pivot (every CTA redundantly) ----
        if (tid < 32) {
            double gv = -1.0;
            int grow = 0x7fffffff, gblk = 0;
            for (int b = tid; b < G; b += 32) {
                const Cand c = a.cand[buf * G + b];
                if (c.val > gv || (c.val == gv && c.row < grow)) { gv = c.val; grow = c.row; gblk = b; }
            }
            for (int off = 16; off; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, gv, off);
                const int orow = __shfl_xor_sync(0xffffffffu, grow, off);
                const int oblk = __shfl_xor_sync(0xffffffffu, gblk, off);
                if (ov > gv || (ov == gv && orow < grow)) { gv = ov; grow = orow; gblk = oblk; }
            }
            if (tid == 0) { s_piv = grow; s_pval = gv; s_pblk = gblk; }
        }
        __syncthreads();
        const int p = s_piv;
        const double pval = s_pval;
        if (tid < w) urow[tid] = a.candrow[((int64_t)buf * G + s_pblk) * PW + tid];
        if (blockIdx.x == 0 && tid == 0) {
            a.ipiv[a.j0 + jj] = a.j0 + p;
            if (pval == 0.0 && *a.info == 0) *a.info = a.j0 + jj + 1;
        }
        __syncthreads();
        if (p != jj && tid < w) {
            if (jj >= r0 && jj < r0 + nr) slab[tid * R + (jj - r0)] = urow[tid];
            if (p >= r0 && p < r0 + nr) slab[tid * R + (p - r0)] = a.diagrow[buf * PW + tid];
        }
        __syncthreads();
        if (pval != 0.0) {
            const c128 inv = cdiv(cmake(1.0, 0.0), urow[jj]);
            for (int r = tid; r < nr; r += 256)
                if (r0 + r > jj) slab[jj * R + r] = cmul(slab[jj * R + r], inv);
            __syncthreads();
            const int nk = w - jj - 1;
            for (int idx = tid; idx < nr * nk; idx += 256) {
                const int r = idx % nr, k = jj + 1 + idx / nr;
                if (r0 + r > jj) {
                    const c128 l = slab[jj * R + r], u = urow[k];
                    c128 v = slab[k * R + r];
                    cfma(v, cmake(-l.x, -l.y), u);
                    slab[k * R + r] = v;
                }
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nr * w; idx += 256) {
        const int k = idx % w, r = idx / w;
        a.A[(int64_t)(r0 + r) * a.lda + k] = slab[k * R + r];
    }
}

// Apply the row interchanges ipiv[k0..k1) of ONE panel to every column outside the panel
// (columns [0, k0) and [k1, n)) of ROW-major A, right after the panel is factored (the LAPACK
// right-looking convention).  Consecutive threads own consecutive columns, so each interchange moves
// contiguous row segments, and a thread's sequential chain is only the panel width (<= 32) long.
__global__ void laswp_kernel(c128* __restrict__ A, int64_t lda, int n, const int* __restrict__ ipiv, int k0, int k1) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n - (k1 - k0)) return;
    if (c >= k0) c += (k1 - k0);          // skip the panel's own columns
    c128* colp = A + c;
    for (int k = k0; k < k1; ++k) {
        const int p = ipiv[k];
        if (p != k) {
            const c128 t = colp[(int64_t)k * lda];
            colp[(int64_t)k * lda] = colp[(int64_t)p * lda];
            colp[(int64_t)p * lda] = t;
        }
    }
}

// Leaf triangular solves on a <=32-row block, one thread per right-hand-side column.
// T(i,k) = Tm[i*sTi + k*sTk] (conj optional); B(i,j) = B[i*sBi + j*sBj].
//   LOWER: forward substitution, UNIT selects unit diagonal.
template <bool LOWER, bool UNIT>
__global__ void __launch_bounds__(128) trsm_leaf_kernel(int h, int ncols, const c128* __restrict__ Tm, int64_t sTi,
                                                       int64_t sTk, int conjT, c128* __restrict__ B, int64_t sBi,
                                                       int64_t sBj) {
    __shared__ c128 sT[32][33];
    for (int idx = threadIdx.x; idx < 32 * 32; idx += 128) {
        const int i = idx & 31, k = idx >> 5;
        c128 v = cmake(0.0, 0.0);
        if (i < h && k < h) {
            v = Tm[i * sTi + k * sTk];
            if (conjT) v.y = -v.y;
        }
        sT[i][k] = v;
    }
    __syncthreads();
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j >= ncols) return;
    c128 x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = (i < h) ? B[i * sBi + j * sBj] : cmake(0.0, 0.0);
    if (LOWER) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (i < h) {
                c128 v = x[i];
#pragma unroll
                for (int k = 0; k < i; ++k) { const c128 t = sT[i][k]; cfma(v, cmake(-t.x, -t.y), x[k]); }
                if (!UNIT) v = cdiv(v, sT[i][i]);
                x[i] = v;
            }
        }
    } else {
#pragma unroll
        for (int i = 31; i >= 0; --i) {
            if (i < h) {
                c128 v = x[i];
#pragma unroll
                for (int k = i + 1; k < 32; ++k) { const c128 t = sT[i][k]; cfma(v, cmake(-t.x, -t.y), x[k]); }
                if (!UNIT) v = cdiv(v, sT[i][i]);
                x[i] = v;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < h) B[i * sBi + j * sBj] = x[i];
}

// perm from the sequential interchanges: perm[i] = source row of permuted row i (single CTA)
__global__ void build_perm_kernel(int n, const int* __restrict__ ipiv, int* __restrict__ perm) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < n; ++k) {
            const int p = ipiv[k];
            if (p != k) { const int t = perm[k]; perm[k] = perm[p]; perm[p] = t; }
        }
    }
}
// dst[i,:] = src[perm[i],:]  (forward) or dst[perm[i],:] = src[i,:] (inverse); row-major n x m
__global__ void permute_rows_kernel(int64_t n, int m, const c128* __restrict__ src, c128* __restrict__ dst,
                                    const int* __restrict__ perm, int inverse) {
    const int64_t total = n * m;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / m;
        const int j = (int)(t % m);
        if (inverse) dst[(int64_t)perm[i] * m + j] = src[t];
        else dst[t] = src[(int64_t)perm[i] * m + j];
    }
}

struct LUWork {
    Cand* cand = nullptr;
    c128* candrow = nullptr;
    c128* diagrow = nullptr;
    int* info = nullptr;
};

// Triangular solve with an h x h triangle T (generic strides) applied to B (h x ncols, generic strides),
// recursive: all large work is ZGEMM.
template <bool LOWER, bool UNIT>
int trsm_rec(feast_ctx* ctx, int h, int ncols, const c128* T, int64_t sTi, int64_t sTk, bool conjT, c128* B,
             int64_t sBi, int64_t sBj) {
    if (h <= 0 || ncols <= 0) return 0;
    if (h <= 32) {
        trsm_leaf_kernel<LOWER, UNIT><<<ceil_div(ncols, 128), 128, 0, ctx->stream>>>(h, ncols, T, sTi, sTk, conjT ? 1 : 0,
                                                                                     B, sBi, sBj);
        KLAUNCH_CHECK(ctx);
        return 0;
    }
    int h1 = ((h / 2 + 31) / 32) * 32;
    if (h1 >= h) h1 = h - 32;
    const int h2 = h - h1;
    if (LOWER) {
        FEAST_TRY((trsm_rec<LOWER, UNIT>(ctx, h1, ncols, T, sTi, sTk, conjT, B, sBi, sBj)));
        // B2 -= T21 * B1
        FEAST_TRY(launch_zgemm(ctx, h2, ncols, h1, hc128(-1, 0), T + h1 * sTi, sTi, sTk, conjT, B, sBi, sBj, hc128(1, 0),
                               B + h1 * sBi, sBi, sBj));
        FEAST_TRY((trsm_rec<LOWER, UNIT>(ctx, h2, ncols, T + h1 * sTi + h1 * sTk, sTi, sTk, conjT, B + h1 * sBi, sBi, sBj)));
    } else {
        FEAST_TRY((trsm_rec<LOWER, UNIT>(ctx, h2, ncols, T + h1 * sTi + h1 * sTk, sTi, sTk, conjT, B + h1 * sBi, sBi, sBj)));
        // B1 -= T12 * B2
        FEAST_TRY(launch_zgemm(ctx, h1, ncols, h2, hc128(-1, 0), T + h1 * sTk, sTi, sTk, conjT, B + h1 * sBi, sBi, sBj,
                               hc128(1, 0), B, sBi, sBj));
        FEAST_TRY((trsm_rec<LOWER, UNIT>(ctx, h1, ncols, T, sTi, sTk, conjT, B, sBi, sBj)));
    }
    return 0;
}

int lu_panel(feast_ctx* ctx, int64_t n, c128* Z, int64_t lda, int j0, int w, int* ipiv, LUWork& wk, int64_t ncols) {
    PanelArgs a;
    a.A = Z + (int64_t)j0 * lda + j0;   // row-major: (row j0, col j0)
    a.lda = lda;
    a.rows = (int)(n - j0);
    a.w = w;
    a.j0 = j0;
    a.ipiv = ipiv;
    a.info = wk.info;
    a.cand = wk.cand;
    a.candrow = wk.candrow;
    a.diagrow = wk.diagrow;
    int G = ceil_div(a.rows, 64);
    if (G > kNumSMs) G = kNumSMs;
    if (G < 1) G = 1;
    a.R = ceil_div(a.rows, G);
    G = ceil_div(a.rows, a.R);
    const size_t smem = sizeof(c128) * (size_t)a.R * w;
    if (smem > 200 * 1024) return feast_fail(ctx, FEAST_ERR_STATE, "dense LU panel exceeds shared memory (n too large)");
    if (!ctx->panel_attr_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(lu_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->panel_attr_set = true;
    }
    void* args[] = {&a};
    CUDA_TRY(ctx, cudaLaunchCooperativeKernel((void*)lu_panel_kernel, dim3(G), dim3(256), args, smem, ctx->stream));
    ctx->launches++;
    if (ncols > w) {   // ncols: columns of the (possibly rectangular) row-major window the interchanges apply to
        laswp_kernel<<<ceil_div(ncols - w, 256), 256, 0, ctx->stream>>>(Z, lda, (int)ncols, ipiv, j0, j0 + w);
        KLAUNCH_CHECK(ctx);
    }
    return 0;
}

int lu_rec(feast_ctx* ctx, int64_t n, c128* Z, int64_t lda, int j0, int w, int* ipiv, LUWork& wk, int64_t ncols) {
    if (w <= PW) return lu_panel(ctx, n, Z, lda, j0, w, ipiv, wk, ncols);
    int h = ((w / 2 + 31) / 32) * 32;
    if (h >= w) h = w - 32;
    const int w2 = w - h;
    FEAST_TRY(lu_rec(ctx, n, Z, lda, j0, h, ipiv, wk, ncols));
    // right part (its rows were already interchanged panel by panel): U12 = L11^-1 A12, A22 -= L21 U12
    c128* A11 = Z + (int64_t)j0 * lda + j0;          // row-major blocks
    c128* A12 = A11 + h;
    c128* A21 = Z + (int64_t)(j0 + h) * lda + j0;
    c128* A22 = A21 + h;
    FEAST_TRY((trsm_rec<true, true>(ctx, h, w2, A11, lda, 1, false, A12, lda, 1)));
    const int mrows = (int)(n - j0 - h);
    FEAST_TRY(launch_zgemm(ctx, mrows, w2, h, hc128(-1, 0), A21, lda, 1, false, A12, lda, 1, hc128(1, 0), A22, lda, 1));
    FEAST_TRY(lu_rec(ctx, n, Z, lda, j0 + h, w2, ipiv, wk, ncols));
    return 0;
}

}  // namespace

int dense_getrf(feast_ctx* ctx, int64_t n, c128* Z, int* ipiv_d, int* info_out) {
    LUWork wk;
    // scratch: carve from red_d (needs 2*148*(16 + 32*16) + 2*32*16 + 4 bytes ~ 160 KB)
    char* base = (char*)ctx->red_d;
    wk.cand = (Cand*)base;                        base += sizeof(Cand) * 2 * kNumSMs;
    wk.candrow = (c128*)base;                     base += sizeof(c128) * 2 * kNumSMs * PW;
    wk.diagrow = (c128*)base;                     base += sizeof(c128) * 2 * PW;
    wk.info = (int*)base;
    CUDA_TRY(ctx, cudaMemsetAsync(wk.info, 0, sizeof(int), ctx->stream));
    // NOTE: the split-K path of launch_zgemm also uses red_d; LU GEMMs have beta = 1 so never split.
    FEAST_TRY(lu_rec(ctx, n, Z, n, 0, (int)n, ipiv_d, wk, n));
    if (info_out) {
        int* h = (int*)ctx->pinned;
        CUDA_TRY(ctx, cudaMemcpyAsync(h, wk.info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        *info_out = *h;
    }
    return 0;
}

// LU with partial pivoting of the first w columns of a ROW-major window of `rows` rows and `ncols` columns (lda >= ncols):
// the row interchanges (ipiv, absolute rows of the window) are applied to all ncols columns, the elimination only to the
// first w (the caller finishes the remaining columns with dense_trsm + launch_zgemm).  Used by the pivoted band LU.
int dense_getrf_rect(feast_ctx* ctx, int rows, int w, int ncols, c128* Z, int64_t lda, int* ipiv_d, int* info_out) {
    LUWork wk;
    char* base = (char*)ctx->red_d;
    wk.cand = (Cand*)base;                        base += sizeof(Cand) * 2 * kNumSMs;
    wk.candrow = (c128*)base;                     base += sizeof(c128) * 2 * kNumSMs * PW;
    wk.diagrow = (c128*)base;                     base += sizeof(c128) * 2 * PW;
    wk.info = (int*)base;
    CUDA_TRY(ctx, cudaMemsetAsync(wk.info, 0, sizeof(int), ctx->stream));
    FEAST_TRY(lu_rec(ctx, rows, Z, lda, 0, w, ipiv_d, wk, ncols));
    if (info_out) {
        int* h = (int*)ctx->pinned;
        CUDA_TRY(ctx, cudaMemcpyAsync(h, wk.info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        *info_out = *h;
    }
    return 0;
}

// B (h x ncols, row-major ldb) <- T^-1 B with the lower-unit or upper-non-unit triangle of a row-major T (ldt)
int dense_trsm(feast_ctx* ctx, bool lower_unit, int h, int ncols, const c128* T, int64_t ldt, c128* B, int64_t ldb) {
    if (lower_unit) return trsm_rec<true, true>(ctx, h, ncols, T, ldt, 1, false, B, ldb, 1);
    return trsm_rec<false, false>(ctx, h, ncols, T, ldt, 1, false, B, ldb, 1);
}

int dense_build_perm(feast_ctx* ctx, int64_t n, const int* ipiv_d, int* perm_d) {
    build_perm_kernel<<<1, 256, 0, ctx->stream>>>((int)n, ipiv_d, perm_d);
    KLAUNCH_CHECK(ctx);
    return 0;
}

namespace {
__global__ void set_identity_blocks_kernel(int64_t n, int nb, c128* __restrict__ D) {
    // D is n x nb (row-major); block k = rows [k*nb, (k+1)*nb): identity on its own diagonal
    const int64_t total = n * nb;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / nb;
        const int j = (int)(t % nb);
        D[t] = ((int)(i % nb) == j) ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
    }
}
}  // namespace

// Explicit inverses of the kDiagNB x kDiagNB diagonal blocks of L (unit lower) and U (upper), computed
// once per factorisation with the recursive TRSM on an identity right-hand side.  They turn every
// diagonal step of the triangular solves into one DMMA GEMM (the recursive solve was launch-bound:
// ~2000 tiny kernels per node solve at n = 16384).
int dense_build_diag_inverses(feast_ctx* ctx, int64_t n, const c128* LU, c128* dinv) {
    const int nb = kDiagNB;
    c128* dL = dinv;
    c128* dU = dinv + (size_t)n * nb;
    const int64_t total = 2 * n * nb;
    int64_t gq = (total + 255) / 256, cap = (int64_t)kNumSMs * 16;
    set_identity_blocks_kernel<<<(int)(gq < cap ? gq : cap), 256, 0, ctx->stream>>>(2 * n, nb, dinv);  // n % nb == 0 not required:
    KLAUNCH_CHECK(ctx);                                                                                // see the re-init below
    if (n % nb != 0) {
        // the U half starts at row n which may not be a multiple of nb: initialise it separately
        int64_t g2 = (n * nb + 255) / 256;
        set_identity_blocks_kernel<<<(int)(g2 < cap ? g2 : cap), 256, 0, ctx->stream>>>(n, nb, dU);
        KLAUNCH_CHECK(ctx);
    }
    for (int64_t k0 = 0; k0 < n; k0 += nb) {
        const int h = (int)((n - k0) < nb ? (n - k0) : nb);
        const c128* Dkk = LU + k0 * n + k0;
        FEAST_TRY((trsm_rec<true, true>(ctx, h, h, Dkk, n, 1, false, dL + k0 * nb, nb, 1)));
        FEAST_TRY((trsm_rec<false, false>(ctx, h, h, Dkk, n, 1, false, dU + k0 * nb, nb, 1)));
    }
    return 0;
}

int dense_getrs(feast_ctx* ctx, int64_t n, const c128* LU, const int* perm_d, const c128* dinv, int m, const c128* Rhs,
                c128* Y, bool conj_transpose, c128* work) {
    if (!work) work = ctx->W2.p;   // n x m scratch
    const int64_t total = n * m;
    int64_t g = (total + 255) / 256, cap = (int64_t)kNumSMs * 16;
    const int grid = (int)(g < cap ? (g < 1 ? 1 : g) : cap);
    if (!conj_transpose) {
        // Y = P * Rhs ; L w = Y ; U y = w      (A = P^T L U  ->  A^-1 = U^-1 L^-1 P)
        permute_rows_kernel<<<grid, 256, 0, ctx->stream>>>(n, m, Rhs, Y, perm_d, 0);
        KLAUNCH_CHECK(ctx);
        if (dinv && work) {
            // blocked solves: W_k = inv(L_kk) Y_k ; Y_below -= L[below,k] W_k ; then Y_k = inv(U_kk) W_k ;
            // W_above -= U[above,k] Y_k.  Two DMMA GEMMs per diagonal block, ping-ponging Y <-> W.
            const int nb = kDiagNB;
            c128* W = work;
            const c128* dL = dinv;
            const c128* dU = dinv + (size_t)n * nb;
            for (int64_t k0 = 0; k0 < n; k0 += nb) {
                const int h = (int)((n - k0) < nb ? (n - k0) : nb);
                FEAST_TRY(launch_zgemm(ctx, h, m, h, hc128(1, 0), dL + k0 * nb, nb, 1, false, Y + k0 * m, m, 1, hc128(0, 0),
                                       W + k0 * m, m, 1));
                const int below = (int)(n - k0 - h);
                if (below > 0)
                    FEAST_TRY(launch_zgemm(ctx, below, m, h, hc128(-1, 0), LU + (k0 + h) * n + k0, n, 1, false, W + k0 * m, m, 1,
                                           hc128(1, 0), Y + (k0 + h) * m, m, 1));
            }
            const int64_t nblk = (n + nb - 1) / nb;
            for (int64_t kb = nblk - 1; kb >= 0; --kb) {
                const int64_t k0 = kb * nb;
                const int h = (int)((n - k0) < nb ? (n - k0) : nb);
                FEAST_TRY(launch_zgemm(ctx, h, m, h, hc128(1, 0), dU + k0 * nb, nb, 1, false, W + k0 * m, m, 1, hc128(0, 0),
                                       Y + k0 * m, m, 1));
                if (k0 > 0)
                    FEAST_TRY(launch_zgemm(ctx, (int)k0, m, h, hc128(-1, 0), LU + k0, n, 1, false, Y + k0 * m, m, 1, hc128(1, 0),
                                           W, m, 1));
            }
            return 0;
        }
        FEAST_TRY((trsm_rec<true, true>(ctx, (int)n, m, LU, n, 1, false, Y, m, 1)));
        FEAST_TRY((trsm_rec<false, false>(ctx, (int)n, m, LU, n, 1, false, Y, m, 1)));
    } else {
        // A^H = U^H L^H P : U^H w = b (lower, non-unit, conj) ; L^H v = w (upper, unit, conj) ; y = P^T v
        c128* tmp = work;
        CUDA_TRY(ctx, cudaMemcpyAsync(tmp, Rhs, sizeof(c128) * total, cudaMemcpyDeviceToDevice, ctx->stream));
        // T := U^H : T(i,k) = conj(U(k,i)) -> strides swapped
        FEAST_TRY((trsm_rec<true, false>(ctx, (int)n, m, LU, 1, n, true, tmp, m, 1)));
        FEAST_TRY((trsm_rec<false, true>(ctx, (int)n, m, LU, 1, n, true, tmp, m, 1)));
        permute_rows_kernel<<<grid, 256, 0, ctx->stream>>>(n, m, tmp, Y, perm_d, 1);
        KLAUNCH_CHECK(ctx);
    }
    return 0;
}
