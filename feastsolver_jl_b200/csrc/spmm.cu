// K4 / K1 / K9: CSR SpMM on row-major complex128 blocks, operator assembly on the
// union pattern and the fused polynomial residual.  HBM-bound kernels:
//   * a group of G lanes owns one CSR row; lane g owns columns g, g+G, ... so every
//     gathered X row is read as contiguous 16*G-byte segments (coalesced LDG.128),
//   * column index / value loads are warp-uniform broadcasts,
//   * persistent grid (multiple of 148 SMs) with a grid-stride loop over rows,
//   * optional fused <X, S X> column reduction via warp shuffles (the COCG <p, Zp>).
// Replaces Julia's CSC mul! (src/feast.jl:42,118,120; src/utils.jl:114) and the
// per-column `T(l_j) * x_j` of src/utils.jl:104-109.
#include <stdlib.h>
#include "kernels.cuh"

namespace {

template <typename VT> struct ValOps;
template <> struct ValOps<double> {
    static __device__ __forceinline__ void fma(c128& acc, double v, c128 x) { rfma(acc, v, x); }
};
template <> struct ValOps<c128> {
    static __device__ __forceinline__ void fma(c128& acc, c128 v, c128 x) { cfma(acc, v, x); }
};

constexpr int kSpmmThreads = 256;

// partials layout: [gridDim.x][2*m] doubles
template <typename VT, int G, int CPL, bool DOT, int MINB>
__global__ void __launch_bounds__(kSpmmThreads, MINB)
spmm_csr_kernel(int n, int m, const int* __restrict__ rowptr, const int* __restrict__ col,
                const VT* __restrict__ val, const c128* __restrict__ X, int ldx, c128* __restrict__ Y, int ldy,
                double* __restrict__ partials) {
    constexpr int RPW = 32 / G;  // rows per warp
    const int lane = threadIdx.x & 31;
    const int g = lane % G;
    const int sub = lane / G;
    const int warp_global = (blockIdx.x * kSpmmThreads + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kSpmmThreads) >> 5;

    c128 dacc[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) dacc[k] = cmake(0.0, 0.0);

    for (int64_t row0 = (int64_t)warp_global * RPW; row0 < n; row0 += (int64_t)nwarps * RPW) {
        const int row = (int)row0 + sub;
        if (row < n) {
            const int e0 = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
            c128 acc[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) acc[k] = cmake(0.0, 0.0);
            int e = e0;
            // 4-way unrolled: issue the index/value loads of four nonzeros, then 4*CPL gathers
            for (; e + 4 <= e1; e += 4) {
                int c0 = __ldg(col + e), c1 = __ldg(col + e + 1), c2 = __ldg(col + e + 2), c3 = __ldg(col + e + 3);
                VT v0 = __ldg(val + e), v1 = __ldg(val + e + 1), v2 = __ldg(val + e + 2), v3 = __ldg(val + e + 3);
                c128 x0[CPL], x1[CPL], x2[CPL], x3[CPL];
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const int j = g + k * G;
                    const bool ok = j < m;
                    x0[k] = ok ? __ldg(X + (int64_t)c0 * ldx + j) : cmake(0, 0);
                    x1[k] = ok ? __ldg(X + (int64_t)c1 * ldx + j) : cmake(0, 0);
                    x2[k] = ok ? __ldg(X + (int64_t)c2 * ldx + j) : cmake(0, 0);
                    x3[k] = ok ? __ldg(X + (int64_t)c3 * ldx + j) : cmake(0, 0);
                }
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    ValOps<VT>::fma(acc[k], v0, x0[k]);
                    ValOps<VT>::fma(acc[k], v1, x1[k]);
                    ValOps<VT>::fma(acc[k], v2, x2[k]);
                    ValOps<VT>::fma(acc[k], v3, x3[k]);
                }
            }
            for (; e < e1; ++e) {
                const int c0 = __ldg(col + e);
                const VT v0 = __ldg(val + e);
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const int j = g + k * G;
                    if (j < m) ValOps<VT>::fma(acc[k], v0, __ldg(X + (int64_t)c0 * ldx + j));
                }
            }
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int j = g + k * G;
                if (j < m) {
                    Y[(int64_t)row * ldy + j] = acc[k];
                    if (DOT) cfma(dacc[k], __ldg(X + (int64_t)row * ldx + j), acc[k]);
                }
            }
        }
    }
    if (DOT) {
        __shared__ double sred[kSpmmThreads / 32][2 * G * CPL];
        const int warp = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            double re = dacc[k].x, im = dacc[k].y;
#pragma unroll
            for (int off = G; off < 32; off <<= 1) {
                re += __shfl_xor_sync(0xffffffffu, re, off);
                im += __shfl_xor_sync(0xffffffffu, im, off);
            }
            if (sub == 0) {
                sred[warp][2 * (g + k * G)] = re;
                sred[warp][2 * (g + k * G) + 1] = im;
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 2 * G * CPL; t += kSpmmThreads) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kSpmmThreads / 32; ++w) s += sred[w][t];
            if ((t >> 1) < m) partials[(int64_t)blockIdx.x * 2 * m + t] = s;
        }
    }
}

// one warp per output: lanes stride over the partial blocks (a single thread walking ~600 partials
// cost 55 us per launch in the first ncu launch list)
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int nblocks, int count,
                                       double* __restrict__ out) {
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (o >= count) return;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += partials[(int64_t)b * count + o];
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) out[o] = s;
}

template <typename VT, int G, int CPL, int MINB>
int spmm_launch_minb(feast_ctx* ctx, int grid, int n, int m, const int* rowptr, const int* col, const VT* val,
                     const c128* X, int ldx, c128* Y, int ldy, c128* dot_out) {
    if (dot_out) {
        spmm_csr_kernel<VT, G, CPL, true, MINB><<<grid, kSpmmThreads, 0, ctx->stream>>>(
            n, m, rowptr, col, val, X, ldx, Y, ldy, ctx->red_d);
        KLAUNCH_CHECK(ctx);
        reduce_partials_kernel<<<ceil_div(2 * m * 32, 128), 128, 0, ctx->stream>>>(ctx->red_d, grid, 2 * m,
                                                                                   (double*)dot_out);
        KLAUNCH_CHECK(ctx);
    } else {
        spmm_csr_kernel<VT, G, CPL, false, MINB><<<grid, kSpmmThreads, 0, ctx->stream>>>(
            n, m, rowptr, col, val, X, ldx, Y, ldy, nullptr);
        KLAUNCH_CHECK(ctx);
    }
    return 0;
}

template <typename VT, int G, int CPL>
int spmm_dispatch_dot(feast_ctx* ctx, int grid, int n, int m, const int* rowptr, const int* col, const VT* val,
                      const c128* X, int ldx, c128* Y, int ldy, c128* dot_out) {
    static const int minb = getenv("FEAST_SPMM_MINB") ? atoi(getenv("FEAST_SPMM_MINB")) : 4;
    if (G == 32 && minb == 5) return spmm_launch_minb<VT, G, CPL, 5>(ctx, grid, n, m, rowptr, col, val, X, ldx, Y, ldy, dot_out);
    if (G == 32 && minb == 6) return spmm_launch_minb<VT, G, CPL, 6>(ctx, grid, n, m, rowptr, col, val, X, ldx, Y, ldy, dot_out);
    if (G == 32 && minb == 3) return spmm_launch_minb<VT, G, CPL, 3>(ctx, grid, n, m, rowptr, col, val, X, ldx, Y, ldy, dot_out);
    return spmm_launch_minb<VT, G, CPL, 4>(ctx, grid, n, m, rowptr, col, val, X, ldx, Y, ldy, dot_out);
}

template <typename VT>
int spmm_dispatch(feast_ctx* ctx, int n, int m, const int* rowptr, const int* col, const VT* val,
                  const c128* X, int ldx, c128* Y, int ldy, c128* dot_out) {
    // grid: persistent, a multiple of the SM count; 8 CTAs of 256 threads fill an SM (2048 threads)
    auto grid_for = [&](int G) {
        int64_t rows_per_block = (int64_t)(kSpmmThreads / 32) * (32 / G);
        int64_t need = (n + rows_per_block - 1) / rows_per_block;
        int64_t cap = (int64_t)kNumSMs * 8;
        int64_t g = need < cap ? need : cap;
        // (the fused-dot variant used to be capped at 4 CTAs/SM: ncu showed it latency-bound at 32 warps/SM)
        return (int)(g < 1 ? 1 : g);
    };
#define SPMM_CASE(G, CPL) return spmm_dispatch_dot<VT, G, CPL>(ctx, grid_for(G), n, m, rowptr, col, val, X, ldx, Y, ldy, dot_out)
    if (m <= 4) SPMM_CASE(4, 1);
    if (m <= 8) SPMM_CASE(8, 1);
    if (m <= 16) SPMM_CASE(16, 1);
    if (m <= 32) SPMM_CASE(32, 1);
    if (m <= 64) SPMM_CASE(32, 2);
    if (m <= 96) SPMM_CASE(32, 3);
    if (m <= 128) SPMM_CASE(32, 4);
#undef SPMM_CASE
    return feast_fail(ctx, FEAST_ERR_STATE, "spmm: column block wider than 128 must be chunked by the caller");
}

}  // namespace

size_t spmm_partials_bytes(int m) { return (size_t)kNumSMs * 8 * 2 * (size_t)(m < 128 ? 128 : m) * sizeof(double); }

int launch_spmm(feast_ctx* ctx, int64_t n, int m, const int* rowptr, const int* col, const double* rvals,
                const c128* cvals, const c128* X, int ldx, c128* Y, int ldy, c128* dot_out) {
    // column blocks wider than 128 are processed in chunks of 128 (registers hold CPL<=4 accumulators)
    for (int j0 = 0; j0 < m; j0 += 128) {
        const int mc = (m - j0) < 128 ? (m - j0) : 128;
        c128* dchunk = dot_out ? dot_out + j0 : nullptr;
        int rc = rvals ? spmm_dispatch<double>(ctx, (int)n, mc, rowptr, col, rvals, X + j0, ldx, Y + j0, ldy, dchunk)
                       : spmm_dispatch<c128>(ctx, (int)n, mc, rowptr, col, cvals, X + j0, ldx, Y + j0, ldy, dchunk);
        if (rc) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------- assembly
namespace {
struct AsmArgs {
    const double* rv[FEAST_MAX_SLOTS];
    const c128* cv[FEAST_MAX_SLOTS];
    c128 coef[FEAST_MAX_SLOTS];
    int nslots;
};

__global__ void assemble_union_kernel(int64_t unnz, AsmArgs a, c128* __restrict__ z) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < unnz; e += (int64_t)gridDim.x * blockDim.x) {
        c128 acc = cmake(0.0, 0.0);
#pragma unroll 1
        for (int i = 0; i < a.nslots; ++i) {
            if (a.rv[i]) rfma(acc, __ldg(a.rv[i] + e), a.coef[i]);
            else if (a.cv[i]) cfma(acc, __ldg(a.cv[i] + e), a.coef[i]);
        }
        z[e] = acc;
    }
}

__global__ void scatter_dense_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                     const c128* __restrict__ zvals, c128* __restrict__ Z) {
    // one warp per row; Z pre-zeroed
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    for (int e = rowptr[warp] + lane; e < rowptr[warp + 1]; e += 32) Z[(int64_t)warp * n + col[e]] = zvals[e];  // row-major
}

// R[row, j] = sum_e ( sum_i lam_j^i a_i[e] ) X[col_e, j]   -- one pass over X, Horner per column
template <int G, int CPL>
__global__ void __launch_bounds__(256)
poly_residual_kernel(int n, int m, AsmArgs a, const int* __restrict__ rowptr, const int* __restrict__ col,
                     const c128* __restrict__ lam, const c128* __restrict__ X, c128* __restrict__ R) {
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31, g = lane % G, sub = lane / G;
    const int warp_global = (blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = (gridDim.x * 256) >> 5;
    c128 lj[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) { int j = g + k * G; lj[k] = j < m ? lam[j] : cmake(0, 0); }
    for (int64_t row0 = (int64_t)warp_global * RPW; row0 < n; row0 += (int64_t)nwarps * RPW) {
        const int row = (int)row0 + sub;
        if (row >= n) continue;
        c128 acc[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) acc[k] = cmake(0, 0);
        for (int e = rowptr[row]; e < rowptr[row + 1]; ++e) {
            const int c0 = __ldg(col + e);
            c128 av[FEAST_MAX_SLOTS];
#pragma unroll
            for (int i = 0; i < FEAST_MAX_SLOTS; ++i) {
                if (i < a.nslots) av[i] = a.rv[i] ? cmake(__ldg(a.rv[i] + e), 0.0) : (a.cv[i] ? __ldg(a.cv[i] + e) : cmake(0, 0));
            }
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int j = g + k * G;
                if (j < m) {
                    c128 t = av[a.nslots - 1];
                    for (int i = a.nslots - 2; i >= 0; --i) t = cadd(cmul(t, lj[k]), av[i]);  // Horner
                    cfma(acc[k], t, __ldg(X + (int64_t)c0 * m + j));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < CPL; ++k) { int j = g + k * G; if (j < m) R[(int64_t)row * m + j] = acc[k]; }
    }
}

// fro2[j] = sum_e | sum_i lam_j^i a_i[e] |^2 ; block handles a slice of e for all j, partials [grid][m]
__global__ void __launch_bounds__(256)
poly_fro_kernel(int64_t unnz, int m, AsmArgs a, const c128* __restrict__ lam, double* __restrict__ partials) {
    extern __shared__ double sacc[];  // m doubles
    for (int j = threadIdx.x; j < m; j += blockDim.x) sacc[j] = 0.0;
    __syncthreads();
    // each warp takes columns j = warp, warp+8, ...; lanes stride over e
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t chunk = (unnz + gridDim.x - 1) / gridDim.x;
    const int64_t e0 = (int64_t)blockIdx.x * chunk, e1 = (e0 + chunk < unnz) ? e0 + chunk : unnz;
    for (int j = warp; j < m; j += 8) {
        const c128 l = lam[j];
        double s = 0.0;
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            c128 t = cmake(0, 0);
            for (int i = a.nslots - 1; i >= 0; --i) {
                c128 av = a.rv[i] ? cmake(__ldg(a.rv[i] + e), 0.0) : (a.cv[i] ? __ldg(a.cv[i] + e) : cmake(0, 0));
                t = cadd(cmul(t, l), av);
            }
            s += cabs2(t);
        }
        for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) sacc[j] = s;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x) partials[(int64_t)blockIdx.x * m + j] = sacc[j];
}
}  // namespace

static void fill_asm(AsmArgs& a, int nslots, const double* const* rv, const c128* const* cv, const hc128* coef) {
    a.nslots = nslots;
    for (int i = 0; i < FEAST_MAX_SLOTS; ++i) {
        a.rv[i] = i < nslots ? rv[i] : nullptr;
        a.cv[i] = i < nslots ? cv[i] : nullptr;
        a.coef[i] = (i < nslots && coef) ? cmake(coef[i].real(), coef[i].imag()) : cmake(0, 0);
    }
}

int launch_assemble_union(feast_ctx* ctx, int64_t unnz, int nslots, const double* const* rv, const c128* const* cv,
                          const hc128* coef, c128* zvals) {
    AsmArgs a;
    fill_asm(a, nslots, rv, cv, coef);
    int grid = ceil_div(unnz, 256);
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    if (grid < 1) grid = 1;
    assemble_union_kernel<<<grid, 256, 0, ctx->stream>>>(unnz, a, zvals);
    KLAUNCH_CHECK(ctx);
    return 0;
}

int launch_scatter_dense(feast_ctx* ctx, int64_t n, const int* rowptr, const int* col, const c128* zvals, c128* Z) {
    CUDA_TRY(ctx, cudaMemsetAsync(Z, 0, sizeof(c128) * n * n, ctx->stream));
    scatter_dense_kernel<<<ceil_div(n * 32, 256), 256, 0, ctx->stream>>>((int)n, rowptr, col, zvals, Z);
    KLAUNCH_CHECK(ctx);
    return 0;
}

int launch_poly_residual(feast_ctx* ctx, int64_t n, int m, int nslots, const int* rowptr, const int* col,
                         int64_t unnz, const double* const* rv, const c128* const* cv, const c128* lam_d,
                         const c128* X, c128* R, double* fro2_d) {
    AsmArgs a;
    fill_asm(a, nslots, rv, cv, nullptr);
    if (m > 128) return feast_fail(ctx, FEAST_ERR_STATE, "polynomial residual supports m0 <= 128");
    int64_t cap = (int64_t)kNumSMs * 8;
#define PR_CASE(G, CPL)                                                                                  \
    {                                                                                                    \
        int64_t need = (n + (8 * (32 / G)) - 1) / (8 * (32 / G));                                        \
        int grid = (int)(need < cap ? need : cap);                                                       \
        poly_residual_kernel<G, CPL><<<grid < 1 ? 1 : grid, 256, 0, ctx->stream>>>((int)n, m, a, rowptr, col, lam_d, X, R); \
    }
    if (m <= 4) PR_CASE(4, 1) else if (m <= 8) PR_CASE(8, 1) else if (m <= 16) PR_CASE(16, 1)
    else if (m <= 32) PR_CASE(32, 1) else if (m <= 64) PR_CASE(32, 2) else if (m <= 96) PR_CASE(32, 3)
    else PR_CASE(32, 4)
#undef PR_CASE
    KLAUNCH_CHECK(ctx);
    if (fro2_d) {
        int grid = ceil_div(unnz, 4096);
        if (grid > kNumSMs * 2) grid = kNumSMs * 2;
        if (grid < 1) grid = 1;
        poly_fro_kernel<<<grid, 256, sizeof(double) * m, ctx->stream>>>(unnz, m, a, lam_d, ctx->red_d);
        KLAUNCH_CHECK(ctx);
        reduce_partials_kernel<<<ceil_div(m * 32, 128), 128, 0, ctx->stream>>>(ctx->red_d, grid, m, fro2_d);
        KLAUNCH_CHECK(ctx);
    }
    return 0;
}
