// Block-tridiagonal direct solver for banded sparse shifted operators Z = sum_i c_i slot_i
// (the robust path for non-symmetric sparse T(z): restarted Krylov stagnates on them, measured).
// With block size b >= half bandwidth the union pattern is block tridiagonal:
//     S_0 = D_0,   S_I = D_I - L_I S_{I-1}^{-1} U_{I-1}          (block Thomas / block LU, pivoting inside S_I)
//     forward  w_I = S_I^{-1} (rhs_I - L_I w_{I-1}),   backward  x_I = w_I - S_I^{-1} U_I x_{I+1}
// Every O(b^3) step is the dense machinery of dense.cu (cooperative panel LU + DMMA ZGEMM); the
// off-diagonal blocks are scattered from the sparse values on the fly and never stored.
// Replaces sparse `lu` + `ldiv!` (UMFPACK upstream) for src/nlfeast.jl:17-28,36-61 at C4 scale.
#include <cmath>
#include <algorithm>

#include "kernels.cuh"

namespace {

// dst (b x b row-major) = block (I, J) of the union-pattern matrix; rows past n get an identity
// diagonal when I == J (padding of the last block keeps S nonsingular)
__global__ void scatter_block_kernel(int n, int b, int I, int J, const int* __restrict__ rowptr, const int* __restrict__ col,
                                     const c128* __restrict__ zvals, c128* __restrict__ dst) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= b) return;
    const int row = I * b + warp;
    if (row >= n) {
        if (I == J && lane == 0) dst[(size_t)warp * b + warp] = cmake(1.0, 0.0);
        return;
    }
    const int c0 = J * b;
    for (int e = rowptr[row] + lane; e < rowptr[row + 1]; e += 32) {
        const int c = col[e] - c0;
        if (col[e] >= 0 && c >= 0 && c < b) dst[(size_t)warp * b + c] = zvals[e];   // col < 0: padding entry
    }
}
// dst[r, :] = (r < rows ? src[r, :] : 0) for a b x m block
__global__ void copy_pad_kernel(int b, int m, int rows, const c128* __restrict__ src, c128* __restrict__ dst) {
    const int total = b * m;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
        dst[t] = (t / m) < rows ? src[t] : cmake(0.0, 0.0);
}
// dst[r, :] = a[r, :] - s[r, :] for r < rows
__global__ void sub_rows_kernel(int m, int rows, const c128* __restrict__ a, const c128* __restrict__ s, c128* __restrict__ dst) {
    const int total = rows * m;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) dst[t] = csub(a[t], s[t]);
}

// r = b - r
__global__ void residual_inplace_kernel(int64_t total, const c128* __restrict__ b, c128* __restrict__ r) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) r[t] = csub(b[t], r[t]);
}
// y += d
__global__ void add_inplace_kernel(int64_t total, c128* __restrict__ y, const c128* __restrict__ d) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) y[t] = cadd(y[t], d[t]);
}

int scatter_block(feast_ctx* ctx, int b, int I, int J, const c128* zvals, c128* dst) {
    CUDA_TRY(ctx, cudaMemsetAsync(dst, 0, sizeof(c128) * (size_t)b * b, ctx->stream));
    scatter_block_kernel<<<ceil_div((int64_t)b * 32, 256), 256, 0, ctx->stream>>>((int)ctx->n, b, I, J, ctx->u_rowptr, ctx->u_col,
                                                                                zvals, dst);
    KLAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace

void band_free(BandFactor& F) {
    if (F.lu) cudaFree(F.lu);
    if (F.piv) cudaFree(F.piv);
    if (F.dinv) cudaFree(F.dinv);
    F = BandFactor();
}

static int band_alloc(feast_ctx* ctx, BandFactor& F) {
    if (F.lu) return 0;
    int b = ((ctx->bandwidth + 31) / 32) * 32;
    if (b < 32) b = 32;
    F.b = b;
    F.nbk = (int)((ctx->n + b - 1) / b);
    const size_t nb = (size_t)F.nbk;
    if (cudaMalloc(&F.lu, sizeof(c128) * nb * b * b) != cudaSuccess || cudaMalloc(&F.piv, sizeof(int) * nb * 2 * b) != cudaSuccess ||
        cudaMalloc(&F.dinv, sizeof(c128) * nb * 2 * b * kDiagNB) != cudaSuccess) {
        cudaGetLastError();
        band_free(F);
        return feast_fail(ctx, FEAST_ERR_OOM, "out of device memory for the banded factorisation (%zu block rows of %d)", nb, b);
    }
    return 0;
}

static int band_work(feast_ctx* ctx, int b, int m) {
    const size_t need = (size_t)5 * b * b + (size_t)4 * b * (m > b ? m : b);
    if (ctx->band_tmp && ctx->band_tmp_elems >= need) return 0;
    if (ctx->band_tmp) { cudaFree(ctx->band_tmp); ctx->band_tmp = nullptr; }
    ctx->band_tmp_elems = need;
    if (cudaMalloc(&ctx->band_tmp, sizeof(c128) * need) != cudaSuccess) {
        cudaGetLastError();
        return feast_fail(ctx, FEAST_ERR_OOM, "out of device memory for the banded solver workspace");
    }
    return 0;
}

int band_factor(feast_ctx* ctx, const c128* zvals, BandFactor& F, int* info) {
    FEAST_TRY(band_alloc(ctx, F));
    const int b = F.b;
    FEAST_TRY(band_work(ctx, b, ctx->m0));
    c128* T1 = ctx->band_tmp;                 // U_{I-1}
    c128* T2 = T1 + (size_t)b * b;            // S_{I-1}^{-1} U_{I-1}
    c128* T3 = T2 + (size_t)b * b;            // L_I
    c128* Wk = T3 + (size_t)b * b;            // getrs scratch (b x b)
    if (info) *info = 0;
    for (int I = 0; I < F.nbk; ++I) {
        c128* S = F.lu + (size_t)I * b * b;
        int* ipiv = F.piv + (size_t)I * 2 * b;
        int* perm = ipiv + b;
        c128* dinv = F.dinv + (size_t)I * 2 * b * kDiagNB;
        FEAST_TRY(scatter_block(ctx, b, I, I, zvals, S));
        if (I > 0) {
            const c128* Sp = F.lu + (size_t)(I - 1) * b * b;
            const int* permp = F.piv + (size_t)(I - 1) * 2 * b + b;
            const c128* dinvp = F.dinv + (size_t)(I - 1) * 2 * b * kDiagNB;
            FEAST_TRY(scatter_block(ctx, b, I - 1, I, zvals, T1));
            FEAST_TRY(dense_getrs(ctx, b, Sp, permp, dinvp, b, T1, T2, false, Wk));
            FEAST_TRY(scatter_block(ctx, b, I, I - 1, zvals, T3));
            FEAST_TRY(launch_zgemm(ctx, b, b, b, hc128(-1, 0), T3, b, 1, false, T2, b, 1, hc128(1, 0), S, b, 1));
        }
        int inf = 0;
        FEAST_TRY(dense_getrf(ctx, b, S, ipiv, &inf));
        if (inf && info && !*info) *info = I * b + inf;
        FEAST_TRY(dense_build_perm(ctx, b, ipiv, perm));
        FEAST_TRY(dense_build_diag_inverses(ctx, b, S, dinv));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int band_solve(feast_ctx* ctx, const BandFactor& F, const c128* zvals, int m, const c128* Rhs, c128* Y) {
    const int b = F.b, nbk = F.nbk;
    const int64_t n = ctx->n;
    FEAST_TRY(band_work(ctx, b, m));
    c128* T1 = ctx->band_tmp;                           // scattered off-diagonal block
    c128* base = T1 + (size_t)5 * b * b;
    const size_t blk = (size_t)b * (m > b ? m : b);
    c128* t_rhs = base;                                  // b x m padded right-hand side / product
    c128* t_sol = base + blk;                            // b x m solve result
    c128* t_wrk = base + 2 * blk;                        // getrs scratch
    c128* t_prev = base + 3 * blk;                       // previous block's vector (padded)
    const int eg = 148 * 4;
    // forward: w_I = S_I^{-1} (rhs_I - L_I w_{I-1}); w is stored in Y
    for (int I = 0; I < nbk; ++I) {
        const int rows = (int)((n - (int64_t)I * b) < b ? (n - (int64_t)I * b) : b);
        copy_pad_kernel<<<eg, 256, 0, ctx->stream>>>(b, m, rows, Rhs + (size_t)I * b * m, t_rhs);
        KLAUNCH_CHECK(ctx);
        if (I > 0) {
            FEAST_TRY(scatter_block(ctx, b, I, I - 1, zvals, T1));
            FEAST_TRY(launch_zgemm(ctx, b, m, b, hc128(-1, 0), T1, b, 1, false, t_prev, m, 1, hc128(1, 0), t_rhs, m, 1));
        }
        FEAST_TRY(dense_getrs(ctx, b, F.lu + (size_t)I * b * b, F.piv + (size_t)I * 2 * b + b, F.dinv + (size_t)I * 2 * b * kDiagNB, m,
                              t_rhs, t_prev, false, t_wrk));
        CUDA_TRY(ctx, cudaMemcpyAsync(Y + (size_t)I * b * m, t_prev, sizeof(c128) * (size_t)rows * m, cudaMemcpyDeviceToDevice,
                                      ctx->stream));
    }
    // backward: x_I = w_I - S_I^{-1} U_I x_{I+1}; t_prev holds x_{I+1} (padded)
    for (int I = nbk - 2; I >= 0; --I) {
        FEAST_TRY(scatter_block(ctx, b, I, I + 1, zvals, T1));
        FEAST_TRY(launch_zgemm(ctx, b, m, b, hc128(1, 0), T1, b, 1, false, t_prev, m, 1, hc128(0, 0), t_rhs, m, 1));
        FEAST_TRY(dense_getrs(ctx, b, F.lu + (size_t)I * b * b, F.piv + (size_t)I * 2 * b + b, F.dinv + (size_t)I * 2 * b * kDiagNB, m,
                              t_rhs, t_sol, false, t_wrk));
        sub_rows_kernel<<<eg, 256, 0, ctx->stream>>>(m, b, Y + (size_t)I * b * m, t_sol, Y + (size_t)I * b * m);
        KLAUNCH_CHECK(ctx);
        CUDA_TRY(ctx, cudaMemcpyAsync(t_prev, Y + (size_t)I * b * m, sizeof(c128) * (size_t)b * m, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return 0;
}

// Solve with iterative refinement against the assembled sparse operator: recovers the digits the elimination loses
// when a Schur complement S_I is ill-conditioned while the shifted operator itself is not (pivoting happens inside
// the S_I only).  A step costs one SpMM and one more pair of sweeps with the same factors; it stops as soon as the
// residual no longer contracts -- which is immediately for the numerically singular operators of C4 at n = 250 000
// (cond >> 1e16: the elimination is backward stable there, see profiles/r1b_c4_full_n250000.json).  work: n x m scratch.
int band_solve_refined(feast_ctx* ctx, const BandFactor& F, const c128* zvals, int m, const c128* Rhs, c128* Y, c128* work,
                       int* steps_out, double* relres_out) {
    const int64_t n = ctx->n, total = n * m;
    FEAST_TRY(band_solve(ctx, F, zvals, m, Rhs, Y));
    double* bn2 = (double*)(ctx->small_d + (size_t)4 * ctx->m0 * ctx->m0);
    double* rn2 = bn2 + m;
    double* h = (double*)ctx->pinned;
    FEAST_TRY(launch_colnorm2(ctx, n, m, Rhs, bn2));
    const int eg = 148 * 8;
    const int max_steps = 4;
    double prev = 0.0, rel = 0.0;
    int steps = 0;
    for (int it = 0; it <= max_steps; ++it) {
        FEAST_TRY(launch_spmm(ctx, n, m, ctx->u_rowptr, ctx->u_col, nullptr, zvals, Y, m, work, m, nullptr));
        residual_inplace_kernel<<<eg, 256, 0, ctx->stream>>>(total, Rhs, work);
        KLAUNCH_CHECK(ctx);
        FEAST_TRY(launch_colnorm2(ctx, n, m, work, rn2));
        CUDA_TRY(ctx, cudaMemcpyAsync(h, bn2, sizeof(double) * 2 * m, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        rel = 0.0;
        for (int j = 0; j < m; ++j)
            if (h[j] > 0.0) rel = std::max(rel, std::sqrt(h[m + j] / h[j]));
        if (!(rel == rel)) break;                                 // NaN: leave the plain solution
        if (rel <= 1e-14 || it == max_steps) break;
        if (it > 0 && rel > 0.25 * prev) break;                    // no longer contracting
        prev = rel;
        FEAST_TRY(band_solve(ctx, F, zvals, m, work, work));       // correction, in place
        add_inplace_kernel<<<eg, 256, 0, ctx->stream>>>(total, Y, work);
        KLAUNCH_CHECK(ctx);
        ++steps;
    }
    if (steps_out) *steps_out = steps;
    if (relres_out) *relres_out = rel;
    return 0;
}
