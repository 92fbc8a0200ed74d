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
                                     const c128* __restrict__ zvals, c128* __restrict__ dst, int64_t ldd) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= b) return;
    const int row = I * b + warp;
    if (row >= n) {
        if (I == J && lane == 0) dst[(size_t)warp * ldd + warp] = cmake(1.0, 0.0);
        return;
    }
    const int c0 = J * b;
    for (int e = rowptr[row] + lane; e < rowptr[row + 1]; e += 32) {
        const int c = col[e] - c0;
        if (col[e] >= 0 && c >= 0 && c < b) dst[(size_t)warp * ldd + c] = zvals[e];   // col < 0: padding entry
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

int scatter_block(feast_ctx* ctx, int b, int I, int J, const c128* zvals, c128* dst, int64_t ldd = 0) {
    if (ldd <= 0) ldd = b;
    CUDA_TRY(ctx, cudaMemset2DAsync(dst, sizeof(c128) * ldd, 0, sizeof(c128) * b, b, ctx->stream));
    scatter_block_kernel<<<ceil_div((int64_t)b * 32, 256), 256, 0, ctx->stream>>>((int)ctx->n, b, I, J, ctx->u_rowptr, ctx->u_col,
                                                                                zvals, dst, ldd);
    KLAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace

void band_free(BandFactor& F) {
    if (F.l21) cudaFree(F.l21);
    if (F.u12) cudaFree(F.u12);
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

int band_factor_pivoted(feast_ctx* ctx, const c128* zvals, BandFactor& F, int* info);
int band_solve_pivoted(feast_ctx* ctx, const BandFactor& F, int m, const c128* Rhs, c128* Y);
bool band_use_pivoted();

int band_factor(feast_ctx* ctx, const c128* zvals, BandFactor& F, int* info) {
    if (F.pivoted || (!F.lu && band_use_pivoted())) return band_factor_pivoted(ctx, zvals, F, info);
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
    if (F.pivoted) return band_solve_pivoted(ctx, F, m, Rhs, Y);
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

// =============================================================================== pivoted band LU (EXPERIMENTAL)
// Band LU with partial pivoting ACROSS adjacent block rows (the block form of zgbtrf with kl = ku = b): step I factors
// the 2b x b panel [S_I ; L_{I+1}] with row interchanges inside the 2b-row window, which fills the block (I, I+2):
//     window (2b x 3b) = [ S_I   U'_I     0        ]   -> rows 0..b-1 : [ L11\U11 | U12 (b x 2b) ]
//                        [ L_I+1 D_{I+1}  U_{I+1}  ]      rows b..2b-1: [ L21     | new [S_{I+1} U'_{I+1}] ]
// 4 b^2 of storage per block row instead of b^2, element growth bounded as in LAPACK's band LU.  Motivation: the
// unpivoted elimination above loses backward stability on the C4 operators beyond ~300 x 300 blocks (measured: 6e-17,
// 4e-13 at 400, 2e-3 at 500; profiles/r1b_c4_full_n250000.json).  The algorithm is checked in numpy (same window
// recurrence, backward error 4e-16 at 500 x 500 blocks); THIS DEVICE CODE HAS NOT RUN ON A GPU YET (the round's GPU
// budget was spent) and is therefore opt-in: FEAST_BAND_PIVOT=1.
namespace {

// each thread owns one column of a (2b x m) row-major window and applies the b sequential interchanges (forward)
__global__ void window_pivots_kernel(c128* __restrict__ Y, int m, int b, const int* __restrict__ ipiv) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    for (int k = 0; k < b; ++k) {
        const int p = ipiv[k];
        if (p != k) {
            const c128 t = Y[(size_t)k * m + j];
            Y[(size_t)k * m + j] = Y[(size_t)p * m + j];
            Y[(size_t)p * m + j] = t;
        }
    }
}

bool band_pivot_enabled() {
    static const bool v = getenv("FEAST_BAND_PIVOT") && atoi(getenv("FEAST_BAND_PIVOT")) != 0;
    return v;
}

int band_alloc_pivoted(feast_ctx* ctx, BandFactor& F) {
    if (F.lu) return 0;
    int b = ((ctx->bandwidth + 31) / 32) * 32;
    if (b < 32) b = 32;
    F.b = b;
    F.nbk = (int)((ctx->n + b - 1) / b);
    F.pivoted = true;
    const size_t nb = (size_t)F.nbk;
    if (cudaMalloc(&F.lu, sizeof(c128) * nb * b * b) != cudaSuccess || cudaMalloc(&F.piv, sizeof(int) * nb * 2 * b) != cudaSuccess ||
        cudaMalloc(&F.l21, sizeof(c128) * nb * b * b) != cudaSuccess || cudaMalloc(&F.u12, sizeof(c128) * nb * 2 * b * b) != cudaSuccess ||
        cudaMalloc(&F.dinv, sizeof(c128) * 2 * b * kDiagNB) != cudaSuccess) {   // diagonal-block inverses of the LAST block only
        cudaGetLastError();
        band_free(F);
        return feast_fail(ctx, FEAST_ERR_OOM, "out of device memory for the pivoted banded factorisation (%zu block rows of %d)", nb, b);
    }
    return 0;
}

// workspace: window 2b x 3b + carry b x 2b (factor); padded vector (nbk + 1) * b x m + b x m scratch (solve)
int band_work_pivoted(feast_ctx* ctx, int b, int nbk, int m) {
    const size_t need = (size_t)6 * b * b + (size_t)2 * b * b + (size_t)(nbk + 2) * b * (size_t)m + (size_t)2 * b * (m > b ? m : b);
    if (ctx->band_tmp && ctx->band_tmp_elems >= need) return 0;
    if (ctx->band_tmp) { cudaFree(ctx->band_tmp); ctx->band_tmp = nullptr; }
    ctx->band_tmp_elems = need;
    if (cudaMalloc(&ctx->band_tmp, sizeof(c128) * need) != cudaSuccess) {
        cudaGetLastError();
        return feast_fail(ctx, FEAST_ERR_OOM, "out of device memory for the pivoted banded solver workspace");
    }
    return 0;
}

}  // namespace

int band_factor_pivoted(feast_ctx* ctx, const c128* zvals, BandFactor& F, int* info) {
    FEAST_TRY(band_alloc_pivoted(ctx, F));
    const int b = F.b, nbk = F.nbk;
    FEAST_TRY(band_work_pivoted(ctx, b, nbk, ctx->m0 > 0 ? ctx->m0 : 1));
    c128* Wn = ctx->band_tmp;                      // window, 2b x 3b row-major (ld 3b)
    c128* carry = Wn + (size_t)6 * b * b;          // [S_I | U'_I], b x 2b row-major (ld 2b)
    const int64_t ldw = 3 * (int64_t)b;
    const size_t rowb = sizeof(c128) * (size_t)b;
    if (info) *info = 0;
    // carry_0 = [D_0 | U_0]
    FEAST_TRY(scatter_block(ctx, b, 0, 0, zvals, carry, 2 * b));
    if (nbk > 1) FEAST_TRY(scatter_block(ctx, b, 0, 1, zvals, carry + b, 2 * b));
    else CUDA_TRY(ctx, cudaMemset2DAsync(carry + b, 2 * rowb, 0, rowb, b, ctx->stream));
    for (int I = 0; I + 1 < nbk; ++I) {
        int* ipiv = F.piv + (size_t)I * 2 * b;
        // assemble the window
        CUDA_TRY(ctx, cudaMemcpy2DAsync(Wn, sizeof(c128) * ldw, carry, 2 * rowb, 2 * rowb, b, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemset2DAsync(Wn + 2 * b, sizeof(c128) * ldw, 0, rowb, b, ctx->stream));
        c128* low = Wn + (size_t)b * ldw;
        FEAST_TRY(scatter_block(ctx, b, I + 1, I, zvals, low, ldw));
        FEAST_TRY(scatter_block(ctx, b, I + 1, I + 1, zvals, low + b, ldw));
        if (I + 2 < nbk) FEAST_TRY(scatter_block(ctx, b, I + 1, I + 2, zvals, low + 2 * b, ldw));
        else CUDA_TRY(ctx, cudaMemset2DAsync(low + 2 * b, sizeof(c128) * ldw, 0, rowb, b, ctx->stream));
        // panel LU of the first b columns over 2b rows, interchanges applied to all 3b columns
        int inf = 0;
        FEAST_TRY(dense_getrf_rect(ctx, 2 * b, b, 3 * b, Wn, ldw, ipiv, &inf));
        if (inf && info && !*info) *info = I * b + inf;
        // U12 = L11^-1 W[0:b, b:3b] ;  W[b:2b, b:3b] -= L21 U12
        FEAST_TRY(dense_trsm(ctx, true, b, 2 * b, Wn, ldw, Wn + b, ldw));
        FEAST_TRY(launch_zgemm(ctx, b, 2 * b, b, hc128(-1, 0), low, ldw, 1, false, Wn + b, ldw, 1, hc128(1, 0), low + b, ldw, 1));
        // store the factors of this block row, keep the updated lower half as the next carry
        CUDA_TRY(ctx, cudaMemcpy2DAsync(F.lu + (size_t)I * b * b, rowb, Wn, sizeof(c128) * ldw, rowb, b, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpy2DAsync(F.l21 + (size_t)I * b * b, rowb, low, sizeof(c128) * ldw, rowb, b, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpy2DAsync(F.u12 + (size_t)I * 2 * b * b, 2 * rowb, Wn + b, sizeof(c128) * ldw, 2 * rowb, b, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpy2DAsync(carry, 2 * rowb, low + b, sizeof(c128) * ldw, 2 * rowb, b, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    {   // last block: square LU of the final Schur complement (first b columns of the carry)
        const int I = nbk - 1;
        c128* S = F.lu + (size_t)I * b * b;
        int* ipiv = F.piv + (size_t)I * 2 * b;
        CUDA_TRY(ctx, cudaMemcpy2DAsync(S, rowb, carry, 2 * rowb, rowb, b, cudaMemcpyDeviceToDevice, ctx->stream));
        int inf = 0;
        FEAST_TRY(dense_getrf(ctx, b, S, ipiv, &inf));
        if (inf && info && !*info) *info = I * b + inf;
        FEAST_TRY(dense_build_perm(ctx, b, ipiv, ipiv + b));
        FEAST_TRY(dense_build_diag_inverses(ctx, b, S, F.dinv));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int band_solve_pivoted(feast_ctx* ctx, const BandFactor& F, int m, const c128* Rhs, c128* Y) {
    const int b = F.b, nbk = F.nbk;
    const int64_t n = ctx->n;
    FEAST_TRY(band_work_pivoted(ctx, b, nbk, m));
    c128* y = ctx->band_tmp + (size_t)8 * b * b;                   // (nbk + 2) * b rows x m, zero padded
    c128* scratch = y + (size_t)(nbk + 2) * b * m;                 // b x max(m, b): solution of the last block
    c128* scratch2 = scratch + (size_t)b * (m > b ? m : b);        // b x max(m, b): getrs work
    CUDA_TRY(ctx, cudaMemsetAsync(y, 0, sizeof(c128) * (size_t)(nbk + 2) * b * m, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(y, Rhs, sizeof(c128) * (size_t)n * m, cudaMemcpyDeviceToDevice, ctx->stream));
    // forward: P_I, L11^-1 on the upper half of the window, L21 update of the lower half
    for (int I = 0; I + 1 < nbk; ++I) {
        c128* top = y + (size_t)I * b * m;
        c128* bot = top + (size_t)b * m;
        window_pivots_kernel<<<ceil_div(m, 128), 128, 0, ctx->stream>>>(top, m, b, F.piv + (size_t)I * 2 * b);
        KLAUNCH_CHECK(ctx);
        FEAST_TRY(dense_trsm(ctx, true, b, m, F.lu + (size_t)I * b * b, b, top, m));
        FEAST_TRY(launch_zgemm(ctx, b, m, b, hc128(-1, 0), F.l21 + (size_t)I * b * b, b, 1, false, top, m, 1, hc128(1, 0), bot, m, 1));
    }
    {   // last block
        const int I = nbk - 1;
        c128* yl = y + (size_t)I * b * m;
        FEAST_TRY(dense_getrs(ctx, b, F.lu + (size_t)I * b * b, F.piv + (size_t)I * 2 * b + b, F.dinv, m, yl, scratch, false, scratch2));
        CUDA_TRY(ctx, cudaMemcpyAsync(yl, scratch, sizeof(c128) * (size_t)b * m, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    // backward: x_I = U11^-1 (y_I - U12 [x_{I+1}; x_{I+2}])   (rows beyond the last block are zero padding)
    for (int I = nbk - 2; I >= 0; --I) {
        c128* yi = y + (size_t)I * b * m;
        FEAST_TRY(launch_zgemm(ctx, b, m, 2 * b, hc128(-1, 0), F.u12 + (size_t)I * 2 * b * b, 2 * b, 1, false, yi + (size_t)b * m, m, 1,
                               hc128(1, 0), yi, m, 1));
        FEAST_TRY(dense_trsm(ctx, false, b, m, F.lu + (size_t)I * b * b, b, yi, m));
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(Y, y, sizeof(c128) * (size_t)n * m, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

bool band_use_pivoted() { return band_pivot_enabled(); }
