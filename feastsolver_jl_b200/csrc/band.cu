// Band LU with partial pivoting across adjacent block rows for banded sparse shifted operators Z = sum_i c_i slot_i
// (the robust path for non-symmetric sparse T(z): restarted Krylov stagnates on them, measured).
// With block size b >= half bandwidth the union pattern is block tridiagonal; step I factors the 2b x b panel
// [S_I ; L_{I+1}] with row interchanges inside the 2b-row window (the block form of zgbtrf with kl = ku = b),
// which fills the block (I, I+2):
//     window (2b x 3b) = [ S_I   U'_I     0        ]   -> rows 0..b-1 : [ L11\U11 | U12 (b x 2b) ]
//                        [ L_I+1 D_{I+1}  U_{I+1}  ]      rows b..2b-1: [ L21     | new [S_{I+1} U'_{I+1}] ]
// 4 b^2 of storage per block row, element growth bounded as in LAPACK's band LU.  Every O(b^3) step is the dense
// machinery of dense.cu (cooperative panel LU + DMMA ZGEMM); the off-diagonal blocks are scattered from the sparse
// values on the fly.  Replaces sparse `lu` + `ldiv!` (UMFPACK upstream) for src/nlfeast.jl:17-28,36-61 at C4 scale.
// History: the first version (round 1) was a block Thomas elimination that pivots inside the Schur complements only;
// on the C4 operators its element growth (1.5e17 at 500 x 500 blocks) cost the DEVICE solve its backward stability
// beyond ~300 x 300 blocks (6e-17 at 300, 4e-13 at 400, 2e-3 at 500) and stalled nlfeast at n = 250 000.  This
// version measures 1.1e-16 / 1.1e-16 / 3.7e-16 at 200 / 400 / 500 blocks on a B200 (profiles/r2_round2_validate.log).
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

int band_alloc(feast_ctx* ctx, BandFactor& F) {
    if (F.lu) return 0;
    int b = ((ctx->bandwidth + 31) / 32) * 32;
    if (b < 32) b = 32;
    F.b = b;
    F.nbk = (int)((ctx->n + b - 1) / b);
    const size_t nb = (size_t)F.nbk;
    if (cudaMalloc(&F.lu, sizeof(c128) * nb * b * b) != cudaSuccess || cudaMalloc(&F.piv, sizeof(int) * nb * 2 * b) != cudaSuccess ||
        cudaMalloc(&F.l21, sizeof(c128) * nb * b * b) != cudaSuccess || cudaMalloc(&F.u12, sizeof(c128) * nb * 2 * b * b) != cudaSuccess ||
        cudaMalloc(&F.dinv, sizeof(c128) * 2 * b * kDiagNB) != cudaSuccess) {   // diagonal-block inverses of the LAST block only
        cudaGetLastError();
        band_free(F);
        return feast_fail(ctx, FEAST_ERR_OOM, "out of device memory for the banded factorisation (%zu block rows of %d)", nb, b);
    }
    return 0;
}

// workspace: window 2b x 3b + carry b x 2b (factor); padded vector (nbk + 1) * b x m + b x m scratch (solve)
int band_work(feast_ctx* ctx, int b, int nbk, int m) {
    const size_t need = (size_t)6 * b * b + (size_t)2 * b * b + (size_t)(nbk + 2) * b * (size_t)m + (size_t)2 * b * (m > b ? m : b);
    if (ctx->band_tmp && ctx->band_tmp_elems >= need) return 0;
    if (ctx->band_tmp) { cudaFree(ctx->band_tmp); ctx->band_tmp = nullptr; }
    ctx->band_tmp_elems = need;
    if (cudaMalloc(&ctx->band_tmp, sizeof(c128) * need) != cudaSuccess) {
        cudaGetLastError();
        return feast_fail(ctx, FEAST_ERR_OOM, "out of device memory for the banded solver workspace");
    }
    return 0;
}

}  // namespace

int band_factor(feast_ctx* ctx, const c128* zvals, BandFactor& F, int* info) {
    FEAST_TRY(band_alloc(ctx, F));
    const int b = F.b, nbk = F.nbk;
    FEAST_TRY(band_work(ctx, b, nbk, ctx->m0 > 0 ? ctx->m0 : 1));
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

int band_solve(feast_ctx* ctx, const BandFactor& F, int m, const c128* Rhs, c128* Y) {
    const int b = F.b, nbk = F.nbk;
    const int64_t n = ctx->n;
    FEAST_TRY(band_work(ctx, b, nbk, m));
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

// Solve with iterative refinement against the assembled sparse operator.  A step costs one SpMM and one more pair of
// sweeps with the same factors; it stops as soon as the residual no longer contracts -- which is immediately for the
// numerically singular operators of C4 at n = 250 000 (cond >> 1e16: the elimination is backward stable there, the
// forward residual is not small; profiles/r1b_c4_full_n250000.json).  relres_out is the achieved max_j ||Z y_j - b_j|| /
// ||b_j|| (the caller turns a value above the inner tolerance, or a non-finite one, into FEAST_WARN_INNER_MAXIT).
// work: n x m scratch.
int band_solve_refined(feast_ctx* ctx, const BandFactor& F, const c128* zvals, int m, const c128* Rhs, c128* Y, c128* work,
                       int* steps_out, double* relres_out) {
    const int64_t n = ctx->n, total = n * m;
    FEAST_TRY(band_solve(ctx, F, m, Rhs, Y));
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
        FEAST_TRY(band_solve(ctx, F, m, work, work));       // correction, in place
        add_inplace_kernel<<<eg, 256, 0, ctx->stream>>>(total, Y, work);
        KLAUNCH_CHECK(ctx);
        ++steps;
    }
    if (steps_out) *steps_out = steps;
    if (relres_out) *relres_out = rel;
    return 0;
}

