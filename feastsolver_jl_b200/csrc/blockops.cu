// K6-K8 building blocks on row-major n x m0 complex128 blocks: tall-skinny Gram (split-K),
// block update Y = X*M, column reductions, fused residual / accumulate epilogues, layout
// conversion, dense assembly.  Replaces qr/mul!/rmul!/broadcast statements of
// src/feast.jl:41-50,68-70,117-127 and src/utils.jl:111-116,166-171.
#include <stdlib.h>
#include "kernels.cuh"

namespace {

// ============================================================================ column reductions
// block = 256 threads laid out as (rows_per_pass = 256/GW) x GW lanes over columns
template <bool CONJ, bool NORM>
__global__ void __launch_bounds__(256)
coldot_kernel(int64_t n, int m, const c128* __restrict__ a, const c128* __restrict__ b,
              double* __restrict__ partials) {
    // thread owns column chunk c = threadIdx.x % cw (cw = min(m,256) rounded to pow2<=256), loops over rows
    extern __shared__ double sm[];  // [256][2]
    int cw = 1;
    while (cw < m && cw < 256) cw <<= 1;
    const int rpp = 256 / cw;
    const int cj = threadIdx.x % cw, rr = threadIdx.x / cw;
    for (int jbase = 0; jbase < m; jbase += cw) {
        const int j = jbase + cj;
        double re = 0.0, im = 0.0;
        if (j < m) {
            for (int64_t i = (int64_t)blockIdx.x * rpp + rr; i < n; i += (int64_t)gridDim.x * rpp) {
                const c128 x = __ldg(a + i * m + j);
                if (NORM) { re = fma(x.x, x.x, re); re = fma(x.y, x.y, re); }
                else {
                    const c128 y = __ldg(b + i * m + j);
                    if (CONJ) { re = fma(x.x, y.x, re); re = fma(x.y, y.y, re); im = fma(x.x, y.y, im); im = fma(-x.y, y.x, im); }
                    else      { re = fma(x.x, y.x, re); re = fma(-x.y, y.y, re); im = fma(x.x, y.y, im); im = fma(x.y, y.x, im); }
                }
            }
        }
        sm[2 * threadIdx.x] = re; sm[2 * threadIdx.x + 1] = im;
        __syncthreads();
        if (rr == 0 && j < m) {
            for (int r = 1; r < rpp; ++r) { re += sm[2 * (r * cw + cj)]; im += sm[2 * (r * cw + cj) + 1]; }
            if (NORM) partials[(int64_t)blockIdx.x * m + j] = re;
            else { partials[(int64_t)blockIdx.x * 2 * m + 2 * j] = re; partials[(int64_t)blockIdx.x * 2 * m + 2 * j + 1] = im; }
        }
        __syncthreads();
    }
}

__global__ void reduce_partials2_kernel(const double* __restrict__ partials, int nblocks, int count,
                                        double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partials[(int64_t)b * count + t];
    out[t] = s;
}

// ============================================================================ elementwise
__global__ void colscale_kernel(int64_t total, int m, c128* __restrict__ a, const c128* __restrict__ s) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
        a[t] = cmul(a[t], __ldg(s + (t % m)));
}
__global__ void colnormalize_kernel(int64_t total, int m, c128* __restrict__ a, const double* __restrict__ nrm2) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const double q = __ldg(nrm2 + (t % m));
        if (q > 0.0) { const double s = 1.0 / sqrt(q); c128 v = a[t]; a[t] = cmake(v.x * s, v.y * s); }
    }
}
__global__ void residual_combine_kernel(int64_t total, int m, c128* __restrict__ R, const c128* __restrict__ BX,
                                        const c128* __restrict__ lam) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const c128 l = __ldg(lam + (t % m));
        R[t] = csub(R[t], cmul(l, __ldg(BX + t)));
    }
}
__global__ void accumulate_kernel(int64_t total, int m, const c128* __restrict__ X, const c128* __restrict__ Y,
                                  const c128* __restrict__ d, c128* __restrict__ Q, c128* __restrict__ Q1, c128 z,
                                  int first_pass) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const c128 dj = __ldg(d + (t % m));
        c128 term = first_pass ? cmul(__ldg(Y + t), dj) : cmul(csub(__ldg(X + t), __ldg(Y + t)), dj);
        Q[t] = cadd(Q[t], term);
        if (Q1) Q1[t] = cadd(Q1[t], cmul(z, term));
    }
}
// column-slice / moment form: Y is a COMPACT n x mloc block holding the solutions of columns j0 .. j0+mloc-1;
// Q_p[:, j0+jj] += z^p * term for p < nmom  (moment accumulators S_p = sum_k w_k z_k^p (...), src/beyn.jl:19-20,52-54)
struct MomentPtrs { c128* q[FEAST_MAX_MOMENTS]; int nmom; };
__global__ void accumulate_slice_kernel(int64_t n, int m, int j0, int mloc, const c128* __restrict__ X, const c128* __restrict__ Y,
                                        const c128* __restrict__ d, MomentPtrs mp, c128 z, int first_pass) {
    const int64_t total = n * mloc;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / mloc;
        const int j = j0 + (int)(t - i * mloc);
        const int64_t g = i * m + j;
        const c128 dj = __ldg(d + j);
        c128 term = first_pass ? cmul(__ldg(Y + t), dj) : cmul(csub(__ldg(X + g), __ldg(Y + t)), dj);
#pragma unroll 1
        for (int p = 0; p < mp.nmom; ++p) {
            mp.q[p][g] = cadd(mp.q[p][g], term);
            term = cmul(z, term);
        }
    }
}
// dst (compact n x mloc) = src[:, j0 : j0 + mloc] of a row-major n x m block
__global__ void gather_cols_kernel(int64_t n, int m, int j0, int mloc, const c128* __restrict__ src, c128* __restrict__ dst) {
    const int64_t total = n * mloc;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / mloc;
        dst[t] = __ldg(src + i * m + j0 + (t - i * mloc));
    }
}
__global__ void add_kernel(int64_t count, const c128* __restrict__ a, c128* __restrict__ y) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) y[t] = cadd(y[t], __ldg(a + t));
}
__global__ void conj_kernel(int64_t count, const c128* __restrict__ s, c128* __restrict__ d) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        const c128 v = s[t];
        d[t] = cmake(v.x, -v.y);
    }
}
__global__ void real_to_complex_kernel(int64_t count, const double* __restrict__ s, c128* __restrict__ d) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x)
        d[t] = cmake(s[t], 0.0);
}

// 32x32 tiled transpose between column-major (ld) and row-major (m contiguous)
template <bool TO_ROWMAJOR>
__global__ void transpose_kernel(int64_t n, int m, const c128* __restrict__ src, c128* __restrict__ dst, int64_t ld,
                                 const int* __restrict__ perm) {
    __shared__ c128 tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    if (TO_ROWMAJOR) {
        // read column-major: consecutive threads along i
        for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
            const int64_t i = i0 + threadIdx.x; const int j = j0 + jj;
            if (i < n && j < m) tile[jj][threadIdx.x] = src[(int64_t)j * ld + (perm ? perm[i] : i)];
        }
        __syncthreads();
        for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
            const int64_t i = i0 + ii; const int j = j0 + threadIdx.x;
            if (i < n && j < m) dst[i * m + j] = tile[threadIdx.x][ii];
        }
    } else {
        for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
            const int64_t i = i0 + ii; const int j = j0 + threadIdx.x;
            if (i < n && j < m) tile[threadIdx.x][ii] = src[i * m + j];
        }
        __syncthreads();
        for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
            const int64_t i = i0 + threadIdx.x; const int j = j0 + jj;
            if (i < n && j < m) dst[(int64_t)j * ld + (perm ? perm[i] : i)] = tile[jj][threadIdx.x];
        }
    }
}

struct DenseAsmArgs {
    const c128* D[FEAST_MAX_SLOTS];
    int kind[FEAST_MAX_SLOTS];
    c128 coef[FEAST_MAX_SLOTS];
    int nslots;
};
__global__ void assemble_dense_kernel(int64_t n, DenseAsmArgs a, c128* __restrict__ Z) {
    const int64_t total = n * n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t % n, j = t / n;
        c128 acc = cmake(0.0, 0.0);
        for (int s = 0; s < a.nslots; ++s) {
            if (a.kind[s] == OP_DENSE) cfma(acc, __ldg(a.D[s] + t), a.coef[s]);
            else if (a.kind[s] == OP_IDENTITY && i == j) acc = cadd(acc, a.coef[s]);
        }
        Z[t] = acc;
    }
}
// fro2[j] for dense polynomial: grid (chunks, m)
__global__ void __launch_bounds__(256)
poly_fro_dense_kernel(int64_t n, int m, DenseAsmArgs a, const c128* __restrict__ lam, double* __restrict__ partials) {
    const int j = blockIdx.y;
    const c128 l = lam[j];
    const int64_t total = n * n;
    double s = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t % n, jj = t / n;
        c128 acc = cmake(0, 0);
        for (int q = a.nslots - 1; q >= 0; --q) {
            c128 v = cmake(0, 0);
            if (a.kind[q] == OP_DENSE) v = __ldg(a.D[q] + t);
            else if (a.kind[q] == OP_IDENTITY && i == jj) v = cmake(1.0, 0.0);
            acc = cadd(cmul(acc, l), v);
        }
        s += cabs2(acc);
    }
    __shared__ double sw[8];
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += sw[w];
        partials[(int64_t)blockIdx.x * m + j] = tot;
    }
}

int ew_grid(int64_t total) {
    int64_t g = (total + 255) / 256;
    int64_t cap = (int64_t)kNumSMs * 16;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}
}  // namespace

// ------------------------------------------------------------------------------- launchers
int launch_gram(feast_ctx* ctx, int64_t n, int m, const c128* A, const c128* B, c128* G_d) {
    // G(i,j) = sum_k conj(A[k,i]) B[k,j];  A(i,k) := A_rm[k*m + i]
    return launch_zgemm(ctx, m, m, n, hc128(1, 0), A, 1, m, true, B, m, 1, hc128(0, 0), G_d, 1, m);
}

int launch_update(feast_ctx* ctx, int64_t n, int m, const c128* X, const c128* M_d, c128* Y) {
    // Y[i,j] = sum_k X[i,k] M[k,j];  M column-major (k + j*m)
    return launch_zgemm(ctx, (int)n, m, m, hc128(1, 0), X, m, 1, false, M_d, 1, m, hc128(0, 0), Y, m, 1);
}

static int reduce_grid(int64_t n, int m) {
    int cw = 1;
    while (cw < m && cw < 256) cw <<= 1;
    int rpp = 256 / cw;
    int64_t need = (n + rpp - 1) / rpp;
    int64_t cap = (int64_t)kNumSMs * 4;
    return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

int launch_coldot(feast_ctx* ctx, int64_t n, int m, const c128* a, const c128* b, bool conj_a, c128* out_d) {
    const int grid = reduce_grid(n, m);
    if (conj_a) coldot_kernel<true, false><<<grid, 256, 512 * sizeof(double), ctx->stream>>>(n, m, a, b, ctx->red_d);
    else coldot_kernel<false, false><<<grid, 256, 512 * sizeof(double), ctx->stream>>>(n, m, a, b, ctx->red_d);
    KLAUNCH_CHECK(ctx);
    reduce_partials2_kernel<<<ceil_div(2 * m, 128), 128, 0, ctx->stream>>>(ctx->red_d, grid, 2 * m, (double*)out_d);
    KLAUNCH_CHECK(ctx);
    return 0;
}

int launch_colnorm2(feast_ctx* ctx, int64_t n, int m, const c128* a, double* out_d) {
    const int grid = reduce_grid(n, m);
    coldot_kernel<true, true><<<grid, 256, 512 * sizeof(double), ctx->stream>>>(n, m, a, a, ctx->red_d);
    KLAUNCH_CHECK(ctx);
    reduce_partials2_kernel<<<ceil_div(m, 128), 128, 0, ctx->stream>>>(ctx->red_d, grid, m, out_d);
    KLAUNCH_CHECK(ctx);
    return 0;
}

int launch_colscale(feast_ctx* ctx, int64_t n, int m, c128* a, const c128* s_d) {
    colscale_kernel<<<ew_grid(n * m), 256, 0, ctx->stream>>>(n * m, m, a, s_d);
    KLAUNCH_CHECK(ctx);
    return 0;
}
int launch_colnormalize(feast_ctx* ctx, int64_t n, int m, c128* a, const double* nrm2_d) {
    colnormalize_kernel<<<ew_grid(n * m), 256, 0, ctx->stream>>>(n * m, m, a, nrm2_d);
    KLAUNCH_CHECK(ctx);
    return 0;
}
int launch_residual_combine(feast_ctx* ctx, int64_t n, int m, c128* R, const c128* BX, const c128* lam_d) {
    residual_combine_kernel<<<ew_grid(n * m), 256, 0, ctx->stream>>>(n * m, m, R, BX, lam_d);
    KLAUNCH_CHECK(ctx);
    return 0;
}
int launch_accumulate(feast_ctx* ctx, int64_t n, int m, const c128* X, const c128* Y, const c128* d_d, c128* Q,
                      c128* Q1, hc128 z, bool first_pass) {
    accumulate_kernel<<<ew_grid(n * m), 256, 0, ctx->stream>>>(n * m, m, X, Y, d_d, Q, Q1, cmake(z.real(), z.imag()),
                                                              first_pass ? 1 : 0);
    KLAUNCH_CHECK(ctx);
    return 0;
}
int launch_accumulate_slice(feast_ctx* ctx, int64_t n, int m, int j0, int mloc, const c128* X, const c128* Y, const c128* d_d,
                            c128* const* moments, int nmom, hc128 z, bool first_pass) {
    if (nmom < 1 || nmom > FEAST_MAX_MOMENTS) return feast_fail(ctx, FEAST_ERR_STATE, "between 1 and %d moment accumulators", FEAST_MAX_MOMENTS);
    MomentPtrs mp;
    mp.nmom = nmom;
    for (int p = 0; p < FEAST_MAX_MOMENTS; ++p) mp.q[p] = p < nmom ? moments[p] : nullptr;
    accumulate_slice_kernel<<<ew_grid(n * mloc), 256, 0, ctx->stream>>>(n, m, j0, mloc, X, Y, d_d, mp, cmake(z.real(), z.imag()),
                                                                       first_pass ? 1 : 0);
    KLAUNCH_CHECK(ctx);
    return 0;
}
int launch_axpy(feast_ctx* ctx, int64_t count, const c128* a, c128* y) {   // y += a
    add_kernel<<<ew_grid(count), 256, 0, ctx->stream>>>(count, a, y);
    KLAUNCH_CHECK(ctx);
    return 0;
}
int launch_gather_cols(feast_ctx* ctx, int64_t n, int m, int j0, int mloc, const c128* src, c128* dst) {
    gather_cols_kernel<<<ew_grid(n * mloc), 256, 0, ctx->stream>>>(n, m, j0, mloc, src, dst);
    KLAUNCH_CHECK(ctx);
    return 0;
}
int launch_colmajor_to_rowmajor(feast_ctx* ctx, int64_t n, int m, const c128* src, int64_t ld, c128* dst, const int* perm) {
    dim3 grid(ceil_div(n, 32), ceil_div(m, 32)), block(32, 8);
    transpose_kernel<true><<<grid, block, 0, ctx->stream>>>(n, m, src, dst, ld, perm);
    KLAUNCH_CHECK(ctx);
    return 0;
}
int launch_rowmajor_to_colmajor(feast_ctx* ctx, int64_t n, int m, const c128* src, c128* dst, int64_t ld, const int* perm) {
    dim3 grid(ceil_div(n, 32), ceil_div(m, 32)), block(32, 8);
    transpose_kernel<false><<<grid, block, 0, ctx->stream>>>(n, m, src, dst, ld, perm);
    KLAUNCH_CHECK(ctx);
    return 0;
}
int launch_conj(feast_ctx* ctx, int64_t count, const c128* src, c128* dst) {
    conj_kernel<<<ew_grid(count), 256, 0, ctx->stream>>>(count, src, dst);
    KLAUNCH_CHECK(ctx);
    return 0;
}
int launch_real_to_complex(feast_ctx* ctx, int64_t count, const double* src, c128* dst) {
    real_to_complex_kernel<<<ew_grid(count), 256, 0, ctx->stream>>>(count, src, dst);
    KLAUNCH_CHECK(ctx);
    return 0;
}
static void fill_dense_args(DenseAsmArgs& a, int nslots, const c128* const* D, const int* kinds, const hc128* coef) {
    a.nslots = nslots;
    for (int i = 0; i < FEAST_MAX_SLOTS; ++i) {
        a.D[i] = i < nslots ? D[i] : nullptr;
        a.kind[i] = i < nslots ? kinds[i] : OP_NONE;
        a.coef[i] = (i < nslots && coef) ? cmake(coef[i].real(), coef[i].imag()) : cmake(0, 0);
    }
}
int launch_assemble_dense(feast_ctx* ctx, int64_t n, int nslots, const c128* const* D, const int* kinds,
                          const hc128* coef, c128* Z) {
    DenseAsmArgs a;
    fill_dense_args(a, nslots, D, kinds, coef);
    assemble_dense_kernel<<<ew_grid(n * n), 256, 0, ctx->stream>>>(n, a, Z);
    KLAUNCH_CHECK(ctx);
    return 0;
}
int launch_poly_fro_dense(feast_ctx* ctx, int64_t n, int m, int nslots, const c128* const* D, const int* kinds,
                          const c128* lam_d, double* fro2_d) {
    DenseAsmArgs a;
    fill_dense_args(a, nslots, D, kinds, nullptr);
    int gx = ceil_div(n * n, 256 * 8);
    if (gx > kNumSMs) gx = kNumSMs;
    if (gx < 1) gx = 1;
    poly_fro_dense_kernel<<<dim3(gx, m), 256, 0, ctx->stream>>>(n, m, a, lam_d, ctx->red_d);
    KLAUNCH_CHECK(ctx);
    reduce_partials2_kernel<<<ceil_div(m, 128), 128, 0, ctx->stream>>>(ctx->red_d, gx, m, fro2_d);
    KLAUNCH_CHECK(ctx);
    return 0;
}

// ------------------------------------------------------------------------------- diagnostics
namespace {
__global__ void count_nonfinite_kernel(int64_t count, const double* __restrict__ p, unsigned long long* out) {
    unsigned long long c = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x)
        if (!isfinite(p[t])) ++c;
    if (c) atomicAdd(out, c);
}
}  // namespace

// FEAST_DEBUG_NAN=1: report blocks that contain non-finite entries (development aid)
int debug_check_finite(feast_ctx* ctx, const void* p, int64_t ndoubles, const char* name) {
    static const bool on = getenv("FEAST_DEBUG_NAN") != nullptr;
    if (!on || !p) return 0;
    unsigned long long* d = nullptr;
    cudaMalloc(&d, sizeof(unsigned long long));
    cudaMemsetAsync(d, 0, sizeof(unsigned long long), ctx->stream);
    count_nonfinite_kernel<<<kNumSMs, 256, 0, ctx->stream>>>(ndoubles, (const double*)p, d);
    unsigned long long h = 0;
    cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (h) fprintf(stderr, "[feast debug] %s: %llu non-finite of %lld doubles\n", name, h, (long long)ndoubles);
    return 0;
}
