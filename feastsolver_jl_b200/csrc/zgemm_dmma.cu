// Complex128 GEMM on the FP64 tensor pipe (DMMA): C = alpha * op(A) * B + beta * C with generic
// element strides.  tcgen05/TMEM has no FP64 kind, so the Blackwell FP64 tensor path is
// mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4); one complex MAC tile = 4 real MMAs on split
// re/im fragments (no 3M trick: it changes rounding).
//
//   CTA tile 64 x 64, K chunk 16, 256 threads = 8 warps (2 x 4), warp tile 32 x 16 = 4 x 2 MMA tiles
//   global -> shared with 16-byte cp.async (zero-fill predicates on the edges), double buffered
//   shared layout [k][m] with row stride 66 elements (== 2 mod 8): the LDS.128 of an MMA fragment
//   (lane -> (k = lane&3, m = lane>>2)) touches every bank exactly once per quarter-warp phase.
// Used by the recursive LU / TRSM (dense.cu), dense A*Q, Q*Xq and the split-K Gram products.
#include "kernels.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, LDS_S = 66, STAGES = 2;

struct DGemmArgs {
    int M, N;
    int64_t K;
    const c128* A; int64_t sAi, sAk; int conjA;
    const c128* B; int64_t sBk, sBj;
    c128* C; int64_t sCi, sCj;
    c128 alpha, beta;
    int64_t kchunk;   // K range per grid.z slice (split-K)
    c128* partial;    // split-K partial output [z][N][M] or nullptr
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int bytes = pred ? 16 : 0;   // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 2) zgemm_dmma_kernel(DGemmArgs g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* sA = reinterpret_cast<c128*>(smem_raw);                 // [STAGES][BK][LDS_S]
    c128* sB = sA + STAGES * BK * LDS_S;                          // [STAGES][BK][LDS_S]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;                      // 2 x 4 warps
    const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * g.kchunk;
    int64_t kend = kbeg + g.kchunk;
    if (kend > g.K) kend = g.K;
    const int nchunks = (int)((kend - kbeg + BK - 1) / BK);

    // loader mapping: 64 x 16 = 1024 elements per tile, 4 per thread; walk the unit-stride dimension
    const bool a_k_contig = (g.sAk == 1);
    const bool b_j_contig = (g.sBj == 1);
    auto load_stage = [&](int stage, int64_t k0) {
        c128* dA = sA + stage * BK * LDS_S;
        c128* dB = sB + stage * BK * LDS_S;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int ii, kk;
            if (a_k_contig) { kk = tid & 15; ii = (tid >> 4) + 16 * r; }
            else            { ii = tid & 63; kk = (tid >> 6) + 4 * r; }
            const int gi = i0 + ii;
            const int64_t gk = k0 + kk;
            const bool ok = (gi < g.M) && (gk < kend);
            cp_async16(dA + kk * LDS_S + ii, ok ? (const void*)(g.A + gi * g.sAi + gk * g.sAk) : (const void*)g.A, ok);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int jj, kk;
            if (b_j_contig) { jj = tid & 63; kk = (tid >> 6) + 4 * r; }
            else            { kk = tid & 15; jj = (tid >> 4) + 16 * r; }
            const int gj = j0 + jj;
            const int64_t gk = k0 + kk;
            const bool ok = (gj < g.N) && (gk < kend);
            cp_async16(dB + kk * LDS_S + jj, ok ? (const void*)(g.B + gk * g.sBk + gj * g.sBj) : (const void*)g.B, ok);
        }
    };

    double cr[4][2][2], ci[4][2][2];   // [m-tile][n-tile][2 columns]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) { cr[a][b][0] = cr[a][b][1] = 0.0; ci[a][b][0] = ci[a][b][1] = 0.0; }

    if (nchunks > 0) { load_stage(0, kbeg); }
    cp_async_commit();
    const int fk = lane & 3, fm = lane >> 2;      // fragment coordinates
    const double asign = g.conjA ? -1.0 : 1.0;
    for (int kc = 0; kc < nchunks; ++kc) {
        if (kc + 1 < nchunks) load_stage((kc + 1) & 1, kbeg + (int64_t)(kc + 1) * BK);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const c128* tA = sA + (kc & 1) * BK * LDS_S + wm * 32 + fm;
        const c128* tB = sB + (kc & 1) * BK * LDS_S + wn * 16 + fm;
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            c128 af[4], bf[2];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = tA[(k4 + fk) * LDS_S + 8 * a];
#pragma unroll
            for (int b = 0; b < 2; ++b) bf[b] = tB[(k4 + fk) * LDS_S + 8 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double ar = af[a].x, ai = asign * af[a].y, nai = -ai;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    dmma(cr[a][b][0], cr[a][b][1], ar, bf[b].x);
                    dmma(cr[a][b][0], cr[a][b][1], nai, bf[b].y);
                    dmma(ci[a][b][0], ci[a][b][1], ar, bf[b].y);
                    dmma(ci[a][b][0], ci[a][b][1], ai, bf[b].x);
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    // epilogue: thread owns C[row = fm][cols 2*fk, 2*fk+1] of every 8x8 tile
    const bool beta_nz = (g.beta.x != 0.0 || g.beta.y != 0.0);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int gi = i0 + wm * 32 + 8 * a + fm;
        if (gi >= g.M) continue;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int gj = j0 + wn * 16 + 8 * b + 2 * fk + q;
                if (gj >= g.N) continue;
                const c128 acc = cmake(cr[a][b][q], ci[a][b][q]);
                if (g.partial) {
                    g.partial[(int64_t)blockIdx.z * g.M * g.N + (int64_t)gj * g.M + gi] = acc;
                } else {
                    c128 r = cmul(g.alpha, acc);
                    c128* cp = g.C + gi * g.sCi + gj * g.sCj;
                    if (beta_nz) r = cadd(r, cmul(g.beta, *cp));
                    *cp = r;
                }
            }
        }
    }
}

__global__ void splitk_reduce_dmma_kernel(const c128* __restrict__ partial, int nz, int M, int N, c128 alpha,
                                          c128* __restrict__ C, int64_t sCi, int64_t sCj) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M * N) return;
    double re = 0.0, im = 0.0;
    for (int z = 0; z < nz; ++z) {
        const c128 v = partial[(int64_t)z * M * N + t];
        re += v.x; im += v.y;
    }
    const int i = t % M, j = t / M;
    C[i * sCi + j * sCj] = cmul(alpha, cmake(re, im));
}

}  // namespace

int launch_zgemm(feast_ctx* ctx, int M, int N, int64_t K, hc128 alpha, const c128* A, int64_t sAi, int64_t sAk,
                 bool conjA, const c128* B, int64_t sBk, int64_t sBj, hc128 beta, c128* C, int64_t sCi, int64_t sCj) {
    if (M <= 0 || N <= 0) return 0;
    DGemmArgs g;
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.sAi = sAi; g.sAk = sAk; g.conjA = conjA ? 1 : 0;
    g.B = B; g.sBk = sBk; g.sBj = sBj;
    g.C = C; g.sCi = sCi; g.sCj = sCj;
    g.alpha = cmake(alpha.real(), alpha.imag());
    g.beta = cmake(beta.real(), beta.imag());
    const int gx = ceil_div(M, BM), gy = ceil_div(N, BN);
    const size_t smem = sizeof(c128) * 2 * STAGES * BK * LDS_S;
    if (!ctx->dmma_attr_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(zgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->dmma_attr_set = true;
    }
    int splitk = 1;
    const bool beta_zero = (beta == hc128(0.0, 0.0));
    if (beta_zero && (int64_t)gx * gy < kNumSMs && K >= 1024) {  // tall-skinny (Gram) / skinny right-hand sides: split K over the SMs
        splitk = (2 * kNumSMs) / (gx * gy);
        int64_t maxsplit = K / 512;
        if (splitk > maxsplit) splitk = (int)maxsplit;
        if (splitk < 1) splitk = 1;
        while (splitk > 1 && (size_t)splitk * M * N * sizeof(c128) > ctx->red_bytes) --splitk;
    }
    if (splitk > 1) {
        int64_t kc = (K + splitk - 1) / splitk;
        kc = ((kc + BK - 1) / BK) * BK;
        splitk = (int)((K + kc - 1) / kc);
        g.kchunk = kc;
        g.partial = (c128*)ctx->red_d;
        zgemm_dmma_kernel<<<dim3(gx, gy, splitk), 256, smem, ctx->stream>>>(g);
        KLAUNCH_CHECK(ctx);
        splitk_reduce_dmma_kernel<<<ceil_div((int64_t)M * N, 256), 256, 0, ctx->stream>>>(g.partial, splitk, M, N, g.alpha,
                                                                                       C, sCi, sCj);
        KLAUNCH_CHECK(ctx);
    } else {
        g.kchunk = K > 0 ? K : 1;
        g.partial = nullptr;
        zgemm_dmma_kernel<<<dim3(gx, gy, 1), 256, smem, ctx->stream>>>(g);
        KLAUNCH_CHECK(ctx);
    }
    return 0;
}
