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
                    const bool ok = j < m;   // c < 0: padding entry of the device layout (value 0)
                    x0[k] = (ok && c0 >= 0) ? __ldg(X + (int64_t)c0 * ldx + j) : cmake(0, 0);
                    x1[k] = (ok && c1 >= 0) ? __ldg(X + (int64_t)c1 * ldx + j) : cmake(0, 0);
                    x2[k] = (ok && c2 >= 0) ? __ldg(X + (int64_t)c2 * ldx + j) : cmake(0, 0);
                    x3[k] = (ok && c3 >= 0) ? __ldg(X + (int64_t)c3 * ldx + j) : cmake(0, 0);
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
                if (c0 < 0) continue;
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


// ------------------------------------------------------------------ tiled SpMM (the Krylov hot kernel)
// One CTA pass = one tile of consecutive matrix rows (tile plan: reorder.cpp) x one slab of G <= 32 columns:
//   * every row of the input block the tile references -- its own rows, then its halo list -- is brought
//     into shared memory by bulk async copies (cp.async.bulk -> UBLKCP, completion on an mbarrier); the tile's
//     CSR slice (values + 16-bit tile-local column numbers) is staged the same way, once per tile;
//   * the products are then computed ENTIRELY out of shared memory (LDS + DFMA): no long-latency load sits in
//     the inner loop, the memory-level parallelism is the copy engine's, not the register file's;
//   * L2->SM traffic per SpMM = (1 + halo/rows) reads of X; with the tile ordering ~2.1 instead of ~5.5.
// 2 CTAs of 512 threads per SM: one computes while the other waits for its copies.
// Measured alternatives at C2 (n = 1e6, m0 = 64, real values 0.515 ms with this kernel):
//   * (round 2) padding entries predicated off (@P LDS.128) and the row's own x value taken from the diagonal entry instead of
//     a second load -- 18 % fewer shared-memory wavefronts on paper: 0.78 ms instead of 0.63 ms for the complex + dot variant
//     (64 registers instead of 58, the extra compares sit in the dependent chain of the loads); reverted;
//   * 3 or 4 smaller CTAs per SM (TileCfg2 / TileCfg1): 0.510 / 0.505 ms although the halo grows to 1.45 / 1.68;
//   * one 1024-thread CTA with a ring of 2..4 tile buffers, copies issued one to three passes ahead: 0.70 .. 1.09 ms;
//   * the same ring fed by four dedicated producer warps (empty/full mbarriers, no CTA barrier): 0.54 ms with per-row
//     bulk copies, 0.74 ms with 16-byte cp.async;
//   * copies + stores alone (products skipped) take 0.36 ms = 3.35 GB through the SM<->L2 fabric at 9.2 TB/s, i.e. the
//     fabric's ceiling: what is left above it is product time that two CTAs per SM do not hide completely.
// Tile configurations (FEAST_TILE_CFG): CTAs per SM x threads, block rows (tile + halo) resident per CTA, staged
// (padded) nonzeros and max rows per tile.  More, smaller CTAs keep more bulk copies in flight per SM (each CTA
// waits for its whole tile before computing) at the price of a larger halo.  Shared memory: 233472 B per SM,
// 1 KB reserved per CTA.
template <int THREADS, int CTAS, int ROWS, int NNZ, int TMAX> struct TileCfg {
    static constexpr int kThreads = THREADS, kCtas = CTAS, kRowsCap = ROWS, kNnzCap = NNZ, kTileMax = TMAX;
    static constexpr int kBudget = 233472 / CTAS - 1024;
};
typedef TileCfg<512, 2, 192, 768, 160> TileCfg0;   // 96 KB of block rows per CTA
typedef TileCfg<256, 4, 96, 384, 80> TileCfg1;     // 48 KB
typedef TileCfg<384, 3, 128, 512, 112> TileCfg2;   // 64 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // bounded spin: a transaction-count mismatch must fail the launch (trap -> cudaErrorLaunchFailure), not hang the device
    uint32_t done = 0;
    for (uint32_t spins = 0; spins < (1u << 26); ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    asm volatile("trap;");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// four consecutive staged values (16-byte aligned: rows are padded to 8 entries)
__device__ __forceinline__ void load_vals4(const double* p, double (&v)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void load_vals4(const c128* p, c128 (&v)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = p[k];
}

template <typename VT, typename CFG> struct TiledSmem {
    static constexpr __host__ __device__ size_t xs_bytes(int G) { return (size_t)CFG::kRowsCap * G * sizeof(c128); }
    static constexpr size_t vs_bytes = (size_t)CFG::kNnzCap * sizeof(VT);
    static constexpr size_t ls_bytes = (size_t)CFG::kNnzCap * sizeof(uint16_t);
    static constexpr size_t rs_bytes = (size_t)(CFG::kTileMax + 4) * sizeof(int);
    static constexpr __host__ __device__ size_t total(int G) { return xs_bytes(G) + vs_bytes + ls_bytes + rs_bytes; }
};

// G lanes own one row of the slab at a time (one accumulator per lane; up to 8 products in flight).
// m <= 2*G: the launch covers one or two slabs.
// EPI (fused epilogues of the multigrid cycle, amg.cu): 0 Y = S X ; 1 Y = C - S X (residual) ; 2 Y = X + w dinv (C - S X)
// (damped-Jacobi post-smoothing step; Y must not alias X: other tiles read X rows as halo).  With DOT the fused column
// reduction is <X, S X> for EPI 0 and <C, Y> otherwise (the <r, z> of the preconditioned COCG recurrence).
struct SpmmEpi { int mode; const c128* C; int ldc; const c128* dinv; double omega; };

template <typename VT, int G, bool DOT, typename CFG, int EPI>
__global__ void __launch_bounds__(CFG::kThreads, CFG::kCtas)
spmm_tiled_kernel(int m, int ntiles, const int* __restrict__ t_ptr, const int* __restrict__ t_hptr,
                  const int* __restrict__ t_hidx, const int* __restrict__ rowptr, const uint16_t* __restrict__ lcol,
                  const VT* __restrict__ val, const c128* __restrict__ X, int ldx, c128* __restrict__ Y, int ldy,
                  double* __restrict__ partials, int dbg, SpmmEpi ep) {
    // dbg (FEAST_SPMM_DEBUG, timing experiments only): 1 = skip the products, 2 = skip the halo copies, 4 = skip the stores
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    c128* xs = (c128*)smem_raw;                                                 // [kRowsCap][SW]
    VT* vs = (VT*)(smem_raw + TiledSmem<VT, CFG>::xs_bytes(G));                      // [kNnzCap]
    uint16_t* ls = (uint16_t*)((unsigned char*)vs + TiledSmem<VT, CFG>::vs_bytes);   // [kNnzCap]
    int* rs = (int*)((unsigned char*)ls + TiledSmem<VT, CFG>::ls_bytes);             // [kTileMax + 1]
    constexpr int UPW = 32 / G;   // rows per warp
    constexpr int kTiledThreads = CFG::kThreads;
    constexpr int NW = kTiledThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane % G, sub = lane / G;

    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t phase = 0;
    c128 dacc[2];
    dacc[0] = cmake(0.0, 0.0);
    dacc[1] = cmake(0.0, 0.0);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int r0 = __ldg(t_ptr + tile), rows = __ldg(t_ptr + tile + 1) - r0;
        const int hp = __ldg(t_hptr + tile);
        const int nref = (dbg & 2) ? rows : rows + __ldg(t_hptr + tile + 1) - hp;   // own rows + halo rows
        const int ea = __ldg(rowptr + r0), eb = __ldg(rowptr + r0 + rows);   // multiples of 8 (padded rows)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int j0 = s * G;
            if (j0 >= m) break;
            const int SW = (m - j0) < G ? (m - j0) : G;
            const uint32_t rowbytes = (uint32_t)SW * (uint32_t)sizeof(c128);
            if (tid == 0) {
                uint32_t bytes = (uint32_t)nref * rowbytes;
                if (s == 0) bytes += (uint32_t)(eb - ea) * (uint32_t)(sizeof(VT) + sizeof(uint16_t));
                mbar_expect_tx(&mbar, bytes);
            }
            __syncthreads();   // the previous pass has finished reading shared memory; the barrier is armed
            // copy t is issued by lane t / NW of warp t % NW: a warp issues its copies one lane at a time
            for (int t = lane * NW + warp; t < nref; t += kTiledThreads) {
                const int srow = t < rows ? r0 + t : __ldg(t_hidx + hp + (t - rows));
                bulk_g2s(xs + (size_t)t * SW, X + (int64_t)srow * ldx + j0, rowbytes, &mbar);
            }
            if (s == 0) {
                if (eb > ea) {
                    if (tid == 32) bulk_g2s(vs, val + ea, (uint32_t)(eb - ea) * (uint32_t)sizeof(VT), &mbar);
                    if (tid == 64) bulk_g2s(ls, lcol + ea, (uint32_t)(eb - ea) * (uint32_t)sizeof(uint16_t), &mbar);
                }
                for (int t = tid; t <= rows; t += kTiledThreads) rs[t] = __ldg(rowptr + r0 + t) - ea;
                __syncthreads();   // row pointers visible
            }
            mbar_wait(&mbar, phase);
            phase ^= 1u;

            const int gg = g < SW ? g : 0;
            for (int lr = warp * UPW + sub; lr < rows; lr += NW * UPW) {
                const int e0 = rs[lr], e1 = rs[lr + 1];
                c128 acc = cmake(0.0, 0.0);
                for (int e = (dbg & 1) ? e1 : e0; e < e1; e += 8) {
                    // 8 tile-local columns in one 16-byte broadcast load; 0xFFFF = padding (only at the end of a row)
                    const uint4 iv = *reinterpret_cast<const uint4*>(ls + e);
                    const unsigned w[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        unsigned lc[4];
                        lc[0] = w[2 * h] & 0xFFFFu; lc[1] = w[2 * h] >> 16; lc[2] = w[2 * h + 1] & 0xFFFFu; lc[3] = w[2 * h + 1] >> 16;
                        if (lc[0] == 0xFFFFu) break;
                        c128 xv[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) xv[k] = xs[(size_t)(lc[k] == 0xFFFFu ? (unsigned)lr : lc[k]) * SW + gg];
                        VT vv[4];
                        load_vals4(vs + e + 4 * h, vv);   // padding values are 0
#pragma unroll
                        for (int k = 0; k < 4; ++k) ValOps<VT>::fma(acc, vv[k], xv[k]);
                    }
                }
                if (g < SW && !(dbg & 4)) {
                    if (EPI == 0) {
                        Y[(int64_t)(r0 + lr) * ldy + j0 + g] = acc;
                        if (DOT) cfma(dacc[s], xs[(size_t)lr * SW + g], acc);
                    } else {
                        const c128 cv = __ldg(ep.C + (int64_t)(r0 + lr) * ep.ldc + j0 + g);
                        c128 outv = csub(cv, acc);
                        if (EPI == 2) {
                            const c128 wd = cscale(ep.omega, __ldg(ep.dinv + r0 + lr));
                            const c128 resid = outv;
                            outv = xs[(size_t)lr * SW + g];
                            cfma(outv, wd, resid);
                        }
                        Y[(int64_t)(r0 + lr) * ldy + j0 + g] = outv;
                        if (DOT) cfma(dacc[s], cv, outv);
                    }
                }
            }
        }
    }
    if (DOT) {
        __syncthreads();
        // the tile buffer is free now: reuse it for the cross-warp reduction.  slot = (warp, sub), 2 slabs each
        double* sred = (double*)smem_raw;   // [NW * UPW][2 slabs][2 * G]
        const int slot = warp * UPW + sub;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            sred[((size_t)slot * 2 + s) * 2 * G + 2 * g] = dacc[s].x;
            sred[((size_t)slot * 2 + s) * 2 * G + 2 * g + 1] = dacc[s].y;
        }
        __syncthreads();
        for (int t = tid; t < 2 * m; t += kTiledThreads) {
            const int j = t >> 1, s = j / G, gq = j - s * G;
            double acc = 0.0;
            for (int sl = 0; sl < NW * UPW; ++sl) acc += sred[((size_t)sl * 2 + s) * 2 * G + 2 * gq + (t & 1)];
            partials[(int64_t)blockIdx.x * 2 * m + t] = acc;
        }
    }
}

// ------------------------------------------------------------------ mixed-precision variant
// Same tiled kernel for blocks stored in complex64 (the reference's `mixed_prec=true`: ComplexF32 solves inside the
// double-precision RII loop, src/feast.jl:19-25).  A 16-byte unit of a row holds TWO complex64 columns, so the copy
// geometry is the one of the complex128 kernel with half as many units per row (m0 = 64 -> one slab of 32 units, one
// pass per tile); products and the fused <p, Zp> accumulate in double, the operator values stay complex128.
// First run on a B200 in round 2 (profiles/r2_round2_validate.log); reachable through feast_set_mixed_precision.
template <bool DOT, typename CFG>
__global__ void __launch_bounds__(CFG::kThreads, CFG::kCtas)
spmm_tiled_f32_kernel(int mu, int ntiles, const int* __restrict__ t_ptr, const int* __restrict__ t_hptr,
                      const int* __restrict__ t_hidx, const int* __restrict__ rowptr, const uint16_t* __restrict__ lcol,
                      const c128* __restrict__ val, const float4* __restrict__ X, int ldx, float4* __restrict__ Y, int ldy,
                      double* __restrict__ partials) {
    constexpr int G = 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    float4* xs = (float4*)smem_raw;                                                    // [kRowsCap][SW] 16-byte units
    c128* vs = (c128*)(smem_raw + TiledSmem<c128, CFG>::xs_bytes(G));                  // [kNnzCap]
    uint16_t* ls = (uint16_t*)((unsigned char*)vs + TiledSmem<c128, CFG>::vs_bytes);   // [kNnzCap]
    int* rs = (int*)((unsigned char*)ls + TiledSmem<c128, CFG>::ls_bytes);             // [kTileMax + 1]
    constexpr int kTiledThreads = CFG::kThreads;
    constexpr int NW = kTiledThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane;

    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t phase = 0;
    c128 dacc[2][2];
#pragma unroll
    for (int s = 0; s < 2; ++s) { dacc[s][0] = cmake(0.0, 0.0); dacc[s][1] = cmake(0.0, 0.0); }

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int r0 = __ldg(t_ptr + tile), rows = __ldg(t_ptr + tile + 1) - r0;
        const int hp = __ldg(t_hptr + tile);
        const int nref = rows + __ldg(t_hptr + tile + 1) - hp;
        const int ea = __ldg(rowptr + r0), eb = __ldg(rowptr + r0 + rows);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int j0 = s * G;
            if (j0 >= mu) break;
            const int SW = (mu - j0) < G ? (mu - j0) : G;
            const uint32_t rowbytes = (uint32_t)SW * 16u;
            if (tid == 0) {
                uint32_t bytes = (uint32_t)nref * rowbytes;
                if (s == 0) bytes += (uint32_t)(eb - ea) * (uint32_t)(sizeof(c128) + sizeof(uint16_t));
                mbar_expect_tx(&mbar, bytes);
            }
            __syncthreads();
            for (int t = lane * NW + warp; t < nref; t += kTiledThreads) {
                const int srow = t < rows ? r0 + t : __ldg(t_hidx + hp + (t - rows));
                bulk_g2s(xs + (size_t)t * SW, X + (int64_t)srow * ldx + j0, rowbytes, &mbar);
            }
            if (s == 0) {
                if (eb > ea) {
                    if (tid == 32) bulk_g2s(vs, val + ea, (uint32_t)(eb - ea) * (uint32_t)sizeof(c128), &mbar);
                    if (tid == 64) bulk_g2s(ls, lcol + ea, (uint32_t)(eb - ea) * (uint32_t)sizeof(uint16_t), &mbar);
                }
                for (int t = tid; t <= rows; t += kTiledThreads) rs[t] = __ldg(rowptr + r0 + t) - ea;
                __syncthreads();
            }
            mbar_wait(&mbar, phase);
            phase ^= 1u;

            const int gg = g < SW ? g : 0;
            for (int lr = warp; lr < rows; lr += NW) {
                const int e0 = rs[lr], e1 = rs[lr + 1];
                c128 acc0 = cmake(0.0, 0.0), acc1 = cmake(0.0, 0.0);
                for (int e = e0; e < e1; e += 8) {
                    const uint4 iv = *reinterpret_cast<const uint4*>(ls + e);
                    const unsigned w[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        unsigned lc[4];
                        lc[0] = w[2 * h] & 0xFFFFu; lc[1] = w[2 * h] >> 16; lc[2] = w[2 * h + 1] & 0xFFFFu; lc[3] = w[2 * h + 1] >> 16;
                        if (lc[0] == 0xFFFFu) break;
                        float4 xv[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) xv[q] = xs[(size_t)(lc[q] == 0xFFFFu ? (unsigned)lr : lc[q]) * SW + gg];
                        c128 vv[4];
                        load_vals4(vs + e + 4 * h, vv);   // padding values are 0
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            cfma(acc0, vv[q], cmake((double)xv[q].x, (double)xv[q].y));
                            cfma(acc1, vv[q], cmake((double)xv[q].z, (double)xv[q].w));
                        }
                    }
                }
                if (g < SW) {
                    const float4 qv = make_float4((float)acc0.x, (float)acc0.y, (float)acc1.x, (float)acc1.y);
                    Y[(int64_t)(r0 + lr) * ldy + j0 + g] = qv;
                    if (DOT) {   // <p, q> of the q that is STORED (the recurrence continues from the rounded block)
                        const float4 own = xs[(size_t)lr * SW + g];
                        cfma(dacc[s][0], cmake((double)own.x, (double)own.y), cmake((double)qv.x, (double)qv.y));
                        cfma(dacc[s][1], cmake((double)own.z, (double)own.w), cmake((double)qv.z, (double)qv.w));
                    }
                }
            }
        }
    }
    if (DOT) {
        __syncthreads();
        double* sred = (double*)smem_raw;   // [NW][2 slabs][G][4]
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            double* o = sred + (((size_t)warp * 2 + s) * G + g) * 4;
            o[0] = dacc[s][0].x; o[1] = dacc[s][0].y; o[2] = dacc[s][1].x; o[3] = dacc[s][1].y;
        }
        __syncthreads();
        // output t = 2 * column + (0: re, 1: im); column c lives in unit c / 2 = s * G + gq, half c & 1
        for (int t = tid; t < 4 * mu; t += kTiledThreads) {
            const int c = t >> 1, unit = c >> 1, s = unit / G, gq = unit - s * G;
            const int comp = 2 * (c & 1) + (t & 1);
            double acc = 0.0;
            for (int wq = 0; wq < NW; ++wq) acc += sred[(((size_t)wq * 2 + s) * G + gq) * 4 + comp];
            partials[(int64_t)blockIdx.x * 4 * mu + t] = acc;
        }
    }
}

template <typename VT, int G, typename CFG, int EPI>
int spmm_tiled_launch(feast_ctx* ctx, int m, const VT* val, const c128* X, int ldx, c128* Y, int ldy, c128* dot_out, SpmmEpi ep) {
    const size_t smem = TiledSmem<VT, CFG>::total(G) < 16384 ? 16384 : TiledSmem<VT, CFG>::total(G);
    static_assert(TiledSmem<VT, CFG>::total(32) <= (size_t)CFG::kBudget, "tile does not fit the CTAs-per-SM target");
    const int ntiles = ctx->ntiles;
    const int grid = ntiles < CFG::kCtas * kNumSMs ? ntiles : CFG::kCtas * kNumSMs;
    static bool attr_done_dot[64] = {}, attr_done[64] = {};   // per device and instantiation (function attributes are per context)
    const int dev = ctx->device & 63;
    static const int dbg = getenv("FEAST_SPMM_DEBUG") ? atoi(getenv("FEAST_SPMM_DEBUG")) : 0;
    if (dot_out) {
        auto kern = spmm_tiled_kernel<VT, G, true, CFG, EPI>;
        if (!attr_done_dot[dev]) {
            CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::kBudget));
            attr_done_dot[dev] = true;
        }
        kern<<<grid, CFG::kThreads, smem, ctx->stream>>>(m, ntiles, ctx->t_ptr, ctx->t_hptr, ctx->t_hidx, ctx->u_rowptr, ctx->u_lcol,
                                                          val, X, ldx, Y, ldy, ctx->red_d, dbg, ep);
        KLAUNCH_CHECK(ctx);
        reduce_partials_kernel<<<ceil_div(2 * m * 32, 128), 128, 0, ctx->stream>>>(ctx->red_d, grid, 2 * m, (double*)dot_out);
        KLAUNCH_CHECK(ctx);
    } else {
        auto kern = spmm_tiled_kernel<VT, G, false, CFG, EPI>;
        if (!attr_done[dev]) {
            CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::kBudget));
            attr_done[dev] = true;
        }
        kern<<<grid, CFG::kThreads, smem, ctx->stream>>>(m, ntiles, ctx->t_ptr, ctx->t_hptr, ctx->t_hidx, ctx->u_rowptr, ctx->u_lcol,
                                                          val, X, ldx, Y, ldy, nullptr, dbg, ep);
        KLAUNCH_CHECK(ctx);
    }
    return 0;
}

template <typename VT, typename CFG, int EPI>
int spmm_tiled_dispatch_e(feast_ctx* ctx, int m, const VT* val, const c128* X, int ldx, c128* Y, int ldy, c128* dot_out, SpmmEpi ep) {
    if (m <= 4) return spmm_tiled_launch<VT, 4, CFG, EPI>(ctx, m, val, X, ldx, Y, ldy, dot_out, ep);
    if (m <= 8) return spmm_tiled_launch<VT, 8, CFG, EPI>(ctx, m, val, X, ldx, Y, ldy, dot_out, ep);
    if (m <= 16) return spmm_tiled_launch<VT, 16, CFG, EPI>(ctx, m, val, X, ldx, Y, ldy, dot_out, ep);
    return spmm_tiled_launch<VT, 32, CFG, EPI>(ctx, m, val, X, ldx, Y, ldy, dot_out, ep);   // m <= 64: one or two slabs of 32 columns
}
template <typename VT, typename CFG>
int spmm_tiled_dispatch_g(feast_ctx* ctx, int m, const VT* val, const c128* X, int ldx, c128* Y, int ldy, c128* dot_out) {
    return spmm_tiled_dispatch_e<VT, CFG, 0>(ctx, m, val, X, ldx, Y, ldy, dot_out, SpmmEpi{0, nullptr, 0, nullptr, 0.0});
}

int tile_cfg_setting() {
    static const int cfg = [] {
        const char* e = getenv("FEAST_TILE_CFG");
        const int v = e ? atoi(e) : 0;
        return (v < 0 || v > 2) ? 0 : v;
    }();
    return cfg;
}

template <typename VT>
int spmm_tiled_dispatch(feast_ctx* ctx, int m, const VT* val, const c128* X, int ldx, c128* Y, int ldy, c128* dot_out) {
    switch (ctx->tile_cfg) {
        case 1: return spmm_tiled_dispatch_g<VT, TileCfg1>(ctx, m, val, X, ldx, Y, ldy, dot_out);
        case 2: return spmm_tiled_dispatch_g<VT, TileCfg2>(ctx, m, val, X, ldx, Y, ldy, dot_out);
        default: return spmm_tiled_dispatch_g<VT, TileCfg0>(ctx, m, val, X, ldx, Y, ldy, dot_out);
    }
}

}  // namespace

size_t spmm_partials_bytes(int m) { return (size_t)kNumSMs * 8 * 2 * (size_t)(m < 128 ? 128 : m) * sizeof(double); }

int spmm_tile_cfg() { return tile_cfg_setting(); }

TileCaps spmm_tile_caps() {
    static const int domain = getenv("FEAST_TILE_DOMAIN") ? atoi(getenv("FEAST_TILE_DOMAIN")) : 65536;
    switch (tile_cfg_setting()) {
        case 1: return TileCaps{TileCfg1::kRowsCap, TileCfg1::kNnzCap, TileCfg1::kTileMax, domain};
        case 2: return TileCaps{TileCfg2::kRowsCap, TileCfg2::kNnzCap, TileCfg2::kTileMax, domain};
        default: return TileCaps{TileCfg0::kRowsCap, TileCfg0::kNnzCap, TileCfg0::kTileMax, domain};
    }
}

int launch_spmm(feast_ctx* ctx, int64_t n, int m, const int* rowptr, const int* col, const double* rvals,
                const c128* cvals, const c128* X, int ldx, c128* Y, int ldy, c128* dot_out) {
    // tiled kernel on the union pattern (column chunks of 64 = two slabs of 32);
    // FEAST_SPMM_TILED=0 selects the row-per-warp kernel (kept for A/B measurements)
    static const bool tiled_off = getenv("FEAST_SPMM_TILED") && atoi(getenv("FEAST_SPMM_TILED")) == 0;
    const bool tiled = !tiled_off && ctx->tiles_ok && rowptr == ctx->u_rowptr && col == ctx->u_col;
    const int chunk = tiled ? 64 : 128;
    for (int j0 = 0; j0 < m; j0 += chunk) {
        const int mc = (m - j0) < chunk ? (m - j0) : chunk;
        c128* dchunk = dot_out ? dot_out + j0 : nullptr;
        int rc;
        if (tiled)
            rc = rvals ? spmm_tiled_dispatch<double>(ctx, mc, rvals, X + j0, ldx, Y + j0, ldy, dchunk)
                       : spmm_tiled_dispatch<c128>(ctx, mc, cvals, X + j0, ldx, Y + j0, ldy, dchunk);
        else
            rc = rvals ? spmm_dispatch<double>(ctx, (int)n, mc, rowptr, col, rvals, X + j0, ldx, Y + j0, ldy, dchunk)
                       : spmm_dispatch<c128>(ctx, (int)n, mc, rowptr, col, cvals, X + j0, ldx, Y + j0, ldy, dchunk);
        if (rc) return rc;
    }
    return 0;
}

// Tiled SpMM with a fused multigrid epilogue on the union pattern (complex values, default tile configuration):
//   mode 1: Y = C - Z X ; mode 2: Y = X + omega dinv (C - Z X).  dot_out (optional): <C, Y> per column (unconjugated).
// Returns 1 when the fused kernel is not applicable (no tile plan / other tile configuration): the caller falls back.
int launch_spmm_epi(feast_ctx* ctx, int m, const c128* zvals, const c128* X, c128* Y, int mode, const c128* C, const c128* dinv,
                    double omega, c128* dot_out) {
    static const bool off = getenv("FEAST_SPMM_EPI") && atoi(getenv("FEAST_SPMM_EPI")) == 0;
    if (off || !ctx->tiles_ok || ctx->tile_cfg != 0) return 1;
    for (int j0 = 0; j0 < m; j0 += 64) {
        const int mc = (m - j0) < 64 ? (m - j0) : 64;
        SpmmEpi ep{mode, C + j0, m, dinv, omega};
        c128* dchunk = dot_out ? dot_out + j0 : nullptr;
        int rc = mode == 1 ? spmm_tiled_dispatch_e<c128, TileCfg0, 1>(ctx, mc, zvals, X + j0, m, Y + j0, m, dchunk, ep)
                           : spmm_tiled_dispatch_e<c128, TileCfg0, 2>(ctx, mc, zvals, X + j0, m, Y + j0, m, dchunk, ep);
        if (rc) return rc;
    }
    return 0;
}

// q (n x m0 complex64) = Z p with the tiled kernel; dot_out (m0 complex128, optional) = sum_i p_ij q_ij.
// Requires the default tile configuration, an even m0 <= 128 and a usable tile plan; returns FEAST_ERR_STATE otherwise.
int launch_spmm_f32(feast_ctx* ctx, int m0, const c128* zvals, const void* X32, void* Y32, c128* dot_out) {
    if (!ctx->tiles_ok || ctx->tile_cfg != 0 || (m0 & 1) || m0 > 128)
        return feast_fail(ctx, FEAST_ERR_STATE, "mixed-precision SpMM needs the default tile plan and an even m0 <= 128");
    typedef TileCfg0 CFG;
    const int mu = m0 / 2;
    const size_t smem = TiledSmem<c128, CFG>::total(32) < 32768 ? 32768 : TiledSmem<c128, CFG>::total(32);
    const int ntiles = ctx->ntiles;
    const int grid = ntiles < CFG::kCtas * kNumSMs ? ntiles : CFG::kCtas * kNumSMs;
    static bool attr_done_dot[64] = {}, attr_done[64] = {};
    const int dev = ctx->device & 63;
    if (dot_out) {
        auto kern = spmm_tiled_f32_kernel<true, CFG>;
        if (!attr_done_dot[dev]) {
            CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::kBudget));
            attr_done_dot[dev] = true;
        }
        kern<<<grid, CFG::kThreads, smem, ctx->stream>>>(mu, ntiles, ctx->t_ptr, ctx->t_hptr, ctx->t_hidx, ctx->u_rowptr, ctx->u_lcol,
                                                          zvals, (const float4*)X32, mu, (float4*)Y32, mu, ctx->red_d);
        KLAUNCH_CHECK(ctx);
        reduce_partials_kernel<<<ceil_div(2 * m0 * 32, 128), 128, 0, ctx->stream>>>(ctx->red_d, grid, 2 * m0, (double*)dot_out);
        KLAUNCH_CHECK(ctx);
    } else {
        auto kern = spmm_tiled_f32_kernel<false, CFG>;
        if (!attr_done[dev]) {
            CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::kBudget));
            attr_done[dev] = true;
        }
        kern<<<grid, CFG::kThreads, smem, ctx->stream>>>(mu, ntiles, ctx->t_ptr, ctx->t_hptr, ctx->t_hidx, ctx->u_rowptr, ctx->u_lcol,
                                                          zvals, (const float4*)X32, mu, (float4*)Y32, mu, nullptr);
        KLAUNCH_CHECK(ctx);
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
    for (int e = rowptr[warp] + lane; e < rowptr[warp + 1]; e += 32)
        if (col[e] >= 0) Z[(int64_t)warp * n + col[e]] = zvals[e];  // row-major; col < 0: padding entry
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
            if (c0 < 0) continue;   // padding entry
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
