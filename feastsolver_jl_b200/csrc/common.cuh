// Shared definitions for libfeast_cuda.so (sm_100a only; no other backend).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <complex>

#include "../../include/feast_cuda.h"

typedef double2 c128;                       // interleaved (re, im) == Julia ComplexF64
typedef std::complex<double> hc128;         // host-side twin

// ------------------------------------------------------------------ complex helpers
__host__ __device__ __forceinline__ c128 cmake(double r, double i) { return make_double2(r, i); }
__host__ __device__ __forceinline__ c128 cadd(c128 a, c128 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ c128 csub(c128 a, c128 b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ c128 cmul(c128 a, c128 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ c128 cconj(c128 a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ c128 cscale(double s, c128 a) { return make_double2(s * a.x, s * a.y); }
// acc += a*b  (4 FMAs)
__device__ __forceinline__ void cfma(c128& acc, c128 a, c128 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a)*b
__device__ __forceinline__ void cfma_conj(c128& acc, c128 a, c128 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
// acc += s*b, s real
__device__ __forceinline__ void rfma(c128& acc, double s, c128 b) {
    acc.x = fma(s, b.x, acc.x); acc.y = fma(s, b.y, acc.y);
}
__host__ __device__ __forceinline__ c128 cdiv(c128 a, c128 b) {
    // Smith's algorithm (robust against overflow of |b|^2)
    if (fabs(b.x) >= fabs(b.y)) {
        double r = b.y / b.x, d = b.x + b.y * r;
        return make_double2((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        double r = b.x / b.y, d = b.x * r + b.y;
        return make_double2((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}
__host__ __device__ __forceinline__ double cabs2(c128 a) { return a.x * a.x + a.y * a.y; }
__host__ __device__ __forceinline__ double cabs1(c128 a) { return fabs(a.x) + fabs(a.y); }  // LAPACK izamax measure

// ------------------------------------------------------------------ operator storage
enum { OP_NONE = 0, OP_IDENTITY = 1, OP_DENSE = 2, OP_CSR = 3 };

struct HostCSR {            // host copy kept until the union pattern is built
    int64_t n = 0, nnz = 0;
    std::vector<int64_t> rowptr;
    std::vector<int> col;
    std::vector<hc128> val;
    bool is_complex = false;
    bool symmetric = false; // S == S^T (values, not conjugated)
};

struct Operator {
    int kind = OP_NONE;
    int64_t n = 0;
    bool is_complex = false;
    bool symmetric = false;
    c128* dense = nullptr;     // OP_DENSE: column-major n x n (device)
    HostCSR host;              // OP_CSR before feast_set_problem
    double* uvals_r = nullptr; // OP_CSR/IDENTITY after set_problem: values on the union pattern (real ...)
    c128*   uvals_c = nullptr; // ... or complex
};

struct BlockVec {            // n x m0 complex block, ROW-MAJOR (m0 contiguous): the
    c128* p = nullptr;       // layout the CSR SpMM gathers coalesced 16*m0-byte rows from
};

struct DenseLU {             // one stored factorisation (column-major, LAPACK getrf layout)
    c128* lu = nullptr;
    int*  ipiv = nullptr;    // device, 0-based absolute row indices (LAPACK interchange sequence)
    int*  perm = nullptr;    // device, perm[i] = source row of permuted row i
    c128* dinv = nullptr;    // explicit inverses of the diagonal blocks of L then U (2 x n x kDiagNB, row-major blocks)
    int64_t n = 0;
};

struct BandFactor {          // band LU with partial pivoting across adjacent block rows (band.cu)
    int b = 0, nbk = 0;      // block size (>= half bandwidth) and number of block rows
    c128* lu = nullptr;      // [nbk][b*b] row-major L11\U11 of each 2b x b panel (last block: LU of the final Schur complement)
    int* piv = nullptr;      // [nbk][2*b] window-relative interchanges (last block: ipiv then perm)
    c128* dinv = nullptr;    // [2*b*kDiagNB] diagonal-block inverses of the last block
    c128* l21 = nullptr;     // [nbk][b*b] multipliers of the lower half of each panel
    c128* u12 = nullptr;     // [nbk][b*2b] row-major (ld 2b): the U rows of block columns I+1, I+2
};

struct feast_factor {        // fine-grained plugin handle
    int kind = 0;            // FEAST_SOLVER_DENSE_LU / FEAST_SOLVER_KRYLOV
    DenseLU lu;
    c128* zvals = nullptr;   // Krylov / banded: assembled union-pattern values
    BandFactor band;
    bool symmetric = false;
    hc128 coef[FEAST_MAX_SLOTS];   // the shift coefficients (the multigrid levels are assembled from them at solve time)
};

struct NcclApi;              // nccl_dl.cpp
struct AmgDev;               // amg.cu

struct feast_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;

    // problem
    int problem = FEAST_PROBLEM_STANDARD;
    int nslots = 0;
    bool problem_ready = false;
    bool storage_dense = false;
    int64_t n = 0;
    Operator ops[FEAST_MAX_SLOTS];
    // union sparsity pattern of all sparse slots (CSR, 32-bit columns)
    int64_t unnz = 0;
    int* u_rowptr = nullptr;
    int* u_col = nullptr;
    bool all_symmetric = false;
    // internal layout of the sparse path (reorder.cpp / spmm.cu): rows cut into tiles whose referenced block rows
    // fit in shared memory; with Krylov inner solves the rows are renumbered so that the tiles are compact
    int* perm_d = nullptr;        // new -> old row map (nullptr: natural order); applied on upload / download
    bool reordered = false;
    bool tiles_ok = false;        // tile plan usable by the tiled SpMM
    int ntiles = 0;
    int tile_cfg = 0;             // tile configuration of the plan (spmm.cu TileCfg*)
    int* t_ptr = nullptr;         // [ntiles + 1] first row of each tile
    int* t_hptr = nullptr;        // [ntiles + 1] halo list offsets
    int* t_hidx = nullptr;        // halo rows per tile
    uint16_t* u_lcol = nullptr;   // [unnz] tile-local column numbers (own rows, then the halo list)
    double halo_ratio = 0.0;      // halo rows per row (diagnostic)
    c128* zvals = nullptr;        // assembled shifted operator on the union pattern
    c128* zvals_pc = nullptr;     // ... at the preconditioner's shift (complex-shifted multigrid, precond_shift != 0)
    double precond_shift = 0.0;   // beta: the multigrid hierarchy is assembled at z + i beta |z| sign(Im z)
    c128* zdense = nullptr;       // assembled dense shifted operator (n x n col-major)
    int*  zpiv = nullptr;
    c128* zdinv = nullptr;        // diagonal-block inverses of the scratch factorisation

    // contour
    std::vector<hc128> znodes, zweights;
    std::vector<int> owner;       // node -> rank
    std::vector<double> node_cost; // measured device ms of the last solve of each node (all ranks, after the all-reduce)
    bool have_costs = false;
    std::vector<double> cost_local;   // this rank's measurements of the running pass
    feast_stats pass_stats;       // statistics of the running pass (node-by-node entries)
    int pass_rc = 0;
    int auto_balance = 1;         // re-shard nodes by measured cost before every contour pass (Krylov / store=0 only)
    // solver
    int solver = FEAST_SOLVER_AUTO, krylov = FEAST_KRYLOV_AUTO;
    double inner_tol = 1e-10;
    int max_inner = 5000;
    int store = 0;
    int precond = FEAST_PRECOND_AUTO;   // Krylov preconditioner request (feast_set_preconditioner)
    AmgDev* amg = nullptr;        // smoothed-aggregation hierarchy (built with the union pattern when applicable)
    std::string amg_why;          // why no hierarchy was built (diagnostic)
    int mixed_prec = 0;           // mixed_prec: complex64 storage of the COCG blocks (feast_set_mixed_precision)
    std::vector<DenseLU> stored;  // per node (only local nodes populated)
    std::vector<BandFactor> bstored; // per node, banded solver
    BandFactor bscratch;          // store=0
    c128* band_tmp = nullptr;     // 5 x b*b + 4 x b*max(b,m0) work blocks of the banded solver
    size_t band_tmp_elems = 0;
    int bandwidth = 0;            // max |i - j| of the union pattern
    bool panel_attr_set = false;
    bool dmma_attr_set = false;
    int dense_threshold = 6000;   // sparse problems up to this n are solved by dense LU

    // subspace blocks
    int m0 = 0;
    BlockVec Q, X, R, Q1, W1, W2;       // Q1: second moment accumulator (polynomial)
    std::vector<BlockVec> mom;          // moment accumulators S_2, S_3, ... (S_0 = Q, S_1 = Q1)
    int nmom = 0;                       // requested number of moments (0: default of the problem kind)
    std::vector<double> last_fro;       // ||T(l_j)||_F of the last polynomial residual
    int shard_mode = FEAST_SHARD_AUTO;  // multi-GPU sharding axis of the contour loop
    bool col_shard = false;             // the running pass shards right-hand-side columns instead of nodes
    int ngroups = 1;                    // column mode: rank groups (a node belongs to one group, see choose_groups in api.cu)
    std::vector<int> gowner;            // column mode: node -> group
    BlockVec Ql, Xl, Rl;                // left subspace of the two-sided driver (dual_gen_feast!)
    BlockVec kx, kr, kp, kq, ks, kt, kv, krh; // Krylov work
    c128* gm_V = nullptr;         // GMRES basis: (gm_restart + 1) blocks
    void* gm_small = nullptr;     // GMRES per-column Hessenberg / rotations
    int gm_restart = 0;
    c128* stage = nullptr;        // n x m0 column-major staging (uploads / downloads)
    // device scratch: [4 m0^2: m0 x m0 matrices][16 m0 + 64: Krylov per-column scalars (krylov.cu carve_scalars)]
    // [6 m0: lambda, conj(lambda), d, dl, 2 m0 doubles of norms][kMaxNodes doubles: node costs].  Every user has its own
    // region (round 1 carved lambda / d / the cost buffer out of the matrix area, which overlapped for m0 == 1).
    c128* small_d = nullptr;
    static constexpr int kMaxNodes = 4096;
    static size_t small_elems(int m) { return (size_t)4 * m * m + 16 * (size_t)m + 64 + 6 * (size_t)m + kMaxNodes / 2 + 16; }
    c128* vec_base() const { return small_d + (size_t)4 * m0 * m0 + 16 * (size_t)m0 + 64; }
    c128* vec_lam() const { return vec_base(); }
    c128* vec_lamc() const { return vec_base() + m0; }
    c128* vec_d() const { return vec_base() + 2 * (size_t)m0; }
    c128* vec_dl() const { return vec_base() + 3 * (size_t)m0; }
    double* vec_nrm() const { return (double*)(vec_base() + 4 * (size_t)m0); }   // 4 m0 doubles
    double* vec_cost() const { return (double*)(vec_base() + 6 * (size_t)m0 + 8); }
    double* red_d = nullptr;      // reduction partials
    size_t red_bytes = 0;
    void* pinned = nullptr;       // pinned host scratch
    size_t pinned_bytes = 0;
    std::vector<hc128> orthR;     // accumulated R factor of the last orthonormalisation (m0 x m0, col-major)

    // multi-GPU
    int nranks = 1, rank = 0;
    void* nccl_comm = nullptr;

    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t sw0 = nullptr, sw1 = nullptr;   // user stopwatch (feast_timer_*)
    cudaEvent_t evn[4] = {nullptr, nullptr, nullptr, nullptr};   // per-node factor / solve brackets of the contour loop
    cudaEvent_t evk[32] = {};                   // SpMM brackets inside the Krylov solves (read back at the convergence checks)
    double phase_ms[3] = {0, 0, 0};
};

// ------------------------------------------------------------------ error plumbing
extern std::string g_last_error;
int feast_fail(feast_ctx* ctx, int code, const char* fmt, ...);

#define CUDA_TRY(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            int code__ = (e__ == cudaErrorMemoryAllocation) ? FEAST_ERR_OOM : FEAST_ERR_CUDA; \
            return feast_fail(ctx, code__, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, \
                              cudaGetErrorString(e__));                                       \
        }                                                                                     \
    } while (0)

#define FEAST_TRY(call)            \
    do {                           \
        int rc__ = (call);         \
        if (rc__ != 0) return rc__; \
    } while (0)

#define KLAUNCH_CHECK(ctx)                                                                     \
    do {                                                                                       \
        (ctx)->launches++;                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess)                                                                \
            return feast_fail(ctx, FEAST_ERR_CUDA, "kernel launch failed at %s:%d: %s", __FILE__, \
                              __LINE__, cudaGetErrorString(e__));                              \
    } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
constexpr int kDiagNB = 256;   // diagonal-block size of the blocked triangular solves (dense.cu)
constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this
