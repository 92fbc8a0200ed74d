// C ABI of libfeast_cuda.so: context, operator upload, and the three per-iteration phases
// (project / recover+residual / contour apply) of the FEAST outer loop, plus the
// fine-grained factorizer / left_divider plugin pair.  See include/feast_cuda.h for the
// reference statements each entry replaces.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <map>

#include <thread>

#include "amg.h"
#include "host_small.h"
#include "kernels.cuh"
#include "nccl_dl.h"

std::string g_last_error;

int feast_fail(feast_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}

#define ARG_CHECK(ctx, cond, k, msg) \
    do { if (!(cond)) return feast_fail(ctx, -(k), "argument %d invalid: %s", (k), msg); } while (0)

// ------------------------------------------------------------------------------- helpers
namespace {

template <typename T>
int dev_alloc(feast_ctx* ctx, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, sizeof(T) * (count ? count : 1));
    if (e != cudaSuccess) {
        cudaGetLastError();
        return feast_fail(ctx, e == cudaErrorMemoryAllocation ? FEAST_ERR_OOM : FEAST_ERR_CUDA,
                          "cudaMalloc of %zu bytes failed: %s", sizeof(T) * count, cudaGetErrorString(e));
    }
    // FEAST_POISON=1 fills every allocation with 0xFF bytes (NaN doubles, negative ints) so
    // that a read of uninitialised device memory shows up deterministically in the tests.
    static const bool poison = getenv("FEAST_POISON") != nullptr;
    if (poison) { cudaMemset(q, 0xFF, sizeof(T) * (count ? count : 1)); cudaDeviceSynchronize(); }
    *p = (T*)q;
    return 0;
}
template <typename T>
void dev_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

bool create_ctx_events(feast_ctx* ctx) {   // on the context's device (cudaSetDevice was called by the caller)
    for (auto& e : ctx->evn) if (cudaEventCreate(&e) != cudaSuccess) return false;
    for (auto& e : ctx->evk) if (cudaEventCreate(&e) != cudaSuccess) return false;
    return true;
}

int bind_device(feast_ctx* ctx) {
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return 0;
}

void free_operator(Operator& op) {
    dev_free(op.dense);
    dev_free(op.uvals_r);
    dev_free(op.uvals_c);
    op = Operator();
}

void drop_stored_factors(feast_ctx* ctx) {
    for (auto& f : ctx->stored) { dev_free(f.lu); dev_free(f.ipiv); dev_free(f.perm); dev_free(f.dinv); }
    ctx->stored.clear();
    for (auto& f : ctx->bstored) band_free(f);
    ctx->bstored.clear();
}

// everything derived from the sparsity pattern and the slot values (device layout of the sparse path)
void free_pattern(feast_ctx* ctx) {
    amg_free(ctx);               // level 0 aliases the union pattern freed below
    dev_free(ctx->u_rowptr);
    dev_free(ctx->u_col);
    dev_free(ctx->perm_d);
    dev_free(ctx->t_ptr);
    dev_free(ctx->t_hptr);
    dev_free(ctx->t_hidx);
    dev_free(ctx->u_lcol);
    ctx->reordered = false;
    ctx->tiles_ok = false;
    ctx->ntiles = 0;
    dev_free(ctx->zvals);
    dev_free(ctx->zvals_pc);
    for (int i = 0; i < FEAST_MAX_SLOTS; ++i) { dev_free(ctx->ops[i].uvals_r); dev_free(ctx->ops[i].uvals_c); }
}

void free_problem_derived(feast_ctx* ctx) {
    free_pattern(ctx);
    dev_free(ctx->zdense);
    dev_free(ctx->zpiv);
    dev_free(ctx->zdinv);
    drop_stored_factors(ctx);
    band_free(ctx->bscratch);
    dev_free(ctx->band_tmp);
    ctx->band_tmp_elems = 0;
    ctx->problem_ready = false;
}

void free_blocks(feast_ctx* ctx) {
    BlockVec* all[] = {&ctx->Q, &ctx->X, &ctx->R, &ctx->Q1, &ctx->W1, &ctx->W2, &ctx->Ql, &ctx->Xl, &ctx->Rl, &ctx->kx, &ctx->kr,
                       &ctx->kp, &ctx->kq, &ctx->ks, &ctx->kt, &ctx->kv, &ctx->krh};
    for (auto* b : all) dev_free(b->p);
    for (auto& b : ctx->mom) dev_free(b.p);
    ctx->mom.clear();
    dev_free(ctx->stage);
    dev_free(ctx->small_d);
    dev_free(ctx->red_d);
    dev_free(ctx->gm_V);
    { char* g = (char*)ctx->gm_small; dev_free(g); ctx->gm_small = nullptr; }
    ctx->gm_restart = 0;
    ctx->red_bytes = 0;
    ctx->m0 = 0;
}

int ensure_block(feast_ctx* ctx, BlockVec& b) {
    if (b.p) return 0;
    return dev_alloc(ctx, &b.p, (size_t)ctx->n * ctx->m0);
}

int ensure_pinned(feast_ctx* ctx, size_t bytes) {
    if (ctx->pinned_bytes >= bytes) return 0;
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    CUDA_TRY(ctx, cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_bytes = bytes;
    return 0;
}

struct PhaseTimer {  // CUDA-event timing of a phase on the library stream
    feast_ctx* ctx; int idx; bool on;
    PhaseTimer(feast_ctx* c, int i) : ctx(c), idx(i), on(true) { cudaEventRecord(ctx->ev0, ctx->stream); }
    double stop() {
        if (!on) return 0.0;
        on = false;
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaEventSynchronize(ctx->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        if (idx >= 0) ctx->phase_ms[idx] += ms;
        return ms;
    }
};

// --------------------------------------------------------------------- operator application
// W = op(slot) * V   (row-major n x m0 blocks)
int apply_slot(feast_ctx* ctx, int slot, const c128* V, c128* W) {
    const Operator& op = ctx->ops[slot];
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    if (op.kind == OP_IDENTITY) {
        CUDA_TRY(ctx, cudaMemcpyAsync(W, V, sizeof(c128) * n * m, cudaMemcpyDeviceToDevice, ctx->stream));
        return 0;
    }
    if (op.kind == OP_DENSE)
        return launch_zgemm(ctx, (int)n, m, n, hc128(1, 0), op.dense, n, 1, false, V, m, 1, hc128(0, 0), W, m, 1);  // dense slots are row-major
    if (op.kind == OP_CSR)
        return launch_spmm(ctx, n, m, ctx->u_rowptr, ctx->u_col, op.uvals_r, op.uvals_c, V, m, W, m, nullptr);
    return feast_fail(ctx, FEAST_ERR_STATE, "operator slot %d is not set", slot);
}

// CSC (possibly 1-based, int64) -> host CSR with sorted columns; also detects S == S^T
int csc_to_host_csr(feast_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval, const void* nzval,
                    int is_complex, int base, HostCSR& out) {
    const int64_t nnz = colptr[n] - base;
    if (nnz < 0) return feast_fail(ctx, -4, "argument 4 invalid: colptr not monotone");
    out.n = n; out.nnz = nnz; out.is_complex = is_complex != 0;
    out.rowptr.assign(n + 1, 0);
    out.col.resize(nnz);
    out.val.resize(nnz);
    for (int64_t e = 0; e < nnz; ++e) {
        const int64_t r = rowval[e] - base;
        if (r < 0 || r >= n) return feast_fail(ctx, -5, "argument 5 invalid: row index out of range");
        out.rowptr[r + 1]++;
    }
    for (int64_t i = 0; i < n; ++i) out.rowptr[i + 1] += out.rowptr[i];
    std::vector<int64_t> cursor(out.rowptr.begin(), out.rowptr.end() - 1);
    const double* rv = (const double*)nzval;
    for (int64_t c = 0; c < n; ++c) {
        for (int64_t e = colptr[c] - base; e < colptr[c + 1] - base; ++e) {
            const int64_t r = rowval[e] - base;
            const int64_t d = cursor[r]++;
            out.col[d] = (int)c;
            out.val[d] = is_complex ? hc128(rv[2 * e], rv[2 * e + 1]) : hc128(rv[e], 0.0);
        }
    }
    // columns visited in increasing order -> each CSR row is already sorted by column.
    // symmetry: compare CSR(S) with CSC(S) == CSR(S^T) when the CSC rows are sorted too
    bool sym = true;
    for (int64_t c = 0; c < n && sym; ++c) {
        const int64_t a0 = colptr[c] - base, a1 = colptr[c + 1] - base;
        if (a1 - a0 != out.rowptr[c + 1] - out.rowptr[c]) { sym = false; break; }
        for (int64_t e = a0; e < a1; ++e) {
            const int64_t d = out.rowptr[c] + (e - a0);
            const hc128 v = is_complex ? hc128(rv[2 * e], rv[2 * e + 1]) : hc128(rv[e], 0.0);
            if (rowval[e] - base != out.col[d] || v != out.val[d]) { sym = false; break; }
        }
    }
    out.symmetric = sym;
    return 0;
}

int effective_solver(const feast_ctx* ctx);

// The rows are renumbered (reorder.cpp) when the inner solves are Krylov: the tiled SpMM is the hot kernel there.
// Direct solvers keep the natural ordering (the banded one relies on it).  FEAST_REORDER=0 disables.
bool want_reorder(const feast_ctx* ctx) {
    static const bool off = getenv("FEAST_REORDER") && atoi(getenv("FEAST_REORDER")) == 0;
    if (off || ctx->storage_dense || ctx->problem == FEAST_PROBLEM_SAMPLED) return false;   // samples change the pattern: natural order
    return effective_solver(ctx) == FEAST_SOLVER_KRYLOV;
}

// The multigrid preconditioner applies to COCG on linear problems whose slots are all real and symmetric.
bool want_amg(const feast_ctx* ctx) {
    static const int64_t min_n = getenv("FEAST_AMG_MIN_N") ? atoll(getenv("FEAST_AMG_MIN_N")) : 20000;
    if (ctx->precond == FEAST_PRECOND_NONE || ctx->storage_dense) return false;
    if (ctx->problem != FEAST_PROBLEM_STANDARD && ctx->problem != FEAST_PROBLEM_GENERALIZED) return false;
    if (effective_solver(ctx) != FEAST_SOLVER_KRYLOV || !ctx->all_symmetric) return false;
    if (ctx->krylov != FEAST_KRYLOV_AUTO && ctx->krylov != FEAST_KRYLOV_COCG) return false;
    for (int s = 0; s < ctx->nslots; ++s)
        if (ctx->ops[s].kind == OP_CSR && ctx->ops[s].host.is_complex) return false;
    return ctx->precond == FEAST_PRECOND_AMG || ctx->n >= min_n;
}

// Build the union CSR pattern over all sparse / identity slots and the per-slot value arrays.
int build_union(feast_ctx* ctx) {
    const int64_t n = ctx->n;
    std::vector<int64_t> rowptr(n + 1, 0);
    std::vector<int> col;
    // pass 1: merged pattern row by row
    std::vector<const HostCSR*> hs;
    bool any_identity = false;
    bool all_sym = true;
    for (int s = 0; s < ctx->nslots; ++s) {
        if (ctx->ops[s].kind == OP_CSR) { hs.push_back(&ctx->ops[s].host); all_sym = all_sym && ctx->ops[s].host.symmetric; }
        else if (ctx->ops[s].kind == OP_IDENTITY) any_identity = true;
        else return feast_fail(ctx, FEAST_ERR_STATE, "slot %d is not set (sparse problem)", s);
    }
    ctx->all_symmetric = all_sym;
    size_t reserve = 0;
    for (auto* h : hs) reserve = std::max(reserve, (size_t)h->nnz);
    col.reserve(reserve + (any_identity ? n : 0));
    std::vector<int64_t> pos(hs.size());
    for (int64_t i = 0; i < n; ++i) {
        for (size_t k = 0; k < hs.size(); ++k) pos[k] = hs[k]->rowptr[i];
        bool diag_pending = any_identity;
        while (true) {
            int best = INT32_MAX;
            for (size_t k = 0; k < hs.size(); ++k)
                if (pos[k] < hs[k]->rowptr[i + 1]) best = std::min(best, hs[k]->col[pos[k]]);
            if (diag_pending && (int)i < best) best = (int)i;
            if (best == INT32_MAX) break;
            if (best == (int)i) diag_pending = false;
            col.push_back(best);
            for (size_t k = 0; k < hs.size(); ++k)
                if (pos[k] < hs[k]->rowptr[i + 1] && hs[k]->col[pos[k]] == best) pos[k]++;
        }
        rowptr[i + 1] = (int64_t)col.size();
    }
    int bw = 0;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) bw = std::max(bw, std::abs((int)i - col[e]));
    ctx->bandwidth = bw;
    const int64_t nnz_nat = (int64_t)col.size();
    // per-slot values on the natural union pattern (the device layout is gathered from them at the end)
    std::vector<std::vector<double>> nat_r(ctx->nslots);
    std::vector<std::vector<hc128>> nat_c(ctx->nslots);
    for (int s = 0; s < ctx->nslots; ++s) {
        Operator& op = ctx->ops[s];
        const bool cplx = (op.kind == OP_CSR) && op.host.is_complex;
        if (cplx) nat_c[s].assign(nnz_nat, hc128(0, 0)); else nat_r[s].assign(nnz_nat, 0.0);
        if (op.kind == OP_IDENTITY) {
            for (int64_t i = 0; i < n; ++i) {
                auto it = std::lower_bound(col.begin() + rowptr[i], col.begin() + rowptr[i + 1], (int)i);
                nat_r[s][it - col.begin()] = 1.0;
            }
            op.symmetric = true;
        } else {
            const HostCSR& h = op.host;
            for (int64_t i = 0; i < n; ++i) {
                int64_t u = rowptr[i];
                for (int64_t e = h.rowptr[i]; e < h.rowptr[i + 1]; ++e) {
                    while (col[u] != h.col[e]) ++u;
                    if (cplx) nat_c[s][u] = h.val[e]; else nat_r[s][u] = h.val[e].real();
                }
            }
            op.symmetric = h.symmetric;
        }
        op.is_complex = cplx;
    }
    // The multigrid hierarchy depends on the natural pattern and values only: its host setup (0.6 s at n = 1e6) runs on a
    // helper thread while this one builds the tile plan and uploads the device layout; joined before the hierarchy's upload.
    const bool amg_wanted = want_amg(ctx);
    AmgHost amg_host;
    std::thread amg_thread;
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } amg_joiner{amg_thread};
    const double* amg_vp[FEAST_MAX_SLOTS] = {};
    if (amg_wanted) {
        for (int s = 0; s < ctx->nslots; ++s) amg_vp[s] = nat_r[s].data();
        amg_thread = std::thread([&] { amg_setup_host(n, rowptr.data(), col.data(), ctx->nslots, amg_vp, amg_max_coarse(), amg_host); });
    }

    // Device layout of the union pattern: rows (re)ordered by the tile plan and every row PADDED to a multiple of
    // 8 entries (column -1, value 0 in every slot), so that each row's 16-bit tile-local column numbers are one
    // aligned 16-byte word per 8 entries and every tile's CSR slice is 16-byte aligned for the bulk copies
    // (spmm.cu).  src[e] = entry of the natural union pattern behind device entry e (-1 for padding).
    ctx->reordered = false;
    ctx->tiles_ok = false;
    const TileCaps caps = spmm_tile_caps();
    ctx->tile_cfg = spmm_tile_cfg();
    TilePlan plan;
    const bool reorder = want_reorder(ctx);
    build_tile_order(n, rowptr.data(), col.data(), reorder, caps, plan);
    std::vector<int> inv;
    if (reorder) {
        inv.resize(n);
        for (int64_t i = 0; i < n; ++i) inv[plan.order[i]] = (int)i;
        ctx->reordered = true;
    }
    std::vector<int64_t> rowptr_f(n + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t old = reorder ? plan.order[i] : i;
        rowptr_f[i + 1] = rowptr_f[i] + ((rowptr[old + 1] - rowptr[old] + 7) & ~(int64_t)7);
    }
    const int64_t unnz = rowptr_f[n];
    if (unnz > INT32_MAX) return feast_fail(ctx, FEAST_ERR_STATE, "union pattern exceeds 2^31 nonzeros");
    ctx->unnz = unnz;
    std::vector<int> col_f((size_t)unnz, -1), src((size_t)unnz, -1);
    {
        std::vector<std::pair<int, int>> rowbuf;
        for (int64_t i = 0; i < n; ++i) {
            const int64_t old = reorder ? plan.order[i] : i;
            int64_t d = rowptr_f[i];
            if (reorder) {
                rowbuf.clear();
                for (int64_t e = rowptr[old]; e < rowptr[old + 1]; ++e) rowbuf.emplace_back(inv[col[e]], (int)e);
                std::sort(rowbuf.begin(), rowbuf.end());
                for (auto& pr : rowbuf) { col_f[d] = pr.first; src[d] = pr.second; ++d; }
            } else {
                for (int64_t e = rowptr[old]; e < rowptr[old + 1]; ++e) { col_f[d] = col[e]; src[d] = (int)e; ++d; }
            }
        }
    }
    std::vector<uint16_t> lcol;
    if (plan.ok && build_tile_halo(n, rowptr_f.data(), col_f.data(), plan, lcol) == 0) ctx->tiles_ok = true;
    ctx->halo_ratio = plan.halo_ratio;
    ctx->ntiles = (int)plan.tile_ptr.size() - 1;

    std::vector<int> rp32(n + 1);
    for (int64_t i = 0; i <= n; ++i) rp32[i] = (int)rowptr_f[i];
    FEAST_TRY(dev_alloc(ctx, &ctx->u_rowptr, n + 1));
    FEAST_TRY(dev_alloc(ctx, &ctx->u_col, unnz));
    // all uploads are ordered on the library stream: it is a NON-BLOCKING stream, so legacy-stream
    // copies from pageable memory would not be ordered against the kernels launched on it
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->u_rowptr, rp32.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->u_col, col_f.data(), sizeof(int) * unnz, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->reordered) {
        FEAST_TRY(dev_alloc(ctx, &ctx->perm_d, n));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->perm_d, plan.order.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (ctx->tiles_ok) {
        FEAST_TRY(dev_alloc(ctx, &ctx->t_ptr, plan.tile_ptr.size()));
        FEAST_TRY(dev_alloc(ctx, &ctx->t_hptr, plan.halo_ptr.size()));
        FEAST_TRY(dev_alloc(ctx, &ctx->t_hidx, plan.halo_idx.size() + 1));
        FEAST_TRY(dev_alloc(ctx, &ctx->u_lcol, unnz));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->t_ptr, plan.tile_ptr.data(), sizeof(int) * plan.tile_ptr.size(), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->t_hptr, plan.halo_ptr.data(), sizeof(int) * plan.halo_ptr.size(), cudaMemcpyHostToDevice, ctx->stream));
        if (!plan.halo_idx.empty())
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->t_hidx, plan.halo_idx.data(), sizeof(int) * plan.halo_idx.size(), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->u_lcol, lcol.data(), sizeof(uint16_t) * unnz, cudaMemcpyHostToDevice, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    // pass 2: per-slot values gathered from the natural union pattern into the device layout
    for (int s = 0; s < ctx->nslots; ++s) {
        Operator& op = ctx->ops[s];
        if (op.is_complex) {
            std::vector<hc128> cv2((size_t)unnz, hc128(0, 0));
            for (int64_t e = 0; e < unnz; ++e) if (src[e] >= 0) cv2[e] = nat_c[s][src[e]];
            FEAST_TRY(dev_alloc(ctx, &op.uvals_c, unnz));
            CUDA_TRY(ctx, cudaMemcpyAsync(op.uvals_c, cv2.data(), sizeof(c128) * unnz, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        } else {
            std::vector<double> rv2((size_t)unnz, 0.0);
            for (int64_t e = 0; e < unnz; ++e) if (src[e] >= 0) rv2[e] = nat_r[s][src[e]];
            FEAST_TRY(dev_alloc(ctx, &op.uvals_r, unnz));
            CUDA_TRY(ctx, cudaMemcpyAsync(op.uvals_r, rv2.data(), sizeof(double) * unnz, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        }
        // host copies are kept so that feast_set_problem can be called again (e.g. switching the
        // problem kind); an identity slot keeps its kind and additionally lives on the union pattern
    }
    FEAST_TRY(dev_alloc(ctx, &ctx->zvals, unnz));
    ctx->amg_why.clear();
    if (amg_wanted) {   // multigrid hierarchy of the Krylov preconditioner: join the host setup, upload (amg.cu)
        std::vector<int> dpos0(n, 0);
        for (int64_t i = 0; i < n; ++i)
            for (int64_t e = rowptr_f[i]; e < rowptr_f[i + 1]; ++e)
                if (col_f[e] == (int)i) dpos0[i] = (int)e;
        amg_thread.join();
        FEAST_TRY(amg_build(ctx, amg_host, ctx->nslots, reorder ? plan.order : std::vector<int>(), dpos0, &ctx->amg_why));
    }
    return 0;
}

int effective_solver(const feast_ctx* ctx) {
    if (ctx->solver != FEAST_SOLVER_AUTO) return ctx->solver;
    if (ctx->storage_dense) return FEAST_SOLVER_DENSE_LU;
    if (ctx->n <= ctx->dense_threshold) return FEAST_SOLVER_DENSE_LU;
    // non-symmetric banded operators (2-D discretisations): restarted Krylov stagnates on the shifted
    // systems, the block-tridiagonal direct solver does not
    if (!ctx->all_symmetric && ctx->bandwidth > 0 && ctx->bandwidth <= 2048) return FEAST_SOLVER_BANDED_LU;
    return FEAST_SOLVER_KRYLOV;
}
// the device layout (row order, tile plan, multigrid hierarchy) follows the solver settings: rebuild it from the host
// copies of the operators; the subspace blocks are dropped
int rebuild_layout(feast_ctx* ctx) {
    FEAST_TRY(bind_device(ctx));
    const int nslots = ctx->nslots;
    free_problem_derived(ctx);
    if (ctx->m0 != 0) free_blocks(ctx);
    ctx->nslots = nslots;
    FEAST_TRY(build_union(ctx));
    ctx->problem_ready = true;
    return 0;
}

int effective_krylov(const feast_ctx* ctx) {
    if (ctx->krylov != FEAST_KRYLOV_AUTO) return ctx->krylov;
    return ctx->all_symmetric ? FEAST_KRYLOV_COCG : FEAST_KRYLOV_GMRES;
}

// Z = sum_i coef[i] * slot_i, assembled either dense (col-major n x n) or on the union pattern
int assemble_dense_Z(feast_ctx* ctx, const hc128* coef, c128* Z) {
    const int64_t n = ctx->n;
    if (ctx->storage_dense) {
        const c128* D[FEAST_MAX_SLOTS];
        int kinds[FEAST_MAX_SLOTS];
        for (int s = 0; s < ctx->nslots; ++s) { D[s] = ctx->ops[s].dense; kinds[s] = ctx->ops[s].kind; }
        return launch_assemble_dense(ctx, n, ctx->nslots, D, kinds, coef, Z);
    }
    const double* rv[FEAST_MAX_SLOTS];
    const c128* cv[FEAST_MAX_SLOTS];
    for (int s = 0; s < ctx->nslots; ++s) { rv[s] = ctx->ops[s].uvals_r; cv[s] = ctx->ops[s].uvals_c; }
    FEAST_TRY(launch_assemble_union(ctx, ctx->unnz, ctx->nslots, rv, cv, coef, ctx->zvals));
    return launch_scatter_dense(ctx, n, ctx->u_rowptr, ctx->u_col, ctx->zvals, Z);
}
int assemble_sparse_Z(feast_ctx* ctx, const hc128* coef, c128* zvals) {
    const double* rv[FEAST_MAX_SLOTS];
    const c128* cv[FEAST_MAX_SLOTS];
    for (int s = 0; s < ctx->nslots; ++s) { rv[s] = ctx->ops[s].uvals_r; cv[s] = ctx->ops[s].uvals_c; }
    return launch_assemble_union(ctx, ctx->unnz, ctx->nslots, rv, cv, coef, zvals);
}

void node_coefs(const feast_ctx* ctx, hc128 z, hc128* coef) {
    if (ctx->problem == FEAST_PROBLEM_POLYNOMIAL) {
        hc128 p(1, 0);
        for (int s = 0; s < ctx->nslots; ++s) { coef[s] = p; p *= z; }
    } else if (ctx->problem == FEAST_PROBLEM_SAMPLED) {   // slot 0 IS T(z_k), evaluated by the caller
        coef[0] = hc128(1, 0);
    } else {  // A - z B   (src/feast.jl:64,141)
        coef[0] = hc128(1, 0);
        coef[1] = -z;
    }
}

int ensure_krylov_work(feast_ctx* ctx, int method) {
    if (method == FEAST_KRYLOV_GMRES) {
        if (ctx->gm_V) return 0;
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
        const size_t blk = sizeof(c128) * (size_t)ctx->n * ctx->m0;
        int64_t R = (int64_t)(0.6 * (double)free_b / (double)blk) - 1;   // basis gets at most 60 % of the free HBM
        if (R > 400) R = 400;   // long recurrences are affordable in 180 GB of HBM and avoid restart stagnation
        if (R > ctx->n) R = ctx->n;
        if (R < 4) return feast_fail(ctx, FEAST_ERR_OOM, "not enough device memory for a GMRES basis");
        ctx->gm_restart = (int)R;
        FEAST_TRY(dev_alloc(ctx, &ctx->gm_V, (size_t)(R + 1) * ctx->n * ctx->m0));
        FEAST_TRY(dev_alloc(ctx, (char**)&ctx->gm_small, gmres_small_bytes(ctx->m0, (int)R)));
        return 0;
    }
    FEAST_TRY(ensure_block(ctx, ctx->kr));
    FEAST_TRY(ensure_block(ctx, ctx->kp));
    FEAST_TRY(ensure_block(ctx, ctx->kq));
    if (ctx->amg && method == FEAST_KRYLOV_COCG) {   // preconditioned COCG: z and the cycle's scratch, coarse-level blocks
        FEAST_TRY(ensure_block(ctx, ctx->ks));
        FEAST_TRY(ensure_block(ctx, ctx->kt));
        FEAST_TRY(amg_ensure_blocks(ctx));
    }
    if (method == FEAST_KRYLOV_BICGSTAB) {
        FEAST_TRY(ensure_block(ctx, ctx->krh));
        FEAST_TRY(ensure_block(ctx, ctx->kv));
        FEAST_TRY(ensure_block(ctx, ctx->ks));
        FEAST_TRY(ensure_block(ctx, ctx->kt));
    }
    return 0;
}

int factor_dense(feast_ctx* ctx, const hc128* coef, DenseLU& f, int* info) {
    const int64_t n = ctx->n;
    f.n = n;
    if (!f.lu) FEAST_TRY(dev_alloc(ctx, &f.lu, (size_t)n * n));
    if (!f.ipiv) FEAST_TRY(dev_alloc(ctx, &f.ipiv, n));
    if (!f.perm) FEAST_TRY(dev_alloc(ctx, &f.perm, n));
    if (!f.dinv) FEAST_TRY(dev_alloc(ctx, &f.dinv, (size_t)2 * n * kDiagNB));
    FEAST_TRY(assemble_dense_Z(ctx, coef, f.lu));
    FEAST_TRY(dense_getrf(ctx, n, f.lu, f.ipiv, info));
    FEAST_TRY(dense_build_perm(ctx, n, f.ipiv, f.perm));
    FEAST_TRY(dense_build_diag_inverses(ctx, n, f.lu, f.dinv));
    return 0;
}

// Iterated Cholesky-QR with column scaling; V <- orth(V), V_in = V_out * Rtot.
// A pass factors the column-scaled Gram matrix G = D^-1 V^H V D^-1 (unit diagonal).  While G is far from the
// identity (or its Cholesky factorisation meets a non-positive pivot) the pass is a SHIFTED Cholesky-QR (Fukaya,
// Kannan, Nakatsukasa, Yamamoto, Yanagisawa 2020): G + s I with s = 11 (m n + m (m + 1)) u ||V D^-1||_2^2, which
// bounds ||R^-1|| by s^-1/2 and brings cond(V) down to ~1e3 per pass for any numerical rank, keeping V_old = V_new R
// backward stable; two or three plain passes then finish.  (Clamping individual pivots instead -- the first version --
// let R^-1 grow without bound: on a filtered FEAST block with singular values 3.7 ... 1e-9 the reconstruction error
// V_out Rtot - V_in reached 1e17, which is what stalled C4 at n = 250 000.)
int orthonormalize(feast_ctx* ctx, BlockVec& V, std::vector<hc128>* Rtot_out) {
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    FEAST_TRY(ensure_block(ctx, ctx->W1));
    FEAST_TRY(ensure_pinned(ctx, sizeof(hc128) * (size_t)m * m + 256));
    std::vector<hc128> G((size_t)m * m), Ri, RD, Rtot, tmp;
    if (Rtot_out) {
        Rtot.assign((size_t)m * m, hc128(0, 0));
        for (int j = 0; j < m; ++j) Rtot[(size_t)j * m + j] = 1.0;
    }
    c128* G_d = ctx->small_d;
    c128* M_d = ctx->small_d + (size_t)m * m;
    const int maxpass = 10;
    double prev_err = -1.0;
    for (int pass = 0; pass < maxpass; ++pass) {
        FEAST_TRY(launch_gram(ctx, n, m, V.p, V.p, G_d));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pinned, G_d, sizeof(c128) * m * m, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(G.data(), ctx->pinned, sizeof(hc128) * m * m);
        if (cholqr_pass(m, (double)n, G, Ri, RD, prev_err)) break;               // orthonormal to rounding (host_small.h)
        // V_new = V * D^-1 * R^-1
        CUDA_TRY(ctx, cudaMemcpyAsync(M_d, Ri.data(), sizeof(c128) * m * m, cudaMemcpyHostToDevice, ctx->stream));
        FEAST_TRY(launch_update(ctx, n, m, V.p, M_d, ctx->W1.p));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // Ri is a host temporary
        std::swap(V.p, ctx->W1.p);
        if (Rtot_out) {  // Rtot <- R * D * Rtot
            matmul_small(m, RD, Rtot, tmp);
            Rtot.swap(tmp);
        }
    }
    if (Rtot_out) Rtot_out->swap(Rtot);
    return 0;
}

int check_ready(feast_ctx* ctx, bool need_subspace) {
    if (!ctx) return feast_fail(nullptr, -1, "argument 1 invalid: null context");
    if (!ctx->problem_ready) return feast_fail(ctx, FEAST_ERR_STATE, "feast_set_problem has not been called");
    if (need_subspace && ctx->m0 == 0) return feast_fail(ctx, FEAST_ERR_STATE, "feast_set_subspace has not been called");
    return bind_device(ctx);
}

int download_block(feast_ctx* ctx, const BlockVec& b, feast_c128* H, int64_t ld) {
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    if (!b.p) return feast_fail(ctx, FEAST_ERR_STATE, "requested block has not been computed yet");
    FEAST_TRY(launch_rowmajor_to_colmajor(ctx, n, m, b.p, ctx->stage, n, ctx->perm_d));
    CUDA_TRY(ctx, cudaMemcpy2DAsync(H, sizeof(c128) * ld, ctx->stage, sizeof(c128) * n, sizeof(c128) * n, m,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Longest-processing-time re-sharding of the contour nodes from the costs measured in the previous
// pass (identical on every rank: the cost vector is all-reduced).  Near-axis / near-spectrum nodes
// take 2-3x more Krylov iterations than the others, and which ones depends on the spectrum, so a
// static map leaves GPUs idle at the all-reduce (measured: 1.8 s of a 3.2 s step at 8 GPUs).
void rebalance_nodes(feast_ctx* ctx) {
    const int nn = (int)ctx->owner.size(), nr = ctx->nranks;
    std::vector<int> order(nn);
    for (int k = 0; k < nn; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ctx->node_cost[a] > ctx->node_cost[b]; });
    std::vector<double> load(nr, 0.0);
    for (int k : order) {
        int best = 0;
        for (int r = 1; r < nr; ++r) if (load[r] < load[best]) best = r;
        ctx->owner[k] = best;
        load[best] += ctx->node_cost[k];
    }
}

// Column sharding in GROUPS (Krylov inner solves, several ranks): the ranks are split into G groups of nranks / G; a node
// belongs to one group, whose ranks each solve m0 / (nranks / G) of its right-hand-side columns.  G = 1 is the pure
// column split (balanced by construction, used in the first pass when no costs are known); larger G means wider, more
// efficient column slices (the per-iteration time has a fixed part: coarse multigrid levels, launch latencies --
// measured 7.15 ms at 64 columns, 4.75 ms at 32) but needs the node costs to balance the groups.  From the costs
// measured in the previous pass every divisor G of nranks is scored by
//     (LPT makespan over G groups) x (columns per rank + fixed), fixed = 12 columns (FEAST_SHARD_FIXED_COLS)
// and the best one is taken; identical on every rank (the cost vector is all-reduced).
int pick_groups(int nn, const double* cost, int nr, int m0, double fixed_cols, int forced, std::vector<int>& gowner) {
    std::vector<int> order(nn);
    for (int k = 0; k < nn; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    double best = -1.0;
    int bestG = 1;
    gowner.assign(nn, 0);
    for (int G = 1; G <= nr; ++G) {
        if (nr % G != 0 || (forced && G != forced)) continue;
        const int gs = nr / G;
        if (m0 < 2 * gs && gs > 1) continue;
        std::vector<double> load(G, 0.0);
        std::vector<int> own(nn, 0);
        for (int k : order) {
            int b = 0;
            for (int g = 1; g < G; ++g) if (load[g] < load[b]) b = g;
            own[k] = b;
            load[b] += cost[k];
        }
        const double score = *std::max_element(load.begin(), load.end()) * ((double)m0 / gs + fixed_cols);
        if (best < 0.0 || score < best) { best = score; bestG = G; gowner = own; }
    }
    return bestG;
}

void choose_groups(feast_ctx* ctx) {
    static const double fixed_cols = getenv("FEAST_SHARD_FIXED_COLS") ? atof(getenv("FEAST_SHARD_FIXED_COLS")) : 12.0;
    static const int forced = getenv("FEAST_SHARD_GROUPS") ? atoi(getenv("FEAST_SHARD_GROUPS")) : 0;
    const int nn = (int)ctx->znodes.size();
    // first pass: no measurements yet -- a static model of the Krylov cost of a node, the inverse distance of the node
    // from the real axis through the contour centre (the same model as partition.node_cost of the Python binding).  On
    // the C2 contour its LPT groups are within 13 % of the balance reached with the measured costs (148 vs 131 iterations
    // per group at 8 ranks), whereas the pure column split of round 2's first version made the first pass 5x slower than
    // the later ones at 8 ranks (2.9 s vs 0.5 s: at 8 columns per rank the iteration is launch / latency bound).
    std::vector<double> model(nn, 1.0);
    if (!ctx->have_costs) {
        hc128 c(0, 0);
        for (auto& z : ctx->znodes) c += z;
        c /= (double)nn;
        double r = 0.0;
        for (auto& z : ctx->znodes) r = std::max(r, std::abs(z - c));
        for (int k = 0; k < nn; ++k) model[k] = 1.0 / (std::fabs(ctx->znodes[k].imag() - c.imag()) / (r > 0 ? r : 1.0) + 0.15);
    }
    ctx->ngroups = pick_groups(nn, ctx->have_costs ? ctx->node_cost.data() : model.data(), ctx->nranks, ctx->m0, fixed_cols, forced, ctx->gowner);
}

// W = op(slot)^H * V.  Dense: one DMMA GEMM on the conjugate-transposed view.  Sparse: supported when
// the slot is (complex-)symmetric, op^H = conj(op): W = conj(op * conj(V)).
int apply_slot_adjoint(feast_ctx* ctx, int slot, const c128* V, c128* W) {
    const Operator& op = ctx->ops[slot];
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    if (op.kind == OP_IDENTITY) {
        CUDA_TRY(ctx, cudaMemcpyAsync(W, V, sizeof(c128) * n * m, cudaMemcpyDeviceToDevice, ctx->stream));
        return 0;
    }
    if (op.kind == OP_DENSE)   // A^H(i,k) = conj(A(k,i)); dense slots are row-major
        return launch_zgemm(ctx, (int)n, m, n, hc128(1, 0), op.dense, 1, n, true, V, m, 1, hc128(0, 0), W, m, 1);
    if (op.kind == OP_CSR) {
        if (!op.symmetric)
            return feast_fail(ctx, FEAST_ERR_STATE, "adjoint of a non-symmetric sparse operator is not supported: pass it dense");
        if (!op.is_complex) return apply_slot(ctx, slot, V, W);   // real symmetric: A^H = A
        FEAST_TRY(ensure_block(ctx, ctx->W2));
        FEAST_TRY(launch_conj(ctx, n * m, V, ctx->W2.p));
        FEAST_TRY(apply_slot(ctx, slot, ctx->W2.p, W));
        return launch_conj(ctx, n * m, W, W);
    }
    return feast_fail(ctx, FEAST_ERR_STATE, "operator slot %d is not set", slot);
}

// Krylov inner solve; COCG with complex64 block storage when mixed precision was requested and is applicable
int krylov_any(feast_ctx* ctx, int method, const hc128* coef, const c128* zvals, const c128* rhs, c128* Y, KrylovResult* kr,
               feast_stats* st, int node = -1) {
    if (ctx->amg && method == FEAST_KRYLOV_COCG && !ctx->mixed_prec) {
        int info = 0;
        // Complex-shifted preconditioner (the "shifted Laplacian" idea): the hierarchy is assembled for
        // z~ = z + i beta |z| sign(Im z) instead of z.  For shifts inside the spectrum (interior slices: many eigenmodes
        // below Re z that the coarse levels no longer resolve) the cycle for the damped operator is a much better
        // preconditioner of the true one (numpy prototype at 40^3, third slice: 1041 -> 372 iterations on the worst node);
        // for the lowest slice beta = 0 is best.  coef[1] carries -z for the linear problems this path serves.
        const c128* zpc = zvals;
        hc128 cpc[FEAST_MAX_SLOTS];
        for (int s = 0; s < FEAST_MAX_SLOTS; ++s) cpc[s] = coef[s];
        if (ctx->precond_shift != 0.0 && ctx->problem != FEAST_PROBLEM_POLYNOMIAL) {
            const hc128 z = -coef[1] / coef[0];
            const hc128 zt = z + hc128(0.0, ctx->precond_shift * std::abs(z) * (z.imag() < 0 ? -1.0 : 1.0));
            cpc[1] = -zt * coef[0];
            if (!ctx->zvals_pc) FEAST_TRY(dev_alloc(ctx, &ctx->zvals_pc, ctx->unnz));
            FEAST_TRY(assemble_sparse_Z(ctx, cpc, ctx->zvals_pc));
            zpc = ctx->zvals_pc;
        }
        FEAST_TRY(amg_assemble(ctx, cpc, zpc, node, &info));
        if (info) return feast_fail(ctx, FEAST_ERR_SINGULAR, "zero pivot at column %d of the coarsest multigrid operator", info);
        if (st) amg_info(ctx, &st->precond_levels, nullptr, 0, nullptr);
        FEAST_TRY(krylov_solve_pcocg(ctx, zvals, zpc, rhs, Y, ctx->inner_tol, ctx->max_inner, kr));
        if (kr->converged || !(kr->relres_max > 0.1)) return 0;
        // Safety net: the aggregation hierarchy assumes an elliptic-type slot 0; on an operator it does not suit, the cycle
        // can fail to reduce the residual at all (relative residual still above 0.1 at max_inner).  Redo the node with the
        // unpreconditioned recurrence rather than hand back a useless solve.
        KrylovResult plain;
        FEAST_TRY(krylov_solve(ctx, method, zvals, rhs, Y, ctx->inner_tol, ctx->max_inner, &plain));
        plain.iters += kr->iters;
        plain.spmm_ms += kr->spmm_ms;
        plain.spmm_launches += kr->spmm_launches;
        *kr = plain;
        if (st) st->precond_levels = 0;
        return 0;
    }
    if (ctx->mixed_prec && method == FEAST_KRYLOV_COCG && (ctx->m0 % 2) == 0 && ctx->m0 <= 128 && ctx->tiles_ok && ctx->tile_cfg == 0)
        return krylov_solve_mixed(ctx, zvals, rhs, Y, ctx->inner_tol, ctx->max_inner, kr);
    return krylov_solve(ctx, method, zvals, rhs, Y, ctx->inner_tol, ctx->max_inner, kr);
}

// One shifted solve  (sum_i coef[i] slot_i) Y = rhs  with m0 right-hand sides: dense LU (stored per node
// when store != 0, node index k >= 0) or Krylov on the union pattern.  `e1` is recorded between the
// factorisation/assembly and the solve.  This is linsolve! (src/utils.jl:175-179) for one contour node.
int solve_shifted(feast_ctx* ctx, int solver, int method, int k, const hc128* coef, const c128* rhs, c128* Y,
                  feast_stats& st, cudaEvent_t e1, int* rc_final, bool adjoint = false) {
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    if (solver == FEAST_SOLVER_DENSE_LU) {
        DenseLU* f;
        DenseLU scratch;
        bool need_factor = true;
        if (ctx->store && k >= 0) {
            f = &ctx->stored[k];
            need_factor = (f->lu == nullptr);
        } else {
            if (!ctx->zdense) FEAST_TRY(dev_alloc(ctx, &ctx->zdense, (size_t)n * n));
            if (!ctx->zpiv) FEAST_TRY(dev_alloc(ctx, &ctx->zpiv, (size_t)2 * n));
            if (!ctx->zdinv) FEAST_TRY(dev_alloc(ctx, &ctx->zdinv, (size_t)2 * n * kDiagNB));
            scratch.lu = ctx->zdense; scratch.ipiv = ctx->zpiv; scratch.perm = ctx->zpiv + n; scratch.dinv = ctx->zdinv;
            f = &scratch;
        }
        if (k == -2) need_factor = false;   // the scratch factorisation of the previous call is reused (adjoint solve)
        if (need_factor) {
            int info = 0;
            FEAST_TRY(factor_dense(ctx, coef, *f, &info));                                // feast.jl:36 / :65 lu
            if (info && !st.info) st.info = info;
        }
        if (e1) cudaEventRecord(e1, ctx->stream);
        FEAST_TRY(ensure_block(ctx, ctx->W2));
        FEAST_TRY(dense_getrs(ctx, n, f->lu, f->perm, f->dinv, m, rhs, Y, adjoint));      // ldiv! (or F' \\ R)
    } else if (solver == FEAST_SOLVER_BANDED_LU) {
        if (adjoint) return feast_fail(ctx, FEAST_ERR_STATE, "adjoint solves are not available with the banded solver");
        if (ctx->storage_dense) return feast_fail(ctx, FEAST_ERR_STATE, "the banded solver needs sparse operators");
        FEAST_TRY(assemble_sparse_Z(ctx, coef, ctx->zvals));
        BandFactor* bf = &ctx->bscratch;
        bool need_factor = true;
        if (ctx->store && k >= 0) {
            if ((int)ctx->bstored.size() != (int)ctx->znodes.size()) ctx->bstored.resize(ctx->znodes.size());
            bf = &ctx->bstored[k];
            need_factor = (bf->lu == nullptr);
        }
        if (need_factor) {
            int info = 0;
            FEAST_TRY(band_factor(ctx, ctx->zvals, *bf, &info));
            if (info && !st.info) st.info = info;
        }
        if (e1) cudaEventRecord(e1, ctx->stream);
        FEAST_TRY(ensure_block(ctx, ctx->W2));
        int steps = 0;
        double rel = 0.0;
        FEAST_TRY(band_solve_refined(ctx, *bf, ctx->zvals, m, rhs, Y, ctx->W2.p, &steps, &rel));
        st.inner_iters_total += steps;                                  // refinement steps
        st.inner_iters_max = std::max(st.inner_iters_max, steps);
        st.inner_relres_max = std::max(st.inner_relres_max, rel);
        if (!(rel <= ctx->inner_tol)) *rc_final = FEAST_WARN_INNER_MAXIT;   // stagnated or non-finite refinement: tell the caller
    } else {
        FEAST_TRY(assemble_sparse_Z(ctx, coef, ctx->zvals));
        if (e1) cudaEventRecord(e1, ctx->stream);
        KrylovResult kr;
        if (adjoint) {
            // Z complex symmetric: Z^H = conj(Z), so Z^H y = b  <=>  Z conj(y) = conj(b)
            if (!ctx->all_symmetric)
                return feast_fail(ctx, FEAST_ERR_STATE, "adjoint Krylov solve needs symmetric operators: pass them dense");
            FEAST_TRY(ensure_block(ctx, ctx->W2));
            FEAST_TRY(launch_conj(ctx, n * m, rhs, ctx->W2.p));
            FEAST_TRY(krylov_any(ctx, method, coef, ctx->zvals, ctx->W2.p, Y, &kr, &st, k));   // same Z (and cached coarse inverse), conjugated data
            FEAST_TRY(launch_conj(ctx, n * m, Y, Y));
        } else {
            FEAST_TRY(krylov_any(ctx, method, coef, ctx->zvals, rhs, Y, &kr, &st, k));
        }
        st.inner_iters_total += kr.iters;
        st.inner_iters_max = std::max(st.inner_iters_max, kr.iters);
        st.inner_relres_max = std::max(st.inner_relres_max, kr.relres_max);
        st.t_spmm_ms += kr.spmm_ms;
        st.spmm_launches += kr.spmm_launches;
        if (!kr.converged) *rc_final = FEAST_WARN_INNER_MAXIT;
    }
    return 0;
}

}  // namespace

// =============================================================================== ABI
extern "C" {

int feast_version(void) { return 100; }

int feast_device_count(int* count) {
    if (!count) return -1;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return feast_fail(nullptr, FEAST_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return 0;
}

int feast_ctx_create(feast_ctx** out, int device) {
    if (!out) return feast_fail(nullptr, -1, "argument 1 invalid: null output pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return feast_fail(nullptr, FEAST_ERR_CUDA,
                          "no CUDA device available (%s); libfeast_cuda has no CPU fallback",
                          e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= count) return feast_fail(nullptr, -2, "argument 2 invalid: device %d of %d", device, count);
    feast_ctx* ctx = new feast_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess || !create_ctx_events(ctx)) {
        int rc = feast_fail(nullptr, FEAST_ERR_CUDA, "context creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return rc;
    }
    if (ensure_pinned(ctx, 1 << 20)) { delete ctx; return FEAST_ERR_CUDA; }
    *out = ctx;
    return 0;
}

int feast_ctx_destroy(feast_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->nccl_comm) { const NcclApi* api = nccl_api(); if (api) api->CommDestroy(ctx->nccl_comm); }
    free_problem_derived(ctx);
    for (int i = 0; i < FEAST_MAX_SLOTS; ++i) free_operator(ctx->ops[i]);
    free_blocks(ctx);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (auto& e : ctx->evn) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->evk) if (e) cudaEventDestroy(e);
    if (ctx->sw0) cudaEventDestroy(ctx->sw0);
    if (ctx->sw1) cudaEventDestroy(ctx->sw1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

const char* feast_last_error(const feast_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

// ------------------------------------------------------------------------- operators
int feast_set_dense(feast_ctx* ctx, int slot, int64_t n, const void* a, int64_t lda, int is_complex) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, slot >= 0 && slot < FEAST_MAX_SLOTS, 2, "slot out of range");
    ARG_CHECK(ctx, n > 0, 3, "n must be positive");
    ARG_CHECK(ctx, a != nullptr, 4, "null matrix");
    ARG_CHECK(ctx, lda >= n, 5, "lda < n");
    FEAST_TRY(bind_device(ctx));
    free_problem_derived(ctx);
    free_operator(ctx->ops[slot]);
    Operator& op = ctx->ops[slot];
    // device storage of dense slots is ROW-major (see dense.cu): upload column-major, transpose once
    FEAST_TRY(dev_alloc(ctx, &op.dense, (size_t)n * n));
    c128* cm = nullptr;
    FEAST_TRY(dev_alloc(ctx, &cm, (size_t)n * n));
    int rc = 0;
    cudaError_t e;
    if (is_complex) {
        e = cudaMemcpy2DAsync(cm, sizeof(c128) * n, a, sizeof(c128) * lda, sizeof(c128) * n, n, cudaMemcpyHostToDevice, ctx->stream);
    } else {
        double* tmp = (double*)op.dense;   // reuse the destination as staging for the real upload
        e = cudaMemcpy2DAsync(tmp, sizeof(double) * n, a, sizeof(double) * lda, sizeof(double) * n, n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) rc = launch_real_to_complex(ctx, n * n, tmp, cm);
    }
    if (e == cudaSuccess && !rc) rc = launch_colmajor_to_rowmajor(ctx, n, (int)n, cm, n, op.dense);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(cm);
    if (e != cudaSuccess) return feast_fail(ctx, FEAST_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
    if (rc) return rc;
    op.kind = OP_DENSE; op.n = n; op.is_complex = true;
    return 0;
}

int feast_set_csc(feast_ctx* ctx, int slot, int64_t n, const int64_t* colptr, const int64_t* rowval, const void* nzval,
                  int is_complex, int index_base) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, slot >= 0 && slot < FEAST_MAX_SLOTS, 2, "slot out of range");
    ARG_CHECK(ctx, n > 0 && n < INT32_MAX, 3, "n out of range");
    ARG_CHECK(ctx, colptr != nullptr, 4, "null colptr");
    ARG_CHECK(ctx, rowval != nullptr || colptr[n] == index_base, 5, "null rowval");
    ARG_CHECK(ctx, nzval != nullptr || colptr[n] == index_base, 6, "null nzval");
    ARG_CHECK(ctx, index_base == 0 || index_base == 1, 8, "index_base must be 0 or 1");
    FEAST_TRY(bind_device(ctx));
    free_problem_derived(ctx);
    free_operator(ctx->ops[slot]);
    Operator& op = ctx->ops[slot];
    FEAST_TRY(csc_to_host_csr(ctx, n, colptr, rowval, nzval, is_complex, index_base, op.host));
    op.kind = OP_CSR; op.n = n; op.is_complex = is_complex != 0;
    return 0;
}

int feast_set_identity(feast_ctx* ctx, int slot, int64_t n) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, slot >= 0 && slot < FEAST_MAX_SLOTS, 2, "slot out of range");
    ARG_CHECK(ctx, n > 0, 3, "n must be positive");
    FEAST_TRY(bind_device(ctx));
    free_problem_derived(ctx);
    free_operator(ctx->ops[slot]);
    ctx->ops[slot].kind = OP_IDENTITY;
    ctx->ops[slot].n = n;
    ctx->ops[slot].symmetric = true;
    return 0;
}

int feast_set_problem(feast_ctx* ctx, int kind, int nslots) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, kind >= FEAST_PROBLEM_STANDARD && kind <= FEAST_PROBLEM_SAMPLED, 2, "unknown problem kind");
    ARG_CHECK(ctx, nslots >= 1 && nslots <= FEAST_MAX_SLOTS, 3, "nslots out of range");
    if (kind == FEAST_PROBLEM_STANDARD) ARG_CHECK(ctx, nslots == 1, 3, "standard problem has one operator");
    if (kind == FEAST_PROBLEM_GENERALIZED) ARG_CHECK(ctx, nslots == 2, 3, "generalized problem has two operators");
    if (kind == FEAST_PROBLEM_POLYNOMIAL) ARG_CHECK(ctx, nslots >= 2, 3, "polynomial problem needs degree >= 1");
    if (kind == FEAST_PROBLEM_SAMPLED) ARG_CHECK(ctx, nslots == 1, 3, "a sampled problem has one operator slot (the current sample)");
    FEAST_TRY(bind_device(ctx));
    free_problem_derived(ctx);
    if (ctx->ops[0].kind == OP_NONE) return feast_fail(ctx, FEAST_ERR_STATE, "slot 0 (A) is not set");
    const int64_t n = ctx->ops[0].n;
    if (kind == FEAST_PROBLEM_STANDARD) {  // B = I implicitly (src/feast.jl:64 `A - I*z`)
        free_operator(ctx->ops[1]);
        ctx->ops[1].kind = OP_IDENTITY; ctx->ops[1].n = n; ctx->ops[1].symmetric = true;
        nslots = 2;
    }
    bool any_dense = false, any_sparse = false;
    for (int s = 0; s < nslots; ++s) {
        const Operator& op = ctx->ops[s];
        if (op.kind == OP_NONE) return feast_fail(ctx, FEAST_ERR_STATE, "slot %d is not set", s);
        if (op.n != n) return feast_fail(ctx, FEAST_ERR_STATE, "slot %d has dimension %lld, expected %lld", s, (long long)op.n, (long long)n);
        any_dense |= op.kind == OP_DENSE;
        any_sparse |= op.kind == OP_CSR;
    }
    if (any_dense && any_sparse)
        return feast_fail(ctx, FEAST_ERR_STATE, "mixing dense and sparse operators is not supported: densify on the caller side");
    if (!any_dense && !any_sparse) return feast_fail(ctx, FEAST_ERR_STATE, "all operators are identities");
    ctx->problem = kind; ctx->nslots = nslots; ctx->n = n; ctx->storage_dense = any_dense;
    if (ctx->m0 != 0) free_blocks(ctx);  // a new problem invalidates the subspace blocks
    if (!any_dense) FEAST_TRY(build_union(ctx));
    ctx->problem_ready = true;
    return 0;
}

// Replace the sample held in slot 0 of a sampled problem (same dimension and storage class).  Subspace blocks, contour
// and stored factorisations are kept; for sparse samples the union pattern is rebuilt (natural row order).
int feast_set_sample_dense(feast_ctx* ctx, int64_t n, const void* a, int64_t lda, int is_complex) {
    FEAST_TRY(check_ready(ctx, false));
    if (ctx->problem != FEAST_PROBLEM_SAMPLED || !ctx->storage_dense)
        return feast_fail(ctx, FEAST_ERR_STATE, "feast_set_sample_dense needs a sampled problem with dense storage");
    ARG_CHECK(ctx, n == ctx->n, 2, "dimension differs from the problem");
    ARG_CHECK(ctx, a != nullptr, 3, "null matrix");
    ARG_CHECK(ctx, lda >= n, 4, "lda < n");
    Operator& op = ctx->ops[0];
    c128* cm = nullptr;
    FEAST_TRY(dev_alloc(ctx, &cm, (size_t)n * n));
    int rc = 0;
    cudaError_t e;
    if (is_complex) {
        e = cudaMemcpy2DAsync(cm, sizeof(c128) * n, a, sizeof(c128) * lda, sizeof(c128) * n, n, cudaMemcpyHostToDevice, ctx->stream);
    } else {
        double* tmp = (double*)op.dense;
        e = cudaMemcpy2DAsync(tmp, sizeof(double) * n, a, sizeof(double) * lda, sizeof(double) * n, n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) rc = launch_real_to_complex(ctx, n * n, tmp, cm);
    }
    if (e == cudaSuccess && !rc) rc = launch_colmajor_to_rowmajor(ctx, n, (int)n, cm, n, op.dense);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(cm);
    if (e != cudaSuccess) return feast_fail(ctx, FEAST_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
    return rc;
}

int feast_set_sample_csc(feast_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval, const void* nzval,
                         int is_complex, int index_base) {
    FEAST_TRY(check_ready(ctx, false));
    if (ctx->problem != FEAST_PROBLEM_SAMPLED || ctx->storage_dense)
        return feast_fail(ctx, FEAST_ERR_STATE, "feast_set_sample_csc needs a sampled problem with sparse storage");
    ARG_CHECK(ctx, n == ctx->n, 2, "dimension differs from the problem");
    ARG_CHECK(ctx, colptr != nullptr, 3, "null colptr");
    ARG_CHECK(ctx, rowval != nullptr || colptr[n] == index_base, 4, "null rowval");
    ARG_CHECK(ctx, nzval != nullptr || colptr[n] == index_base, 5, "null nzval");
    ARG_CHECK(ctx, index_base == 0 || index_base == 1, 7, "index_base must be 0 or 1");
    Operator& op = ctx->ops[0];
    FEAST_TRY(csc_to_host_csr(ctx, n, colptr, rowval, nzval, is_complex, index_base, op.host));
    op.is_complex = is_complex != 0;
    free_pattern(ctx);
    return build_union(ctx);
}

// ------------------------------------------------------------------------- contour / solver / comm
int feast_set_contour(feast_ctx* ctx, int nnodes, const feast_c128* z, const feast_c128* w) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, nnodes >= 1 && nnodes <= feast_ctx::kMaxNodes, 2, "need between 1 and 4096 nodes");
    ARG_CHECK(ctx, z != nullptr, 3, "null nodes");
    ARG_CHECK(ctx, w != nullptr, 4, "null weights");
    ctx->znodes.resize(nnodes);
    ctx->zweights.resize(nnodes);
    for (int k = 0; k < nnodes; ++k) { ctx->znodes[k] = hc128(z[k].re, z[k].im); ctx->zweights[k] = hc128(w[k].re, w[k].im); }
    // default owners: round-robin pairs node k with node k + nnodes/2 on the same rank when possible
    ctx->owner.assign(nnodes, 0);
    for (int k = 0; k < nnodes; ++k) ctx->owner[k] = k % ctx->nranks;
    ctx->node_cost.assign(nnodes, 0.0);
    ctx->have_costs = false;
    drop_stored_factors(ctx);   // the factors belong to the old nodes
    amg_drop_cache(ctx);
    return 0;
}

int feast_set_solver(feast_ctx* ctx, int kind, int krylov, double inner_tol, int max_inner, int store) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, kind >= 0 && kind <= 3, 2, "unknown solver kind");
    ARG_CHECK(ctx, krylov >= 0 && krylov <= 3, 3, "unknown Krylov method");
    ARG_CHECK(ctx, inner_tol > 0 && inner_tol < 1, 4, "inner_tol must be in (0,1)");
    ARG_CHECK(ctx, max_inner >= 1, 5, "max_inner must be positive");
    if ((store != 0) != (ctx->store != 0) || kind != ctx->solver) drop_stored_factors(ctx);
    ctx->solver = kind; ctx->krylov = krylov; ctx->inner_tol = inner_tol; ctx->max_inner = max_inner; ctx->store = store;
    // the internal row ordering follows the solver kind (Krylov: tiled; direct: natural).  A change after
    // feast_set_problem rebuilds the device operators from the host copies and drops the subspace blocks.
    if (ctx->problem_ready && !ctx->storage_dense && (want_reorder(ctx) != ctx->reordered || want_amg(ctx) != (ctx->amg != nullptr)))
        FEAST_TRY(rebuild_layout(ctx));
    return 0;
}

int feast_set_preconditioner(feast_ctx* ctx, int kind) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, kind >= FEAST_PRECOND_NONE && kind <= FEAST_PRECOND_AUTO, 2, "unknown preconditioner kind");
    ctx->precond = kind;
    if (ctx->problem_ready && !ctx->storage_dense && want_amg(ctx) != (ctx->amg != nullptr)) FEAST_TRY(rebuild_layout(ctx));
    return 0;
}

// beta of the complex-shifted preconditioner (0: the hierarchy is assembled at the node itself)
int feast_set_preconditioner_shift(feast_ctx* ctx, double beta) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, beta >= 0.0 && beta <= 4.0, 2, "beta must be in [0, 4]");
    if (beta != ctx->precond_shift) amg_drop_cache(ctx);   // the cached coarse inverses belong to the old shift
    ctx->precond_shift = beta;
    return 0;
}

int feast_preconditioner_info(const feast_ctx* ctx, int* nlevels, int* sizes, int cap, double* setup_seconds) {
    if (!ctx) return feast_fail(nullptr, -1, "argument 1 invalid: null context");
    return amg_info(ctx, nlevels, sizes, cap, setup_seconds);
}

int feast_set_mixed_precision(feast_ctx* ctx, int on) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ctx->mixed_prec = on ? 1 : 0;
    return 0;
}

// host-only: the group choice of the column-sharded contour loop for given node costs (CPU tests of the N > 1 logic)
int feast_debug_pick_groups(int nnodes, const double* cost, int nranks, int m0, int* group_of_node) {
    if (nnodes < 1 || !cost || nranks < 1 || m0 < 1 || !group_of_node) return -1;
    std::vector<int> own;
    const int G = pick_groups(nnodes, cost, nranks, m0, 12.0, 0, own);
    for (int k = 0; k < nnodes; ++k) group_of_node[k] = own[k];
    return G;
}

int feast_set_sharding(feast_ctx* ctx, int mode) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, mode >= FEAST_SHARD_AUTO && mode <= FEAST_SHARD_COLUMNS, 2, "unknown sharding mode");
    ctx->shard_mode = mode;
    return 0;
}

// Moment accumulators of the following contour passes: S_p = sum_k w_k z_k^p (...) for p < nmom (0 restores the
// default: 1 for linear problems, 2 for nonlinear ones).  S_0 and S_1 are the blocks Q and Q1 of the other entries.
int feast_set_moments(feast_ctx* ctx, int nmom) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, nmom >= 0 && nmom <= FEAST_MAX_MOMENTS, 2, "between 0 (default) and 8 moments");
    ctx->nmom = nmom;
    return 0;
}

static const c128* named_block(feast_ctx* ctx, int id) {   // >= 0: moment S_id ; -1: X ; -2: R
    if (id == -1) return ctx->X.p;
    if (id == -2) return ctx->R.p;
    if (id == 0) return ctx->Q.p;
    if (id == 1) return ctx->Q1.p;
    if (id >= 2 && id - 2 < (int)ctx->mom.size()) return ctx->mom[id - 2].p;
    return nullptr;
}

// G (m0 x m0, column-major) = block(a)^H block(b); block ids: p >= 0 moment S_p, -1 the subspace block X, -2 R.
// (Y' * S_p of block_SS!, src/beyn.jl:65-68; the Gram matrices behind the tall SVD of the block-Hankel moment matrix)
int feast_block_gram(feast_ctx* ctx, int a, int b, feast_c128* G) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, G != nullptr, 4, "null output");
    const c128 *A = named_block(ctx, a), *B = named_block(ctx, b);
    if (!A || !B) return feast_fail(ctx, FEAST_ERR_STATE, "block %d or %d does not exist (yet)", a, b);
    const int m = ctx->m0;
    FEAST_TRY(launch_gram(ctx, ctx->n, m, A, B, ctx->small_d));
    CUDA_TRY(ctx, cudaMemcpyAsync(G, ctx->small_d, sizeof(c128) * m * m, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// X = sum_{p < nblk} S_p W_p, W = [W_0; W_1; ...] (nblk*m0 x m0, column-major, ldw >= nblk*m0)
// (X = S[:, 1:K] * V * Xq of block_SS!, src/beyn.jl:87; Y = U[1:N, :] * vectors of nlfeast_moments!, nlfeast.jl:231)
int feast_moment_combine(feast_ctx* ctx, int nblk, const feast_c128* W, int64_t ldw) {
    FEAST_TRY(check_ready(ctx, true));
    const int m = ctx->m0;
    ARG_CHECK(ctx, nblk >= 1 && nblk <= FEAST_MAX_MOMENTS, 2, "nblk out of range");
    ARG_CHECK(ctx, W != nullptr, 3, "null W");
    ARG_CHECK(ctx, ldw >= (int64_t)nblk * m, 4, "ldw < nblk * m0");
    FEAST_TRY(ensure_block(ctx, ctx->W1));
    FEAST_TRY(ensure_block(ctx, ctx->W2));
    std::vector<hc128> Wp((size_t)m * m);
    const int64_t total = ctx->n * m;
    for (int p = 0; p < nblk; ++p) {
        const c128* S = named_block(ctx, p);
        if (!S) return feast_fail(ctx, FEAST_ERR_STATE, "moment %d has not been accumulated", p);
        for (int j = 0; j < m; ++j)
            for (int i = 0; i < m; ++i) Wp[(size_t)j * m + i] = hc128(W[(size_t)j * ldw + (size_t)p * m + i].re, W[(size_t)j * ldw + (size_t)p * m + i].im);
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->small_d, Wp.data(), sizeof(c128) * m * m, cudaMemcpyHostToDevice, ctx->stream));
        FEAST_TRY(launch_update(ctx, ctx->n, m, S, ctx->small_d, p == 0 ? ctx->W2.p : ctx->W1.p));
        if (p > 0) FEAST_TRY(launch_axpy(ctx, total, ctx->W1.p, ctx->W2.p));     // W2 += W1
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // Wp is reused
    }
    std::swap(ctx->X.p, ctx->W2.p);
    return 0;
}

// ||T(l_j)||_F of the last feast_recover_residual of a polynomial problem (absolute residual = res * fro)
int feast_last_fro(feast_ctx* ctx, double* fro) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, fro != nullptr, 2, "null output");
    for (int j = 0; j < ctx->m0; ++j) fro[j] = j < (int)ctx->last_fro.size() ? ctx->last_fro[j] : 0.0;
    return 0;
}

int feast_layout_info(const feast_ctx* ctx, int* info4, double* halo) {
    if (!ctx) return feast_fail(nullptr, -1, "argument 1 invalid: null context");
    if (info4) { info4[0] = ctx->reordered ? 1 : 0; info4[1] = ctx->ntiles; info4[2] = ctx->bandwidth; info4[3] = ctx->tiles_ok ? 1 : 0; }
    if (halo) *halo = ctx->halo_ratio;
    return 0;
}

int feast_comm_unique_id(void* id128) {
    if (!id128) return feast_fail(nullptr, -1, "argument 1 invalid: null id buffer");
    const NcclApi* api = nccl_api();
    if (!api) return feast_fail(nullptr, FEAST_ERR_NCCL, "libnccl.so.2 could not be loaded");
    NcclUid id;
    int rc = api->GetUniqueId(&id);
    if (rc) return feast_fail(nullptr, FEAST_ERR_NCCL, "ncclGetUniqueId: %s", api->GetErrorString ? api->GetErrorString(rc) : "error");
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int feast_comm_init(feast_ctx* ctx, int nranks, int rank, const void* id128) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, nranks >= 1, 2, "nranks must be positive");
    ARG_CHECK(ctx, rank >= 0 && rank < nranks, 3, "rank out of range");
    FEAST_TRY(bind_device(ctx));
    ctx->nranks = nranks; ctx->rank = rank;
    if (nranks > 1) {
        ARG_CHECK(ctx, id128 != nullptr, 4, "null unique id");
        const NcclApi* api = nccl_api();
        if (!api) return feast_fail(ctx, FEAST_ERR_NCCL, "libnccl.so.2 could not be loaded");
        NcclUid id;
        memcpy(&id, id128, sizeof(id));
        int rc = api->CommInitRank(&ctx->nccl_comm, nranks, id, rank);
        if (rc) return feast_fail(ctx, FEAST_ERR_NCCL, "ncclCommInitRank: %s", api->GetErrorString ? api->GetErrorString(rc) : "error");
        // NCCL sets up its channels / peer connections lazily at the first collective (seconds at 8 ranks): do that here,
        // on a private stream and with a message large enough to take the bandwidth algorithm of the real all-reduce, so
        // that it overlaps the host-side layout build (the binding runs this entry on a helper thread) instead of
        // landing inside the first contour pass.
        static const bool warm = !(getenv("FEAST_NCCL_WARMUP") && atoi(getenv("FEAST_NCCL_WARMUP")) == 0);
        if (warm) {
            cudaStream_t ws = nullptr;
            double* buf = nullptr;
            const size_t count = (size_t)8 << 20;   // 64 MB
            if (cudaStreamCreateWithFlags(&ws, cudaStreamNonBlocking) == cudaSuccess && cudaMalloc((void**)&buf, count * sizeof(double)) == cudaSuccess) {
                cudaMemsetAsync(buf, 0, count * sizeof(double), ws);
                rc = api->AllReduce(buf, buf, count, kNcclDouble, kNcclSum, ctx->nccl_comm, ws);
                cudaStreamSynchronize(ws);
            }
            if (buf) cudaFree(buf);
            if (ws) cudaStreamDestroy(ws);
            cudaGetLastError();
            if (rc) return feast_fail(ctx, FEAST_ERR_NCCL, "ncclAllReduce (warm-up): %s", api->GetErrorString ? api->GetErrorString(rc) : "error");
        }
    }
    for (size_t k = 0; k < ctx->owner.size(); ++k) ctx->owner[k] = (int)(k % nranks);
    return 0;
}

int feast_set_node_owners(feast_ctx* ctx, int nnodes, const int* owner) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, nnodes == (int)ctx->znodes.size(), 2, "nnodes differs from the contour");
    for (int k = 0; k < nnodes; ++k) {
        const int o = owner ? owner[k] : k % ctx->nranks;
        ARG_CHECK(ctx, o >= 0 && o < ctx->nranks, 3, "owner rank out of range");
        ctx->owner[k] = o;
    }
    return 0;
}

// ------------------------------------------------------------------------- subspace
int feast_set_subspace(feast_ctx* ctx, int64_t n, int m0, const feast_c128* X, int64_t ldx) {
    FEAST_TRY(check_ready(ctx, false));
    ARG_CHECK(ctx, n == ctx->n, 2, "Incorrect dimensions of X, must match A");  // src/feast.jl:15-16
    ARG_CHECK(ctx, m0 >= 1 && m0 <= n, 3, "m0 out of range");
    ARG_CHECK(ctx, X != nullptr, 4, "null X");
    ARG_CHECK(ctx, ldx >= n, 5, "ldx < n");
    if (ctx->m0 != m0) {
        free_blocks(ctx);
        ctx->m0 = m0;
        FEAST_TRY(dev_alloc(ctx, &ctx->stage, (size_t)n * m0));
        FEAST_TRY(dev_alloc(ctx, &ctx->small_d, feast_ctx::small_elems(m0)));
        // reduction scratch: split-K Gram partials (2*148 slices of m0 x m0) or SpMM/col-dot partials
        size_t red = std::max((size_t)2 * kNumSMs * m0 * m0 * sizeof(c128), spmm_partials_bytes(m0));
        red = std::max(red, (size_t)kNumSMs * 4 * 3 * (size_t)std::max(m0, 256) * sizeof(double));
        red = std::max(red, (size_t)kNumSMs * 4 * 8 * 2 * (size_t)m0 * sizeof(double));   // GMRES multi-dot partials
        red = std::max(red, (size_t)1 << 20);
        FEAST_TRY(dev_alloc(ctx, &ctx->red_d, red / sizeof(double)));
        ctx->red_bytes = red;
        FEAST_TRY(ensure_pinned(ctx, sizeof(hc128) * ((size_t)4 * m0 * m0 + 16 * m0) + 4096));
    }
    FEAST_TRY(ensure_block(ctx, ctx->Q));
    FEAST_TRY(ensure_block(ctx, ctx->X));
    FEAST_TRY(ensure_block(ctx, ctx->R));
    CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->stage, sizeof(c128) * n, X, sizeof(c128) * ldx, sizeof(c128) * n, m0,
                                    cudaMemcpyHostToDevice, ctx->stream));
    FEAST_TRY(launch_colmajor_to_rowmajor(ctx, n, m0, ctx->stage, n, ctx->Q.p, ctx->perm_d));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->X.p, ctx->Q.p, sizeof(c128) * n * m0, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Overwrite the block X only (Q and the moment accumulators keep their contents): the probe block of block_SS!
// (src/beyn.jl:45) is uploaded this way after the contour pass.
int feast_set_X(feast_ctx* ctx, const feast_c128* X, int64_t ldx) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, X != nullptr, 2, "null X");
    ARG_CHECK(ctx, ldx >= ctx->n, 3, "ldx < n");
    const int64_t n = ctx->n;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->stage, sizeof(c128) * n, X, sizeof(c128) * ldx, sizeof(c128) * n, ctx->m0,
                                    cudaMemcpyHostToDevice, ctx->stream));
    FEAST_TRY(launch_colmajor_to_rowmajor(ctx, n, ctx->m0, ctx->stage, n, ctx->X.p, ctx->perm_d));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int feast_get_X(feast_ctx* ctx, feast_c128* X, int64_t ldx) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, X != nullptr, 2, "null X");
    ARG_CHECK(ctx, ldx >= ctx->n, 3, "ldx < n");
    return download_block(ctx, ctx->X, X, ldx);
}
int feast_get_Q(feast_ctx* ctx, feast_c128* Q, int64_t ldq) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, Q != nullptr, 2, "null Q");
    ARG_CHECK(ctx, ldq >= ctx->n, 3, "ldq < n");
    return download_block(ctx, ctx->Q, Q, ldq);
}
int feast_get_R(feast_ctx* ctx, feast_c128* R, int64_t ldr) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, R != nullptr, 2, "null R");
    ARG_CHECK(ctx, ldr >= ctx->n, 3, "ldr < n");
    return download_block(ctx, ctx->R, R, ldr);
}

// ------------------------------------------------------------------------- phases
int feast_project(feast_ctx* ctx, feast_c128* Aq, feast_c128* Bq) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, Aq != nullptr, 2, "null Aq");
    if (ctx->problem == FEAST_PROBLEM_POLYNOMIAL || ctx->problem == FEAST_PROBLEM_SAMPLED)
        return feast_fail(ctx, FEAST_ERR_STATE, "feast_project applies to linear problems; use feast_beyn_reduce");
    if (ctx->problem == FEAST_PROBLEM_GENERALIZED) ARG_CHECK(ctx, Bq != nullptr, 3, "null Bq");
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    PhaseTimer tm(ctx, 0);
    debug_check_finite(ctx, ctx->Q.p, 2 * n * m, "project: Q in");
    FEAST_TRY(orthonormalize(ctx, ctx->Q, nullptr));                       // feast.jl:41 / :117
    debug_check_finite(ctx, ctx->Q.p, 2 * n * m, "project: Q orth");
    FEAST_TRY(ensure_block(ctx, ctx->W1));
    c128* G_d = ctx->small_d;
    FEAST_TRY(apply_slot(ctx, 0, ctx->Q.p, ctx->R.p));                     // R = A Q      feast.jl:42
    debug_check_finite(ctx, ctx->R.p, 2 * n * m, "project: R = A Q");
    FEAST_TRY(launch_gram(ctx, n, m, ctx->Q.p, ctx->R.p, G_d));            // Aq = Q' R    feast.jl:43
    debug_check_finite(ctx, G_d, 2 * m * m, "project: Aq");
    CUDA_TRY(ctx, cudaMemcpyAsync(Aq, G_d, sizeof(c128) * m * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (Bq) {
        c128* G2_d = ctx->small_d + (size_t)m * m;
        FEAST_TRY(apply_slot(ctx, 1, ctx->Q.p, ctx->W1.p));                // R = B Q      feast.jl:120
        debug_check_finite(ctx, ctx->W1.p, 2 * n * m, "project: B Q");
        FEAST_TRY(launch_gram(ctx, n, m, ctx->Q.p, ctx->W1.p, G2_d));      // Bq = Q' R    feast.jl:121
        debug_check_finite(ctx, G2_d, 2 * m * m, "project: Bq");
        CUDA_TRY(ctx, cudaMemcpyAsync(Bq, G2_d, sizeof(c128) * m * m, cudaMemcpyDeviceToHost, ctx->stream));
    }
    tm.stop();
    return 0;
}

int feast_recover_residual(feast_ctx* ctx, const feast_c128* Xq, const feast_c128* lambda, double* res) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, lambda != nullptr, 3, "null lambda");
    ARG_CHECK(ctx, res != nullptr, 4, "null res");
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    PhaseTimer tm(ctx, 1);
    c128* M_d = ctx->small_d;
    c128* lam_d = ctx->vec_lam();
    double* nrm_d = ctx->vec_nrm();
    double* fro_d = nrm_d + m;
    CUDA_TRY(ctx, cudaMemcpyAsync(lam_d, lambda, sizeof(c128) * m, cudaMemcpyHostToDevice, ctx->stream));
    if (Xq) {   // Xq == NULL: X is taken as it is (feast_moment_combine has produced it)
        CUDA_TRY(ctx, cudaMemcpyAsync(M_d, Xq, sizeof(c128) * m * m, cudaMemcpyHostToDevice, ctx->stream));
        FEAST_TRY(launch_update(ctx, n, m, ctx->Q.p, M_d, ctx->X.p));      // X = Q Xq            feast.jl:48
    }
    debug_check_finite(ctx, ctx->X.p, 2 * n * m, "recover: X = Q Xq");
    FEAST_TRY(launch_colnorm2(ctx, n, m, ctx->X.p, nrm_d));
    FEAST_TRY(launch_colnormalize(ctx, n, m, ctx->X.p, nrm_d));            // x_j /= ||x_j||      utils.jl:113
    debug_check_finite(ctx, ctx->X.p, 2 * n * m, "recover: X normalised");
    double* hres = (double*)ctx->pinned;
    if (ctx->problem == FEAST_PROBLEM_SAMPLED) {
        // the caller evaluates T(l_j) and finishes column by column with feast_sampled_residual
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        for (int j = 0; j < m; ++j) res[j] = 0.0;
    } else if (ctx->problem != FEAST_PROBLEM_POLYNOMIAL) {
        FEAST_TRY(apply_slot(ctx, 0, ctx->X.p, ctx->R.p));                 // R = A X
        const c128* BX = ctx->X.p;
        // after build_union an identity B is an OP_CSR slot with unit diagonal: skip the SpMM
        const bool b_identity = (ctx->problem == FEAST_PROBLEM_STANDARD);
        if (!b_identity) {
            FEAST_TRY(ensure_block(ctx, ctx->W1));
            FEAST_TRY(apply_slot(ctx, 1, ctx->X.p, ctx->W1.p));
            BX = ctx->W1.p;
        }
        FEAST_TRY(launch_residual_combine(ctx, n, m, ctx->R.p, BX, lam_d)); // R_j = (A - l_j B) x_j  utils.jl:114
        FEAST_TRY(launch_colnorm2(ctx, n, m, ctx->R.p, nrm_d));            // res_j = ||R_j||      utils.jl:168
        CUDA_TRY(ctx, cudaMemcpyAsync(hres, nrm_d, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        for (int j = 0; j < m; ++j) res[j] = std::sqrt(hres[j]);
    } else {
        if (ctx->storage_dense) {
            // R = sum_i (A_i X) diag(l^i)
            FEAST_TRY(ensure_block(ctx, ctx->W1));
            std::vector<hc128> pw(m, hc128(1, 0));
            c128* pw_d = ctx->small_d + (size_t)m * m;
            CUDA_TRY(ctx, cudaMemsetAsync(ctx->R.p, 0, sizeof(c128) * n * m, ctx->stream));
            for (int s = 0; s < ctx->nslots; ++s) {
                FEAST_TRY(apply_slot(ctx, s, ctx->X.p, ctx->W1.p));
                CUDA_TRY(ctx, cudaMemcpyAsync(pw_d, pw.data(), sizeof(c128) * m, cudaMemcpyHostToDevice, ctx->stream));
                // R += W1 * diag(pw): accumulate with X := W1, Y := 0-block trick -> use first_pass form
                FEAST_TRY(launch_accumulate(ctx, n, m, nullptr, ctx->W1.p, pw_d, ctx->R.p, nullptr, hc128(0, 0), true));
                CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
                for (int j = 0; j < m; ++j) pw[j] *= hc128(lambda[j].re, lambda[j].im);
            }
            const c128* D[FEAST_MAX_SLOTS];
            int kinds[FEAST_MAX_SLOTS];
            for (int s = 0; s < ctx->nslots; ++s) { D[s] = ctx->ops[s].dense; kinds[s] = ctx->ops[s].kind; }
            FEAST_TRY(launch_poly_fro_dense(ctx, n, m, ctx->nslots, D, kinds, lam_d, fro_d));
        } else {
            const double* rv[FEAST_MAX_SLOTS];
            const c128* cv[FEAST_MAX_SLOTS];
            for (int s = 0; s < ctx->nslots; ++s) { rv[s] = ctx->ops[s].uvals_r; cv[s] = ctx->ops[s].uvals_c; }
            FEAST_TRY(launch_poly_residual(ctx, n, m, ctx->nslots, ctx->u_rowptr, ctx->u_col, ctx->unnz, rv, cv, lam_d,
                                           ctx->X.p, ctx->R.p, fro_d));   // R_j = T(l_j) x_j   utils.jl:107
        }
        FEAST_TRY(launch_colnorm2(ctx, n, m, ctx->R.p, nrm_d));
        CUDA_TRY(ctx, cudaMemcpyAsync(hres, nrm_d, sizeof(double) * 2 * m, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->last_fro.resize(m);
        for (int j = 0; j < m; ++j) {
            ctx->last_fro[j] = std::sqrt(hres[m + j]);
            res[j] = std::sqrt(hres[j]) / ctx->last_fro[j];  // utils.jl:154
        }
    }
    tm.stop();
    return 0;
}

// ---- the contour loop (src/feast.jl:57-71, src/nlfeast.jl:36-61) in three pieces: begin (zero the accumulators,
// re-shard the nodes), one node (shifted solve + weighted accumulation), end (all-reduce).  feast_contour_apply runs
// all of it in one call; the sampled-operator entries (opaque T(z) closures evaluated by the caller) drive the pieces.
static int contour_prepare(feast_ctx* ctx, const feast_c128* lambda, int first_pass, int* solver, int* method) {
    ARG_CHECK(ctx, lambda != nullptr || first_pass, 2, "null lambda");
    if (ctx->znodes.empty()) return feast_fail(ctx, FEAST_ERR_STATE, "feast_set_contour has not been called");
    const bool nep = ctx->problem == FEAST_PROBLEM_POLYNOMIAL || ctx->problem == FEAST_PROBLEM_SAMPLED;
    if (first_pass && !nep)
        return feast_fail(ctx, -3, "argument 3 invalid: first_pass applies to nonlinear (polynomial / sampled) problems only");
    *solver = effective_solver(ctx);
    *method = effective_krylov(ctx);
    if (*solver == FEAST_SOLVER_KRYLOV && ctx->storage_dense)
        return feast_fail(ctx, FEAST_ERR_STATE, "Krylov inner solves need sparse operators");
    FEAST_TRY(ensure_block(ctx, ctx->W1));
    if (nep) FEAST_TRY(ensure_block(ctx, ctx->Q1));
    if (*solver == FEAST_SOLVER_KRYLOV) FEAST_TRY(ensure_krylov_work(ctx, *method));
    const int nnodes = (int)ctx->znodes.size();
    if (ctx->store && (int)ctx->stored.size() != nnodes) ctx->stored.resize(nnodes);
    // Krylov solves have no per-node factorisation to keep together, so with several ranks the COLUMNS of every node's
    // right-hand side are sharded instead of the nodes: every rank runs all nodes on its m0 / nranks columns.  The load
    // is then balanced by construction -- with the multigrid preconditioner the near-axis nodes need ~10x the iterations
    // of the others, which no node -> rank map can balance at 2 nodes per rank -- and the vector work scales with the columns.
    static const char* shard_env = getenv("FEAST_SHARD");
    int mode = ctx->shard_mode;
    if (mode == FEAST_SHARD_AUTO && shard_env) mode = !strcmp(shard_env, "nodes") ? FEAST_SHARD_NODES : (!strcmp(shard_env, "columns") ? FEAST_SHARD_COLUMNS : mode);
    ctx->col_shard = ctx->nranks > 1 && *solver == FEAST_SOLVER_KRYLOV && ctx->problem != FEAST_PROBLEM_SAMPLED &&
                     mode != FEAST_SHARD_NODES && ctx->m0 >= 2 * ctx->nranks;
    const int nmom = ctx->nmom > 0 ? ctx->nmom : (nep ? 2 : 1);
    if (nmom > 1) FEAST_TRY(ensure_block(ctx, ctx->Q1));
    if ((int)ctx->mom.size() < nmom - 2) ctx->mom.resize(nmom - 2);
    for (int p = 2; p < nmom; ++p) FEAST_TRY(ensure_block(ctx, ctx->mom[p - 2]));
    if (ctx->col_shard) FEAST_TRY(ensure_block(ctx, ctx->W2));
    return 0;
}

static int contour_nmom(const feast_ctx* ctx) {
    const bool nep = ctx->problem == FEAST_PROBLEM_POLYNOMIAL || ctx->problem == FEAST_PROBLEM_SAMPLED;
    return ctx->nmom > 0 ? ctx->nmom : (nep ? 2 : 1);
}
static c128* moment_block(feast_ctx* ctx, int p) { return p == 0 ? ctx->Q.p : (p == 1 ? ctx->Q1.p : ctx->mom[p - 2].p); }

static bool contour_can_move_nodes(const feast_ctx* ctx, int solver) {   // stored factors / caller-evaluated samples pin nodes to ranks
    return ctx->nranks > 1 && ctx->auto_balance && !ctx->col_shard && ctx->problem != FEAST_PROBLEM_SAMPLED &&
           !(solver != FEAST_SOLVER_KRYLOV && ctx->store);
}

static int contour_begin(feast_ctx* ctx, int solver) {
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    const bool nep = ctx->problem == FEAST_PROBLEM_POLYNOMIAL || ctx->problem == FEAST_PROBLEM_SAMPLED;
    (void)nep;
    for (int p = 0; p < contour_nmom(ctx); ++p)                                                // feast.jl:58, nlfeast.jl:32-33
        CUDA_TRY(ctx, cudaMemsetAsync(moment_block(ctx, p), 0, sizeof(c128) * n * m, ctx->stream));
    if (contour_can_move_nodes(ctx, solver) && ctx->have_costs) rebalance_nodes(ctx);
    if (ctx->col_shard) choose_groups(ctx);
    ctx->cost_local.assign(ctx->znodes.size(), 0.0);
    return 0;
}

// does this rank take part in node k of the running pass?
static bool contour_my_node(const feast_ctx* ctx, int k) {
    if (!ctx->col_shard) return ctx->owner[k] == ctx->rank;
    const int gs = ctx->nranks / ctx->ngroups;
    return ctx->gowner[k] == ctx->rank / gs;
}

static int contour_node(feast_ctx* ctx, int k, const feast_c128* lambda, int first_pass, int solver, int method, feast_stats& st,
                        int* rc_final) {
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    const bool nep = ctx->problem == FEAST_PROBLEM_POLYNOMIAL || ctx->problem == FEAST_PROBLEM_SAMPLED;
    cudaEvent_t e0 = ctx->evn[0], e1 = ctx->evn[1], e2 = ctx->evn[2];
    c128* d_d = ctx->vec_d();
    std::vector<hc128> d(m);
    hc128 coef[FEAST_MAX_SLOTS];
    (void)nep;
    const c128* rhs = first_pass ? ctx->X.p : ctx->R.p;
    st.nodes_local++;
    st.col_sharded = ctx->col_shard ? ctx->nranks / ctx->ngroups : 0;
    const hc128 z = ctx->znodes[k], w = ctx->zweights[k];
    node_coefs(ctx, z, coef);
    for (int j = 0; j < m; ++j)
        d[j] = first_pass ? w : w / (z - hc128(lambda[j].re, lambda[j].im));              // feast.jl:60,69
    CUDA_TRY(ctx, cudaMemcpyAsync(d_d, d.data(), sizeof(c128) * m, cudaMemcpyHostToDevice, ctx->stream));
    cudaEventRecord(e0, ctx->stream);
    int j0 = 0, mloc = m;
    if (ctx->col_shard && ctx->ngroups < ctx->nranks) {   // this rank's column slice of the right-hand side, as a compact block
        const int gs = ctx->nranks / ctx->ngroups, sl = ctx->rank % gs;
        j0 = (int)((int64_t)sl * m / gs);
        mloc = (int)((int64_t)(sl + 1) * m / gs) - j0;
        FEAST_TRY(launch_gather_cols(ctx, n, m, j0, mloc, rhs, ctx->W2.p));
        rhs = ctx->W2.p;
    }
    ctx->m0 = mloc;          // the solvers size everything by ctx->m0: a compact n x mloc system
    const int src = solve_shifted(ctx, solver, method, k, coef, rhs, ctx->W1.p, st, e1, rc_final);
    ctx->m0 = m;
    FEAST_TRY(src);
    debug_check_finite(ctx, rhs, 2 * n * mloc, "contour: rhs");
    debug_check_finite(ctx, ctx->W1.p, 2 * n * mloc, "contour: solve result");
    // Q += (X - Y) diag(w/(z - l))  [feast.jl:68-70] ; nonlinear: Q0, Q1 [nlfeast.jl:56-58] ; moments S_p [beyn.jl:19-20,52-54]
    c128* mp[FEAST_MAX_MOMENTS];
    const int nmom = contour_nmom(ctx);
    for (int p = 0; p < nmom; ++p) mp[p] = moment_block(ctx, p);
    FEAST_TRY(launch_accumulate_slice(ctx, n, m, j0, mloc, ctx->X.p, ctx->W1.p, d_d, mp, nmom, z, first_pass != 0));
    cudaEventRecord(e2, ctx->stream);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // d (host vector) goes out of scope
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, e0, e1);
    cudaEventElapsedTime(&b, e1, e2);
    st.t_factor_ms += a;
    st.t_solve_ms += b;
    ctx->cost_local[k] = (double)a + (double)b;
    return 0;
}

static int contour_end(feast_ctx* ctx, int solver, feast_stats& st) {
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    const bool nep = ctx->problem == FEAST_PROBLEM_POLYNOMIAL || ctx->problem == FEAST_PROBLEM_SAMPLED;
    const int nnodes = (int)ctx->znodes.size();
    if (ctx->nranks > 1) {                                                                     // NC1
        const NcclApi* api = nccl_api();
        cudaEvent_t e0 = ctx->evn[0], e3 = ctx->evn[3];
        const bool can_move_nodes = contour_can_move_nodes(ctx, solver) || (ctx->col_shard && ctx->auto_balance);
        cudaEventRecord(e0, ctx->stream);
        (void)nep;
        int rc = 0;
        for (int p = 0; p < contour_nmom(ctx) && !rc; ++p)
            rc = api->AllReduce(moment_block(ctx, p), moment_block(ctx, p), (size_t)2 * n * m, kNcclDouble, kNcclSum, ctx->nccl_comm, ctx->stream);
        if (rc) return feast_fail(ctx, FEAST_ERR_NCCL, "ncclAllReduce: %s", api->GetErrorString ? api->GetErrorString(rc) : "error");
        cudaEventRecord(e3, ctx->stream);
        if (can_move_nodes) {   // share the measured per-node costs (nnodes doubles) for the next pass
            double* cbuf = ctx->vec_cost();
            CUDA_TRY(ctx, cudaMemcpyAsync(cbuf, ctx->cost_local.data(), sizeof(double) * nnodes, cudaMemcpyHostToDevice, ctx->stream));
            rc = api->AllReduce(cbuf, cbuf, (size_t)nnodes, kNcclDouble, kNcclSum, ctx->nccl_comm, ctx->stream);
            if (rc) return feast_fail(ctx, FEAST_ERR_NCCL, "ncclAllReduce (node costs) failed");
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->node_cost.data(), cbuf, sizeof(double) * nnodes, cudaMemcpyDeviceToHost, ctx->stream));
            // (column mode: the sum runs over the nranks / ngroups ranks of a node's group, all at the same slice width)
        }
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (can_move_nodes) ctx->have_costs = true;
        float c = 0;
        cudaEventElapsedTime(&c, e0, e3);
        st.t_reduce_ms = c;
    }
    return 0;
}

static int contour_finish(feast_ctx* ctx, const feast_stats& st, int rc_final) {
    if (st.info) return feast_fail(ctx, FEAST_ERR_SINGULAR, "zero pivot at column %d of a shifted factorisation", st.info);
    if (rc_final == FEAST_WARN_INNER_MAXIT)
        feast_fail(ctx, FEAST_WARN_INNER_MAXIT, "inner solve did not reach inner_tol=%.1e (max_inner=%d, relres %.3e)", ctx->inner_tol,
                   ctx->max_inner, st.inner_relres_max);
    return rc_final;
}

int feast_contour_apply(feast_ctx* ctx, const feast_c128* lambda, int first_pass, feast_stats* stats) {
    FEAST_TRY(check_ready(ctx, true));
    if (ctx->problem == FEAST_PROBLEM_SAMPLED)
        return feast_fail(ctx, FEAST_ERR_STATE, "sampled operators are applied node by node: use feast_contour_node");
    int solver = 0, method = 0;
    FEAST_TRY(contour_prepare(ctx, lambda, first_pass, &solver, &method));
    feast_stats st;
    memset(&st, 0, sizeof(st));
    PhaseTimer tm(ctx, 2);
    int rc_final = 0;
    FEAST_TRY(contour_begin(ctx, solver));
    const int nnodes = (int)ctx->znodes.size();
    for (int k = 0; k < nnodes; ++k) {
        if (!contour_my_node(ctx, k)) continue;
        FEAST_TRY(contour_node(ctx, k, lambda, first_pass, solver, method, st, &rc_final));
    }
    FEAST_TRY(contour_end(ctx, solver, st));
    st.t_total_ms = tm.stop();
    if (stats) *stats = st;
    return contour_finish(ctx, st, rc_final);
}

// Sampled operators (the closure form nlfeast!(T::Function, ...), src/nlfeast.jl:2-4): slot 0 holds the caller's
// evaluation of T at ONE point.  For every contour node the caller uploads T(z_k) with feast_set_sample_* and calls
// feast_contour_node (phase bit 1: first node of the pass, bit 2: last node); with store != 0 the factorisation of a
// node is kept, so later passes need no sample for it (feast_node_needs_sample).
int feast_contour_node(feast_ctx* ctx, int k, const feast_c128* lambda, int first_pass, int phase, feast_stats* stats) {
    FEAST_TRY(check_ready(ctx, true));
    if (ctx->problem != FEAST_PROBLEM_SAMPLED)
        return feast_fail(ctx, FEAST_ERR_STATE, "feast_contour_node applies to sampled problems (FEAST_PROBLEM_SAMPLED)");
    int solver = 0, method = 0;
    FEAST_TRY(contour_prepare(ctx, lambda, first_pass, &solver, &method));
    ARG_CHECK(ctx, k >= -1 && k < (int)ctx->znodes.size(), 2, "node index out of range");
    PhaseTimer tm(ctx, 2);
    if (phase & 1) {
        memset(&ctx->pass_stats, 0, sizeof(ctx->pass_stats));
        ctx->pass_rc = 0;
        FEAST_TRY(contour_begin(ctx, solver));
    }
    if (k >= 0) FEAST_TRY(contour_node(ctx, k, lambda, first_pass, solver, method, ctx->pass_stats, &ctx->pass_rc));   // k = -1: no local node
    if (phase & 2) FEAST_TRY(contour_end(ctx, solver, ctx->pass_stats));
    ctx->pass_stats.t_total_ms += tm.stop();
    if (stats) *stats = ctx->pass_stats;
    return (phase & 2) ? contour_finish(ctx, ctx->pass_stats, ctx->pass_rc) : 0;
}

int feast_node_needs_sample(const feast_ctx* ctx, int k) {
    if (!ctx || k < 0 || k >= (int)ctx->znodes.size()) return -1;
    if (!ctx->store) return 1;
    if (k < (int)ctx->stored.size() && ctx->stored[k].lu) return 0;
    if (k < (int)ctx->bstored.size() && ctx->bstored[k].lu) return 0;
    return 1;
}

// Column j of the nonlinear residual with the current sample T(l_j) in slot 0 (src/utils.jl:104-109,151-157):
// R[:, j] = T x_j ; res = ||R_j|| / fro  (fro = ||T(l_j)||_F, computed by the caller who holds the matrix).
int feast_sampled_residual(feast_ctx* ctx, int j, double fro, double* res) {
    FEAST_TRY(check_ready(ctx, true));
    if (ctx->problem != FEAST_PROBLEM_SAMPLED) return feast_fail(ctx, FEAST_ERR_STATE, "sampled problems only");
    ARG_CHECK(ctx, j >= 0 && j < ctx->m0, 2, "column out of range");
    ARG_CHECK(ctx, fro > 0.0, 3, "norm must be positive");
    ARG_CHECK(ctx, res != nullptr, 4, "null output");
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    const Operator& op = ctx->ops[0];
    PhaseTimer tm(ctx, 1);
    if (op.kind == OP_DENSE)
        FEAST_TRY(launch_zgemm(ctx, (int)n, 1, n, hc128(1, 0), op.dense, n, 1, false, ctx->X.p + j, m, 1, hc128(0, 0), ctx->R.p + j, m, 1));
    else
        FEAST_TRY(launch_spmm(ctx, n, 1, ctx->u_rowptr, ctx->u_col, op.uvals_r, op.uvals_c, ctx->X.p + j, m, ctx->R.p + j, m, nullptr));
    double* nrm_d = ctx->vec_nrm();
    FEAST_TRY(launch_colnorm2(ctx, n, m, ctx->R.p, nrm_d));
    double* h = (double*)ctx->pinned;
    CUDA_TRY(ctx, cudaMemcpyAsync(h, nrm_d + j, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *res = std::sqrt(h[0]) / fro;
    tm.stop();
    return 0;
}

// Stochastic eigenvalue-count estimate (src/stochastic.jl:2-33): the current subspace block X holds the
// probe vectors; est = Re sum_k w_k tr(X' (z_k B - A)^-1 X) / m0, node-sharded like the contour loop.
int feast_estimate_count(feast_ctx* ctx, double* est, feast_stats* stats) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, est != nullptr, 2, "null output");
    if (ctx->problem == FEAST_PROBLEM_POLYNOMIAL || ctx->problem == FEAST_PROBLEM_SAMPLED)
        return feast_fail(ctx, FEAST_ERR_STATE, "feast_estimate_count applies to linear problems");
    if (ctx->znodes.empty()) return feast_fail(ctx, FEAST_ERR_STATE, "feast_set_contour has not been called");
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    const int solver = effective_solver(ctx), method = effective_krylov(ctx);
    if (solver == FEAST_SOLVER_KRYLOV && ctx->storage_dense)
        return feast_fail(ctx, FEAST_ERR_STATE, "Krylov inner solves need sparse operators");
    FEAST_TRY(ensure_block(ctx, ctx->W1));
    if (solver == FEAST_SOLVER_KRYLOV) FEAST_TRY(ensure_krylov_work(ctx, method));
    const int nnodes = (int)ctx->znodes.size();
    if (ctx->store && (int)ctx->stored.size() != nnodes) ctx->stored.resize(nnodes);
    feast_stats st;
    memset(&st, 0, sizeof(st));
    int rc_final = 0;
    c128* dots_d = ctx->small_d;   // m complex
    hc128 acc(0.0, 0.0);
    std::vector<hc128> dots(m);
    for (int k = 0; k < nnodes; ++k) {
        if (ctx->owner[k] != ctx->rank) continue;
        st.nodes_local++;
        const hc128 z = ctx->znodes[k], w = ctx->zweights[k];
        hc128 coef[FEAST_MAX_SLOTS] = {hc128(-1, 0), z};                     // B z - A  (stochastic.jl:24)
        FEAST_TRY(solve_shifted(ctx, solver, method, ctx->store ? -1 : -1, coef, ctx->X.p, ctx->W1.p, st, nullptr, &rc_final));
        FEAST_TRY(launch_coldot(ctx, n, m, ctx->X.p, ctx->W1.p, true, dots_d));   // diag of X' Y  (tr(P), :26-27)
        CUDA_TRY(ctx, cudaMemcpyAsync(dots.data(), dots_d, sizeof(c128) * m, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        hc128 tr(0.0, 0.0);
        for (int j = 0; j < m; ++j) tr += dots[j];
        acc += tr * w / (double)m;
    }
    double part[2] = {acc.real(), acc.imag()};
    if (ctx->nranks > 1) {
        const NcclApi* api = nccl_api();
        double* buf = (double*)ctx->small_d;
        CUDA_TRY(ctx, cudaMemcpyAsync(buf, part, sizeof(part), cudaMemcpyHostToDevice, ctx->stream));
        int rc = api->AllReduce(buf, buf, 2, kNcclDouble, kNcclSum, ctx->nccl_comm, ctx->stream);
        if (rc) return feast_fail(ctx, FEAST_ERR_NCCL, "ncclAllReduce failed");
        CUDA_TRY(ctx, cudaMemcpyAsync(part, buf, sizeof(part), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *est = part[0];                                                          // return real(est), :32
    if (stats) *stats = st;
    if (st.info) return feast_fail(ctx, FEAST_ERR_SINGULAR, "zero pivot at column %d of a shifted factorisation", st.info);
    return rc_final;
}

// ------------------------------------------------------------------------- two-sided driver (dual_gen_feast!)
// src/feast.jl:165-257.  Right blocks live in Q/X/R, left blocks in Ql/Xl/Rl.  One LU per node serves both the
// solve with A - zB and with its adjoint (the reference factors twice, feast.jl:185-194).
int feast_dual_set_subspace(feast_ctx* ctx, int64_t n, int m0, const feast_c128* Xr, int64_t ldr, const feast_c128* Xl,
                            int64_t ldl) {
    FEAST_TRY(feast_set_subspace(ctx, n, m0, Xr, ldr));
    ARG_CHECK(ctx, Xl != nullptr, 6, "null Xl");
    ARG_CHECK(ctx, ldl >= n, 7, "ldl < n");
    FEAST_TRY(ensure_block(ctx, ctx->Ql));
    FEAST_TRY(ensure_block(ctx, ctx->Xl));
    FEAST_TRY(ensure_block(ctx, ctx->Rl));
    CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->stage, sizeof(c128) * n, Xl, sizeof(c128) * ldl, sizeof(c128) * n, m0,
                                    cudaMemcpyHostToDevice, ctx->stream));
    FEAST_TRY(launch_colmajor_to_rowmajor(ctx, n, m0, ctx->stage, n, ctx->Ql.p, ctx->perm_d));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->Xl.p, ctx->Ql.p, sizeof(c128) * n * m0, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// G = Ql' B Qr   (feast.jl:199, the argument of svd!)
int feast_dual_project(feast_ctx* ctx, feast_c128* G) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, G != nullptr, 2, "null G");
    if (!ctx->Ql.p) return feast_fail(ctx, FEAST_ERR_STATE, "feast_dual_set_subspace has not been called");
    if (ctx->problem == FEAST_PROBLEM_POLYNOMIAL || ctx->problem == FEAST_PROBLEM_SAMPLED)
        return feast_fail(ctx, FEAST_ERR_STATE, "linear problems only");
    const int m = ctx->m0;
    PhaseTimer tm(ctx, 0);
    FEAST_TRY(ensure_block(ctx, ctx->W1));
    FEAST_TRY(apply_slot(ctx, 1, ctx->Q.p, ctx->W1.p));
    FEAST_TRY(launch_gram(ctx, ctx->n, m, ctx->Ql.p, ctx->W1.p, ctx->small_d));
    CUDA_TRY(ctx, cudaMemcpyAsync(G, ctx->small_d, sizeof(c128) * m * m, cudaMemcpyDeviceToHost, ctx->stream));
    tm.stop();
    return 0;
}

// Qr <- Qr Mr ; Ql <- Ql Ml (feast.jl:200-201) ; Aq = Ql' A Qr ; Bq = Ql' B Qr (feast.jl:202-205)
int feast_dual_rotate(feast_ctx* ctx, const feast_c128* Mr, const feast_c128* Ml, feast_c128* Aq, feast_c128* Bq) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, Mr != nullptr, 2, "null Mr");
    ARG_CHECK(ctx, Ml != nullptr, 3, "null Ml");
    ARG_CHECK(ctx, Aq != nullptr, 4, "null Aq");
    ARG_CHECK(ctx, Bq != nullptr, 5, "null Bq");
    if (!ctx->Ql.p) return feast_fail(ctx, FEAST_ERR_STATE, "feast_dual_set_subspace has not been called");
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    PhaseTimer tm(ctx, 0);
    FEAST_TRY(ensure_block(ctx, ctx->W1));
    c128* M_d = ctx->small_d;
    c128* G_d = ctx->small_d + (size_t)m * m;
    CUDA_TRY(ctx, cudaMemcpyAsync(M_d, Mr, sizeof(c128) * m * m, cudaMemcpyHostToDevice, ctx->stream));
    FEAST_TRY(launch_update(ctx, n, m, ctx->Q.p, M_d, ctx->W1.p));
    std::swap(ctx->Q.p, ctx->W1.p);
    CUDA_TRY(ctx, cudaMemcpyAsync(M_d, Ml, sizeof(c128) * m * m, cudaMemcpyHostToDevice, ctx->stream));
    FEAST_TRY(launch_update(ctx, n, m, ctx->Ql.p, M_d, ctx->W1.p));
    std::swap(ctx->Ql.p, ctx->W1.p);
    FEAST_TRY(apply_slot(ctx, 0, ctx->Q.p, ctx->R.p));
    FEAST_TRY(launch_gram(ctx, n, m, ctx->Ql.p, ctx->R.p, G_d));
    CUDA_TRY(ctx, cudaMemcpyAsync(Aq, G_d, sizeof(c128) * m * m, cudaMemcpyDeviceToHost, ctx->stream));
    FEAST_TRY(apply_slot(ctx, 1, ctx->Q.p, ctx->W1.p));
    FEAST_TRY(launch_gram(ctx, n, m, ctx->Ql.p, ctx->W1.p, G_d));
    CUDA_TRY(ctx, cudaMemcpyAsync(Bq, G_d, sizeof(c128) * m * m, cudaMemcpyDeviceToHost, ctx->stream));
    tm.stop();
    return 0;
}

// Xr = Qr Xqr, Xl = Ql Xql (feast.jl:209,212); update_R! on both sides (feast.jl:213-214, the left one with
// conj(lambda), see the oracle docstring); resr_j = ||Rr_j|| (feast.jl:215)
int feast_dual_recover_residual(feast_ctx* ctx, const feast_c128* Xqr, const feast_c128* Xql, const feast_c128* lambda,
                                double* resr) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, Xqr != nullptr, 2, "null Xqr");
    ARG_CHECK(ctx, Xql != nullptr, 3, "null Xql");
    ARG_CHECK(ctx, lambda != nullptr, 4, "null lambda");
    ARG_CHECK(ctx, resr != nullptr, 5, "null resr");
    if (!ctx->Ql.p) return feast_fail(ctx, FEAST_ERR_STATE, "feast_dual_set_subspace has not been called");
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    PhaseTimer tm(ctx, 1);
    FEAST_TRY(ensure_block(ctx, ctx->W1));
    c128* M_d = ctx->small_d;
    c128* lam_d = ctx->vec_lam();
    c128* lamc_d = ctx->vec_lamc();
    double* nrm_d = ctx->vec_nrm();
    std::vector<hc128> lamc(m);
    for (int j = 0; j < m; ++j) lamc[j] = hc128(lambda[j].re, -lambda[j].im);
    CUDA_TRY(ctx, cudaMemcpyAsync(lam_d, lambda, sizeof(c128) * m, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(lamc_d, lamc.data(), sizeof(c128) * m, cudaMemcpyHostToDevice, ctx->stream));
    // right side
    CUDA_TRY(ctx, cudaMemcpyAsync(M_d, Xqr, sizeof(c128) * m * m, cudaMemcpyHostToDevice, ctx->stream));
    FEAST_TRY(launch_update(ctx, n, m, ctx->Q.p, M_d, ctx->X.p));
    FEAST_TRY(launch_colnorm2(ctx, n, m, ctx->X.p, nrm_d));
    FEAST_TRY(launch_colnormalize(ctx, n, m, ctx->X.p, nrm_d));
    FEAST_TRY(apply_slot(ctx, 0, ctx->X.p, ctx->R.p));
    FEAST_TRY(apply_slot(ctx, 1, ctx->X.p, ctx->W1.p));
    FEAST_TRY(launch_residual_combine(ctx, n, m, ctx->R.p, ctx->W1.p, lam_d));
    FEAST_TRY(launch_colnorm2(ctx, n, m, ctx->R.p, nrm_d));
    double* hres = (double*)ctx->pinned;
    CUDA_TRY(ctx, cudaMemcpyAsync(hres, nrm_d, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
    // left side
    CUDA_TRY(ctx, cudaMemcpyAsync(M_d, Xql, sizeof(c128) * m * m, cudaMemcpyHostToDevice, ctx->stream));
    FEAST_TRY(launch_update(ctx, n, m, ctx->Ql.p, M_d, ctx->Xl.p));
    FEAST_TRY(launch_colnorm2(ctx, n, m, ctx->Xl.p, nrm_d));
    FEAST_TRY(launch_colnormalize(ctx, n, m, ctx->Xl.p, nrm_d));
    FEAST_TRY(apply_slot_adjoint(ctx, 0, ctx->Xl.p, ctx->Rl.p));
    FEAST_TRY(apply_slot_adjoint(ctx, 1, ctx->Xl.p, ctx->W1.p));
    FEAST_TRY(launch_residual_combine(ctx, n, m, ctx->Rl.p, ctx->W1.p, lamc_d));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < m; ++j) resr[j] = std::sqrt(hres[j]);
    tm.stop();
    return 0;
}

// Qr = sum_k (Xr - (A - z_k B)^-1 Rr) diag(w_k/(z_k - l)) ; Ql = sum_k (Xl - (A - z_k B)^-H Rl) diag(conj(w_k/(z_k - l)))
// (feast.jl:225-247), node-sharded, both accumulators all-reduced.
int feast_dual_contour_apply(feast_ctx* ctx, const feast_c128* lambda, feast_stats* stats) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, lambda != nullptr, 2, "null lambda");
    if (!ctx->Ql.p) return feast_fail(ctx, FEAST_ERR_STATE, "feast_dual_set_subspace has not been called");
    if (ctx->znodes.empty()) return feast_fail(ctx, FEAST_ERR_STATE, "feast_set_contour has not been called");
    const int64_t n = ctx->n;
    const int m = ctx->m0;
    const int solver = effective_solver(ctx), method = effective_krylov(ctx);
    if (solver == FEAST_SOLVER_KRYLOV && ctx->storage_dense)
        return feast_fail(ctx, FEAST_ERR_STATE, "Krylov inner solves need sparse operators");
    FEAST_TRY(ensure_block(ctx, ctx->W1));
    FEAST_TRY(ensure_block(ctx, ctx->W2));
    if (solver == FEAST_SOLVER_KRYLOV) FEAST_TRY(ensure_krylov_work(ctx, method));
    feast_stats st;
    memset(&st, 0, sizeof(st));
    PhaseTimer tm(ctx, 2);
    int rc_final = 0;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->Q.p, 0, sizeof(c128) * n * m, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->Ql.p, 0, sizeof(c128) * n * m, ctx->stream));
    const int nnodes = (int)ctx->znodes.size();
    if (ctx->store && (int)ctx->stored.size() != nnodes) ctx->stored.resize(nnodes);
    c128* d_d = ctx->vec_d();
    c128* dl_d = ctx->vec_dl();
    std::vector<hc128> d(m), dl(m);
    hc128 coef[FEAST_MAX_SLOTS];
    for (int k = 0; k < nnodes; ++k) {
        if (ctx->owner[k] != ctx->rank) continue;
        st.nodes_local++;
        const hc128 z = ctx->znodes[k], w = ctx->zweights[k];
        node_coefs(ctx, z, coef);
        for (int j = 0; j < m; ++j) {
            d[j] = w / (z - hc128(lambda[j].re, lambda[j].im));   // feast.jl:227,233
            dl[j] = std::conj(d[j]);                               // feast.jl:236,245
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(d_d, d.data(), sizeof(c128) * m, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(dl_d, dl.data(), sizeof(c128) * m, cudaMemcpyHostToDevice, ctx->stream));
        FEAST_TRY(solve_shifted(ctx, solver, method, k, coef, ctx->R.p, ctx->W1.p, st, nullptr, &rc_final, false));
        FEAST_TRY(launch_accumulate(ctx, n, m, ctx->X.p, ctx->W1.p, d_d, ctx->Q.p, nullptr, z, false));
        FEAST_TRY(solve_shifted(ctx, solver, method, ctx->store ? k : -2, coef, ctx->Rl.p, ctx->W1.p, st, nullptr, &rc_final, true));
        FEAST_TRY(launch_accumulate(ctx, n, m, ctx->Xl.p, ctx->W1.p, dl_d, ctx->Ql.p, nullptr, z, false));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (ctx->nranks > 1) {
        const NcclApi* api = nccl_api();
        int rc = api->AllReduce(ctx->Q.p, ctx->Q.p, (size_t)2 * n * m, kNcclDouble, kNcclSum, ctx->nccl_comm, ctx->stream);
        if (!rc) rc = api->AllReduce(ctx->Ql.p, ctx->Ql.p, (size_t)2 * n * m, kNcclDouble, kNcclSum, ctx->nccl_comm, ctx->stream);
        if (rc) return feast_fail(ctx, FEAST_ERR_NCCL, "ncclAllReduce failed");
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    st.t_total_ms = tm.stop();
    if (stats) *stats = st;
    if (st.info) return feast_fail(ctx, FEAST_ERR_SINGULAR, "zero pivot at column %d of a shifted factorisation", st.info);
    return rc_final;
}

int feast_dual_get(feast_ctx* ctx, feast_c128* Xr, int64_t ldr, feast_c128* Xl, int64_t ldl) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, Xr != nullptr, 2, "null Xr");
    ARG_CHECK(ctx, ldr >= ctx->n, 3, "ldr < n");
    ARG_CHECK(ctx, Xl != nullptr, 4, "null Xl");
    ARG_CHECK(ctx, ldl >= ctx->n, 5, "ldl < n");
    FEAST_TRY(download_block(ctx, ctx->X, Xr, ldr));
    return download_block(ctx, ctx->Xl, Xl, ldl);
}

int feast_orthonormalize_X(feast_ctx* ctx) {
    FEAST_TRY(check_ready(ctx, true));
    FEAST_TRY(orthonormalize(ctx, ctx->X, nullptr));  // nlfeast.jl:12-13
    return 0;
}

int feast_beyn_reduce(feast_ctx* ctx, feast_c128* Rf, feast_c128* G1) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, Rf != nullptr, 2, "null Rf");
    ARG_CHECK(ctx, G1 != nullptr, 3, "null G1");
    if (!ctx->Q1.p) return feast_fail(ctx, FEAST_ERR_STATE, "feast_contour_apply (polynomial) has not produced Q1 yet");
    const int m = ctx->m0;
    PhaseTimer tm(ctx, 0);
    std::vector<hc128> Rtot;
    FEAST_TRY(orthonormalize(ctx, ctx->Q, &Rtot));                  // Q0 = U Rtot  (tall svd! of utils.jl:70 -> QR + small SVD)
    memcpy(Rf, Rtot.data(), sizeof(hc128) * m * m);
    c128* G_d = ctx->small_d;
    FEAST_TRY(launch_gram(ctx, ctx->n, m, ctx->Q.p, ctx->Q1.p, G_d));  // U' Q1        utils.jl:71
    CUDA_TRY(ctx, cudaMemcpyAsync(G1, G_d, sizeof(c128) * m * m, cudaMemcpyDeviceToHost, ctx->stream));
    tm.stop();
    return 0;
}

// ------------------------------------------------------------------------- plugin path
int feast_factorize(feast_ctx* ctx, const feast_c128* coef, int ncoef, feast_factor** out) {
    FEAST_TRY(check_ready(ctx, false));
    ARG_CHECK(ctx, coef != nullptr, 2, "null coefficients");
    ARG_CHECK(ctx, ncoef >= 1 && ncoef <= ctx->nslots, 3, "ncoef out of range");
    ARG_CHECK(ctx, out != nullptr, 4, "null output");
    hc128 cf[FEAST_MAX_SLOTS];
    for (int s = 0; s < FEAST_MAX_SLOTS; ++s) cf[s] = s < ncoef ? hc128(coef[s].re, coef[s].im) : hc128(0, 0);
    feast_factor* F = new feast_factor();
    F->kind = effective_solver(ctx);
    for (int s = 0; s < FEAST_MAX_SLOTS; ++s) F->coef[s] = cf[s];
    if (F->kind == FEAST_SOLVER_DENSE_LU) {
        if (!ctx->red_d) {  // getrf scratch when no subspace has been set yet
            FEAST_TRY(dev_alloc(ctx, &ctx->red_d, ((size_t)1 << 20) / sizeof(double)));
            ctx->red_bytes = (size_t)1 << 20;
        }
        int info = 0;
        int rc = factor_dense(ctx, cf, F->lu, &info);
        if (rc) { delete F; return rc; }
        if (info) {
            dev_free(F->lu.lu); dev_free(F->lu.ipiv); dev_free(F->lu.perm); dev_free(F->lu.dinv);
            delete F;
            return feast_fail(ctx, FEAST_ERR_SINGULAR, "zero pivot at column %d", info);
        }
    } else {
        int rc = dev_alloc(ctx, &F->zvals, ctx->unnz);
        if (!rc) rc = assemble_sparse_Z(ctx, cf, F->zvals);
        if (!rc && F->kind == FEAST_SOLVER_BANDED_LU) {
            if (!ctx->red_d) {
                rc = dev_alloc(ctx, &ctx->red_d, ((size_t)1 << 20) / sizeof(double));
                ctx->red_bytes = (size_t)1 << 20;
            }
            int info = 0;
            if (!rc) rc = band_factor(ctx, F->zvals, F->band, &info);
            if (!rc && info) rc = feast_fail(ctx, FEAST_ERR_SINGULAR, "zero pivot at column %d", info);
        }
        if (rc) { band_free(F->band); dev_free(F->zvals); delete F; return rc; }
        F->symmetric = ctx->all_symmetric;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = F;
    return 0;
}

int feast_solve(feast_ctx* ctx, const feast_factor* F, int64_t n, int nrhs, const feast_c128* Bm, int64_t ldb,
                feast_c128* Y, int64_t ldy, int conj_transpose) {
    FEAST_TRY(check_ready(ctx, false));
    ARG_CHECK(ctx, F != nullptr, 2, "null factor");
    ARG_CHECK(ctx, n == ctx->n, 3, "dimension mismatch");
    ARG_CHECK(ctx, nrhs >= 1, 4, "nrhs must be positive");
    ARG_CHECK(ctx, Bm != nullptr, 5, "null right-hand side");
    ARG_CHECK(ctx, ldb >= n, 6, "ldb < n");
    ARG_CHECK(ctx, Y != nullptr, 7, "null solution");
    ARG_CHECK(ctx, ldy >= n, 8, "ldy < n");
    if (ctx->m0 != nrhs) {
        // the block workspace is sized by m0: (re)size it through a zero subspace of the right width
        std::vector<hc128> zero((size_t)n * nrhs, hc128(0, 0));
        FEAST_TRY(feast_set_subspace(ctx, n, nrhs, (const feast_c128*)zero.data(), n));
    }
    FEAST_TRY(ensure_block(ctx, ctx->W1));
    FEAST_TRY(ensure_block(ctx, ctx->W2));
    const int m = nrhs;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->stage, sizeof(c128) * n, Bm, sizeof(c128) * ldb, sizeof(c128) * n, m,
                                    cudaMemcpyHostToDevice, ctx->stream));
    FEAST_TRY(ensure_block(ctx, ctx->R));
    c128* rhs = ctx->R.p;
    FEAST_TRY(launch_colmajor_to_rowmajor(ctx, n, m, ctx->stage, n, rhs, ctx->perm_d));
    int rc_final = 0;
    if (F->kind == FEAST_SOLVER_DENSE_LU) {
        FEAST_TRY(dense_getrs(ctx, n, F->lu.lu, F->lu.perm, F->lu.dinv, m, rhs, ctx->W1.p, conj_transpose != 0));
    } else if (F->kind == FEAST_SOLVER_BANDED_LU) {
        if (conj_transpose) return feast_fail(ctx, FEAST_ERR_STATE, "adjoint solves are not available with the banded solver");
        FEAST_TRY(band_solve_refined(ctx, F->band, F->zvals, m, rhs, ctx->W1.p, ctx->W2.p, nullptr, nullptr));
    } else {
        if (conj_transpose && !F->symmetric)
            return feast_fail(ctx, FEAST_ERR_STATE, "adjoint Krylov solve needs a symmetric operator in this build");
        const int method = effective_krylov(ctx);
        FEAST_TRY(ensure_krylov_work(ctx, method));
        KrylovResult kr;
        if (conj_transpose) {
            // Z symmetric: Z^H = conj(Z)  ->  Z^H y = b  <=>  Z conj(y) = conj(b)
            return feast_fail(ctx, FEAST_ERR_STATE, "adjoint Krylov solve is not implemented yet");
        }
        FEAST_TRY(krylov_any(ctx, method, F->coef, F->zvals, rhs, ctx->W1.p, &kr, nullptr));
        if (!kr.converged) rc_final = FEAST_WARN_INNER_MAXIT;
    }
    FEAST_TRY(launch_rowmajor_to_colmajor(ctx, n, m, ctx->W1.p, ctx->stage, n, ctx->perm_d));
    CUDA_TRY(ctx, cudaMemcpy2DAsync(Y, sizeof(c128) * ldy, ctx->stage, sizeof(c128) * n, sizeof(c128) * n, m,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return rc_final;
}

int feast_factor_free(feast_ctx* ctx, feast_factor* F) {  // finalize!(F), src/utils.jl:173
    if (!F) return 0;
    if (ctx) cudaSetDevice(ctx->device);
    dev_free(F->lu.lu); dev_free(F->lu.ipiv); dev_free(F->lu.perm); dev_free(F->lu.dinv);
    dev_free(F->zvals);
    band_free(F->band);
    delete F;
    return 0;
}

// ------------------------------------------------------------------------- kernel-level entries
int feast_apply_operator(feast_ctx* ctx, int slot, int which, feast_c128* Y, int64_t ldy, int reps, float* ms) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, slot >= 0 && slot < ctx->nslots, 2, "slot out of range");
    ARG_CHECK(ctx, which == 0 || which == 1, 3, "which must be 0 (Q) or 1 (X)");
    if (reps < 1) reps = 1;
    const c128* V = which == 0 ? ctx->Q.p : ctx->X.p;
    FEAST_TRY(apply_slot(ctx, slot, V, ctx->R.p));  // warm-up / result
    if (ms) {
        cudaEventRecord(ctx->ev0, ctx->stream);
        for (int r = 0; r < reps; ++r) FEAST_TRY(apply_slot(ctx, slot, V, ctx->R.p));
        cudaEventRecord(ctx->ev1, ctx->stream);
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
        float t = 0;
        cudaEventElapsedTime(&t, ctx->ev0, ctx->ev1);
        *ms = t / reps;
    }
    if (Y) {
        ARG_CHECK(ctx, ldy >= ctx->n, 5, "ldy < n");
        return download_block(ctx, ctx->R, Y, ldy);
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int feast_kernel_bench(feast_ctx* ctx, int which, int reps, float* ms) {
    FEAST_TRY(check_ready(ctx, true));
    ARG_CHECK(ctx, which == 0 || which == 1, 2, "which must be 0 (direction kernel) or 1 (residual update kernel)");
    ARG_CHECK(ctx, reps >= 1, 3, "reps must be positive");
    ARG_CHECK(ctx, ms != nullptr, 4, "null output");
    return krylov_kernel_bench(ctx, which, reps, ms);
}

int feast_sync(feast_ctx* ctx) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    FEAST_TRY(bind_device(ctx));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int feast_timer_start(feast_ctx* ctx) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    FEAST_TRY(bind_device(ctx));
    if (!ctx->sw0) { CUDA_TRY(ctx, cudaEventCreate(&ctx->sw0)); CUDA_TRY(ctx, cudaEventCreate(&ctx->sw1)); }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->sw0, ctx->stream));
    return 0;
}

int feast_timer_stop(feast_ctx* ctx, float* ms) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    ARG_CHECK(ctx, ms != nullptr, 2, "null output");
    if (!ctx->sw0) return feast_fail(ctx, FEAST_ERR_STATE, "feast_timer_start has not been called");
    FEAST_TRY(bind_device(ctx));
    CUDA_TRY(ctx, cudaEventRecord(ctx->sw1, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->sw1));
    CUDA_TRY(ctx, cudaEventElapsedTime(ms, ctx->sw0, ctx->sw1));
    return 0;
}

int64_t feast_launch_count(const feast_ctx* ctx) { return ctx ? ctx->launches : 0; }

int feast_phase_times(feast_ctx* ctx, double* ms3, int reset) {
    ARG_CHECK(ctx, ctx != nullptr, 1, "null context");
    if (ms3) for (int i = 0; i < 3; ++i) ms3[i] = ctx->phase_ms[i];
    if (reset) for (int i = 0; i < 3; ++i) ctx->phase_ms[i] = 0.0;
    return 0;
}

}  // extern "C"
