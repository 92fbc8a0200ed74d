// Kernel launchers (all enqueue on ctx->stream and return a feast status code).
#pragma once
#include "common.cuh"
#include "reorder.h"

// ---- spmm.cu ---------------------------------------------------------------
// Y(n x m, row-major ld=ldy) = S * X(n x m, row-major ld=ldx); S = CSR(rowptr,col,val),
// real (rvals) or complex (cvals) values.  If dot_out != nullptr additionally returns
// dot_out[j] = sum_i X[i,j] * Y[i,j]  (UNconjugated bilinear form, the COCG <p, Zp>).
int launch_spmm(feast_ctx* ctx, int64_t n, int m, const int* rowptr, const int* col,
                const double* rvals, const c128* cvals, const c128* X, int ldx, c128* Y, int ldy,
                c128* dot_out);
// fused multigrid epilogues on the union pattern (see spmm.cu); returns 1 when not applicable (caller falls back)
int launch_spmm_epi(feast_ctx* ctx, int m, const c128* zvals, const c128* X, c128* Y, int mode, const c128* C, const c128* dinv,
                    double omega, c128* dot_out);
// mixed-precision variant: blocks stored as complex64 (n x m0, m0 even), products accumulated in double
int launch_spmm_f32(feast_ctx* ctx, int m0, const c128* zvals, const void* X32, void* Y32, c128* dot_out);
// zvals[e] = sum_i coef[i] * slotvals_i[e]   (K1 / K9: shifted / polynomial assembly)
int launch_assemble_union(feast_ctx* ctx, int64_t unnz, int nslots, const double* const* rv,
                          const c128* const* cv, const hc128* coef, c128* zvals);
// R[:,j] = sum_i lam_j^i (S_i X)[:,j] on the union pattern; fro2[j] = ||T(lam_j)||_F^2
int launch_poly_residual(feast_ctx* ctx, int64_t n, int m, int nslots, const int* rowptr, const int* col,
                         int64_t unnz, const double* const* rv, const c128* const* cv, const c128* lam_d,
                         const c128* X, c128* R, double* fro2_d);
// dense Z (n x n col-major) = scatter of union-pattern values
int launch_scatter_dense(feast_ctx* ctx, int64_t n, const int* rowptr, const int* col, const c128* zvals, c128* Z);

// ---- blockops.cu -----------------------------------------------------------
// C(M x N) = alpha * op(A)(M x K) * B(K x N) + beta * C with arbitrary element strides.
// conjA: use conj(A(i,k)).  splitk > 1: K is split over grid.z, partials reduced (beta must be 0).
int launch_zgemm(feast_ctx* ctx, int M, int N, int64_t K, hc128 alpha, const c128* A, int64_t sAi, int64_t sAk,
                 bool conjA, const c128* B, int64_t sBk, int64_t sBj, hc128 beta, c128* C, int64_t sCi, int64_t sCj);
// G(m x m, column-major ld m, device) = A^H B over n rows (A, B row-major n x m): split-K tall-skinny Gram
int launch_gram(feast_ctx* ctx, int64_t n, int m, const c128* A, const c128* B, c128* G_d);
// Y(n x m) = X(n x m) * Mx(m x m column-major, device); X, Y row-major, must not alias
int launch_update(feast_ctx* ctx, int64_t n, int m, const c128* X, const c128* M_d, c128* Y);
// out[j] = sum_i conj?(a[i,j]) * b[i,j]   (device result, m complex)
int launch_coldot(feast_ctx* ctx, int64_t n, int m, const c128* a, const c128* b, bool conj_a, c128* out_d);
// nrm2[j] = sum_i |a[i,j]|^2 (device result, m doubles)
int launch_colnorm2(feast_ctx* ctx, int64_t n, int m, const c128* a, double* out_d);
// a[:,j] *= s[j] (complex per-column scale, device vector)
int launch_colscale(feast_ctx* ctx, int64_t n, int m, c128* a, const c128* s_d);
// a[:,j] *= 1/sqrt(nrm2[j])   (0 columns left untouched)
int launch_colnormalize(feast_ctx* ctx, int64_t n, int m, c128* a, const double* nrm2_d);
// R = AX - BX * diag(lam); BX may be nullptr meaning B = I (uses X)
int launch_residual_combine(feast_ctx* ctx, int64_t n, int m, c128* AX_inout_R, const c128* BX, const c128* lam_d);
// Q += (X - Y) * diag(d)     [linear accumulate, src/feast.jl:68-70]; if Q1 != nullptr also Q1 += z*(...)
// first_pass: term = Y * w (nlfeast.jl:39-45, d[j] = w for all j)
int launch_accumulate(feast_ctx* ctx, int64_t n, int m, const c128* X, const c128* Y, const c128* d_d,
                      c128* Q, c128* Q1, hc128 z, bool first_pass);
// Column-slice / moment form of the accumulation: Y is a compact n x mloc block (solutions of columns j0 .. j0+mloc-1),
// moments[p][:, j0+jj] += z^p * term for p < nmom.
int launch_accumulate_slice(feast_ctx* ctx, int64_t n, int m, int j0, int mloc, const c128* X, const c128* Y, const c128* d_d,
                            c128* const* moments, int nmom, hc128 z, bool first_pass);
int launch_axpy(feast_ctx* ctx, int64_t count, const c128* a, c128* y);   // y += a
// dst (compact n x mloc) = src[:, j0 : j0 + mloc]
int launch_gather_cols(feast_ctx* ctx, int64_t n, int m, int j0, int mloc, const c128* src, c128* dst);
// layout conversion between host column-major (ld) and device row-major blocks
// perm (device, new -> old row map, may be nullptr): device row i holds host row perm[i]
int launch_colmajor_to_rowmajor(feast_ctx* ctx, int64_t n, int m, const c128* src, int64_t ld, c128* dst, const int* perm = nullptr);
int launch_rowmajor_to_colmajor(feast_ctx* ctx, int64_t n, int m, const c128* src, c128* dst, int64_t ld, const int* perm = nullptr);
// ---- spmm.cu: capacities of the tiled SpMM (shared-memory rows / staged nonzeros / rows per tile)
struct TileCaps;
TileCaps spmm_tile_caps();
int spmm_tile_cfg();   // tile configuration the capacities belong to (FEAST_TILE_CFG)
int launch_real_to_complex(feast_ctx* ctx, int64_t count, const double* src, c128* dst);
int launch_conj(feast_ctx* ctx, int64_t count, const c128* src, c128* dst);  // dst = conj(src), may alias
// Z(n x n) = sum_i coef[i] * D_i (dense col-major slots; identity slots add coef to the diagonal)
int launch_assemble_dense(feast_ctx* ctx, int64_t n, int nslots, const c128* const* D, const int* kinds,
                          const hc128* coef, c128* Z);
// fro2[j] = || sum_i lam_j^i D_i ||_F^2 for dense slots
int launch_poly_fro_dense(feast_ctx* ctx, int64_t n, int m, int nslots, const c128* const* D, const int* kinds,
                          const c128* lam_d, double* fro2_d);

// ---- krylov.cu ----------------------------------------------------------------
struct KrylovResult { int iters; double relres_max; bool converged; double spmm_ms; int spmm_launches; };
// Solve Z Y = Rhs for all m columns with pseudo-block COCG (complex symmetric Z) or BiCGStab.
// Z given by the union pattern + zvals.  Y, Rhs row-major n x m.  Y is overwritten (zero start).
int krylov_solve(feast_ctx* ctx, int method, const c128* zvals, const c128* Rhs, c128* Y,
                 double tol, int maxit, KrylovResult* out);

// COCG with complex64 storage of the Krylov blocks (mixed_prec); needs m0 even and the default tile plan
int krylov_solve_mixed(feast_ctx* ctx, const c128* zvals, const c128* Rhs, c128* Y, double tol, int maxit, KrylovResult* out);
// COCG preconditioned by the smoothed-aggregation V-cycle (needs ctx->amg assembled for the node)
int krylov_solve_pcocg(feast_ctx* ctx, const c128* zvals, const c128* zvals_pc, const c128* Rhs, c128* Y, double tol, int maxit,
                       KrylovResult* out);
int krylov_kernel_bench(feast_ctx* ctx, int which, int reps, float* ms);
int gmres_solve(feast_ctx* ctx, const c128* zvals, const c128* Rhs, c128* Y, double tol, int maxit, KrylovResult* out);
size_t gmres_small_bytes(int m, int R);

// ---- amg.cu: smoothed-aggregation preconditioner of the Krylov inner solves
struct AmgHost;
int amg_max_coarse();
int amg_build(feast_ctx* ctx, AmgHost& H, int nslots, const std::vector<int>& order, const std::vector<int>& dpos0, std::string* why);
void amg_free(feast_ctx* ctx);
int amg_ensure_blocks(feast_ctx* ctx);
int amg_assemble(feast_ctx* ctx, const hc128* coef, const c128* zvals0, int node, int* info);   // node >= 0: cache the coarse inverse
void amg_drop_cache(feast_ctx* ctx);
int amg_apply(feast_ctx* ctx, const c128* zvals0, const c128* r, c128* y, c128* t, c128** out, c128* dot_rz, bool* dot_done);
int amg_info(const feast_ctx* ctx, int* nlevels, int* sizes, int cap, double* setup_seconds);

// ---- dense.cu ------------------------------------------------------------------
// In-place LU with partial pivoting of ROW-major n x n Z (Z(i,j) = Z[i*n + j]); ipiv device 0-based.
int dense_getrf(feast_ctx* ctx, int64_t n, c128* Z, int* ipiv_d, int* info_out);
// Solve op(LU) Y = Rhs for m right-hand sides held ROW-MAJOR (n x m); result in Y (row-major).
// perm_d from dense_build_perm (perm[i] = source row of permuted row i).
// dinv: diagonal-block inverses from dense_build_diag_inverses (may be nullptr -> recursive TRSM)
int dense_getrs(feast_ctx* ctx, int64_t n, const c128* LU, const int* perm_d, const c128* dinv, int m, const c128* Rhs,
                c128* Y, bool conj_transpose, c128* work = nullptr);   // work: n x m scratch (default ctx->W2)
// ---- band.cu: block-tridiagonal direct solver for banded sparse operators
struct BandFactor;
int band_factor(feast_ctx* ctx, const c128* zvals, BandFactor& F, int* info);
int band_solve(feast_ctx* ctx, const BandFactor& F, int m, const c128* Rhs, c128* Y);
void band_free(BandFactor& F);
// band_solve + iterative refinement against the assembled sparse operator (work: n x m scratch)
int band_solve_refined(feast_ctx* ctx, const BandFactor& F, const c128* zvals, int m, const c128* Rhs, c128* Y, c128* work,
                       int* steps_out, double* relres_out);
int dense_build_diag_inverses(feast_ctx* ctx, int64_t n, const c128* LU, c128* dinv);
// pivoted LU of the first w columns of a rows x ncols row-major window (interchanges applied to all ncols columns)
int dense_getrf_rect(feast_ctx* ctx, int rows, int w, int ncols, c128* Z, int64_t lda, int* ipiv_d, int* info_out);
// B (h x ncols, row-major ldb) <- T^-1 B, T = lower-unit or upper-non-unit triangle of a row-major matrix (ldt)
int dense_trsm(feast_ctx* ctx, bool lower_unit, int h, int ncols, const c128* T, int64_t ldt, c128* B, int64_t ldb);
int dense_build_perm(feast_ctx* ctx, int64_t n, const int* ipiv_d, int* perm_d);
size_t spmm_partials_bytes(int m);
int debug_check_finite(feast_ctx* ctx, const void* p, int64_t ndoubles, const char* name);
