/*
 * feast_cuda.h -- C ABI of libfeast_cuda.so
 *
 * B200-native (sm_100a) replacement for the contour-quadrature hot path of
 * spacedome/FEASTSolver.jl.  The reference is pure Julia; the only native seam
 * it has is the plugin pair `factorizer` / `left_divider` funnelled through
 * `linsolve!` (src/utils.jl:173-179) plus direct LinearAlgebra calls inside the
 * drivers.  Each entry point below cites the reference statements it replaces
 * (file:line into the reference tree).  The Julia-side bindings (`ccall`) are
 * in julia/FEASTSolverB200.jl and INTEGRATION.md; the Python ctypes binding is
 * feastsolver_jl_b200/_lib.py.
 *
 * Conventions
 *   - complex numbers are interleaved (re, im) doubles == Julia ComplexF64 ==
 *     cuDoubleComplex; host matrices are COLUMN-MAJOR with explicit leading
 *     dimension (Julia `Matrix`), sparse matrices are CSC (Julia SparseMatrixCSC:
 *     colptr/rowval/nzval, index base given by the caller, 1 for Julia).
 *   - every function returns int: 0 ok; -k = k-th argument invalid (LAPACK
 *     style, cf. src/lapack.jl:77,90); FEAST_ERR_* > 0 otherwise.  Nothing
 *     throws, exits or aborts across the boundary.  feast_last_error() gives
 *     the message the Julia shim passes to error().
 *   - the caller owns all host arrays; they need only stay valid for the
 *     duration of the call.  The library owns all device memory, streams and
 *     NCCL communicators behind feast_ctx.
 *   - a context is single-threaded by contract; distinct contexts are
 *     independent.  There is NO CPU fallback: without a CUDA device every
 *     compute entry fails with FEAST_ERR_CUDA.
 *   - the m0 x m0 reduced eigenproblem / SVD is NOT in the ABI: the caller does
 *     it with its own LAPACK (eigen!/svd! in Julia, scipy in the harness).
 */
#ifndef FEAST_CUDA_H
#define FEAST_CUDA_H

#include <stdint.h>

#if defined(__GNUC__)
#define FEAST_API __attribute__((visibility("default")))
#else
#define FEAST_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct feast_ctx feast_ctx;
typedef struct { double re, im; } feast_c128;

/* ---- status codes --------------------------------------------------- */
#define FEAST_OK                 0
#define FEAST_ERR_CUDA           1000  /* CUDA runtime error / no device       */
#define FEAST_ERR_NCCL           1001  /* NCCL error / libnccl not loadable    */
#define FEAST_ERR_OOM            1002  /* out of device memory                 */
#define FEAST_ERR_STATE          1003  /* call sequence invalid (e.g. no A set)*/
#define FEAST_ERR_SINGULAR       1004  /* exact zero pivot (getrf info>0); the */
                                       /* column is in feast_stats.info        */
#define FEAST_WARN_INNER_MAXIT   2000  /* Krylov hit max_inner (non fatal:     */
                                       /* result produced, stats carry resid.) */

/* ---- enums ---------------------------------------------------------- */
enum { FEAST_SOLVER_AUTO = 0, FEAST_SOLVER_DENSE_LU = 1, FEAST_SOLVER_KRYLOV = 2,
       FEAST_SOLVER_BANDED_LU = 3 /* block-tridiagonal direct solver for banded sparse operators */ };
enum { FEAST_KRYLOV_AUTO = 0, FEAST_KRYLOV_COCG = 1, FEAST_KRYLOV_BICGSTAB = 2, FEAST_KRYLOV_GMRES = 3 };
enum { FEAST_PROBLEM_STANDARD = 0,    /* A x = l x          feast!      src/feast.jl:10-80   */
       FEAST_PROBLEM_GENERALIZED = 1, /* A x = l B x        gen_feast!  src/feast.jl:89-156  */
       FEAST_PROBLEM_POLYNOMIAL = 2,  /* sum l^i A_i x = 0  nlfeast!    src/nlfeast.jl:2-84  */
       FEAST_PROBLEM_SAMPLED = 3      /* T(l) x = 0 with T an opaque callable of the caller (the closure form
                                         nlfeast!(T::Function, ...), src/nlfeast.jl:2-4): slot 0 holds T at ONE point,
                                         re-uploaded by the caller for every node / Ritz value                       */ };
/* Preconditioner of the Krylov inner solves: a smoothed-aggregation multigrid V-cycle built from the real symmetric
 * operator slots (linear problems, COCG).  AUTO = use it when applicable and n is large enough to pay for the setup. */
enum { FEAST_PRECOND_NONE = 0, FEAST_PRECOND_AMG = 1, FEAST_PRECOND_AUTO = 2 };
/* multi-GPU sharding axis of the contour loop: nodes (src/feast.jl:34 threads over them), or the columns of every
 * node's right-hand side (Krylov inner solves only: no factorisation to keep together, perfect balance); AUTO picks
 * columns for Krylov solves and nodes for the direct solvers */
enum { FEAST_SHARD_AUTO = 0, FEAST_SHARD_NODES = 1, FEAST_SHARD_COLUMNS = 2 };
#define FEAST_MAX_MOMENTS 8          /* moment accumulators S_p = sum_k w_k z_k^p (...) of one contour pass */
#define FEAST_MAX_SLOTS 8            /* slot 0 = A (or A_0), 1 = B (or A_1), ... A_7 */

typedef struct {
    int    nodes_local;       /* contour nodes solved by this rank                  */
    int    inner_iters_total; /* Krylov iterations summed over local nodes          */
    int    inner_iters_max;   /* max over local nodes                               */
    int    info;              /* LU: first zero-pivot column (1-based) or 0         */
    double inner_relres_max;  /* max over nodes/columns of ||r||/||b|| reached      */
    double t_factor_ms;       /* device time in factorisation / operator assembly   */
    double t_solve_ms;        /* device time in solves + fused accumulation         */
    double t_reduce_ms;       /* device time in the NCCL all-reduce                 */
    double t_total_ms;
    double t_spmm_ms;         /* Krylov: device time inside the SpMM launches       */
    int64_t spmm_launches;    /* Krylov: number of SpMM launches                    */
    int    precond_levels;    /* levels of the multigrid preconditioner in use (0: unpreconditioned) */
    int    col_sharded;       /* 0: nodes sharded over the ranks; s >= 1: column mode, every node ran on s ranks (a group),
                                 each solving m0 / s of its right-hand-side columns                              */
} feast_stats;

/* ---- library / context ---------------------------------------------- */
FEAST_API int  feast_version(void);
FEAST_API int  feast_device_count(int* count);
/* One context drives ONE GPU (one process per GPU under torchrun, or several
 * contexts in one Julia process).  */
FEAST_API int  feast_ctx_create(feast_ctx** out, int device);
FEAST_API int  feast_ctx_destroy(feast_ctx* ctx);
FEAST_API const char* feast_last_error(const feast_ctx* ctx); /* ctx may be NULL: last global error */

/* ---- contour constructors (host, scalar) ----------------------------
 * Replace src/contour.jl:26-31, 33-44, 47-63, 66-86 (same node order, weights,
 * divisibility errors -> -3, "Invalid corners" -> -1).  z, w: caller arrays of N. */
FEAST_API int  feast_contour_circular_trapezoidal(feast_c128 c, double r, int N, feast_c128* z, feast_c128* w);
FEAST_API int  feast_contour_circular_gauss(feast_c128 c, double r, int N, feast_c128* z, feast_c128* w);
FEAST_API int  feast_contour_rectangular_gauss(feast_c128 bottom_left, feast_c128 top_right, int N, feast_c128* z, feast_c128* w);
FEAST_API int  feast_contour_rectangular_trapezoidal(feast_c128 bottom_left, feast_c128 top_right, int N, feast_c128* z, feast_c128* w);
/* Gauss-Legendre rule on [-1,1] (FastGaussQuadrature.gausslegendre, call sites contour.jl:37,52) */
FEAST_API int  feast_gauss_legendre(int n, double* x, double* w);

/* ---- operators (uploaded once, device resident) ---------------------
 * Replace the `A`, `B`, `T(z)` arguments of the drivers.  Slots: 0 = A, 1 = B for
 * linear problems; i = A_i for polynomial problems T(z) = sum z^i A_i
 * (test/butterfly.jl:61, test/polynomial.jl:9-11).                           */
FEAST_API int  feast_set_dense(feast_ctx* ctx, int slot, int64_t n, const void* a, int64_t lda, int is_complex);
FEAST_API int  feast_set_csc(feast_ctx* ctx, int slot, int64_t n, const int64_t* colptr, const int64_t* rowval,
                   const void* nzval, int is_complex, int index_base);
FEAST_API int  feast_set_identity(feast_ctx* ctx, int slot, int64_t n); /* B = I (UniformScaling) */
/* kind: FEAST_PROBLEM_*; nslots = 1 (standard), 2 (generalized), degree+1 (polynomial). */
FEAST_API int  feast_set_problem(feast_ctx* ctx, int kind, int nslots);

/* ---- contour, solver, sharding --------------------------------------- */
FEAST_API int  feast_set_contour(feast_ctx* ctx, int nnodes, const feast_c128* z, const feast_c128* w);
/* kind: FEAST_SOLVER_*; krylov: FEAST_KRYLOV_*; inner_tol relative to ||R_j||;
 * store != 0 keeps LU factors of all local nodes (src/feast.jl:28-38).        */
FEAST_API int  feast_set_solver(feast_ctx* ctx, int kind, int krylov, double inner_tol, int max_inner, int store);
/* Multi-GPU: contour nodes are sharded over ranks (the axis the reference threads
 * over, src/feast.jl:34, src/nlfeast.jl:19,36); Q is summed with ncclAllReduce.
 * id128: 128-byte ncclUniqueId produced on rank 0 and broadcast by the caller.   */
FEAST_API int  feast_comm_unique_id(void* id128);
FEAST_API int  feast_comm_init(feast_ctx* ctx, int nranks, int rank, const void* id128);
/* owner[k] = rank that solves node k (NULL -> balanced default). */
FEAST_API int  feast_set_node_owners(feast_ctx* ctx, int nnodes, const int* owner);

/* ---- subspace --------------------------------------------------------- */
/* Upload X0 (n x m0 column-major, ldx >= n): becomes Q (and X).  src/feast.jl:21 */
FEAST_API int  feast_set_subspace(feast_ctx* ctx, int64_t n, int m0, const feast_c128* X, int64_t ldx);
FEAST_API int  feast_set_X(feast_ctx* ctx, const feast_c128* X, int64_t ldx); /* overwrite the block X only */
FEAST_API int  feast_get_X(feast_ctx* ctx, feast_c128* X, int64_t ldx);   /* normalised Ritz vectors */
FEAST_API int  feast_get_Q(feast_ctx* ctx, feast_c128* Q, int64_t ldq);   /* current basis / accumulator */
FEAST_API int  feast_get_R(feast_ctx* ctx, feast_c128* R, int64_t ldr);   /* residual vectors */

/* ---- per-iteration phases (coarse-grained, the graded path) ----------- */
/* Q <- orth(Q); Aq = Q'AQ [, Bq = Q'BQ] (m0 x m0 column-major, ld m0).
 * Replaces src/feast.jl:41-43 and :117-121 (qr, mul!, mul!).  Bq may be NULL
 * for the standard problem.                                                  */
FEAST_API int  feast_project(feast_ctx* ctx, feast_c128* Aq, feast_c128* Bq);
/* X = Q Xq; x_j /= ||x_j||; R_j = (A - l_j B) x_j (or T(l_j) x_j); res_j = ||R_j||
 * (absolute; polynomial: relative to ||T(l_j)||_F).  Replaces src/feast.jl:48-50,
 * :125-127, src/utils.jl:104-116,151-157,166-171.                             */
/* Xq may be NULL: X is then taken as it is (normalised in place), e.g. after feast_moment_combine */
FEAST_API int  feast_recover_residual(feast_ctx* ctx, const feast_c128* Xq, const feast_c128* lambda, double* res);
/* The hot loop.  Linear: Q = sum_k (X - (A - z_k B)^-1 R) diag(w_k/(z_k - l_j))
 * (src/feast.jl:57-71, :134-147).  Polynomial: Q0, Q1 of src/nlfeast.jl:36-61;
 * first_pass != 0 selects the Beyn pass w_k T(z_k)^-1 X (nlfeast.jl:39-45).
 * Node-sharded + all-reduced when a communicator is attached.  stats may be NULL.
 * Returns FEAST_WARN_INNER_MAXIT (non-fatal) when a Krylov solve stopped early. */
FEAST_API int  feast_contour_apply(feast_ctx* ctx, const feast_c128* lambda, int first_pass, feast_stats* stats);
/* Polynomial problems: Q0 = U Rf (U orthonormal, replaces the tall svd!(Q0) of
 * src/utils.jl:70); returns Rf (m0 x m0) and G1 = U' Q1 (utils.jl:71).  The caller
 * finishes the m0 x m0 SVD/eig (utils.jl:72-76) and calls feast_recover_residual
 * with Xq = Ur * vecs.                                                          */
/* sampled problems (opaque T(z)): replace the sample in slot 0, apply ONE contour node with it (phase bit 1 = first
 * node of the pass, bit 2 = last; k = -1: this rank owns no node, collective part only), finish the residual of
 * column j with the sample T(l_j) (fro = ||T(l_j)||_F from the caller).  src/nlfeast.jl:36-61, src/utils.jl:104-109,151-157 */
FEAST_API int  feast_set_sample_dense(feast_ctx* ctx, int64_t n, const void* a, int64_t lda, int is_complex);
FEAST_API int  feast_set_sample_csc(feast_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval,
                                    const void* nzval, int is_complex, int index_base);
FEAST_API int  feast_contour_node(feast_ctx* ctx, int k, const feast_c128* lambda, int first_pass, int phase, feast_stats* stats);
FEAST_API int  feast_node_needs_sample(const feast_ctx* ctx, int k);   /* 0 when a stored factorisation of node k exists */
FEAST_API int  feast_sampled_residual(feast_ctx* ctx, int j, double fro, double* res);
FEAST_API int  feast_set_sharding(feast_ctx* ctx, int mode);                 /* FEAST_SHARD_* */
/* moment machinery of the one-shot contour solvers (src/beyn.jl:2-94) and nlfeast_moments! (src/nlfeast.jl:173-318) */
FEAST_API int  feast_set_moments(feast_ctx* ctx, int nmom);
FEAST_API int  feast_block_gram(feast_ctx* ctx, int a, int b, feast_c128* G); /* ids: p >= 0 moment S_p, -1 X, -2 R */
FEAST_API int  feast_moment_combine(feast_ctx* ctx, int nblk, const feast_c128* W, int64_t ldw);
FEAST_API int  feast_last_fro(feast_ctx* ctx, double* fro);
FEAST_API int  feast_beyn_reduce(feast_ctx* ctx, feast_c128* Rf, feast_c128* G1);
/* contour_estimate_eig (src/stochastic.jl:2-33): with the probe vectors uploaded by
 * feast_set_subspace, est = Re sum_k w_k tr(X' (z_k B - A)^-1 X) / m0 (node-sharded).   */
FEAST_API int  feast_estimate_count(feast_ctx* ctx, double* est, feast_stats* stats);
/* ---- two-sided driver dual_gen_feast! (src/feast.jl:165-257) -------------------------
 * Right blocks are the ones of the one-sided path; left blocks Ql/Xl/Rl are added.  The m0 x m0
 * SVD bi-orthogonalisation (feast.jl:199-201) and the two reduced eigenproblems (:206,:210) stay
 * on the host.  One LU per node serves A - zB and its adjoint (getrs 'C').                 */
FEAST_API int  feast_dual_set_subspace(feast_ctx* ctx, int64_t n, int m0, const feast_c128* Xr, int64_t ldr,
                             const feast_c128* Xl, int64_t ldl);
FEAST_API int  feast_dual_project(feast_ctx* ctx, feast_c128* G);            /* G = Ql' B Qr          feast.jl:199     */
FEAST_API int  feast_dual_rotate(feast_ctx* ctx, const feast_c128* Mr, const feast_c128* Ml,
                       feast_c128* Aq, feast_c128* Bq);                /* feast.jl:200-205                       */
FEAST_API int  feast_dual_recover_residual(feast_ctx* ctx, const feast_c128* Xqr, const feast_c128* Xql,
                                 const feast_c128* lambda, double* resr);  /* feast.jl:207-215                  */
FEAST_API int  feast_dual_contour_apply(feast_ctx* ctx, const feast_c128* lambda, feast_stats* stats); /* :225-247 */
FEAST_API int  feast_dual_get(feast_ctx* ctx, feast_c128* Xr, int64_t ldr, feast_c128* Xl, int64_t ldl);
/* nlfeast! first statement: X <- thin Q of X (src/nlfeast.jl:12-13).           */
FEAST_API int  feast_orthonormalize_X(feast_ctx* ctx);

/* ---- fine-grained plugin path (works with the UNMODIFIED reference drivers) --
 * factorizer(C) / left_divider(Y, F, X) / finalize!(F)  (src/utils.jl:173-179):
 * feast_factorize builds F for the shifted operator sum_i coef[i] * slot_i
 * (coef = {1, -z} gives A - zB, feast.jl:64,141); feast_solve is ldiv!.        */
typedef struct feast_factor feast_factor;
FEAST_API int  feast_factorize(feast_ctx* ctx, const feast_c128* coef, int ncoef, feast_factor** out);
FEAST_API int  feast_solve(feast_ctx* ctx, const feast_factor* F, int64_t n, int nrhs,
                 const feast_c128* Bm, int64_t ldb, feast_c128* Y, int64_t ldy, int conj_transpose);
FEAST_API int  feast_factor_free(feast_ctx* ctx, feast_factor* F);

/* ---- kernel-level entries used by the parity tests and bench.py ---------- */
/* Y = op(slot) * V for the current n x m0 block held in Q (which=0) or X (1);
 * result left in R and optionally downloaded.  Times `reps` launches with CUDA
 * events on the library stream and returns the mean in *ms (NULL ok).        */
FEAST_API int  feast_apply_operator(feast_ctx* ctx, int slot, int which, feast_c128* Y, int64_t ldy, int reps, float* ms);
/* synchronise the library stream */
/* measurement only: mean device time of `reps` stand-alone launches of a vector kernel of the Krylov iteration on the
 * resident work blocks (which = 0: x += alpha p ; p = z + beta p, 1: r -= alpha q + norm partials) */
FEAST_API int  feast_kernel_bench(feast_ctx* ctx, int which, int reps, float* ms);
FEAST_API int  feast_sync(feast_ctx* ctx);
/* CUDA-event stopwatch on the library stream (the stream every kernel of this context is
 * launched on): start records an event, stop records a second one, synchronises and returns
 * the elapsed device-timeline milliseconds (host gaps between launches included).     */
FEAST_API int  feast_timer_start(feast_ctx* ctx);
FEAST_API int  feast_timer_stop(feast_ctx* ctx, float* ms);
/* number of kernels this context has launched so far (bench.py gpu_launches) */
FEAST_API int64_t feast_launch_count(const feast_ctx* ctx);
/* device-timed phases since the last reset (ms): [0]=project [1]=recover [2]=contour_apply */
FEAST_API int  feast_phase_times(feast_ctx* ctx, double* ms3, int reset);
/* The reference's `mixed_prec=true` (src/feast.jl:19-25, ComplexF32 solves inside the double-precision RII loop) for the
 * Krylov path: the COCG blocks are stored in complex64 (half the HBM traffic), arithmetic stays double; the recurrence is
 * the unpreconditioned one.  Applies when the inner solver is COCG, m0 is even and the default tile plan is in use;
 * ignored otherwise.  (Validated on a B200 in round 2: tests/test_gpu_parity.py::test_feast_mixed_prec_krylov.) */
FEAST_API int  feast_set_mixed_precision(feast_ctx* ctx, int on);
/* kind: FEAST_PRECOND_*.  Takes effect immediately (the device layout is rebuilt if a hierarchy has to be added or dropped). */
FEAST_API int  feast_set_preconditioner(feast_ctx* ctx, int kind);
/* complex-shifted preconditioner: the hierarchy is assembled at z + i*beta*|z|*sign(Im z) instead of the node z
 * (beta = 0 default; ~0.5 for contours in the interior of the spectrum) */
FEAST_API int  feast_set_preconditioner_shift(feast_ctx* ctx, double beta);
/* levels in use (0 = none), their sizes (up to cap entries) and the host setup time of the hierarchy */
FEAST_API int  feast_preconditioner_info(const feast_ctx* ctx, int* nlevels, int* sizes, int cap, double* setup_seconds);
/* Internal layout of the sparse path (no reference counterpart: UMFPACK reorders internally as
 * well).  The rows are cut into tiles whose referenced rows of the n x m0 block fit in shared
 * memory (tiled SpMM); with Krylov inner solves the rows are renumbered so that the tiles are
 * compact.  info[0] = 1 if renumbered, info[1] = number of tiles, info[2] = max |i - j| of the
 * union pattern in the NATURAL ordering, info[3] = 1 if the tiled SpMM is in use; halo = rows
 * fetched from outside a tile per row.  Blocks cross the ABI in the caller's ordering.          */
FEAST_API int  feast_layout_info(const feast_ctx* ctx, int* info4, double* halo);
/* Host-only (no device): the tile plan the library would build for a 0-based CSR pattern with the
 * given capacities; order[i_new] = i_old (may be NULL).  Returns 1 if a single row exceeds them,
 * 2 / 3 if the plan fails its self-check (capacities / tile-local column numbers).              */
FEAST_API int  feast_debug_tile_plan(int64_t n, const int64_t* rowptr, const int* col, int reorder, int rows_cap,
                                     int nnz_cap, int tile_max, int domain_rows, int* order, int* ntiles,
                                     double* halo_ratio);
/* Host-only restatement of the library's orthonormalisation (iterated, shifted Cholesky-QR; replaces
 * `qr(Q).Q`, src/feast.jl:41, and the tall `svd!(Q0)` of src/utils.jl:70) for CPU regression tests:
 * V (n x m column-major, ldv) is overwritten with the orthonormal factor, V_in = V_out * Rtot.     */
/* host-only: smoothed-aggregation hierarchy of the Krylov preconditioner (amg_setup.cpp), for CPU tests of the setup */
FEAST_API void* feast_debug_amg_build(int64_t n, const int64_t* rowptr, const int* col, int nslots, const double* vals_flat,
                                      int max_coarse, int* nlevels, double* seconds);
FEAST_API int  feast_debug_amg_level_info(const void* handle, int lev, int* n, int* nnz, int* nc, int* pnnz, double* rho);
FEAST_API int  feast_debug_amg_level_get(const void* handle, int lev, int* rowptr, int* col, double* vals_flat,
                                         int* p_rowptr, int* p_col, double* p_val);
FEAST_API void feast_debug_amg_free(void* handle);
/* host-only: rank groups of the column-sharded contour loop for given node costs; returns the number of groups */
FEAST_API int  feast_debug_pick_groups(int nnodes, const double* cost, int nranks, int m0, int* group_of_node);
FEAST_API int  feast_debug_cholqr(int64_t n, int m, feast_c128* V, int64_t ldv, feast_c128* Rtot, int* passes);

#ifdef __cplusplus
}
#endif
#endif /* FEAST_CUDA_H */
