/*
 * libhsbp -- C-ABI of the B200-native (sm_100a) hybridized-SBP solve path.
 *
 * This is the drop-in boundary for ONE path of brittany-erickson/HybridSBP: the
 * block-local curvilinear SBP operator, the per-block local solves, the trace
 * (lambda) operators and solve, and the BP1 rate-and-state stage.  The reference
 * is pure Julia with no FFI of its own; every entry point below names the
 * reference interface (file:line under the reference tree) whose *work* it
 * replaces, and INTEGRATION.md shows the `ccall` stubs a maintainer would add.
 *
 * Conventions (same as the reference):
 *   - values are fp64, ids / sizes / offsets are int64;
 *   - a block field is (Nr+1) x (Ns+1), column-major, r fastest; blocks are
 *     concatenated in block order (the reference's `vstarts` layout);
 *   - ids held in arrays (FToE, FToLF, EToS ...) are 1-based exactly as the
 *     reference produces them; the library converts internally;
 *   - boundary-condition codes: 0 locked interface, 1 Dirichlet, 2 Neumann,
 *     >= 7 jump interface (global_curved.jl:13-16);
 *   - face data of one block is laid out face 1..4, face k having Ns+1 points
 *     (k = 1, 2) or Nr+1 points (k = 3, 4).
 *
 * Every function returns 0 on success, < 0 for an argument error, > 0 for a
 * CUDA / NCCL failure; nothing throws or exits.  `hsbp_last_error` gives the
 * text.  Pointers named *_dev are device pointers (from hsbp_malloc or any CUDA
 * allocation of the same device); all others are host pointers which the
 * library only reads/writes during the call.
 *
 * There is no CPU fallback anywhere behind this header.
 */
#ifndef HSBP_H
#define HSBP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSBP_BC_LOCKED    0
#define HSBP_BC_DIRICHLET 1
#define HSBP_BC_NEUMANN   2
#define HSBP_BC_JUMP      7

#define HSBP_OK            0
#define HSBP_ERR_ARG      -1
#define HSBP_ERR_STATE    -2
#define HSBP_ERR_UNSUPP   -3
#define HSBP_ERR_CUDA      1
#define HSBP_ERR_NCCL      2

typedef struct hsbp_ctx    hsbp_ctx;     /* one device + its streams                         */
typedef struct hsbp_blocks hsbp_blocks;  /* block-local operators of a set of blocks          */
typedef struct hsbp_trace  hsbp_trace;   /* trace (lambda) operators + Schur-complement solve */

/* ---- context / memory --------------------------------------------------- */
int  hsbp_version(void);
int  hsbp_ctx_create(int device, hsbp_ctx **ctx);
int  hsbp_ctx_destroy(hsbp_ctx *ctx);
const char *hsbp_last_error(hsbp_ctx *ctx);      /* valid until the next call on ctx */
int  hsbp_malloc(hsbp_ctx *ctx, size_t bytes, void **dptr);
int  hsbp_free(hsbp_ctx *ctx, void *dptr);
int  hsbp_h2d(hsbp_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int  hsbp_d2h(hsbp_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
int  hsbp_memset0(hsbp_ctx *ctx, void *dst_dev, size_t bytes);
int  hsbp_sync(hsbp_ctx *ctx);
int  hsbp_host_register(hsbp_ctx *ctx, void *host, size_t bytes);    /* pin a caller buffer */
int  hsbp_host_unregister(hsbp_ctx *ctx, void *host);
/* the CUDA stream all work of ctx is issued on (a cudaStream_t), for callers that time with events */
void *hsbp_stream(hsbp_ctx *ctx);
/* elapsed-time helpers on that stream (CUDA events owned by the library) */
int  hsbp_timer_start(hsbp_ctx *ctx);
int  hsbp_timer_stop(hsbp_ctx *ctx, double *milliseconds);  /* synchronises */

/* ---- block-local operator: replaces locoperator's assembled M-tilde ------
 * reference: locoperator, global_curved.jl:211-506 (sparse M-tilde :470-486,
 * F_k :455-458, tau :418-442); variable_diagonal_sbp_D2, diagonal_sbp.jl:474-764.
 * Nothing is assembled: the library keeps crr/css/crs, tau and bc codes and
 * applies M-tilde matrix-free.  p in {2, 4, 6} (global_curved.jl:402-416).      */
int  hsbp_blocks_create(hsbp_ctx *ctx, int p, int64_t nblocks,
                        const int64_t *Nr, const int64_t *Ns, hsbp_blocks **blocks);
int  hsbp_blocks_destroy(hsbp_blocks *blocks);
int64_t hsbp_blocks_num_volume_points(const hsbp_blocks *blocks);   /* VNp  */
int64_t hsbp_blocks_num_face_points(const hsbp_blocks *blocks);     /* sum over blocks and faces */
/* coefficient fields from create_metrics (global_curved.jl:160-162), concatenated by block */
int  hsbp_blocks_set_metrics(hsbp_blocks *blocks, const double *crr, const double *css,
                             const double *crs);
/* same, from device memory (e.g. synthetic meshes generated on the device) */
int  hsbp_blocks_set_metrics_dev(hsbp_blocks *blocks, const double *crr_dev,
                                 const double *css_dev, const double *crs_dev);
/* Geometry on the device.  The reference evaluates transfinite_blend (global_curved.jl:19-51) and
 * create_metrics (:136-209) on the host with callbacks per block; here the host supplies only the O(N) edge data:
 *   hsbp_blocks_blend_dev         x, x_r, x_s of every block from its four edge curves sampled at the grid points; edges_dev
 *                                 holds per block [a1(s_j) | a2(s_j) | a3(r_i) | a4(r_i)] followed by the same four for the
 *                                 derivatives a1', a2', a3', a4' (twice the block-face layout); corner consistency as :25
 *   hsbp_blocks_set_geometry_dev  create_metrics: J (> 0 asserted, :157), crr / css / crs into the blocks' coefficient fields
 *                                 from x_r, x_s, y_r, y_s; optional outputs J (volume layout) and sJ, nx, ny (block-face
 *                                 layout, :166-200) -- pass NULL to skip
 *   hsbp_blocks_set_synthetic_warp  the analytic warped mesh of SURVEY.md section 8d for blocks numbered bx + nbx * by, block
 *                                 columns bx0 .. bx0 + nbx - 1 of a mesh whose warp period is L: nothing crosses PCIe       */
int  hsbp_blocks_blend_dev(hsbp_blocks *blocks, const double *edges_dev, double *x_dev, double *xr_dev, double *xs_dev);
int  hsbp_blocks_set_geometry_dev(hsbp_blocks *blocks, const double *xr_dev, const double *xs_dev, const double *yr_dev,
                                  const double *ys_dev, double *J_dev, double *sJ_dev, double *nx_dev, double *ny_dev);
int  hsbp_blocks_set_synthetic_warp(hsbp_blocks *blocks, int64_t nbx, int64_t bx0, double L, double A, double *x_dev, double *y_dev);
/* LFToB of every block: 4 * nblocks codes (argument LFToB of locoperator, :212) */
int  hsbp_blocks_set_bc(hsbp_blocks *blocks, const int64_t *bctype);
/* penalty tau_1..4 on the device as global_curved.jl:418-437 (psi_min, l nearest lines) */
int  hsbp_blocks_compute_tau(hsbp_blocks *blocks, double tauscale);
/* or supply / read back tau (face layout, see top) */
int  hsbp_blocks_set_tau(hsbp_blocks *blocks, const double *tau);
int  hsbp_blocks_get_tau(hsbp_blocks *blocks, double *tau);

/* y = M-tilde u for all blocks (the SpMV `lop[e].M̃ * u`, global_curved.jl:470-492)      */
int  hsbp_apply(hsbp_blocks *blocks, const double *u_dev, double *y_dev);
/* y = M-tilde u and energy[e] = u_e . (M-tilde u)_e per block (host array of nblocks doubles) from the same pass: the sweep
 * kernel leaves the chunk sums, a small kernel adds the closure lines -- the p . A p of a CG step on the reference's
 * `lop[e].M̃` without a second pass over the vectors.  Line-marching kernel only (HSBP_ERR_UNSUPP otherwise).           */
int  hsbp_apply_energy(hsbp_blocks *blocks, const double *u_dev, double *y_dev, double *energy);
/* same through host buffers: H2D of u, apply, D2H of y inside the call                 */
int  hsbp_apply_host(hsbp_blocks *blocks, const double *u, double *y);
/* hsbp_apply with CUDA events between its stages.  ms[0] is always the dominant volume kernel:
 *   line-marching variant: ms[0] k_sweep (volume + folded face terms), ms[1] k_face_prep, ms[2] 0
 *   generic variant:       ms[0] two-pass volume kernels, ms[1] face gather, ms[2] face scatter
 * Synchronises; for benchmarking.                                                                   */
int  hsbp_apply_timed(hsbp_blocks *blocks, const double *u_dev, double *y_dev, double *ms);
/* which kernel variant hsbp_apply last used: 0 generic two-pass kernels, 1 line-marching TMA kernel (k_sweep)     */
int  hsbp_apply_variant(const hsbp_blocks *blocks);
/* force the generic kernels (testing) */
int  hsbp_blocks_force_generic(hsbp_blocks *blocks, int on);
/* tuning / testing knobs: "force_generic" (0/1), "sweep_chunks_per_side" (0 = heuristic),
 * "sweep_points_per_thread" (0 = heuristic, 2, 4), "sweep_fold_faces" (1), "sweep_deep" (1: css / crs windows in
 * shared-memory rings, 0: in registers), "sweep_p6_regs" (128 / 168), "fdm_gemm" (arithmetic of the fast-diagonalisation
 * preconditioner's GEMMs, hand-written kernels: 0 fp64 on the fp64 tensor pipe (mma.sync f64), 3 TF32 on tcgen05 with TMEM
 * accumulators (blocks of 128 / 256 points per direction, fp64 otherwise); -1 cuBLAS TF32, for comparison in tests only;
 * set before hsbp_local_setup), "fdm_tc_variant" (0: two fused GEMM pairs with TMA operands and a TMEM operand, 1: four
 * single-GEMM launches; testing), "fdm_eig_lib" (1: eigen-decompositions of the setup by cuSOLVER instead of the batched
 * Jacobi kernel; comparison in tests only), "fdm_no_skip" (1: converged blocks stay in the kernels of the PCG iteration),
 * "fdm_no_fused_dot" (1: p . A p by a separate pass instead of inside the sweep kernel),
 * "band_no_stream" (1: plain-load banded solve kernel) */
int  hsbp_blocks_set_option(hsbp_blocks *blocks, const char *name, int64_t value);

/* face operators of the blocks, block-face layout (no inter-block coupling):
 *   ft = F_k^T u                 (rows of Fbar^T before orientation, global_curved.jl:455-458)
 *   y += F_k v                   (columns; used by locbcarray!, global_curved.jl:596-623)
 *   tr = HfI_FT_k u              (traction operator, global_curved.jl:460-463)               */
int  hsbp_face_FT(hsbp_blocks *blocks, const double *u_dev, double *ft_dev);
int  hsbp_face_F_add(hsbp_blocks *blocks, const double *v_dev, double alpha, double *y_dev);
int  hsbp_face_traction(hsbp_blocks *blocks, const double *u_dev, double *tr_dev);

/* ---- per-block local solves: replaces the `factorization` plugin ---------------------------
 * reference: SBPLocalOperator1 stores factors[e] = factorization(lop[e].M̃) (global_curved.jl:672-703,
 * plugin supplied as x -> cholesky(Symmetric(x)) at square_circle.jl:299, BP1.jl:78) and every use is
 * `F \ g` (global_curved.jl:734, square_circle.jl:383, odefun.jl:43).
 *   HSBP_LOCAL_PCG       batched matrix-free Jacobi-PCG on the block operator (any block size)
 *   HSBP_LOCAL_CHOLESKY  batched dense fp64 Cholesky of M-tilde_e (small blocks)
 *   HSBP_LOCAL_BAND      batched banded fp64 Cholesky of M-tilde_e (points numbered r-fastest, half-bandwidth
 *                        about WB*(Nr+1), WB = 2 / 5 / 8 for p = 2 / 4 / 6): the direct solver for blocks whose band
 *                        fits in device memory, e.g. the single 201 x 201 block of seas/BP1 (BP1.jl:78)
 *   HSBP_LOCAL_FDM       batched matrix-free PCG preconditioned by the exact inverse of the separable part of
 *                        M-tilde_e (fast diagonalisation: four dense fp64 GEMMs per block and iteration); blocks of one
 *                        size, meant for large, mildly varying blocks (256 x 256 points) where Jacobi-PCG needs O(N)
 *                        iterations
 * tol is the relative residual ||g - M u|| / ||g|| per block (PCG variants only).                         */
#define HSBP_LOCAL_PCG      1
#define HSBP_LOCAL_CHOLESKY 2
#define HSBP_LOCAL_BAND     3
#define HSBP_LOCAL_FDM      4
typedef struct {
  int64_t iterations_max;     /* PCG: largest iteration count over blocks (0 for Cholesky) */
  int64_t iterations_sum;     /* PCG: sum over blocks                                      */
  int64_t failed_blocks;      /* blocks that did not reach tol within maxit                */
  double  max_rel_residual;   /* PCG: max over blocks of ||r|| / ||g||                     */
} hsbp_local_stats;
int  hsbp_local_setup(hsbp_blocks *blocks, int mode, double tol, int64_t maxit);
int  hsbp_local_solve(hsbp_blocks *blocks, const double *g_dev, double *u_dev, hsbp_local_stats *stats);
/* z = P^-1 r of the fast-diagonalisation preconditioner alone (HSBP_LOCAL_FDM; testing / profiling hook) */
int  hsbp_local_precondition(hsbp_blocks *blocks, const double *r_dev, double *z_dev);

/* ---- the `factorization` plugin at the reference's own seam -------------------------------------------------------
 * reference: SBPLocalOperator1 calls `factorization(lop[e].M̃)` on the ASSEMBLED sparse matrix of every block, after
 * probing the plugin with a 1 x 1 matrix to learn the factor type (global_curved.jl:681, 698; plugin supplied as
 * x -> cholesky(Symmetric(x)) at square_circle.jl:299, BP1.jl:78); later it only uses `F \ g` (:734,
 * square_circle.jl:383, odefun.jl:43) and `F' \ S` (:774).  hsbp_factor is that object for a host that keeps the
 * reference's assembly: one symmetric positive definite matrix given by its CSC arrays (index_base 1: straight from
 * the reference's host language; only the lower triangle is read), stored as a band -- with points numbered r-fastest
 * M̃_e is banded -- and factorised on the device with the banded Cholesky kernels (DMMA trailing update).  Any size from
 * the 1 x 1 probe upwards.  hsbp_factor_solve: nrhs right-hand sides one after the other, host arrays.               */
typedef struct hsbp_factor hsbp_factor;
int  hsbp_factor_create(hsbp_ctx *ctx, int64_t n, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                        int index_base, hsbp_factor **factor);
int  hsbp_factor_destroy(hsbp_factor *factor);
int64_t hsbp_factor_size(const hsbp_factor *factor);
int  hsbp_factor_solve(hsbp_factor *factor, const double *g, double *u, int64_t nrhs);
int  hsbp_factor_solve_dev(hsbp_factor *factor, const double *g_dev, double *u_dev);

/* ---- trace (lambda) operators and Schur-complement solve ------------------------------------
 * reference: gloλoperator (global_curved.jl:510-565) builds FToλstarts, the sparse Fbar^T and the
 * diagonal D; assembleλmatrix (:743-797) forms B = D - Fbar^T M̃^-1 Fbar explicitly and
 * square_circle.jl:314,377 factorises and solves it.  Here Fbar / Fbar^T / B are applied matrix-free
 * and B lambda = b is solved by a GPU-resident preconditioned CG.
 * Connectivity arrays are exactly connectivityarrays' outputs (global_curved.jl:82-132), 1-based:
 * FToE, FToLF are 2 x nfaces (column-major), EToO (uint8 / Bool) and EToS are 4 x nblocks.       */
int  hsbp_trace_create(hsbp_blocks *blocks, int64_t nfaces, const int64_t *FToB, const int64_t *FToE,
                       const int64_t *FToLF, const uint8_t *EToO, const int64_t *EToS, hsbp_trace **trace);
int  hsbp_trace_destroy(hsbp_trace *trace);     /* before hsbp_blocks_destroy of its blocks: a trace points into them */
int64_t hsbp_trace_num_lambda(const hsbp_trace *trace);                       /* lambda-Np */
int  hsbp_trace_get_starts(const hsbp_trace *trace, int64_t *FTolambdastarts);   /* nfaces+1, 1-based */
int  hsbp_trace_get_D(hsbp_trace *trace, double *D);                          /* host, lambda-Np */
/* Static condensation: form the dense per-block matrices S_e = F_e^T M̃_e^-1 F_e (F_e = [F_1 .. F_4] of block e, size
 * 2(Nr+1) + 2(Ns+1)) once, with one batched local solve per face point -- the products assembleλmatrix computes block by
 * block (global_curved.jl:759-790, `F' \ F` slices) without assembling the global sparse B.  Afterwards every
 * hsbp_trace_schur_apply / CG iteration of hsbp_trace_solve is one dense matrix-vector product per block (scatter and
 * gather fused) instead of a local solve; the right-hand side and the back-substitution still use the local solver.
 * enable = 0 frees the matrices and returns to the matrix-free path.  Call after hsbp_local_setup.  Fails (and keeps
 * nothing) if a local solve did not reach its tolerance.                                                          */
int  hsbp_trace_condense(hsbp_trace *trace, int enable);

/* ---- multi-GPU: one context (= one device, one NCCL rank) per GPU ---------------------------------------------
 * The reference is single-process, single-thread.  What shards is its block structure: blocks are
 * independent given lambda (global_curved.jl:732-737) and every face couples exactly two blocks (:525-554).  The host
 * partitions the blocks, creates one hsbp_blocks / hsbp_trace per device from the *local* connectivity (FToE = 0 for the
 * side of an interface face that lives on another device: such a cut face carries lambda on both devices) and tells
 * the library which faces are cut.  From then on every trace call below is collective over the communicator: the
 * partial Fbar^T contributions of the cut faces travel by ncclSend / ncclRecv (one message per partner), CG scalars
 * and the coarse-level data by ncclAllReduce, all on the context's stream.
 *   hsbp_comm_unique_id   128 bytes (an ncclUniqueId) made by one rank; the host hands them to the others
 *   hsbp_comm_init        every rank, same id
 * hsbp_trace_set_partition (every rank, right after hsbp_trace_create, also ranks without cut faces):
 *   faces[c]    1-based id (position in this trace's FToB) of cut face c
 *   partner[c]  rank that holds the other side
 *   gamma[c]    0-based index of the face among ALL cut faces of the mesh (any numbering both sides agree on, e.g. by
 *               increasing global face id); n_gamma_total = number of cut faces of the whole mesh
 * It completes D = Hf (tau- + tau+) on the cut faces with the partner's half.                                      */
int  hsbp_comm_unique_id(void *id128);
int  hsbp_comm_init(hsbp_ctx *ctx, const void *id128, int rank, int world);
int  hsbp_comm_destroy(hsbp_ctx *ctx);
int  hsbp_comm_rank(const hsbp_ctx *ctx);
int  hsbp_comm_world(const hsbp_ctx *ctx);
int  hsbp_comm_allreduce_sum(hsbp_ctx *ctx, double *x_dev, int64_t n);        /* in place, blocking */
int  hsbp_trace_set_partition(hsbp_trace *trace, int64_t ncut, const int64_t *faces, const int64_t *partner,
                              const int64_t *gamma, int64_t n_gamma_total);

/* Preconditioner of the CG on B (the reference factorises B directly, square_circle.jl:314; here B is only applied).
 * First level:
 *   HSBP_PRECOND_JACOBI       D = Hf (tau- + tau+)                      (default)
 *   HSBP_PRECOND_FACE_BLOCKS  block-Jacobi with the exact diagonal blocks B_ff = D_f - S_e-[f,f] - S_e+[f,f], kept as
 *                             explicit inverses (one dense matrix-vector product per face); needs hsbp_trace_condense.
 *                             On a cut face the partner's S_e[f,f] is fetched by send / recv.
 * Second level (hsbp_trace_coarse_setup, modes = 1..3, 0 = off): additive coarse space of `modes` Legendre polynomials
 * per face.  With two modes the CG iteration count no longer grows with the number of blocks across the mesh.  The
 * coarse matrix Z^T B Z is eliminated rank by rank (dense inverse of the rank-interior part, a small replicated Schur
 * complement on the cut faces), so its cost per iteration does not grow with the number of GPUs.  Call after the
 * first level; collective.
 * hsbp_trace_precond_apply: z = P^-1 r with both levels (collective).                                                */
#define HSBP_PRECOND_JACOBI      0
#define HSBP_PRECOND_FACE_BLOCKS 1
int  hsbp_trace_precond_setup(hsbp_trace *trace, int kind);
int  hsbp_trace_coarse_setup(hsbp_trace *trace, int modes);
int64_t hsbp_trace_coarse_size(const hsbp_trace *trace);      /* coarse dofs: this rank's interior ones + all cut-face ones */
int  hsbp_trace_precond_apply(hsbp_trace *trace, const double *r_dev, double *z_dev);
/* "cg_chunk" (iterations enqueued between two looks at the device's status word, default 4),
 * "cg_lookahead" (chunks the host runs ahead of the device, default 1), "cg_graph" (1: with condensed blocks a chunk of
 * iterations is captured into a CUDA graph once and replayed; 0: plain launches), "cg_p2p" (1, default: the exchanges of the
 * iteration loop -- cut-face parts, partial sums of the CG scalars -- go straight into the partners' device memory over
 * NVLink from small kernels of the same graph; 0, or peer mapping not possible: ncclSend / ncclRecv / ncclAllReduce) */
int  hsbp_trace_set_option(hsbp_trace *trace, const char *name, int64_t value);
/* which way the last hsbp_trace_solve exchanged data inside its loop: 0 one rank, 1 NCCL, 2 peer memory */
int  hsbp_trace_comm_path(const hsbp_trace *trace);
int  hsbp_trace_FbarT(hsbp_trace *trace, const double *u_dev, double *lam_dev);             /* lam = Fbar^T u (this device's blocks) */
int  hsbp_trace_Fbar_add(hsbp_trace *trace, const double *lam_dev, double alpha, double *y_dev); /* y += a Fbar lam */
int  hsbp_trace_schur_apply(hsbp_trace *trace, const double *lam_dev, double *out_dev);     /* out = B lam (collective) */
/* b = gdelta - Fbar^T M̃^-1 g   (LocalToGLobalRHS!, global_curved.jl:730-740); collective             */
int  hsbp_trace_rhs(hsbp_trace *trace, const double *g_dev, const double *gdelta_dev, double *b_dev);
/* statistics of the local solves of the last trace call (rhs, schur_apply, condense, coarse_setup, solve) */
int  hsbp_trace_last_local_stats(hsbp_trace *trace, hsbp_local_stats *stats);
typedef struct {
  int64_t outer_iterations;
  int64_t converged;              /* 1 if the CG residual reached tol * ||b|| and every local solve reached its tolerance */
  double  rel_residual;           /* ||r|| / ||b|| of the CG recurrence at exit                                            */
  int64_t inner_iterations_sum;   /* over all local solves of the call (PCG)          */
  int64_t inner_iterations_max;
  int64_t local_solves;
  double  true_rel_residual;      /* ||b - B lambda|| / ||b|| recomputed with one more application of B after the loop   */
  int64_t failed_local_blocks;    /* block solves (rhs, matvecs, back-substitution) that missed their tolerance         */
  double  max_local_rel_residual;
  int64_t coarse_dofs;            /* size of the second-level problem seen by this rank, 0 = one level                   */
  int64_t issued_iterations;      /* iterations enqueued; those after convergence return at once on the device           */
  double  b_norm;                 /* ||b||                                                                               */
  double  cg_loop_ms;             /* device time of the iteration loop alone (CUDA events on the library's stream)       */
} hsbp_trace_stats;
/* lambda = B^-1 (gdelta - Fbar^T M̃^-1 g), u = M̃^-1 (g - Fbar lambda)   (square_circle.jl:376-388).
 * Device-resident preconditioned CG: all scalars stay on the device, the host enqueues iterations ahead and looks at a
 * mapped status word every "cg_chunk" iterations; two reductions per iteration on a partitioned mesh (p.q, and one
 * vector with r.z, r.r and the coarse-level data).  Collective.                                                     */
int  hsbp_trace_solve(hsbp_trace *trace, const double *g_dev, const double *gdelta_dev,
                      double *lambda_dev, double *u_dev, double tol, int64_t maxit, hsbp_trace_stats *stats);

/* ---- SEAS BP1: ODE right-hand side with rate-and-state friction ---------------------------------
 * reference: odefun (seas/BP1/odefun.jl:8-121) = boundary data (:36-42, locbcarray_mod! global_curved.jl:569-592)
 * -> local solve (:43) -> shear traction on the fault (computetraction_mod, global_curved.jl:627-634) -> per fault
 * node bracketed Newton on rateandstate (newtbndv, global_curved.jl:1031-1075, odefun.jl:69-96) -> d(psi)/dt (:101).
 * The state vector is [psi; delta], the output [dpsi/dt; V], each 2 * (points of the fault face).
 * Failures are not errors: stats.rejected != 0 is the reference's reject_step flag (BP1.jl:149-159).            */
typedef struct hsbp_bp1 hsbp_bp1;
typedef struct {
  double Vp;                      /* plate rate; loading face carries t * Vp / 2            (odefun.jl:36)   */
  double mu_shear;                /* shear modulus                                            (BP1.jl:24)      */
  double sigma_n, eta, V0, tau_z0, Dc, f0, b;   /* rate-and-state parameters                 (BP1.jl:8-23,104) */
  double ftol, atolx, rtolx;      /* newtbndv tolerances, 1e-9 in odefun.jl:83-85                              */
  int64_t maxiter;                /* 500                                                      (global_curved.jl:1041) */
} hsbp_bp1_params;
typedef struct {
  int64_t rejected;               /* any failure below, or a local solve that did not converge                  */
  int64_t failure_bits;           /* 1: tau NaN (odefun.jl:73), 2: root find failed (:91), 4: dpsi not finite (:102) */
  int64_t failed_nodes;
  int64_t newton_iterations_max;
  int64_t local_iterations;       /* PCG iterations of the local solve                                           */
} hsbp_bp1_stats;
/* block, fault_face, loading_face are 1-based; a and sJ live on the fault face (RSa of BP1.jl:96-102, lop.sJ) */
int  hsbp_bp1_create(hsbp_blocks *blocks, int64_t block, int64_t fault_face, int64_t loading_face,
                     const double *a, const double *sJ, const hsbp_bp1_params *params, hsbp_bp1 **bp1);
int  hsbp_bp1_destroy(hsbp_bp1 *bp1);
/* Condense the local solve onto the fault (optional, after hsbp_local_setup): the displacement enters odefun only through
 * the traction on the fault face and u = M̃^-1 ge is linear in the boundary data (odefun.jl:36-43), so
 *   HfI_FT_f u = -1/2 Tf delta - (t Vp / 2) tl,   Tf = HfI_FT_f M̃^-1 F_f,  tl = HfI_FT_f M̃^-1 F_l 1
 * -- the same per-block products assembleλmatrix forms (global_curved.jl:759-790), for the two Dirichlet faces of the BP1
 * block.  Costs (fault points + 1) local solves once; afterwards hsbp_bp1_rhs is one small kernel (dense matrix-vector
 * product + root find per node) and hsbp_bp1_get_u does the local solve on demand.  enable = 0 switches back.         */
int  hsbp_bp1_condense(hsbp_bp1 *bp1, int enable);
int  hsbp_bp1_rhs(hsbp_bp1 *bp1, double t, const double *psi_delta, double *dpsi_V, hsbp_bp1_stats *stats);
int  hsbp_bp1_get_u(hsbp_bp1 *bp1, double *u);      /* displacement of the last rhs call, host array of VNp */

/* ---- rate-and-state fault stage on a multiblock mesh ---------------------------------------
 * On a multiblock mesh (seas/BP1/meshes/BP1_v1.inp) the fault is a set of jump interfaces: slip enters through the jump
 * branch of locbcarray! (global_curved.jl:614-617), the displacement comes from the trace solve (square_circle.jl:376-388)
 * and the shear stress from computetraction (global_curved.jl:638-644).  The chain is linear in (slip, time): the stress
 * change at the n fault nodes is dtau = A delta + t b.  hsbp_fault_rhs evaluates [dpsi/dt; V] from [psi; delta] with A, b
 * formed once by the host layer (n + 1 trace solves); hsbp_fault_stage takes the dtau of one trace solve per call.  The
 * per-node root find and state evolution are those of hsbp_bp1_rhs (odefun.jl:69-108).  A, b may be NULL (stage only). */
typedef struct hsbp_fault hsbp_fault;
int  hsbp_fault_create(hsbp_ctx *ctx, int64_t n, const double *A, const double *b, const double *a,
                       const hsbp_bp1_params *params, hsbp_fault **fault);
int  hsbp_fault_destroy(hsbp_fault *fault);
int  hsbp_fault_rhs(hsbp_fault *fault, double t, const double *psi_delta, double *dpsi_V, hsbp_bp1_stats *stats);
int  hsbp_fault_stage(hsbp_fault *fault, const double *dtau_dev, const double *psi_delta, double *dpsi_V, hsbp_bp1_stats *stats);

/* ---- measured fp64 denominators of the device (benchmarks; not on the solve path) -------------------------------
 * sustained fp64 FMA rate of the CUDA cores, sustained mma.sync.m8n8k4.f64 rate (the tensor instruction of the dense and
 * banded factorisations and of the Gauss-Jordan inversions), cuBLAS DGEMM n^3 (library number, denominator only).       */
int  hsbp_peak_fp64_fma(hsbp_ctx *ctx, double *tflops);
int  hsbp_peak_fp64_dmma(hsbp_ctx *ctx, double *tflops);
int  hsbp_peak_dgemm(hsbp_ctx *ctx, int64_t n, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* HSBP_H */
