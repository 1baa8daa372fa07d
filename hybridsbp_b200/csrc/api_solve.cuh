// C-ABI: local solves (K2), trace operators (K3) and the Schur-complement CG (K4).
// Included by hsbp.cu (unity build).
#pragma once
#include "k_solve.cuh"

struct hsbp_trace {
  hsbp_blocks *blocks = nullptr;
  int64_t nfaces = 0, nlam_faces = 0, lNp = 0;
  std::vector<int64_t> starts;          // nfaces + 1, 1-based (FToλstarts)
  std::vector<hsbp::LamFace> h_faces;
  hsbp::LamFace *d_faces = nullptr;
  double *d_D = nullptr;
  double *d_ft = nullptr, *d_fv = nullptr;          // block-face scratch (FNp)
  double *d_w = nullptr, *d_z = nullptr;            // volume scratch (VNp)
  double *d_r = nullptr, *d_p = nullptr, *d_q = nullptr, *d_zz = nullptr;   // lambda scratch
  double *d_partial = nullptr, *d_dots = nullptr;
  hsbp_local_stats acc = {0, 0, 0, 0.0};
  int64_t local_solves = 0;
  // static condensation (hsbp_trace_condense): dense S_e = F_e^T M̃_e^-1 F_e of every block
  double *d_S = nullptr;
  int64_t *d_S_off = nullptr;
  int max_nf = 0;
  // face-block preconditioner (hsbp_trace_precond_setup): Cholesky factors of the diagonal blocks B_ff
  double *d_pc = nullptr, *d_pc_work = nullptr;
  void *d_pc_desc = nullptr;
};

namespace {

using namespace hsbp;

int fdm_setup(hsbp_blocks *b);                                                        // api_fdm.cuh
int fdm_solve(hsbp_blocks *b, const double *g, double *x, hsbp_local_stats *stats);

inline dim3 vec_grid(int64_t n) {
  return dim3((unsigned)std::max<int64_t>(1, std::min<int64_t>((n + VEC_THREADS - 1) / VEC_THREADS, 148 * 8)));
}

int local_alloc(hsbp_blocks *b) {
  hsbp_ctx *ctx = b->ctx;
  const size_t vb = (size_t)b->VNp * sizeof(double);
  if (!b->d_dinv) {
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_dinv, vb));
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_pr, vb));
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_pp, vb));
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_pAp, vb));
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_pcg, b->nblocks * sizeof(PcgState)));
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_nactive, 2 * sizeof(int)));
  }
  return HSBP_OK;
}

// Jacobi preconditioner: exact diagonal of M-tilde by probing with c x c coloured unit vectors
template <int P> int probe_diagonal(hsbp_blocks *b) {
  hsbp_ctx *ctx = b->ctx;
  // farthest coupling of M-tilde in one direction: closure block (M-1), Neumann G^T G (2W), cross terms
  const int c = std::max(Sbp<P>::M - 1, 2 * Sbp<P>::W) + 1;
  const dim3 grid = gen_grid(b);
  double *u = b->d_pp, *y = b->d_pAp;
  for (int cj = 0; cj < c; ++cj)
    for (int ci = 0; ci < c; ++ci) {
      k_color_vector<P><<<grid, GEN_THREADS, 0, ctx->stream>>>(b->d_desc, c, ci, cj, u);
      int rc = apply_async(b, u, y);
      if (rc) return rc;
      k_color_pick<P><<<grid, GEN_THREADS, 0, ctx->stream>>>(b->d_desc, c, ci, cj, y, b->d_dinv);
    }
  return check_launch(ctx, "probe_diagonal");
}

int pcg_solve(hsbp_blocks *b, const double *g, double *x, hsbp_local_stats *stats) {
  hsbp_ctx *ctx = b->ctx;
  const double tol2 = b->local_tol * b->local_tol;
  k_pcg_init<<<(unsigned)b->nblocks, 1024, 0, ctx->stream>>>(b->d_desc, g, b->d_dinv, x, b->d_pr, b->d_pp, (PcgState *)b->d_pcg, tol2);
  int rc = check_launch(ctx, "k_pcg_init");
  if (rc) return rc;
  const int check_every = 8;
  int h_active = 1;
  int64_t it = 0;
  while (it < b->local_maxit) {
    int slot = 0;
    for (int k = 0; k < check_every && it < b->local_maxit; ++k, ++it) {
      if ((rc = apply_async(b, b->d_pp, b->d_pAp))) return rc;
      slot = (int)(it & 1);
      HSBP_CUDA(ctx, cudaMemsetAsync(b->d_nactive + slot, 0, sizeof(int), ctx->stream));
      k_pcg_update<<<(unsigned)b->nblocks, 1024, 0, ctx->stream>>>(b->d_desc, b->d_dinv, b->d_pAp, x, b->d_pr, b->d_pp,
                                                                  (PcgState *)b->d_pcg, tol2, b->d_nactive + slot);
    }
    if ((rc = check_launch(ctx, "k_pcg_update"))) return rc;
    HSBP_CUDA(ctx, cudaMemcpyAsync(&h_active, b->d_nactive + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_active == 0) break;
  }
  if (stats) {
    std::vector<PcgState> st(b->nblocks);
    HSBP_CUDA(ctx, cudaMemcpyAsync(st.data(), b->d_pcg, b->nblocks * sizeof(PcgState), cudaMemcpyDeviceToHost, ctx->stream));
    HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    hsbp_local_stats s = {0, 0, 0, 0.0};
    for (auto &q : st) {
      s.iterations_max = std::max<int64_t>(s.iterations_max, q.iters);
      s.iterations_sum += q.iters;
      s.failed_blocks += q.active ? 1 : 0;
      if (q.g2 > 0) s.max_rel_residual = std::max(s.max_rel_residual, sqrt(q.rr / q.g2));
    }
    *stats = s;
  }
  return HSBP_OK;
}

int local_solve_impl(hsbp_blocks *b, const double *g, double *u, hsbp_local_stats *stats) {
  hsbp_ctx *ctx = b->ctx;
  if (b->local_mode == 0) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_local_solve: call hsbp_local_setup first");
  if (!g || !u || g == u) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_local_solve: bad pointers (in-place solve is not supported)");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  if (b->local_mode == HSBP_LOCAL_CHOLESKY) return chol_solve(b, g, u, stats);
  if (b->local_mode == HSBP_LOCAL_BAND) return band_solve(b, g, u, stats);
  if (b->local_mode == HSBP_LOCAL_FDM) return fdm_solve(b, g, u, stats);
  return pcg_solve(b, g, u, stats);
}

int dots(hsbp_trace *t, int64_t n, const double *x0, const double *y0, const double *x1, const double *y1,
         const double *x2, const double *y2, double out[3]) {
  hsbp_ctx *ctx = t->blocks->ctx;
  k_dot3_partial<<<DOT_BLOCKS, VEC_THREADS, 0, ctx->stream>>>(n, x0, y0, x1, y1, x2, y2, t->d_partial);
  k_dot3_final<<<1, 256, 0, ctx->stream>>>(DOT_BLOCKS, t->d_partial, t->d_dots);
  int rc = check_launch(ctx, "k_dot3");
  if (rc) return rc;
  HSBP_CUDA(ctx, cudaMemcpyAsync(out, t->d_dots, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HSBP_OK;
}

int trace_FbarT(hsbp_trace *t, const double *u, double *lam) {
  int rc = hsbp_face_FT(t->blocks, u, t->d_ft);
  if (rc) return rc;
  hsbp_ctx *ctx = t->blocks->ctx;
  if (t->nlam_faces)
    k_lam_gather<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_faces, t->d_ft, lam);
  return check_launch(ctx, "k_lam_gather");
}

int trace_Fbar_add(hsbp_trace *t, const double *lam, double alpha, double *y) {
  hsbp_ctx *ctx = t->blocks->ctx;
  HSBP_CUDA(ctx, cudaMemsetAsync(t->d_fv, 0, (size_t)t->blocks->FNp * sizeof(double), ctx->stream));
  if (t->nlam_faces)
    k_lam_scatter<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_faces, lam, t->d_fv);
  int rc = check_launch(ctx, "k_lam_scatter");
  if (rc) return rc;
  return hsbp_face_F_add(t->blocks, t->d_fv, alpha, y);
}

void accumulate(hsbp_trace *t, const hsbp_local_stats &s) {
  t->acc.iterations_max = std::max(t->acc.iterations_max, s.iterations_max);
  t->acc.iterations_sum += s.iterations_sum;
  t->acc.failed_blocks += s.failed_blocks;
  t->acc.max_rel_residual = std::max(t->acc.max_rel_residual, s.max_rel_residual);
  t->local_solves += 1;
}

// out = D o lam - Fbar^T M^-1 Fbar lam
int schur_apply(hsbp_trace *t, const double *lam, double *out) {
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  if (t->d_S) {                      // condensed: scatter, one dense matrix-vector product per block, gather
    HSBP_CUDA(ctx, cudaMemsetAsync(t->d_fv, 0, (size_t)b->FNp * sizeof(double), ctx->stream));
    if (t->nlam_faces) k_lam_scatter<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_faces, lam, t->d_fv);
    k_cond_gemv<<<dim3(16, (unsigned)b->nblocks), 256, (size_t)t->max_nf * sizeof(double), ctx->stream>>>(
        b->d_desc, t->d_S_off, t->d_S, t->d_fv, t->d_ft);
    if (t->nlam_faces) k_lam_gather<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_faces, t->d_ft, out);
    k_ewise<<<vec_grid(t->lNp), VEC_THREADS, 0, ctx->stream>>>(t->lNp, t->d_D, lam, t->d_zz, 0);
    k_axpby<<<vec_grid(t->lNp), VEC_THREADS, 0, ctx->stream>>>(t->lNp, 1.0, t->d_zz, -1.0, out, out);
    return check_launch(ctx, "schur_apply (condensed)");
  }
  HSBP_CUDA(ctx, cudaMemsetAsync(t->d_w, 0, (size_t)b->VNp * sizeof(double), ctx->stream));
  int rc = trace_Fbar_add(t, lam, 1.0, t->d_w);
  if (rc) return rc;
  hsbp_local_stats s;
  if ((rc = local_solve_impl(b, t->d_w, t->d_z, &s))) return rc;
  accumulate(t, s);
  if ((rc = trace_FbarT(t, t->d_z, out))) return rc;
  // out = D*lam - out
  k_ewise<<<vec_grid(t->lNp), VEC_THREADS, 0, ctx->stream>>>(t->lNp, t->d_D, lam, t->d_zz, 0);
  k_axpby<<<vec_grid(t->lNp), VEC_THREADS, 0, ctx->stream>>>(t->lNp, 1.0, t->d_zz, -1.0, out, out);
  return check_launch(ctx, "schur_apply");
}

// S_e = F_e^T M̃_e^-1 F_e column by column, all blocks in lockstep: one local solve per face point of a block
int trace_condense(hsbp_trace *t) {
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  if (b->local_mode == 0) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_condense: call hsbp_local_setup first");
  std::vector<int64_t> off(b->nblocks);
  int64_t total = 0;
  int max_nf = 0;
  for (int64_t e = 0; e < b->nblocks; ++e) {
    const BlockDesc &d = b->h_desc[e];
    const int nf = 2 * (d.Ns + 1) + 2 * (d.Nr + 1);
    off[e] = total; total += (int64_t)nf * nf; max_nf = std::max(max_nf, nf);
  }
  size_t free_b = 0, total_b = 0;
  HSBP_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  if ((size_t)total * sizeof(double) > free_b / 2 || (size_t)max_nf * sizeof(double) > 48 * 1024)
    HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "hsbp_trace_condense: the condensed blocks do not fit");
  cudaFree(t->d_S); cudaFree(t->d_S_off); t->d_S = nullptr; t->d_S_off = nullptr;
  double *S = nullptr;
  int64_t *S_off = nullptr;
  HSBP_CUDA(ctx, cudaMalloc((void **)&S, (size_t)total * sizeof(double)));
  if (cudaMalloc((void **)&S_off, b->nblocks * sizeof(int64_t)) != cudaSuccess) { cudaFree(S); HSBP_FAIL(ctx, HSBP_ERR_CUDA, "out of device memory"); }
  auto fail = [&](int rc) { cudaFree(S); cudaFree(S_off); return rc; };
  if (cudaMemcpyAsync(S_off, off.data(), b->nblocks * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
      cudaMemsetAsync(t->d_fv, 0, (size_t)b->FNp * sizeof(double), ctx->stream) != cudaSuccess)
    return fail(HSBP_ERR_CUDA);
  int rc;
  for (int c = 0; c < max_nf; ++c) {
    k_cond_unit<<<(unsigned)((b->nblocks + 127) / 128), 128, 0, ctx->stream>>>(b->d_desc, b->nblocks, c, t->d_fv);
    if (cudaMemsetAsync(t->d_w, 0, (size_t)b->VNp * sizeof(double), ctx->stream) != cudaSuccess) return fail(HSBP_ERR_CUDA);
    if ((rc = hsbp_face_F_add(b, t->d_fv, 1.0, t->d_w))) return fail(rc);
    hsbp_local_stats s;
    if ((rc = local_solve_impl(b, t->d_w, t->d_z, &s))) return fail(rc);
    accumulate(t, s);
    if ((rc = hsbp_face_FT(b, t->d_z, t->d_ft))) return fail(rc);
    k_cond_store<<<(unsigned)b->nblocks, 256, 0, ctx->stream>>>(b->d_desc, S_off, c, t->d_ft, S);
  }
  k_cond_sym<<<dim3((unsigned)b->nblocks, 16), 256, 0, ctx->stream>>>(b->d_desc, S_off, S);
  if ((rc = check_launch(ctx, "trace_condense"))) return fail(rc);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(HSBP_ERR_CUDA);
  t->d_S = S; t->d_S_off = S_off; t->max_nf = max_nf;
  return HSBP_OK;
}

void precond_free(hsbp_trace *t) {
  cudaFree(t->d_pc); cudaFree(t->d_pc_work); cudaFree(t->d_pc_desc);
  t->d_pc = nullptr; t->d_pc_work = nullptr; t->d_pc_desc = nullptr;
}

// dense Cholesky of every B_ff (batched, the panel / DMMA trailing-update kernels of the dense local solver)
int precond_faceblocks(hsbp_trace *t, int64_t ncut = 0, const int64_t *cut_faces = nullptr, const double *partner_dev = nullptr) {
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  if (!t->d_S) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_precond_setup: the face-block preconditioner needs hsbp_trace_condense first");
  precond_free(t);
  const int64_t nf = t->nlam_faces;
  if (nf == 0) return HSBP_OK;
  static_assert(sizeof(FaceBlock) == sizeof(CholBlock), "descriptor layout");
  std::vector<FaceBlock> fbs(nf);
  int64_t off = 0, woff = 0;
  int maxld = 0;
  for (int64_t i = 0; i < nf; ++i) {
    const LamFace &f = t->h_faces[i];
    const int ld = (f.nl + CH_NB - 1) / CH_NB * CH_NB;
    fbs[i].off = off; fbs[i].np = f.nl; fbs[i].ld = ld; fbs[i].voff = f.loff; fbs[i].woff = woff;
    off += (int64_t)ld * ld; woff += ld; maxld = std::max(maxld, ld);
  }
  HSBP_CUDA(ctx, cudaMalloc((void **)&t->d_pc, (size_t)off * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&t->d_pc_work, (size_t)woff * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&t->d_pc_desc, nf * sizeof(FaceBlock)));
  HSBP_CUDA(ctx, cudaMemcpyAsync(t->d_pc_desc, fbs.data(), nf * sizeof(FaceBlock), cudaMemcpyHostToDevice, ctx->stream));
  // cut faces whose partner contribution is known: map lambda-face index -> offset into partner_dev
  int64_t *d_pidx = nullptr;
  if (ncut > 0) {
    std::vector<int64_t> pidx(nf, -1);
    int64_t po = 0;
    for (int64_t c = 0; c < ncut; ++c) {
      const int64_t lfi = cut_faces[c];
      if (lfi < 0 || lfi >= nf) { precond_free(t); HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_precond_setup_cut: bad face index"); }
      pidx[lfi] = po;
      po += (int64_t)t->h_faces[lfi].nl * t->h_faces[lfi].nl;
    }
    HSBP_CUDA(ctx, cudaMalloc((void **)&d_pidx, nf * sizeof(int64_t)));
    HSBP_CUDA(ctx, cudaMemcpyAsync(d_pidx, pidx.data(), nf * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  k_faceblock_fill<<<(unsigned)nf, 256, 0, ctx->stream>>>(t->d_faces, (const FaceBlock *)t->d_pc_desc, b->d_desc, t->d_S_off, t->d_S,
                                                         t->d_D, t->d_pc, d_pidx, partner_dev);
  int *d_flag = nullptr;
  std::vector<int> flag(nf, 0);
  HSBP_CUDA(ctx, cudaMalloc((void **)&d_flag, nf * sizeof(int)));
  cudaMemsetAsync(d_flag, 0, nf * sizeof(int), ctx->stream);
  const CholBlock *dcb = (const CholBlock *)t->d_pc_desc;
  for (int k0 = 0; k0 < maxld; k0 += CH_NB) {
    k_chol_panel<<<(unsigned)nf, CH_THREADS, 0, ctx->stream>>>(dcb, t->d_pc, k0, d_flag);
    const int nt = (maxld - k0 - CH_NB) / CH_NB;
    if (nt > 0) k_chol_update<<<dim3(nt, nt, (unsigned)nf), CH_THREADS, 0, ctx->stream>>>(dcb, t->d_pc, k0);
  }
  cudaError_t e1 = cudaGetLastError();
  if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(flag.data(), d_flag, nf * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_flag); cudaFree(d_pidx);
  if (e1 != cudaSuccess) { precond_free(t); ctx->err = std::string("precond_faceblocks: ") + cudaGetErrorString(e1); return HSBP_ERR_CUDA; }
  for (int64_t i = 0; i < nf; ++i)
    if (flag[i]) { precond_free(t); HSBP_FAIL(ctx, HSBP_ERR_ARG, "face-block preconditioner: a diagonal block of B is not positive definite"); }
  return HSBP_OK;
}

// z = preconditioner^-1 r: Cholesky solves with the face blocks, or r / D (Jacobi)
int precond_apply(hsbp_trace *t, const double *r, double *z) {
  hsbp_ctx *ctx = t->blocks->ctx;
  if (t->d_pc && t->nlam_faces)
    k_chol_solve<<<(unsigned)t->nlam_faces, CH_THREADS, 0, ctx->stream>>>((const CholBlock *)t->d_pc_desc, t->d_pc, r, z, t->d_pc_work);
  else
    k_ewise<<<vec_grid(t->lNp), VEC_THREADS, 0, ctx->stream>>>(t->lNp, r, t->d_D, z, 1);
  return check_launch(ctx, "precond_apply");
}

int trace_rhs(hsbp_trace *t, const double *g, const double *gd, double *bl) {
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  hsbp_local_stats s;
  int rc = local_solve_impl(b, g, t->d_z, &s);
  if (rc) return rc;
  accumulate(t, s);
  if ((rc = trace_FbarT(t, t->d_z, bl))) return rc;
  k_axpby<<<vec_grid(t->lNp), VEC_THREADS, 0, ctx->stream>>>(t->lNp, 1.0, gd, -1.0, bl, bl);
  return check_launch(ctx, "trace_rhs");
}

}  // namespace

extern "C" {

int hsbp_local_setup(hsbp_blocks *b, int mode, double tol, int64_t maxit) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!b->have_metrics || !b->have_tau) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_local_setup: metrics / tau not set");
  if (mode != HSBP_LOCAL_PCG && mode != HSBP_LOCAL_CHOLESKY && mode != HSBP_LOCAL_BAND && mode != HSBP_LOCAL_FDM) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_local_setup: unknown mode");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  b->local_tol = tol > 0 ? tol : 1e-13;
  b->local_maxit = maxit > 0 ? maxit : 100000;
  int rc;
  if (mode == HSBP_LOCAL_CHOLESKY) {
    if ((rc = chol_setup(b))) return rc;
  } else if (mode == HSBP_LOCAL_BAND) {
    if ((rc = band_setup(b))) return rc;
  } else if (mode == HSBP_LOCAL_FDM) {
    if ((rc = fdm_setup(b))) return rc;
  } else {
    if ((rc = local_alloc(b))) return rc;
    rc = dispatch_p(b->p, [&](auto Pc) { return probe_diagonal<decltype(Pc)::value>(b); });
    if (rc) return rc;
    HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  b->local_mode = mode;
  return HSBP_OK;
}

int hsbp_local_solve(hsbp_blocks *b, const double *g_dev, double *u_dev, hsbp_local_stats *stats) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_local_stats s;
  int rc = local_solve_impl(b, g_dev, u_dev, &s);
  if (rc == HSBP_OK && stats) *stats = s;
  return rc;
}

int hsbp_trace_create(hsbp_blocks *b, int64_t nfaces, const int64_t *FToB, const int64_t *FToE, const int64_t *FToLF,
                      const uint8_t *EToO, const int64_t *EToS, hsbp_trace **out) {
  if (!b || !out) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  *out = nullptr;
  if (nfaces <= 0 || !FToB || !FToE || !FToLF || !EToO || !EToS) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_create: bad arguments");
  if (!b->have_tau) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_create: tau not set");
  hsbp_trace *t = new (std::nothrow) hsbp_trace();
  if (!t) HSBP_FAIL(ctx, HSBP_ERR_STATE, "out of host memory");
  t->blocks = b; t->nfaces = nfaces;
  t->starts.assign(nfaces + 1, 1);
  auto fail = [&](const char *m) { ctx->err = m; delete t; return HSBP_ERR_ARG; };
  auto fstart = [&](int64_t e, int k) {
    const BlockDesc &d = b->h_desc[e];
    const int64_t nsp = d.Ns + 1, nrp = d.Nr + 1;
    return d.foff + (k == 0 ? 0 : k == 1 ? nsp : k == 2 ? 2 * nsp : 2 * nsp + nrp);
  };
  for (int64_t f = 0; f < nfaces; ++f) {
    const int64_t bc = FToB[f];
    if (bc == HSBP_BC_DIRICHLET || bc == HSBP_BC_NEUMANN) { t->starts[f + 1] = t->starts[f]; continue; }   // :521-524
    if (!(bc == HSBP_BC_LOCKED || bc >= HSBP_BC_JUMP)) return fail("invalid bc");
    // a side whose FToE entry is 0 lives on another device (partitioned mesh): the face still carries lambda here
    const int64_t em = FToE[2 * f] - 1, ep = FToE[2 * f + 1] - 1;
    const int km = (int)FToLF[2 * f] - 1, kp = (int)FToLF[2 * f + 1] - 1;
    if (em < 0 && ep < 0) return fail("hsbp_trace_create: interface face without a local block");
    if (em >= b->nblocks || (em >= 0 && (km < 0 || km > 3))) return fail("hsbp_trace_create: bad FToE / FToLF");
    if (ep >= b->nblocks || (ep >= 0 && (kp < 0 || kp > 3))) return fail("hsbp_trace_create: bad FToE / FToLF");
    int nl = -1;
    if (em >= 0) { const BlockDesc &dm = b->h_desc[em]; nl = (km <= 1 ? dm.Ns : dm.Nr) + 1; }
    if (ep >= 0) {
      const BlockDesc &dp = b->h_desc[ep];
      const int nlp = (kp <= 1 ? dp.Ns : dp.Nr) + 1;
      if (nl >= 0 && nl != nlp) return fail("non-conforming interface (global_curved.jl:528)");
      nl = nlp;
    }
    if (em >= 0 && (!EToO[km + 4 * em] || EToS[km + 4 * em] != 1)) return fail("minus side must be oriented with the face (global_curved.jl:531)");
    if (ep >= 0 && EToS[kp + 4 * ep] != 2) return fail("plus side must have EToS == 2 (global_curved.jl:539)");
    LamFace lf;
    lf.em = (int32_t)em; lf.km = km; lf.ep = (int32_t)ep; lf.kp = kp;
    lf.flip = (ep >= 0 && !EToO[kp + 4 * ep]) ? 1 : 0; lf.nl = nl;
    lf.loff = t->starts[f] - 1; lf.fm = em >= 0 ? fstart(em, km) : 0; lf.fp = ep >= 0 ? fstart(ep, kp) : 0;
    t->h_faces.push_back(lf);
    t->starts[f + 1] = t->starts[f] + nl;
  }
  t->nlam_faces = (int64_t)t->h_faces.size();
  t->lNp = t->starts[nfaces] - 1;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaSuccess;
  auto A = [&](void **p_, size_t n) { if (e == cudaSuccess) e = cudaMalloc(p_, n ? n : 8); };
  const size_t lb = (size_t)t->lNp * sizeof(double), fb = (size_t)b->FNp * sizeof(double), vb = (size_t)b->VNp * sizeof(double);
  A((void **)&t->d_faces, t->h_faces.size() * sizeof(LamFace));
  A((void **)&t->d_D, lb); A((void **)&t->d_ft, fb); A((void **)&t->d_fv, fb);
  A((void **)&t->d_w, vb); A((void **)&t->d_z, vb);
  A((void **)&t->d_r, lb); A((void **)&t->d_p, lb); A((void **)&t->d_q, lb); A((void **)&t->d_zz, lb);
  A((void **)&t->d_partial, 3 * DOT_BLOCKS * sizeof(double)); A((void **)&t->d_dots, 3 * sizeof(double));
  if (e == cudaSuccess && t->nlam_faces)
    e = cudaMemcpyAsync(t->d_faces, t->h_faces.data(), t->h_faces.size() * sizeof(LamFace), cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    ctx->err = std::string("hsbp_trace_create: ") + cudaGetErrorString(e);
    hsbp_trace_destroy(t);
    return HSBP_ERR_CUDA;
  }
  if (t->nlam_faces) {
    int rc = dispatch_p(b->p, [&](auto Pc) {
      k_lam_D<decltype(Pc)::value><<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_faces, b->d_desc, b->d_tau, t->d_D);
      return check_launch(ctx, "k_lam_D");
    });
    if (rc) { hsbp_trace_destroy(t); return rc; }
  }
  cudaStreamSynchronize(ctx->stream);
  *out = t;
  return HSBP_OK;
}

int hsbp_trace_destroy(hsbp_trace *t) {
  if (!t) return HSBP_ERR_ARG;
  cudaSetDevice(t->blocks->ctx->device);
  cudaStreamSynchronize(t->blocks->ctx->stream);
  cudaFree(t->d_faces); cudaFree(t->d_D); cudaFree(t->d_ft); cudaFree(t->d_fv); cudaFree(t->d_w); cudaFree(t->d_z);
  cudaFree(t->d_r); cudaFree(t->d_p); cudaFree(t->d_q); cudaFree(t->d_zz); cudaFree(t->d_partial); cudaFree(t->d_dots);
  cudaFree(t->d_S); cudaFree(t->d_S_off);
  precond_free(t);
  delete t;
  return HSBP_OK;
}

int64_t hsbp_trace_num_lambda(const hsbp_trace *t) { return t ? t->lNp : -1; }

int hsbp_trace_get_starts(const hsbp_trace *t, int64_t *s) {
  if (!t || !s) return HSBP_ERR_ARG;
  memcpy(s, t->starts.data(), t->starts.size() * sizeof(int64_t));
  return HSBP_OK;
}

int hsbp_trace_get_D(hsbp_trace *t, double *D) {
  if (!t || !D) return HSBP_ERR_ARG;
  return hsbp_d2h(t->blocks->ctx, D, t->d_D, (size_t)t->lNp * sizeof(double));
}

int hsbp_trace_set_D(hsbp_trace *t, const double *D) {
  if (!t || !D) return HSBP_ERR_ARG;
  return hsbp_h2d(t->blocks->ctx, t->d_D, D, (size_t)t->lNp * sizeof(double));
}

int hsbp_trace_FbarT(hsbp_trace *t, const double *u_dev, double *lam_dev) {
  if (!t) return HSBP_ERR_ARG;
  if (!u_dev || !lam_dev) HSBP_FAIL(t->blocks->ctx, HSBP_ERR_ARG, "hsbp_trace_FbarT: null pointer");
  return trace_FbarT(t, u_dev, lam_dev);
}

int hsbp_trace_Fbar_add(hsbp_trace *t, const double *lam_dev, double alpha, double *y_dev) {
  if (!t) return HSBP_ERR_ARG;
  if (!y_dev || !lam_dev) HSBP_FAIL(t->blocks->ctx, HSBP_ERR_ARG, "hsbp_trace_Fbar_add: null pointer");
  HSBP_CUDA(t->blocks->ctx, cudaSetDevice(t->blocks->ctx->device));
  return trace_Fbar_add(t, lam_dev, alpha, y_dev);
}

int hsbp_trace_schur_apply(hsbp_trace *t, const double *lam_dev, double *out_dev) {
  if (!t) return HSBP_ERR_ARG;
  if (!out_dev || !lam_dev || out_dev == lam_dev) HSBP_FAIL(t->blocks->ctx, HSBP_ERR_ARG, "hsbp_trace_schur_apply: bad pointers");
  HSBP_CUDA(t->blocks->ctx, cudaSetDevice(t->blocks->ctx->device));
  return schur_apply(t, lam_dev, out_dev);
}

int hsbp_trace_condense(hsbp_trace *t, int enable) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = t->blocks->ctx;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!enable) {
    HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(t->d_S); cudaFree(t->d_S_off); t->d_S = nullptr; t->d_S_off = nullptr;
    return HSBP_OK;
  }
  precond_free(t);                   // factors of an older S
  return trace_condense(t);
}

int hsbp_trace_precond_setup(hsbp_trace *t, int kind) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = t->blocks->ctx;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  if (kind == HSBP_PRECOND_JACOBI) {
    HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    precond_free(t);
    return HSBP_OK;
  }
  if (kind != HSBP_PRECOND_FACE_BLOCKS) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_precond_setup: unknown kind");
  return precond_faceblocks(t);
}

// lambda-face index (position among the faces that carry lambda) of face f (0-based position in the FToB order), or -1
static int64_t lam_face_index(const hsbp_trace *t, int64_t f) {
  if (f < 0 || f >= t->nfaces || t->starts[f + 1] == t->starts[f]) return -1;
  int64_t k = 0;
  for (int64_t g = 0; g < f; ++g) k += t->starts[g + 1] > t->starts[g] ? 1 : 0;
  return k;
}

int hsbp_trace_precond_cut_own(hsbp_trace *t, int64_t ncut, const int64_t *faces, double *out_dev) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  if (ncut < 0 || (ncut > 0 && (!faces || !out_dev))) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_precond_cut_own: bad arguments");
  if (!t->d_S) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_precond_cut_own: needs hsbp_trace_condense first");
  if (ncut == 0) return HSBP_OK;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<int64_t> idx(ncut), ooff(ncut);
  int64_t o = 0;
  for (int64_t c = 0; c < ncut; ++c) {
    const int64_t k = lam_face_index(t, faces[c] - 1);
    if (k < 0) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_precond_cut_own: face carries no lambda");
    const LamFace &f = t->h_faces[k];
    if ((f.em >= 0) == (f.ep >= 0)) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_precond_cut_own: not a cut face");
    idx[c] = k; ooff[c] = o; o += (int64_t)f.nl * f.nl;
  }
  int64_t *d_idx = nullptr;
  HSBP_CUDA(ctx, cudaMalloc((void **)&d_idx, 2 * ncut * sizeof(int64_t)));
  cudaMemcpyAsync(d_idx, idx.data(), ncut * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d_idx + ncut, ooff.data(), ncut * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
  k_faceblock_own<<<(unsigned)ncut, 256, 0, ctx->stream>>>(t->d_faces, b->d_desc, t->d_S_off, t->d_S, d_idx, d_idx + ncut, out_dev);
  cudaError_t e1 = cudaGetLastError();
  if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_idx);
  if (e1 != cudaSuccess) { ctx->err = std::string("hsbp_trace_precond_cut_own: ") + cudaGetErrorString(e1); return HSBP_ERR_CUDA; }
  return HSBP_OK;
}

int hsbp_trace_precond_setup_cut(hsbp_trace *t, int64_t ncut, const int64_t *faces, const double *partner_dev) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = t->blocks->ctx;
  if (ncut < 0 || (ncut > 0 && (!faces || !partner_dev))) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_precond_setup_cut: bad arguments");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<int64_t> idx(ncut);
  for (int64_t c = 0; c < ncut; ++c) {
    idx[c] = lam_face_index(t, faces[c] - 1);
    if (idx[c] < 0) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_precond_setup_cut: face carries no lambda");
  }
  return precond_faceblocks(t, ncut, idx.data(), partner_dev);
}

int hsbp_trace_precond_apply(hsbp_trace *t, const double *r_dev, double *z_dev) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = t->blocks->ctx;
  if (!r_dev || !z_dev || r_dev == z_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_precond_apply: bad pointers");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  return precond_apply(t, r_dev, z_dev);
}

int hsbp_trace_rhs(hsbp_trace *t, const double *g_dev, const double *gd_dev, double *b_dev) {
  if (!t) return HSBP_ERR_ARG;
  if (!g_dev || !gd_dev || !b_dev) HSBP_FAIL(t->blocks->ctx, HSBP_ERR_ARG, "hsbp_trace_rhs: null pointer");
  HSBP_CUDA(t->blocks->ctx, cudaSetDevice(t->blocks->ctx->device));
  return trace_rhs(t, g_dev, gd_dev, b_dev);
}

int hsbp_trace_solve(hsbp_trace *t, const double *g_dev, const double *gd_dev, double *lam, double *u_dev,
                     double tol, int64_t maxit, hsbp_trace_stats *stats) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  if (!g_dev || !gd_dev || !lam || !u_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_solve: null pointer");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  t->acc = {0, 0, 0, 0.0};
  t->local_solves = 0;
  const int64_t n = t->lNp;
  const dim3 lg = vec_grid(n);
  double *r = t->d_r, *p = t->d_p, *q = t->d_q, *z = t->d_zz;
  int rc;
  hsbp_trace_stats st = {0, 0, 0.0, 0, 0, 0};
  if (n > 0) {
    if ((rc = trace_rhs(t, g_dev, gd_dev, r))) return rc;                  // r = b (lambda0 = 0)
    HSBP_CUDA(ctx, cudaMemsetAsync(lam, 0, n * sizeof(double), ctx->stream));
    if ((rc = precond_apply(t, r, p))) return rc;                          // p = z = preconditioned residual
    double d3[3];
    if ((rc = dots(t, n, r, p, r, r, nullptr, nullptr, d3))) return rc;
    double rz = d3[0];
    const double b2 = d3[1];
    double rr = b2;
    if (b2 > 0) {
      while (st.outer_iterations < maxit) {
        if ((rc = schur_apply(t, p, q))) return rc;                        // note: uses t->d_zz as scratch before z is needed
        if ((rc = dots(t, n, p, q, nullptr, nullptr, nullptr, nullptr, d3))) return rc;
        const double alpha = rz / d3[0];
        k_axpby<<<lg, VEC_THREADS, 0, ctx->stream>>>(n, 1.0, lam, alpha, p, lam);
        k_axpby<<<lg, VEC_THREADS, 0, ctx->stream>>>(n, 1.0, r, -alpha, q, r);
        if ((rc = precond_apply(t, r, z))) return rc;
        if ((rc = dots(t, n, r, z, r, r, nullptr, nullptr, d3))) return rc;
        st.outer_iterations += 1;
        rr = d3[1];
        if (sqrt(rr / b2) <= tol) { st.converged = 1; break; }
        const double beta = d3[0] / rz;
        rz = d3[0];
        k_axpby<<<lg, VEC_THREADS, 0, ctx->stream>>>(n, 1.0, z, beta, p, p);
      }
      st.rel_residual = sqrt(rr / b2);
    } else {
      st.converged = 1;
    }
  } else {
    st.converged = 1;
  }
  // u = M^-1 (g - Fbar lambda)
  HSBP_CUDA(ctx, cudaMemcpyAsync(t->d_w, g_dev, (size_t)b->VNp * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  if (n > 0 && (rc = trace_Fbar_add(t, lam, -1.0, t->d_w))) return rc;
  hsbp_local_stats s;
  if ((rc = local_solve_impl(b, t->d_w, u_dev, &s))) return rc;
  accumulate(t, s);
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  st.inner_iterations_sum = t->acc.iterations_sum;
  st.inner_iterations_max = t->acc.iterations_max;
  st.local_solves = t->local_solves;
  if (stats) *stats = st;
  return HSBP_OK;
}

}  // extern "C"
