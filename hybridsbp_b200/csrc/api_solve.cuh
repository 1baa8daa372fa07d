// C-ABI: local solves (K2), trace operators (K3) and the Schur-complement CG (K4).
// Included by hsbp.cu (unity build).
#pragma once
#include "k_solve.cuh"
#include "k_cg.cuh"

struct hsbp_trace {
  hsbp_blocks *blocks = nullptr;
  int64_t nfaces = 0, nlam_faces = 0, lNp = 0;
  std::vector<int64_t> starts;          // nfaces + 1, 1-based (FToλstarts)
  std::vector<int64_t> face2lam;        // per face (FToB order): index among the faces that carry lambda, -1
  std::vector<hsbp::LamFace> h_faces;
  hsbp::LamFace *d_faces = nullptr;
  std::vector<hsbp::LamFaceX> h_fx;     // partition / preconditioner data of every lambda face (k_cg.cuh)
  hsbp::LamFaceX *d_fx = nullptr;
  int64_t *d_f2l = nullptr;             // FNp: lambda index seen from every block-face point (orientation applied), -1
  int32_t *d_blk_lf = nullptr;          // 4 * nblocks: lambda-face index of every block face, -1
  int max_nl = 0;
  double *d_D = nullptr;
  double *d_ft = nullptr, *d_fv = nullptr;          // block-face scratch (FNp)
  double *d_w = nullptr, *d_z = nullptr;            // volume scratch (VNp)
  double *d_r = nullptr, *d_p = nullptr, *d_q = nullptr, *d_zz = nullptr, *d_b = nullptr;   // lambda scratch
  hsbp_local_stats acc = {0, 0, 0, 0.0};
  int64_t local_solves = 0;
  // static condensation (hsbp_trace_condense): dense S_e = F_e^T M̃_e^-1 F_e of every block
  double *d_S = nullptr;
  int64_t *d_S_off = nullptr;
  int max_nf = 0;
  // partitioned mesh (hsbp_trace_set_partition): message layout of the cut faces
  bool partitioned = false;
  int64_t n_gamma = 0;                  // cut faces of the whole mesh
  std::vector<int> peers;               // partner ranks, increasing
  std::vector<int64_t> peer_off, peer_cnt, peer_boff, peer_bcnt;    // per peer: vector message (doubles), face-block message
  std::vector<int32_t> cut_faces;       // lambda-face indices of this rank's cut faces in message order
  int64_t msg_len = 0, bmsg_len = 0;
  double *d_send = nullptr, *d_recv = nullptr;
  // first level of the preconditioner: explicit inverses of the diagonal blocks B_ff (hsbp_trace_precond_setup)
  int precond_kind = 0;
  double *d_binv = nullptr;
  // second level (hsbp_trace_coarse_setup)
  int cmodes = 0, nI = 0, ldI = 0, nGq = 0, nGt = 0, ldG = 0;
  double *d_AII = nullptr, *d_E = nullptr, *d_ET = nullptr, *d_SG = nullptr;
  int64_t *d_gidx = nullptr;
  double *d_bI = nullptr, *d_bG = nullptr, *d_t = nullptr, *d_ey = nullptr, *d_cG = nullptr;
  // CG state
  double *d_facepart = nullptr, *d_part1 = nullptr, *d_red1 = nullptr, *d_red2_in = nullptr, *d_red2_out = nullptr;
  hsbp::CgState *d_state = nullptr;
  hsbp::CgStatus *h_status = nullptr, *d_status = nullptr;
  int cg_chunk = 4, cg_lookahead = 1, cg_graph = 1;
  void *graph_exec = nullptr;           // cudaGraphExec_t of one chunk of CG iterations (condensed blocks), for graph_lam / graph_K
  const double *graph_lam = nullptr;
  int graph_K = 0;
  void *ev_a = nullptr, *ev_b = nullptr;    // CUDA events around the iteration loop (hsbp_trace_stats::cg_loop_ms)
  uint64_t blocks_generation = 0;       // generation of the blocks' operator the condensed / preconditioner data belong to
  // peer-memory exchange of the iteration loop (api_p2p.cuh); null: NCCL
  void *p2p = nullptr;
  int p2p_want = 1;                     // option "cg_p2p"
  bool graph_p2p = false;               // the captured graph belongs to this path
};

namespace {

using namespace hsbp;

int trace_build_maps(hsbp_trace *t);                                                  // api_cg.cuh
void trace_free_solver(hsbp_trace *t);
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
  t->face2lam.assign(nfaces, -1);
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
    t->face2lam[f] = (int64_t)t->h_faces.size();
    t->h_faces.push_back(lf);
    t->max_nl = std::max(t->max_nl, nl);
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
  A((void **)&t->d_r, lb); A((void **)&t->d_p, lb); A((void **)&t->d_q, lb); A((void **)&t->d_zz, lb); A((void **)&t->d_b, lb);
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
  {
    int rc = trace_build_maps(t);
    if (rc) { hsbp_trace_destroy(t); return rc; }
  }
  cudaStreamSynchronize(ctx->stream);
  t->blocks_generation = b->generation;
  *out = t;
  return HSBP_OK;
}

int hsbp_trace_destroy(hsbp_trace *t) {
  if (!t) return HSBP_ERR_ARG;
  cudaSetDevice(t->blocks->ctx->device);
  cudaStreamSynchronize(t->blocks->ctx->stream);
  cudaFree(t->d_faces); cudaFree(t->d_D); cudaFree(t->d_ft); cudaFree(t->d_fv); cudaFree(t->d_w); cudaFree(t->d_z);
  cudaFree(t->d_r); cudaFree(t->d_p); cudaFree(t->d_q); cudaFree(t->d_zz); cudaFree(t->d_b);
  cudaFree(t->d_S); cudaFree(t->d_S_off);
  trace_free_solver(t);
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

}  // extern "C"
