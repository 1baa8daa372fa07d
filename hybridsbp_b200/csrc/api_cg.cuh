// C-ABI: the trace (lambda) system -- partitioned meshes, two-level preconditioner and the device-resident CG (K4).
// Included by hsbp.cu (unity build) after api_solve.cuh.
//
// reference: the trace solve of square_circle.jl:376-388 (lambda = BF \ b_lambda with BF = cholesky(B), u = M̃^-1 (g - Fbar lambda));
// B = D - Fbar^T M̃^-1 Fbar of assembleλmatrix (global_curved.jl:743-797) is never formed: it is applied block by block and the
// system is solved by preconditioned CG.  Blocks couple only through faces shared by exactly two blocks (global_curved.jl:525-554)
// and are independent given lambda (:732-737), which is what lets the mesh be partitioned across GPUs.
#pragma once
#include <chrono>
#include <thread>

#include "k_cg.cuh"
#include "k_dense.cuh"
#include "api_p2p.cuh"

namespace {

using namespace hsbp;

template <class T> int upload_vec(hsbp_ctx *ctx, T **dptr, const std::vector<T> &h) {
  cudaFree(*dptr); *dptr = nullptr;
  HSBP_CUDA(ctx, cudaMalloc((void **)dptr, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) HSBP_CUDA(ctx, cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // h may be a temporary
  return HSBP_OK;
}

int upload_fx(hsbp_trace *t) { return upload_vec(t->blocks->ctx, &t->d_fx, t->h_fx); }

void graph_free(hsbp_trace *t) {
  if (t->graph_exec) { cudaGraphExecDestroy((cudaGraphExec_t)t->graph_exec); t->graph_exec = nullptr; }
}

void coarse_free(hsbp_trace *t) {
  graph_free(t);
  cudaFree(t->d_AII); cudaFree(t->d_E); cudaFree(t->d_ET); cudaFree(t->d_SG); cudaFree(t->d_gidx);
  cudaFree(t->d_bI); cudaFree(t->d_bG); cudaFree(t->d_t); cudaFree(t->d_ey); cudaFree(t->d_cG);
  cudaFree(t->d_red2_in); cudaFree(t->d_red2_out);
  t->d_AII = t->d_E = t->d_ET = t->d_SG = nullptr; t->d_gidx = nullptr;
  t->d_bI = t->d_bG = t->d_t = t->d_ey = t->d_cG = nullptr; t->d_red2_in = t->d_red2_out = nullptr;
  t->cmodes = 0; t->nI = t->ldI = t->nGq = t->nGt = t->ldG = 0;
}

void precond_free(hsbp_trace *t) {
  graph_free(t);
  cudaFree(t->d_binv); t->d_binv = nullptr;
  t->precond_kind = HSBP_PRECOND_JACOBI;
  for (auto &x : t->h_fx) { x.binv_off = -1; x.binv_ld = 0; }
}

void trace_free_solver(hsbp_trace *t) {
  p2p_free(t);
  coarse_free(t);
  cudaFree(t->d_binv); cudaFree(t->d_fx); cudaFree(t->d_f2l); cudaFree(t->d_blk_lf);
  cudaFree(t->d_send); cudaFree(t->d_recv); cudaFree(t->d_facepart); cudaFree(t->d_part1); cudaFree(t->d_red1);
  cudaFree(t->d_state);
  if (t->h_status) cudaFreeHost(t->h_status);
  if (t->ev_a) cudaEventDestroy((cudaEvent_t)t->ev_a);
  if (t->ev_b) cudaEventDestroy((cudaEvent_t)t->ev_b);
}

// reduction buffers of the second all-reduce: 3 scalars + one entry per coarse dof on a cut face of the whole mesh
int alloc_red2(hsbp_trace *t) {
  hsbp_ctx *ctx = t->blocks->ctx;
  cudaFree(t->d_red2_in); cudaFree(t->d_red2_out); t->d_red2_in = t->d_red2_out = nullptr;
  const size_t n = 3 + (size_t)t->nGt;
  HSBP_CUDA(ctx, cudaMalloc((void **)&t->d_red2_in, n * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&t->d_red2_out, n * sizeof(double)));
  HSBP_CUDA(ctx, cudaMemsetAsync(t->d_red2_in, 0, n * sizeof(double), ctx->stream));
  HSBP_CUDA(ctx, cudaMemsetAsync(t->d_red2_out, 0, n * sizeof(double), ctx->stream));
  return HSBP_OK;
}

// index maps between block faces and lambda, default (unpartitioned) face data, CG scratch
int trace_build_maps(hsbp_trace *t) {
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  std::vector<int64_t> f2l((size_t)b->FNp, -1);
  std::vector<int32_t> blk_lf((size_t)(4 * b->nblocks), -1);
  t->h_fx.assign(t->h_faces.size(), LamFaceX());
  int nI = 0;
  for (size_t i = 0; i < t->h_faces.size(); ++i) {
    const LamFace &f = t->h_faces[i];
    for (int n = 0; n < f.nl; ++n) {
      if (f.em >= 0) f2l[f.fm + n] = f.loff + n;
      if (f.ep >= 0) f2l[f.fp + (f.flip ? f.nl - 1 - n : n)] = f.loff + n;
    }
    if (f.em >= 0) blk_lf[4 * (size_t)f.em + f.km] = (int32_t)i;
    if (f.ep >= 0) blk_lf[4 * (size_t)f.ep + f.kp] = (int32_t)i;
    LamFaceX &x = t->h_fx[i];
    x.msg_off = -1; x.binv_off = -1; x.gamma = -1; x.binv_ld = 0; x.owned = 1; x.cI = nI++; x.cG = -1;
  }
  int rc;
  if ((rc = upload_vec(ctx, &t->d_f2l, f2l))) return rc;
  if ((rc = upload_vec(ctx, &t->d_blk_lf, blk_lf))) return rc;
  if ((rc = upload_fx(t))) return rc;
  const size_t nf = std::max<size_t>(t->h_faces.size(), 1);
  HSBP_CUDA(ctx, cudaMalloc((void **)&t->d_facepart, 2 * nf * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&t->d_part1, nf * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&t->d_red1, 2 * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&t->d_state, sizeof(CgState)));
  HSBP_CUDA(ctx, cudaMemsetAsync(t->d_state, 0, sizeof(CgState), ctx->stream));
  HSBP_CUDA(ctx, cudaHostAlloc((void **)&t->h_status, sizeof(CgStatus), cudaHostAllocMapped));
  HSBP_CUDA(ctx, cudaHostGetDevicePointer((void **)&t->d_status, t->h_status, 0));
  memset(t->h_status, 0, sizeof(CgStatus));
  return alloc_red2(t);
}

int stale_check(hsbp_trace *t, const char *who) {
  if (t->blocks_generation != t->blocks->generation)
    HSBP_FAIL(t->blocks->ctx, HSBP_ERR_STATE, std::string(who) + ": the blocks' operator changed after this trace was built (metrics / bc / tau); "
                                                                 "create the trace again");
  return HSBP_OK;
}

// ft = F_e^T M̃_e^-1 F_e fv for every block (block-face vectors): condensed blocks or one batched local solve
int blockface_op(hsbp_trace *t, const double *fv, double *ft) {
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  if (t->d_S) {
    k_cond_gemv<<<dim3(16, (unsigned)b->nblocks), 256, (size_t)t->max_nf * sizeof(double), ctx->stream>>>(b->d_desc, t->d_S_off, t->d_S, fv, ft);
    return check_launch(ctx, "k_cond_gemv");
  }
  HSBP_CUDA(ctx, cudaMemsetAsync(t->d_w, 0, (size_t)b->VNp * sizeof(double), ctx->stream));
  int rc = hsbp_face_F_add(b, fv, 1.0, t->d_w);
  if (rc) return rc;
  hsbp_local_stats s;
  if ((rc = local_solve_impl(b, t->d_w, t->d_z, &s))) return rc;
  accumulate(t, s);
  return hsbp_face_FT(b, t->d_z, ft);
}

// ft = (block-face operator)(x seen from the block faces); kernels return at once when the CG has finished unless force
int lam_blockface(hsbp_trace *t, const double *x, int force) {
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  if (t->d_S) {
    const int rows = t->max_nf;
    const unsigned gx = (unsigned)std::max(1, std::min(16, (rows + 31) / 32));
    k_cg_gemv<<<dim3(gx, (unsigned)b->nblocks), 256, (size_t)t->max_nf * sizeof(double), ctx->stream>>>(
        t->d_state, force, b->d_desc, t->d_S_off, t->d_S, t->d_f2l, x, t->d_ft);
    return check_launch(ctx, "k_cg_gemv");
  }
  k_cg_scatter<<<vec_grid(b->FNp), VEC_THREADS, 0, ctx->stream>>>(t->d_state, 1, b->FNp, t->d_f2l, x, t->d_fv);
  int rc = check_launch(ctx, "k_cg_scatter");
  if (rc) return rc;
  return blockface_op(t, t->d_fv, t->d_ft);
}

int exchange_vec(hsbp_trace *t) {
  return comm_exchange(t->blocks->ctx, t->peers, t->peer_off, t->peer_cnt, t->d_send, t->d_recv);
}

// out = B x on this rank's lambda (all ranks call; cut faces are completed by the exchange)
int dist_schur_apply(hsbp_trace *t, const double *x, double *out) {
  hsbp_ctx *ctx = t->blocks->ctx;
  if (t->nlam_faces == 0) return HSBP_OK;
  int rc = lam_blockface(t, x, 1);
  if (rc) return rc;
  k_cg_q<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_state, 1, 0, t->d_faces, t->d_fx, t->d_D, x, nullptr, t->d_ft, out,
                                                           t->d_send, t->d_part1, t->d_red1);
  if ((rc = check_launch(ctx, "k_cg_q"))) return rc;
  if (t->partitioned) {
    if ((rc = exchange_vec(t))) return rc;
    k_cg_finish<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_faces, t->d_fx, t->d_send, t->d_recv, out, 0);
    rc = check_launch(ctx, "k_cg_finish");
  }
  return rc;
}

// b = gdelta - Fbar^T M̃^-1 g   (LocalToGLobalRHS!, global_curved.jl:730-740)
int dist_rhs(hsbp_trace *t, const double *g, const double *gd, double *bl) {
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  hsbp_local_stats s;
  int rc = local_solve_impl(b, g, t->d_z, &s);
  if (rc) return rc;
  accumulate(t, s);
  if (t->nlam_faces == 0) return HSBP_OK;
  if ((rc = hsbp_face_FT(b, t->d_z, t->d_ft))) return rc;
  k_cg_q<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_state, 1, 1, t->d_faces, t->d_fx, t->d_D, nullptr, gd, t->d_ft, bl,
                                                           t->d_send, t->d_part1, t->d_red1);
  if ((rc = check_launch(ctx, "k_cg_q"))) return rc;
  if (t->partitioned) {
    if ((rc = exchange_vec(t))) return rc;
    k_cg_finish<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_faces, t->d_fx, t->d_send, t->d_recv, bl, 0);
    rc = check_launch(ctx, "k_cg_finish");
  }
  return rc;
}

// the preconditioner stages of one CG iteration (after r is final): first level + partial sums, coarse level,
// reduction, scalars, and either the new search direction (zonly = 0) or just z (zonly = 1)
int precond_stages(hsbp_trace *t, int force, int zonly, double *lam, double *r, double *z, double *p, const double *q,
                   bool use_p2p = false) {
  hsbp_ctx *ctx = t->blocks->ctx;
  const unsigned nf = (unsigned)t->nlam_faces;
  const size_t smem = (2 * (size_t)t->max_nl + 256) * sizeof(double);
  const bool multi = ctx->world > 1;                 // a single rank reads the reduction inputs directly: nothing to sum
  const double *red1 = multi ? t->d_red1 + 1 : t->d_red1, *red2 = multi ? t->d_red2_out : t->d_red2_in;
  k_cg_update<<<nf, 256, smem, ctx->stream>>>(t->d_state, force, t->d_faces, t->d_fx, t->d_D, t->d_binv, red1, p, q, t->d_send,
                                              t->d_recv, lam, r, z, t->cmodes, t->d_bI, t->d_bG, t->d_facepart, t->d_red2_in);
  int rc = check_launch(ctx, "k_cg_update");
  if (rc) return rc;
  if (t->cmodes > 0) {
    const int rows = t->nI + t->nGq;
    k_cg_coarse<<<(unsigned)std::max(1, (rows + 7) / 8), 256, 0, ctx->stream>>>(t->d_state, force & 1, t->nI, t->ldI, t->nGq, t->d_AII, t->d_E,
                                                                               t->d_bI, t->d_bG, t->d_gidx, t->d_t, t->d_ey, t->d_facepart,
                                                                               (int)nf, t->d_red2_in);
    if ((rc = check_launch(ctx, "k_cg_coarse"))) return rc;
  }
  if (multi && use_p2p) {                              // partial sums straight into every rank's mailbox (api_p2p.cuh)
    P2PDev *pp = ((P2PHost *)t->p2p)->d_dev;
    const int n2 = 3 + (t->cmodes > 0 ? t->nGt : 0);
    k_p2p_push_b<<<1, 1024, 0, ctx->stream>>>(t->d_state, force & 1, pp, t->d_red2_in, n2);
    k_p2p_wait_b<<<1, 1024, 0, ctx->stream>>>(t->d_state, force & 1, pp, t->d_red2_out, n2);
    if ((rc = check_launch(ctx, "k_p2p_push_b / k_p2p_wait_b"))) return rc;
  } else if (multi && (rc = comm_allreduce(ctx, t->d_red2_in, t->d_red2_out, 3 + (size_t)(t->cmodes > 0 ? t->nGt : 0)))) return rc;
  const int nGt = t->cmodes > 0 ? t->nGt : 0;
  k_cg_scalars<<<(unsigned)std::max(1, (nGt + 7) / 8), 256, 0, ctx->stream>>>(t->d_state, force & 1, nGt, t->ldG, t->d_SG, red2, t->d_cG,
                                                                             t->d_status);
  if ((rc = check_launch(ctx, "k_cg_scalars"))) return rc;
  k_cg_p<<<nf, 128, 0, ctx->stream>>>(t->d_state, force & 1, zonly, t->d_faces, t->d_fx, t->cmodes, t->nI, t->nGq, t->d_t, t->d_ET, t->d_gidx,
                                      t->d_cG, z, p);
  return check_launch(ctx, "k_cg_p");
}

int trace_large_smem(hsbp_trace *t) {
  hsbp_ctx *ctx = t->blocks->ctx;
  const size_t smem = (2 * (size_t)t->max_nl + 256) * sizeof(double);
  if (smem > 200 * 1024) HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "trace solve: faces with more than 12 000 points are not supported");
  HSBP_CUDA(ctx, hsbp_smem_optin(ctx, k_cg_update, std::max<size_t>(smem, 48 * 1024)));      // (these kernels also have static shared memory)
  const size_t s2 = (size_t)t->max_nf * sizeof(double);
  if (s2 > 200 * 1024) HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "trace solve: blocks with more than 25 000 face points are not supported");
  HSBP_CUDA(ctx, hsbp_smem_optin(ctx, k_cg_gemv, std::max<size_t>(s2, 48 * 1024)));
  HSBP_CUDA(ctx, hsbp_smem_optin(ctx, k_cond_gemv, std::max<size_t>(s2, 48 * 1024)));
  return HSBP_OK;
}

// wait until the status word says that iteration `target` is complete or the CG has finished; false on a CUDA error
bool wait_iteration(hsbp_trace *t, int target) {
  volatile CgStatus *s = t->h_status;
  int spins = 0;
  while (true) {
    if (s->iter >= target || s->done_iter >= 0) return true;
    if ((++spins & 1023) == 0) {
      if (cudaStreamQuery(t->blocks->ctx->stream) != cudaErrorNotReady) {      // everything enqueued has run (or the device failed)
        return s->iter >= target || s->done_iter >= 0;
      }
      std::this_thread::yield();
    }
  }
}

// Preconditioned CG on B lambda = b with r = b on entry (d_r), lambda = 0.  Returns the statistics of the iteration.
int cg_run(hsbp_trace *t, double *lam, double tol, int64_t maxit, hsbp_trace_stats *st) {
  hsbp_ctx *ctx = t->blocks->ctx;
  int rc;
  if ((rc = trace_large_smem(t))) return rc;
  CgState h;
  memset(&h, 0, sizeof(h));
  h.tol2 = tol * tol; h.maxit = (int32_t)std::min<int64_t>(maxit, 1 << 30); h.init = 1; h.done_iter = -1;
  HSBP_CUDA(ctx, cudaMemcpyAsync(t->d_state, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  t->h_status->iter = -1; t->h_status->done_iter = -1; t->h_status->converged = 0; t->h_status->rr = 0; t->h_status->b2 = 0;
  double *r = t->d_r, *p = t->d_p, *q = t->d_q, *z = t->d_zz;
  // z = P^-1 r, p = z, r.z, b2 = r.r
  if ((rc = precond_stages(t, 0, 0, lam, r, z, p, q))) return rc;
  // an iterative local solver inside the matvec synchronises the stream anyway and makes every iteration expensive:
  // check after every iteration there; otherwise run ahead of the device by whole chunks
  const bool sync_mode = !t->d_S && (t->blocks->local_mode == HSBP_LOCAL_PCG || t->blocks->local_mode == HSBP_LOCAL_FDM);
  const int K = sync_mode ? 1 : std::max(1, t->cg_chunk), L = sync_mode ? 0 : std::max(0, t->cg_lookahead);
  int64_t issued = 0;
  if (!t->ev_a) {
    HSBP_CUDA(ctx, cudaEventCreate((cudaEvent_t *)&t->ev_a));
    HSBP_CUDA(ctx, cudaEventCreate((cudaEvent_t *)&t->ev_b));
  }
  HSBP_CUDA(ctx, cudaEventRecord((cudaEvent_t)t->ev_a, ctx->stream));
  // peer-memory path of the loop's exchanges: set up collectively (every rank reaches this point), NCCL if it is not available
  if (ctx->world > 1 && t->partitioned && t->p2p_want) {
    int rcp = p2p_setup(t);
    if (rcp) return rcp;
  } else if (t->p2p) {
    p2p_free(t);
  }
  const bool use_p2p = t->p2p != nullptr;
  // one iteration, enqueued: block-face products, q and the local part of p.q, exchange + reduction, the preconditioner stages
  auto enqueue_iteration = [&]() -> int {
    int rc_;
    if ((rc_ = lam_blockface(t, p, 0))) return rc_;
    k_cg_q<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_state, 0, 0, t->d_faces, t->d_fx, t->d_D, p, nullptr, t->d_ft, q, t->d_send,
                                                             t->d_part1, t->d_red1);
    if ((rc_ = check_launch(ctx, "k_cg_q"))) return rc_;
    if (use_p2p) {                                   // cut-face parts and the partial of p.q over NVLink, no NCCL call in the loop
      P2PDev *pp = ((P2PHost *)t->p2p)->d_dev;
      k_p2p_push_a<<<1, 1024, 0, ctx->stream>>>(t->d_state, 0, pp, t->d_send, t->d_red1);
      k_p2p_wait_a<<<1, 1024, 0, ctx->stream>>>(t->d_state, 0, pp, t->d_recv, t->d_red1);
      if ((rc_ = check_launch(ctx, "k_p2p_push_a / k_p2p_wait_a"))) return rc_;
      return precond_stages(t, 0, 0, lam, r, z, p, q, true);
    }
    if (t->partitioned && (rc_ = exchange_vec(t))) return rc_;
    if (ctx->world > 1 && (rc_ = comm_allreduce(ctx, t->d_red1, t->d_red1 + 1, 1))) return rc_;
    return precond_stages(t, 0, 0, lam, r, z, p, q);
  };
  // With condensed blocks an iteration is a fixed sequence of kernels and NCCL calls: a chunk of K iterations is captured once
  // into a CUDA graph and replayed (one launch per chunk instead of 6 + 3 per iteration).
  const bool want_graph = t->d_S != nullptr && t->cg_graph != 0;
  if (want_graph && (!t->graph_exec || t->graph_lam != lam || t->graph_K != K || t->graph_p2p != use_p2p)) {
    if (t->graph_exec) { cudaGraphExecDestroy((cudaGraphExec_t)t->graph_exec); t->graph_exec = nullptr; }
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      int rcg = HSBP_OK;
      for (int k = 0; k < K && rcg == HSBP_OK; ++k) rcg = enqueue_iteration();
      cudaError_t ec = cudaStreamEndCapture(ctx->stream, &g);
      cudaGraphExec_t ge = nullptr;
      if (rcg == HSBP_OK && ec == cudaSuccess && g && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess) {
        t->graph_exec = ge; t->graph_lam = lam; t->graph_K = K; t->graph_p2p = use_p2p;
      } else {
        t->cg_graph = 0;                      // capture is not possible here (e.g. an NCCL build without graph support): plain launches
        cudaGetLastError();
      }
      if (g) cudaGraphDestroy(g);
    } else {
      t->cg_graph = 0;
      cudaGetLastError();
    }
  }
  const bool use_graph = want_graph && t->cg_graph != 0 && t->graph_exec != nullptr;
  for (int64_t c = 0;; ++c) {
    // every rank takes this decision on the same data: the state of the iteration after (c - L) K iterations
    const int64_t seen = (c - L) * K;
    if (seen >= 0) {
      if (!wait_iteration(t, (int)std::min<int64_t>(seen, h.maxit))) break;
      const int di = t->h_status->done_iter;
      if (di >= 0 && di <= seen) break;
      if (t->h_status->iter < std::min<int64_t>(seen, h.maxit) && di < 0) {
        HSBP_FAIL(ctx, HSBP_ERR_CUDA, "trace CG: the device stopped making progress");
      }
    }
    if (use_graph) {
      HSBP_CUDA(ctx, cudaGraphLaunch((cudaGraphExec_t)t->graph_exec, ctx->stream));
      issued += K;
    } else {
      for (int k = 0; k < K; ++k, ++issued)
        if ((rc = enqueue_iteration())) return rc;
    }
  }
  HSBP_CUDA(ctx, cudaEventRecord((cudaEvent_t)t->ev_b, ctx->stream));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  {
    float ms = 0.f;
    HSBP_CUDA(ctx, cudaEventElapsedTime(&ms, (cudaEvent_t)t->ev_a, (cudaEvent_t)t->ev_b));
    st->cg_loop_ms = ms;
  }
  if (use_p2p) {
    P2PDev hd;
    HSBP_CUDA(ctx, cudaMemcpy(&hd, ((P2PHost *)t->p2p)->d_dev, sizeof(hd), cudaMemcpyDeviceToHost));
    if (hd.error) HSBP_FAIL(ctx, HSBP_ERR_NCCL, "trace CG: a partner's peer-memory flag did not arrive (rank lost or out of step)");
  }
  HSBP_CUDA(ctx, cudaMemcpy(&h, t->d_state, sizeof(h), cudaMemcpyDeviceToHost));
  st->outer_iterations = h.iter;
  st->converged = h.converged;
  st->rel_residual = h.b2 > 0 ? sqrt(h.rr / h.b2) : 0.0;
  st->issued_iterations = issued;
  st->b_norm = sqrt(h.b2);
  return HSBP_OK;
}

int require_comm_consistency(hsbp_trace *t, const char *who) {
  hsbp_ctx *ctx = t->blocks->ctx;
  if (ctx->world > 1 && !t->partitioned)
    HSBP_FAIL(ctx, HSBP_ERR_STATE, std::string(who) + ": the context has a communicator: call hsbp_trace_set_partition first (every rank, also one without cut faces)");
  return HSBP_OK;
}

}  // namespace

extern "C" {

int hsbp_trace_set_option(hsbp_trace *t, const char *name, int64_t value) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = t->blocks->ctx;
  if (!name) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_set_option: null name");
  const std::string n(name);
  if (n == "cg_chunk") t->cg_chunk = (int)std::max<int64_t>(1, value);
  else if (n == "cg_graph") t->cg_graph = value ? 1 : 0;
  else if (n == "cg_lookahead") t->cg_lookahead = (int)std::max<int64_t>(0, value);
  else if (n == "cg_p2p") t->p2p_want = value ? 1 : 0;
  else HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_set_option: unknown option " + n);
  return HSBP_OK;
}

int hsbp_trace_comm_path(const hsbp_trace *t) {
  if (!t) return -1;
  if (t->blocks->ctx->world <= 1) return 0;
  return t->p2p ? 2 : 1;
}

int hsbp_trace_last_local_stats(hsbp_trace *t, hsbp_local_stats *stats) {
  if (!t || !stats) return HSBP_ERR_ARG;
  *stats = t->acc;
  return HSBP_OK;
}

int hsbp_trace_set_partition(hsbp_trace *t, int64_t ncut, const int64_t *faces, const int64_t *partner, const int64_t *gamma,
                             int64_t n_gamma_total) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  if (ncut < 0 || n_gamma_total < ncut || (ncut > 0 && (!faces || !partner || !gamma)))
    HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_set_partition: bad arguments");
  if (t->partitioned) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_set_partition: already partitioned");
  if (ncut > 0 && ctx->world < 2) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_set_partition: cut faces need a communicator (hsbp_comm_init)");
  if (ctx->world > 1 && t->nlam_faces == 0)
    HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "hsbp_trace_set_partition: a rank whose blocks touch no interface face cannot take part in a partitioned solve");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  struct Cut { int64_t gamma; int peer; int32_t lf; };
  std::vector<Cut> cuts((size_t)ncut);
  std::vector<char> listed(t->h_faces.size(), 0);
  for (int64_t c = 0; c < ncut; ++c) {
    const int64_t f = faces[c] - 1;
    if (f < 0 || f >= t->nfaces || t->face2lam[f] < 0) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_set_partition: face carries no lambda");
    const int32_t lfi = (int32_t)t->face2lam[f];
    const LamFace &lf = t->h_faces[lfi];
    if ((lf.em >= 0) == (lf.ep >= 0)) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_set_partition: not a cut face (both sides are local)");
    if (partner[c] < 0 || partner[c] >= ctx->world || partner[c] == ctx->rank) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_set_partition: bad partner rank");
    if (gamma[c] < 0 || gamma[c] >= n_gamma_total || listed[lfi]) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_set_partition: bad cut-face index");
    listed[lfi] = 1;
    cuts[c] = {gamma[c], (int)partner[c], lfi};
  }
  for (size_t i = 0; i < t->h_faces.size(); ++i)
    if (((t->h_faces[i].em >= 0) != (t->h_faces[i].ep >= 0)) && !listed[i])
      HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_set_partition: a face with one local side is missing from the list");
  // message order: partners by rank, the faces of one partner by their global cut-face index (both ranks agree on it)
  std::sort(cuts.begin(), cuts.end(), [](const Cut &a, const Cut &c) { return a.peer != c.peer ? a.peer < c.peer : a.gamma < c.gamma; });
  t->peers.clear(); t->peer_off.clear(); t->peer_cnt.clear(); t->peer_boff.clear(); t->peer_bcnt.clear(); t->cut_faces.clear();
  int64_t off = 0, boff = 0;
  int nI = 0;
  for (auto &x : t->h_fx) { x.msg_off = -1; x.gamma = -1; x.cG = -1; x.owned = 1; x.cI = -1; }
  for (size_t c = 0; c < cuts.size(); ++c) {
    if (t->peers.empty() || t->peers.back() != cuts[c].peer) {
      t->peers.push_back(cuts[c].peer); t->peer_off.push_back(off); t->peer_cnt.push_back(0); t->peer_boff.push_back(boff); t->peer_bcnt.push_back(0);
    }
    const LamFace &lf = t->h_faces[cuts[c].lf];
    LamFaceX &x = t->h_fx[cuts[c].lf];
    x.msg_off = off; x.gamma = cuts[c].gamma; x.cG = (int32_t)c; x.owned = lf.em >= 0 ? 1 : 0;
    t->cut_faces.push_back(cuts[c].lf);
    off += lf.nl; boff += (int64_t)lf.nl * lf.nl;
    t->peer_cnt.back() += lf.nl; t->peer_bcnt.back() += (int64_t)lf.nl * lf.nl;
  }
  for (auto &x : t->h_fx) if (x.cG < 0) x.cI = nI++;
  t->msg_len = off; t->bmsg_len = boff; t->n_gamma = n_gamma_total;
  int rc = upload_fx(t);
  if (rc) return rc;
  cudaFree(t->d_send); cudaFree(t->d_recv); t->d_send = t->d_recv = nullptr;
  HSBP_CUDA(ctx, cudaMalloc((void **)&t->d_send, std::max<size_t>((size_t)off, 1) * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&t->d_recv, std::max<size_t>((size_t)off, 1) * sizeof(double)));
  t->partitioned = true;
  // D = Hf (tau_minus + tau_plus): add the partner's half on every cut face (global_curved.jl:556-557)
  if (t->nlam_faces && t->msg_len > 0) {
    k_cg_pack<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_faces, t->d_fx, t->d_D, t->d_send);
    if ((rc = check_launch(ctx, "k_cg_pack"))) return rc;
    if ((rc = exchange_vec(t))) return rc;
    k_cg_finish<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_faces, t->d_fx, t->d_send, t->d_recv, t->d_D, 1);
    if ((rc = check_launch(ctx, "k_cg_finish"))) return rc;
  }
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // preconditioners built before the partition was known do not know the partner's contributions
  precond_free(t);
  coarse_free(t);
  if ((rc = upload_fx(t))) return rc;
  return alloc_red2(t);
}

}  // extern "C"

namespace {

// First level: explicit inverses of the diagonal blocks B_ff = D_f - S_e-[f, f] - S_e+[f, f] (orientation applied on the
// plus side).  On a cut face the other rank's S_e[f, f] arrives by send / recv; D_f - (own + partner) is a commutative
// sum, so both ranks invert the same bits and their copies of lambda stay identical.
int precond_faceblocks(hsbp_trace *t) {
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  if (!t->d_S) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_precond_setup: the face-block preconditioner needs hsbp_trace_condense first");
  precond_free(t);
  const int64_t nf = t->nlam_faces;
  if (nf == 0) { t->precond_kind = HSBP_PRECOND_FACE_BLOCKS; return HSBP_OK; }
  static_assert(sizeof(FaceBlock) == sizeof(CholBlock), "descriptor layout");
  std::vector<CholBlock> fbs((size_t)nf);
  int64_t off = 0, woff = 0;
  for (int64_t i = 0; i < nf; ++i) {
    const LamFace &f = t->h_faces[i];
    const int ld = (f.nl + GJ_NB - 1) / GJ_NB * GJ_NB;
    fbs[i].off = off; fbs[i].np = f.nl; fbs[i].ld = ld; fbs[i].voff = f.loff; fbs[i].woff = woff;
    off += (int64_t)ld * ld; woff += ld;
  }
  CholBlock *d_fb = nullptr;
  int64_t *d_idx = nullptr, *d_pidx = nullptr;
  double *bsend = nullptr, *brecv = nullptr;
  auto cleanup = [&]() { cudaFree(d_fb); cudaFree(d_idx); cudaFree(d_pidx); cudaFree(bsend); cudaFree(brecv); };
  auto fail = [&](int rc) { cleanup(); precond_free(t); return rc; };
  int rc;
  if ((rc = upload_vec(ctx, &d_fb, fbs))) return fail(rc);
  if (cudaMalloc((void **)&t->d_binv, (size_t)off * sizeof(double)) != cudaSuccess) { ctx->err = "face-block preconditioner: out of device memory"; return fail(HSBP_ERR_CUDA); }
  const int64_t ncut = (int64_t)t->cut_faces.size();
  if (ncut > 0) {
    // pack this rank's S_e[f, f] (lambda orientation) in message order, exchange, hand the partner's blocks to the fill kernel
    std::vector<int64_t> idx(2 * (size_t)ncut), pidx((size_t)nf, -1);
    int64_t o = 0;
    for (int64_t c = 0; c < ncut; ++c) {
      const LamFace &f = t->h_faces[t->cut_faces[c]];
      idx[c] = t->cut_faces[c]; idx[ncut + c] = o; pidx[t->cut_faces[c]] = o;
      o += (int64_t)f.nl * f.nl;
    }
    if ((rc = upload_vec(ctx, &d_idx, idx)) || (rc = upload_vec(ctx, &d_pidx, pidx))) return fail(rc);
    if (cudaMalloc((void **)&bsend, (size_t)o * sizeof(double)) != cudaSuccess || cudaMalloc((void **)&brecv, (size_t)o * sizeof(double)) != cudaSuccess) {
      ctx->err = "face-block preconditioner: out of device memory"; return fail(HSBP_ERR_CUDA);
    }
    k_faceblock_own<<<(unsigned)ncut, 256, 0, ctx->stream>>>(t->d_faces, b->d_desc, t->d_S_off, t->d_S, d_idx, d_idx + ncut, bsend);
    if ((rc = check_launch(ctx, "k_faceblock_own"))) return fail(rc);
    if ((rc = comm_exchange(ctx, t->peers, t->peer_boff, t->peer_bcnt, bsend, brecv))) return fail(rc);
  }
  k_faceblock_fill<<<(unsigned)nf, 256, 0, ctx->stream>>>(t->d_faces, (const FaceBlock *)d_fb, b->d_desc, t->d_S_off, t->d_S, t->d_D, t->d_binv,
                                                         d_pidx, brecv);
  if ((rc = check_launch(ctx, "k_faceblock_fill"))) return fail(rc);
  if ((rc = dense_spd_inverse_batched(ctx, fbs, d_fb, t->d_binv, "face-block preconditioner (a diagonal block of B)"))) return fail(rc);
  for (int64_t i = 0; i < nf; ++i) { t->h_fx[i].binv_off = fbs[i].off; t->h_fx[i].binv_ld = fbs[i].ld; }
  cleanup();
  t->precond_kind = HSBP_PRECOND_FACE_BLOCKS;
  return upload_fx(t);
}

// Second level: Z = `modes` Legendre polynomials per face (in the lambda orientation), A_c = Z^T B Z eliminated rank by rank.
//   I   coarse dofs of this rank's uncut faces -> A_II (dense, inverted in place),
//   G   coarse dofs of the cut faces of the whole mesh -> S_G = A_GG - sum_ranks A_GI A_II^-1 A_IG (replicated, inverted),
//   E = A_II^-1 A_IG couples the two.  Z^T B Z comes block by block from T_e = Z_e^T (F_e^T M̃_e^-1 F_e) Z_e, 4 * modes
//   applications of the block-face operator for all blocks at once, and Z^T D Z on the faces.
int coarse_setup(hsbp_trace *t, int modes) {
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  coarse_free(t);
  int rc;
  if (modes <= 0) return alloc_red2(t);
  if (modes > 3) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_coarse_setup: at most 3 modes per face");
  if (!t->d_S && b->local_mode == 0) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_coarse_setup: call hsbp_local_setup (or hsbp_trace_condense) first");
  if ((rc = trace_large_smem(t))) return rc;
  int nIf = 0;
  for (auto &x : t->h_fx) if (x.cI >= 0) nIf++;
  const int nI = modes * nIf, nGq = modes * (int)t->cut_faces.size(), nGt = modes * (int)t->n_gamma;
  const int ldI = (nI + GJ_NB - 1) / GJ_NB * GJ_NB, ldG = (nGt + GJ_NB - 1) / GJ_NB * GJ_NB;
  if (nI > 16384 || nGt > 16384) HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "hsbp_trace_coarse_setup: coarse problem too large for the dense elimination (use fewer modes)");
  const int nq = 4 * modes;
  double *T = nullptr, *AIG = nullptr, *AGG = nullptr, *SGp = nullptr;
  auto cleanup = [&]() { cudaFree(T); cudaFree(AIG); cudaFree(AGG); cudaFree(SGp); };
  auto fail = [&](int rc_) { cleanup(); coarse_free(t); alloc_red2(t); return rc_; };
  auto dalloc = [&](double **p, size_t n) {
    if (cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(double)) != cudaSuccess) return false;
    return cudaMemsetAsync(*p, 0, std::max<size_t>(n, 1) * sizeof(double), ctx->stream) == cudaSuccess;
  };
  if (!dalloc(&T, (size_t)b->nblocks * nq * nq) || !dalloc(&AIG, (size_t)nI * nGq) || !dalloc(&AGG, (size_t)nGq * nGq) || !dalloc(&SGp, (size_t)nGq * nGq) ||
      !dalloc(&t->d_AII, (size_t)ldI * ldI) || !dalloc(&t->d_E, (size_t)nI * nGq) || !dalloc(&t->d_ET, (size_t)nI * nGq) ||
      !dalloc(&t->d_SG, (size_t)ldG * ldG) || !dalloc(&t->d_bI, nI) || !dalloc(&t->d_bG, nGq) || !dalloc(&t->d_t, nI) ||
      !dalloc(&t->d_ey, nGq) || !dalloc(&t->d_cG, nGt)) {
    ctx->err = "hsbp_trace_coarse_setup: out of device memory";
    return fail(HSBP_ERR_CUDA);
  }
  std::vector<int64_t> gidx((size_t)nGq);
  for (size_t c = 0; c < t->cut_faces.size(); ++c)
    for (int m = 0; m < modes; ++m) gidx[modes * c + m] = (int64_t)modes * t->h_fx[t->cut_faces[c]].gamma + m;
  if ((rc = upload_vec(ctx, &t->d_gidx, gidx))) return fail(rc);
  // T_e, one column (local face k, mode m) at a time for all blocks
  for (int k = 0; k < 4; ++k)
    for (int m = 0; m < modes; ++m) {
      k_coarse_unit<<<(unsigned)b->nblocks, 128, 0, ctx->stream>>>(b->d_desc, t->d_faces, t->d_f2l, t->d_blk_lf, k, m, t->d_fv);
      if ((rc = check_launch(ctx, "k_coarse_unit"))) return fail(rc);
      if ((rc = blockface_op(t, t->d_fv, t->d_ft))) return fail(rc);
      k_coarse_project<<<(unsigned)b->nblocks, 128, 0, ctx->stream>>>(b->d_desc, t->d_faces, t->d_f2l, t->d_blk_lf, modes, k * modes + m, t->d_ft, T);
      if ((rc = check_launch(ctx, "k_coarse_project"))) return fail(rc);
    }
  k_coarse_assemble<<<(unsigned)b->nblocks, 64, 0, ctx->stream>>>(t->d_fx, t->d_blk_lf, modes, T, nI, ldI, nGq, t->d_AII, AIG, AGG);
  if (t->nlam_faces)
    k_coarse_diag<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_faces, t->d_fx, modes, t->d_D, ldI, nGq, t->d_AII, AGG);
  if (ldI > nI) k_dense_pad_identity<<<1, 32, 0, ctx->stream>>>(t->d_AII, nI, ldI);
  if ((rc = check_launch(ctx, "coarse assembly"))) return fail(rc);
  if (nI > 0) {
    std::vector<CholBlock> cb(1);
    cb[0].off = 0; cb[0].np = nI; cb[0].ld = ldI; cb[0].voff = 0; cb[0].woff = 0;
    CholBlock *d_cb = nullptr;
    if ((rc = upload_vec(ctx, &d_cb, cb))) return fail(rc);
    rc = dense_spd_inverse_batched(ctx, cb, d_cb, t->d_AII, "coarse matrix A_II");
    cudaFree(d_cb);
    if (rc) return fail(rc);
  }
  if (nGq > 0 && nI > 0) {
    k_dense_gemm_nn<<<dim3((unsigned)((nI + 255) / 256), (unsigned)nGq), 256, 0, ctx->stream>>>(nI, nGq, nI, t->d_AII, ldI, AIG, nI, t->d_E, nI);
    k_dense_transpose<<<(unsigned)(((int64_t)nI * nGq + 255) / 256), 256, 0, ctx->stream>>>(nI, nGq, t->d_E, t->d_ET);
    if ((rc = check_launch(ctx, "coarse E"))) return fail(rc);
  }
  if (nGq > 0) {
    // this rank's part of S_G: A_GG(own) - A_IG^T E
    k_dense_sub_atb<<<(unsigned)(((int64_t)nGq * nGq + 7) / 8), 256, 0, ctx->stream>>>(nGq, nGq, nI, AIG, nI, t->d_E, nI, AGG, SGp, nGq);
    k_coarse_place<<<(unsigned)((nGq * nGq + 255) / 256), 256, 0, ctx->stream>>>(nGq, t->d_gidx, SGp, ldG, t->d_SG);
    if ((rc = check_launch(ctx, "coarse S_G"))) return fail(rc);
  }
  if (nGt > 0) {
    if ((rc = comm_allreduce(ctx, t->d_SG, t->d_SG, (size_t)ldG * ldG))) return fail(rc);
    if (ldG > nGt) k_dense_pad_identity<<<1, 32, 0, ctx->stream>>>(t->d_SG, nGt, ldG);
    std::vector<CholBlock> cb(1);
    cb[0].off = 0; cb[0].np = nGt; cb[0].ld = ldG; cb[0].voff = 0; cb[0].woff = 0;
    CholBlock *d_cb = nullptr;
    if ((rc = upload_vec(ctx, &d_cb, cb))) return fail(rc);
    rc = dense_spd_inverse_batched(ctx, cb, d_cb, t->d_SG, "coarse matrix S_Gamma");
    cudaFree(d_cb);
    if (rc) return fail(rc);
  }
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cleanup();
  t->cmodes = modes; t->nI = nI; t->ldI = ldI; t->nGq = nGq; t->nGt = nGt; t->ldG = ldG;
  return alloc_red2(t);
}

}  // namespace

extern "C" {

int hsbp_trace_schur_apply(hsbp_trace *t, const double *lam_dev, double *out_dev) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = t->blocks->ctx;
  if (!out_dev || !lam_dev || out_dev == lam_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_schur_apply: bad pointers");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = stale_check(t, "hsbp_trace_schur_apply")) || (rc = require_comm_consistency(t, "hsbp_trace_schur_apply")) || (rc = trace_large_smem(t))) return rc;
  t->acc = {0, 0, 0, 0.0};
  if ((rc = dist_schur_apply(t, lam_dev, out_dev))) return rc;
  if (t->acc.failed_blocks) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_schur_apply: a local solve did not converge (hsbp_trace_last_local_stats)");
  return HSBP_OK;
}

int hsbp_trace_rhs(hsbp_trace *t, const double *g_dev, const double *gd_dev, double *b_dev) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = t->blocks->ctx;
  if (!g_dev || !gd_dev || !b_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_rhs: null pointer");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = stale_check(t, "hsbp_trace_rhs")) || (rc = require_comm_consistency(t, "hsbp_trace_rhs"))) return rc;
  t->acc = {0, 0, 0, 0.0};
  if ((rc = dist_rhs(t, g_dev, gd_dev, b_dev))) return rc;
  if (t->acc.failed_blocks) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_rhs: a local solve did not converge (hsbp_trace_last_local_stats)");
  return HSBP_OK;
}

int hsbp_trace_condense(hsbp_trace *t, int enable) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = t->blocks->ctx;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  precond_free(t);                   // inverses / coarse matrices of an older S
  coarse_free(t);
  int rc = upload_fx(t);
  if (rc || (rc = alloc_red2(t))) return rc;
  if (!enable) {
    cudaFree(t->d_S); cudaFree(t->d_S_off); t->d_S = nullptr; t->d_S_off = nullptr;
    return HSBP_OK;
  }
  if ((rc = stale_check(t, "hsbp_trace_condense"))) return rc;
  t->acc = {0, 0, 0, 0.0};
  if ((rc = trace_condense(t))) return rc;
  if (t->acc.failed_blocks) {
    cudaFree(t->d_S); cudaFree(t->d_S_off); t->d_S = nullptr; t->d_S_off = nullptr;
    HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_condense: a local solve did not reach its tolerance; the condensed blocks were discarded "
                                   "(raise maxit / choose another local solver)");
  }
  return HSBP_OK;
}

int hsbp_trace_precond_setup(hsbp_trace *t, int kind) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = t->blocks->ctx;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = stale_check(t, "hsbp_trace_precond_setup")) || (rc = require_comm_consistency(t, "hsbp_trace_precond_setup"))) return rc;
  if (kind == HSBP_PRECOND_JACOBI) {
    HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    precond_free(t);
    return upload_fx(t);
  }
  if (kind != HSBP_PRECOND_FACE_BLOCKS) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_precond_setup: unknown kind");
  return precond_faceblocks(t);
}

int hsbp_trace_coarse_setup(hsbp_trace *t, int modes) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = t->blocks->ctx;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = stale_check(t, "hsbp_trace_coarse_setup")) || (rc = require_comm_consistency(t, "hsbp_trace_coarse_setup"))) return rc;
  t->acc = {0, 0, 0, 0.0};
  if ((rc = coarse_setup(t, modes))) return rc;
  if (t->acc.failed_blocks) {
    coarse_free(t); alloc_red2(t);
    HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_coarse_setup: a local solve did not reach its tolerance");
  }
  return HSBP_OK;
}

int64_t hsbp_trace_coarse_size(const hsbp_trace *t) { return t ? (int64_t)t->nI + t->nGt : -1; }

// z = P^-1 r with the preconditioner hsbp_trace_solve uses (both levels); collective on a partitioned mesh
int hsbp_trace_precond_apply(hsbp_trace *t, const double *r_dev, double *z_dev) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = t->blocks->ctx;
  if (!r_dev || !z_dev || r_dev == z_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_precond_apply: bad pointers");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = stale_check(t, "hsbp_trace_precond_apply")) || (rc = require_comm_consistency(t, "hsbp_trace_precond_apply")) || (rc = trace_large_smem(t))) return rc;
  if (t->nlam_faces == 0) return HSBP_OK;
  return precond_stages(t, 3, 1, nullptr, const_cast<double *>(r_dev), z_dev, nullptr, nullptr);
}

/* lambda = B^-1 (gdelta - Fbar^T M̃^-1 g), u = M̃^-1 (g - Fbar lambda)   (square_circle.jl:376-388) */
int hsbp_trace_solve(hsbp_trace *t, const double *g_dev, const double *gd_dev, double *lam, double *u_dev,
                     double tol, int64_t maxit, hsbp_trace_stats *stats) {
  if (!t) return HSBP_ERR_ARG;
  hsbp_blocks *b = t->blocks;
  hsbp_ctx *ctx = b->ctx;
  if (!g_dev || !gd_dev || !lam || !u_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_trace_solve: null pointer");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = stale_check(t, "hsbp_trace_solve")) || (rc = require_comm_consistency(t, "hsbp_trace_solve"))) return rc;
  if (b->local_mode == 0) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_trace_solve: call hsbp_local_setup first");
  t->acc = {0, 0, 0, 0.0};
  t->local_solves = 0;
  const int64_t n = t->lNp;
  hsbp_trace_stats st;
  memset(&st, 0, sizeof(st));
  st.coarse_dofs = t->cmodes > 0 ? (int64_t)t->nI + t->nGt : 0;
  const bool collective = ctx->world > 1;
  // HSBP_TRACE_TIMING=1: phase times of this call on stderr (synchronising; diagnostics only)
  const bool timing = getenv("HSBP_TRACE_TIMING") != nullptr;
  auto tnow = [&]() { if (timing) cudaStreamSynchronize(ctx->stream); return std::chrono::steady_clock::now(); };
  auto tph = tnow();
  auto tprint = [&](const char *what) {
    if (!timing) return;
    auto t1 = tnow();
    fprintf(stderr, "[trace_solve] %-34s %8.2f ms\n", what, 1e3 * std::chrono::duration<double>(t1 - tph).count());
    tph = t1;
  };
  if (n > 0 || collective) {
    if ((rc = dist_rhs(t, g_dev, gd_dev, t->d_r))) return rc;                        // r = b (lambda0 = 0)
    tprint("right-hand side (local solve)");
    if (n > 0) {
      HSBP_CUDA(ctx, cudaMemcpyAsync(t->d_b, t->d_r, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      HSBP_CUDA(ctx, cudaMemsetAsync(lam, 0, n * sizeof(double), ctx->stream));
    }
    if ((rc = cg_run(t, lam, tol, maxit, &st))) return rc;
    tprint("CG");
    // true residual ||b - B lambda|| / ||b|| with one more application of B (the recurrence can drift when B is applied
    // through inexact local solves)
    if ((rc = dist_schur_apply(t, lam, t->d_q))) return rc;
    if (t->nlam_faces) {
      k_cg_resid<<<(unsigned)t->nlam_faces, 128, 0, ctx->stream>>>(t->d_state, t->d_faces, t->d_fx, t->d_b, t->d_q, t->d_part1, t->d_red1);
      if ((rc = check_launch(ctx, "k_cg_resid"))) return rc;
    } else {
      HSBP_CUDA(ctx, cudaMemsetAsync(t->d_red1, 0, sizeof(double), ctx->stream));
    }
    if ((rc = comm_allreduce(ctx, t->d_red1, t->d_red1 + 1, 1))) return rc;
    double rt = 0.0;
    HSBP_CUDA(ctx, cudaMemcpyAsync(&rt, t->d_red1 + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    st.true_rel_residual = st.b_norm > 0 ? sqrt(rt) / st.b_norm : 0.0;
  } else {
    st.converged = 1;
  }
  tprint("true residual");
  // u = M^-1 (g - Fbar lambda)
  HSBP_CUDA(ctx, cudaMemcpyAsync(t->d_w, g_dev, (size_t)b->VNp * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  if (n > 0 && (rc = trace_Fbar_add(t, lam, -1.0, t->d_w))) return rc;
  hsbp_local_stats s;
  if ((rc = local_solve_impl(b, t->d_w, u_dev, &s))) return rc;
  accumulate(t, s);
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  tprint("back-substitution (local solve)");
  st.inner_iterations_sum = t->acc.iterations_sum;
  st.inner_iterations_max = t->acc.iterations_max;
  st.local_solves = t->local_solves;
  st.failed_local_blocks = t->acc.failed_blocks;
  st.max_local_rel_residual = t->acc.max_rel_residual;
  if (st.failed_local_blocks) st.converged = 0;
  if (stats) *stats = st;
  return HSBP_OK;
}

}  // extern "C"
