// C-ABI: SEAS BP1 ODE right-hand side (K5).  Included by hsbp.cu (unity build).
#pragma once
#include "k_bp1.cuh"

struct hsbp_bp1 {
  hsbp_blocks *blocks = nullptr;
  int64_t block = 0;                 // 0-based
  int kf = 0, kl = 0;                // 0-based local faces: fault, loading
  int nf = 0, nl = 0;
  int64_t off_fault = 0, off_load = 0;
  hsbp_bp1_params prm;
  double *d_a = nullptr, *d_sJ = nullptr, *d_state = nullptr, *d_out = nullptr;
  double *d_v = nullptr, *d_ge = nullptr, *d_u = nullptr;
  int *d_flags = nullptr;
};

extern "C" {

int hsbp_bp1_create(hsbp_blocks *b, int64_t block, int64_t fault_face, int64_t loading_face, const double *a,
                    const double *sJ, const hsbp_bp1_params *prm, hsbp_bp1 **out) {
  if (!b || !out) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  *out = nullptr;
  if (!a || !sJ || !prm) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_bp1_create: null pointer");
  if (block < 1 || block > b->nblocks || fault_face < 1 || fault_face > 4 || loading_face < 1 || loading_face > 4 ||
      fault_face == loading_face)
    HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_bp1_create: bad block / face ids (1-based)");
  const BlockDesc &d = b->h_desc[block - 1];
  if (d.bc[fault_face - 1] != HSBP_BC_DIRICHLET || d.bc[loading_face - 1] != HSBP_BC_DIRICHLET)
    HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_bp1_create: fault and loading faces must be Dirichlet faces (BP1.jl:73)");
  hsbp_bp1 *f = new (std::nothrow) hsbp_bp1();
  if (!f) HSBP_FAIL(ctx, HSBP_ERR_STATE, "out of host memory");
  f->blocks = b; f->block = block - 1; f->kf = (int)fault_face - 1; f->kl = (int)loading_face - 1; f->prm = *prm;
  const int Nrp = d.Nr + 1, Nsp = d.Ns + 1;
  auto fstart = [&](int k) { return (int64_t)(k < 2 ? k * Nsp : 2 * Nsp + (k - 2) * Nrp); };
  f->nf = f->kf < 2 ? Nsp : Nrp; f->nl = f->kl < 2 ? Nsp : Nrp;
  f->off_fault = d.foff + fstart(f->kf); f->off_load = d.foff + fstart(f->kl);
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaSuccess;
  auto A = [&](void **p_, size_t n) { if (e == cudaSuccess) e = cudaMalloc(p_, n); };
  A((void **)&f->d_a, f->nf * sizeof(double)); A((void **)&f->d_sJ, f->nf * sizeof(double));
  A((void **)&f->d_state, 2 * f->nf * sizeof(double)); A((void **)&f->d_out, 2 * f->nf * sizeof(double));
  A((void **)&f->d_v, (size_t)b->FNp * sizeof(double));
  A((void **)&f->d_ge, (size_t)b->VNp * sizeof(double)); A((void **)&f->d_u, (size_t)b->VNp * sizeof(double));
  A((void **)&f->d_flags, 4 * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_a, a, f->nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_sJ, sJ, f->nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->d_v, 0, (size_t)b->FNp * sizeof(double), ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    ctx->err = std::string("hsbp_bp1_create: ") + cudaGetErrorString(e);
    hsbp_bp1_destroy(f);
    return HSBP_ERR_CUDA;
  }
  *out = f;
  return HSBP_OK;
}

int hsbp_bp1_destroy(hsbp_bp1 *f) {
  if (!f) return HSBP_ERR_ARG;
  cudaSetDevice(f->blocks->ctx->device);
  cudaStreamSynchronize(f->blocks->ctx->stream);
  cudaFree(f->d_a); cudaFree(f->d_sJ); cudaFree(f->d_state); cudaFree(f->d_out);
  cudaFree(f->d_v); cudaFree(f->d_ge); cudaFree(f->d_u); cudaFree(f->d_flags);
  delete f;
  return HSBP_OK;
}

// dpsi_V = odefun(psi_delta, t)   (seas/BP1/odefun.jl:8-121); everything between the two small host copies
// runs on the device in one stream.
int hsbp_bp1_rhs(hsbp_bp1 *f, double t, const double *psi_delta, double *dpsi_V, hsbp_bp1_stats *stats) {
  if (!f) return HSBP_ERR_ARG;
  hsbp_blocks *b = f->blocks;
  hsbp_ctx *ctx = b->ctx;
  if (!psi_delta || !dpsi_V) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_bp1_rhs: null pointer");
  if (b->local_mode == 0) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_bp1_rhs: call hsbp_local_setup first");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  const int nf = f->nf;
  HSBP_CUDA(ctx, cudaMemcpyAsync(f->d_state, psi_delta, 2 * nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  HSBP_CUDA(ctx, cudaMemsetAsync(f->d_flags, 0, 4 * sizeof(int), ctx->stream));
  // boundary data -> ge = - sum_k F_k v_k   (locbcarray_mod!, global_curved.jl:569-592; Neumann data is zero)
  const int nmax = std::max(nf, f->nl);
  hsbp::k_bp1_bc<<<(nmax + 127) / 128, 128, 0, ctx->stream>>>(nf, f->off_fault, f->d_state, f->nl, f->off_load,
                                                              t * f->prm.Vp / 2.0, f->d_v);
  HSBP_CUDA(ctx, cudaMemsetAsync(f->d_ge, 0, (size_t)b->VNp * sizeof(double), ctx->stream));
  int rc = hsbp_face_F_add(b, f->d_v, -1.0, f->d_ge);
  if (rc) return rc;
  // u = M-tilde^-1 ge   (odefun.jl:43)
  hsbp_local_stats ls = {0, 0, 0, 0.0};
  if ((rc = hsbp_local_solve(b, f->d_ge, f->d_u, &ls))) return rc;
  // traction operator on all faces (small), then the fault stage
  if ((rc = hsbp_face_traction(b, f->d_u, b->d_fa))) return rc;
  hsbp::Bp1Dev dp;
  dp.mu_shear = f->prm.mu_shear; dp.sigma_n = f->prm.sigma_n; dp.eta = f->prm.eta; dp.V0 = f->prm.V0;
  dp.tau_z0 = f->prm.tau_z0; dp.Dc = f->prm.Dc; dp.f0 = f->prm.f0; dp.b = f->prm.b;
  dp.ftol = f->prm.ftol; dp.atolx = f->prm.atolx; dp.rtolx = f->prm.rtolx; dp.maxiter = (int)f->prm.maxiter;
  hsbp::k_bp1_fault<<<(nf + 127) / 128, 128, 0, ctx->stream>>>(nf, b->d_fa + f->off_fault, b->d_tau + f->off_fault,
                                                               f->d_sJ, f->d_a, f->d_state, f->d_out, dp, f->d_flags);
  cudaError_t e1 = cudaGetLastError();
  if (e1 != cudaSuccess) { ctx->err = std::string("k_bp1_fault: ") + cudaGetErrorString(e1); return HSBP_ERR_CUDA; }
  int flags[4] = {0, 0, 0, 0};
  HSBP_CUDA(ctx, cudaMemcpyAsync(dpsi_V, f->d_out, 2 * nf * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  HSBP_CUDA(ctx, cudaMemcpyAsync(flags, f->d_flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (stats) {
    stats->rejected = (flags[0] != 0 || ls.failed_blocks != 0) ? 1 : 0;
    stats->failure_bits = flags[0];
    stats->failed_nodes = flags[2];
    stats->newton_iterations_max = flags[1];
    stats->local_iterations = ls.iterations_max;
  }
  return HSBP_OK;
}

// displacement field of the last hsbp_bp1_rhs call (device -> host), e.g. for output
int hsbp_bp1_get_u(hsbp_bp1 *f, double *u) {
  if (!f || !u) return HSBP_ERR_ARG;
  return hsbp_d2h(f->blocks->ctx, u, f->d_u, (size_t)f->blocks->VNp * sizeof(double));
}

}  // extern "C"
