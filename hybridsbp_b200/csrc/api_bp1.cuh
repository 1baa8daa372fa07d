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
  // condensed fault operator (hsbp_bp1_condense): Tf = HfI_FT_f M̃^-1 F_f, tl = HfI_FT_f M̃^-1 F_l 1
  double *d_Tf = nullptr, *d_tl = nullptr;
  bool condensed = false;
  double *h_io = nullptr, *d_io = nullptr;       // mapped host memory: [psi; delta | dpsi; V]
  int *h_flags = nullptr, *d_hflags = nullptr;   // mapped host memory: status words of the last launch
  std::vector<double> last_state;
  double last_t = 0.0;
  bool u_valid = false;
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
  if (e == cudaSuccess) e = cudaHostAlloc((void **)&f->h_io, 4 * f->nf * sizeof(double), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&f->d_io, f->h_io, 0);
  if (e == cudaSuccess) e = cudaHostAlloc((void **)&f->h_flags, 4 * sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&f->d_hflags, f->h_flags, 0);
  f->last_state.assign(2 * f->nf, 0.0);
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
  cudaFree(f->d_Tf); cudaFree(f->d_tl);
  if (f->h_io) cudaFreeHost(f->h_io);
  if (f->h_flags) cudaFreeHost(f->h_flags);
  delete f;
  return HSBP_OK;
}

}  // extern "C"

namespace {

hsbp::Bp1Dev bp1_dev_params(const hsbp_bp1 *f) {
  hsbp::Bp1Dev dp;
  dp.mu_shear = f->prm.mu_shear; dp.sigma_n = f->prm.sigma_n; dp.eta = f->prm.eta; dp.V0 = f->prm.V0;
  dp.tau_z0 = f->prm.tau_z0; dp.Dc = f->prm.Dc; dp.f0 = f->prm.f0; dp.b = f->prm.b;
  dp.ftol = f->prm.ftol; dp.atolx = f->prm.atolx; dp.rtolx = f->prm.rtolx; dp.maxiter = (int)f->prm.maxiter;
  return dp;
}

// boundary data on the device (d_state = [psi; delta]) -> ge = - sum_k F_k v_k -> u = M̃^-1 ge -> traction operator on all faces
// (locbcarray_mod!, global_curved.jl:569-592; odefun.jl:36-43; Neumann data is zero)
int bp1_displacement(hsbp_bp1 *f, double t, hsbp_local_stats *ls) {
  hsbp_blocks *b = f->blocks;
  hsbp_ctx *ctx = b->ctx;
  const int nf = f->nf, nmax = std::max(nf, f->nl);
  hsbp::k_bp1_bc<<<(nmax + 127) / 128, 128, 0, ctx->stream>>>(nf, f->off_fault, f->d_state, f->nl, f->off_load,
                                                              t * f->prm.Vp / 2.0, f->d_v);
  HSBP_CUDA(ctx, cudaMemsetAsync(f->d_ge, 0, (size_t)b->VNp * sizeof(double), ctx->stream));
  int rc = hsbp_face_F_add(b, f->d_v, -1.0, f->d_ge);
  if (rc) return rc;
  if ((rc = hsbp_local_solve(b, f->d_ge, f->d_u, ls))) return rc;
  f->u_valid = true;
  return hsbp_face_traction(b, f->d_u, b->d_fa);
}

}  // namespace

extern "C" {

// Condense the local solve onto the fault: nf + 1 local solves once, afterwards hsbp_bp1_rhs is one small kernel
// (k_bp1_fault_condensed).  enable = 0 returns to one local solve per call.
int hsbp_bp1_condense(hsbp_bp1 *f, int enable) {
  if (!f) return HSBP_ERR_ARG;
  hsbp_blocks *b = f->blocks;
  hsbp_ctx *ctx = b->ctx;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(f->d_Tf); cudaFree(f->d_tl); f->d_Tf = f->d_tl = nullptr;
  f->condensed = false;
  if (!enable) return HSBP_OK;
  if (b->local_mode == 0) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_bp1_condense: call hsbp_local_setup first");
  const int nf = f->nf;
  HSBP_CUDA(ctx, cudaMalloc((void **)&f->d_Tf, (size_t)nf * nf * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&f->d_tl, (size_t)nf * sizeof(double)));
  HSBP_CUDA(ctx, cudaMemsetAsync(f->d_v, 0, (size_t)b->FNp * sizeof(double), ctx->stream));
  int64_t failed = 0;
  int rc = HSBP_OK;
  for (int m = 0; m <= nf && rc == HSBP_OK; ++m) {
    // column m < nf: unit value at fault node m; column nf: ones on the loading face
    if (m > 0) hsbp::k_fill<<<1, 32, 0, ctx->stream>>>(f->d_v + f->off_fault + m - 1, 1, 0.0);
    if (m < nf) hsbp::k_fill<<<1, 32, 0, ctx->stream>>>(f->d_v + f->off_fault + m, 1, 1.0);
    else hsbp::k_fill<<<(f->nl + 255) / 256, 256, 0, ctx->stream>>>(f->d_v + f->off_load, f->nl, 1.0);
    HSBP_CUDA(ctx, cudaMemsetAsync(f->d_ge, 0, (size_t)b->VNp * sizeof(double), ctx->stream));
    if ((rc = hsbp_face_F_add(b, f->d_v, 1.0, f->d_ge))) break;
    hsbp_local_stats ls = {0, 0, 0, 0.0};
    if ((rc = hsbp_local_solve(b, f->d_ge, f->d_u, &ls))) break;
    failed += ls.failed_blocks;
    if ((rc = hsbp_face_traction(b, f->d_u, b->d_fa))) break;
    HSBP_CUDA(ctx, cudaMemcpyAsync(m < nf ? f->d_Tf + (size_t)nf * m : f->d_tl, b->d_fa + f->off_fault, nf * sizeof(double),
                                   cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if (rc == HSBP_OK) {
    cudaError_t e = cudaMemsetAsync(f->d_v, 0, (size_t)b->FNp * sizeof(double), ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("hsbp_bp1_condense: ") + cudaGetErrorString(e); rc = HSBP_ERR_CUDA; }
  }
  f->u_valid = false;
  if (rc == HSBP_OK && failed) { ctx->err = "hsbp_bp1_condense: a local solve did not reach its tolerance"; rc = HSBP_ERR_STATE; }
  if (rc != HSBP_OK) { cudaFree(f->d_Tf); cudaFree(f->d_tl); f->d_Tf = f->d_tl = nullptr; return rc; }
  f->condensed = true;
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
  const hsbp::Bp1Dev dp = bp1_dev_params(f);
  memcpy(f->last_state.data(), psi_delta, 2 * nf * sizeof(double));
  f->last_t = t;
  if (f->condensed) {
    // state and result travel through mapped host memory: one launch, one synchronisation, no copies
    memcpy(f->h_io, psi_delta, 2 * nf * sizeof(double));
    f->h_flags[0] = f->h_flags[1] = f->h_flags[2] = f->h_flags[3] = 0;
    hsbp::k_bp1_fault_condensed<<<(nf + 63) / 64, 64, 2 * nf * sizeof(double), ctx->stream>>>(
        nf, f->d_Tf, f->d_tl, t * f->prm.Vp / 2.0, b->d_tau + f->off_fault, f->d_sJ, f->d_a, f->d_io, f->d_io + 2 * nf, dp, f->d_hflags);
    cudaError_t e1 = cudaGetLastError();
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
    if (e1 != cudaSuccess) { ctx->err = std::string("k_bp1_fault_condensed: ") + cudaGetErrorString(e1); return HSBP_ERR_CUDA; }
    memcpy(dpsi_V, f->h_io + 2 * nf, 2 * nf * sizeof(double));
    f->u_valid = false;
    if (stats) {
      stats->rejected = f->h_flags[0] != 0 ? 1 : 0;
      stats->failure_bits = f->h_flags[0];
      stats->failed_nodes = f->h_flags[2];
      stats->newton_iterations_max = f->h_flags[1];
      stats->local_iterations = 0;
    }
    return HSBP_OK;
  }
  HSBP_CUDA(ctx, cudaMemcpyAsync(f->d_state, psi_delta, 2 * nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  HSBP_CUDA(ctx, cudaMemsetAsync(f->d_flags, 0, 4 * sizeof(int), ctx->stream));
  hsbp_local_stats ls = {0, 0, 0, 0.0};
  int rc = bp1_displacement(f, t, &ls);
  if (rc) return rc;
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

// displacement field of the last hsbp_bp1_rhs call (device -> host), e.g. for output; with the condensed fault operator
// the local solve is done here, on demand
int hsbp_bp1_get_u(hsbp_bp1 *f, double *u) {
  if (!f || !u) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = f->blocks->ctx;
  if (!f->u_valid) {
    HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
    HSBP_CUDA(ctx, cudaMemcpyAsync(f->d_state, f->last_state.data(), 2 * f->nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    hsbp_local_stats ls = {0, 0, 0, 0.0};
    int rc = bp1_displacement(f, f->last_t, &ls);
    if (rc) return rc;
  }
  return hsbp_d2h(ctx, u, f->d_u, (size_t)f->blocks->VNp * sizeof(double));
}

}  // extern "C"
