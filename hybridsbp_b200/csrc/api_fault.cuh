// C-ABI: rate-and-state fault stage on the fault nodes of a MULTIBLOCK mesh.
//
// The single-block benchmark (seas/BP1/odefun.jl) imposes slip through a Dirichlet face; on a multiblock mesh such as
// seas/BP1/meshes/BP1_v1.inp the fault is a set of jump interfaces: slip enters through the jump branch of locbcarray!
// (global_curved.jl:614-617, in_jump as in square_circle.jl:335-350), the displacement comes from the trace solve
// (square_circle.jl:376-388) and the shear stress from computetraction (global_curved.jl:638-644).  That chain is linear in
// (slip, time), so the stress change at the n fault nodes is   dtau = A delta + t b.   The host layer
// (hybridsbp_b200/bp1_multiblock.py) either forms A and b once with n + 1 trace solves (K4) and hands them over, or does one
// trace solve per right-hand-side evaluation and passes dtau; the per-node work -- bracketed Newton on rateandstate and the
// state evolution (global_curved.jl:1031-1075, odefun.jl:69-108) -- is the same device function as the single-block stage (K5).
#pragma once
#include "k_bp1.cuh"

struct hsbp_fault {
  hsbp_ctx *ctx = nullptr;
  int n = 0;
  hsbp_bp1_params prm;
  double *d_A = nullptr, *d_b = nullptr, *d_a = nullptr;
  double *h_io = nullptr, *d_io = nullptr;       // mapped host memory: [psi; delta | dpsi; V]
  int *h_flags = nullptr, *d_hflags = nullptr;
};

namespace {

hsbp::Bp1Dev fault_dev_params(const hsbp_bp1_params &p) {
  hsbp::Bp1Dev dp;
  dp.mu_shear = p.mu_shear; dp.sigma_n = p.sigma_n; dp.eta = p.eta; dp.V0 = p.V0; dp.tau_z0 = p.tau_z0; dp.Dc = p.Dc; dp.f0 = p.f0;
  dp.b = p.b; dp.ftol = p.ftol; dp.atolx = p.atolx; dp.rtolx = p.rtolx; dp.maxiter = (int)p.maxiter;
  return dp;
}

int fault_launch(hsbp_fault *f, double t, const double *dtau_dev, const double *psi_delta, double *dpsi_V, hsbp_bp1_stats *stats) {
  hsbp_ctx *ctx = f->ctx;
  const int n = f->n;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  memcpy(f->h_io, psi_delta, 2 * n * sizeof(double));
  f->h_flags[0] = f->h_flags[1] = f->h_flags[2] = f->h_flags[3] = 0;
  hsbp::k_fault_linear<<<(n + 63) / 64, 64, 2 * n * sizeof(double), ctx->stream>>>(n, f->d_A, f->d_b, t, dtau_dev, f->d_a, f->d_io,
                                                                                   f->d_io + 2 * n, fault_dev_params(f->prm), f->d_hflags);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { ctx->err = std::string("k_fault_linear: ") + cudaGetErrorString(e); return HSBP_ERR_CUDA; }
  memcpy(dpsi_V, f->h_io + 2 * n, 2 * n * sizeof(double));
  if (stats) {
    stats->rejected = f->h_flags[0] != 0 ? 1 : 0;
    stats->failure_bits = f->h_flags[0];
    stats->failed_nodes = f->h_flags[2];
    stats->newton_iterations_max = f->h_flags[1];
    stats->local_iterations = 0;
  }
  return HSBP_OK;
}

}  // namespace

extern "C" {

int hsbp_fault_destroy(hsbp_fault *f) {
  if (!f) return HSBP_ERR_ARG;
  cudaSetDevice(f->ctx->device);
  cudaStreamSynchronize(f->ctx->stream);
  cudaFree(f->d_A); cudaFree(f->d_b); cudaFree(f->d_a);
  if (f->h_io) cudaFreeHost(f->h_io);
  if (f->h_flags) cudaFreeHost(f->h_flags);
  delete f;
  return HSBP_OK;
}

// A (n x n, column-major) and b (n) may be NULL when every call brings its own dtau (hsbp_fault_stage)
int hsbp_fault_create(hsbp_ctx *ctx, int64_t n, const double *A, const double *b, const double *a, const hsbp_bp1_params *prm,
                      hsbp_fault **out) {
  if (!ctx || !out) return HSBP_ERR_ARG;
  *out = nullptr;
  if (n < 1 || n > 20000 || !a || !prm || ((A == nullptr) != (b == nullptr))) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_fault_create: bad arguments");
  hsbp_fault *f = new (std::nothrow) hsbp_fault();
  if (!f) HSBP_FAIL(ctx, HSBP_ERR_STATE, "out of host memory");
  f->ctx = ctx; f->n = (int)n; f->prm = *prm;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaMalloc((void **)&f->d_a, n * sizeof(double));
  if (e == cudaSuccess && A) e = cudaMalloc((void **)&f->d_A, (size_t)n * n * sizeof(double));
  if (e == cudaSuccess && b) e = cudaMalloc((void **)&f->d_b, n * sizeof(double));
  if (e == cudaSuccess) e = cudaHostAlloc((void **)&f->h_io, 4 * n * sizeof(double), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&f->d_io, f->h_io, 0);
  if (e == cudaSuccess) e = cudaHostAlloc((void **)&f->h_flags, 4 * sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&f->d_hflags, f->h_flags, 0);
  if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_a, a, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && A) e = cudaMemcpyAsync(f->d_A, A, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && b) e = cudaMemcpyAsync(f->d_b, b, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { ctx->err = std::string("hsbp_fault_create: ") + cudaGetErrorString(e); hsbp_fault_destroy(f); return HSBP_ERR_CUDA; }
  *out = f;
  return HSBP_OK;
}

// [dpsi/dt; V] at the fault nodes for the state [psi; delta] at time t, with dtau = A delta + t b
int hsbp_fault_rhs(hsbp_fault *f, double t, const double *psi_delta, double *dpsi_V, hsbp_bp1_stats *stats) {
  if (!f) return HSBP_ERR_ARG;
  if (!psi_delta || !dpsi_V) HSBP_FAIL(f->ctx, HSBP_ERR_ARG, "hsbp_fault_rhs: null pointer");
  if (!f->d_A) HSBP_FAIL(f->ctx, HSBP_ERR_STATE, "hsbp_fault_rhs: the fault was created without A and b (use hsbp_fault_stage)");
  return fault_launch(f, t, nullptr, psi_delta, dpsi_V, stats);
}

// the same with the stress change of this evaluation given in device memory (e.g. from a trace solve)
int hsbp_fault_stage(hsbp_fault *f, const double *dtau_dev, const double *psi_delta, double *dpsi_V, hsbp_bp1_stats *stats) {
  if (!f) return HSBP_ERR_ARG;
  if (!dtau_dev || !psi_delta || !dpsi_V) HSBP_FAIL(f->ctx, HSBP_ERR_ARG, "hsbp_fault_stage: null pointer");
  return fault_launch(f, 0.0, dtau_dev, psi_delta, dpsi_V, stats);
}

}  // extern "C"
