// C-ABI: the reference's `factorization` plugin at its own seam.
//
// reference: SBPLocalOperator1(lop, Nr, Ns, factorization) calls `factorization(lop[e].M̃)` on the ASSEMBLED sparse
// matrix of every block -- after probing the plugin with a 1 x 1 matrix to learn the factor type (global_curved.jl:681,
// 698; supplied as x -> cholesky(Symmetric(x)) at square_circle.jl:299, seas/BP1/BP1.jl:78) -- and afterwards only uses
// `F \ g` (:734, square_circle.jl:383, odefun.jl:43) and `F' \ S` with a sparse right-hand side block (:774).
// hsbp_factor is that object for a host that keeps the reference's assembly untouched: it takes the CSC arrays of
// one symmetric positive definite matrix, stores its lower band (with points numbered r-fastest M̃_e is banded,
// api_band.cuh) and factorises it with the banded Cholesky kernels (DMMA trailing update); solves are the streamed
// banded sweeps.  Any size from the 1 x 1 probe upwards.
#pragma once
#include "api_band.cuh"

struct hsbp_factor {
  hsbp_ctx *ctx = nullptr;
  int64_t n = 0;
  hsbp::BandBlock bb;
  hsbp::BandBlock *d_bb = nullptr;
  double *d_band = nullptr, *d_work = nullptr, *d_inv = nullptr, *d_g = nullptr, *d_x = nullptr;
  int64_t rhs_cap = 0;
  int stream_stages = 0;
};

namespace {

int factor_solve_dev(hsbp_factor *f, const double *g, double *x) {
  hsbp_ctx *ctx = f->ctx;
  using namespace hsbp;
  if (f->stream_stages >= 2) {
    const int nst = f->stream_stages;
    const size_t sm = (size_t)nst * BS_PB * f->bb.ld * 8 + (size_t)nst * BS_PB * BS_PB * 8 + (size_t)BS_WIN * 8 + 2 * BS_PB * 8 +
                      nst * sizeof(uint64_t);
    HSBP_CUDA(ctx, hsbp_smem_optin(ctx, k_band_solve_stream<BS_PB>, sm));
    k_band_solve_stream<BS_PB><<<1, BS_THREADS, sm, ctx->stream>>>(f->d_bb, f->d_band, f->d_inv, g, x, f->d_work, nst, f->bb.ld);
  } else {
    k_band_solve<<<1, CH_THREADS, 0, ctx->stream>>>(f->d_bb, f->d_band, g, x, f->d_work);
  }
  return check_launch(ctx, "hsbp_factor solve");
}

}  // namespace

extern "C" {

int hsbp_factor_destroy(hsbp_factor *f) {
  if (!f) return HSBP_ERR_ARG;
  cudaSetDevice(f->ctx->device);
  cudaStreamSynchronize(f->ctx->stream);
  cudaFree(f->d_bb); cudaFree(f->d_band); cudaFree(f->d_work); cudaFree(f->d_inv); cudaFree(f->d_g); cudaFree(f->d_x);
  delete f;
  return HSBP_OK;
}

// colptr (n + 1), rowval, nzval: compressed sparse columns with `index_base` (1 for arrays straight from the reference's
// host language, 0 for C-style); only the lower triangle (row >= column) is read.
int hsbp_factor_create(hsbp_ctx *ctx, int64_t n, const int64_t *colptr, const int64_t *rowval, const double *nzval, int index_base,
                       hsbp_factor **out) {
  using namespace hsbp;
  if (!ctx || !out) return HSBP_ERR_ARG;
  *out = nullptr;
  if (n < 1 || n > (1 << 30) || !colptr || !rowval || !nzval || (index_base != 0 && index_base != 1))
    HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_factor_create: bad arguments");
  int64_t kd = 0;
  for (int64_t c = 0; c < n; ++c)
    for (int64_t k = colptr[c] - index_base; k < colptr[c + 1] - index_base; ++k) {
      const int64_t r = rowval[k] - index_base;
      if (r < 0 || r >= n) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_factor_create: row index out of range");
      if (r >= c) kd = std::max(kd, r - c);
    }
  hsbp_factor *f = new (std::nothrow) hsbp_factor();
  if (!f) HSBP_FAIL(ctx, HSBP_ERR_STATE, "out of host memory");
  f->ctx = ctx; f->n = n;
  BandBlock &q = f->bb;
  memset(&q, 0, sizeof(q));
  q.np = (int32_t)n; q.npad = (int32_t)((n + CH_NB - 1) / CH_NB * CH_NB);
  q.kd = (int32_t)kd; q.ld = (int32_t)((kd + 1 + 31) / 32 * 32);
  q.Nrp = (int32_t)n; q.Nsp = 1; q.off = 0; q.voff = 0; q.woff = 0; q.ioff = 0;
  const size_t nb = (size_t)q.npad * q.ld;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  if (nb * sizeof(double) > free_b / 2) { delete f; HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "hsbp_factor_create: the band does not fit in device memory"); }
  std::vector<double> AB;
  try { AB.assign(nb, 0.0); } catch (...) { delete f; HSBP_FAIL(ctx, HSBP_ERR_STATE, "out of host memory"); }
  for (int64_t c = n; c < q.npad; ++c) AB[(size_t)c * q.ld] = 1.0;          // identity on the pad
  for (int64_t c = 0; c < n; ++c)
    for (int64_t k = colptr[c] - index_base; k < colptr[c + 1] - index_base; ++k) {
      const int64_t r = rowval[k] - index_base;
      if (r >= c) AB[(size_t)c * q.ld + (size_t)(r - c)] += nzval[k];          // duplicates are summed, as sparse() does
    }
  cudaError_t e = cudaSuccess;
  auto A = [&](void **p_, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p_, std::max<size_t>(bytes, 8)); };
  A((void **)&f->d_bb, sizeof(BandBlock)); A((void **)&f->d_band, nb * sizeof(double));
  A((void **)&f->d_work, (size_t)q.npad * sizeof(double)); A((void **)&f->d_inv, (size_t)q.npad * BS_PB * sizeof(double));
  int *d_flag = nullptr;
  A((void **)&d_flag, sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_bb, &q, sizeof(BandBlock), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_band, AB.data(), nb * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream);
  int flag = 0;
  if (e == cudaSuccess) {
    const int ntmax = (q.kd + CH_NB - 1) / CH_NB + 1;
    for (int k0 = 0; k0 < q.npad; k0 += CH_NB) {
      k_band_panel<<<1, CH_THREADS, 0, ctx->stream>>>(f->d_bb, f->d_band, k0, d_flag);
      const int nt = std::min(ntmax, (q.npad - k0 - CH_NB) / CH_NB);
      if (nt > 0) k_band_update<<<dim3(nt, nt, 1), CH_THREADS, 0, ctx->stream>>>(f->d_bb, f->d_band, k0);
    }
    k_band_invdiag<BS_PB><<<dim3((unsigned)(q.npad / BS_PB), 1), BS_PB, 0, ctx->stream>>>(f->d_bb, f->d_band, f->d_inv);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(d_flag);
  if (e != cudaSuccess) { ctx->err = std::string("hsbp_factor_create: ") + cudaGetErrorString(e); hsbp_factor_destroy(f); return HSBP_ERR_CUDA; }
  if (flag) { hsbp_factor_destroy(f); HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_factor_create: the matrix is not positive definite"); }
  {
    const size_t fixed = (size_t)BS_WIN * 8 + 2 * BS_PB * 8 + 64;
    const size_t per_stage = (size_t)BS_PB * q.ld * 8 + (size_t)BS_PB * BS_PB * 8;
    const int nst = (int)std::min<size_t>(4, (ctx->smem_optin > fixed ? (ctx->smem_optin - fixed) / per_stage : 0));
    f->stream_stages = (nst >= 2 && q.kd >= BS_PB && q.kd + 2 * BS_PB <= BS_WIN) ? nst : 0;
  }
  *out = f;
  return HSBP_OK;
}

int64_t hsbp_factor_size(const hsbp_factor *f) { return f ? f->n : -1; }

// u = A^-1 g for nrhs right-hand sides stored one after the other (host arrays, leading dimension n): `F \ g`
int hsbp_factor_solve(hsbp_factor *f, const double *g, double *u, int64_t nrhs) {
  if (!f) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = f->ctx;
  if (!g || !u || nrhs < 1) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_factor_solve: bad arguments");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  if (f->rhs_cap < f->n) {
    cudaFree(f->d_g); cudaFree(f->d_x); f->d_g = f->d_x = nullptr; f->rhs_cap = 0;
    HSBP_CUDA(ctx, cudaMalloc((void **)&f->d_g, (size_t)f->n * sizeof(double)));
    HSBP_CUDA(ctx, cudaMalloc((void **)&f->d_x, (size_t)f->n * sizeof(double)));
    f->rhs_cap = f->n;
  }
  for (int64_t k = 0; k < nrhs; ++k) {
    HSBP_CUDA(ctx, cudaMemcpyAsync(f->d_g, g + k * f->n, (size_t)f->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    int rc = factor_solve_dev(f, f->d_g, f->d_x);
    if (rc) return rc;
    HSBP_CUDA(ctx, cudaMemcpyAsync(u + k * f->n, f->d_x, (size_t)f->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return HSBP_OK;
}

int hsbp_factor_solve_dev(hsbp_factor *f, const double *g_dev, double *u_dev) {
  if (!f) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = f->ctx;
  if (!g_dev || !u_dev || g_dev == u_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_factor_solve_dev: bad pointers");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  return factor_solve_dev(f, g_dev, u_dev);
}

}  // extern "C"
