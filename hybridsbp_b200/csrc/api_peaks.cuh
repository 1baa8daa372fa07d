// Measured fp64 denominators of this device, next to the driver's HBM / bf16 numbers in MEASURED_PEAKS.json:
//   hsbp_peak_fp64_fma    sustained fp64 FMA rate of the CUDA cores (what bounds K1 besides HBM: ~ 100-230 flop per point)
//   hsbp_peak_fp64_dmma   sustained rate of mma.sync.m8n8k4.f64 issued from registers (the instruction K2a's trailing update,
//                         the banded factorisation and the Gauss-Jordan inversions use)
//   hsbp_peak_dgemm       cuBLAS DGEMM n x n x n -- a library number, used only as the denominator SURVEY.md section 8d names
// Benchmarks, not part of the solve path.  Included by hsbp.cu after api_fdm.cuh (cuBLAS handle).
#pragma once

namespace hsbp {

// 8 independent FMA chains per thread, `iters` rounds: 16 * iters flop per thread
__global__ void __launch_bounds__(256) k_peak_fma(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678) out[0] = s;               // keeps the chains alive
}
// 4 independent accumulator pairs per warp, `iters` rounds of m8n8k4: 2 * 8 * 8 * 4 = 512 flop per instruction
__global__ void __launch_bounds__(256) k_peak_dmma(double *out, int iters, double a, double b) {
  double c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const double av = a + threadIdx.x, bv = b - threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    dmma_m8n8k4(c[0], c[1], av, bv); dmma_m8n8k4(c[2], c[3], av, bv);
    dmma_m8n8k4(c[4], c[5], av, bv); dmma_m8n8k4(c[6], c[7], av, bv);
  }
  const double s = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
  if (s == 12345.678) out[0] = s;
}

}  // namespace hsbp

namespace {

template <class Launch> int peak_time(hsbp_ctx *ctx, Launch &&launch, double *ms_best) {
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  launch();                                     // warm-up
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double best = 1e300;
  for (int rep = 0; rep < 5; ++rep) {
    HSBP_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    launch();
    HSBP_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    HSBP_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float f = 0.f;
    HSBP_CUDA(ctx, cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
    best = std::min(best, (double)f);
  }
  HSBP_CUDA(ctx, cudaGetLastError());
  *ms_best = best;
  return HSBP_OK;
}

}  // namespace

extern "C" {

int hsbp_peak_fp64_fma(hsbp_ctx *ctx, double *tflops) {
  if (!ctx || !tflops) return HSBP_ERR_ARG;
  double *d = nullptr;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaMalloc((void **)&d, 8));
  const int iters = 20000, grid = ctx->sm_count * 8, threads = 256;
  double ms = 0;
  int rc = peak_time(ctx, [&]() { hsbp::k_peak_fma<<<grid, threads, 0, ctx->stream>>>(d, iters, 1.0000001, 1e-9); }, &ms);
  cudaFree(d);
  if (rc) return rc;
  *tflops = 16.0 * iters * (double)grid * threads / (ms * 1e-3) / 1e12;
  return HSBP_OK;
}

int hsbp_peak_fp64_dmma(hsbp_ctx *ctx, double *tflops) {
  if (!ctx || !tflops) return HSBP_ERR_ARG;
  double *d = nullptr;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaMalloc((void **)&d, 8));
  const int iters = 20000, grid = ctx->sm_count * 8, threads = 256;
  double ms = 0;
  int rc = peak_time(ctx, [&]() { hsbp::k_peak_dmma<<<grid, threads, 0, ctx->stream>>>(d, iters, 1.0000001, 1e-9); }, &ms);
  cudaFree(d);
  if (rc) return rc;
  *tflops = 4.0 * 512.0 * iters * (double)grid * (threads / 32) / (ms * 1e-3) / 1e12;
  return HSBP_OK;
}

int hsbp_peak_dgemm(hsbp_ctx *ctx, int64_t n, double *tflops) {
  if (!ctx || !tflops || n < 64 || n > 32768) return HSBP_ERR_ARG;
  FdmLibs *libs = nullptr;
  int rc = fdm_libs(ctx, &libs);
  if (rc) return rc;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  double *A = nullptr, *B = nullptr, *C = nullptr;
  const size_t bytes = (size_t)n * n * sizeof(double);
  HSBP_CUDA(ctx, cudaMalloc((void **)&A, bytes));
  if (cudaMalloc((void **)&B, bytes) != cudaSuccess || cudaMalloc((void **)&C, bytes) != cudaSuccess) {
    cudaFree(A); cudaFree(B); cudaFree(C);
    HSBP_FAIL(ctx, HSBP_ERR_CUDA, "hsbp_peak_dgemm: out of device memory");
  }
  hsbp::k_fill<<<1024, 256, 0, ctx->stream>>>(A, n * n, 1.0 / 3.0);
  hsbp::k_fill<<<1024, 256, 0, ctx->stream>>>(B, n * n, 1.0 / 7.0);
  const double one = 1.0, zero = 0.0;
  cublasStatus_t bs = CUBLAS_STATUS_SUCCESS;
  double ms = 0;
  rc = peak_time(ctx, [&]() {
    cublasStatus_t s_ = cublasDgemm(libs->blas, CUBLAS_OP_N, CUBLAS_OP_N, (int)n, (int)n, (int)n, &one, A, (int)n, B, (int)n, &zero, C, (int)n);
    if (s_ != CUBLAS_STATUS_SUCCESS) bs = s_;
  }, &ms);
  cudaFree(A); cudaFree(B); cudaFree(C);
  if (rc) return rc;
  if (bs != CUBLAS_STATUS_SUCCESS) HSBP_FAIL(ctx, HSBP_ERR_CUDA, "hsbp_peak_dgemm: cuBLAS DGEMM failed");
  *tflops = 2.0 * (double)n * n * n / (ms * 1e-3) / 1e12;
  return HSBP_OK;
}

}  // extern "C"
