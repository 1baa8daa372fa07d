// Small dense fp64 building blocks of the trace preconditioner: batched in-place inversion of symmetric positive
// definite matrices (face blocks B_ff, the coarse matrices A_II and S_Gamma), so that applying a preconditioner is a
// bandwidth-bound matrix-vector product instead of two latency-bound triangular sweeps.
//
// Block Gauss-Jordan without pivoting (every Schur complement of an SPD matrix is SPD), panels of 32:
//   for every pivot tile k:   P = A_kk^-1
//                             A_kj <- P A_kj            (j != k)
//                             A_ij <- A_ij - A_ik A_kj  (i, j != k; rank-32 update on the fp64 tensor pipe, DMMA)
//                             A_ik <- -A_ik P           (i != k)
//                             A_kk <- P
// Storage as the dense Cholesky of the local solver (api_chol.cuh): column-major, leading dimension = size rounded up
// to 32, the pad being an identity block.
#pragma once
#include "api_chol.cuh"

namespace hsbp {

constexpr int GJ_NB = 32;
constexpr int GJ_THREADS = 256;

// P = A_kk^-1 for every matrix of the batch: written to A_kk and to Pw[batch][32 x 32]
__global__ void __launch_bounds__(GJ_THREADS)
k_gj_pivot(const CholBlock *__restrict__ cb, double *__restrict__ Aall, int k0, double *__restrict__ Pw, int *__restrict__ flag) {
  __shared__ double M[GJ_NB][GJ_NB + 1];
  __shared__ double colj[GJ_NB];
  __shared__ double pinv;
  const CholBlock b = cb[blockIdx.x];
  if (k0 >= b.ld) return;
  double *A = Aall + b.off;
  const int ld = b.ld, tid = threadIdx.x;
  for (int idx = tid; idx < GJ_NB * GJ_NB; idx += GJ_THREADS) {
    const int i = idx % GJ_NB, j = idx / GJ_NB;
    M[i][j] = A[(int64_t)(k0 + j) * ld + k0 + i];
  }
  __syncthreads();
  for (int j = 0; j < GJ_NB; ++j) {
    if (tid < GJ_NB) colj[tid] = M[tid][j];
    if (tid == 0) {
      const double piv = M[j][j];
      if (!(piv > 0.0)) flag[blockIdx.x] = 1;
      pinv = 1.0 / piv;
    }
    __syncthreads();
    if (tid < GJ_NB) M[j][tid] = (tid == j ? 1.0 : M[j][tid]) * pinv;
    __syncthreads();
    for (int idx = tid; idx < GJ_NB * GJ_NB; idx += GJ_THREADS) {
      const int i = idx % GJ_NB, c = idx / GJ_NB;
      if (i != j) M[i][c] = (c == j ? 0.0 : M[i][c]) - colj[i] * M[j][c];
    }
    __syncthreads();
  }
  double *P = Pw + (int64_t)blockIdx.x * GJ_NB * GJ_NB;
  for (int idx = tid; idx < GJ_NB * GJ_NB; idx += GJ_THREADS) {
    const int i = idx % GJ_NB, j = idx / GJ_NB;
    A[(int64_t)(k0 + j) * ld + k0 + i] = M[i][j];
    P[idx] = M[i][j];
  }
}

// grid = (tiles, batch): tile t != k handles the column tile A_tk (saved to Cw, replaced by -A_tk P) and the row
// tile A_kt (replaced by P A_kt).  Cw: ld x 32 per matrix at 32 * woff.
__global__ void __launch_bounds__(GJ_THREADS)
k_gj_panels(const CholBlock *__restrict__ cb, double *__restrict__ Aall, int k0, const double *__restrict__ Pw,
            double *__restrict__ Cw) {
  __shared__ double P[GJ_NB][GJ_NB + 1];
  __shared__ double X[GJ_NB][GJ_NB + 1];
  const CholBlock b = cb[blockIdx.y];
  const int t0 = blockIdx.x * GJ_NB;
  if (k0 >= b.ld || t0 >= b.ld || t0 == k0) return;
  double *A = Aall + b.off;
  const int ld = b.ld, tid = threadIdx.x;
  const double *Pg = Pw + (int64_t)blockIdx.y * GJ_NB * GJ_NB;
  double *C = Cw + (int64_t)b.woff * GJ_NB + (int64_t)blockIdx.x * GJ_NB * GJ_NB;
  for (int idx = tid; idx < GJ_NB * GJ_NB; idx += GJ_THREADS) {
    const int i = idx % GJ_NB, j = idx / GJ_NB;
    P[i][j] = Pg[idx];
    const double x = A[(int64_t)(k0 + j) * ld + t0 + i];     // column tile A_tk
    X[i][j] = x;
    C[idx] = x;
  }
  __syncthreads();
  for (int idx = tid; idx < GJ_NB * GJ_NB; idx += GJ_THREADS) {
    const int i = idx % GJ_NB, j = idx / GJ_NB;
    double s = 0.0;
#pragma unroll 8
    for (int m = 0; m < GJ_NB; ++m) s += X[i][m] * P[m][j];
    A[(int64_t)(k0 + j) * ld + t0 + i] = -s;
  }
  __syncthreads();
  for (int idx = tid; idx < GJ_NB * GJ_NB; idx += GJ_THREADS) {
    const int i = idx % GJ_NB, j = idx / GJ_NB;
    X[i][j] = A[(int64_t)(t0 + j) * ld + k0 + i];            // row tile A_kt
  }
  __syncthreads();
  for (int idx = tid; idx < GJ_NB * GJ_NB; idx += GJ_THREADS) {
    const int i = idx % GJ_NB, j = idx / GJ_NB;
    double s = 0.0;
#pragma unroll 8
    for (int m = 0; m < GJ_NB; ++m) s += P[i][m] * X[m][j];
    A[(int64_t)(t0 + j) * ld + k0 + i] = s;
  }
}

// grid = (tiles, tiles, batch): A_ij -= C_i (A_kj), i, j != k, C_i the saved column tile
__global__ void __launch_bounds__(GJ_THREADS)
k_gj_update(const CholBlock *__restrict__ cb, double *__restrict__ Aall, int k0, const double *__restrict__ Cw) {
  __shared__ double Ci[GJ_NB][GJ_NB + 1];          // [row][m]
  __shared__ double Rj[GJ_NB][GJ_NB + 1];          // [col][m]  (transposed so that both operands are read along m)
  const CholBlock b = cb[blockIdx.z];
  const int i0 = blockIdx.x * GJ_NB, j0 = blockIdx.y * GJ_NB;
  if (k0 >= b.ld || i0 >= b.ld || j0 >= b.ld || i0 == k0 || j0 == k0) return;
  double *A = Aall + b.off;
  const int ld = b.ld, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const double *C = Cw + (int64_t)b.woff * GJ_NB + (int64_t)blockIdx.x * GJ_NB * GJ_NB;
  for (int idx = tid; idx < GJ_NB * GJ_NB; idx += GJ_THREADS) {
    const int i = idx % GJ_NB, m = idx / GJ_NB;
    Ci[i][m] = C[idx];
    Rj[m][i] = A[(int64_t)(j0 + m) * ld + k0 + i];           // A_kj[i = row in the pivot tile][m = column]  ->  Rj[col][row]
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int st = wid * 2 + s, si = (st >> 2) * 8, sj = (st & 3) * 8;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int k = 0; k < GJ_NB; k += 4)
      dmma_m8n8k4(c0, c1, Ci[si + (lane >> 2)][k + (lane & 3)], Rj[sj + (lane >> 2)][k + (lane & 3)]);
    const int gi = i0 + si + (lane >> 2);
    const int gj = j0 + sj + (lane & 3) * 2;
    double *c = A + (int64_t)gj * ld + gi;
    c[0] -= c0;
    c[ld] -= c1;
  }
}

// A <- (A + A^T) / 2 on the np x np part (grid = (slices, batch))
__global__ void __launch_bounds__(256)
k_dense_sym(const CholBlock *__restrict__ cb, double *__restrict__ Aall) {
  const CholBlock b = cb[blockIdx.y];
  double *A = Aall + b.off;
  const int64_t n2 = (int64_t)b.np * b.np;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n2; idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx % b.np), c = (int)(idx / b.np);
    if (r <= c) continue;
    const double v = 0.5 * (A[r + (int64_t)b.ld * c] + A[c + (int64_t)b.ld * r]);
    A[r + (int64_t)b.ld * c] = v;
    A[c + (int64_t)b.ld * r] = v;
  }
}

// identity on the pad of one ld x ld matrix whose np x np part was filled elsewhere (the rest of the pad is zero)
__global__ void k_dense_pad_identity(double *__restrict__ A, int np, int ld) {
  const int i = np + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ld) A[i + (int64_t)ld * i] = 1.0;
}

// C (m x n, ldc) = A (m x k, lda; symmetric part is not assumed) * B (k x n, ldb); setup only, one thread per row of a column
__global__ void __launch_bounds__(256)
k_dense_gemm_nn(int m, int n, int k, const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb,
                double *__restrict__ C, int ldc) {
  const int j = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m || j >= n) return;
  double s = 0.0;
  for (int l = 0; l < k; ++l) s += A[i + (int64_t)lda * l] * B[l + (int64_t)ldb * j];
  C[i + (int64_t)ldc * j] = s;
}
// C (m x n) = C0 - A^T B with A (k x m, lda), B (k x n, ldb): one warp per entry
__global__ void __launch_bounds__(256)
k_dense_sub_atb(int m, int n, int k, const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb,
                const double *__restrict__ C0, double *__restrict__ C, int ldc) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)m * n) return;
  const int i = (int)(w % m), j = (int)(w / m);
  double s = 0.0;
  for (int l = lane; l < k; l += 32) s += A[l + (int64_t)lda * i] * B[l + (int64_t)ldb * j];
  s = warp_sum(s);
  if (lane == 0) C[i + (int64_t)ldc * j] = C0[i + (int64_t)ldc * j] - s;
}
// B (n x m) = A^T for A (m x n)
__global__ void k_dense_transpose(int m, int n, const double *__restrict__ A, double *__restrict__ B) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)m * n) return;
  const int i = (int)(idx % m), j = (int)(idx / m);
  B[j + (int64_t)n * i] = A[idx];
}

}  // namespace hsbp

namespace {

using namespace hsbp;

// In-place inverse of every SPD matrix of the batch (descriptors on the device in d_cb, copies in h_cb); symmetrised.
// Pw / Cw are scratch of 1024 doubles per matrix and 32 * sum(ld) doubles.  Returns HSBP_ERR_ARG with a message if
// a pivot is not positive.
int dense_spd_inverse_batched(hsbp_ctx *ctx, const std::vector<CholBlock> &h_cb, const CholBlock *d_cb, double *A, const char *what) {
  const int64_t nb = (int64_t)h_cb.size();
  if (nb == 0) return HSBP_OK;
  int maxld = 0;
  int64_t wsum = 0;
  for (const CholBlock &c : h_cb) { maxld = std::max(maxld, (int)c.ld); wsum = std::max<int64_t>(wsum, c.woff + c.ld); }
  double *Pw = nullptr, *Cw = nullptr;
  int *d_flag = nullptr;
  HSBP_CUDA(ctx, cudaMalloc((void **)&Pw, (size_t)nb * GJ_NB * GJ_NB * sizeof(double)));
  cudaError_t e = cudaMalloc((void **)&Cw, (size_t)wsum * GJ_NB * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_flag, nb * sizeof(int));
  if (e == cudaSuccess) e = cudaMemsetAsync(d_flag, 0, nb * sizeof(int), ctx->stream);
  if (e != cudaSuccess) { cudaFree(Pw); cudaFree(Cw); cudaFree(d_flag); ctx->err = std::string(what) + ": " + cudaGetErrorString(e); return HSBP_ERR_CUDA; }
  const int nt = maxld / GJ_NB;
  // the batch index is a grid dimension: y / z are limited to 65535
  for (int64_t b0 = 0; b0 < nb; b0 += 32768) {
    const unsigned cnt = (unsigned)std::min<int64_t>(32768, nb - b0);
    for (int k0 = 0; k0 < maxld; k0 += GJ_NB) {
      k_gj_pivot<<<cnt, GJ_THREADS, 0, ctx->stream>>>(d_cb + b0, A, k0, Pw + b0 * GJ_NB * GJ_NB, d_flag + b0);
      if (nt > 1) {
        k_gj_panels<<<dim3(nt, cnt), GJ_THREADS, 0, ctx->stream>>>(d_cb + b0, A, k0, Pw + b0 * GJ_NB * GJ_NB, Cw);
        k_gj_update<<<dim3(nt, nt, cnt), GJ_THREADS, 0, ctx->stream>>>(d_cb + b0, A, k0, Cw);
      }
    }
    k_dense_sym<<<dim3((unsigned)std::max(1, std::min(64, maxld * maxld / 4096)), cnt), 256, 0, ctx->stream>>>(d_cb + b0, A);
  }
  std::vector<int> flag(nb, 0);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(flag.data(), d_flag, nb * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(Pw); cudaFree(Cw); cudaFree(d_flag);
  if (e != cudaSuccess) { ctx->err = std::string(what) + ": " + cudaGetErrorString(e); return HSBP_ERR_CUDA; }
  for (int64_t i = 0; i < nb; ++i)
    if (flag[i]) HSBP_FAIL(ctx, HSBP_ERR_ARG, std::string(what) + ": matrix is not positive definite");
  return HSBP_OK;
}

}  // namespace
