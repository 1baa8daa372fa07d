// K2a: batched dense fp64 Cholesky local solver for small blocks.
//
// Replaces the reference's per-block `cholesky(Symmetric(M-tilde_e))` + `F \ g`
// (global_curved.jl:698, 734; plugin at square_circle.jl:299, BP1.jl:78) where a dense factor fits:
//   setup   dense M-tilde_e is formed on the device by applying the matrix-free operator to unit vectors
//           (one apply per column index, all blocks at once), then factorised in place, one CTA per block,
//           right-looking in panels of 32: per panel one kernel for the diagonal block (CUDA cores, shared memory)
//           and the panel TRSM (one row per thread), and one kernel with a CTA per 32 x 32 tile of the trailing matrix
//           of every block for C -= L21 L21^T on the fp64 tensor pipe (mma.sync m8n8k4 f64, DMMA)
//           -- the only place tensor cores are used, as BASELINE.json's north star asks.
//   solve   L y = g, L^T x = y, one CTA per block, panels of 32 (column reads are coalesced).
// Storage: block e at chol_off[e], column-major, leading dimension ld_e = Np_e rounded up to 32 (the pad is an
// identity block, so every panel is full).
#pragma once
#include "k_solve.cuh"

namespace hsbp {

constexpr int CH_NB = 32;          // panel width
constexpr int CH_THREADS = 256;

struct CholBlock {
  int64_t off;     // offset of the dense matrix
  int32_t np, ld;  // true size, padded size / leading dimension
  int64_t voff;    // offset of the block in volume vectors
  int64_t woff;    // offset of the block's padded work vector
};

// unit vectors: u = e_c in every block that has a point c (0 elsewhere: the vector is cleared by the caller once
// and the previous unit entry is removed here)
__global__ void k_chol_unit(const CholBlock *__restrict__ cb, int64_t nblocks, int c, double *__restrict__ u) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nblocks) return;
  const CholBlock b = cb[e];
  if (c > 0 && c - 1 < b.np) u[b.voff + c - 1] = 0.0;
  if (c < b.np) u[b.voff + c] = 1.0;
}
// column c of every dense matrix <- y_e ; pad rows / columns get the identity
__global__ void k_chol_store_col(const CholBlock *__restrict__ cb, int c, const double *__restrict__ y, double *__restrict__ A) {
  const CholBlock b = cb[blockIdx.x];
  if (c >= b.ld) return;
  double *col = A + b.off + (int64_t)c * b.ld;
  for (int i = threadIdx.x; i < b.ld; i += blockDim.x)
    col[i] = (c < b.np && i < b.np) ? y[b.voff + i] : (i == c ? 1.0 : 0.0);
}

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// One panel step of the in-place lower Cholesky of every block (right-looking, panel width 32):
//   k_chol_panel   diagonal block (CUDA cores, shared memory) and the panel below it (X L^T = A21, one row per thread)
//   k_chol_update  trailing update C[ti][tj] -= L21[ti] L21[tj]^T, one CTA per 32 x 32 tile of the lower triangle of
//                  every block, on the fp64 tensor pipe (mma.sync m8n8k4 f64)
// flag[e] = 1 if a pivot was not positive.
__global__ void __launch_bounds__(CH_THREADS)
k_chol_panel(const CholBlock *__restrict__ cb, double *__restrict__ Aall, int k0, int *__restrict__ flag) {
  __shared__ double D[CH_NB][CH_NB + 1];            // diagonal block / its factor
  const CholBlock b = cb[blockIdx.x];
  if (k0 >= b.ld) return;
  double *A = Aall + b.off;
  const int ld = b.ld, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  bool bad = false;
  for (int idx = tid; idx < CH_NB * CH_NB; idx += CH_THREADS) {
    const int i = idx % CH_NB, j = idx / CH_NB;
    D[i][j] = A[(int64_t)(k0 + j) * ld + k0 + i];
  }
  __syncthreads();
  if (wid == 0) {
    for (int j = 0; j < CH_NB; ++j) {
      const double djj = D[j][j];
      if (!(djj > 0.0)) bad = true;
      const double l = sqrt(djj);
      __syncwarp();
      if (lane >= j) D[lane][j] = (lane == j) ? l : D[lane][j] / l;
      __syncwarp();
      for (int c = j + 1; c < CH_NB; ++c)
        if (lane >= c) D[lane][c] -= D[lane][j] * D[c][j];
      __syncwarp();
    }
    if (bad) flag[blockIdx.x] = 1;
  }
  __syncthreads();
  for (int idx = tid; idx < CH_NB * CH_NB; idx += CH_THREADS) {
    const int i = idx % CH_NB, j = idx / CH_NB;
    A[(int64_t)(k0 + j) * ld + k0 + i] = (i >= j) ? D[i][j] : 0.0;
  }
  for (int i = k0 + CH_NB + tid; i < ld; i += CH_THREADS) {
    double x[CH_NB];
#pragma unroll
    for (int j = 0; j < CH_NB; ++j) x[j] = A[(int64_t)(k0 + j) * ld + i];
#pragma unroll
    for (int j = 0; j < CH_NB; ++j) {
      double s = x[j];
#pragma unroll
      for (int c = 0; c < j; ++c) s -= x[c] * D[j][c];
      x[j] = s / D[j][j];
    }
#pragma unroll
    for (int j = 0; j < CH_NB; ++j) A[(int64_t)(k0 + j) * ld + i] = x[j];
  }
}

// grid = (tiles, tiles, blocks); tile (ti, tj), tj <= ti, of the trailing matrix behind panel k0
__global__ void __launch_bounds__(CH_THREADS)
k_chol_update(const CholBlock *__restrict__ cb, double *__restrict__ Aall, int k0) {
  __shared__ double Ti[CH_NB][CH_NB + 1];           // L21 tiles, [row][k]
  __shared__ double Tj[CH_NB][CH_NB + 1];
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (tj > ti) return;
  const CholBlock b = cb[blockIdx.z];
  const int ld = b.ld, m0 = k0 + CH_NB;
  if (m0 + (ti + 1) * CH_NB > ld) return;
  double *A = Aall + b.off;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int idx = tid; idx < CH_NB * CH_NB; idx += CH_THREADS) {
    const int i = idx % CH_NB, k = idx / CH_NB;
    Ti[i][k] = A[(int64_t)(k0 + k) * ld + m0 + ti * CH_NB + i];
    Tj[i][k] = A[(int64_t)(k0 + k) * ld + m0 + tj * CH_NB + i];
  }
  __syncthreads();
  // 16 sub-tiles of 8 x 8, two per warp
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int st = wid * 2 + s, si = (st >> 2) * 8, sj = (st & 3) * 8;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int k = 0; k < CH_NB; k += 4)
      dmma_m8n8k4(c0, c1, Ti[si + (lane >> 2)][k + (lane & 3)], Tj[sj + (lane >> 2)][k + (lane & 3)]);
    const int gi = m0 + ti * CH_NB + si + (lane >> 2);
    const int gj = m0 + tj * CH_NB + sj + (lane & 3) * 2;
    double *c = A + (int64_t)gj * ld + gi;
    c[0] -= c0;
    c[ld] -= c1;
  }
}

// x_e = (L L^T)^-1 g_e for every block; work: one padded vector per block (CholBlock::woff)
__global__ void __launch_bounds__(CH_THREADS)
k_chol_solve(const CholBlock *__restrict__ cb, const double *__restrict__ Aall, const double *__restrict__ g,
             double *__restrict__ x, double *__restrict__ work) {
  __shared__ double D[CH_NB][CH_NB + 1];
  __shared__ double xb[CH_NB];
  const CholBlock b = cb[blockIdx.x];
  const double *A = Aall + b.off;
  double *r = work + b.woff;
  const int ld = b.ld, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < ld; i += CH_THREADS) r[i] = i < b.np ? g[b.voff + i] : 0.0;
  __syncthreads();
  // ---- L y = g ---------------------------------------------------------------------------------
  for (int k0 = 0; k0 < ld; k0 += CH_NB) {
    for (int idx = tid; idx < CH_NB * CH_NB; idx += CH_THREADS) {
      const int i = idx % CH_NB, j = idx / CH_NB;
      D[i][j] = A[(int64_t)(k0 + j) * ld + k0 + i];
    }
    __syncthreads();
    if (wid == 0) {
      double v = r[k0 + lane];
      for (int j = 0; j < CH_NB; ++j) {
        const double yj = __shfl_sync(0xffffffffu, v, j) / D[j][j];
        if (lane == j) v = yj;
        if (lane > j) v -= D[lane][j] * yj;
      }
      xb[lane] = v;
      r[k0 + lane] = v;
    }
    __syncthreads();
    for (int i = k0 + CH_NB + tid; i < ld; i += CH_THREADS) {
      double s = r[i];
#pragma unroll 8
      for (int j = 0; j < CH_NB; ++j) s -= A[(int64_t)(k0 + j) * ld + i] * xb[j];
      r[i] = s;
    }
    __syncthreads();
  }
  // ---- L^T x = y -------------------------------------------------------------------------------
  for (int k0 = ld - CH_NB; k0 >= 0; k0 -= CH_NB) {
    for (int c = wid; c < CH_NB; c += CH_THREADS / 32) {      // s_c = L[k0+32.., k0+c] . x[k0+32..]
      const double *col = A + (int64_t)(k0 + c) * ld;
      double s = 0.0;
      for (int i = k0 + CH_NB + lane; i < ld; i += 32) s += col[i] * r[i];
      s = warp_sum(s);
      if (lane == 0) xb[c] = r[k0 + c] - s;
    }
    for (int idx = tid; idx < CH_NB * CH_NB; idx += CH_THREADS) {
      const int i = idx % CH_NB, j = idx / CH_NB;
      D[i][j] = A[(int64_t)(k0 + j) * ld + k0 + i];
    }
    __syncthreads();
    if (wid == 0) {
      double v = xb[lane];
      for (int j = CH_NB - 1; j >= 0; --j) {
        const double xj = __shfl_sync(0xffffffffu, v, j) / D[j][j];
        if (lane == j) v = xj;
        if (lane < j) v -= D[j][lane] * xj;
      }
      r[k0 + lane] = v;
    }
    __syncthreads();
  }
  for (int i = tid; i < b.np; i += CH_THREADS) x[b.voff + i] = r[i];
}

}  // namespace hsbp

namespace {

using namespace hsbp;

int chol_setup(hsbp_blocks *b) {
  hsbp_ctx *ctx = b->ctx;
  // sizes
  std::vector<CholBlock> cbs(b->nblocks);
  int64_t off = 0, woff = 0;
  int maxld = 0;
  for (int64_t e = 0; e < b->nblocks; ++e) {
    const BlockDesc &d = b->h_desc[e];
    const int64_t np = (int64_t)(d.Nr + 1) * (d.Ns + 1);
    if (np > 8192) HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "dense Cholesky local solver: block too large (use HSBP_LOCAL_PCG)");
    const int ld = (int)((np + CH_NB - 1) / CH_NB * CH_NB);
    cbs[e].off = off; cbs[e].woff = woff; cbs[e].np = (int)np; cbs[e].ld = ld; cbs[e].voff = d.voff;
    off += (int64_t)ld * ld; woff += ld;
    maxld = std::max(maxld, ld);
  }
  size_t free_b = 0, total_b = 0;
  HSBP_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  if ((size_t)off * sizeof(double) > free_b / 2)
    HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "dense Cholesky local solver: factors do not fit in device memory (use HSBP_LOCAL_PCG)");
  cudaFree(b->d_chol); cudaFree(b->d_chol_off); cudaFree(b->d_chol_work);
  b->d_chol = nullptr; b->d_chol_off = nullptr; b->d_chol_work = nullptr;
  HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_chol, (size_t)off * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_chol_off, b->nblocks * sizeof(CholBlock)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_chol_work, (size_t)woff * sizeof(double)));
  HSBP_CUDA(ctx, cudaMemcpyAsync(b->d_chol_off, cbs.data(), b->nblocks * sizeof(CholBlock), cudaMemcpyHostToDevice, ctx->stream));
  const CholBlock *dcb = (const CholBlock *)b->d_chol_off;
  // dense M-tilde: column c of every block = M-tilde e_c
  double *u = nullptr, *y = nullptr;
  HSBP_CUDA(ctx, cudaMalloc((void **)&u, (size_t)b->VNp * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&y, (size_t)b->VNp * sizeof(double)));
  HSBP_CUDA(ctx, cudaMemsetAsync(u, 0, (size_t)b->VNp * sizeof(double), ctx->stream));
  int rc = HSBP_OK;
  for (int c = 0; c < maxld && rc == HSBP_OK; ++c) {
    k_chol_unit<<<(unsigned)((b->nblocks + 127) / 128), 128, 0, ctx->stream>>>(dcb, b->nblocks, c, u);
    rc = apply_async(b, u, y);
    k_chol_store_col<<<(unsigned)b->nblocks, 128, 0, ctx->stream>>>(dcb, c, y, b->d_chol);
  }
  int *d_flag = nullptr;
  std::vector<int> flag(b->nblocks, 0);
  cudaError_t e1 = cudaMalloc((void **)&d_flag, b->nblocks * sizeof(int));
  if (rc == HSBP_OK && e1 == cudaSuccess) {
    cudaMemsetAsync(d_flag, 0, b->nblocks * sizeof(int), ctx->stream);
    for (int k0 = 0; k0 < maxld; k0 += CH_NB) {
      k_chol_panel<<<(unsigned)b->nblocks, CH_THREADS, 0, ctx->stream>>>(dcb, b->d_chol, k0, d_flag);
      const int nt = (maxld - k0 - CH_NB) / CH_NB;
      if (nt > 0)
        k_chol_update<<<dim3(nt, nt, (unsigned)b->nblocks), CH_THREADS, 0, ctx->stream>>>(dcb, b->d_chol, k0);
    }
    e1 = cudaGetLastError();
    if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(flag.data(), d_flag, b->nblocks * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(d_flag); cudaFree(u); cudaFree(y);
  if (rc) return rc;
  if (e1 != cudaSuccess) { ctx->err = std::string("chol_setup: ") + cudaGetErrorString(e1); return HSBP_ERR_CUDA; }
  for (int64_t e = 0; e < b->nblocks; ++e)
    if (flag[e]) HSBP_FAIL(ctx, HSBP_ERR_ARG, "dense Cholesky: M-tilde of a block is not positive definite");
  return HSBP_OK;
}

int chol_solve(hsbp_blocks *b, const double *g, double *x, hsbp_local_stats *stats) {
  hsbp_ctx *ctx = b->ctx;
  if (!b->d_chol) HSBP_FAIL(ctx, HSBP_ERR_STATE, "dense Cholesky local solver: not set up");
  k_chol_solve<<<(unsigned)b->nblocks, CH_THREADS, 0, ctx->stream>>>((const CholBlock *)b->d_chol_off, b->d_chol, g, x,
                                                                    b->d_chol_work);
  if (stats) { stats->iterations_max = 0; stats->iterations_sum = 0; stats->failed_blocks = 0; stats->max_rel_residual = 0.0; }
  return check_launch(ctx, "k_chol_solve");
}

}  // namespace
