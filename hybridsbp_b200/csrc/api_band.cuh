// K2c: batched banded fp64 Cholesky local solver.
//
// The reference keeps one sparse Cholesky factor of M-tilde_e per block for the whole run and only back-solves
// (`factors[e] = cholesky(Symmetric(M̃_e))`, global_curved.jl:698; `F \ g`, :734, square_circle.jl:383,
// seas/BP1/odefun.jl:43).  With points numbered r-fastest, M-tilde_e is banded: a point couples to points at most
// WB lines away in s and WB points away in r (closure blocks, boundary-derivative stencils, Neumann corrections and
// the mixed-derivative corners), so the half-bandwidth is kd = WB * (Nr+1) + WB.  For blocks that are too large for
// the dense factor (api_chol.cuh) but whose band fits in memory -- every block of configs 1-3 -- this is the
// direct-solver replacement of the plugin; the 256 x 256-point blocks of config 4 (674 MB of band each) stay with PCG.
//
//   setup   the band is filled by applying the matrix-free operator to (2 WB + 1)^2 coloured probe vectors (all blocks
//           at once), then factorised in place, right-looking in panels of 32 columns: per panel one kernel for the
//           diagonal block + panel TRSM and one kernel with a CTA per 32 x 32 tile of the trailing band window, on the
//           fp64 tensor pipe (mma.sync m8n8k4 f64) -- no fill occurs outside the band.
//   solve   L y = g, L^T x = y, one CTA per block, panels of 32.
// Storage (LAPACK lower band): AB[c * ld + d] = A[c + d][c], d = 0 .. kd, ld = kd + 1 rounded up to 32;
// the matrix is padded to a multiple of 32 columns with an identity block.
#pragma once
#include "api_chol.cuh"

namespace hsbp {

struct BandBlock {
  int64_t off;       // offset of the band storage
  int32_t np, npad;  // true size, padded size
  int32_t kd, ld;    // half-bandwidth, leading dimension
  int32_t Nrp, Nsp;
  int64_t voff;      // offset of the block in volume vectors
  int64_t woff;      // offset of the block's padded work vector
  int64_t ioff;      // offset of the block's inverted diagonal blocks (npad / BS_PB panels of BS_PB x BS_PB)
};

template <int P> struct BandWidth {
  using S = Sbp<P>;
  // farthest coupling of M-tilde in one direction (points / lines)
  static constexpr int WB = (S::M - 1 > S::BN - 1 ? S::M - 1 : S::BN - 1) > (2 * S::HALF > S::NB - 1 ? 2 * S::HALF : S::NB - 1)
                                ? (S::M - 1 > S::BN - 1 ? S::M - 1 : S::BN - 1)
                                : (2 * S::HALF > S::NB - 1 ? 2 * S::HALF : S::NB - 1);
};

// identity on the pad, zero elsewhere
__global__ void k_band_init(const BandBlock *__restrict__ bb, double *__restrict__ AB) {
  const BandBlock b = bb[blockIdx.x];
  double *A = AB + b.off;
  const int64_t n = (int64_t)b.npad * b.ld;
  for (int64_t idx = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.y * blockDim.x) {
    const int c = (int)(idx / b.ld), d = (int)(idx - (int64_t)c * b.ld);
    A[idx] = (c >= b.np && d == 0) ? 1.0 : 0.0;
  }
}

// y = M-tilde u for the probe vector of colour (ci, cj) modulo C: every row (i', j') of y belongs to the one probe
// column (i0, j0) within WB of it; the lower-triangle entries go into the band
__global__ void k_band_pick(const BandBlock *__restrict__ bb, int C, int WB, int ci, int cj, const double *__restrict__ y,
                            double *__restrict__ AB) {
  const BandBlock b = bb[blockIdx.x];
  double *A = AB + b.off;
  for (int64_t idx = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; idx < b.np; idx += (int64_t)gridDim.y * blockDim.x) {
    const int jp = (int)(idx / b.Nrp), ip = (int)(idx - (int64_t)jp * b.Nrp);
    // i0 in [ip - WB, ip + WB] with i0 % C == ci  (C = 2 WB + 1: exactly one candidate)
    int i0 = ip - WB + (((ci - (ip - WB)) % C) + C) % C;
    int j0 = jp - WB + (((cj - (jp - WB)) % C) + C) % C;
    if (i0 < 0 || i0 >= b.Nrp || j0 < 0 || j0 >= b.Nsp) continue;
    const int64_t c = i0 + (int64_t)b.Nrp * j0;
    if (idx >= c) A[c * b.ld + (idx - c)] = y[b.voff + idx];
  }
}

__device__ __forceinline__ double band_get(const double *A, int ld, int kd, int r, int c) {
  const int d = r - c;
  return (d >= 0 && d <= kd) ? A[(int64_t)c * ld + d] : 0.0;
}

// panel step k0: diagonal block (CUDA cores) and the rows below it inside the band (X L^T = A21, one row per thread)
__global__ void __launch_bounds__(CH_THREADS)
k_band_panel(const BandBlock *__restrict__ bb, double *__restrict__ AB, int k0, int *__restrict__ flag) {
  __shared__ double D[CH_NB][CH_NB + 1];
  const BandBlock b = bb[blockIdx.x];
  if (k0 >= b.npad) return;
  double *A = AB + b.off;
  const int ld = b.ld, kd = b.kd, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  bool bad = false;
  for (int idx = tid; idx < CH_NB * CH_NB; idx += CH_THREADS) {
    const int i = idx % CH_NB, j = idx / CH_NB;
    D[i][j] = band_get(A, ld, kd, k0 + i, k0 + j);
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
    if (i >= j && i - j <= kd) A[(int64_t)(k0 + j) * ld + (i - j)] = D[i][j];
  }
  const int rend = min(b.npad, k0 + CH_NB + kd);              // rows beyond have no entry in these columns
  for (int r = k0 + CH_NB + tid; r < rend; r += CH_THREADS) {
    double x[CH_NB];
#pragma unroll
    for (int j = 0; j < CH_NB; ++j) x[j] = band_get(A, ld, kd, r, k0 + j);
#pragma unroll
    for (int j = 0; j < CH_NB; ++j) {
      double s = x[j];
#pragma unroll
      for (int c = 0; c < j; ++c) s -= x[c] * D[j][c];
      x[j] = s / D[j][j];
    }
#pragma unroll
    for (int j = 0; j < CH_NB; ++j)
      if (r - (k0 + j) <= kd) A[(int64_t)(k0 + j) * ld + (r - k0 - j)] = x[j];
  }
}

// trailing update of the band window behind panel k0; grid = (tiles, tiles, blocks), tile (ti, tj), tj <= ti
__global__ void __launch_bounds__(CH_THREADS)
k_band_update(const BandBlock *__restrict__ bb, double *__restrict__ AB, int k0) {
  __shared__ double Ti[CH_NB][CH_NB + 1];
  __shared__ double Tj[CH_NB][CH_NB + 1];
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (tj > ti) return;
  const BandBlock b = bb[blockIdx.z];
  const int ld = b.ld, kd = b.kd, m0 = k0 + CH_NB;
  const int r0 = m0 + ti * CH_NB, c0 = m0 + tj * CH_NB;
  if (r0 >= b.npad || r0 >= m0 + kd) return;                    // rows without entries in the panel's columns
  if (r0 - (c0 + CH_NB - 1) > kd) return;                       // tile entirely outside the band
  double *A = AB + b.off;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int idx = tid; idx < CH_NB * CH_NB; idx += CH_THREADS) {
    const int i = idx % CH_NB, k = idx / CH_NB;
    Ti[i][k] = (r0 + i < b.npad) ? band_get(A, ld, kd, r0 + i, k0 + k) : 0.0;
    Tj[i][k] = (c0 + i < b.npad) ? band_get(A, ld, kd, c0 + i, k0 + k) : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int st = wid * 2 + s, si = (st >> 2) * 8, sj = (st & 3) * 8;
    double q0 = 0.0, q1 = 0.0;
#pragma unroll
    for (int k = 0; k < CH_NB; k += 4)
      dmma_m8n8k4(q0, q1, Ti[si + (lane >> 2)][k + (lane & 3)], Tj[sj + (lane >> 2)][k + (lane & 3)]);
    const int gi = r0 + si + (lane >> 2);
    const int gj = c0 + sj + (lane & 3) * 2;
    if (gi < b.npad) {
      if (gi - gj >= 0 && gi - gj <= kd) A[(int64_t)gj * ld + (gi - gj)] -= q0;
      if (gi - gj - 1 >= 0 && gi - gj - 1 <= kd && gj + 1 < b.npad) A[(int64_t)(gj + 1) * ld + (gi - gj - 1)] -= q1;
    }
  }
}

// x_e = (L L^T)^-1 g_e for every block
__global__ void __launch_bounds__(CH_THREADS)
k_band_solve(const BandBlock *__restrict__ bb, const double *__restrict__ AB, const double *__restrict__ g,
             double *__restrict__ x, double *__restrict__ work) {
  __shared__ double D[CH_NB][CH_NB + 1];
  __shared__ double xb[CH_NB];
  const BandBlock b = bb[blockIdx.x];
  const double *A = AB + b.off;
  double *r = work + b.woff;
  const int ld = b.ld, kd = b.kd, npad = b.npad, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < npad; i += CH_THREADS) r[i] = i < b.np ? g[b.voff + i] : 0.0;
  __syncthreads();
  // ---- L y = g ---------------------------------------------------------------------------------
  for (int k0 = 0; k0 < npad; k0 += CH_NB) {
    for (int idx = tid; idx < CH_NB * CH_NB; idx += CH_THREADS) {
      const int i = idx % CH_NB, j = idx / CH_NB;
      D[i][j] = band_get(A, ld, kd, k0 + i, k0 + j);
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
    const int rend = min(npad, k0 + CH_NB + kd);
    for (int i = k0 + CH_NB + tid; i < rend; i += CH_THREADS) {
      double s = r[i];
#pragma unroll 8
      for (int j = 0; j < CH_NB; ++j) {
        const int d = i - (k0 + j);
        if (d <= kd) s -= A[(int64_t)(k0 + j) * ld + d] * xb[j];
      }
      r[i] = s;
    }
    __syncthreads();
  }
  // ---- L^T x = y -------------------------------------------------------------------------------
  for (int k0 = npad - CH_NB; k0 >= 0; k0 -= CH_NB) {
    for (int c = wid; c < CH_NB; c += CH_THREADS / 32) {      // s_c = L[k0+32.., k0+c] . x[k0+32..]  (inside the band)
      const double *col = A + (int64_t)(k0 + c) * ld;
      const int iend = min(npad, k0 + c + kd + 1);
      double s = 0.0;
      for (int i = k0 + CH_NB + lane; i < iend; i += 32) s += col[i - (k0 + c)] * r[i];
      s = warp_sum(s);
      if (lane == 0) xb[c] = r[k0 + c] - s;
    }
    for (int idx = tid; idx < CH_NB * CH_NB; idx += CH_THREADS) {
      const int i = idx % CH_NB, j = idx / CH_NB;
      D[i][j] = band_get(A, ld, kd, k0 + i, k0 + j);
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

// ---- streamed solve ---------------------------------------------------------------------------------------------
// The solve above touches the band through ordinary loads with several block-wide barriers per panel: 30 ms for the single
// 201 x 201 block of BP1 (134 MB of band, read twice).  k_band_solve_stream keeps the same data layout -- the BS_PB columns
// of a panel are contiguous in memory -- and streams the panels through a shared-memory ring with bulk copies (TMA) that
// run ahead of the arithmetic; the active part of the right-hand side lives in a circular shared-memory window, and the
// triangular solves with the diagonal blocks are matrix-vector products with their precomputed inverses, so the only
// serial chain per panel is one 16-term dot product per lane.
constexpr int BS_PB = 16;             // panel width of the streamed solve
constexpr int BS_THREADS = 512;
constexpr int BS_WIN = 2048;          // circular window of the right-hand side / solution (doubles); needs kd + 2 BS_PB <= BS_WIN

// inverse of every PB x PB lower-triangular diagonal block of the factor, row-major [i][j]
template <int PB>
__global__ void __launch_bounds__(PB)
k_band_invdiag(const BandBlock *__restrict__ bb, const double *__restrict__ AB, double *__restrict__ inv) {
  __shared__ double D[PB][PB + 1];
  const BandBlock b = bb[blockIdx.y];
  const int k0 = blockIdx.x * PB;
  if (k0 >= b.npad) return;
  const double *A = AB + b.off;
  const int c = threadIdx.x;
  for (int i = 0; i < PB; ++i) D[i][c] = band_get(A, b.ld, b.kd, k0 + i, k0 + c);
  __syncthreads();
  double x[PB];                    // column c of the inverse: D x = e_c
#pragma unroll
  for (int i = 0; i < PB; ++i) {
    double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < PB; ++j)
      if (j < i) s -= D[i][j] * x[j];
    x[i] = (i >= c) ? s / D[i][i] : 0.0;
  }
  double *out = inv + b.ioff + (int64_t)blockIdx.x * PB * PB;
#pragma unroll
  for (int i = 0; i < PB; ++i) out[i * PB + c] = x[i];
}

// x_e = (L L^T)^-1 g_e, one CTA per block.  Dynamic shared memory: nst panel stages of PB * ld doubles, nst inverse
// blocks, the window, 2 * PB doubles, nst mbarriers.  PB: panel width (16; 8 or 4 when a stage of 16 columns of a wide band
// would not leave room for two stages, e.g. p = 6 blocks of 137 x 137 points).
template <int PB>
__global__ void __launch_bounds__(BS_THREADS, 1)
k_band_solve_stream(const BandBlock *__restrict__ bb, const double *__restrict__ AB, const double *__restrict__ inv,
                    const double *__restrict__ g, double *__restrict__ x, double *__restrict__ work, int nst, int maxld) {
  extern __shared__ __align__(128) unsigned char band_smem[];
  const BandBlock b = bb[blockIdx.x];
  const double *A = AB + b.off;
  const double *Iv = inv + b.ioff;
  double *y = work + b.woff;                                  // forward result (global, npad doubles)
  const int ld = b.ld, kd = b.kd, npad = b.npad, np = b.np, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int npan = npad / PB;
  double *pan = reinterpret_cast<double *>(band_smem);        // [nst][PB * maxld]
  double *ivs = pan + (size_t)nst * PB * maxld;            // [nst][PB * PB]
  double *win = ivs + (size_t)nst * PB * PB;            // [BS_WIN]
  double *pv = win + BS_WIN;                                  // [2 * PB]: panel solution / column dots
  uint64_t *full = reinterpret_cast<uint64_t *>(pv + 2 * PB);
  const uint32_t pan_bytes = (uint32_t)(PB * ld) * 8u, inv_bytes = (uint32_t)(PB * PB) * 8u;
  if (tid == 0) {
    for (int s = 0; s < nst; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int k, int st) {                           // panel k into stage st (thread 0)
    mbar_expect_tx(&full[st], pan_bytes + inv_bytes);
    bulk_g2s(pan + (size_t)st * PB * maxld, A + (int64_t)k * PB * ld, pan_bytes, &full[st]);
    bulk_g2s(ivs + (size_t)st * PB * PB, Iv + (int64_t)k * PB * PB, inv_bytes, &full[st]);
  };
  // ================= L y = g: panels in increasing order =================
  if (tid == 0) {
    fence_proxy_async();
    for (int k = 0; k < nst && k < npan; ++k) issue(k, k);
  }
  for (int i = tid; i < PB + kd && i < BS_WIN; i += BS_THREADS) win[i] = (i < np) ? g[b.voff + i] : 0.0;
  __syncthreads();
  int st = 0;
  uint32_t parity = 0;
  for (int k = 0; k < npan; ++k) {
    const int k0 = k * PB;
    // rows that enter the window for the next panel: k0 + PB + kd .. + PB - 1 (loaded early, stored late)
    double gnew = 0.0;
    const int inew = k0 + PB + kd + (tid - 32);
    if (wid == 1 && lane < PB && inew < np) gnew = g[b.voff + inew];
    mbar_wait(&full[st], parity);
    const double *P = pan + (size_t)st * PB * maxld;
    const double *Iq = ivs + (size_t)st * PB * PB;
    if (wid == 0 && lane < PB) {                           // y_panel = inv(D) r_panel
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < PB; ++j) s += Iq[lane * PB + j] * win[(k0 + j) & (BS_WIN - 1)];
      pv[lane] = s;
      y[k0 + lane] = s;
    }
    __syncthreads();
    for (int t = tid; t < kd; t += BS_THREADS) {              // r_i -= sum_j L[i][k0+j] y_j, i = k0 + PB + t
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int d = t + PB - j;
        if (d <= kd) s += P[j * ld + d] * pv[j];
      }
      win[(k0 + PB + t) & (BS_WIN - 1)] -= s;
    }
    if (wid == 1 && lane < PB) win[inew & (BS_WIN - 1)] = gnew;
    __syncthreads();
    if (tid == 0 && k + nst < npan) { fence_proxy_async(); issue(k + nst, st); }
    if (++st == nst) { st = 0; parity ^= 1u; }
  }
  // ================= L^T x = y: panels in decreasing order =================
  __syncthreads();
  for (int i = tid; i < BS_WIN; i += BS_THREADS) win[i] = 0.0;   // x beyond the last row is zero
  // the stage / parity sequence simply continues: panel (npan - 1 - m) is the (npan + m)-th use of the ring
  if (tid == 0) {
    fence_proxy_async();
    int s2 = st;
    for (int m = 0; m < nst && m < npan; ++m) { issue(npan - 1 - m, s2); if (++s2 == nst) s2 = 0; }
  }
  __syncthreads();
  for (int m = 0; m < npan; ++m) {
    const int k = npan - 1 - m, k0 = k * PB;
    double yv = 0.0;
    if (wid == 0 && lane < PB) yv = y[k0 + lane];           // early load
    mbar_wait(&full[st], parity);
    const double *P = pan + (size_t)st * PB * maxld;
    const double *Iq = ivs + (size_t)st * PB * PB;
    if (wid < PB) {                                             // s_j = L[k0+PB.., k0+j] . x[k0+PB..], one warp per column
      const int j = wid;                                        // BS_THREADS / 32 >= PB warps
      double s = 0.0;
      const int tend = kd - PB + j;                          // d = t + PB - j <= kd
      for (int t = lane; t <= tend; t += 32) s += P[j * ld + t + PB - j] * win[(k0 + PB + t) & (BS_WIN - 1)];
      s = warp_sum(s);
      if (lane == 0) pv[PB + j] = s;
    }
    __syncthreads();
    if (wid == 0 && lane < PB) {                             // x_panel = inv(D)^T (y_panel - s)
      pv[lane] = yv - pv[PB + lane];
      __syncwarp((1u << PB) - 1u);
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < PB; ++j)
        if (j >= lane) s += Iq[j * PB + lane] * pv[j];
      win[(k0 + lane) & (BS_WIN - 1)] = s;
      if (k0 + lane < np) x[b.voff + k0 + lane] = s;
    }
    __syncthreads();
    if (tid == 0 && m + nst < npan) { fence_proxy_async(); issue(npan - 1 - (m + nst), st); }
    if (++st == nst) { st = 0; parity ^= 1u; }
  }
}

}  // namespace hsbp

namespace {

using namespace hsbp;

template <int P> int band_setup_p(hsbp_blocks *b) {
  hsbp_ctx *ctx = b->ctx;
  constexpr int WB = BandWidth<P>::WB, C = 2 * WB + 1;
  std::vector<BandBlock> bbs(b->nblocks);
  int64_t off = 0, woff = 0, ioff = 0;
  int maxnpad = 0, maxkd = 0, minkd = 1 << 30, maxld = 0;
  for (int64_t e = 0; e < b->nblocks; ++e) {
    const BlockDesc &d = b->h_desc[e];
    BandBlock &q = bbs[e];
    q.Nrp = d.Nr + 1; q.Nsp = d.Ns + 1;
    q.np = q.Nrp * q.Nsp;
    q.npad = (q.np + CH_NB - 1) / CH_NB * CH_NB;
    q.kd = std::min(q.np - 1, WB * q.Nrp + WB);
    q.ld = (q.kd + 1 + 31) / 32 * 32;
    q.off = off; q.woff = woff; q.voff = d.voff;
    q.ioff = ioff;
    off += (int64_t)q.npad * q.ld; woff += q.npad; ioff += (int64_t)q.npad * BS_PB;
    maxnpad = std::max(maxnpad, q.npad); maxkd = std::max(maxkd, q.kd); minkd = std::min(minkd, q.kd);
    maxld = std::max(maxld, q.ld);
  }
  size_t free_b = 0, total_b = 0;
  HSBP_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  if ((size_t)off * sizeof(double) > free_b / 2)
    HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "banded Cholesky local solver: the bands do not fit in device memory (use HSBP_LOCAL_PCG)");
  cudaFree(b->d_band); cudaFree(b->d_band_desc); cudaFree(b->d_band_work); cudaFree(b->d_band_inv);
  b->d_band = nullptr; b->d_band_desc = nullptr; b->d_band_work = nullptr; b->d_band_inv = nullptr;
  HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_band_inv, (size_t)ioff * sizeof(double)));
  // streamed solve: at least two panel stages must fit in shared memory, the window must hold a panel's rows
  {
    // widest panel (16, 8, 4 columns) that leaves room for two stages
    b->band_stream_stages = 0; b->band_pb = BS_PB;
    for (int pb = BS_PB; pb >= 4 && b->band_stream_stages == 0; pb /= 2) {
      const size_t fixed = (size_t)BS_WIN * 8 + 2 * pb * 8 + 64 + 2048;
      const size_t per_stage = (size_t)pb * maxld * 8 + (size_t)pb * pb * 8;
      const int nst = (int)std::min<size_t>(4, (ctx->smem_optin > fixed ? (ctx->smem_optin - fixed) / per_stage : 0));
      if (nst >= 2 && minkd >= pb && maxkd + 2 * pb <= BS_WIN) { b->band_stream_stages = nst; b->band_pb = pb; }
    }
    b->band_maxld = maxld;
    b->band_maxnpad = maxnpad;
  }
  HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_band, (size_t)off * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_band_desc, b->nblocks * sizeof(BandBlock)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_band_work, (size_t)woff * sizeof(double)));
  HSBP_CUDA(ctx, cudaMemcpyAsync(b->d_band_desc, bbs.data(), b->nblocks * sizeof(BandBlock), cudaMemcpyHostToDevice, ctx->stream));
  const BandBlock *dbb = (const BandBlock *)b->d_band_desc;
  const dim3 grid = gen_grid(b);
  k_band_init<<<dim3((unsigned)b->nblocks, 64), 256, 0, ctx->stream>>>(dbb, b->d_band);
  double *u = nullptr, *y = nullptr;
  HSBP_CUDA(ctx, cudaMalloc((void **)&u, (size_t)b->VNp * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&y, (size_t)b->VNp * sizeof(double)));
  int rc = HSBP_OK;
  for (int cj = 0; cj < C && rc == HSBP_OK; ++cj)
    for (int ci = 0; ci < C && rc == HSBP_OK; ++ci) {
      k_color_vector<P><<<grid, GEN_THREADS, 0, ctx->stream>>>(b->d_desc, C, ci, cj, u);
      rc = apply_async(b, u, y);
      k_band_pick<<<dim3((unsigned)b->nblocks, 64), 256, 0, ctx->stream>>>(dbb, C, WB, ci, cj, y, b->d_band);
    }
  int *d_flag = nullptr;
  std::vector<int> flag(b->nblocks, 0);
  cudaError_t e1 = cudaMalloc((void **)&d_flag, b->nblocks * sizeof(int));
  if (rc == HSBP_OK && e1 == cudaSuccess) {
    cudaMemsetAsync(d_flag, 0, b->nblocks * sizeof(int), ctx->stream);
    const int ntmax = (maxkd + CH_NB - 1) / CH_NB + 1;
    for (int k0 = 0; k0 < maxnpad; k0 += CH_NB) {
      k_band_panel<<<(unsigned)b->nblocks, CH_THREADS, 0, ctx->stream>>>(dbb, b->d_band, k0, d_flag);
      const int nt = std::min(ntmax, (maxnpad - k0 - CH_NB) / CH_NB);
      if (nt > 0)
        k_band_update<<<dim3(nt, nt, (unsigned)b->nblocks), CH_THREADS, 0, ctx->stream>>>(dbb, b->d_band, k0);
    }
    if (b->band_pb == 16) k_band_invdiag<16><<<dim3((unsigned)(maxnpad / 16), (unsigned)b->nblocks), 16, 0, ctx->stream>>>(dbb, b->d_band, b->d_band_inv);
    else if (b->band_pb == 8) k_band_invdiag<8><<<dim3((unsigned)(maxnpad / 8), (unsigned)b->nblocks), 8, 0, ctx->stream>>>(dbb, b->d_band, b->d_band_inv);
    else k_band_invdiag<4><<<dim3((unsigned)(maxnpad / 4), (unsigned)b->nblocks), 4, 0, ctx->stream>>>(dbb, b->d_band, b->d_band_inv);
    e1 = cudaGetLastError();
    if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(flag.data(), d_flag, b->nblocks * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(d_flag); cudaFree(u); cudaFree(y);
  if (rc) return rc;
  if (e1 != cudaSuccess) { ctx->err = std::string("band_setup: ") + cudaGetErrorString(e1); return HSBP_ERR_CUDA; }
  for (int64_t e = 0; e < b->nblocks; ++e)
    if (flag[e]) HSBP_FAIL(ctx, HSBP_ERR_ARG, "banded Cholesky: M-tilde of a block is not positive definite");
  return HSBP_OK;
}

int band_setup(hsbp_blocks *b) {
  return dispatch_p(b->p, [&](auto Pc) { return band_setup_p<decltype(Pc)::value>(b); });
}

int band_solve(hsbp_blocks *b, const double *g, double *x, hsbp_local_stats *stats) {
  hsbp_ctx *ctx = b->ctx;
  if (!b->d_band) HSBP_FAIL(ctx, HSBP_ERR_STATE, "banded Cholesky local solver: not set up");
  if (b->band_stream_stages >= 2 && !b->band_no_stream) {
    const int nst = b->band_stream_stages, pb = b->band_pb;
    const size_t sm = (size_t)nst * pb * b->band_maxld * 8 + (size_t)nst * pb * pb * 8 + (size_t)BS_WIN * 8 +
                      2 * pb * 8 + nst * sizeof(uint64_t);
    auto go = [&](auto kern) -> int {
      HSBP_CUDA(ctx, hsbp_smem_optin(ctx, kern, sm));
      kern<<<(unsigned)b->nblocks, BS_THREADS, sm, ctx->stream>>>((const BandBlock *)b->d_band_desc, b->d_band, b->d_band_inv, g, x,
                                                                  b->d_band_work, nst, b->band_maxld);
      return HSBP_OK;
    };
    const int rcl = pb == 16 ? go(k_band_solve_stream<16>) : (pb == 8 ? go(k_band_solve_stream<8>) : go(k_band_solve_stream<4>));
    if (rcl) return rcl;
  } else {
    k_band_solve<<<(unsigned)b->nblocks, CH_THREADS, 0, ctx->stream>>>((const BandBlock *)b->d_band_desc, b->d_band, g, x,
                                                                      b->d_band_work);
  }
  if (stats) { stats->iterations_max = 0; stats->iterations_sum = 0; stats->failed_blocks = 0; stats->max_rel_residual = 0.0; }
  return check_launch(ctx, "k_band_solve");
}

}  // namespace
