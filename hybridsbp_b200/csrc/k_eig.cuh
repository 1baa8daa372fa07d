// Batched symmetric eigen-decomposition for the setup of the fast-diagonalisation preconditioner (K2d): one CTA per matrix,
// one-sided (Hestenes) Jacobi in blocks of columns that live in shared memory.
//
// The matrices are the standard forms H^-1/2 A H^-1/2 of the collapsed 1-D operators (api_fdm.cuh): symmetric positive
// definite, n <= 256 (config 4), thousands of them.  Rotating columns of G = A V from the right until they are orthogonal gives
// A V = U S with V orthogonal; for a symmetric positive semi-definite A the columns of V are its eigenvectors and the column
// norms of G its eigenvalues.  V is a product of plane rotations: orthogonal to rounding whatever the spectrum looks like,
// which is what the preconditioner needs (its eigenvalues are Rayleigh quotients taken afterwards).
//
// A sweep visits every pair of column blocks (16 columns each): the 2 x 16 columns of G and of V are loaded into shared memory
// (128 KB at n = 256), all 496 pairs among them are rotated in 31 round-robin steps of 16 disjoint pairs (one pair per warp and
// step: three dot products by warp shuffles, then the plane rotation of both column pairs), and the columns go back.  Per sweep a
// matrix moves 31 MB through L2 instead of the 535 MB of a column-pair-at-a-time sweep.  Converged pairs (|g_p . g_q| <= tol
// |g_p| |g_q|) are skipped; the loop ends with the first sweep without a rotation.
//
// The reference has no counterpart (it factorises M-tilde, global_curved.jl:698); this replaces the cuSOLVER syevd calls of
// round 1 (2048 decompositions one after the other: 4.6 s at config 4).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hsbp {

constexpr int EIG_CB = 16;                   // columns per block (8 / 4 for matrices whose 2 x 16 columns of G and V exceed shared memory)
constexpr int EIG_THREADS = 512;             // 16 warps: one column pair per warp and round-robin step
inline size_t eig_smem_bytes(int n, int cb = EIG_CB) { return (size_t)4 * cb * n * sizeof(double); }

__device__ __forceinline__ double eig_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// A: [batch][n x n] column-major; in: symmetric matrix, out: eigenvectors as columns, eigenvalues ascending
// lam: [batch][n]; work: [batch][n x n] scratch; info: incremented once per matrix that did not converge in max_sweeps
template <int CB>
__global__ void __launch_bounds__(EIG_THREADS, 1)
k_jacobi_eig(int n, double *__restrict__ A, double *__restrict__ lam, double *__restrict__ work, int max_sweeps, double tol,
             int *__restrict__ info) {
  extern __shared__ __align__(16) double eig_sm[];
  constexpr int M2 = 2 * CB, RR = 2 * CB - 1;
  double *Gs = eig_sm;                         // [2 CB][n]
  double *Vs = eig_sm + (size_t)M2 * n;        // [2 CB][n]
  __shared__ int s_rot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = EIG_THREADS / 32;
  double *Ab = A + (int64_t)blockIdx.x * n * n, *Gb = work + (int64_t)blockIdx.x * n * n;
  for (int idx = tid; idx < n * n; idx += EIG_THREADS) Gb[idx] = Ab[idx];
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += EIG_THREADS) Ab[idx] = (idx % n == idx / n) ? 1.0 : 0.0;
  __syncthreads();
  const int nbk = (n + CB - 1) / CB;
  bool converged = false;
  for (int sweep = 0; sweep < max_sweeps && !converged; ++sweep) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int I = 0; I < nbk; ++I)
      for (int J = I + 1; J < nbk; ++J) {
        for (int idx = tid; idx < M2 * n; idx += EIG_THREADS) {
          const int c = idx / n, i = idx - c * n;
          const int gc = c < CB ? I * CB + c : J * CB + (c - CB);
          Gs[idx] = gc < n ? Gb[(int64_t)gc * n + i] : 0.0;
          Vs[idx] = gc < n ? Ab[(int64_t)gc * n + i] : 0.0;
        }
        __syncthreads();
        for (int step = 0; step < RR; ++step) {
          for (int k = warp; k < CB; k += nwarp) {                 // round-robin: column RR sits, the others walk the circle
            const int p = k == 0 ? RR : (step + k) % RR, q = (step + RR - k) % RR;
            double *gp = Gs + (size_t)p * n, *gq = Gs + (size_t)q * n;
            double a = 0.0, b = 0.0, c = 0.0;
            for (int i = lane; i < n; i += 32) {
              const double x = gp[i], y = gq[i];
              a = fma(x, x, a); b = fma(y, y, b); c = fma(x, y, c);
            }
            a = eig_warp_sum(a); b = eig_warp_sum(b); c = eig_warp_sum(c);
            if (fabs(c) > tol * sqrt(a * b)) {                     // the same value on every lane
              // tan of the rotation angle: t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (b - a) / (2 c), written with one
              // root and one division
              const double d = b - a;
              const double t = copysign(2.0 * c, d * c) / (fabs(d) + sqrt(fma(d, d, 4.0 * c * c)));
              const double cs = rsqrt(fma(t, t, 1.0)), sn = cs * t;
              double *vp = Vs + (size_t)p * n, *vq = Vs + (size_t)q * n;
              for (int i = lane; i < n; i += 32) {
                const double x = gp[i], y = gq[i];
                gp[i] = cs * x - sn * y; gq[i] = sn * x + cs * y;
                const double vx = vp[i], vy = vq[i];
                vp[i] = cs * vx - sn * vy; vq[i] = sn * vx + cs * vy;
              }
              if (lane == 0) s_rot = 1;
            }
          }
          __syncthreads();
        }
        for (int idx = tid; idx < M2 * n; idx += EIG_THREADS) {
          const int c = idx / n, i = idx - c * n;
          const int gc = c < CB ? I * CB + c : J * CB + (c - CB);
          if (gc < n) { Gb[(int64_t)gc * n + i] = Gs[idx]; Ab[(int64_t)gc * n + i] = Vs[idx]; }
        }
        __syncthreads();
      }
    converged = s_rot == 0;
    __syncthreads();
  }
  if (!converged && tid == 0) atomicAdd(info, 1);
  // eigenvalues = column norms of G; sort ascending (rank by counting), permute the columns of V through the scratch
  double *ls = eig_sm;                          // [n]
  int *rank = reinterpret_cast<int *>(eig_sm + n);
  for (int k = warp; k < n; k += nwarp) {
    double a = 0.0;
    for (int i = lane; i < n; i += 32) { const double x = Gb[(int64_t)k * n + i]; a = fma(x, x, a); }
    a = eig_warp_sum(a);
    if (lane == 0) ls[k] = sqrt(a);
  }
  __syncthreads();
  for (int k = tid; k < n; k += EIG_THREADS) {
    const double lk = ls[k];
    int r = 0;
    for (int j = 0; j < n; ++j) r += (ls[j] < lk || (ls[j] == lk && j < k)) ? 1 : 0;
    rank[k] = r;
    lam[(int64_t)blockIdx.x * n + r] = lk;
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
    const int k = idx / n, i = idx - k * n;
    Gb[(int64_t)rank[k] * n + i] = Ab[idx];
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += EIG_THREADS) Ab[idx] = Gb[idx];
}

}  // namespace hsbp
