// 1-D summation-by-parts building blocks on the device (fp64).
//
// Everything the block operator needs from reference diagonal_sbp.jl is evaluated
// on the fly from the published coefficient tables (sbp_tables_gen.h):
//   first derivative  Q = H*D, Q^T     diagonal_sbp.jl:67-161  (h cancels in H*D)
//   variable-coefficient stiffness M   diagonal_sbp.jl:474-746 (before the 1/h of :746)
//   boundary derivative row BS         diagonal_sbp.jl:507,511,591
// Nothing is ever assembled.  Lines are addressed through accessor functors so the
// same code serves the r and the s direction and the mirrored far-end closure.
#pragma once
#include "sbp_tables_gen.h"

namespace hsbp {

__constant__ double c_d1_d2[3] = HSBP_D1_D_2;
__constant__ double c_d1_d4[5] = HSBP_D1_D_4;
__constant__ double c_d1_d6[7] = HSBP_D1_D_6;
__constant__ double c_d1_bd2[HSBP_D1_BM_2 * HSBP_D1_BN_2] = HSBP_D1_BD_2;
__constant__ double c_d1_bd4[HSBP_D1_BM_4 * HSBP_D1_BN_4] = HSBP_D1_BD_4;
__constant__ double c_d1_bd6[HSBP_D1_BM_6 * HSBP_D1_BN_6] = HSBP_D1_BD_6;
__constant__ double c_d1_hw2[HSBP_D1_BM_2] = HSBP_D1_HW_2;
__constant__ double c_d1_hw4[HSBP_D1_BM_4] = HSBP_D1_HW_4;
__constant__ double c_d1_hw6[HSBP_D1_BM_6] = HSBP_D1_HW_6;
__constant__ double c_d2_bs2[3] = {1.5, -2.0, 0.5};
__constant__ double c_d2_bs4[4] = HSBP_D2_BS_4;
__constant__ double c_d2_bs6[5] = HSBP_D2_BS_6;
__constant__ double c_d2_T2[2] = {0.5, 0.5};
__constant__ double c_d2_T4[HSBP_D2_M_4 * HSBP_D2_M_4 * HSBP_D2_NK_4] = HSBP_D2_T_4;
__constant__ double c_d2_T6[HSBP_D2_M_6 * HSBP_D2_M_6 * HSBP_D2_NK_6] = HSBP_D2_T_6;

template <int P> struct Sbp;

template <> struct Sbp<2> {
  static constexpr int HALF = 1, BM = 1, BN = 2, M = 1, NK = 2, NB = 3, W = 1, LPSI = 2;
  __device__ static const double *d() { return c_d1_d2; }
  __device__ static const double *bd() { return c_d1_bd2; }
  __device__ static const double *hwt() { return c_d1_hw2; }
  __device__ static const double *bs() { return c_d2_bs2; }
  __device__ static const double *T() { return c_d2_T2; }
  // M[i][i+o] for an interior-type entry (diagonal_sbp.jl:495-503)
  template <class B> __device__ static double mint(int i, int o, B b) {
    if (o == -1) return -0.5 * (b(i - 1) + b(i));
    if (o == 1) return -0.5 * (b(i) + b(i + 1));
    return 0.5 * (b(i - 1) + 2.0 * b(i) + b(i + 1));
  }
};

template <> struct Sbp<4> {
  static constexpr int HALF = 2, BM = 4, BN = 6, M = 6, NK = 8, NB = 4, W = 3, LPSI = 4;
  __device__ static const double *d() { return c_d1_d4; }
  __device__ static const double *bd() { return c_d1_bd4; }
  __device__ static const double *hwt() { return c_d1_hw4; }
  __device__ static const double *bs() { return c_d2_bs4; }
  __device__ static const double *T() { return c_d2_T4; }
  // diagonal_sbp.jl:567-582
  template <class B> __device__ static double mint(int i, int o, B b) {
    switch (o) {
      case -2: return 0.125 * (b(i) + b(i - 2)) - (1.0 / 6.0) * b(i - 1);
      case -1: return -(1.0 / 6.0) * (b(i + 1) + b(i - 2)) - 0.5 * (b(i) + b(i - 1));
      case 0:  return (1.0 / 24.0) * (b(i + 2) + b(i - 2)) + (5.0 / 6.0) * (b(i + 1) + b(i - 1)) + 0.75 * b(i);
      case 1:  return -(1.0 / 6.0) * (b(i + 2) + b(i - 1)) - 0.5 * (b(i + 1) + b(i));
      default: return 0.125 * (b(i + 2) + b(i)) - (1.0 / 6.0) * b(i + 1);
    }
  }
};

template <> struct Sbp<6> {
  static constexpr int HALF = 3, BM = 6, BN = 9, M = 9, NK = 12, NB = 5, W = 5, LPSI = 7;
  __device__ static const double *d() { return c_d1_d6; }
  __device__ static const double *bd() { return c_d1_bd6; }
  __device__ static const double *hwt() { return c_d1_hw6; }
  __device__ static const double *bs() { return c_d2_bs6; }
  __device__ static const double *T() { return c_d2_T6; }
  // diagonal_sbp.jl:719-727 (weights kept exactly as the reference has them)
  template <class B> __device__ static double mint(int i, int o, B b) {
    switch (o) {
      case -3: return -(11.0 / 360.0) * (b(i - 3) + b(i)) + (1.0 / 40.0) * (b(i - 2) + b(i - 1));
      case -2: return (1.0 / 20.0) * (b(i - 3) + b(i + 1)) + (7.0 / 40.0) * (b(i - 2) + b(i)) - (3.0 / 10.0) * b(i - 1);
      case -1: return -(1.0 / 40.0) * (b(i - 3) + b(i + 2)) - (3.0 / 10.0) * (b(i - 2) + b(i + 1)) - (17.0 / 40.0) * (b(i - 1) + b(i));
      case 0:  return (1.0 / 180.0) * (b(i - 3) + b(i + 3)) + 0.125 * (b(i - 2) + b(i + 2)) + (19.0 / 20.0) * (b(i - 1) + b(i + 1)) + (101.0 / 180.0) * b(i);
      case 1:  return -(1.0 / 40.0) * (b(i - 2) + b(i + 3)) - (3.0 / 10.0) * (b(i - 1) + b(i + 2)) - (17.0 / 40.0) * (b(i) + b(i + 1));
      case 2:  return (1.0 / 20.0) * (b(i - 1) + b(i + 3)) + (7.0 / 40.0) * (b(i) + b(i + 2)) - (3.0 / 10.0) * b(i + 1);
      default: return -(11.0 / 360.0) * (b(i) + b(i + 3)) + (1.0 / 40.0) * (b(i + 1) + b(i + 2));
    }
  }
};

// norm weight of point i on an (N+1)-point line, in units of h (diagonal_sbp.jl:135-139)
template <int P> __device__ __forceinline__ double hweight(int i, int N) {
  using S = Sbp<P>;
  if (i < S::BM) return S::hwt()[i];
  if (i > N - S::BM) return S::hwt()[N - i];
  return 1.0;
}

// entry (k, j) of Q = H*D (pure number: the h of H cancels the 1/h of D)
template <int P> __device__ __forceinline__ double q_entry(int k, int j, int N) {
  using S = Sbp<P>;
  if (k < S::BM) return j < S::BN ? S::hwt()[k] * S::bd()[k * S::BN + j] : 0.0;
  if (k > N - S::BM) {
    const int kk = N - k, jj = N - j;
    return jj < S::BN ? -S::hwt()[kk] * S::bd()[kk * S::BN + jj] : 0.0;
  }
  const int o = j - k;
  return (o >= -S::HALF && o <= S::HALF) ? S::d()[o + S::HALF] : 0.0;
}

// (Q u)_k
template <int P, class U> __device__ __forceinline__ double q_apply(int k, int N, U u) {
  using S = Sbp<P>;
  double acc = 0.0;
  if (k < S::BM) {
#pragma unroll
    for (int j = 0; j < S::BN; ++j) acc += S::bd()[k * S::BN + j] * u(j);
    return S::hwt()[k] * acc;
  }
  if (k > N - S::BM) {
    const int kk = N - k;
#pragma unroll
    for (int j = 0; j < S::BN; ++j) acc += S::bd()[kk * S::BN + j] * u(N - j);
    return -S::hwt()[kk] * acc;
  }
#pragma unroll
  for (int o = -S::HALF; o <= S::HALF; ++o)
    if (o != 0) acc += S::d()[o + S::HALF] * u(k + o);
  return acc;
}

// (Q^T t)_j = sum_k Q[k][j] t(k)
template <int P, class Tt> __device__ __forceinline__ double qt_apply(int j, int N, Tt t) {
  using S = Sbp<P>;
  double acc = 0.0;
  if (j >= S::BN + S::HALF - 1 && j <= N - (S::BN + S::HALF - 1)) {
    // far from both closures: Q^T = -Q there
#pragma unroll
    for (int o = -S::HALF; o <= S::HALF; ++o)
      if (o != 0) acc += S::d()[S::HALF - o] * t(j + o);   // Q[j+o][j] = d[(j-(j+o)) + HALF]
    return acc;
  }
  const int k0 = max(0, j - S::W), k1 = min(N, j + S::W);
  for (int k = k0; k <= k1; ++k) {
    const double c = q_entry<P>(k, j, N);
    if (c != 0.0) acc += c * t(k);       // never touch t outside the sparsity pattern
  }
  return acc;
}

// (M(b) u)_i of the variable-coefficient stiffness matrix, before the division by h
template <int P, class B, class U> __device__ __forceinline__ double m_apply(int i, int N, B b, U u) {
  using S = Sbp<P>;
  if (i >= S::M && i <= N - S::M) {
    double acc = 0.0;
#pragma unroll
    for (int o = -S::HALF; o <= S::HALF; ++o) acc += S::mint(i, o, b) * u(i + o);
    return acc;
  }
  // closure row: mirror the far end onto the near end
  const bool far = i > N - S::M;
  const int ii = far ? N - i : i;
  auto bb = [&](int k) { return far ? b(N - k) : b(k); };
  auto uu = [&](int k) { return far ? u(N - k) : u(k); };
  double acc = 0.0;
  const double *T = S::T() + ii * S::M * S::NK;
  for (int j = 0; j < S::M; ++j) {
    double c = 0.0;
#pragma unroll
    for (int k = 0; k < S::NK; ++k) c += T[j * S::NK + k] * bb(k);
    acc += c * uu(j);
  }
  // interior-type entries that stick out of the closure block (rows M-HALF .. M-1)
  for (int j = S::M; j <= ii + S::HALF; ++j) acc += S::mint(ii, j - ii, bb) * uu(j);
  return acc;
}

}  // namespace hsbp
