// Generic (any block shape, p = 2/4/6) kernels of the block operator.
//
//   M-tilde u = A-tilde u + sum_k C-tilde_k u - sum_{k Neumann} F_k (tau_k H_k)^-1 F_k^T u
//   (reference locoperator, global_curved.jl:356, 444-447, 470, 478-486)
//
// is evaluated matrix-free as
//
//   y  = Arr u + Ass u + Qs^T[crs o (Qr u)] + Qr^T[crs o (Qs u)]      (volume part)
//   y += sum_k G_k^T alpha_k + L_k^T beta_k                            (face part)
//
// with a_k = L_k u (face restriction), g_k = G_k u (global_curved.jl:450-453) and
//   (alpha, beta) = (-a, tau H a - g)           Dirichlet / interface faces
//   (alpha, beta) = (-g / (tau H), 0)           Neumann faces
// which is algebraically what :444-447 and :478-486 assemble.
// The same face primitives give F_k^T u = g - tau H a and F_k v = G_k^T v - L_k^T tau H v.
#pragma once
#include "hsbp_internal.h"
#include "sbp1d.cuh"

namespace hsbp {

constexpr int GEN_THREADS = 256;

// ---- volume part, pass 1: t = crs o (Qr u), w = crs o (Qs u) -------------------------------
template <int P>
__global__ void __launch_bounds__(GEN_THREADS)
k_cross_pre(const BlockDesc *__restrict__ desc, const double *__restrict__ crs,
            const double *__restrict__ u, double *__restrict__ t, double *__restrict__ w) {
  const BlockDesc d = desc[blockIdx.x];
  const int Nrp = d.Nr + 1, Nsp = d.Ns + 1;
  const int64_t np = (int64_t)Nrp * Nsp;
  const double *ub = u + d.voff;
  for (int64_t idx = (int64_t)blockIdx.y * GEN_THREADS + threadIdx.x; idx < np;
       idx += (int64_t)gridDim.y * GEN_THREADS) {
    const int j = (int)(idx / Nrp), i = (int)(idx - (int64_t)j * Nrp);
    const double c = crs[d.voff + idx];
    const double qr = q_apply<P>(i, d.Nr, [&](int l) { return ub[l + (int64_t)Nrp * j]; });
    const double qs = q_apply<P>(j, d.Ns, [&](int l) { return ub[i + (int64_t)Nrp * l]; });
    t[d.voff + idx] = c * qr;
    w[d.voff + idx] = c * qs;
  }
}

// ---- volume part, pass 2 ------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(GEN_THREADS)
k_vol_apply(const BlockDesc *__restrict__ desc, const double *__restrict__ crr,
            const double *__restrict__ css, const double *__restrict__ u,
            const double *__restrict__ t, const double *__restrict__ w, double *__restrict__ y) {
  const BlockDesc d = desc[blockIdx.x];
  const int Nrp = d.Nr + 1, Nsp = d.Ns + 1;
  const int64_t np = (int64_t)Nrp * Nsp;
  const double hr = 2.0 / d.Nr, hs = 2.0 / d.Ns;
  const double *ub = u + d.voff, *tb = t + d.voff, *wb = w + d.voff;
  const double *rrb = crr + d.voff, *ssb = css + d.voff;
  for (int64_t idx = (int64_t)blockIdx.y * GEN_THREADS + threadIdx.x; idx < np;
       idx += (int64_t)gridDim.y * GEN_THREADS) {
    const int j = (int)(idx / Nrp), i = (int)(idx - (int64_t)j * Nrp);
    const int64_t row = (int64_t)Nrp * j;
    // Hs[j] * A(crr[:, j]) u[:, j]      (global_curved.jl:261-268)
    const double rr = m_apply<P>(i, d.Nr, [&](int l) { return rrb[l + row]; },
                                 [&](int l) { return ub[l + row]; });
    // Hr[i] * A(css[i, :]) u[i, :]      (global_curved.jl:313-322)
    const double ss = m_apply<P>(j, d.Ns, [&](int l) { return ssb[i + (int64_t)Nrp * l]; },
                                 [&](int l) { return ub[i + (int64_t)Nrp * l]; });
    const double sr = qt_apply<P>(j, d.Ns, [&](int l) { return tb[i + (int64_t)Nrp * l]; });   // :352
    const double rs = qt_apply<P>(i, d.Nr, [&](int l) { return wb[l + row]; });                 // :353
    y[d.voff + idx] = (hs * hweight<P>(j, d.Ns) / hr) * rr + (hr * hweight<P>(i, d.Nr) / hs) * ss + sr + rs;
  }
}

// ---- face geometry helper -----------------------------------------------------------------
struct FaceGeom {
  int Nt;          // tangential N
  int Nn;          // normal N
  int nf;          // points on the face
  int64_t fstart;  // start of the face inside the block's face storage
  double ht, hn;
  int sgn;         // -1 faces 1,3 ; +1 faces 2,4
};
__device__ __forceinline__ FaceGeom face_geom(const BlockDesc &d, int k) {
  FaceGeom g;
  const int Nrp = d.Nr + 1, Nsp = d.Ns + 1;
  if (k < 2) { g.Nt = d.Ns; g.Nn = d.Nr; g.nf = Nsp; g.fstart = (int64_t)k * Nsp; }
  else       { g.Nt = d.Nr; g.Nn = d.Ns; g.nf = Nrp; g.fstart = 2 * (int64_t)Nsp + (int64_t)(k - 2) * Nrp; }
  g.ht = 2.0 / g.Nt; g.hn = 2.0 / g.Nn;
  g.sgn = (k & 1) ? 1 : -1;
  return g;
}
// volume index (inside the block) of the point at normal offset m from face point n
__device__ __forceinline__ int64_t face_vol(const BlockDesc &d, int k, int n, int m) {
  const int Nrp = d.Nr + 1;
  switch (k) {
    case 0: return m + (int64_t)Nrp * n;
    case 1: return (d.Nr - m) + (int64_t)Nrp * n;
    case 2: return n + (int64_t)Nrp * m;
    default: return n + (int64_t)Nrp * (d.Ns - m);
  }
}

enum FaceMode { FACE_APPLY = 0, FACE_FT = 1, FACE_TRACTION = 2 };

// ---- face gather: a = L u, g = G u, then the mode's combination ----------------------------
// grid.x = 4 * nblocks
template <int P>
__global__ void __launch_bounds__(128)
k_face_gather(const BlockDesc *__restrict__ desc, const double *__restrict__ crr,
              const double *__restrict__ css, const double *__restrict__ crs,
              const double *__restrict__ tau, const double *__restrict__ u,
              double *__restrict__ fa, double *__restrict__ fb, int mode) {
  using S = Sbp<P>;
  const int e = blockIdx.x >> 2, k = blockIdx.x & 3;
  const BlockDesc d = desc[e];
  const FaceGeom fg = face_geom(d, k);
  const double *ub = u + d.voff;
  const double *cnn = (k < 2 ? crr : css) + d.voff;
  const double *cx = crs + d.voff;
  for (int n = threadIdx.x; n < fg.nf; n += blockDim.x) {
    const int64_t f0 = face_vol(d, k, n, 0);
    const double a = ub[f0];
    double bsu = 0.0;
#pragma unroll
    for (int m = 0; m < S::NB; ++m) bsu += S::bs()[m] * ub[face_vol(d, k, n, m)];
    const double Hf = fg.ht * hweight<P>(n, fg.Nt);
    const double qt = q_apply<P>(n, fg.Nt, [&](int l) { return ub[face_vol(d, k, l, 0)]; });
    const double g = (Hf / fg.hn) * cnn[f0] * bsu + fg.sgn * cx[f0] * qt;
    const int64_t fi = d.foff + fg.fstart + n;
    const double tH = tau[fi] * Hf;
    if (mode == FACE_APPLY) {
      if (d.bc[k] == HSBP_BC_NEUMANN) { fa[fi] = -g / tH; fb[fi] = 0.0; }
      else                            { fa[fi] = -a;      fb[fi] = tH * a - g; }
    } else if (mode == FACE_FT) {
      fa[fi] = g - tH * a;
    } else {
      fa[fi] = g / Hf - tau[fi] * a;
    }
  }
}

// alpha = s v, beta = -s tau H v   (so that scatter adds s * F_k v)
template <int P>
__global__ void __launch_bounds__(128)
k_face_prep_F(const BlockDesc *__restrict__ desc, const double *__restrict__ tau,
              const double *__restrict__ v, double s, double *__restrict__ fa, double *__restrict__ fb) {
  const int e = blockIdx.x >> 2, k = blockIdx.x & 3;
  const BlockDesc d = desc[e];
  const FaceGeom fg = face_geom(d, k);
  for (int n = threadIdx.x; n < fg.nf; n += blockDim.x) {
    const int64_t fi = d.foff + fg.fstart + n;
    const double Hf = fg.ht * hweight<P>(n, fg.Nt);
    const double vv = v[fi];
    fa[fi] = s * vv;
    fb[fi] = -s * tau[fi] * Hf * vv;
  }
}

// ---- face scatter: y += G_k^T alpha_k + L_k^T beta_k, faces in order (deterministic) -------
// grid.x = nblocks; faces are processed one after the other because they share corner points
template <int P>
__global__ void __launch_bounds__(256)
k_face_scatter(const BlockDesc *__restrict__ desc, const double *__restrict__ crr,
               const double *__restrict__ css, const double *__restrict__ crs,
               const double *__restrict__ fa, const double *__restrict__ fb, double *__restrict__ y) {
  using S = Sbp<P>;
  const BlockDesc d = desc[blockIdx.x];
  double *yb = y + d.voff;
  const double *cx = crs + d.voff;
  for (int k = 0; k < 4; ++k) {
    const FaceGeom fg = face_geom(d, k);
    const double *cnn = (k < 2 ? crr : css) + d.voff;
    const double *al = fa + d.foff + fg.fstart;
    const double *be = fb + d.foff + fg.fstart;
    for (int n = threadIdx.x; n < fg.nf; n += blockDim.x) {
      const int64_t f0 = face_vol(d, k, n, 0);
      const double Hf = fg.ht * hweight<P>(n, fg.Nt);
      const double cN = (Hf / fg.hn) * cnn[f0] * al[n];
      // tangential part: (Q_t^T (crs_face o alpha))_n lands on the face point itself
      const double qt = qt_apply<P>(n, fg.Nt, [&](int l) { return cx[face_vol(d, k, l, 0)] * al[l]; });
      yb[f0] += S::bs()[0] * cN + fg.sgn * qt + be[n];
#pragma unroll
      for (int m = 1; m < S::NB; ++m) yb[face_vol(d, k, n, m)] += S::bs()[m] * cN;
    }
    __syncthreads();
  }
}

// ---- penalty parameters tau_k (global_curved.jl:402-437) -----------------------------------
__device__ __forceinline__ void penalty_consts(int p, double &beta, double &alpha) {
  if (p == 2) { beta = 0.363636363; alpha = 1.0 / 2.0; }
  else if (p == 4) { beta = 0.2505765857; alpha = 17.0 / 48.0; }
  else { beta = 0.1878687080; alpha = 13649.0 / 43200.0; }
}

template <int P>
__global__ void __launch_bounds__(128)
k_compute_tau(const BlockDesc *__restrict__ desc, const double *__restrict__ crr,
              const double *__restrict__ css, const double *__restrict__ crs, double tauscale,
              double *__restrict__ tau, int *__restrict__ bad) {
  using S = Sbp<P>;
  const int e = blockIdx.x >> 2, k = blockIdx.x & 3;
  const BlockDesc d = desc[e];
  const FaceGeom fg = face_geom(d, k);
  double beta, alpha;
  penalty_consts(P, beta, alpha);
  const double *rr = crr + d.voff, *ss = css + d.voff, *rs = crs + d.voff;
  for (int n = threadIdx.x; n < fg.nf; n += blockDim.x) {
    // Evaluated with explicitly rounded operations (no FMA contraction) in the reference's
    // operation order: psi_min is a difference of nearly equal numbers for strongly anisotropic
    // tensors, so a contracted multiply-add would change tau in the 10th digit.
    double psi = 1e300;
    for (int m = 0; m < S::LPSI; ++m) {
      const int64_t v = face_vol(d, k, n, m);
      const double a = rr[v], b = ss[v], c = rs[v];
      const double dab = __dsub_rn(a, b);
      const double disc = __dadd_rn(__dmul_rn(dab, dab), __dmul_rn(4.0, __dmul_rn(c, c)));
      const double pm = __dmul_rn(__dsub_rn(__dadd_rn(a, b), __dsqrt_rn(disc)), 0.5);   // :418
      psi = fmin(psi, pm);
    }
    if (!(psi > 0.0)) atomicExch(bad, 1);                                        // :419
    const int64_t f0 = face_vol(d, k, n, 0);
    const double cn = (k < 2 ? rr : ss)[f0], cx = rs[f0];
    const double num = __dadd_rn(__ddiv_rn(__dmul_rn(cn, cn), beta), __ddiv_rn(__dmul_rn(cx, cx), alpha));
    tau[d.foff + fg.fstart + n] = __ddiv_rn(__dmul_rn(__ddiv_rn(__dmul_rn(2.0, tauscale), fg.hn), num), psi);
  }
}

// the reference asserts psi_min > 0 at every point of the block (global_curved.jl:418-419), not only in the layers the face
// penalties look at: same expression, same operation order, whole volume
__global__ void __launch_bounds__(256)
k_psi_min_check(int64_t n, const double *__restrict__ crr, const double *__restrict__ css, const double *__restrict__ crs,
                int *__restrict__ bad) {
  bool any = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double a = crr[i], b = css[i], c = crs[i];
    const double dab = __dsub_rn(a, b);
    const double disc = __dadd_rn(__dmul_rn(dab, dab), __dmul_rn(4.0, __dmul_rn(c, c)));
    const double pm = __dmul_rn(__dsub_rn(__dadd_rn(a, b), __dsqrt_rn(disc)), 0.5);
    any |= !(pm > 0.0);
  }
  if (__syncthreads_or(any) && threadIdx.x == 0) atomicExch(bad, 1);
}

}  // namespace hsbp
