// Solver-side kernels: vector primitives, the batched per-block PCG update (K2b), the trace-space
// gather / scatter between block faces and lambda (K3), and the Schur-complement pieces (K4).
//
// Reference being replaced:
//   local solves   factorization(M-tilde) and F \ g      global_curved.jl:698, 734; square_circle.jl:383
//   Fbar^T, D      glolambdaoperator                     global_curved.jl:510-565
//   B lambda       assemblelambdamatrix + cholesky(B)    global_curved.jl:743-797; square_circle.jl:313-377
#pragma once
#include "hsbp_internal.h"
#include "k_generic.cuh"

namespace hsbp {

// ---- small vector kernels (lambda space and volume space) ----------------------------------
constexpr int VEC_THREADS = 256;
constexpr int DOT_BLOCKS = 296;     // 2 per SM; partial sums are combined in a fixed order

__global__ void k_fill(double *x, int64_t n, double v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}
// z = a*x + b*y   (z may alias x or y)
__global__ void k_axpby(int64_t n, double a, const double *x, double b, const double *y, double *z) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    z[i] = a * x[i] + b * y[i];
}
// z = x * y (elementwise), or z = x / y
__global__ void k_ewise(int64_t n, const double *x, const double *y, double *z, int divide) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    z[i] = divide ? x[i] / y[i] : x[i] * y[i];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the CTA; result valid in every thread; scratch >= 32 doubles
__device__ __forceinline__ double cta_sum(double v, double *scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();                 // protect scratch from the previous use
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double t = (lane < nw) ? scratch[lane] : 0.0;
  t = warp_sum(t);
  return t;
}

// partial[b] = sum over this CTA's slice of x*y ; up to 3 dot products at once
__global__ void __launch_bounds__(VEC_THREADS)
k_dot3_partial(int64_t n, const double *x0, const double *y0, const double *x1, const double *y1,
               const double *x2, const double *y2, double *partial) {
  __shared__ double scratch[32];
  double s0 = 0, s1 = 0, s2 = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    s0 += x0[i] * y0[i];
    if (x1) s1 += x1[i] * y1[i];
    if (x2) s2 += x2[i] * y2[i];
  }
  s0 = cta_sum(s0, scratch); s1 = cta_sum(s1, scratch); s2 = cta_sum(s2, scratch);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = s0; partial[gridDim.x + blockIdx.x] = s1; partial[2 * gridDim.x + blockIdx.x] = s2;
  }
}
__global__ void k_dot3_final(int nb, const double *partial, double *out) {
  __shared__ double scratch[32];
  for (int q = 0; q < 3; ++q) {
    double s = 0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partial[q * nb + i];
    s = cta_sum(s, scratch);
    if (threadIdx.x == 0) out[q] = s;
  }
}

// ---- probing the diagonal of M-tilde (setup of the Jacobi preconditioner) -------------------
template <int P>
__global__ void k_color_vector(const BlockDesc *__restrict__ desc, int c, int ci, int cj, double *__restrict__ u) {
  const BlockDesc d = desc[blockIdx.x];
  const int Nrp = d.Nr + 1;
  const int64_t np = (int64_t)Nrp * (d.Ns + 1);
  for (int64_t idx = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; idx < np; idx += (int64_t)gridDim.y * blockDim.x) {
    const int j = (int)(idx / Nrp), i = (int)(idx - (int64_t)j * Nrp);
    u[d.voff + idx] = (i % c == ci && j % c == cj) ? 1.0 : 0.0;
  }
}
template <int P>
__global__ void k_color_pick(const BlockDesc *__restrict__ desc, int c, int ci, int cj,
                             const double *__restrict__ y, double *__restrict__ dinv) {
  const BlockDesc d = desc[blockIdx.x];
  const int Nrp = d.Nr + 1;
  const int64_t np = (int64_t)Nrp * (d.Ns + 1);
  for (int64_t idx = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; idx < np; idx += (int64_t)gridDim.y * blockDim.x) {
    const int j = (int)(idx / Nrp), i = (int)(idx - (int64_t)j * Nrp);
    if (i % c == ci && j % c == cj) dinv[d.voff + idx] = 1.0 / y[d.voff + idx];
  }
}

// ---- batched PCG, one CTA per block ---------------------------------------------------------
struct PcgState {      // per block
  double rz, g2, rr;
  double alpha;        // last step length (flexible beta of the FDM-preconditioned variant)
  int32_t active, iters;
};

// x = 0, r = g, z = Dinv r, p = z, rz = r.z, g2 = g.g
__global__ void __launch_bounds__(1024)
k_pcg_init(const BlockDesc *__restrict__ desc, const double *__restrict__ g, const double *__restrict__ dinv,
           double *__restrict__ x, double *__restrict__ r, double *__restrict__ p, PcgState *__restrict__ st,
           double tol2) {
  __shared__ double scratch[32];
  const BlockDesc d = desc[blockIdx.x];
  const int64_t np = (int64_t)(d.Nr + 1) * (d.Ns + 1);
  double s_rz = 0, s_gg = 0;
  for (int64_t i = threadIdx.x; i < np; i += blockDim.x) {
    const double gi = g[d.voff + i], zi = dinv[d.voff + i] * gi;
    x[d.voff + i] = 0.0; r[d.voff + i] = gi; p[d.voff + i] = zi;
    s_rz += gi * zi; s_gg += gi * gi;
  }
  s_rz = cta_sum(s_rz, scratch); s_gg = cta_sum(s_gg, scratch);
  if (threadIdx.x == 0) {
    PcgState s; s.rz = s_rz; s.g2 = s_gg; s.rr = s_gg; s.iters = 0;
    s.active = (s_gg > 0.0 && s_gg > tol2 * s_gg) ? 1 : 0;     // g == 0 -> x = 0 (global_curved.jl:733)
    st[blockIdx.x] = s;
  }
}

// one PCG iteration's vector work for every still-active block (Ap = M-tilde p was just computed)
__global__ void __launch_bounds__(1024)
k_pcg_update(const BlockDesc *__restrict__ desc, const double *__restrict__ dinv, const double *__restrict__ Ap,
             double *__restrict__ x, double *__restrict__ r, double *__restrict__ p, PcgState *__restrict__ st,
             double tol2, int *__restrict__ nactive) {
  __shared__ double scratch[32];
  const BlockDesc d = desc[blockIdx.x];
  PcgState s = st[blockIdx.x];
  if (!s.active) return;
  const int64_t np = (int64_t)(d.Nr + 1) * (d.Ns + 1);
  const int64_t o = d.voff;
  // blocks whose offset and size are even take 16-byte accesses, two independent pairs per trip (the loops are bandwidth bound)
  const bool vec2 = ((o | np) & 1) == 0;
  const int64_t n2 = np >> 1;
  const double2 *p2c = reinterpret_cast<const double2 *>(p + o), *A2 = reinterpret_cast<const double2 *>(Ap + o);
  const double2 *d2 = reinterpret_cast<const double2 *>(dinv + o);
  double pAp = 0;
  if (vec2) {
#pragma unroll 2
    for (int64_t i = threadIdx.x; i < n2; i += blockDim.x) {
      const double2 pv = p2c[i], av = A2[i];
      pAp = fma(pv.x, av.x, pAp); pAp = fma(pv.y, av.y, pAp);
    }
  } else {
    for (int64_t i = threadIdx.x; i < np; i += blockDim.x) pAp += p[o + i] * Ap[o + i];
  }
  pAp = cta_sum(pAp, scratch);
  const double alpha = s.rz / pAp;
  double rr = 0, rz = 0;
  if (vec2) {
    double2 *x2 = reinterpret_cast<double2 *>(x + o), *r2 = reinterpret_cast<double2 *>(r + o);
#pragma unroll 2
    for (int64_t i = threadIdx.x; i < n2; i += blockDim.x) {
      const double2 pv = p2c[i], av = A2[i], dv = d2[i];
      double2 xv = x2[i], rv = r2[i];
      xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
      rv.x = fma(-alpha, av.x, rv.x); rv.y = fma(-alpha, av.y, rv.y);
      x2[i] = xv; r2[i] = rv;
      rr = fma(rv.x, rv.x, rr); rr = fma(rv.y, rv.y, rr);
      rz = fma(rv.x * rv.x, dv.x, rz); rz = fma(rv.y * rv.y, dv.y, rz);
    }
  } else {
    for (int64_t i = threadIdx.x; i < np; i += blockDim.x) {
      x[o + i] += alpha * p[o + i];
      const double ri = r[o + i] - alpha * Ap[o + i];
      r[o + i] = ri;
      rr += ri * ri; rz += ri * ri * dinv[o + i];
    }
  }
  rr = cta_sum(rr, scratch); rz = cta_sum(rz, scratch);
  const double beta = rz / s.rz;
  const bool done = !(rr > tol2 * s.g2);
  if (vec2) {
    double2 *p2 = reinterpret_cast<double2 *>(p + o);
    const double2 *r2 = reinterpret_cast<const double2 *>(r + o);
#pragma unroll 2
    for (int64_t i = threadIdx.x; i < n2; i += blockDim.x) {
      const double2 dv = d2[i], rv = r2[i];
      double2 pv = p2[i];
      pv.x = done ? 0.0 : fma(beta, pv.x, dv.x * rv.x);
      pv.y = done ? 0.0 : fma(beta, pv.y, dv.y * rv.y);
      p2[i] = pv;
    }
  } else {
    for (int64_t i = threadIdx.x; i < np; i += blockDim.x)
      p[o + i] = done ? 0.0 : dinv[o + i] * r[o + i] + beta * p[o + i];
  }
  if (threadIdx.x == 0) {
    s.rz = rz; s.rr = rr; s.iters += 1; s.active = done ? 0 : 1;
    st[blockIdx.x] = s;
    if (!done) atomicAdd(nactive, 1);
  }
}

// ---- trace space <-> block faces -------------------------------------------------------------
struct LamFace {       // one face that carries lambda
  int32_t em, km, ep, kp;   // 0-based block and local face of the minus / plus side; -1: that side lives on another
                            // device (cut face of a partitioned mesh) -- its contribution is exchanged by the host layer
  int32_t flip, nl;
  int64_t loff;             // 0-based offset in lambda vectors
  int64_t fm, fp;           // offsets of the two block faces in block-face vectors
};

// lam = (F^T u)_minus + orient((F^T u)_plus)        rows of Fbar^T (global_curved.jl:533-553)
__global__ void k_lam_gather(const LamFace *__restrict__ lf, const double *__restrict__ ft, double *__restrict__ lam) {
  const LamFace f = lf[blockIdx.x];
  for (int n = threadIdx.x; n < f.nl; n += blockDim.x) {
    double v = 0.0;
    if (f.em >= 0) v = ft[f.fm + n];
    if (f.ep >= 0) v += ft[f.fp + (f.flip ? f.nl - 1 - n : n)];
    lam[f.loff + n] = v;
  }
}
// block-face vector v = lambda seen from each block face (0 on faces without lambda)
__global__ void k_lam_scatter(const LamFace *__restrict__ lf, const double *__restrict__ lam, double *__restrict__ v) {
  const LamFace f = lf[blockIdx.x];
  for (int n = threadIdx.x; n < f.nl; n += blockDim.x) {
    const double l = lam[f.loff + n];
    if (f.em >= 0) v[f.fm + n] = l;
    if (f.ep >= 0) v[f.fp + (f.flip ? f.nl - 1 - n : n)] = l;
  }
}
// D = Hf o (tau_minus + orient(tau_plus))            global_curved.jl:556-557
template <int P>
__global__ void k_lam_D(const LamFace *__restrict__ lf, const BlockDesc *__restrict__ desc,
                        const double *__restrict__ tau, double *__restrict__ D) {
  const LamFace f = lf[blockIdx.x];
  const BlockDesc d = desc[f.em >= 0 ? f.em : f.ep];
  const FaceGeom fg = face_geom(d, f.em >= 0 ? f.km : f.kp);     // same points, same norm on both sides
  for (int n = threadIdx.x; n < f.nl; n += blockDim.x) {
    double t = 0.0;
    if (f.em >= 0) t = tau[f.fm + n];
    if (f.ep >= 0) t += tau[f.fp + (f.flip ? f.nl - 1 - n : n)];
    D[f.loff + n] = fg.ht * hweight<P>(n, fg.Nt) * t;            // the norm weights are mirror symmetric
  }
}

// ---- static condensation: dense S_e = F_e^T M̃_e^-1 F_e per block ---------------------------------------------
// (what assembleλmatrix forms block by block, global_curved.jl:743-797: F' \ F slices and their products)
__device__ __forceinline__ int block_nf(const BlockDesc &d) { return 2 * (d.Ns + 1) + 2 * (d.Nr + 1); }
// block-face vector with a one at local face point c of every block (the previous one is cleared)
__global__ void k_cond_unit(const BlockDesc *__restrict__ desc, int64_t nblocks, int c, double *__restrict__ v) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nblocks) return;
  const BlockDesc d = desc[e];
  const int nf = block_nf(d);
  if (c > 0 && c - 1 < nf) v[d.foff + c - 1] = 0.0;
  if (c < nf) v[d.foff + c] = 1.0;
}
__global__ void k_cond_store(const BlockDesc *__restrict__ desc, const int64_t *__restrict__ soff, int c,
                             const double *__restrict__ ft, double *__restrict__ S) {
  const BlockDesc d = desc[blockIdx.x];
  const int nf = block_nf(d);
  if (c >= nf) return;
  double *col = S + soff[blockIdx.x] + (int64_t)c * nf;
  for (int r = threadIdx.x; r < nf; r += blockDim.x) col[r] = ft[d.foff + r];
}
// S <- (S + S^T) / 2: the columns come from iterative solves, the matrix-vector kernel reads rows as columns
__global__ void k_cond_sym(const BlockDesc *__restrict__ desc, const int64_t *__restrict__ soff, double *__restrict__ S) {
  const BlockDesc d = desc[blockIdx.x];
  const int nf = block_nf(d);
  double *Sb = S + soff[blockIdx.x];
  for (int64_t idx = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; idx < (int64_t)nf * nf; idx += (int64_t)gridDim.y * blockDim.x) {
    const int r = (int)(idx % nf), c = (int)(idx / nf);
    if (r <= c) continue;
    const double v = 0.5 * (Sb[r + (int64_t)nf * c] + Sb[c + (int64_t)nf * r]);
    Sb[r + (int64_t)nf * c] = v;
    Sb[c + (int64_t)nf * r] = v;
  }
}
// y_e = S_e v_e in the block-face layout: one warp per row (= column, S is symmetric), coalesced 8-byte reads;
// grid = (row groups, blocks).  Algorithmic traffic: 8 nf^2 bytes per block.
__global__ void __launch_bounds__(256)
k_cond_gemv(const BlockDesc *__restrict__ desc, const int64_t *__restrict__ soff, const double *__restrict__ S,
            const double *__restrict__ v, double *__restrict__ y) {
  extern __shared__ double sv[];
  const BlockDesc d = desc[blockIdx.y];
  const int nf = block_nf(d);
  const double *Sb = S + soff[blockIdx.y];
  for (int c = threadIdx.x; c < nf; c += blockDim.x) sv[c] = v[d.foff + c];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = blockIdx.x * nw + wid; r < nf; r += gridDim.x * nw) {
    const double *col = Sb + (int64_t)nf * r;
    double s0 = 0.0, s1 = 0.0;
    int c = lane;
    for (; c + 32 < nf; c += 64) { s0 += col[c] * sv[c]; s1 += col[c + 32] * sv[c + 32]; }
    if (c < nf) s0 += col[c] * sv[c];
    const double s = warp_sum(s0 + s1);
    if (lane == 0) y[d.foff + r] = s;
  }
}

// ---- face-block preconditioner of the trace system ----------------------------------------------------------
// A_f = B_ff = D_f - S_{e-}[f, f] - orient(S_{e+}[f, f]): the diagonal block of B = D - Fbar^T M̃^-1 Fbar that belongs to
// face f, from the condensed blocks; padded to ld with the identity.  A cut face of a partitioned mesh (one side on another
// device) keeps only D_f: both devices must apply the same preconditioner to their copies of lambda.
struct FaceBlock {      // layout-compatible with CholBlock (api_chol.cuh): off, np, ld, voff, woff
  int64_t off;
  int32_t np, ld;
  int64_t voff;
  int64_t woff;
};
// partner (optional): for cut faces, the other device's contribution S_{e'}[f, f] in lambda orientation, packed in
// the order of the cut faces (pidx[face] = offset into partner, -1: none -> the face keeps only D_f)
__global__ void __launch_bounds__(256)
k_faceblock_fill(const LamFace *__restrict__ lf, const FaceBlock *__restrict__ fb, const BlockDesc *__restrict__ desc,
                 const int64_t *__restrict__ soff, const double *__restrict__ S, const double *__restrict__ D,
                 double *__restrict__ A, const int64_t *__restrict__ pidx, const double *__restrict__ partner) {
  const LamFace f = lf[blockIdx.x];
  const FaceBlock q = fb[blockIdx.x];
  double *Ab = A + q.off;
  const bool cut = f.em < 0 || f.ep < 0;
  const double *Pn = (cut && pidx && pidx[blockIdx.x] >= 0) ? partner + pidx[blockIdx.x] : nullptr;
  const bool diag_only = cut && !Pn;
  const double *Sm = nullptr, *Sp = nullptr;
  int nfm = 0, nfp = 0, om = 0, op = 0;
  if (f.em >= 0) { const BlockDesc dm = desc[f.em]; nfm = block_nf(dm); om = (int)(f.fm - dm.foff); Sm = S + soff[f.em]; }
  if (f.ep >= 0) { const BlockDesc dp = desc[f.ep]; nfp = block_nf(dp); op = (int)(f.fp - dp.foff); Sp = S + soff[f.ep]; }
  for (int idx = threadIdx.x; idx < q.ld * q.ld; idx += blockDim.x) {
    const int i = idx % q.ld, j = idx / q.ld;
    double v = 0.0;
    if (i < q.np && j < q.np) {
      if (i == j) v = D[f.loff + i];
      if (!diag_only) {                                       // D - (one side + other side): the sum of the two sides is
                                                              // commutative, so both devices of a cut face build the same block
        const int ip = f.flip ? q.np - 1 - i : i, jp = f.flip ? q.np - 1 - j : j;
        const double cm = Sm ? Sm[(om + i) + (int64_t)nfm * (om + j)] : Pn[i + (int64_t)q.np * j];
        const double cp = Sp ? Sp[(op + ip) + (int64_t)nfp * (op + jp)] : Pn[i + (int64_t)q.np * j];
        v -= cm + cp;
      }
    } else if (i == j) {
      v = 1.0;
    }
    Ab[idx] = v;
  }
}
// this device's contribution to B_ff of a cut face, in lambda orientation, dense nl x nl
__global__ void __launch_bounds__(256)
k_faceblock_own(const LamFace *__restrict__ lf, const BlockDesc *__restrict__ desc, const int64_t *__restrict__ soff,
                const double *__restrict__ S, const int64_t *__restrict__ faces, const int64_t *__restrict__ ooff,
                double *__restrict__ out) {
  const LamFace f = lf[faces[blockIdx.x]];
  double *O = out + ooff[blockIdx.x];
  const bool minus = f.em >= 0;
  const BlockDesc d = desc[minus ? f.em : f.ep];
  const int nf = block_nf(d), o = (int)((minus ? f.fm : f.fp) - d.foff), nl = f.nl;
  const double *Sb = S + soff[minus ? f.em : f.ep];
  const bool flip = !minus && f.flip;
  for (int idx = threadIdx.x; idx < nl * nl; idx += blockDim.x) {
    const int i = idx % nl, j = idx / nl;
    const int ii = flip ? nl - 1 - i : i, jj = flip ? nl - 1 - j : j;
    O[idx] = Sb[(o + ii) + (int64_t)nf * (o + jj)];
  }
}

}  // namespace hsbp
