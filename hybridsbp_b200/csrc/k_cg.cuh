// K4: kernels of the device-resident, two-level preconditioned CG on the trace system
//     B lambda = b,   B = D - Fbar^T M̃^-1 Fbar      (global_curved.jl:743-797; solved directly at square_circle.jl:314, 377)
// over a mesh whose blocks may be partitioned across GPUs (lambda replicated on cut faces).
//
// One iteration (condensed blocks S_e = F_e^T M̃_e^-1 F_e, face-block + coarse preconditioner):
//   k_cg_gemv     ft_e = S_e (p seen from the block's faces)                    scatter fused into the product
//   k_cg_q        q = D o p - own side(s) of Fbar^T..., cut faces: own part -> send buffer; p.q partial -> 1 double
//   [NCCL]        send / recv of the cut-face parts, all-reduce of p.q           (rank-local part is known before the exchange)
//   k_cg_update   alpha; q completed on cut faces; lambda += alpha p; r -= alpha q; z1 = B_ff^-1 r;
//                 partial r.z1, r.r and the coarse restriction Z^T r per face
//   k_cg_coarse   t = A_II^-1 b_I, E^T b_I (one tall matrix-vector product), then [r.z1, r.r, b_I.t, y] -> reduction buffer
//   [NCCL]        ONE all-reduce of 3 + (coarse dofs on cut faces) doubles
//   k_cg_scalars  c_G = S_G^-1 y, r.z = r.z1 + b_I.t + y.c_G, beta, convergence flag, status word for the host
//   k_cg_p        z = z1 + Z c (c_I = t - E c_G), p = z + beta p
// All scalars live in device memory (CgState); the host only enqueues and polls a mapped status word every few
// iterations.  Once `done` is set every kernel returns at once, so iterations enqueued ahead are no-ops.
#pragma once
#include "k_solve.cuh"

namespace hsbp {

struct LamFaceX {          // per face that carries lambda, next to LamFace
  int64_t msg_off;         // offset of the face in the send / recv buffers (cut faces), -1 otherwise
  int64_t binv_off;        // explicit inverse of the diagonal block B_ff (first-level preconditioner), -1: Jacobi with D
  int64_t gamma;           // global index among all cut faces of the mesh, -1
  int32_t binv_ld;
  int32_t owned;           // this rank counts the face in inner products (uncut, or the minus side lives here)
  int32_t cI;              // index among this rank's uncut lambda faces, -1 for a cut face
  int32_t cG;              // index among this rank's cut faces, -1 for an uncut face
};

struct CgState {
  double rz, b2, rr, beta, tol2, rr_last;
  int32_t iter, done, converged, maxit, init, done_iter;
  unsigned int ticket[4];
};
struct CgStatus {          // mapped host memory, written by k_cg_scalars
  int32_t iter, done_iter, converged, pad;
  double rr, b2;
};

__device__ __forceinline__ double legendre_mode(int m, int n, int nl) {
  const double s = nl > 1 ? -1.0 + 2.0 * (double)n / (double)(nl - 1) : 0.0;
  return m == 0 ? 1.0 : (m == 1 ? s : 0.5 * (3.0 * s * s - 1.0));
}

// true for exactly one CTA of the grid: the one that arrives last (all writes of the others are visible to it)
__device__ __forceinline__ bool last_cta(unsigned int *ticket, unsigned int ncta) {
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    last = (t == ncta - 1);
    if (last) *ticket = 0;
  }
  __syncthreads();
  if (last) __threadfence();
  return last;
}

// x seen from every block-face point: v[fi] = x[f2l[fi]] (0 where the face carries no lambda)
__global__ void k_cg_scatter(const CgState *__restrict__ st, int force, int64_t n, const int64_t *__restrict__ f2l,
                             const double *__restrict__ x, double *__restrict__ v) {
  if (!force && st->done) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = f2l[i];
    v[i] = l >= 0 ? x[l] : 0.0;
  }
}

// ft_e = S_e v_e with v gathered from the lambda vector x through f2l; one warp per row (S_e is symmetric: a row is a
// contiguous column), grid = (row groups, blocks).  Algorithmic traffic 8 nf^2 bytes per block.
__global__ void __launch_bounds__(256)
k_cg_gemv(const CgState *__restrict__ st, int force, const BlockDesc *__restrict__ desc, const int64_t *__restrict__ soff,
          const double *__restrict__ S, const int64_t *__restrict__ f2l, const double *__restrict__ x, double *__restrict__ ft) {
  if (!force && st->done) return;
  extern __shared__ double sv[];
  const BlockDesc d = desc[blockIdx.y];
  const int nf = block_nf(d);
  const double *Sb = S + soff[blockIdx.y];
  for (int c = threadIdx.x; c < nf; c += blockDim.x) {
    const int64_t l = f2l[d.foff + c];
    sv[c] = l >= 0 ? x[l] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = blockIdx.x * nw + wid; r < nf; r += gridDim.x * nw) {
    const double *col = Sb + (int64_t)nf * r;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = lane;
    for (; c + 96 < nf; c += 128) {
      s0 += col[c] * sv[c]; s1 += col[c + 32] * sv[c + 32]; s2 += col[c + 64] * sv[c + 64]; s3 += col[c + 96] * sv[c + 96];
    }
    for (; c < nf; c += 32) s0 += col[c] * sv[c];
    const double s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) ft[d.foff + r] = s;
  }
}

// One CTA per lambda face.  mode 0 (matvec): base = D o x, out = base - own, partial of x.(owned base - own);
// mode 1 (right-hand side): base = given vector, out = base - own.  Cut faces: own -> send buffer, out = base
// (completed by k_cg_finish / k_cg_update after the exchange).  The last CTA adds the partials up in face order.
__global__ void __launch_bounds__(128)
k_cg_q(CgState *__restrict__ st, int force, int mode, const LamFace *__restrict__ lf, const LamFaceX *__restrict__ lx,
       const double *__restrict__ D, const double *__restrict__ x, const double *__restrict__ basev,
       const double *__restrict__ ft, double *__restrict__ out, double *__restrict__ send,
       double *__restrict__ partial, double *__restrict__ red) {
  if (!force && st->done) return;
  __shared__ double scratch[32];
  const LamFace f = lf[blockIdx.x];
  const LamFaceX fx = lx[blockIdx.x];
  double s = 0.0;
  for (int n = threadIdx.x; n < f.nl; n += blockDim.x) {
    double own = 0.0;
    if (f.em >= 0) own = ft[f.fm + n];
    if (f.ep >= 0) own += ft[f.fp + (f.flip ? f.nl - 1 - n : n)];
    const double xn = mode == 0 ? x[f.loff + n] : 0.0;
    const double base = mode == 0 ? D[f.loff + n] * xn : basev[f.loff + n];
    if (fx.msg_off >= 0) { send[fx.msg_off + n] = own; out[f.loff + n] = base; }
    else out[f.loff + n] = base - own;
    s += xn * ((fx.owned ? base : 0.0) - own);
  }
  if (mode != 0) return;
  s = cta_sum(s, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
  if (last_cta(&st->ticket[0], gridDim.x)) {
    double t = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += partial[i];
    t = cta_sum(t, scratch);
    if (threadIdx.x == 0) red[0] = t;
  }
}

// send buffer <- x on the cut faces
__global__ void __launch_bounds__(128)
k_cg_pack(const LamFace *__restrict__ lf, const LamFaceX *__restrict__ lx, const double *__restrict__ x, double *__restrict__ send) {
  const LamFace f = lf[blockIdx.x];
  const LamFaceX fx = lx[blockIdx.x];
  if (fx.msg_off < 0) return;
  for (int n = threadIdx.x; n < f.nl; n += blockDim.x) send[fx.msg_off + n] = x[f.loff + n];
}

// out = out - (send + recv) on the cut faces (the sum of the two sides is commutative: both ranks get the same bits);
// mode 1: out = send + recv (completion of D)
__global__ void __launch_bounds__(128)
k_cg_finish(const LamFace *__restrict__ lf, const LamFaceX *__restrict__ lx, const double *__restrict__ send,
            const double *__restrict__ recv, double *__restrict__ out, int mode) {
  const LamFace f = lf[blockIdx.x];
  const LamFaceX fx = lx[blockIdx.x];
  if (fx.msg_off < 0) return;
  for (int n = threadIdx.x; n < f.nl; n += blockDim.x) {
    const double both = send[fx.msg_off + n] + recv[fx.msg_off + n];
    out[f.loff + n] = mode == 1 ? both : out[f.loff + n] - both;
  }
}

// One CTA per lambda face: step along p, first-level preconditioner, per-face partial sums.
//   bc: coarse restriction, uncut faces at bI[q cI + m], cut faces at bG[q cG + m] (counted by the owner only)
//   facepart[2 f], [2 f + 1]: owner-counted r.z1, r.r
__global__ void __launch_bounds__(256)
k_cg_update(CgState *__restrict__ st, int force, const LamFace *__restrict__ lf, const LamFaceX *__restrict__ lx,
            const double *__restrict__ D, const double *__restrict__ binv, const double *__restrict__ red1,
            const double *__restrict__ p, const double *__restrict__ q, const double *__restrict__ send,
            const double *__restrict__ recv, double *__restrict__ lam, double *__restrict__ r, double *__restrict__ z,
            int modes, double *__restrict__ bI, double *__restrict__ bG, double *__restrict__ facepart, double *__restrict__ red2) {
  if (!force && st->done) return;
  extern __shared__ double sm[];
  __shared__ double scratch[32];
  const LamFace f = lf[blockIdx.x];
  const LamFaceX fx = lx[blockIdx.x];
  const int nl = f.nl, tid = threadIdx.x, T = blockDim.x;
  double *rs = sm, *zs = sm + nl, *part = sm + 2 * nl;
  const bool init = st->init != 0 || (force & 2);     // force bit 1: apply the preconditioner only, no step
  const double alpha = init ? 0.0 : st->rz / red1[0];
  for (int n = tid; n < nl; n += T) {
    double rn = r[f.loff + n];
    if (!init) {
      double qn = q[f.loff + n];
      if (fx.msg_off >= 0) qn -= send[fx.msg_off + n] + recv[fx.msg_off + n];
      lam[f.loff + n] += alpha * p[f.loff + n];
      rn -= alpha * qn;
      r[f.loff + n] = rn;
    }
    rs[n] = rn;
  }
  __syncthreads();
  if (fx.binv_off >= 0) {
    const double *A = binv + fx.binv_off;
    const int ld = fx.binv_ld;
    int G = 1;
    while (2 * G * nl <= T) G *= 2;
    if (G == 1) {
      for (int i = tid; i < nl; i += T) {
        double s0 = 0.0, s1 = 0.0;
        int j = 0;
        for (; j + 1 < nl; j += 2) { s0 += A[i + (int64_t)ld * j] * rs[j]; s1 += A[i + (int64_t)ld * (j + 1)] * rs[j + 1]; }
        if (j < nl) s0 += A[i + (int64_t)ld * j] * rs[j];
        zs[i] = s0 + s1;
      }
    } else {
      if (tid < G * nl) {
        const int g = tid / nl, i = tid - g * nl;
        double s = 0.0;
        for (int j = g; j < nl; j += G) s += A[i + (int64_t)ld * j] * rs[j];
        part[g * nl + i] = s;
      }
      __syncthreads();
      if (tid < nl) {
        double s = 0.0;
        for (int g = 0; g < G; ++g) s += part[g * nl + tid];
        zs[tid] = s;
      }
    }
  } else {
    for (int n = tid; n < nl; n += T) zs[n] = rs[n] / D[f.loff + n];
  }
  __syncthreads();
  double a = 0.0, b = 0.0, c0 = 0.0, c1 = 0.0, c2 = 0.0;
  for (int n = tid; n < nl; n += T) {
    const double rn = rs[n], zn = zs[n];
    z[f.loff + n] = zn;
    a += rn * zn; b += rn * rn;
    if (modes > 0) c0 += rn;
    if (modes > 1) c1 += legendre_mode(1, n, nl) * rn;
    if (modes > 2) c2 += legendre_mode(2, n, nl) * rn;
  }
  a = cta_sum(a, scratch); b = cta_sum(b, scratch);
  if (modes > 0) c0 = cta_sum(c0, scratch);
  if (modes > 1) c1 = cta_sum(c1, scratch);
  if (modes > 2) c2 = cta_sum(c2, scratch);
  if (tid == 0) {
    facepart[2 * blockIdx.x] = fx.owned ? a : 0.0;
    facepart[2 * blockIdx.x + 1] = fx.owned ? b : 0.0;
    if (modes > 0) {
      double *dst = fx.cI >= 0 ? bI + (int64_t)modes * fx.cI : bG + (int64_t)modes * fx.cG;
      const double w = (fx.cI >= 0 || fx.owned) ? 1.0 : 0.0;
      dst[0] = w * c0;
      if (modes > 1) dst[1] = w * c1;
      if (modes > 2) dst[2] = w * c2;
    }
  }
  if (modes == 0) {          // no coarse level: this kernel fills the reduction buffer itself
    if (last_cta(&st->ticket[1], gridDim.x)) {
      double s0 = 0.0, s1 = 0.0;
      for (int i = tid; i < (int)gridDim.x; i += T) { s0 += facepart[2 * i]; s1 += facepart[2 * i + 1]; }
      s0 = cta_sum(s0, scratch); s1 = cta_sum(s1, scratch);
      if (tid == 0) { red2[0] = s0; red2[1] = s1; red2[2] = 0.0; }
    }
  }
}

// rows 0 .. nI-1: t = A_II^-1 b_I; rows nI .. nI+nGq-1: ey = E^T b_I (column j of E is contiguous).  One warp per row.
// The last CTA fills the reduction buffer: [r.z1, r.r, b_I.t, y at the global positions of this rank's cut-face dofs].
__global__ void __launch_bounds__(256)
k_cg_coarse(CgState *__restrict__ st, int force, int nI, int ldI, int nGq, const double *__restrict__ AIIinv,
            const double *__restrict__ E, const double *__restrict__ bI, const double *__restrict__ bG,
            const int64_t *__restrict__ gidx, double *__restrict__ t, double *__restrict__ ey,
            const double *__restrict__ facepart, int nfaces, double *__restrict__ red2) {
  if (!force && st->done) return;
  __shared__ double scratch[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int row = blockIdx.x * nw + wid;
  if (row < nI + nGq) {
    const double *col = row < nI ? AIIinv + (int64_t)ldI * row : E + (int64_t)nI * (row - nI);
    double s0 = 0.0, s1 = 0.0;
    int c = lane;
    for (; c + 32 < nI; c += 64) { s0 += col[c] * bI[c]; s1 += col[c + 32] * bI[c + 32]; }
    if (c < nI) s0 += col[c] * bI[c];
    const double s = warp_sum(s0 + s1);
    if (lane == 0) { if (row < nI) t[row] = s; else ey[row - nI] = s; }
  }
  if (last_cta(&st->ticket[1], gridDim.x)) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < nfaces; i += blockDim.x) { s0 += facepart[2 * i]; s1 += facepart[2 * i + 1]; }
    for (int i = threadIdx.x; i < nI; i += blockDim.x) s2 += bI[i] * t[i];
    s0 = cta_sum(s0, scratch); s1 = cta_sum(s1, scratch); s2 = cta_sum(s2, scratch);
    if (threadIdx.x == 0) { red2[0] = s0; red2[1] = s1; red2[2] = s2; }
    for (int j = threadIdx.x; j < nGq; j += blockDim.x) red2[3 + gidx[j]] = bG[j] - ey[j];
  }
}

// c_G = S_G^-1 y (one warp per row), then the scalars of the iteration in the last CTA
__global__ void __launch_bounds__(256)
k_cg_scalars(CgState *__restrict__ st, int force, int nGt, int ldG, const double *__restrict__ SGinv,
             const double *__restrict__ red2, double *__restrict__ cG, CgStatus *__restrict__ status) {
  if (!force && st->done) return;
  __shared__ double scratch[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const double *y = red2 + 3;
  const int row = blockIdx.x * nw + wid;
  if (row < nGt) {
    const double *col = SGinv + (int64_t)ldG * row;
    double s = 0.0;
    for (int c = lane; c < nGt; c += 32) s += col[c] * y[c];
    s = warp_sum(s);
    if (lane == 0) cG[row] = s;
  }
  if (last_cta(&st->ticket[2], gridDim.x)) {
    double yc = 0.0;
    for (int i = threadIdx.x; i < nGt; i += blockDim.x) yc += y[i] * cG[i];
    yc = cta_sum(yc, scratch);
    if (threadIdx.x == 0) {
      const double rz_new = red2[0] + red2[2] + yc, rr = red2[1];
      if (force) {                         // a stand-alone preconditioner application: only r.z is recorded
        st->rr_last = rz_new;
      } else if (st->init) {
        st->b2 = rr; st->rr = rr; st->rz = rz_new; st->beta = 0.0; st->init = 0; st->iter = 0;
        if (!(rr > 0.0)) { st->done = 1; st->converged = 1; st->done_iter = 0; }
      } else {
        st->iter += 1;
        st->beta = rz_new / st->rz;
        st->rz = rz_new; st->rr = rr;
        if (rr <= st->tol2 * st->b2) { st->done = 1; st->converged = 1; st->done_iter = st->iter; }
        else if (st->iter >= st->maxit || !(rr == rr)) { st->done = 1; st->done_iter = st->iter; }
      }
      if (!force) {
        status->rr = st->rr; status->b2 = st->b2; status->converged = st->converged;
        status->done_iter = st->done ? st->done_iter : -1;
        __threadfence_system();
        *(volatile int32_t *)&status->iter = st->iter;
        __threadfence_system();
      }
    }
  }
}

// One CTA per lambda face: coarse correction and the new search direction (zonly: just z = z1 + Z c)
__global__ void __launch_bounds__(128)
k_cg_p(const CgState *__restrict__ st, int force, int zonly, const LamFace *__restrict__ lf, const LamFaceX *__restrict__ lx,
       int modes, int nI, int nGq, const double *__restrict__ t, const double *__restrict__ ET,
       const int64_t *__restrict__ gidx, const double *__restrict__ cG, double *__restrict__ z, double *__restrict__ p) {
  if (!force && st->done) return;
  __shared__ double cf[4];
  const LamFace f = lf[blockIdx.x];
  const LamFaceX fx = lx[blockIdx.x];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (modes > 0 && wid < modes) {
    double c;
    if (fx.cI >= 0) {
      const int i = modes * fx.cI + wid;
      double s = 0.0;                         // (E c_G)_i over this rank's cut-face dofs
      for (int j = lane; j < nGq; j += 32) s += ET[j + (int64_t)nGq * i] * cG[gidx[j]];
      s = warp_sum(s);
      c = t[i] - s;
    } else {
      c = cG[(int64_t)modes * fx.gamma + wid];
    }
    if (lane == 0) cf[wid] = c;
  }
  __syncthreads();
  const double beta = (zonly || force) ? 0.0 : st->beta;
  for (int n = threadIdx.x; n < f.nl; n += blockDim.x) {
    double zn = z[f.loff + n];
    if (modes > 0) zn += cf[0];
    if (modes > 1) zn += legendre_mode(1, n, f.nl) * cf[1];
    if (modes > 2) zn += legendre_mode(2, n, f.nl) * cf[2];
    if (zonly) z[f.loff + n] = zn;
    else p[f.loff + n] = beta != 0.0 ? zn + beta * p[f.loff + n] : zn;      // beta == 0: p may be uninitialised
  }
}

// sum over owned faces of (b - q)^2 -> red[0] (true residual), one CTA per face
__global__ void __launch_bounds__(128)
k_cg_resid(CgState *__restrict__ st, const LamFace *__restrict__ lf, const LamFaceX *__restrict__ lx,
           const double *__restrict__ b, const double *__restrict__ q, double *__restrict__ partial, double *__restrict__ red) {
  __shared__ double scratch[32];
  const LamFace f = lf[blockIdx.x];
  const LamFaceX fx = lx[blockIdx.x];
  double s = 0.0;
  if (fx.owned)
    for (int n = threadIdx.x; n < f.nl; n += blockDim.x) { const double d = b[f.loff + n] - q[f.loff + n]; s += d * d; }
  s = cta_sum(s, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
  if (last_cta(&st->ticket[0], gridDim.x)) {
    double t = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += partial[i];
    t = cta_sum(t, scratch);
    if (threadIdx.x == 0) red[0] = t;
  }
}

// ---- coarse space setup -------------------------------------------------------------------------------------------
// block-face vector: Legendre mode m (in the lambda orientation of the face) on local face k of every block whose face
// k carries lambda, zero elsewhere
__global__ void __launch_bounds__(128)
k_coarse_unit(const BlockDesc *__restrict__ desc, const LamFace *__restrict__ lf, const int64_t *__restrict__ f2l,
              const int32_t *__restrict__ blk_lf, int k, int m, double *__restrict__ v) {
  const BlockDesc d = desc[blockIdx.x];
  const int nf = block_nf(d);
  const int lfi = blk_lf[4 * blockIdx.x + k];
  const int64_t nsp = d.Ns + 1, nrp = d.Nr + 1;
  const int64_t fs = k == 0 ? 0 : k == 1 ? nsp : k == 2 ? 2 * nsp : 2 * nsp + nrp;
  const int fn = (int)(k < 2 ? nsp : nrp);
  for (int c = threadIdx.x; c < nf; c += blockDim.x) {
    double val = 0.0;
    if (lfi >= 0 && c >= fs && c < fs + fn) {
      const LamFace f = lf[lfi];
      val = legendre_mode(m, (int)(f2l[d.foff + c] - f.loff), f.nl);
    }
    v[d.foff + c] = val;
  }
}
// T_e[(k', m'), col] = (mode m' on face k') . ft_e ;  T: nblocks x (4 q) x (4 q), column-major per block
__global__ void __launch_bounds__(128)
k_coarse_project(const BlockDesc *__restrict__ desc, const LamFace *__restrict__ lf, const int64_t *__restrict__ f2l,
                 const int32_t *__restrict__ blk_lf, int modes, int col, const double *__restrict__ ft, double *__restrict__ T) {
  const BlockDesc d = desc[blockIdx.x];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nq = 4 * modes;
  const int64_t nsp = d.Ns + 1, nrp = d.Nr + 1;
  for (int row = wid; row < nq; row += nw) {
    const int k = row / modes, m = row - k * modes;
    const int lfi = blk_lf[4 * blockIdx.x + k];
    double s = 0.0;
    if (lfi >= 0) {
      const LamFace f = lf[lfi];
      const int64_t fs = d.foff + (k == 0 ? 0 : k == 1 ? nsp : k == 2 ? 2 * nsp : 2 * nsp + nrp);
      for (int n = lane; n < f.nl; n += 32) s += legendre_mode(m, (int)(f2l[fs + n] - f.loff), f.nl) * ft[fs + n];
      s = warp_sum(s);
    }
    if (lane == 0) T[(int64_t)blockIdx.x * nq * nq + row + (int64_t)nq * col] = s;
  }
}
// scatter -T_e into A_II (dense, ld ldI), A_IG (nI x nGq) and this rank's part of A_GG (nGq x nGq)
__global__ void k_coarse_assemble(const LamFaceX *__restrict__ lx, const int32_t *__restrict__ blk_lf, int modes,
                                  const double *__restrict__ T, int nI, int ldI, int nGq, double *__restrict__ AII,
                                  double *__restrict__ AIG, double *__restrict__ AGG) {
  const int nq = 4 * modes;
  for (int idx = threadIdx.x; idx < nq * nq; idx += blockDim.x) {
    const int row = idx % nq, col = idx / nq;
    const int kr = row / modes, mr = row - kr * modes, kc = col / modes, mc = col - kc * modes;
    const int fr = blk_lf[4 * blockIdx.x + kr], fc = blk_lf[4 * blockIdx.x + kc];
    if (fr < 0 || fc < 0) continue;
    const LamFaceX xr = lx[fr], xc = lx[fc];
    const double v = -T[(int64_t)blockIdx.x * nq * nq + idx];
    if (xr.cI >= 0 && xc.cI >= 0) atomicAdd(&AII[(modes * xr.cI + mr) + (int64_t)ldI * (modes * xc.cI + mc)], v);
    else if (xr.cI >= 0 && xc.cG >= 0) atomicAdd(&AIG[(modes * xr.cI + mr) + (int64_t)nI * (modes * xc.cG + mc)], v);
    else if (xr.cG >= 0 && xc.cG >= 0) atomicAdd(&AGG[(modes * xr.cG + mr) + (int64_t)nGq * (modes * xc.cG + mc)], v);
  }
}
// + Z^T D Z (owner of the face), one CTA per lambda face
__global__ void __launch_bounds__(128)
k_coarse_diag(const LamFace *__restrict__ lf, const LamFaceX *__restrict__ lx, int modes, const double *__restrict__ D,
              int ldI, int nGq, double *__restrict__ AII, double *__restrict__ AGG) {
  __shared__ double scratch[32];
  const LamFace f = lf[blockIdx.x];
  const LamFaceX fx = lx[blockIdx.x];
  if (fx.cI < 0 && !fx.owned) return;
  for (int a = 0; a < modes; ++a)
    for (int b = 0; b < modes; ++b) {
      double s = 0.0;
      for (int n = threadIdx.x; n < f.nl; n += blockDim.x) s += legendre_mode(a, n, f.nl) * legendre_mode(b, n, f.nl) * D[f.loff + n];
      s = cta_sum(s, scratch);
      if (threadIdx.x == 0) {
        if (fx.cI >= 0) atomicAdd(&AII[(modes * fx.cI + a) + (int64_t)ldI * (modes * fx.cI + b)], s);
        else atomicAdd(&AGG[(modes * fx.cG + a) + (int64_t)nGq * (modes * fx.cG + b)], s);
      }
    }
}
// this rank's part of S_Gamma (nGq x nGq) placed into the global matrix (ldG, zero elsewhere)
__global__ void k_coarse_place(int nGq, const int64_t *__restrict__ gidx, const double *__restrict__ part, int ldG, double *__restrict__ SG) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nGq * nGq) return;
  const int i = idx % nGq, j = idx / nGq;
  SG[gidx[i] + (int64_t)ldG * gidx[j]] = part[idx];
}

}  // namespace hsbp
