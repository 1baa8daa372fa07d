// K2d: batched PCG on M-tilde_e with a fast-diagonalisation (separable) preconditioner -- the local solver for
// blocks that are too large for a direct factor (256 x 256 points: 65 536 unknowns per block, 1024 blocks).
//
// The reference keeps a sparse Cholesky factor of every block (`factorization(M̃_e)`, global_curved.jl:698) and
// back-solves (`F \ g`, :734, square_circle.jl:383).  Matrix-free, the same solve is a CG on M-tilde_e; Jacobi-PCG needs
// O(N) iterations per solve.  Here the preconditioner is the exact inverse of a separable operator
//     P_e = Ar_e (x) Hs + Hr (x) As_e                                                   (tensor product, r fastest)
// whose 1-D factors are read off the matrix-free operator by collapsing it against a profile w in the other direction,
//     (I (x) w^T) M̃_e (I (x) w)   ( = Ar_e (w^T Hs w) + Hr (w^T As_e w) when M̃_e is separable ).
// For a block whose coefficients do not vary (crr, css constant, crs = 0, tau constant along faces) P_e = M̃_e; on the
// smoothly warped blocks of the synthetic mesh the coefficients vary by a few per cent inside a block and the mixed
// term is a fraction of the diagonal ones, so kappa(P^-1 M̃) = O(1) (measured 4 - 10) instead of O(N^2).
// With the generalised eigen-decompositions  Ar Vr = Hr Vr Lr,  Vr^T Hr Vr = I  (and s likewise)
//     P^-1 = (Vs (x) Vr) diag(1 / d_ij) (Vs (x) Vr)^T,   d_ij = mr_i + ms_j - c,
// i.e. for the block's residual as an (Nr+1) x (Ns+1) matrix R:  Z = Vr [ (Vr^T R Vs) o Dinv ] Vs^T  -- four dense
// GEMMs per block, hand-written: two launches of k_fdm_pair (tcgen05 TF32, TMA operands, the intermediate product stays in
// tensor memory; k_tcgemm.cuh) or k_dgemm_batched (fp64, mma.sync) -- cuBLAS only as a comparison knob for tests.
//
//   setup   4 (2 WB + 1) operator applications with coloured probe vectors (all blocks at once); symmetric
//           eigen-decompositions of H^-1/2 A H^-1/2 by the batched Jacobi kernel k_jacobi_eig (k_eig.cuh; cuSOLVER syevd as a
//           comparison knob).  The collapse against the constant carries the
//           other direction's face penalties as a large shift sigma Hr, which leaves the eigenvectors alone but would
//           make lr_i + ls_j - c a difference of large numbers; the eigenvalues are therefore Rayleigh quotients of
//           the operator collapsed against the other direction's lowest mode (a shift of the smallest eigenvalue only).
//   solve   PCG per block, all blocks at once: M̃ p (k_sweep, which also leaves p . M̃ p), update of x / r, z = P^-1 r, update of p;
//           blocks that have converged drop out of every kernel.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cublas_v2.h>
#include <cusolverDn.h>
#include "api_band.cuh"
#include "k_tcgemm.cuh"
#include "k_eig.cuh"

namespace hsbp {

// probe vector: colour (index % C == ci) along direction dir (0: r, 1: s) times a profile in the other direction --
// constant one (w == nullptr), or the block's vector w[block * ldw + index] (lowest mode of the other direction)
__global__ void k_fdm_probe(const BlockDesc *__restrict__ desc, int dir, int C, int ci, const double *__restrict__ w,
                            int64_t ldw, double *__restrict__ u) {
  const BlockDesc d = desc[blockIdx.x];
  const int Nrp = d.Nr + 1;
  const int64_t np = (int64_t)Nrp * (d.Ns + 1);
  const double *wb = w ? w + (int64_t)blockIdx.x * ldw : nullptr;
  for (int64_t idx = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; idx < np; idx += (int64_t)gridDim.y * blockDim.x) {
    const int j = (int)(idx / Nrp), i = (int)(idx - (int64_t)j * Nrp);
    const double prof = wb ? wb[dir == 0 ? j : i] : 1.0;
    u[d.voff + idx] = ((dir == 0 ? i : j) % C == ci) ? prof : 0.0;
  }
}

// collapse y = M-tilde u of a probe onto direction `dir` against the same profile (scale * 1 or the block's vector w)
// and store the entries of the 1-D operator it exposes: row l of the collapsed vector belongs to the one probe column
// l0 = ci (mod C) within WB of l.  A: [block][n x n] column-major (n = Nr+1 or Ns+1).
__global__ void __launch_bounds__(256)
k_fdm_collapse(const BlockDesc *__restrict__ desc, int dir, int C, int WB, int ci, const double *__restrict__ w, int64_t ldw,
               double scale, const double *__restrict__ y, double *__restrict__ A) {
  const BlockDesc d = desc[blockIdx.x];
  const int Nrp = d.Nr + 1, Nsp = d.Ns + 1;
  const double *yb = y + d.voff;
  const double *wb = w ? w + (int64_t)blockIdx.x * ldw : nullptr;
  const int n = dir == 0 ? Nrp : Nsp;
  double *Ab = A + (int64_t)blockIdx.x * n * n;
  if (dir == 0) {
    for (int i = threadIdx.x; i < Nrp; i += blockDim.x) {      // sum over s, coalesced in i
      double s = 0.0;
      for (int j = 0; j < Nsp; ++j) s += (wb ? wb[j] : scale) * yb[i + (int64_t)Nrp * j];
      const int i0 = i - WB + (((ci - (i - WB)) % C) + C) % C;
      if (i0 >= 0 && i0 < Nrp) Ab[i + (int64_t)n * i0] = s;
    }
  } else {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = wid; j < Nsp; j += nw) {                        // sum over r, one warp per line
      double s = 0.0;
      for (int i = lane; i < Nrp; i += 32) s += (wb ? wb[i] : scale) * yb[i + (int64_t)Nrp * j];
      s = warp_sum(s);
      if (lane == 0) {
        const int j0 = j - WB + (((ci - (j - WB)) % C) + C) % C;
        if (j0 >= 0 && j0 < Nsp) Ab[j + (int64_t)n * j0] = s;
      }
    }
  }
}

// mu[block][a] = V[:, a]^T T[:, a]  with T = A V: the Rayleigh quotients of the (H-orthonormal) columns of V
__global__ void __launch_bounds__(256)
k_fdm_rayleigh(int n, const double *__restrict__ V, const double *__restrict__ T, double *__restrict__ mu) {
  const double *Vb = V + (int64_t)blockIdx.x * n * n, *Tb = T + (int64_t)blockIdx.x * n * n;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int a = wid; a < n; a += nw) {
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += Vb[i + (int64_t)n * a] * Tb[i + (int64_t)n * a];
    s = warp_sum(s);
    if (lane == 0) mu[(int64_t)blockIdx.x * n + a] = s;
  }
}

// A <- H^-1/2 sym(A) H^-1/2 (standard form of the generalised problem A v = l H v); one CTA per block
template <int P>
__global__ void k_fdm_standard_form(int n, double *__restrict__ A) {
  double *Ab = A + (int64_t)blockIdx.x * n * n;
  const double h = 2.0 / (n - 1);
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx % n, k = idx / n;
    if (i < k) continue;
    const double v = 0.5 * (Ab[i + (int64_t)n * k] + Ab[k + (int64_t)n * i]) /
                     sqrt(h * hweight<P>(i, n - 1) * h * hweight<P>(k, n - 1));
    Ab[i + (int64_t)n * k] = v;
    Ab[k + (int64_t)n * i] = v;
  }
}
// eigenvectors of the standard form -> generalised eigenvectors V = H^-1/2 Q
template <int P>
__global__ void k_fdm_scale_vectors(int n, double *__restrict__ V) {
  double *Vb = V + (int64_t)blockIdx.x * n * n;
  const double h = 2.0 / (n - 1);
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx % n;
    Vb[idx] /= sqrt(h * hweight<P>(i, n - 1));
  }
}
// Dinv[i, j] = 1 / max(mr_i + ms_j - c, floor),  c = (mr_0 + ms_0) / 2: both mr_0 and ms_0 are the Rayleigh quotient of
// the lowest tensor-product mode, which the sum would otherwise count twice
__global__ void k_fdm_dinv(const BlockDesc *__restrict__ desc, const double *__restrict__ mr, const double *__restrict__ ms,
                           double *__restrict__ dinv) {
  const BlockDesc d = desc[blockIdx.x];
  const int Nrp = d.Nr + 1, Nsp = d.Ns + 1;
  const double *mrb = mr + (int64_t)blockIdx.x * Nrp, *msb = ms + (int64_t)blockIdx.x * Nsp;
  const double c = 0.5 * (mrb[0] + msb[0]);
  const double floor_ = 1e-10 * fabs(mrb[Nrp - 1] + msb[Nsp - 1]);       // ascending eigenvalue order
  const int64_t np = (int64_t)Nrp * Nsp;
  for (int64_t idx = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; idx < np; idx += (int64_t)gridDim.y * blockDim.x) {
    const int j = (int)(idx / Nrp), i = (int)(idx - (int64_t)j * Nrp);
    dinv[d.voff + idx] = 1.0 / fmax(mrb[i] + msb[j] - c, floor_);
  }
}

// PCG with a general preconditioner, part 1 (after Ap = M̃ p): x += alpha p, r -= alpha Ap, rr = r.r
__global__ void __launch_bounds__(1024)
k_fpcg_update1(const BlockDesc *__restrict__ desc, const double *__restrict__ Ap, double *__restrict__ x,
               double *__restrict__ r, const double *__restrict__ p, PcgState *__restrict__ st, double tol2,
               float *__restrict__ r32,        // r32 (optional): r rounded to TF32, the tensor-core operand of z = P^-1 r
               const double *__restrict__ dotpart, int nch, int closure_pts) {
  // dotpart (optional): p . Ap of every (block, chunk) as k_sweep left it (SweepParams::dot) -- everything but the first
  // closure_pts points of either s-end of the block, which are summed here: one pass over the vectors instead of two
  __shared__ double scratch[32];
  const BlockDesc d = desc[blockIdx.x];
  PcgState s = st[blockIdx.x];
  if (!s.active) return;
  const int64_t np = (int64_t)(d.Nr + 1) * (d.Ns + 1), o = d.voff;
  double pAp = 0;
  if (dotpart != nullptr) {
    for (int64_t i = threadIdx.x; i < 2 * (int64_t)closure_pts; i += blockDim.x) {
      const int64_t k = i < closure_pts ? i : np - 2 * (int64_t)closure_pts + i;
      pAp += p[o + k] * Ap[o + k];
    }
    if (threadIdx.x == 0)
      for (int c = 0; c < nch; ++c) pAp += dotpart[(int64_t)blockIdx.x * nch + c];
  } else {
    for (int64_t i = threadIdx.x; i < np; i += blockDim.x) pAp += p[o + i] * Ap[o + i];
  }
  pAp = cta_sum(pAp, scratch);
  const double alpha = s.rz / pAp;
  double rr = 0;
  if (((o | np) & 1) == 0) {                                     // 16-byte accesses, two independent pairs per trip
    const double2 *p2 = reinterpret_cast<const double2 *>(p + o), *A2 = reinterpret_cast<const double2 *>(Ap + o);
    double2 *x2 = reinterpret_cast<double2 *>(x + o), *r2 = reinterpret_cast<double2 *>(r + o);
    float2 *f2 = r32 ? reinterpret_cast<float2 *>(r32 + o) : nullptr;
    const int64_t n2 = np >> 1;
#pragma unroll 2
    for (int64_t i = threadIdx.x; i < n2; i += blockDim.x) {
      const double2 pv = p2[i], av = A2[i];
      double2 xv = x2[i], rv = r2[i];
      xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
      rv.x = fma(-alpha, av.x, rv.x); rv.y = fma(-alpha, av.y, rv.y);
      x2[i] = xv; r2[i] = rv;
      if (f2) f2[i] = make_float2(hsbp::tc::round_tf32((float)rv.x), hsbp::tc::round_tf32((float)rv.y));
      rr = fma(rv.x, rv.x, rr); rr = fma(rv.y, rv.y, rr);
    }
  } else {
    for (int64_t i = threadIdx.x; i < np; i += blockDim.x) {
      x[o + i] += alpha * p[o + i];
      const double ri = r[o + i] - alpha * Ap[o + i];
      r[o + i] = ri;
      if (r32) r32[o + i] = hsbp::tc::round_tf32((float)ri);
      rr += ri * ri;
    }
  }
  rr = cta_sum(rr, scratch);
  if (threadIdx.x == 0) {
    s.rr = rr; s.iters += 1; s.alpha = alpha;
    if (!(rr > tol2 * s.g2)) s.active = 2;                       // converged: part 2 retires the block
    st[blockIdx.x] = s;
  }
}
// part 2 (after z = P^-1 r): p = z + beta p with the flexible (Polak-Ribiere) beta
//   beta = z_new . (r_new - r_old) / (z_old . r_old) = -alpha (z_new . A p) / rz_old,
// which equals r_new.z_new / rz_old for an exact, fixed preconditioner and stays a descent direction when the
// preconditioner is applied in reduced precision.  init != 0: first direction p = z, rz = r.z
__global__ void __launch_bounds__(1024)
k_fpcg_update2(const BlockDesc *__restrict__ desc, const double *__restrict__ r, const double *__restrict__ z,
               const double *__restrict__ Ap, double *__restrict__ p, PcgState *__restrict__ st, int init,
               int *__restrict__ nactive) {
  __shared__ double scratch[32];
  const BlockDesc d = desc[blockIdx.x];
  PcgState s = st[blockIdx.x];
  if (!s.active) return;
  const int64_t np = (int64_t)(d.Nr + 1) * (d.Ns + 1), o = d.voff;
  if (s.active == 2) {                                           // retire: a zero direction keeps later applies harmless
    for (int64_t i = threadIdx.x; i < np; i += blockDim.x) p[o + i] = 0.0;
    if (threadIdx.x == 0) { s.active = 0; st[blockIdx.x] = s; }
    return;
  }
  double rz = 0, zAp = 0;
  const bool vec2 = ((o | np) & 1) == 0;                         // 16-byte accesses
  const double2 *z2 = reinterpret_cast<const double2 *>(z + o);
  const int64_t n2 = np >> 1;
  if (vec2) {
    const double2 *r2 = reinterpret_cast<const double2 *>(r + o), *A2 = reinterpret_cast<const double2 *>(Ap + o);
#pragma unroll 2
    for (int64_t i = threadIdx.x; i < n2; i += blockDim.x) {
      const double2 zv = z2[i], rv = r2[i];
      rz = fma(rv.x, zv.x, rz); rz = fma(rv.y, zv.y, rz);
      if (!init) { const double2 av = A2[i]; zAp = fma(zv.x, av.x, zAp); zAp = fma(zv.y, av.y, zAp); }
    }
  } else {
    for (int64_t i = threadIdx.x; i < np; i += blockDim.x) {
      const double zi = z[o + i];
      rz += r[o + i] * zi;
      if (!init) zAp += zi * Ap[o + i];
    }
  }
  rz = cta_sum(rz, scratch); zAp = cta_sum(zAp, scratch);
  const double beta = init ? 0.0 : -s.alpha * zAp / s.rz;
  if (vec2) {
    double2 *p2 = reinterpret_cast<double2 *>(p + o);
#pragma unroll 2
    for (int64_t i = threadIdx.x; i < n2; i += blockDim.x) {
      const double2 zv = z2[i];
      double2 pv = p2[i];
      pv.x = fma(beta, pv.x, zv.x); pv.y = fma(beta, pv.y, zv.y);
      p2[i] = pv;
    }
  } else {
    for (int64_t i = threadIdx.x; i < np; i += blockDim.x) p[o + i] = z[o + i] + beta * p[o + i];
  }
  if (threadIdx.x == 0) {
    s.rz = rz;
    st[blockIdx.x] = s;
    atomicAdd(nactive, 1);
  }
}
// fp64 <-> fp32 copies for the reduced-precision preconditioner
__global__ void k_f64_to_f32(int64_t n, const double *__restrict__ x, float *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = (float)x[i];
}
__global__ void k_f64_to_tf32(int64_t n, const double *__restrict__ x, float *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = hsbp::tc::round_tf32((float)x[i]);
}
__global__ void k_f32_to_f64(int64_t n, const float *__restrict__ x, double *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = (double)x[i];
}
__global__ void k_mul_f32(int64_t n, float *__restrict__ x, const float *__restrict__ d) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] *= d[i];
}
// x = 0, r = g, g2 = rr = g.g
__global__ void __launch_bounds__(1024)
k_fpcg_init(const BlockDesc *__restrict__ desc, const double *__restrict__ g, double *__restrict__ x,
            double *__restrict__ r, double *__restrict__ p, PcgState *__restrict__ st, float *__restrict__ r32) {
  __shared__ double scratch[32];
  const BlockDesc d = desc[blockIdx.x];
  const int64_t np = (int64_t)(d.Nr + 1) * (d.Ns + 1), o = d.voff;
  double gg = 0;
  for (int64_t i = threadIdx.x; i < np; i += blockDim.x) {
    const double gi = g[o + i];
    x[o + i] = 0.0; r[o + i] = gi; p[o + i] = 0.0;
    if (r32) r32[o + i] = hsbp::tc::round_tf32((float)gi);
    gg += gi * gi;
  }
  gg = cta_sum(gg, scratch);
  if (threadIdx.x == 0) {
    PcgState s; s.rz = 0.0; s.g2 = gg; s.rr = gg; s.iters = 0;
    s.active = gg > 0.0 ? 1 : 0;                                  // g == 0 -> x = 0 (global_curved.jl:733)
    st[blockIdx.x] = s;
  }
}

}  // namespace hsbp

namespace {

using namespace hsbp;

struct FdmLibs {                 // per-context library handles (created on first use)
  cublasHandle_t blas = nullptr;
  cusolverDnHandle_t solver = nullptr;
};

#define HSBP_BLAS(ctx, call)                                                                        \
  do {                                                                                              \
    cublasStatus_t _s = (call);                                                                     \
    if (_s != CUBLAS_STATUS_SUCCESS) {                                                              \
      (ctx)->err = std::string(#call) + ": cuBLAS status " + std::to_string((int)_s);                 \
      return HSBP_ERR_CUDA;                                                                         \
    }                                                                                               \
  } while (0)
#define HSBP_SOLVER(ctx, call)                                                                      \
  do {                                                                                              \
    cusolverStatus_t _s = (call);                                                                   \
    if (_s != CUSOLVER_STATUS_SUCCESS) {                                                            \
      (ctx)->err = std::string(#call) + ": cuSOLVER status " + std::to_string((int)_s);               \
      return HSBP_ERR_CUDA;                                                                         \
    }                                                                                               \
  } while (0)

// the tcgen05 kernel computes 128 x N tiles with K in blocks of 32
inline bool fdm_tc_shapes_ok(int Nrp, int Nsp) { return (Nrp == 128 || Nrp == 256) && (Nsp == 128 || Nsp == 256); }

int fdm_libs(hsbp_ctx *ctx, FdmLibs **out) {
  if (!ctx->fdm_libs) {
    FdmLibs *l = new (std::nothrow) FdmLibs();
    if (!l) HSBP_FAIL(ctx, HSBP_ERR_STATE, "out of host memory");
    ctx->fdm_libs = l;
    HSBP_BLAS(ctx, cublasCreate(&l->blas));
    HSBP_BLAS(ctx, cublasSetStream(l->blas, ctx->stream));
    HSBP_SOLVER(ctx, cusolverDnCreate(&l->solver));
    HSBP_SOLVER(ctx, cusolverDnSetStream(l->solver, ctx->stream));
  }
  *out = (FdmLibs *)ctx->fdm_libs;
  return HSBP_OK;
}

void fdm_libs_destroy(hsbp_ctx *ctx) {
  FdmLibs *l = (FdmLibs *)ctx->fdm_libs;
  if (!l) return;
  if (l->blas) cublasDestroy(l->blas);
  if (l->solver) cusolverDnDestroy(l->solver);
  delete l;
  ctx->fdm_libs = nullptr;
}

// 3-D tensor map over a batch of `rows` x `inner` fp32 matrices (inner index contiguous): box = 32 x box_rows x 1 with the
// 128-byte swizzle the tensor-core descriptors of k_fdm_pair expect
int fdm_encode_map(hsbp_ctx *ctx, void *tm_out, const float *base, int inner, int rows, int64_t nb, int box_rows) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn enc = nullptr;
  if (!enc) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
      HSBP_FAIL(ctx, HSBP_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    enc = (encode_fn)fn;
  }
  const cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)nb};
  const cuuint64_t gstr[2] = {(cuuint64_t)inner * 4, (cuuint64_t)inner * rows * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc((CUtensorMap *)tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) HSBP_FAIL(ctx, HSBP_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return HSBP_OK;
}

template <int P> int fdm_setup_p(hsbp_blocks *b) {
  hsbp_ctx *ctx = b->ctx;
  if (!b->uniform) HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "fast-diagonalisation PCG needs blocks of one size (use HSBP_LOCAL_PCG)");
  FdmLibs *libs = nullptr;
  int rc = fdm_libs(ctx, &libs);
  if (rc) return rc;
  if ((rc = local_alloc(b))) return rc;
  const int Nrp = b->max_Nr + 1, Nsp = b->max_Ns + 1;
  const int64_t nb = b->nblocks;
  const size_t vb = (size_t)b->VNp * sizeof(double);
  if (!b->d_fdm_vr) {
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_fdm_vr, (size_t)nb * Nrp * Nrp * sizeof(double)));
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_fdm_vs, (size_t)nb * Nsp * Nsp * sizeof(double)));
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_fdm_z, vb));
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_fdm_t, vb));
  }
  double *d_lr = nullptr, *d_ls = nullptr, *d_work = nullptr, *d_a2 = nullptr, *d_t2 = nullptr;
  int *d_info = nullptr;
  auto cleanup = [&]() { cudaFree(d_lr); cudaFree(d_ls); cudaFree(d_work); cudaFree(d_info); cudaFree(d_a2); cudaFree(d_t2); };
  const int nmax = std::max(Nrp, Nsp);
  HSBP_CUDA(ctx, cudaMalloc((void **)&d_lr, (size_t)nb * Nrp * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&d_ls, (size_t)nb * Nsp * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&d_info, sizeof(int)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&d_a2, (size_t)nb * nmax * nmax * sizeof(double)));
  HSBP_CUDA(ctx, cudaMalloc((void **)&d_t2, (size_t)nb * nmax * nmax * sizeof(double)));
  HSBP_CUDA(ctx, cudaMemsetAsync(b->d_fdm_vr, 0, (size_t)nb * Nrp * Nrp * sizeof(double), ctx->stream));
  HSBP_CUDA(ctx, cudaMemsetAsync(b->d_fdm_vs, 0, (size_t)nb * Nsp * Nsp * sizeof(double), ctx->stream));
  // HSBP_FDM_TIMING=1: phase times of this setup on stderr (synchronising; diagnostics only)
  const bool timing = getenv("HSBP_FDM_TIMING") != nullptr;
  auto tnow = [&]() { if (timing) cudaStreamSynchronize(ctx->stream); return std::chrono::steady_clock::now(); };
  auto tprint = [&](const char *what, std::chrono::steady_clock::time_point t0) {
    if (timing) fprintf(stderr, "[fdm_setup] %-28s %8.3f s\n", what, std::chrono::duration<double>(tnow() - t0).count());
  };
  auto tph = tnow();
  // ---- round 1: 1-D operators collapsed against the constant; their eigenvectors are the transform ---------
  // (Ar + sigma_s Hr and As + sigma_r Hs have the eigenvectors of Ar, As; the shifts sigma -- the collapsed penalty
  // terms of the other direction's faces -- are large, so the eigenvalues are taken from round 2 instead)
  constexpr int WB = BandWidth<P>::WB, C = 2 * WB + 1;
  const dim3 grid = gen_grid(b);
  double *u = b->d_pp, *y = b->d_pAp;
  for (int dir = 0; dir < 2 && rc == HSBP_OK; ++dir)
    for (int ci = 0; ci < C && rc == HSBP_OK; ++ci) {
      k_fdm_probe<<<grid, GEN_THREADS, 0, ctx->stream>>>(b->d_desc, dir, C, ci, nullptr, 0, u);
      rc = apply_async(b, u, y);
      k_fdm_collapse<<<(unsigned)nb, 256, 0, ctx->stream>>>(b->d_desc, dir, C, WB, ci, nullptr, 0, 0.5, y,
                                                           dir == 0 ? b->d_fdm_vr : b->d_fdm_vs);
    }
  if (rc == HSBP_OK) {
    k_fdm_standard_form<P><<<(unsigned)nb, 256, 0, ctx->stream>>>(Nrp, b->d_fdm_vr);
    k_fdm_standard_form<P><<<(unsigned)nb, 256, 0, ctx->stream>>>(Nsp, b->d_fdm_vs);
    rc = check_launch(ctx, "fdm probes");
  }
  if (rc) { cleanup(); return rc; }
  tprint("round 1 probes", tph); tph = tnow();
  if (b->fdm_eig_lib == 0) {
    // hand-written batched Jacobi eigensolver (k_eig.cuh), one CTA per matrix, all blocks of a direction in one launch
    int *d_bad = d_info;
    HSBP_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    for (int dir = 0; dir < 2; ++dir) {
      const int n = dir == 0 ? Nrp : Nsp;
      double *Am = dir == 0 ? b->d_fdm_vr : b->d_fdm_vs, *lm = dir == 0 ? d_lr : d_ls;
      auto go = [&](auto kern, int cb) -> int {          // widest column block whose 4 cb columns fit in shared memory
        const size_t sm = eig_smem_bytes(n, cb);
        HSBP_CUDA(ctx, hsbp_smem_optin(ctx, kern, sm));
        kern<<<(unsigned)nb, EIG_THREADS, sm, ctx->stream>>>(n, Am, lm, d_t2, 30, 1e-15, d_bad);
        return HSBP_OK;
      };
      const size_t room = (size_t)ctx->smem_optin - 2048;
      int rce;
      if (eig_smem_bytes(n, 16) <= room) rce = go(k_jacobi_eig<16>, 16);
      else if (eig_smem_bytes(n, 8) <= room) rce = go(k_jacobi_eig<8>, 8);
      else if (eig_smem_bytes(n, 4) <= room) rce = go(k_jacobi_eig<4>, 4);
      else { cleanup(); HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "fast-diagonalisation setup: block too large for the batched eigensolver"); }
      if (rce) { cleanup(); return rce; }
    }
    int bad = 0;
    cudaError_t e2 = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
    if (e2 != cudaSuccess || bad) {
      cleanup();
      HSBP_FAIL(ctx, HSBP_ERR_CUDA, e2 != cudaSuccess ? std::string("fast-diagonalisation setup: k_jacobi_eig: ") + cudaGetErrorString(e2)
                                                      : "fast-diagonalisation setup: " + std::to_string(bad) + " eigen-decompositions did not converge");
    }
  } else {
    // cuSOLVER syevd, one matrix after the other: kept only as a comparison for tests (fdm_eig_lib option)
    int lwork_r = 0, lwork_s = 0;
    cusolverStatus_t cs = cusolverDnDsyevd_bufferSize(libs->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, Nrp,
                                                      b->d_fdm_vr, Nrp, d_lr, &lwork_r);
    if (cs == CUSOLVER_STATUS_SUCCESS)
      cs = cusolverDnDsyevd_bufferSize(libs->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, Nsp, b->d_fdm_vs, Nsp,
                                       d_ls, &lwork_s);
    const int lwork = std::max(lwork_r, lwork_s);
    cudaError_t e1 = cs == CUSOLVER_STATUS_SUCCESS ? cudaMalloc((void **)&d_work, (size_t)lwork * sizeof(double)) : cudaSuccess;
    for (int64_t e = 0; e < nb && cs == CUSOLVER_STATUS_SUCCESS && e1 == cudaSuccess; ++e) {
      cs = cusolverDnDsyevd(libs->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, Nrp,
                            b->d_fdm_vr + (size_t)e * Nrp * Nrp, Nrp, d_lr + (size_t)e * Nrp, d_work, lwork, d_info);
      if (cs == CUSOLVER_STATUS_SUCCESS)
        cs = cusolverDnDsyevd(libs->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, Nsp,
                              b->d_fdm_vs + (size_t)e * Nsp * Nsp, Nsp, d_ls + (size_t)e * Nsp, d_work, lwork, d_info);
    }
    if (cs != CUSOLVER_STATUS_SUCCESS || e1 != cudaSuccess) {
      cleanup();
      HSBP_FAIL(ctx, HSBP_ERR_CUDA, "fast-diagonalisation setup: cuSOLVER syevd failed (status " + std::to_string((int)cs) + ")");
    }
  }
  tprint("eigen-decompositions", tph); tph = tnow();
  k_fdm_scale_vectors<P><<<(unsigned)nb, 256, 0, ctx->stream>>>(Nrp, b->d_fdm_vr);
  k_fdm_scale_vectors<P><<<(unsigned)nb, 256, 0, ctx->stream>>>(Nsp, b->d_fdm_vs);
  // ---- round 2: eigenvalues.  Collapse against the lowest mode w0 of the other direction (w0^T H w0 = 1):
  //   (I (x) w0^T) M̃ (I (x) w0) = Ar + (w0^T As w0) Hr,  a shift of the size of the smallest eigenvalue only;
  //   mr_a = v_a^T [.] v_a are the Rayleigh quotients of the tensor modes v_a (x) w0 (and ms_a likewise).
  for (int dir = 0; dir < 2 && rc == HSBP_OK; ++dir) {
    const int n = dir == 0 ? Nrp : Nsp;
    const double *prof = dir == 0 ? b->d_fdm_vs : b->d_fdm_vr;           // column 0 of the other direction's vectors
    const int64_t ldw = dir == 0 ? (int64_t)Nsp * Nsp : (int64_t)Nrp * Nrp;
    double *V = dir == 0 ? b->d_fdm_vr : b->d_fdm_vs;
    HSBP_CUDA(ctx, cudaMemsetAsync(d_a2, 0, (size_t)nb * n * n * sizeof(double), ctx->stream));
    for (int ci = 0; ci < C && rc == HSBP_OK; ++ci) {
      k_fdm_probe<<<grid, GEN_THREADS, 0, ctx->stream>>>(b->d_desc, dir, C, ci, prof, ldw, u);
      rc = apply_async(b, u, y);
      k_fdm_collapse<<<(unsigned)nb, 256, 0, ctx->stream>>>(b->d_desc, dir, C, WB, ci, prof, ldw, 1.0, y, d_a2);
    }
    if (rc) break;
    {                                                       // T = A V (fp64, mma.sync.m8n8k4.f64)
      hsbp::tc::DgemmParams g;
      g.A = d_a2; g.B = V; g.C = d_t2; g.scale = nullptr; g.strideA = g.strideB = g.strideC = (int64_t)n * n; g.M = g.N = g.K = n;
      g.sam = 1; g.sak = n; g.sbn = n; g.sbk = 1; g.scm = 1; g.scn = n;
      hsbp::tc::k_dgemm_batched<<<dim3((unsigned)((n + 63) / 64), (unsigned)((n + 63) / 64), (unsigned)nb), 256, 0, ctx->stream>>>(g);
    }
    k_fdm_rayleigh<<<(unsigned)nb, 256, 0, ctx->stream>>>(n, V, d_t2, dir == 0 ? d_lr : d_ls);
  }
  if (rc) { cleanup(); return rc; }
  k_fdm_dinv<<<dim3((unsigned)nb, 16), 256, 0, ctx->stream>>>(b->d_desc, d_lr, d_ls, b->d_dinv);
  tprint("round 2 (Rayleigh quotients)", tph); tph = tnow();
  if (b->fdm_gemm != 0) {                                        // fp32 copies for the TF32 application (and their transposes:
                                                                 // the tensor-core kernel wants every operand contiguous in k)
    if (!b->d_fdm_vr32) {
      cudaMalloc((void **)&b->d_fdm_vr32, (size_t)nb * Nrp * Nrp * sizeof(float));
      cudaMalloc((void **)&b->d_fdm_vs32, (size_t)nb * Nsp * Nsp * sizeof(float));
      cudaMalloc((void **)&b->d_fdm_vrT32, (size_t)nb * Nrp * Nrp * sizeof(float));
      cudaMalloc((void **)&b->d_fdm_vsT32, (size_t)nb * Nsp * Nsp * sizeof(float));
      cudaMalloc((void **)&b->d_fdm_dinv32, (size_t)b->VNp * sizeof(float));
      cudaMalloc((void **)&b->d_fdm_dinvT32, (size_t)b->VNp * sizeof(float));
      cudaMalloc((void **)&b->d_fdm_a32, (size_t)b->VNp * sizeof(float));
      cudaMalloc((void **)&b->d_fdm_b32, (size_t)b->VNp * sizeof(float));
    }
    if (!b->d_fdm_vr32 || !b->d_fdm_vs32 || !b->d_fdm_vrT32 || !b->d_fdm_vsT32 || !b->d_fdm_dinv32 || !b->d_fdm_dinvT32 || !b->d_fdm_a32 ||
        !b->d_fdm_b32) {
      cleanup();
      HSBP_FAIL(ctx, HSBP_ERR_CUDA, "fast-diagonalisation setup: out of device memory (fp32 copies)");
    }
    k_f64_to_f32<<<vec_grid((int64_t)nb * Nrp * Nrp), VEC_THREADS, 0, ctx->stream>>>((int64_t)nb * Nrp * Nrp, b->d_fdm_vr, b->d_fdm_vr32);
    k_f64_to_f32<<<vec_grid((int64_t)nb * Nsp * Nsp), VEC_THREADS, 0, ctx->stream>>>((int64_t)nb * Nsp * Nsp, b->d_fdm_vs, b->d_fdm_vs32);
    k_f64_to_f32<<<vec_grid(b->VNp), VEC_THREADS, 0, ctx->stream>>>(b->VNp, b->d_dinv, b->d_fdm_dinv32);
    if (b->fdm_gemm == 3) {                                      // static tensor-core operands: rounded to TF32 once
      hsbp::tc::k_round_tf32<<<vec_grid((int64_t)nb * Nrp * Nrp), VEC_THREADS, 0, ctx->stream>>>((int64_t)nb * Nrp * Nrp, b->d_fdm_vr32);
      hsbp::tc::k_round_tf32<<<vec_grid((int64_t)nb * Nsp * Nsp), VEC_THREADS, 0, ctx->stream>>>((int64_t)nb * Nsp * Nsp, b->d_fdm_vs32);
    }
    hsbp::tc::k_transpose_f32<<<dim3(64, (unsigned)nb), 256, 0, ctx->stream>>>(Nrp, Nrp, b->d_fdm_vr32, b->d_fdm_vrT32);
    hsbp::tc::k_transpose_f32<<<dim3(64, (unsigned)nb), 256, 0, ctx->stream>>>(Nsp, Nsp, b->d_fdm_vs32, b->d_fdm_vsT32);
    hsbp::tc::k_transpose_f32<<<dim3(64, (unsigned)nb), 256, 0, ctx->stream>>>(Nrp, Nsp, b->d_fdm_dinv32, b->d_fdm_dinvT32);
    b->fdm_tm_valid = false;
    if (b->fdm_gemm == 3 && fdm_tc_shapes_ok(Nrp, Nsp)) {       // tensor maps of k_fdm_pair: (a) Vr, r32, Vs   (b) Vr^T, T3, Vs^T
      rc = fdm_encode_map(ctx, b->fdm_tm[0], b->d_fdm_vr32, Nrp, Nrp, nb, 128);
      if (!rc) rc = fdm_encode_map(ctx, b->fdm_tm[1], b->d_fdm_a32, Nrp, Nsp, nb, Nsp);
      if (!rc) rc = fdm_encode_map(ctx, b->fdm_tm[2], b->d_fdm_vs32, Nsp, Nsp, nb, Nsp);
      if (!rc) rc = fdm_encode_map(ctx, b->fdm_tm[3], b->d_fdm_vrT32, Nrp, Nrp, nb, 128);
      if (!rc) rc = fdm_encode_map(ctx, b->fdm_tm[4], b->d_fdm_b32, Nrp, Nsp, nb, Nsp);
      if (!rc) rc = fdm_encode_map(ctx, b->fdm_tm[5], b->d_fdm_vsT32, Nsp, Nsp, nb, Nsp);
      if (rc) { cleanup(); return rc; }
      b->fdm_tm_valid = true;
    }
  }
  cudaError_t e1 = cudaGetLastError();
  if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
  tprint("fp32 copies, tensor maps", tph); tph = tnow();
  cleanup();
  tprint("cleanup (cudaFree)", tph);
  if (e1 != cudaSuccess) { ctx->err = std::string("fdm_setup: ") + cudaGetErrorString(e1); return HSBP_ERR_CUDA; }
  return HSBP_OK;
}

int fdm_setup(hsbp_blocks *b) {
  return dispatch_p(b->p, [&](auto Pc) { return fdm_setup_p<decltype(Pc)::value>(b); });
}

// z = P^-1 r for all blocks: Z = Vr [ (Vr^T R Vs) o Dinv ] Vs^T -- four batched GEMMs per application, hand-written:
//   fdm_gemm = 3 (default of the synthetic-mesh drivers)  tcgen05.mma kind::tf32, accumulators in TMEM (k_tcgemm.cuh); r is
//                 converted to fp32 while it is staged, the `o Dinv` rides in the epilogue of the second GEMM, z comes out
//                 as fp64; PCG itself (x, r, p, the operator) stays fp64 and the flexible beta keeps it convergent with the
//                 inexactly applied preconditioner.  Needs Nr+1, Ns+1 in {128, 256}.
//   fdm_gemm = 0  fp64 on the fp64 tensor pipe (mma.sync.m8n8k4.f64), any block size.
//   fdm_gemm = -1 the strided-batched cuBLAS TF32 GEMMs of round 1 -- kept only so that tests can compare the hand-written
//                 kernels with a library result; never a default.
int fdm_precondition(hsbp_blocks *b, const double *r, double *z, bool r32_ready = false, const int *active = nullptr,
                     int active_stride = 0) {
  hsbp_ctx *ctx = b->ctx;
  const int Nrp = b->max_Nr + 1, Nsp = b->max_Ns + 1;
  const int nb = (int)b->nblocks;
  const long long sv = (long long)Nrp * Nsp, sr = (long long)Nrp * Nrp, ss = (long long)Nsp * Nsp;
  if (b->fdm_gemm == 3 && b->fdm_tm_valid && b->fdm_tc_variant == 0) {
    // two launches of k_fdm_pair (k_tcgemm.cuh): T3 = ((Vr^T R) Vs) o Dinv, then Z = (Vr T3) Vs^T; r as TF32-rounded fp32 in
    // d_fdm_a32 (written by the PCG update kernels, or converted here), T3 in d_fdm_b32
    using namespace hsbp::tc;
    if (!r32_ready) k_f64_to_tf32<<<vec_grid(b->VNp), VEC_THREADS, 0, ctx->stream>>>(b->VNp, r, b->d_fdm_a32);
    PairParams pp;
    pp.active = active; pp.active_stride = active_stride; pp.M = Nrp; pp.N = Nsp; pp.strideO = sv;
    const CUtensorMap *tm = reinterpret_cast<const CUtensorMap *>(b->fdm_tm);
    const dim3 grid((unsigned)(Nrp / BM), (unsigned)nb);
    HSBP_CUDA(ctx, hsbp_smem_optin(ctx, k_fdm_pair<false>, pair_smem_bytes()));
    HSBP_CUDA(ctx, hsbp_smem_optin(ctx, k_fdm_pair<true>, pair_smem_bytes()));
    pp.out = b->d_fdm_b32; pp.scale = b->d_fdm_dinv32;
    k_fdm_pair<false><<<grid, THREADS, pair_smem_bytes(), ctx->stream>>>(tm[0], tm[1], tm[2], pp);
    pp.out = z; pp.scale = nullptr;
    k_fdm_pair<true><<<grid, THREADS, pair_smem_bytes(), ctx->stream>>>(tm[3], tm[4], tm[5], pp);
    return check_launch(ctx, "fdm_precondition (tcgen05 TF32, fused pairs)");
  }
  if (b->fdm_gemm == 3 && fdm_tc_shapes_ok(Nrp, Nsp)) {
    using namespace hsbp::tc;
    float *t1 = b->d_fdm_a32, *t3 = b->d_fdm_b32;
    auto launch = [&](const void *A, int64_t sA, int lda, const void *B, int64_t sB, int ldb, int bf64, int M, int N, int K, void *out,
                      int ldo, int mode, const float *scale) {
      GemmParams g;
      g.A = A; g.B = B; g.out = out; g.scale = scale; g.strideA = sA; g.strideB = sB; g.strideO = sv;
      g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldo = ldo; g.b_is_f64 = bf64; g.mode = mode;
      k_tc_gemm<<<dim3((unsigned)(M / BM), (unsigned)nb), THREADS, gemm_smem_bytes(N), ctx->stream>>>(g);
    };
    HSBP_CUDA(ctx, hsbp_smem_optin(ctx, k_tc_gemm, gemm_smem_bytes(256)));
    // T1 = Vr^T R            A(m, k) = Vr[k + Nrp m]   B(n, k) = R[k + Nrp n] (fp64)       -> row-major (m, n)
    launch(b->d_fdm_vr32, sr, Nrp, r, sv, Nrp, 1, Nrp, Nsp, Nrp, t1, Nsp, OUT_ROWMAJOR_F32, nullptr);
    // T3 = (T1 Vs) o Dinv    A = T1 row-major          B(n, k) = Vs[k + Nsp n]              -> row-major, scaled
    launch(t1, sv, Nsp, b->d_fdm_vs32, ss, Nsp, 0, Nrp, Nsp, Nsp, t3, Nsp, OUT_ROWMAJOR_F32_SCALED, b->d_fdm_dinvT32);
    // W = T3 Vs^T            A = T3 row-major          B(n, k) = Vs[n + Nsp k] = VsT[k + Nsp n]   -> column-major (m + Nrp n)
    launch(t3, sv, Nsp, b->d_fdm_vsT32, ss, Nsp, 0, Nrp, Nsp, Nsp, t1, Nrp, OUT_COLMAJOR_F32, nullptr);
    // Z = Vr W               A(m, k) = Vr[m + Nrp k] = VrT[k + Nrp m]   B(n, k) = W[k + Nrp n]    -> column-major fp64
    launch(b->d_fdm_vrT32, sr, Nrp, t1, sv, Nrp, 0, Nrp, Nsp, Nrp, z, Nrp, OUT_COLMAJOR_F64, nullptr);
    return check_launch(ctx, "fdm_precondition (tcgen05 TF32)");
  }
  if (b->fdm_gemm == -1) {
    FdmLibs *libs = nullptr;
    int rc = fdm_libs(ctx, &libs);
    if (rc) return rc;
    const float one = 1.0f, zero = 0.0f;
    float *a = b->d_fdm_a32, *t = b->d_fdm_b32;
    const cublasComputeType_t ct = CUBLAS_COMPUTE_32F_FAST_TF32;
    k_f64_to_f32<<<vec_grid(b->VNp), VEC_THREADS, 0, ctx->stream>>>(b->VNp, r, a);
    HSBP_BLAS(ctx, cublasGemmStridedBatchedEx(libs->blas, CUBLAS_OP_T, CUBLAS_OP_N, Nrp, Nsp, Nrp, &one, b->d_fdm_vr32, CUDA_R_32F,
                                              Nrp, sr, a, CUDA_R_32F, Nrp, sv, &zero, t, CUDA_R_32F, Nrp, sv, nb, ct,
                                              CUBLAS_GEMM_DEFAULT));
    HSBP_BLAS(ctx, cublasGemmStridedBatchedEx(libs->blas, CUBLAS_OP_N, CUBLAS_OP_N, Nrp, Nsp, Nsp, &one, t, CUDA_R_32F, Nrp, sv,
                                              b->d_fdm_vs32, CUDA_R_32F, Nsp, ss, &zero, a, CUDA_R_32F, Nrp, sv, nb, ct,
                                              CUBLAS_GEMM_DEFAULT));
    k_mul_f32<<<vec_grid(b->VNp), VEC_THREADS, 0, ctx->stream>>>(b->VNp, a, b->d_fdm_dinv32);
    HSBP_BLAS(ctx, cublasGemmStridedBatchedEx(libs->blas, CUBLAS_OP_N, CUBLAS_OP_N, Nrp, Nsp, Nrp, &one, b->d_fdm_vr32, CUDA_R_32F,
                                              Nrp, sr, a, CUDA_R_32F, Nrp, sv, &zero, t, CUDA_R_32F, Nrp, sv, nb, ct,
                                              CUBLAS_GEMM_DEFAULT));
    HSBP_BLAS(ctx, cublasGemmStridedBatchedEx(libs->blas, CUBLAS_OP_N, CUBLAS_OP_T, Nrp, Nsp, Nsp, &one, t, CUDA_R_32F, Nrp, sv,
                                              b->d_fdm_vs32, CUDA_R_32F, Nsp, ss, &zero, a, CUDA_R_32F, Nrp, sv, nb, ct,
                                              CUBLAS_GEMM_DEFAULT));
    k_f32_to_f64<<<vec_grid(b->VNp), VEC_THREADS, 0, ctx->stream>>>(b->VNp, a, z);
    return check_launch(ctx, "fdm_precondition (cuBLAS, comparison only)");
  }
  // fp64 on the fp64 tensor pipe; all four operands are used where they lie (general strides)
  {
    using namespace hsbp::tc;
    double *t = b->d_fdm_t;
    auto launch = [&](const double *A, int64_t sA, int64_t sam, int64_t sak, const double *B, int64_t sB, int64_t sbn, int64_t sbk, int M,
                      int N, int K, double *Cc, const double *scale) {
      DgemmParams g;
      g.A = A; g.B = B; g.C = Cc; g.scale = scale; g.strideA = sA; g.strideB = sB; g.strideC = sv; g.M = M; g.N = N; g.K = K;
      g.sam = sam; g.sak = sak; g.sbn = sbn; g.sbk = sbk; g.scm = 1; g.scn = M;
      k_dgemm_batched<<<dim3((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64), (unsigned)nb), 256, 0, ctx->stream>>>(g);
    };
    launch(b->d_fdm_vr, sr, Nrp, 1, r, sv, Nrp, 1, Nrp, Nsp, Nrp, t, nullptr);                      // T1 = Vr^T R
    launch(t, sv, 1, Nrp, b->d_fdm_vs, ss, Nsp, 1, Nrp, Nsp, Nsp, z, b->d_dinv);                  // T3 = (T1 Vs) o Dinv
    launch(z, sv, 1, Nrp, b->d_fdm_vs, ss, 1, Nsp, Nrp, Nsp, Nsp, t, nullptr);                     // W = T3 Vs^T
    launch(b->d_fdm_vr, sr, 1, Nrp, t, sv, Nrp, 1, Nrp, Nsp, Nrp, z, nullptr);                     // Z = Vr W
    return check_launch(ctx, "fdm_precondition (fp64 DMMA)");
  }
}

int fdm_solve(hsbp_blocks *b, const double *g, double *x, hsbp_local_stats *stats) {
  hsbp_ctx *ctx = b->ctx;
  if (!b->d_fdm_vr) HSBP_FAIL(ctx, HSBP_ERR_STATE, "fast-diagonalisation PCG: not set up");
  const double tol2 = b->local_tol * b->local_tol;
  PcgState *st = (PcgState *)b->d_pcg;
  double *r = b->d_pr, *p = b->d_pp, *Ap = b->d_pAp, *z = b->d_fdm_z;
  const unsigned nb = (unsigned)b->nblocks;
  int rc;
  // converged blocks drop out of every kernel of the iteration: the update kernels, the operator apply and the preconditioner
  // all look at PcgState::active
  const int *act = reinterpret_cast<const int *>(reinterpret_cast<const char *>(st) + offsetof(PcgState, active));
  const int act_stride = (int)(sizeof(PcgState) / sizeof(int));
  const bool pair = b->fdm_gemm == 3 && b->fdm_tm_valid && b->fdm_tc_variant == 0;
  float *r32 = pair ? b->d_fdm_a32 : nullptr;
  struct SkipGuard {                                   // hsbp_apply of other callers must see every block again
    hsbp_blocks *b;
    ~SkipGuard() { b->skip_flags = nullptr; b->skip_stride = 0; b->sweep_dot_out = nullptr; }
  } guard{b};
  k_fpcg_init<<<nb, 1024, 0, ctx->stream>>>(b->d_desc, g, x, r, p, st, r32);
  if ((rc = fdm_precondition(b, r, z, pair, act, act_stride))) return rc;
  if (!b->fdm_no_skip) { b->skip_flags = act; b->skip_stride = act_stride; }
  // p . Ap comes out of the sweep kernel when it produces the final y in one pass (fused faces)
  const bool fused_dot = !b->fdm_no_fused_dot && sweep_dot_eligible(b);
  if (fused_dot) {
    if (!b->d_sweep_dot) HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_sweep_dot, (size_t)b->nblocks * 65 * sizeof(double)));
    b->sweep_dot_out = b->d_sweep_dot;
  }
  const int closure_pts = dispatch_p(b->p, [&](auto Pc) { return SweepTab<decltype(Pc)::value>::BM; }) * (b->max_Nr + 1);
  HSBP_CUDA(ctx, cudaMemsetAsync(b->d_nactive, 0, 2 * sizeof(int), ctx->stream));
  k_fpcg_update2<<<nb, 1024, 0, ctx->stream>>>(b->d_desc, r, z, Ap, p, st, 1, b->d_nactive);
  const int check_every = 4;
  int h_active = 1;
  int64_t it = 0;
  while (it < b->local_maxit) {
    int slot = 0;
    for (int k = 0; k < check_every && it < b->local_maxit; ++k, ++it) {
      if ((rc = apply_async(b, p, Ap))) return rc;
      k_fpcg_update1<<<nb, 1024, 0, ctx->stream>>>(b->d_desc, Ap, x, r, p, st, tol2, r32, fused_dot ? b->d_sweep_dot : nullptr,
                                                   b->sweep_nch, closure_pts);
      if ((rc = fdm_precondition(b, r, z, pair, b->fdm_no_skip ? nullptr : act, act_stride))) return rc;
      slot = (int)(it & 1);
      HSBP_CUDA(ctx, cudaMemsetAsync(b->d_nactive + slot, 0, sizeof(int), ctx->stream));
      k_fpcg_update2<<<nb, 1024, 0, ctx->stream>>>(b->d_desc, r, z, Ap, p, st, 0, b->d_nactive + slot);
    }
    if ((rc = check_launch(ctx, "k_fpcg_update"))) return rc;
    HSBP_CUDA(ctx, cudaMemcpyAsync(&h_active, b->d_nactive + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_active == 0) break;
  }
  if (stats) {
    std::vector<PcgState> hs(b->nblocks);
    HSBP_CUDA(ctx, cudaMemcpyAsync(hs.data(), st, b->nblocks * sizeof(PcgState), cudaMemcpyDeviceToHost, ctx->stream));
    HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    hsbp_local_stats s = {0, 0, 0, 0.0};
    for (auto &q : hs) {
      s.iterations_max = std::max<int64_t>(s.iterations_max, q.iters);
      s.iterations_sum += q.iters;
      s.failed_blocks += q.active ? 1 : 0;
      if (q.g2 > 0) s.max_rel_residual = std::max(s.max_rel_residual, sqrt(q.rr / q.g2));
    }
    *stats = s;
  }
  return HSBP_OK;
}

}  // namespace

extern "C" {

// z = P^-1 r of the fast-diagonalisation preconditioner alone (testing / profiling hook; HSBP_LOCAL_FDM must be set up)
int hsbp_local_precondition(hsbp_blocks *b, const double *r_dev, double *z_dev) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!r_dev || !z_dev || r_dev == z_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_local_precondition: bad pointers");
  if (b->local_mode != HSBP_LOCAL_FDM || !b->d_fdm_vr) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_local_precondition: set up HSBP_LOCAL_FDM first");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  return fdm_precondition(b, r_dev, z_dev);
}

}  // extern "C"

