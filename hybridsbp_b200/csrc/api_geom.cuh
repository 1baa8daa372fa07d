// C-ABI: block geometry on the device -- transfinite blend, curvilinear metric terms, and the synthetic warped mesh.
//
// reference: transfinite_blend (global_curved.jl:19-51), create_metrics (:136-209).  The reference evaluates them on the host
// with callbacks per block; at config-4 scale that is O(VNp) host work plus 24 B per point of host-to-device traffic.  Here the
// host only supplies the O(N) edge curves (or nothing at all for the analytic synthetic mesh), everything per grid point is a
// kernel.  The host-side mirrors (hybridsbp_b200/host.py) stay for meshes whose maps are arbitrary callbacks.
#pragma once
#include "hsbp_internal.h"
#include "k_generic.cuh"

namespace hsbp {

// x, x_r, x_s of every block from its four edge curves sampled at the grid points.  edges: per block (block-face layout twice)
//   [a1(s_j) | a2(s_j) | a3(r_i) | a4(r_i)]  then  [a1'(s_j) | a2'(s_j) | a3'(r_i) | a4'(r_i)],   edge k = local face k
__global__ void __launch_bounds__(GEN_THREADS)
k_blend(const BlockDesc *__restrict__ desc, const double *__restrict__ edges, double *__restrict__ x, double *__restrict__ xr,
        double *__restrict__ xs, int *__restrict__ bad) {
  const BlockDesc d = desc[blockIdx.x];
  const int Nrp = d.Nr + 1, Nsp = d.Ns + 1;
  const int64_t np = (int64_t)Nrp * Nsp;
  const double *a1 = edges + 2 * d.foff, *a2 = a1 + Nsp, *a3 = a2 + Nsp, *a4 = a3 + Nrp;
  const double *a1s = a4 + Nrp, *a2s = a1s + Nsp, *a3r = a2s + Nsp, *a4r = a3r + Nrp;
  const double c11 = a1[0], c21 = a2[0], c12 = a1[d.Ns], c22 = a2[d.Ns];
  if (blockIdx.y == 0 && threadIdx.x == 0) {          // the edge curves must meet at the corners (global_curved.jl:25)
    const double sc = fmax(fmax(fabs(c11), fabs(c21)), fmax(fabs(c12), fabs(c22))) + 1e-300;
    const double dev = fmax(fmax(fabs(c11 - a3[0]), fabs(c21 - a3[d.Nr])), fmax(fabs(c12 - a4[0]), fabs(c22 - a4[d.Nr])));
    if (!(dev <= 1.5e-8 * sc + 1e-8)) atomicExch(bad, 1);    // isapprox-like: relative sqrt(eps), small absolute slack
  }
  for (int64_t idx = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; idx < np; idx += (int64_t)gridDim.y * blockDim.x) {
    const int j = (int)(idx / Nrp), i = (int)(idx - (int64_t)j * Nrp);
    const double r = -1.0 + 2.0 * (double)i / (double)d.Nr, s = -1.0 + 2.0 * (double)j / (double)d.Ns;
    const double rp = 1.0 + r, rm = 1.0 - r, sp = 1.0 + s, sm = 1.0 - s;
    const int64_t o = d.voff + idx;
    x[o] = (rp * a2[j] + rm * a1[j] + sp * a4[i] + sm * a3[i]) / 2.0 - (rp * sp * c22 + rm * sp * c12 + rp * sm * c21 + rm * sm * c11) / 4.0;
    xr[o] = (a2[j] - a1[j] + sp * a4r[i] + sm * a3r[i]) / 2.0 - (sp * (c22 - c12) + sm * (c21 - c11)) / 4.0;
    xs[o] = (rp * a2s[j] + rm * a1s[j] + a4[i] - a3[i]) / 2.0 - (rp * (c22 - c21) + rm * (c12 - c11)) / 4.0;
  }
}

// create_metrics (global_curved.jl:154-162): J, the contravariant terms and the coefficient tensor of the transformed Laplacian
__global__ void __launch_bounds__(256)
k_metrics(int64_t n, const double *__restrict__ xr, const double *__restrict__ xs, const double *__restrict__ yr,
          const double *__restrict__ ys, double *__restrict__ crr, double *__restrict__ css, double *__restrict__ crs,
          double *__restrict__ Jout, int *__restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double a = xr[i], b = xs[i], c = yr[i], e = ys[i];
    const double J = a * e - b * c;
    if (!(J > 0.0)) atomicExch(bad, 1);                       // @assert minimum(J) > 0 (global_curved.jl:157)
    const double rx = e / J, sx = -c / J, ry = -b / J, sy = a / J;
    crr[i] = J * (rx * rx + ry * ry);
    crs[i] = J * (sx * rx + sy * ry);
    css[i] = J * (sx * sx + sy * sy);
    if (Jout) Jout[i] = J;
  }
}

// surface Jacobian and outward unit normal of the four faces (global_curved.jl:166-200), block-face layout; grid.x = 4 * nblocks
__global__ void __launch_bounds__(128)
k_face_normals(const BlockDesc *__restrict__ desc, const double *__restrict__ xr, const double *__restrict__ xs,
               const double *__restrict__ yr, const double *__restrict__ ys, double *__restrict__ sJ, double *__restrict__ nx,
               double *__restrict__ ny) {
  const int e = blockIdx.x >> 2, k = blockIdx.x & 3;
  const BlockDesc d = desc[e];
  const FaceGeom fg = face_geom(d, k);
  for (int n = threadIdx.x; n < fg.nf; n += blockDim.x) {
    const int64_t v = d.voff + face_vol(d, k, n, 0);
    double a, b;
    switch (k) {
      case 0: a = -ys[v]; b = xs[v]; break;
      case 1: a = ys[v]; b = -xs[v]; break;
      case 2: a = yr[v]; b = -xr[v]; break;
      default: a = -yr[v]; b = xr[v]; break;
    }
    const double l = hypot(a, b);
    const int64_t fi = d.foff + fg.fstart + n;
    sJ[fi] = l; nx[fi] = a / l; ny[fi] = b / l;
  }
}

// the synthetic warped mesh of SURVEY.md section 8d (hybridsbp_b200/synthetic.py): block (bx0 + e % nbx, e / nbx) of a grid of
// unit blocks,  x = xi + A sin(k xi) sin(k eta),  y = eta - A sin(k xi) sin(k eta),  k = 2 pi / L
__global__ void __launch_bounds__(GEN_THREADS)
k_synthetic_warp(const BlockDesc *__restrict__ desc, int nbx, int bx0, double L, double A, double *__restrict__ x,
                 double *__restrict__ y, double *__restrict__ xr, double *__restrict__ xs, double *__restrict__ yr,
                 double *__restrict__ ys) {
  const BlockDesc d = desc[blockIdx.x];
  const int Nrp = d.Nr + 1, Nsp = d.Ns + 1;
  const int64_t np = (int64_t)Nrp * Nsp;
  const int bx = bx0 + (int)(blockIdx.x % nbx), by = (int)(blockIdx.x / nbx);
  const double kk = 2.0 * 3.14159265358979323846 / L;
  for (int64_t idx = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; idx < np; idx += (int64_t)gridDim.y * blockDim.x) {
    const int j = (int)(idx / Nrp), i = (int)(idx - (int64_t)j * Nrp);
    const double r = -1.0 + 2.0 * (double)i / (double)d.Nr, s = -1.0 + 2.0 * (double)j / (double)d.Ns;
    const double xi = bx + (r + 1.0) / 2.0, et = by + (s + 1.0) / 2.0;
    double sx_, cx_, se_, ce_;
    sincos(kk * xi, &sx_, &cx_);
    sincos(kk * et, &se_, &ce_);
    const double w = A * sx_ * se_, wr = (A * kk / 2.0) * cx_ * se_, ws = (A * kk / 2.0) * sx_ * ce_;
    const int64_t o = d.voff + idx;
    if (x) { x[o] = xi + w; y[o] = et - w; }
    xr[o] = 0.5 + wr; xs[o] = ws; yr[o] = -wr; ys[o] = 0.5 - ws;
  }
}

}  // namespace hsbp

namespace {

int geom_set_metrics(hsbp_blocks *b, const double *xr, const double *xs, const double *yr, const double *ys, double *J, double *sJ,
                     double *nx, double *ny) {
  using namespace hsbp;
  hsbp_ctx *ctx = b->ctx;
  int *d_bad = nullptr;
  HSBP_CUDA(ctx, cudaMalloc(&d_bad, sizeof(int)));
  HSBP_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  k_metrics<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(b->VNp, xr, xs, yr, ys, b->d_crr, b->d_css, b->d_crs, J, d_bad);
  if (sJ && nx && ny) k_face_normals<<<(unsigned)(4 * b->nblocks), 128, 0, ctx->stream>>>(b->d_desc, xr, xs, yr, ys, sJ, nx, ny);
  int bad = 0;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_bad);
  if (e != cudaSuccess) { ctx->err = std::string("create_metrics on the device: ") + cudaGetErrorString(e); return HSBP_ERR_CUDA; }
  if (bad) HSBP_FAIL(ctx, HSBP_ERR_ARG, "non-positive Jacobian (global_curved.jl:157)");
  b->have_metrics = true;
  b->sweep_scaled_valid = false;
  b->rim_valid = false;
  operator_changed(b);
  return HSBP_OK;
}

}  // namespace

extern "C" {

int hsbp_blocks_blend_dev(hsbp_blocks *b, const double *edges_dev, double *x_dev, double *xr_dev, double *xs_dev) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!edges_dev || !x_dev || !xr_dev || !xs_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_blocks_blend_dev: null pointer");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  int *d_bad = nullptr;
  HSBP_CUDA(ctx, cudaMalloc(&d_bad, sizeof(int)));
  HSBP_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  hsbp::k_blend<<<gen_grid(b), hsbp::GEN_THREADS, 0, ctx->stream>>>(b->d_desc, edges_dev, x_dev, xr_dev, xs_dev, d_bad);
  int bad = 0;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_bad);
  if (e != cudaSuccess) { ctx->err = std::string("hsbp_blocks_blend_dev: ") + cudaGetErrorString(e); return HSBP_ERR_CUDA; }
  if (bad) HSBP_FAIL(ctx, HSBP_ERR_ARG, "edge curves do not meet at the corners (global_curved.jl:25)");
  return HSBP_OK;
}

int hsbp_blocks_set_geometry_dev(hsbp_blocks *b, const double *xr_dev, const double *xs_dev, const double *yr_dev, const double *ys_dev,
                                 double *J_dev, double *sJ_dev, double *nx_dev, double *ny_dev) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!xr_dev || !xs_dev || !yr_dev || !ys_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_blocks_set_geometry_dev: null pointer");
  if ((sJ_dev || nx_dev || ny_dev) && !(sJ_dev && nx_dev && ny_dev)) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_blocks_set_geometry_dev: sJ, nx, ny go together");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  return geom_set_metrics(b, xr_dev, xs_dev, yr_dev, ys_dev, J_dev, sJ_dev, nx_dev, ny_dev);
}

int hsbp_blocks_set_synthetic_warp(hsbp_blocks *b, int64_t nbx, int64_t bx0, double L, double A, double *x_dev, double *y_dev) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (nbx < 1 || !(L > 0) || b->nblocks % nbx != 0 || ((x_dev == nullptr) != (y_dev == nullptr)))
    HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_blocks_set_synthetic_warp: bad arguments");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  // the four derivative fields live in the scratch vectors of the generic kernels plus two temporaries
  double *t4[4] = {nullptr, nullptr, nullptr, nullptr};
  const size_t vb = (size_t)b->VNp * sizeof(double);
  for (int i = 0; i < 4; ++i)
    if (cudaMalloc((void **)&t4[i], vb) != cudaSuccess) {
      for (int j = 0; j < 4; ++j) cudaFree(t4[j]);
      HSBP_FAIL(ctx, HSBP_ERR_CUDA, "hsbp_blocks_set_synthetic_warp: out of device memory");
    }
  hsbp::k_synthetic_warp<<<gen_grid(b), hsbp::GEN_THREADS, 0, ctx->stream>>>(b->d_desc, (int)nbx, (int)bx0, L, A, x_dev, y_dev, t4[0], t4[1],
                                                                          t4[2], t4[3]);
  int rc = check_launch(ctx, "k_synthetic_warp");
  if (rc == HSBP_OK) rc = geom_set_metrics(b, t4[0], t4[1], t4[2], t4[3], nullptr, nullptr, nullptr, nullptr);
  for (int j = 0; j < 4; ++j) cudaFree(t4[j]);
  return rc;
}

}  // extern "C"
