// libhsbp: C-ABI entry points (include/hsbp.h).  Unity build: the kernels are header files.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>

#include "hsbp_internal.h"
#include "k_generic.cuh"
#include "k_sweep.cuh"
#include "k_solve.cuh"

using namespace hsbp;

namespace {
thread_local std::string g_noctx_err;

template <class F> int dispatch_p(int p, F &&f) {
  switch (p) {
#ifndef HSBP_ONLY_P4          // experiment builds: one order only, to cut the compile time
    case 2: return f(std::integral_constant<int, 2>{});
    case 6: return f(std::integral_constant<int, 6>{});
#endif
    case 4: return f(std::integral_constant<int, 4>{});
  }
  return HSBP_ERR_UNSUPP;
}

void fdm_libs_destroy(hsbp_ctx *ctx);      // api_fdm.cuh

// M-tilde changed (metrics, boundary conditions or tau): everything derived from the old operator is stale -- local
// factors must be set up again, a trace refuses its condensed blocks / preconditioners until they are rebuilt
void operator_changed(hsbp_blocks *b) {
  b->generation += 1;
  b->local_mode = 0;
}

int check_launch(hsbp_ctx *ctx, const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
    return HSBP_ERR_CUDA;
  }
  return HSBP_OK;
}
}  // namespace

extern "C" {

int hsbp_version(void) { return 200; }

int hsbp_ctx_create(int device, hsbp_ctx **out) {
  if (!out) return HSBP_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return HSBP_ERR_CUDA;
  hsbp_ctx *ctx = new (std::nothrow) hsbp_ctx();
  if (!ctx) return HSBP_ERR_STATE;
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
      prop.major < 10) {     // sm_100a only: fail loudly, there is no fallback
    delete ctx;
    return HSBP_ERR_UNSUPP;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream[0], cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream[1], cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->copy_ev[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->copy_ev[1], cudaEventDisableTiming) != cudaSuccess) {
    delete ctx;
    return HSBP_ERR_CUDA;
  }
  *out = ctx;
  return HSBP_OK;
}

int hsbp_ctx_destroy(hsbp_ctx *ctx) {
  if (!ctx) return HSBP_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  hsbp_comm_destroy(ctx);
  fdm_libs_destroy(ctx);
  cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->copy_ev[0]); cudaEventDestroy(ctx->copy_ev[1]);
  cudaStreamDestroy(ctx->copy_stream[0]); cudaStreamDestroy(ctx->copy_stream[1]);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return HSBP_OK;
}

const char *hsbp_last_error(hsbp_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

int hsbp_malloc(hsbp_ctx *ctx, size_t bytes, void **dptr) {
  if (!ctx || !dptr) return HSBP_ERR_ARG;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 8));
  return HSBP_OK;
}
int hsbp_free(hsbp_ctx *ctx, void *dptr) {
  if (!ctx) return HSBP_ERR_ARG;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  HSBP_CUDA(ctx, cudaFree(dptr));
  return HSBP_OK;
}
int hsbp_h2d(hsbp_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (!ctx || (bytes && (!dst || !src))) return HSBP_ERR_ARG;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HSBP_OK;
}
int hsbp_d2h(hsbp_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (!ctx || (bytes && (!dst || !src))) return HSBP_ERR_ARG;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HSBP_OK;
}
int hsbp_memset0(hsbp_ctx *ctx, void *dst, size_t bytes) {
  if (!ctx || (bytes && !dst)) return HSBP_ERR_ARG;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaMemsetAsync(dst, 0, bytes, ctx->stream));
  return HSBP_OK;
}
int hsbp_sync(hsbp_ctx *ctx) {
  if (!ctx) return HSBP_ERR_ARG;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HSBP_OK;
}
int hsbp_host_register(hsbp_ctx *ctx, void *host, size_t bytes) {
  if (!ctx || !host) return HSBP_ERR_ARG;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaHostRegister(host, bytes, cudaHostRegisterDefault));
  return HSBP_OK;
}
int hsbp_host_unregister(hsbp_ctx *ctx, void *host) {
  if (!ctx || !host) return HSBP_ERR_ARG;
  HSBP_CUDA(ctx, cudaHostUnregister(host));
  return HSBP_OK;
}
void *hsbp_stream(hsbp_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int hsbp_timer_start(hsbp_ctx *ctx) {
  if (!ctx) return HSBP_ERR_ARG;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return HSBP_OK;
}
int hsbp_timer_stop(hsbp_ctx *ctx, double *ms) {
  if (!ctx || !ms) return HSBP_ERR_ARG;
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  HSBP_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  float f = 0.f;
  HSBP_CUDA(ctx, cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
  *ms = (double)f;
  return HSBP_OK;
}

// ---------------------------------------------------------------------------------------------
int hsbp_blocks_create(hsbp_ctx *ctx, int p, int64_t nblocks, const int64_t *Nr, const int64_t *Ns,
                       hsbp_blocks **out) {
  if (!ctx || !out) return HSBP_ERR_ARG;
  *out = nullptr;
  if (p != 2 && p != 4 && p != 6) HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "unknown order (p must be 2, 4 or 6)");
  if (nblocks <= 0 || !Nr || !Ns) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_blocks_create: bad arguments");
  // smallest grid the closures fit on (diagonal_sbp.jl:129-131, 741-743)
  const int minN = (p == 2) ? 2 : (p == 4 ? 11 : 17);
  hsbp_blocks *b = new (std::nothrow) hsbp_blocks();
  if (!b) HSBP_FAIL(ctx, HSBP_ERR_STATE, "out of host memory");
  b->ctx = ctx; b->p = p; b->nblocks = nblocks;
  b->h_desc.resize(nblocks);
  int64_t voff = 0, foff = 0;
  b->uniform = true;
  for (int64_t e = 0; e < nblocks; ++e) {
    if (Nr[e] < minN || Ns[e] < minN || Nr[e] > (1 << 24) || Ns[e] > (1 << 24)) {
      delete b;
      HSBP_FAIL(ctx, HSBP_ERR_ARG, "Grid not big enough to support the operator");
    }
    BlockDesc &d = b->h_desc[e];
    memset(&d, 0, sizeof(d));
    d.Nr = (int32_t)Nr[e]; d.Ns = (int32_t)Ns[e]; d.voff = voff; d.foff = foff;
    for (int k = 0; k < 4; ++k) d.bc[k] = HSBP_BC_DIRICHLET;   // locoperator's default LFToB (:212-213)
    voff += (Nr[e] + 1) * (Ns[e] + 1);
    foff += 2 * (Nr[e] + 1) + 2 * (Ns[e] + 1);
    if (Nr[e] != Nr[0] || Ns[e] != Ns[0]) b->uniform = false;
    b->max_Nr = std::max<int>(b->max_Nr, d.Nr);
    b->max_Ns = std::max<int>(b->max_Ns, d.Ns);
  }
  b->VNp = voff; b->FNp = foff;
  cudaSetDevice(ctx->device);
  const size_t vb = (size_t)b->VNp * sizeof(double), fb = (size_t)b->FNp * sizeof(double);
  cudaError_t e = cudaSuccess;
  auto A = [&](void **p_, size_t n) { if (e == cudaSuccess) e = cudaMalloc(p_, n); };
  A((void **)&b->d_desc, nblocks * sizeof(BlockDesc));
  A((void **)&b->d_crr, vb); A((void **)&b->d_css, vb); A((void **)&b->d_crs, vb);
  A((void **)&b->d_tau, fb); A((void **)&b->d_fa, fb + 64); A((void **)&b->d_fb, fb + 64);   // (k_sweep's phantom point reads one entry past a face)
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(b->d_desc, b->h_desc.data(), nblocks * sizeof(BlockDesc), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    ctx->err = std::string("hsbp_blocks_create: ") + cudaGetErrorString(e);
    hsbp_blocks_destroy(b);
    return HSBP_ERR_CUDA;
  }
  *out = b;
  return HSBP_OK;
}

int hsbp_blocks_destroy(hsbp_blocks *b) {
  if (!b) return HSBP_ERR_ARG;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  cudaFree(b->d_desc); cudaFree(b->d_crr); cudaFree(b->d_css); cudaFree(b->d_crs);
  cudaFree(b->d_crr_s); cudaFree(b->d_css_s); cudaFree(b->d_rtab); cudaFree(b->d_rim);
  cudaFree(b->d_crs_p); cudaFree(b->d_sweep_dot);
  cudaFree(b->d_tau); cudaFree(b->d_fa); cudaFree(b->d_fb); cudaFree(b->d_t); cudaFree(b->d_w);
  cudaFree(b->d_stage_u); cudaFree(b->d_stage_y);
  for (cudaEvent_t ev : b->pipe_ev) if (ev) cudaEventDestroy(ev);
  cudaFree(b->d_dinv); cudaFree(b->d_pr); cudaFree(b->d_pp); cudaFree(b->d_pAp); cudaFree(b->d_pcg);
  cudaFree(b->d_nactive); cudaFree(b->d_chol); cudaFree(b->d_chol_off); cudaFree(b->d_chol_work);
  cudaFree(b->d_band); cudaFree(b->d_band_desc); cudaFree(b->d_band_work); cudaFree(b->d_band_inv);
  cudaFree(b->d_fdm_vr); cudaFree(b->d_fdm_vs); cudaFree(b->d_fdm_z); cudaFree(b->d_fdm_t);
  cudaFree(b->d_fdm_vr32); cudaFree(b->d_fdm_vs32); cudaFree(b->d_fdm_dinv32); cudaFree(b->d_fdm_a32); cudaFree(b->d_fdm_b32);
  cudaFree(b->d_fdm_vrT32); cudaFree(b->d_fdm_vsT32); cudaFree(b->d_fdm_dinvT32);
  delete b;
  return HSBP_OK;
}

int64_t hsbp_blocks_num_volume_points(const hsbp_blocks *b) { return b ? b->VNp : -1; }
int64_t hsbp_blocks_num_face_points(const hsbp_blocks *b) { return b ? b->FNp : -1; }

static int set_metrics(hsbp_blocks *b, const double *crr, const double *css, const double *crs, cudaMemcpyKind kind) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!crr || !css || !crs) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_blocks_set_metrics: null pointer");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t vb = (size_t)b->VNp * sizeof(double);
  HSBP_CUDA(ctx, cudaMemcpyAsync(b->d_crr, crr, vb, kind, ctx->stream));
  HSBP_CUDA(ctx, cudaMemcpyAsync(b->d_css, css, vb, kind, ctx->stream));
  HSBP_CUDA(ctx, cudaMemcpyAsync(b->d_crs, crs, vb, kind, ctx->stream));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  b->have_metrics = true;
  b->sweep_scaled_valid = false;
  b->rim_valid = false;
  operator_changed(b);
  return HSBP_OK;
}
int hsbp_blocks_set_metrics(hsbp_blocks *b, const double *crr, const double *css, const double *crs) {
  return set_metrics(b, crr, css, crs, cudaMemcpyHostToDevice);
}
int hsbp_blocks_set_metrics_dev(hsbp_blocks *b, const double *crr, const double *css, const double *crs) {
  return set_metrics(b, crr, css, crs, cudaMemcpyDeviceToDevice);
}

int hsbp_blocks_set_bc(hsbp_blocks *b, const int64_t *bctype) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!bctype) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_blocks_set_bc: null pointer");
  for (int64_t i = 0; i < 4 * b->nblocks; ++i) {
    const int64_t c = bctype[i];
    if (!(c == HSBP_BC_DIRICHLET || c == HSBP_BC_NEUMANN || c == HSBP_BC_LOCKED || c >= HSBP_BC_JUMP))
      HSBP_FAIL(ctx, HSBP_ERR_ARG, "invalid bc");                      // global_curved.jl:480-484
  }
  for (int64_t e = 0; e < b->nblocks; ++e)
    for (int k = 0; k < 4; ++k) b->h_desc[e].bc[k] = (int32_t)std::min<int64_t>(bctype[4 * e + k], 1 << 30);
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  HSBP_CUDA(ctx, cudaMemcpyAsync(b->d_desc, b->h_desc.data(), b->nblocks * sizeof(BlockDesc), cudaMemcpyHostToDevice, ctx->stream));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  b->have_bc = true;
  b->rim_valid = false;
  operator_changed(b);
  return HSBP_OK;
}

int hsbp_blocks_compute_tau(hsbp_blocks *b, double tauscale) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!b->have_metrics) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_blocks_compute_tau: set metrics first");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  int *d_bad = nullptr;
  HSBP_CUDA(ctx, cudaMalloc(&d_bad, sizeof(int)));
  HSBP_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  int rc = dispatch_p(b->p, [&](auto P) {
    k_compute_tau<decltype(P)::value><<<(unsigned)(4 * b->nblocks), 128, 0, ctx->stream>>>(
        b->d_desc, b->d_crr, b->d_css, b->d_crs, tauscale, b->d_tau, d_bad);
    int rc2 = check_launch(ctx, "k_compute_tau");
    if (rc2) return rc2;
    k_psi_min_check<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(b->VNp, b->d_crr, b->d_css, b->d_crs, d_bad);   // :419, whole block
    return check_launch(ctx, "k_psi_min_check");
  });
  int bad = 0;
  if (rc == HSBP_OK) {
    cudaError_t e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = HSBP_ERR_CUDA; }
  }
  cudaFree(d_bad);
  if (rc != HSBP_OK) return rc;
  if (bad) HSBP_FAIL(ctx, HSBP_ERR_ARG, "coefficient tensor is not positive definite (psi_min <= 0)");
  b->have_tau = true;
  b->rim_valid = false;
  operator_changed(b);
  return HSBP_OK;
}

int hsbp_blocks_set_tau(hsbp_blocks *b, const double *tau) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!tau) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_blocks_set_tau: null pointer");
  int rc = hsbp_h2d(ctx, b->d_tau, tau, (size_t)b->FNp * sizeof(double));
  if (rc == HSBP_OK) { b->have_tau = true; b->rim_valid = false; operator_changed(b); }
  return rc;
}
int hsbp_blocks_get_tau(hsbp_blocks *b, double *tau) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!tau) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_blocks_get_tau: null pointer");
  if (!b->have_tau) HSBP_FAIL(ctx, HSBP_ERR_STATE, "tau not set");
  return hsbp_d2h(ctx, tau, b->d_tau, (size_t)b->FNp * sizeof(double));
}

int hsbp_blocks_force_generic(hsbp_blocks *b, int on) {
  if (!b) return HSBP_ERR_ARG;
  b->force_generic = on;
  return HSBP_OK;
}
int hsbp_apply_variant(const hsbp_blocks *b) { return b ? b->last_variant : -1; }
int hsbp_blocks_set_option(hsbp_blocks *b, const char *name, int64_t value) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!name) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_blocks_set_option: null name");
  const std::string n(name);
  if (n == "force_generic") b->force_generic = (int)value;
  else if (n == "sweep_chunks_per_side") b->sweep_ncs_override = (int)value;
  else if (n == "sweep_points_per_thread") b->sweep_r_override = (int)value;
  else if (n == "sweep_fold_faces") b->sweep_fold_faces = (int)value;
  else if (n == "sweep_deep") b->sweep_deep = (int)value;
  else if (n == "fdm_gemm") b->fdm_gemm = (int)value;
  else if (n == "fdm_tc_variant") b->fdm_tc_variant = (int)value;
  else if (n == "fdm_no_skip") b->fdm_no_skip = (int)value;
  else if (n == "sweep_no_pdl") b->sweep_no_pdl = (int)value;
  else if (n == "host_groups") b->host_groups = (int)value;
  else if (n == "fdm_no_fused_dot") b->fdm_no_fused_dot = (int)value;
  else if (n == "fdm_eig_lib") b->fdm_eig_lib = (int)value;
  else if (n == "sweep_p6_regs") b->sweep_p6_regs = (int)value;
  else if (n == "band_no_stream") b->band_no_stream = (int)value;
  else HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_blocks_set_option: unknown option " + n);
  return HSBP_OK;
}

}  // extern "C"

static int ensure_scratch(hsbp_blocks *b) {
  hsbp_ctx *ctx = b->ctx;
  if (!b->d_t) {
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_t, (size_t)b->VNp * sizeof(double)));
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_w, (size_t)b->VNp * sizeof(double)));
  }
  return HSBP_OK;
}

static dim3 gen_grid(const hsbp_blocks *b) {
  const int64_t maxnp = (int64_t)(b->max_Nr + 1) * (b->max_Ns + 1);
  const int64_t tiles = std::min<int64_t>((maxnp + GEN_THREADS - 1) / GEN_THREADS, 65535);
  return dim3((unsigned)b->nblocks, (unsigned)tiles);
}

// volume part with the generic two-pass kernels
template <int P> static int vol_generic(hsbp_blocks *b, const double *u, double *y) {
  hsbp_ctx *ctx = b->ctx;
  int rc = ensure_scratch(b);
  if (rc) return rc;
  const dim3 grid = gen_grid(b);
  k_cross_pre<P><<<grid, GEN_THREADS, 0, ctx->stream>>>(b->d_desc, b->d_crs, u, b->d_t, b->d_w);
  if ((rc = check_launch(ctx, "k_cross_pre"))) return rc;
  k_vol_apply<P><<<grid, GEN_THREADS, 0, ctx->stream>>>(b->d_desc, b->d_crr, b->d_css, u, b->d_t, b->d_w, y);
  return check_launch(ctx, "k_vol_apply");
}

static int apply_async(hsbp_blocks *b, const double *u, double *y, cudaEvent_t *evs = nullptr) {
  hsbp_ctx *ctx = b->ctx;
  if (!b->have_metrics || !b->have_tau) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_apply: metrics / tau not set");
  if (!u || !y || u == y) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_apply: bad pointers");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  return dispatch_p(b->p, [&](auto Pc) {
    constexpr int P = decltype(Pc)::value;
    int rc;
    if (evs) cudaEventRecord(evs[0], ctx->stream);
    if (!b->force_generic && sweep_eligible<P>(b)) {
      b->last_variant = 1;
      if (b->sweep_fold_faces) {
        // k_edge_prep (rim of every block: face terms + r-end closure rows; only reads u), then one pass
        // that produces y = M-tilde u
        rc = vol_sweep<P>(b, u, y, true, evs ? evs[1] : nullptr);
        if (evs) { cudaEventRecord(evs[2], ctx->stream); cudaEventRecord(evs[3], ctx->stream); }
        return rc;
      }
      rc = vol_sweep<P>(b, u, y, false);                   // volume part only; face terms by the two generic kernels
    } else {
      rc = vol_generic<P>(b, u, y);
      b->last_variant = 0;
    }
    if (rc) return rc;
    if (evs) cudaEventRecord(evs[1], ctx->stream);
    k_face_gather<P><<<(unsigned)(4 * b->nblocks), 128, 0, ctx->stream>>>(
        b->d_desc, b->d_crr, b->d_css, b->d_crs, b->d_tau, u, b->d_fa, b->d_fb, FACE_APPLY);
    if ((rc = check_launch(ctx, "k_face_gather"))) return rc;
    if (evs) cudaEventRecord(evs[2], ctx->stream);
    k_face_scatter<P><<<(unsigned)b->nblocks, 256, 0, ctx->stream>>>(
        b->d_desc, b->d_crr, b->d_css, b->d_crs, b->d_fa, b->d_fb, y);
    if (evs) cudaEventRecord(evs[3], ctx->stream);
    return check_launch(ctx, "k_face_scatter");
  });
}

// the sweep kernel can leave u . M-tilde u per chunk when it produces the final y in one pass (fused faces, deep two-point kernel)
static bool sweep_dot_eligible(hsbp_blocks *b) {
  if (b->force_generic || !b->sweep_fold_faces || !b->sweep_deep) return false;
  if (dispatch_p(b->p, [&](auto Pc) { return sweep_eligible<decltype(Pc)::value>(b) ? 1 : 0; }) != 1) return false;
  return sweep_points_per_thread(b) == 2;
}

// out[e] = chunk sums of k_sweep (SweepParams::dot) + u . y on the first closure_pts points of either s-end of the block
__global__ void __launch_bounds__(256)
k_block_dot_finish(const BlockDesc *__restrict__ desc, const double *__restrict__ u, const double *__restrict__ y,
                   const double *__restrict__ part, int nch, int closure_pts, double *__restrict__ out) {
  __shared__ double scratch[32];
  const BlockDesc d = desc[blockIdx.x];
  const int64_t np = (int64_t)(d.Nr + 1) * (d.Ns + 1), o = d.voff;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < 2 * (int64_t)closure_pts; i += blockDim.x) {
    const int64_t k = i < closure_pts ? i : np - 2 * (int64_t)closure_pts + i;
    s += u[o + k] * y[o + k];
  }
  if (threadIdx.x == 0)
    for (int c = 0; c < nch; ++c) s += part[(int64_t)blockIdx.x * nch + c];
  s = cta_sum(s, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

extern "C" {

int hsbp_apply_energy(hsbp_blocks *b, const double *u_dev, double *y_dev, double *energy) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!energy) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_apply_energy: null pointer");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!sweep_dot_eligible(b)) HSBP_FAIL(ctx, HSBP_ERR_UNSUPP, "hsbp_apply_energy: needs the line-marching kernel (uniform blocks of >= 32 points per direction)");
  if (!b->d_sweep_dot) HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_sweep_dot, (size_t)b->nblocks * 65 * sizeof(double)));
  b->sweep_dot_out = b->d_sweep_dot;
  int rc = apply_async(b, u_dev, y_dev);
  b->sweep_dot_out = nullptr;
  if (rc) return rc;
  const int closure_pts = dispatch_p(b->p, [&](auto Pc) { return SweepTab<decltype(Pc)::value>::BM; }) * (b->max_Nr + 1);
  double *d_out = b->d_sweep_dot + (size_t)b->nblocks * 64;
  k_block_dot_finish<<<(unsigned)b->nblocks, 256, 0, ctx->stream>>>(b->d_desc, u_dev, y_dev, b->d_sweep_dot, b->sweep_nch, closure_pts, d_out);
  if ((rc = check_launch(ctx, "k_block_dot_finish"))) return rc;
  return hsbp_d2h(ctx, energy, d_out, (size_t)b->nblocks * sizeof(double));
}

int hsbp_apply(hsbp_blocks *b, const double *u_dev, double *y_dev) {
  if (!b) return HSBP_ERR_ARG;
  return apply_async(b, u_dev, y_dev);
}

int hsbp_apply_timed(hsbp_blocks *b, const double *u_dev, double *y_dev, double *ms) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!ms) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_apply_timed: null pointer");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaEvent_t evs[4];
  for (int i = 0; i < 4; ++i) HSBP_CUDA(ctx, cudaEventCreate(&evs[i]));
  int rc = apply_async(b, u_dev, y_dev, evs);
  if (rc == HSBP_OK) {
    cudaError_t e = cudaEventSynchronize(evs[3]);
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) {
      float f = 0.f;
      e = cudaEventElapsedTime(&f, evs[i], evs[i + 1]);
      ms[i] = f;
    }
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = HSBP_ERR_CUDA; }
    if (b->last_variant == 1 && b->sweep_fold_faces) std::swap(ms[0], ms[1]);     // launch order there: face preparation, then k_sweep
  }
  for (int i = 0; i < 4; ++i) cudaEventDestroy(evs[i]);
  return rc;
}

int hsbp_apply_host(hsbp_blocks *b, const double *u, double *y) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!u || !y) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_apply_host: null pointer");
  if (!b->have_metrics || !b->have_tau) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_apply_host: metrics / tau not set");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t vb = (size_t)b->VNp * sizeof(double);
  if (!b->d_stage_u) {
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_stage_u, vb));
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_stage_y, vb));
  }
  // Blocks are independent in M-tilde u, so the call is pipelined over groups of blocks: the H2D copy of
  // group g+1, the kernels of group g and the D2H copy of group g-1 overlap (three streams, PCIe is full duplex).
  const bool pipelined = !b->force_generic && b->sweep_fold_faces && b->nblocks >= 16 &&
                         dispatch_p(b->p, [&](auto Pc) { return sweep_eligible<decltype(Pc)::value>(b) ? 1 : 0; }) == 1;
  if (!pipelined) {
    HSBP_CUDA(ctx, cudaMemcpyAsync(b->d_stage_u, u, vb, cudaMemcpyHostToDevice, ctx->stream));
    int rc = apply_async(b, b->d_stage_u, b->d_stage_y);
    if (rc) return rc;
    HSBP_CUDA(ctx, cudaMemcpyAsync(y, b->d_stage_y, vb, cudaMemcpyDeviceToHost, ctx->stream));
    HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HSBP_OK;
  }
  const int ngroups = (int)std::min<int64_t>(b->host_groups > 0 ? b->host_groups : 16, b->nblocks / 8);
  const int64_t per = (b->nblocks + ngroups - 1) / ngroups;
  const int64_t np = (int64_t)(b->max_Nr + 1) * (b->max_Ns + 1);
  if ((int)b->pipe_ev.size() < 2 * ngroups) {
    const size_t old = b->pipe_ev.size();
    b->pipe_ev.resize(2 * ngroups, nullptr);
    for (size_t i = old; i < b->pipe_ev.size(); ++i)
      HSBP_CUDA(ctx, cudaEventCreateWithFlags(&b->pipe_ev[i], cudaEventDisableTiming));
  }
  b->last_variant = 1;
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // earlier work on the compute stream (setup) is done
  int rc = HSBP_OK;
  for (int g = 0; g < ngroups && rc == HSBP_OK; ++g) {
    const int64_t e0 = g * per, ne = std::min<int64_t>(per, b->nblocks - e0);
    if (ne <= 0) break;
    const size_t off = (size_t)(e0 * np), nbytes = (size_t)(ne * np) * sizeof(double);
    HSBP_CUDA(ctx, cudaMemcpyAsync(b->d_stage_u + off, u + off, nbytes, cudaMemcpyHostToDevice, ctx->copy_stream[0]));
    HSBP_CUDA(ctx, cudaEventRecord(b->pipe_ev[2 * g], ctx->copy_stream[0]));
    HSBP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, b->pipe_ev[2 * g], 0));
    rc = dispatch_p(b->p, [&](auto Pc) {
      return vol_sweep<decltype(Pc)::value>(b, b->d_stage_u, b->d_stage_y, true, nullptr, e0, ne);
    });
    if (rc) break;
    HSBP_CUDA(ctx, cudaEventRecord(b->pipe_ev[2 * g + 1], ctx->stream));
    HSBP_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream[1], b->pipe_ev[2 * g + 1], 0));
    HSBP_CUDA(ctx, cudaMemcpyAsync(y + off, b->d_stage_y + off, nbytes, cudaMemcpyDeviceToHost, ctx->copy_stream[1]));
  }
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream[0]));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  HSBP_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream[1]));
  return rc;
}

int hsbp_face_FT(hsbp_blocks *b, const double *u_dev, double *ft_dev) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!b->have_metrics || !b->have_tau) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_face_FT: metrics / tau not set");
  if (!u_dev || !ft_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_face_FT: null pointer");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  return dispatch_p(b->p, [&](auto Pc) {
    k_face_gather<decltype(Pc)::value><<<(unsigned)(4 * b->nblocks), 128, 0, ctx->stream>>>(
        b->d_desc, b->d_crr, b->d_css, b->d_crs, b->d_tau, u_dev, ft_dev, nullptr, FACE_FT);
    return check_launch(ctx, "k_face_gather");
  });
}

int hsbp_face_traction(hsbp_blocks *b, const double *u_dev, double *tr_dev) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!b->have_metrics || !b->have_tau) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_face_traction: metrics / tau not set");
  if (!u_dev || !tr_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_face_traction: null pointer");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  return dispatch_p(b->p, [&](auto Pc) {
    k_face_gather<decltype(Pc)::value><<<(unsigned)(4 * b->nblocks), 128, 0, ctx->stream>>>(
        b->d_desc, b->d_crr, b->d_css, b->d_crs, b->d_tau, u_dev, tr_dev, nullptr, FACE_TRACTION);
    return check_launch(ctx, "k_face_gather");
  });
}

int hsbp_face_F_add(hsbp_blocks *b, const double *v_dev, double alpha, double *y_dev) {
  if (!b) return HSBP_ERR_ARG;
  hsbp_ctx *ctx = b->ctx;
  if (!b->have_metrics || !b->have_tau) HSBP_FAIL(ctx, HSBP_ERR_STATE, "hsbp_face_F_add: metrics / tau not set");
  if (!v_dev || !y_dev) HSBP_FAIL(ctx, HSBP_ERR_ARG, "hsbp_face_F_add: null pointer");
  HSBP_CUDA(ctx, cudaSetDevice(ctx->device));
  return dispatch_p(b->p, [&](auto Pc) {
    constexpr int P = decltype(Pc)::value;
    k_face_prep_F<P><<<(unsigned)(4 * b->nblocks), 128, 0, ctx->stream>>>(b->d_desc, b->d_tau, v_dev, alpha, b->d_fa, b->d_fb);
    int rc = check_launch(ctx, "k_face_prep_F");
    if (rc) return rc;
    k_face_scatter<P><<<(unsigned)b->nblocks, 256, 0, ctx->stream>>>(b->d_desc, b->d_crr, b->d_css, b->d_crs, b->d_fa, b->d_fb, y_dev);
    return check_launch(ctx, "k_face_scatter");
  });
}

}  // extern "C"

#include "api_geom.cuh"
#include "api_comm.cuh"
#include "api_chol.cuh"
#include "k_dense.cuh"
#include "api_band.cuh"
#include "api_factor.cuh"
#include "api_solve.cuh"
#include "api_cg.cuh"
#include "api_fdm.cuh"
#include "api_bp1.cuh"
#include "api_fault.cuh"
#include "api_peaks.cuh"
