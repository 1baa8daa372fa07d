// Internal data structures of libhsbp (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/hsbp.h"

struct hsbp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream[2] = {nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t copy_ev[2] = {nullptr, nullptr};
  int sm_count = 148;
  size_t smem_optin = 0;
  std::vector<std::pair<const void *, size_t>> smem_optin_done;   // kernels whose dynamic shared-memory opt-in was made on THIS device
  void *comm = nullptr;             // ncclComm_t of this context (api_comm.cuh); one rank per context
  int rank = 0, world = 1;
  void *fdm_libs = nullptr;         // cuBLAS / cuSOLVER handles of the fast-diagonalisation preconditioner (api_fdm.cuh)
  std::string err;
};

// Per-block descriptor on the device (one per block, 64 bytes).
struct BlockDesc {
  int32_t Nr, Ns;        // grid sizes (points are Nr+1, Ns+1)
  int64_t voff;          // 0-based offset of the block in volume vectors
  int64_t foff;          // 0-based offset of the block's face 1 in face vectors
  int32_t bc[4];         // boundary-condition code of local faces 1..4
  int32_t pad[2];
  // face k of this block starts at foff + fstart(k):
  //   k=0: 0, k=1: Ns+1, k=2: 2(Ns+1), k=3: 2(Ns+1)+(Nr+1)
};

struct hsbp_blocks {
  hsbp_ctx *ctx = nullptr;
  int p = 0;
  int64_t nblocks = 0;
  int64_t VNp = 0;       // volume points
  int64_t FNp = 0;       // face points (all blocks, all four faces)
  std::vector<BlockDesc> h_desc;
  BlockDesc *d_desc = nullptr;
  double *d_crr = nullptr, *d_css = nullptr, *d_crs = nullptr;
  double *d_rim = nullptr;                         // static r-face data for k_edge_prep (k_rim_build)
  bool rim_valid = false;
  double *d_rtab = nullptr;                        // r-end table of the line-marching kernel (k_edge_prep)
  double *d_crr_s = nullptr, *d_css_s = nullptr;   // norm-weighted copies for the line-marching kernel (lazy)
  double *d_crs_p = nullptr;                       // odd line lengths: pitched copy (even pitch, zero pad) of crs
  bool sweep_scaled_valid = false;
  double *d_tau = nullptr;          // FNp
  double *d_fa = nullptr;           // FNp scratch: alpha (or F^T u)
  double *d_fb = nullptr;           // FNp scratch: beta
  double *d_t = nullptr, *d_w = nullptr;   // VNp scratch of the generic kernels (lazy)
  bool have_metrics = false, have_bc = false, have_tau = false;
  bool uniform = false;             // all blocks share (Nr, Ns)
  int max_Nr = 0, max_Ns = 0;
  int force_generic = 0;
  int sweep_r_override = 0;         // points per thread of the line-marching kernel (0 = heuristic, 2 or 4)
  int sweep_deep = 1;               // 1: css / crs of older lines come from deeper shared-memory rings, 0: register windows (k_sweep.cuh)
  int last_sweep_ctas_per_sm = 0;
  int sweep_p6_regs = 168;          // register cap of the p = 6, 2-points-per-thread deep-ring kernel (168: 0.98 ms, 128: 1.13 ms)
  int sweep_fold_faces = 1;         // fold the face terms into k_sweep (0: separate gather / scatter kernels)
  int sweep_ncs_override = 0;       // chunks per side of the line-marching kernel (0 = heuristic)
  int last_variant = -1;
  uint64_t generation = 0;          // bumped whenever the operator changes (metrics, bc, tau): factors, condensed blocks and
                                    // preconditioners built for an older generation are stale
  // local solves
  int local_mode = 0;
  double local_tol = 1e-13;
  int64_t local_maxit = 100000;
  double *d_dinv = nullptr, *d_pr = nullptr, *d_pp = nullptr, *d_pAp = nullptr;
  void *d_pcg = nullptr;
  int *d_nactive = nullptr;
  double *d_chol = nullptr;                 // dense factors, block e at chol_off[e], leading dimension Np_e
  std::vector<int64_t> chol_off;
  void *d_chol_off = nullptr;               // CholBlock descriptors
  double *d_chol_work = nullptr;
  double *d_band = nullptr;                 // banded factors (api_band.cuh), LAPACK lower-band storage per block
  void *d_band_desc = nullptr;              // BandBlock descriptors
  double *d_band_work = nullptr;
  double *d_band_inv = nullptr;             // inverted diagonal blocks of the banded factors (streamed solve)
  int band_stream_stages = 0, band_maxld = 0, band_maxnpad = 0, band_no_stream = 0;
  int band_pb = 16;                // panel width of the streamed banded solve (16, or 8 / 4 for wide bands)
  double *d_fdm_vr = nullptr, *d_fdm_vs = nullptr;   // generalised eigenvectors of the collapsed 1-D operators (api_fdm.cuh)
  double *d_fdm_z = nullptr, *d_fdm_t = nullptr;     // preconditioned residual, GEMM scratch
  float *d_fdm_vr32 = nullptr, *d_fdm_vs32 = nullptr, *d_fdm_dinv32 = nullptr, *d_fdm_a32 = nullptr, *d_fdm_b32 = nullptr;
  float *d_fdm_vrT32 = nullptr, *d_fdm_vsT32 = nullptr, *d_fdm_dinvT32 = nullptr;    // transposes: every tensor-core operand contiguous in k
  int host_groups = 0;              // groups of blocks hsbp_apply_host pipelines its copies and kernels over (0: 16)
  int sweep_no_pdl = 0;             // 1: k_sweep without programmatic dependent launch behind k_edge_prep (testing / timing)
  double *d_sweep_dot = nullptr;    // [nblocks][<= 64] chunk sums of p . Ap (FDM-PCG)
  int fdm_no_fused_dot = 0;         // 1: p . Ap by a separate pass of the update kernel (testing)
  double *sweep_dot_out = nullptr;  // set around a batched PCG: k_sweep leaves u . M-tilde u per (block, chunk) here (SweepParams::dot)
  int sweep_nch = 0;                // chunks per block of the last k_sweep launch
  const int *skip_flags = nullptr;  // set around a batched PCG: blocks with skip_flags[e * skip_stride] == 0 are left out of
  int skip_stride = 0;              // hsbp_apply's sweep kernels and the preconditioner (converged blocks)
  int fdm_eig_lib = 0;              // 1: eigen-decompositions of the FDM setup by cuSOLVER syevd (comparison only)
  int fdm_no_skip = 0;              // 1: converged blocks stay in the kernels of the FDM-PCG iteration (testing / timing)
  int fdm_tc_variant = 0;           // 0: k_fdm_pair (two fused GEMM pairs, TMA operands); 1: four single-GEMM launches (testing)
  alignas(64) unsigned char fdm_tm[6][128];   // CUtensorMap x 6 (k_fdm_pair operands), valid after hsbp_local_setup(HSBP_LOCAL_FDM)
  bool fdm_tm_valid = false;
  int fdm_gemm = 0;                 // arithmetic of the preconditioner's GEMMs: 0 fp64 (default), 1 fp32, 2 fp32 emulated on BF16 tensor
                                    // cores, 3 TF32 tensor cores (2.1x faster solves on smooth blocks, the synthetic-mesh drivers opt in;
                                    // on strongly varying coefficients the TF32 noise can stall the already weak preconditioner)
  // pinned staging for hsbp_apply_host (lazy)
  double *d_stage_u = nullptr, *d_stage_y = nullptr;
  std::vector<cudaEvent_t> pipe_ev;          // per block group: H2D done, kernels done
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember per context which kernels have it
template <class K> inline cudaError_t hsbp_smem_optin(hsbp_ctx *ctx, K kernel, size_t bytes) {
  const void *key = reinterpret_cast<const void *>(kernel);
  for (auto &k : ctx->smem_optin_done)
    if (k.first == key) {
      if (k.second >= bytes) return cudaSuccess;
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e == cudaSuccess) k.second = bytes;
      return e;
    }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) ctx->smem_optin_done.push_back({key, bytes});
  return e;
}

#define HSBP_CUDA(ctx, call)                                                     \
  do {                                                                           \
    cudaError_t _e = (call);                                                     \
    if (_e != cudaSuccess) {                                                     \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e);           \
      return HSBP_ERR_CUDA;                                                      \
    }                                                                            \
  } while (0)

#define HSBP_FAIL(ctx, code, msg) \
  do {                            \
    (ctx)->err = (msg);           \
    return (code);                \
  } while (0)
