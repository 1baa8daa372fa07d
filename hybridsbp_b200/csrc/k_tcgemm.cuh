// Batched 256-class GEMMs of the fast-diagonalisation preconditioner (K2d) on the 5th-generation tensor cores:
// tcgen05.mma kind::tf32 with the accumulator in tensor memory (TMEM), hand-written for sm_100a.
//
//   D_b[m][n] = sum_k A_b[m][k] * B_b[n][k]         b = block of the batch, both operands "K-major": a row of A (of B)
//                                                   is contiguous in k
// One CTA computes a 128 x N tile (N <= 256, multiple of 16): the whole accumulator is 128 lanes x N columns of TMEM.
// K runs in blocks of 32 through a two-stage shared-memory ring; operands are written by all threads in the canonical
// no-swizzle K-major layout of the tensor core (8 x 16-byte core matrices; 16-byte chunk c of row r at c * rows * 16 + r * 16),
// converting fp64 sources to fp32 on the way (the MMA reads the top 19 bits: TF32).  One elected thread issues the four
// K = 8 MMAs of a block and commits them to the stage's mbarrier; loads of the next block overlap with them.  The epilogue
// reads TMEM with tcgen05.ld (32 lanes x 32 columns per warp and call) and writes row-major fp32 (optionally scaled
// elementwise -- the "o Dinv" between the GEMM pairs), column-major fp32 or column-major fp64.
//
// The preconditioner z = Vr [ (Vr^T R Vs) o Dinv ] Vs^T was four launches of this kernel in the first version; the default is
// now k_fdm_pair below (two chained GEMMs per launch, TMA operands, TMEM operand), this kernel stays as the comparison variant
// ("fdm_tc_variant" = 1).  The reference has no counterpart -- it factorises M-tilde (global_curved.jl:698) -- this is the
// engine of the batched PCG local solver.
#pragma once
#include <cuda.h>            // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

namespace hsbp {
namespace tc {

constexpr int BM = 128;          // rows of the output tile = TMEM lanes
constexpr int BK = 32;           // K per stage (8 chunks of 16 bytes)
constexpr int NSTAGE = 2;
constexpr int THREADS = 256;

enum OutMode { OUT_ROWMAJOR_F32 = 0, OUT_ROWMAJOR_F32_SCALED = 1, OUT_COLMAJOR_F32 = 2, OUT_COLMAJOR_F64 = 3 };

struct GemmParams {
  const void *A;       // fp32, A[b * strideA + m * lda + k]
  const void *B;       // fp32 or fp64 (b_is_f64), B[b * strideB + n * ldb + k]
  void *out;           // see OutMode; [b * strideO + ...] with leading dimension ldo
  const float *scale;  // OUT_ROWMAJOR_F32_SCALED: out[m][n] = acc * scale[b * strideO + m * ldo + n]
  int64_t strideA, strideB, strideO;
  int M, N, K;         // per block; M multiple of 128, N multiple of 16 and <= 256, K multiple of 32
  int lda, ldb, ldo;
  int b_is_f64;
  int mode;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// round to nearest TF32 (the MMA itself truncates the low 13 mantissa bits of an fp32 operand): every value that will be
// read as a tensor-core operand is rounded once, where it is produced
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ void mbar_init_(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait_(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity)
      : "memory");
}

// shared-memory matrix descriptor, K-major, no swizzle: core matrices of 8 rows x 16 bytes are contiguous (128 B);
// SBO = distance between 8-row groups, LBO = distance between the 16-byte chunks along K (both in bytes here)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // descriptor version of sm_100
  return d;                                     // base offset 0, layout type 0 (no swizzle)
}

// instruction descriptor: D fp32, A and B TF32, both K-major, M x N
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                                 // c_format: F32
  d |= 2u << 7;                                 // a_format: TF32
  d |= 2u << 10;                                // b_format: TF32
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// shared memory per CTA: NSTAGE * (BM + N) * BK * 4 bytes + barriers
__host__ __device__ inline size_t gemm_smem_bytes(int N) { return (size_t)NSTAGE * (BM + N) * BK * 4 + 1024; }

// epilogue of one 128 x N tile: TMEM -> registers -> global, in the layout the consumer of the result wants
__device__ __forceinline__ void tc_epilogue(const GemmParams &p, uint32_t tmem_d, int mt, int64_t b, int warp, int lane, int N) {
  // epilogue: warp w reads lanes 32 (w % 4) .. +31 (= rows of the tile), column chunks (w / 4), (w / 4) + 2, ...
  const int row = mt * BM + (warp & 3) * 32 + lane;
  const int64_t ob = b * p.strideO;
  for (int c0 = (warp >> 2) * 32; c0 < N; c0 += 64) {
    uint32_t v[32];
    tmem_ld32(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int ncol = min(32, N - c0);
    if (p.mode == OUT_ROWMAJOR_F32 || p.mode == OUT_ROWMAJOR_F32_SCALED) {
      float *o = reinterpret_cast<float *>(p.out) + ob + (int64_t)row * p.ldo + c0;
      const float *sc = p.mode == OUT_ROWMAJOR_F32_SCALED ? p.scale + ob + (int64_t)row * p.ldo + c0 : nullptr;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j < ncol) {
          float4 w = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          if (sc) {
            const float4 q = *reinterpret_cast<const float4 *>(sc + j);
            w.x *= q.x; w.y *= q.y; w.z *= q.z; w.w *= q.w;
          }
          w.x = round_tf32(w.x); w.y = round_tf32(w.y); w.z = round_tf32(w.z); w.w = round_tf32(w.w);   // operand of the next GEMM
          *reinterpret_cast<float4 *>(o + j) = w;
        }
      }
    } else if (p.mode == OUT_COLMAJOR_F32) {
      float *o = reinterpret_cast<float *>(p.out) + ob + row;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncol) o[(int64_t)(c0 + j) * p.ldo] = round_tf32(__uint_as_float(v[j]));
    } else {
      double *o = reinterpret_cast<double *>(p.out) + ob + row;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncol) o[(int64_t)(c0 + j) * p.ldo] = (double)__uint_as_float(v[j]);
    }
  }
}

__global__ void __launch_bounds__(THREADS, 2)
k_tc_gemm(GemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t bar_free[NSTAGE];        // the MMAs that read a stage have completed
  __shared__ uint64_t bar_acc;                 // all MMAs of the tile have completed
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, K = p.K;
  const int mt = blockIdx.x;                    // 128-row tile
  const int64_t b = blockIdx.y;
  const uint32_t a_stage_bytes = BM * BK * 4, b_stage_bytes = (uint32_t)N * BK * 4;
  unsigned char *sA = smem_raw, *sB = smem_raw + NSTAGE * a_stage_bytes;

  // TMEM: 128 lanes x ncols columns of fp32 (power of two >= 32)
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init_(&bar_free[s], 1);
    mbar_init_(&bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;

  const float *Ag = reinterpret_cast<const float *>(p.A) + b * p.strideA + (int64_t)mt * BM * p.lda;
  const float *Bg32 = reinterpret_cast<const float *>(p.B) + b * p.strideB;
  const double *Bg64 = reinterpret_cast<const double *>(p.B) + b * p.strideB;
  const uint32_t idesc = make_idesc(BM, N);
  const int nkb = K / BK;
  // a warp copies 8 rows x 4 chunks per trip: shared-memory stores of the 8 rows of one chunk are 128 contiguous bytes
  const int lr = lane & 7, lc = lane >> 3;       // row within the group of 8, chunk within the group of 4

  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb % NSTAGE;
    if (kb >= NSTAGE) mbar_wait_(&bar_free[s], (uint32_t)((kb / NSTAGE - 1) & 1));
    unsigned char *sa = sA + (size_t)s * a_stage_bytes, *sb = sB + (size_t)s * b_stage_bytes;
    const int k0 = kb * BK;
    // Unit of work = (row group of 8, chunk group of 4); a warp takes one unit per trip.  All global loads of a phase are
    // issued before the first shared-memory store, so a phase exposes ONE memory latency (not one per trip).
    constexpr int NW = THREADS / 32, TA = (BM / 8) * 2 / NW;       // 4 trips for A
    {                                                              // phase 1: A and the first 128 rows of B
      float4 ra[TA];
#pragma unroll
      for (int i = 0; i < TA; ++i) {
        const int u = warp + i * NW, r = (u >> 1) * 8 + lr, c = (u & 1) * 4 + lc;
        ra[i] = *reinterpret_cast<const float4 *>(Ag + (int64_t)r * p.lda + k0 + c * 4);
      }
      if (p.b_is_f64) {
        double2 rb[TA][2];
#pragma unroll
        for (int i = 0; i < TA; ++i) {
          const int u = warp + i * NW, r = (u >> 1) * 8 + lr, c = (u & 1) * 4 + lc;
          const double2 *src = reinterpret_cast<const double2 *>(Bg64 + (int64_t)r * p.ldb + k0 + c * 4);
          rb[i][0] = src[0]; rb[i][1] = src[1];
        }
#pragma unroll
        for (int i = 0; i < TA; ++i) {
          const int u = warp + i * NW, r = (u >> 1) * 8 + lr, c = (u & 1) * 4 + lc;
          *reinterpret_cast<float4 *>(sb + (size_t)c * N * 16 + (size_t)r * 16) = make_float4(
              round_tf32((float)rb[i][0].x), round_tf32((float)rb[i][0].y), round_tf32((float)rb[i][1].x), round_tf32((float)rb[i][1].y));
        }
      } else {
        float4 rb[TA];
#pragma unroll
        for (int i = 0; i < TA; ++i) {
          const int u = warp + i * NW, r = (u >> 1) * 8 + lr, c = (u & 1) * 4 + lc;
          rb[i] = *reinterpret_cast<const float4 *>(Bg32 + (int64_t)r * p.ldb + k0 + c * 4);
        }
#pragma unroll
        for (int i = 0; i < TA; ++i) {
          const int u = warp + i * NW, r = (u >> 1) * 8 + lr, c = (u & 1) * 4 + lc;
          *reinterpret_cast<float4 *>(sb + (size_t)c * N * 16 + (size_t)r * 16) = rb[i];
        }
      }
#pragma unroll
      for (int i = 0; i < TA; ++i) {
        const int u = warp + i * NW, r = (u >> 1) * 8 + lr, c = (u & 1) * 4 + lc;
        *reinterpret_cast<float4 *>(sa + (size_t)c * BM * 16 + (size_t)r * 16) = ra[i];
      }
    }
    if (N > 128) {                                                 // phase 2: rows 128 .. N-1 of B
      const int nu = (N / 8) * 2;
      if (p.b_is_f64) {
        double2 rb[TA][2];
#pragma unroll
        for (int i = 0; i < TA; ++i) {
          const int u = (BM / 8) * 2 + warp + i * NW, r = (u >> 1) * 8 + lr, c = (u & 1) * 4 + lc;
          if (u < nu) {
            const double2 *src = reinterpret_cast<const double2 *>(Bg64 + (int64_t)r * p.ldb + k0 + c * 4);
            rb[i][0] = src[0]; rb[i][1] = src[1];
          }
        }
#pragma unroll
        for (int i = 0; i < TA; ++i) {
          const int u = (BM / 8) * 2 + warp + i * NW, r = (u >> 1) * 8 + lr, c = (u & 1) * 4 + lc;
          if (u < nu)
            *reinterpret_cast<float4 *>(sb + (size_t)c * N * 16 + (size_t)r * 16) = make_float4(
                round_tf32((float)rb[i][0].x), round_tf32((float)rb[i][0].y), round_tf32((float)rb[i][1].x), round_tf32((float)rb[i][1].y));
        }
      } else {
        float4 rb[TA];
#pragma unroll
        for (int i = 0; i < TA; ++i) {
          const int u = (BM / 8) * 2 + warp + i * NW, r = (u >> 1) * 8 + lr, c = (u & 1) * 4 + lc;
          if (u < nu) rb[i] = *reinterpret_cast<const float4 *>(Bg32 + (int64_t)r * p.ldb + k0 + c * 4);
        }
#pragma unroll
        for (int i = 0; i < TA; ++i) {
          const int u = (BM / 8) * 2 + warp + i * NW, r = (u >> 1) * 8 + lr, c = (u & 1) * 4 + lc;
          if (u < nu) *reinterpret_cast<float4 *>(sb + (size_t)c * N * 16 + (size_t)r * 16) = rb[i];
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_addr(sa), b0 = smem_addr(sb);
#pragma unroll
      for (int j = 0; j < BK / 8; ++j) {           // K = 8 per instruction: chunks 2j, 2j+1
        const uint64_t ad = make_desc(a0 + (uint32_t)(2 * j) * BM * 16, BM * 16, 128);
        const uint64_t bd = make_desc(b0 + (uint32_t)(2 * j) * (uint32_t)N * 16, (uint32_t)N * 16, 128);
        mma_tf32(tmem_d, ad, bd, idesc, (kb > 0 || j > 0) ? 1u : 0u);
      }
      mma_commit(&bar_free[s]);
      if (kb == nkb - 1) mma_commit(&bar_acc);
    }
  }
  mbar_wait_(&bar_acc, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  tc_epilogue(p, tmem_d, mt, b, warp, lane, N);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(ncols) : "memory");
}


// ---- two chained GEMMs per launch: TMA operands, second GEMM reads its A operand from tensor memory ---------------------------
//   D1[128 x N]  = A1[m-tile, K = M] * B1[N, K = M]^T          both operands from shared memory (SS)
//   D2[128 x N]  = D1[128, K = N]    * B2[N, K = N]^T          A = the accumulator of the first GEMM, where it lies in TMEM (TS)
// The preconditioner  Z = Vr [ (Vr^T R Vs) o Dinv ] Vs^T  is two launches:
//   (a)  T3 = ((Vr^T R) Vs) o Dinv      A1 = Vr (rows m, k contiguous), B1 = R as fp32, B2 = Vs;       out fp32, column-major, TF32-rounded
//   (b)  Z  = (Vr T3) Vs^T              A1 = Vr^T, B1 = T3 (column-major: contiguous in the contraction index), B2 = Vs^T;  out fp64
// so T1 = Vr^T R and Vr T3 never leave the SM (each is 128 KB of TMEM per CTA).
// Every operand tile is one 3-D TMA box (32 fp32 = 128 bytes along k, all rows, one block of the batch) written with the 128-byte
// swizzle the tensor core reads; a 4-stage ring of 48 KB stages, filled by one elected lane of warp 0, drained by one elected
// lane of warp 1 which issues the MMAs (4 x K = 8 per stage) and commits each stage back to the producer.  TMEM: columns
// [0, N) hold D1, [N, 2N) D2.  All eight warps run the epilogue (tcgen05.ld, 32 lanes x 32 columns per call), lanes = rows of the
// tile, so that every store instruction of a warp writes 32 consecutive values of a column.
constexpr int PSTG = 4;
constexpr int PA_BYTES = BM * 128;                 // A tile: 128 rows x 128 bytes
constexpr int PB_BYTES = 256 * 128;                // B tile: up to 256 rows x 128 bytes
constexpr int PSTAGE_BYTES = PA_BYTES + PB_BYTES;
__host__ __device__ inline size_t pair_smem_bytes() { return (size_t)PSTG * PSTAGE_BYTES + 1024; }

struct PairParams {
  void *out;              // (a) fp32 / (b) fp64, [b * strideO + m + M * n]
  const float *scale;     // (a) Dinv as fp32, indexed like out; (b) unused
  const int *active;      // optional per-block flag (block b at active[b * active_stride]); blocks with 0 are skipped
  int active_stride;
  int M, N;               // block is M x N (M = Nr+1 = rows and K of the first GEMM, N = Ns+1 = columns and K of the second)
  int64_t strideO;
};

// K-major operand tile with 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset
  d |= (uint64_t)1 << 46;                         // descriptor version of sm_100
  d |= (uint64_t)2 << 61;                         // layout type: SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t *v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one_() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}

// OUT_F64: launch (b), fp64 output, no scaling.  Between the GEMMs warps 4..7 round the accumulator of the first one to the
// nearest TF32 in place (tcgen05.ld / tcgen05.st) -- as an operand the tensor core would truncate it -- and the issuing lane
// waits for them: the second GEMM must not read D1 before the first one has committed (measured: back-to-back issue without
// the commit / wait reads a partial accumulator).
template <bool OUT_F64>
__global__ void __launch_bounds__(THREADS, 1)
k_fdm_pair(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
           const __grid_constant__ CUtensorMap tmB2, const PairParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t bar_full[PSTG], bar_empty[PSTG], bar_acc1, bar_mid, bar_acc2;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x, b = blockIdx.y;
  if (p.active != nullptr && p.active[(int64_t)b * p.active_stride] == 0) return;       // converged block (uniform per CTA)
  const int N = p.N, nk1 = p.M / BK, nk2 = N / BK, nit = nk1 + nk2;
  const uint32_t s0 = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const uint32_t ncols = 2 * N <= 256 ? 256u : 512u;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < PSTG; ++s) { mbar_init_(&bar_full[s], 1); mbar_init_(&bar_empty[s], 1); }
    mbar_init_(&bar_acc1, 1); mbar_init_(&bar_mid, 128); mbar_init_(&bar_acc2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_x = tmem_base_s, tmem_y = tmem_base_s + (uint32_t)N;

  if (warp == 0) {
    if (elect_one_()) {                                            // ---- producer: TMA into the ring ----
      for (int it = 0; it < nit; ++it) {
        const int s = it % PSTG;
        if (it >= PSTG) mbar_wait_(&bar_empty[s], (uint32_t)((it / PSTG - 1) & 1));
        const uint32_t sa = s0 + (uint32_t)s * PSTAGE_BYTES, sb = sa + PA_BYTES;
        if (it < nk1) {
          mbar_expect_tx_(&bar_full[s], (uint32_t)(PA_BYTES + N * 128));
          tma_load_3d(sa, &tmA1, it * BK, mt * BM, b, &bar_full[s]);
          tma_load_3d(sb, &tmB1, it * BK, 0, b, &bar_full[s]);
        } else {
          mbar_expect_tx_(&bar_full[s], (uint32_t)(N * 128));
          tma_load_3d(sb, &tmB2, (it - nk1) * BK, 0, b, &bar_full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one_()) {                                            // ---- MMA issuer ----
      const uint32_t idesc = make_idesc(BM, N);
      for (int it = 0; it < nit; ++it) {
        const int s = it % PSTG;
        mbar_wait_(&bar_full[s], (uint32_t)((it / PSTG) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = s0 + (uint32_t)s * PSTAGE_BYTES, sb = sa + PA_BYTES;
        if (it < nk1) {
#pragma unroll
          for (int j = 0; j < BK / 8; ++j)
            mma_tf32(tmem_x, make_desc_sw128(sa + 32u * j), make_desc_sw128(sb + 32u * j), idesc, (it > 0 || j > 0) ? 1u : 0u);
        } else {
          const int kk = it - nk1;
          if (kk == 0) {                                    // D1 rounded in place by warps 4..7
            mbar_wait_(&bar_mid, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
#pragma unroll
          for (int j = 0; j < BK / 8; ++j)
            mma_tf32_ts(tmem_y, tmem_x + (uint32_t)(kk * BK + 8 * j), make_desc_sw128(sb + 32u * j), idesc, (kk > 0 || j > 0) ? 1u : 0u);
        }
        mma_commit(&bar_empty[s]);
        if (it == nk1 - 1) mma_commit(&bar_acc1);
        if (it == nit - 1) mma_commit(&bar_acc2);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {                                   // ---- D1 -> nearest TF32, in place ----
    mbar_wait_(&bar_acc1, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tl = tmem_x + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tl + (uint32_t)c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(round_tf32(__uint_as_float(v[j])));
      tmem_st32(tl + (uint32_t)c0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive_(&bar_mid);
  }

  // ---- epilogue: D2 -> global, column-major (lanes = rows: coalesced) ----
  mbar_wait_(&bar_acc2, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int row = mt * BM + (warp & 3) * 32 + lane;
    const uint32_t tl = tmem_y + ((uint32_t)((warp & 3) * 32) << 16);
    const int64_t ob = (int64_t)b * p.strideO + row;
    for (int c0 = (warp >> 2) * 32; c0 < N; c0 += 64) {
      uint32_t v[32];
      tmem_ld32(tl + (uint32_t)c0, v);
      if constexpr (OUT_F64) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        double *o = reinterpret_cast<double *>(p.out) + ob + (int64_t)c0 * p.M;
#pragma unroll
        for (int j = 0; j < 32; ++j) o[(int64_t)j * p.M] = (double)__uint_as_float(v[j]);
      } else {
        const float *sc = p.scale + ob + (int64_t)c0 * p.M;
        float q[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) q[j] = __ldg(sc + (int64_t)j * p.M);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float *o = reinterpret_cast<float *>(p.out) + ob + (int64_t)c0 * p.M;
#pragma unroll
        for (int j = 0; j < 32; ++j) o[(int64_t)j * p.M] = round_tf32(__uint_as_float(v[j]) * q[j]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(ncols) : "memory");
}

}  // namespace tc
}  // namespace hsbp

// ---- fp64 variant on the fp64 tensor pipe (mma.sync.m8n8k4.f64), any sizes and strides ----------------------------------------
//   C_b(m, n) = [scale_b(m, n) *] sum_k A_b(m, k) B_b(n, k),   X(i, j) at X[i * s0 + j * s1]
// 64 x 64 tile per CTA, 8 warps: warp w owns rows 8 w .. 8 w + 7 and the 8 column tiles.  Used by the fp64 mode of the
// preconditioner (strongly varying coefficients, where TF32 noise stalls PCG) and by the Rayleigh quotients of its setup.
namespace hsbp {
namespace tc {

struct DgemmParams {
  const double *A, *B;
  double *C;
  const double *scale;          // same indexing as C, or null
  int64_t strideA, strideB, strideC;
  int M, N, K;
  int64_t sam, sak, sbn, sbk, scm, scn;
};

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256)
k_dgemm_batched(DgemmParams p) {
  constexpr int T = 64, BKD = 16;
  __shared__ double As[T][BKD + 1], Bs[T][BKD + 1];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int m0 = blockIdx.x * T, n0 = blockIdx.y * T;
  const int64_t b = blockIdx.z;
  const double *A = p.A + b * p.strideA, *B = p.B + b * p.strideB;
  double acc[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
  for (int k0 = 0; k0 < p.K; k0 += BKD) {
    for (int idx = tid; idx < T * BKD; idx += 256) {
      int i, k;
      if (p.sam == 1) { i = idx % T; k = idx / T; } else { k = idx % BKD; i = idx / BKD; }
      As[i][k] = (m0 + i < p.M && k0 + k < p.K) ? A[(int64_t)(m0 + i) * p.sam + (int64_t)(k0 + k) * p.sak] : 0.0;
      if (p.sbn == 1) { i = idx % T; k = idx / T; } else { k = idx % BKD; i = idx / BKD; }
      Bs[i][k] = (n0 + i < p.N && k0 + k < p.K) ? B[(int64_t)(n0 + i) * p.sbn + (int64_t)(k0 + k) * p.sbk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BKD; k += 4) {
      const double a = As[8 * w + (lane >> 2)][k + (lane & 3)];
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma884(acc[j][0], acc[j][1], a, Bs[8 * j + (lane >> 2)][k + (lane & 3)]);
    }
    __syncthreads();
  }
  const int m = m0 + 8 * w + (lane >> 2);
  if (m >= p.M) return;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int n = n0 + 8 * j + (lane & 3) * 2 + q;
      if (n < p.N) {
        const int64_t o = b * p.strideC + (int64_t)m * p.scm + (int64_t)n * p.scn;
        p.C[o] = p.scale ? acc[j][q] * p.scale[o] : acc[j][q];
      }
    }
}

// x <- nearest TF32 value (static operands of the tensor-core GEMMs, once at setup)
__global__ void k_round_tf32(int64_t n, float *__restrict__ x) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = round_tf32(x[i]);
}

// out[b][j + n i] = in[b][i + m j]  (m x n column-major -> its transpose), fp32
__global__ void k_transpose_f32(int m, int n, const float *__restrict__ in, float *__restrict__ out) {
  const float *I = in + (int64_t)blockIdx.y * m * n;
  float *O = out + (int64_t)blockIdx.y * m * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (int64_t)m * n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % m), j = (int)(idx / m);
    O[j + (int64_t)n * i] = I[idx];
  }
}

}  // namespace tc
}  // namespace hsbp
