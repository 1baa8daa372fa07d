// Fast path of the volume part  y = A-tilde u  for uniform blocks with an even number of r-points
// (lines are then 16-byte aligned and can be moved by the TMA engine with cp.async.bulk).
//
//   k_march : all output lines j in [M, Ns+1-M)   -- interior in s, every r (r-closures included)
//   k_strip : the 2*M closure lines at the two s-ends of every block
//
// k_march streams a block line by line in s.  One CTA owns the full r-extent of a block and a
// chunk of output lines; an elected thread keeps NSTAGE lines of (u, crr, css, crs) in flight
// into a shared-memory ring with 1-D bulk copies that complete on mbarriers.  Every thread owns
// R consecutive r-points and keeps, in registers, sliding windows over s of u, css, crs and of the
// output accumulators, so each input element is read from HBM once per chunk (plus HALF halo
// lines at the chunk ends) and everything that needs r-neighbours is done through the shared
// line at arrival time:
//   at arrival of line jn : rr(jn) = scale * M(crr(:,jn)) u(:,jn)         -> acc[jn]
//                           t(jn)  = crs o (Qr u(:,jn)); acc[jn+o] += Qs[jn][jn+o] * t  (Qs^T t, push form)
//   at output of line jo = jn - HALF (window complete):
//                           ss(jo) from the u / css windows, w(jo) = crs o (Qs u)(jo),
//                           exchange w through shared memory, rs = Qr^T w, store y(:, jo).
// Algorithmic traffic: 40 B per point (u, crr, css, crs in, y out).  DESIGN.md section 4.
#pragma once
#include "hsbp_internal.h"
#include "sbp1d.cuh"

namespace hsbp {

// ---- PTX helpers: mbarrier + 1-D bulk copy (TMA) ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

constexpr int MARCH_NSTAGE = 4;
constexpr int MARCH_R = 2;

struct MarchParams {
  const double *crr, *css, *crs, *u;
  double *y;
  int Nr, Ns;         // uniform block size
  int nchunks;        // chunks of output lines per block
  int chunk;          // output lines per chunk
  int64_t nblocks;
};

template <int P> struct MarchCfg {
  using S = Sbp<P>;
  static constexpr int H = S::HALF;
  static constexpr int L = 2 * H + 1;      // window length
};

// interior-in-s stiffness row from the windows: sum_o M[jo][jo+o] u(jo+o), b = css window
template <int P, class B, class U> __device__ __forceinline__ double m_row_interior(B b, U u) {
  using S = Sbp<P>;
  double acc = 0.0;
#pragma unroll
  for (int o = -S::HALF; o <= S::HALF; ++o) acc += S::mint(0, o, b) * u(o);
  return acc;
}

// one marching step; PH = jn mod L selects the physical register slots at compile time
template <int P, int PH>
__device__ __forceinline__ void march_step(
    const int jn, const int j_out0, const int j_out1, const int i0, const bool own, const int Nr, const int Ns,
    const int Nrp, const double sc_rr, const double (&sc_ss)[MARCH_R],
    const double *__restrict__ su, const double *__restrict__ srr, const double *__restrict__ sss,
    const double *__restrict__ srs, double *__restrict__ swb, double *__restrict__ yblk,
    double (&wu)[MarchCfg<P>::L][MARCH_R], double (&wb)[MarchCfg<P>::L][MARCH_R],
    double (&wc)[MarchCfg<P>::L][MARCH_R], double (&acc)[MarchCfg<P>::L][MARCH_R]) {
  using S = Sbp<P>;
  constexpr int H = MarchCfg<P>::H, L = MarchCfg<P>::L, R = MARCH_R;
  constexpr int SN = PH;                              // slot of the new line jn
  constexpr int SO = (PH - H + L) % L;                // slot of the output line jo = jn - H
  const int jo = jn - H;
  const bool out = (jo >= j_out0) && (jo < j_out1);

  // ---- arrival of line jn ---------------------------------------------------------------
  if (own) {
    const double d_lo = S::d()[0];
    (void)d_lo;
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int i = i0 + q;
      if (i < Nrp) {
        wu[SN][q] = su[i];
        wb[SN][q] = sss[i];
        wc[SN][q] = srs[i];
        const double rr = m_apply<P>(i, Nr, [&](int l) { return srr[l]; }, [&](int l) { return su[l]; });
        const double t = srs[i] * q_apply<P>(i, Nr, [&](int l) { return su[l]; });
        // acc slot of line jn was reset when line jn-L left the window
        acc[SN][q] += sc_rr * rr;
        // push Qs^T t: line jn+o receives Qs[jn][jn+o] * t = d[o+H] * t (interior rows only, see header)
#pragma unroll
        for (int o = -H; o <= H; ++o)
          if (o != 0) acc[(PH + o + L) % L][q] += S::d()[o + H] * t;
      }
    }
  }
  // ---- w(jo) = crs(jo) * (Qs u)(jo) from the windows, exchanged through shared memory ------
  if (out && own) {
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int i = i0 + q;
      if (i < Nrp) {
        double qs = 0.0;
#pragma unroll
        for (int o = -H; o <= H; ++o)
          if (o != 0) qs += S::d()[o + H] * wu[(SO + o + L) % L][q];
        swb[i] = wc[SO][q] * qs;
      }
    }
  }
  __syncthreads();
  // (the caller refills the ring stage of line jn right after this barrier)
  if (out && own) {
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int i = i0 + q;
      if (i < Nrp) {
        const double ss = m_row_interior<P>([&](int k) { return wb[(SO + k + L) % L][q]; },
                                            [&](int k) { return wu[(SO + k + L) % L][q]; });
        const double rs = qt_apply<P>(i, Nr, [&](int l) { return swb[l]; });
        yblk[(int64_t)Nrp * jo + i] = acc[SO][q] + sc_ss[q] * ss + rs;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < R; ++q) acc[SO][q] = 0.0;        // slot now belongs to line jo + L
}

template <int P>
__global__ void __launch_bounds__(512)
k_march(const MarchParams prm) {
  using S = Sbp<P>;
  constexpr int H = MarchCfg<P>::H, L = MarchCfg<P>::L, R = MARCH_R, NST = MARCH_NSTAGE;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int Nr = prm.Nr, Ns = prm.Ns, Nrp = Nr + 1, Nsp = Ns + 1;
  const int LW = (Nrp + 3) & ~1;                     // padded line length in doubles (even)
  double *ring = reinterpret_cast<double *>(smem_raw);               // NST * 4 * LW
  double *wbuf = ring + (size_t)NST * 4 * LW;                        // 2 * LW
  uint64_t *full = reinterpret_cast<uint64_t *>(wbuf + 2 * LW);      // NST

  const int64_t e = blockIdx.x / prm.nchunks;
  const int c = (int)(blockIdx.x - e * prm.nchunks);
  const int j_first = S::M, j_last = Nsp - S::M;     // output lines of k_march: [j_first, j_last)
  const int j_out0 = j_first + c * prm.chunk;
  const int j_out1 = min(j_last, j_out0 + prm.chunk);
  if (j_out0 >= j_out1) return;
  const int jlo = j_out0 - H, jhi = j_out1 - 1 + H;  // input lines [jlo, jhi]
  const int nlines = jhi - jlo + 1;
  const int64_t voff = e * (int64_t)Nrp * Nsp;
  const uint32_t line_bytes = (uint32_t)Nrp * 8u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int n) {      // line jlo + n into stage n % NST  (thread 0 only)
    const int st = n % NST;
    double *dst = ring + (size_t)st * 4 * LW;
    const int64_t g = voff + (int64_t)Nrp * (jlo + n);
    mbar_expect_tx(&full[st], 4u * line_bytes);
    bulk_g2s(dst, prm.u + g, line_bytes, &full[st]);
    bulk_g2s(dst + LW, prm.crr + g, line_bytes, &full[st]);
    bulk_g2s(dst + 2 * LW, prm.css + g, line_bytes, &full[st]);
    bulk_g2s(dst + 3 * LW, prm.crs + g, line_bytes, &full[st]);
  };
  if (threadIdx.x == 0)
    for (int n = 0; n < NST && n < nlines; ++n) issue(n);

  const int i0 = threadIdx.x * R;
  const bool own = i0 < Nrp;
  const double hr = 2.0 / Nr, hs = 2.0 / Ns;
  const double sc_rr = hs / hr;                       // Hs[j]/hr with Hs[j] = hs on interior lines
  double sc_ss[R];
#pragma unroll
  for (int q = 0; q < R; ++q) sc_ss[q] = (i0 + q < Nrp) ? hr * hweight<P>(i0 + q, Nr) / hs : 0.0;
  double wu[L][R], wb[L][R], wc[L][R], acc[L][R];
#pragma unroll
  for (int k = 0; k < L; ++k)
#pragma unroll
    for (int q = 0; q < R; ++q) { wu[k][q] = 0.0; wb[k][q] = 0.0; wc[k][q] = 0.0; acc[k][q] = 0.0; }
  double *yblk = prm.y + voff;

  // lines are processed in groups of L so that register slots are compile-time constants;
  // the first group starts at a line index that is a multiple of L (dummy steps before jlo are skipped)
  const int jstart = (jlo / L) * L;
  for (int jb = jstart; jb <= jhi; jb += L) {
#define HSBP_MARCH_STEP(PH_)                                                                          \
    if (PH_ < L) {                                                                                   \
      const int jn = jb + PH_;                                                                        \
      if (jn >= jlo && jn <= jhi) {                                                                   \
        const int n = jn - jlo, st = n % NST;                                                         \
        mbar_wait(&full[st], (uint32_t)((n / NST) & 1));                                              \
        const double *sb = ring + (size_t)st * 4 * LW;                                                \
        march_step<P, (PH_ < L ? PH_ : 0)>(jn, j_out0, j_out1, i0, own, Nr, Ns, Nrp, sc_rr, sc_ss, sb, sb + LW,  \
                                           sb + 2 * LW, sb + 3 * LW, wbuf + (size_t)(n & 1) * LW, yblk, wu, wb, \
                                           wc, acc);                                                  \
        if (threadIdx.x == 0 && n + NST < nlines) { fence_proxy_async(); issue(n + NST); }            \
      }                                                                                               \
    }
    HSBP_MARCH_STEP(0) HSBP_MARCH_STEP(1) HSBP_MARCH_STEP(2) HSBP_MARCH_STEP(3)
    HSBP_MARCH_STEP(4) HSBP_MARCH_STEP(5) HSBP_MARCH_STEP(6)
#undef HSBP_MARCH_STEP
  }
}

// ---- s-end strips: output lines [0, M) and [Ns+1-M, Ns+1) of every block ---------------------
// grid.x = 2 * nblocks; the tile (all r, NL = M + HALF lines) lives in shared memory and the
// generic 1-D operators are evaluated on it.
template <int P>
__global__ void __launch_bounds__(256)
k_strip(const MarchParams prm) {
  using S = Sbp<P>;
  constexpr int NL = S::M + S::HALF;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Nr = prm.Nr, Ns = prm.Ns, Nrp = Nr + 1, Nsp = Ns + 1;
  double *su = reinterpret_cast<double *>(smem_raw);      // [NL][Nrp] each
  double *srr = su + (size_t)NL * Nrp;
  double *sss = srr + (size_t)NL * Nrp;
  double *srs = sss + (size_t)NL * Nrp;
  double *st = srs + (size_t)NL * Nrp;                    // t  [NL][Nrp]
  double *sw = st + (size_t)NL * Nrp;                     // w  [M][Nrp]
  const int64_t e = blockIdx.x >> 1;
  const bool tail = blockIdx.x & 1;
  const int lbase = tail ? Nsp - NL : 0;                  // first line held in the tile
  const int obase = tail ? Nsp - S::M : 0;                // first output line
  const int64_t voff = e * (int64_t)Nrp * Nsp;
  const int ntile = NL * Nrp;
  {
    const int64_t g0 = voff + (int64_t)Nrp * lbase;
    for (int idx = threadIdx.x; idx < ntile; idx += blockDim.x) {
      su[idx] = prm.u[g0 + idx]; srr[idx] = prm.crr[g0 + idx];
      sss[idx] = prm.css[g0 + idx]; srs[idx] = prm.crs[g0 + idx];
    }
  }
  __syncthreads();
  // t = crs o (Qr u) on every held line; w = crs o (Qs u) on the output lines
  for (int idx = threadIdx.x; idx < ntile; idx += blockDim.x) {
    const int l = idx / Nrp, i = idx - l * Nrp;
    const double *ul = su + (size_t)l * Nrp;
    st[idx] = srs[idx] * q_apply<P>(i, Nr, [&](int k) { return ul[k]; });
  }
  for (int idx = threadIdx.x; idx < S::M * Nrp; idx += blockDim.x) {
    const int lo = idx / Nrp, i = idx - lo * Nrp;
    const int j = obase + lo;
    const double qs = q_apply<P>(j, Ns, [&](int k) { return su[(size_t)(k - lbase) * Nrp + i]; });
    sw[idx] = srs[(size_t)(j - lbase) * Nrp + i] * qs;
  }
  __syncthreads();
  const double hr = 2.0 / Nr, hs = 2.0 / Ns;
  for (int idx = threadIdx.x; idx < S::M * Nrp; idx += blockDim.x) {
    const int lo = idx / Nrp, i = idx - lo * Nrp;
    const int j = obase + lo;
    const double *ul = su + (size_t)(j - lbase) * Nrp, *bl = srr + (size_t)(j - lbase) * Nrp;
    const double rr = m_apply<P>(i, Nr, [&](int k) { return bl[k]; }, [&](int k) { return ul[k]; });
    const double ss = m_apply<P>(j, Ns, [&](int k) { return sss[(size_t)(k - lbase) * Nrp + i]; },
                                 [&](int k) { return su[(size_t)(k - lbase) * Nrp + i]; });
    const double sr = qt_apply<P>(j, Ns, [&](int k) { return st[(size_t)(k - lbase) * Nrp + i]; });
    const double rs = qt_apply<P>(i, Nr, [&](int k) { return sw[(size_t)lo * Nrp + k]; });
    prm.y[voff + (int64_t)Nrp * j + i] =
        (hs * hweight<P>(j, Ns) / hr) * rr + (hr * hweight<P>(i, Nr) / hs) * ss + sr + rs;
  }
}

template <int P> static size_t march_smem(int Nrp) {
  const int LW = (Nrp + 3) & ~1;
  return (size_t)(MARCH_NSTAGE * 4 + 2) * LW * sizeof(double) + MARCH_NSTAGE * sizeof(uint64_t);
}
template <int P> static size_t strip_smem(int Nrp) {
  using S = Sbp<P>;
  return (size_t)((S::M + S::HALF) * 5 + S::M) * Nrp * sizeof(double);
}

template <int P> static bool march_eligible(const hsbp_blocks *b) {
  using S = Sbp<P>;
  if (!b->uniform) return false;
  const int Nrp = b->max_Nr + 1, Nsp = b->max_Ns + 1;
  if (Nrp & 1) return false;                                   // 16-byte aligned lines for the bulk copies
  if (Nrp > 2 * 512) return false;
  if (Nsp < 2 * S::M + 2 * S::HALF + 1) return false;          // needs interior lines between the strips
  if (strip_smem<P>(Nrp) > b->ctx->smem_optin || march_smem<P>(Nrp) > b->ctx->smem_optin) return false;
  return true;
}

template <int P> static int vol_march(hsbp_blocks *b, const double *u, double *y) {
  using S = Sbp<P>;
  hsbp_ctx *ctx = b->ctx;
  if (((uintptr_t)u & 15) || ((uintptr_t)y & 15)) {
    ctx->err = "hsbp_apply: u / y must be 16-byte aligned for the marching kernel";
    return HSBP_ERR_ARG;
  }
  MarchParams prm;
  prm.crr = b->d_crr; prm.css = b->d_css; prm.crs = b->d_crs; prm.u = u; prm.y = y;
  prm.Nr = b->max_Nr; prm.Ns = b->max_Ns; prm.nblocks = b->nblocks;
  const int Nrp = prm.Nr + 1, Nsp = prm.Ns + 1;
  const int nout = Nsp - 2 * S::M;
  // chunks of about 64 output lines, at least enough CTAs to fill the machine a few times over
  int nchunks = std::max(1, (nout + 63) / 64);
  while ((int64_t)nchunks * b->nblocks < 4LL * ctx->sm_count && nout / (nchunks * 2) >= 16) nchunks *= 2;
  prm.nchunks = nchunks;
  prm.chunk = (nout + nchunks - 1) / nchunks;
  const int nthreads = std::min(512, ((Nrp + MARCH_R - 1) / MARCH_R + 31) & ~31);
  const size_t sm = march_smem<P>(Nrp), ss = strip_smem<P>(Nrp);
  static bool attr_set[8] = {false};
  if (!attr_set[P]) {
    if (cudaFuncSetAttribute(k_march<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin) != cudaSuccess ||
        cudaFuncSetAttribute(k_strip<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin) != cudaSuccess) {
      ctx->err = "cudaFuncSetAttribute(max dynamic shared memory) failed";
      return HSBP_ERR_CUDA;
    }
    attr_set[P] = true;
  }
  k_march<P><<<(unsigned)(b->nblocks * nchunks), nthreads, sm, ctx->stream>>>(prm);
  cudaError_t e1 = cudaGetLastError();
  k_strip<P><<<(unsigned)(2 * b->nblocks), 256, ss, ctx->stream>>>(prm);
  cudaError_t e2 = cudaGetLastError();
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    ctx->err = std::string("k_march/k_strip: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2);
    return HSBP_ERR_CUDA;
  }
  return HSBP_OK;
}

}  // namespace hsbp
