// Line-marching kernel for the volume part  y = A-tilde u  of uniform blocks
//   A u = Arr u + Ass u + Qs^T[crs o (Qr u)] + Qr^T[crs o (Qs u)]      (global_curved.jl:261-353)
// including every closure (r- and s-ends), so that one launch covers all points of all blocks.
//
// Work decomposition
//   CTA    = one block x one chunk of s-lines; the CTA spans the full r-extent of the block, every
//            thread owns R consecutive r-points.  The lower half of a block is marched upwards from
//            line 0, the upper half downwards from line Ns (the operators are mirror-symmetric:
//            M is persymmetric, Q changes sign), so closures only ever occur at the *start* of a march.
//   lines  = streamed through a shared-memory ring by 1-D bulk copies (TMA, cp.async.bulk) that
//            complete on mbarriers; NST-1 lines of (u, crr, css, crs) are in flight per CTA.
//   s-dir  = register windows per point: u (2H+1 lines), scaled css (2H), crs (H+1) and 2H+1
//            accumulators; the output lags the newest line by H lines.
//   r-dir  = neighbours come from the shared line (u, crr) or, for w = crs o (Qs u), from a
//            double-buffered shared line written by the owners (one __syncthreads per line).
//
// Arithmetic.  The variable-coefficient stiffness matrix M(b) (diagonal_sbp.jl:474-746) is symmetric
// with zero row sums, so (M u)_i = sum_j M_ij (u_j - u_i); only the couplings M_ij are formed and each
// is used for both rows ("pair form", tools/sbp_coeffs.py).  Closure rows:
//   r-ends   every G lines the CTA evaluates the MC closure rows of M(crr) u and the BM closure rows of
//            Qr u for the next G lines, one (line, end) per thread, from global memory (L2) into a small
//            shared table; the edge lanes pick their values up when the line arrives.
//   s-ends   closure rows of M(css) u and Qs u are evaluated directly from global memory at output
//            time; contributions of the dense BM x BM closure block of Qs^T, which reach further than
//            the H-line lag, are added to y by read-modify-write from the thread that owns the points.
// The algorithm is emulated step by step on the CPU in tools/proto_sweep.py and checked there against
// the oracle's assembled operator (tests/test_sweep_algorithm.py).
//
// Algorithmic traffic: 40 B per point (u, crr, css, crs in, y out); DESIGN.md section 4.
#pragma once
#include "hsbp_internal.h"
#include "sbp1d.cuh"
#include "sweep_tables_gen.h"

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

constexpr int SW_NST = 3;              // ring stages: one being consumed, two in flight
constexpr int SW_MAX_THREADS = 256;

struct SweepParams {
  const double *crr, *css, *crs, *u;
  double *y;
  int Nr, Ns;         // uniform block size
  int ncs;            // chunks per side (a block is 2*ncs CTAs)
  int K;              // lines [0, K) are marched upwards, lines [K, Ns] downwards
  int per_up, per_dn; // output lines per chunk on either side
};

template <int P> struct SweepCfg {
  using T = SweepTab<P>;
  static constexpr int H = T::H, W = 2 * T::H + 1, LB = 2 * T::H;
  static constexpr int PAD = (T::H + 1) & ~1;               // halo of a shared line, even (16-byte vectors)
  static constexpr int CLW = T::MC + T::BM;                 // r-closure table: MC rows of M u, BM rows of Q u
  __device__ static const double *hw() { return P == 2 ? c_sw_hw2 : (P == 4 ? c_sw_hw4 : c_sw_hw6); }
  __device__ static const double *Qc() { return P == 2 ? c_sw_Qc2 : (P == 4 ? c_sw_Qc4 : c_sw_Qc6); }
  template <int O> __device__ static constexpr double D() {
    if constexpr (O == 1) return T::D1;
    else if constexpr (O == 2) return T::D2;
    else return T::D3;
  }
};

// coupling M[a][a+O] of the interior stencil; b(s) returns the coefficient at index a+s
// (diagonal_sbp.jl:495-503, 567-582, 719-727; tools/sbp_coeffs.py INTERIOR_PAIR)
template <int P, int O, class B> __device__ __forceinline__ double pair_coef(B b) {
  if constexpr (P == 2) {
    return -0.5 * (b(0) + b(1));
  } else if constexpr (P == 4) {
    if constexpr (O == 1) return -(1.0 / 6.0) * (b(2) + b(-1)) - 0.5 * (b(1) + b(0));
    else return 0.125 * (b(2) + b(0)) - (1.0 / 6.0) * b(1);
  } else {
    if constexpr (O == 1) return -(1.0 / 40.0) * (b(-2) + b(3)) - (3.0 / 10.0) * (b(-1) + b(2)) - (17.0 / 40.0) * (b(0) + b(1));
    else if constexpr (O == 2) return (1.0 / 20.0) * (b(-1) + b(3)) + (7.0 / 40.0) * (b(0) + b(2)) - (3.0 / 10.0) * b(1);
    else return -(11.0 / 360.0) * (b(0) + b(3)) + (1.0 / 40.0) * (b(1) + b(2));
  }
}

// compile-time loop over the offsets 1..H
template <int O, int H, class F> __device__ __forceinline__ void for_offsets(F &&f) {
  if constexpr (O <= H) {
    f(std::integral_constant<int, O>{});
    for_offsets<O + 1, H>(f);
  }
}

// ---- rarely executed closure evaluations, kept out of line so that their temporaries do not add to
// the register footprint of the marching loop ----------------------------------------------------
// r-end closure rows of one line: cl[0..MC) = rows of M(crr) u, cl[MC..MC+BM) = rows of Qr u.
// pb / pu point at the end point of the line, sg = +1 (near end) or -1 (far end, mirrored: Q flips sign).
template <int P>
__device__ __noinline__ void sweep_rclosure(const double *__restrict__ pb, const double *__restrict__ pu, int sg,
                                            double *cl) {
  using T = SweepTab<P>;
  double b[T::NK], uu[T::NK], qq[T::BM];
#pragma unroll
  for (int k = 0; k < T::NK; ++k) { b[k] = __ldg(pb + sg * k); uu[k] = __ldg(pu + sg * k); }
  d2_closure_rows<P>(b, uu, cl);
  q_closure_rows<P>(uu, qq);
#pragma unroll
  for (int k = 0; k < T::BM; ++k) cl[T::MC + k] = sg < 0 ? -qq[k] : qq[k];
}
// s-end closure row `row` of M(css) u for one column, straight from memory (pb / pu: the column's point on
// marching line 0, lstride: distance between marching lines)
template <int P>
__device__ __noinline__ double sweep_sclosure_row(int row, const double *__restrict__ pb, const double *__restrict__ pu,
                                                  int64_t lstride) {
  using T = SweepTab<P>;
  double b[T::NK], uu[T::NK];
#pragma unroll
  for (int k = 0; k < T::NK; ++k) { b[k] = __ldg(pb + k * lstride); uu[k] = __ldg(pu + k * lstride); }
  return d2_closure_row<P>(row, b, uu);
}

template <int P, int R>
__global__ void __launch_bounds__(SW_MAX_THREADS)
k_sweep(const SweepParams prm) {
  using T = SweepTab<P>;
  using C = SweepCfg<P>;
  constexpr int H = C::H, W = C::W, LB = C::LB, PAD = C::PAD, CLW = C::CLW, NST = SW_NST;
  constexpr int MC = T::MC, NK = T::NK, BM = T::BM, BN = T::BN;
  constexpr int NV = R + 2 * PAD;                         // values of a line a thread looks at
  static_assert(R % 2 == 0 && PAD >= H, "layout");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int Nr = prm.Nr, Ns = prm.Ns, Nrp = Nr + 1, Nsp = Ns + 1;
  const int LW = Nrp + 2 * PAD;
  const int G = nthreads >> 1;                            // lines per r-closure table refill
  double *ring = reinterpret_cast<double *>(smem_raw);    // [NST][4][LW]
  double *wbuf = ring + (size_t)NST * 4 * LW;             // [2][LW]
  double *clbuf = wbuf + 2 * LW;                          // [G][2][CLW]
  uint64_t *full = reinterpret_cast<uint64_t *>(clbuf + (size_t)G * 2 * CLW);

  // ---- which chunk --------------------------------------------------------------------------
  const int nch = 2 * prm.ncs;
  const int64_t e = blockIdx.x / nch;
  const int c = (int)(blockIdx.x - e * nch);
  const bool up = c < prm.ncs;
  const int cc = up ? c : c - prm.ncs;
  const int nside = up ? prm.K : Nsp - prm.K;
  const int per = up ? prm.per_up : prm.per_dn;
  const int o0 = cc * per, o1 = min(nside, o0 + per);     // output lines [o0, o1), marching coordinates
  if (o0 >= o1) return;
  const bool prologue = (o0 == 0);
  const int jstart = prologue ? 0 : o0 - H, jend = o1 - 1 + H;
  const int nlines = jend - jstart + 1;
  const double sig = up ? 1.0 : -1.0;
  const int64_t lstride = up ? (int64_t)Nrp : -(int64_t)Nrp;
  const int64_t base = e * (int64_t)Nrp * Nsp + (up ? 0 : (int64_t)Ns * Nrp);   // marching line j starts at base + j*lstride
  const uint32_t line_bytes = (uint32_t)Nrp * 8u;

  // ---- one-time setup: zero the halos of the shared lines, barriers --------------------------
  for (int idx = tid; idx < (NST * 4 + 2) * 2 * PAD; idx += nthreads) {
    const int line = idx / (2 * PAD), k = idx - line * (2 * PAD);
    ring[(size_t)line * LW + (k < PAD ? k : Nrp + k)] = 0.0;
  }
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int n) {      // marching line jstart + n into stage n % NST  (thread 0 only)
    const int st = n % NST;
    double *dst = ring + (size_t)st * 4 * LW + PAD;
    const int64_t g = base + (int64_t)(jstart + n) * lstride;
    mbar_expect_tx(&full[st], 4u * line_bytes);
    bulk_g2s(dst, prm.u + g, line_bytes, &full[st]);
    bulk_g2s(dst + LW, prm.crr + g, line_bytes, &full[st]);
    bulk_g2s(dst + 2 * LW, prm.css + g, line_bytes, &full[st]);
    bulk_g2s(dst + 3 * LW, prm.crs + g, line_bytes, &full[st]);
  };
  if (tid == 0) {
    fence_proxy_async();
    for (int n = 0; n < NST && n < nlines; ++n) issue(n);
  }

  const int i0 = tid * R;
  const bool own = i0 < Nrp;
  const double hr = 2.0 / Nr, hs = 2.0 / Ns;
  const double *hwt = C::hw();
  const double *qc = C::Qc();
  double sc_ss[R];                                        // Hr[i] / hs  (global_curved.jl:313-322)
#pragma unroll
  for (int q = 0; q < R; ++q) {
    const int i = i0 + q;
    const double hwi = i < BM ? hwt[i] : (i > Nr - BM && i <= Nr ? hwt[Nr - i] : 1.0);
    sc_ss[q] = hr * hwi / hs;
  }
  const double *gu = prm.u + base + i0;                   // + j*lstride: this thread's points on marching line j
  const double *gss = prm.css + base + i0;
  double *gy = prm.y + base + i0;

  double uw[W][R], bw[LB][R], cw[H + 1][R], acc[W][R];
#pragma unroll
  for (int q = 0; q < R; ++q) {
#pragma unroll
    for (int k = 0; k < W; ++k) { uw[k][q] = 0.0; acc[k][q] = 0.0; }
#pragma unroll
    for (int k = 0; k < LB; ++k) bw[k][q] = 0.0;
#pragma unroll
    for (int k = 0; k <= H; ++k) cw[k][q] = 0.0;
  }
  if (prologue && own) {                                  // lines that collect read-modify-write contributions
    for (int l = 0; l < BM; ++l)
#pragma unroll
      for (int q = 0; q < R; ++q) gy[l * lstride + q] = 0.0;
  }

#pragma unroll 1
  for (int n = 0; n < nlines; ++n) {
    const int j = jstart + n;
    const int st = n % NST;
    const int ng = n % G;

    // ---- r-closure table for lines j .. j+G-1: thread -> (line, end) ---------------------------
    if (ng == 0) {
      const int jj = j + (tid >> 1), side = tid & 1;
      if (jj <= jend) {
        const int64_t g = base + (int64_t)jj * lstride + (side ? Nr : 0);
        sweep_rclosure<P>(prm.crr + g, prm.u + g, side ? -1 : 1, clbuf + (size_t)tid * CLW);   // (tid>>1)*2 + side == tid
      }
      __syncthreads();
    }

    mbar_wait(&full[st], (uint32_t)((n / NST) & 1));
    const double *sb = ring + (size_t)st * 4 * LW;
    const int jo = j - H;
    const bool outp = (jo >= o0) && (jo < o1);
    double *wb = wbuf + (size_t)(n & 1) * LW;

    if (own) {
      // ---- shift the windows, take in line j ------------------------------------------------
      double U[NV], Bq[NV];
#pragma unroll
      for (int k = 0; k < NV / 2; ++k) {
        const double2 a = *reinterpret_cast<const double2 *>(sb + i0 + 2 * k);
        const double2 b2 = *reinterpret_cast<const double2 *>(sb + LW + i0 + 2 * k);
        U[2 * k] = a.x; U[2 * k + 1] = a.y; Bq[2 * k] = b2.x; Bq[2 * k + 1] = b2.y;
      }
#pragma unroll
      for (int q = 0; q < R; ++q) {
#pragma unroll
        for (int k = 0; k < W - 1; ++k) { uw[k][q] = uw[k + 1][q]; acc[k][q] = acc[k + 1][q]; }
#pragma unroll
        for (int k = 0; k < LB - 1; ++k) bw[k][q] = bw[k + 1][q];
#pragma unroll
        for (int k = 0; k < H; ++k) cw[k][q] = cw[k + 1][q];
        uw[W - 1][q] = U[PAD + q];
        acc[W - 1][q] = 0.0;
      }
#pragma unroll
      for (int k = 0; k < R / 2; ++k) {
        const double2 a = *reinterpret_cast<const double2 *>(sb + 2 * LW + PAD + i0 + 2 * k);
        const double2 b2 = *reinterpret_cast<const double2 *>(sb + 3 * LW + PAD + i0 + 2 * k);
        bw[LB - 1][2 * k] = a.x * sc_ss[2 * k]; bw[LB - 1][2 * k + 1] = a.y * sc_ss[2 * k + 1];
        cw[H][2 * k] = b2.x; cw[H][2 * k + 1] = b2.y;
      }

      // ---- r-direction on line j: rr = M(crr) u (pair form), qr = Qr u -------------------------
      double rr[R], qr[R];
#pragma unroll
      for (int q = 0; q < R; ++q) { rr[q] = 0.0; qr[q] = 0.0; }
      for_offsets<1, H>([&](auto Oc) {
        constexpr int O = decltype(Oc)::value;
        double f[R + O];
#pragma unroll
        for (int k = 0; k < R + O; ++k) {                 // pair (a, a+O), a = i0 - O + k
          const int ia = PAD - O + k;
          const double cf = pair_coef<P, O>([&](int s) { return Bq[ia + s]; });
          f[k] = cf * (U[ia + O] - U[ia]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) {
          rr[q] += f[q + O] - f[q];
          qr[q] += C::template D<O>() * (U[PAD + q + O] - U[PAD + q - O]);
        }
      });
      if (i0 < MC) {                                      // near r-end: closure rows from the table
        const double *cl = clbuf + (size_t)(2 * ng) * CLW;
#pragma unroll
        for (int q = 0; q < R; ++q) {
          if (i0 + q < MC) rr[q] = cl[i0 + q];
          if (i0 + q < BM) qr[q] = cl[MC + i0 + q];
        }
      }
      if (i0 + R > Nrp - MC) {                            // far r-end
        const double *cl = clbuf + (size_t)(2 * ng + 1) * CLW;
#pragma unroll
        for (int q = 0; q < R; ++q) {
          const int m = Nr - (i0 + q);
          if (m < MC) rr[q] = cl[m];
          if (m < BM) qr[q] = cl[MC + m];
        }
      }
      const double scrr = (hs / hr) * (j < BM ? hwt[j] : 1.0);       // Hs[j] / hr  (global_curved.jl:261-268)
      double t[R];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        acc[H][q] = fma(scrr, rr[q], acc[H][q]);
        t[q] = cw[H][q] * qr[q];                          // t = crs o (Qr u) on line j
      }
      // ---- Qs^T t, pushed from row j: (Qs^T t)(l) += Qs[j][l] t(j) ------------------------------
      if (j >= BM) {
        for_offsets<1, H>([&](auto Oc) {
          constexpr int O = decltype(Oc)::value;
          const double d = sig * C::template D<O>();
#pragma unroll
          for (int q = 0; q < R; ++q) {
            acc[H + O][q] = fma(d, t[q], acc[H + O][q]);
            acc[H - O][q] = fma(-d, t[q], acc[H - O][q]);
          }
        });
      } else {
        for (int l = 0; l < BM; ++l) {                    // dense closure block: straight into y
          const double d = sig * qc[j * BN + l];
          if (d != 0.0) {
            double *yl = gy + l * lstride;
#pragma unroll
            for (int q = 0; q < R; ++q) yl[q] = fma(d, t[q], yl[q]);
          }
        }
        for_offsets<1, H>([&](auto Oc) {
          constexpr int O = decltype(Oc)::value;
          const int l = j + O;
          if (l >= BM && l < BN) {
            const double d = sig * qc[j * BN + l];
#pragma unroll
            for (int q = 0; q < R; ++q) acc[H + O][q] = fma(d, t[q], acc[H + O][q]);
          }
        });
      }
      // ---- s-direction stiffness: pairs (a, a+O), a = j-H; rows a <-> acc[0], a+O <-> acc[O] ------
      {
        const int a = j - H;
        const bool row_a_closure = prologue && (a < MC);
        for_offsets<1, H>([&](auto Oc) {
          constexpr int O = decltype(Oc)::value;
          if (!(prologue && (a + O < MC))) {
#pragma unroll
            for (int q = 0; q < R; ++q) {
              const double cf = pair_coef<P, O>([&](int s) { return bw[H - 1 + s][q]; });
              const double f = cf * (uw[H + O][q] - uw[H][q]);
              if (!row_a_closure) acc[0][q] += f;
              acc[O][q] -= f;
            }
          }
        });
      }
      // ---- w = crs o (Qs u) on the output line, shared with the r-neighbours ------------------
      if (outp) {
        double qs[R];
#pragma unroll
        for (int q = 0; q < R; ++q) qs[q] = 0.0;
        if (jo >= BM) {
          for_offsets<1, H>([&](auto Oc) {
            constexpr int O = decltype(Oc)::value;
            const double d = sig * C::template D<O>();
#pragma unroll
            for (int q = 0; q < R; ++q) qs[q] = fma(d, uw[H + O][q] - uw[H - O][q], qs[q]);
          });
        } else {
          for (int l = 0; l < BN; ++l) {
            const double d = sig * qc[jo * BN + l];
            if (d != 0.0) {
              const double *ul = gu + l * lstride;
#pragma unroll
              for (int q = 0; q < R; ++q) qs[q] = fma(d, ul[q], qs[q]);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < R / 2; ++k)
          *reinterpret_cast<double2 *>(wb + PAD + i0 + 2 * k) =
              make_double2(cw[0][2 * k] * qs[2 * k], cw[0][2 * k + 1] * qs[2 * k + 1]);
      }
    }
    __syncthreads();
    if (tid == 0 && n + NST < nlines) { fence_proxy_async(); issue(n + NST); }    // stage st is free again
    if (own && outp) {
      // ---- rs = Qr^T w, then the output line ------------------------------------------------
      double Wv[NV], val[R];
#pragma unroll
      for (int k = 0; k < NV / 2; ++k) {
        const double2 a = *reinterpret_cast<const double2 *>(wb + i0 + 2 * k);
        Wv[2 * k] = a.x; Wv[2 * k + 1] = a.y;
      }
#pragma unroll
      for (int q = 0; q < R; ++q) val[q] = 0.0;
      for_offsets<1, H>([&](auto Oc) {
        constexpr int O = decltype(Oc)::value;
#pragma unroll
        for (int q = 0; q < R; ++q) val[q] = fma(-C::template D<O>(), Wv[PAD + q + O] - Wv[PAD + q - O], val[q]);
      });
      if (i0 < BM) {                                      // near r-end rows of Qr^T
        const double *w0 = wb + PAD;
#pragma unroll
        for (int q = 0; q < R; ++q)
          if (i0 + q < BM) val[q] = qt_closure_row<P>(i0 + q, w0);
      }
      if (i0 + R > Nrp - BM) {                            // far r-end: mirrored, sign flipped
        double wr[BN];
#pragma unroll
        for (int k = 0; k < BN; ++k) wr[k] = wb[PAD + Nr - k];
#pragma unroll
        for (int q = 0; q < R; ++q) {
          const int m = Nr - (i0 + q);
          if (m < BM) val[q] = -qt_closure_row<P>(m, wr);
        }
      }
#pragma unroll
      for (int q = 0; q < R; ++q) val[q] += acc[0][q];
      if (prologue && jo < MC) {                          // closure row jo of M(css) u, straight from memory
#pragma unroll
        for (int q = 0; q < R; ++q)
          val[q] = fma(sc_ss[q], sweep_sclosure_row<P>(jo, gss + q, gu + q, lstride), val[q]);
      }
      double *yl = gy + jo * lstride;
      if (prologue && jo < BM) {
#pragma unroll
        for (int q = 0; q < R; ++q) yl[q] += val[q];
      } else {
#pragma unroll
        for (int k = 0; k < R / 2; ++k)
          *reinterpret_cast<double2 *>(yl + 2 * k) = make_double2(val[2 * k], val[2 * k + 1]);
      }
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
template <int P> static size_t sweep_smem(int Nrp, int nthreads) {
  using C = SweepCfg<P>;
  const int LW = Nrp + 2 * C::PAD;
  return (size_t)(SW_NST * 4 + 2) * LW * sizeof(double) + (size_t)(nthreads / 2) * 2 * C::CLW * sizeof(double) +
         SW_NST * sizeof(uint64_t);
}

static int sweep_points_per_thread(const hsbp_blocks *b) {
  const int Nrp = b->max_Nr + 1;
  if (b->sweep_r_override == 2 || (b->sweep_r_override == 4 && Nrp % 4 == 0)) return b->sweep_r_override;
  return (Nrp % 4 == 0 && Nrp >= 128) ? 4 : 2;
}

template <int P> static bool sweep_eligible(const hsbp_blocks *b) {
  if (!b->uniform) return false;
  const int Nrp = b->max_Nr + 1, Nsp = b->max_Ns + 1;
  if (Nrp & 1) return false;                                   // 16-byte aligned lines for the bulk copies
  if (Nrp < 32 || Nsp < 32) return false;                      // room for both closures and a chunk per side
  const int R = sweep_points_per_thread(b);
  const int nthreads = ((Nrp / R) + 31) & ~31;
  if (nthreads > SW_MAX_THREADS) return false;
  return sweep_smem<P>(Nrp, nthreads) <= b->ctx->smem_optin;
}

template <int P, int R> static int sweep_launch(hsbp_blocks *b, const double *u, double *y) {
  hsbp_ctx *ctx = b->ctx;
  const int Nrp = b->max_Nr + 1, Nsp = b->max_Ns + 1;
  const int nthreads = ((Nrp / R) + 31) & ~31;
  const size_t sm = sweep_smem<P>(Nrp, nthreads);
  static bool attr_set = false;
  static int ctas_per_sm = 1;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_sweep<P, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin) != cudaSuccess) {
      ctx->err = "cudaFuncSetAttribute(max dynamic shared memory) failed";
      return HSBP_ERR_CUDA;
    }
    attr_set = true;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_sweep<P, R>, nthreads, sm) != cudaSuccess || ctas_per_sm < 1)
    ctas_per_sm = 1;
  // chunks per side: enough CTAs to fill the machine in an almost whole number of waves, chunks of
  // at least 16 output lines (each chunk re-reads 2H halo lines)
  const int64_t slots = (int64_t)ctx->sm_count * ctas_per_sm;
  const int K = Nsp / 2;
  int best = 1;
  double best_eff = -1.0;
  for (int ncs = 1; ncs <= 16; ++ncs) {
    const int per = (K + ncs - 1) / ncs;
    if (ncs > 1 && per < 16) break;
    const int64_t ctas = b->nblocks * 2 * ncs;
    const double waves = (double)ctas / (double)slots;
    const double eff = waves / std::ceil(waves) * ((double)per / (per + 2 * SweepCfg<P>::H));
    if (eff > best_eff + 1e-9) { best_eff = eff; best = ncs; }
  }
  if (b->sweep_ncs_override > 0) best = b->sweep_ncs_override;
  SweepParams prm;
  prm.crr = b->d_crr; prm.css = b->d_css; prm.crs = b->d_crs; prm.u = u; prm.y = y;
  prm.Nr = b->max_Nr; prm.Ns = b->max_Ns; prm.ncs = best; prm.K = K;
  prm.per_up = (K + best - 1) / best;
  prm.per_dn = (Nsp - K + best - 1) / best;
  k_sweep<P, R><<<(unsigned)(b->nblocks * 2 * best), nthreads, sm, ctx->stream>>>(prm);
  cudaError_t e1 = cudaGetLastError();
  if (e1 != cudaSuccess) {
    ctx->err = std::string("k_sweep: ") + cudaGetErrorString(e1);
    return HSBP_ERR_CUDA;
  }
  return HSBP_OK;
}

template <int P> static int vol_sweep(hsbp_blocks *b, const double *u, double *y) {
  hsbp_ctx *ctx = b->ctx;
  if (((uintptr_t)u & 15) || ((uintptr_t)y & 15)) {
    ctx->err = "hsbp_apply: u / y must be 16-byte aligned for the line-marching kernel";
    return HSBP_ERR_ARG;
  }
  return sweep_points_per_thread(b) == 4 ? sweep_launch<P, 4>(b, u, y) : sweep_launch<P, 2>(b, u, y);
}

}  // namespace hsbp
