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
//   r-ends   the MCX closure rows of M(crr) u (with the face terms of faces 1, 2 already added) and the BM
//            closure rows of Qr u are prepared per (line, end) by k_edge_prep into a small table; the row of
//            each staged line rides along in the TMA ring and the edge lanes pick their values from it.
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
#include "k_generic.cuh"
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
// the same with shared-window addresses that were converted once
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s_s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared-memory access through 32-bit shared-window addresses
__device__ __forceinline__ double2 lds128(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double lds64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
// hide how a loop invariant was computed: the compiler then keeps it in a register instead of re-deriving it
__device__ __forceinline__ uint32_t opaque(uint32_t x) { asm volatile("" : "+r"(x)); return x; }
__device__ __forceinline__ int opaque(int x) { asm volatile("" : "+r"(x)); return x; }
__device__ __forceinline__ int64_t opaque(int64_t x) { asm volatile("" : "+l"(x)); return x; }
__device__ __forceinline__ double opaque(double x) { asm volatile("" : "+d"(x)); return x; }

// one lane of a fully active warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}
// 8-byte asynchronous copy global -> shared and its completion as one (pre-counted) arrival on an mbarrier
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

#ifndef SW_NST_OVERRIDE
#define SW_NST_OVERRIDE 3
#endif
constexpr int SW_NW = 2;                  // w buffers
constexpr int SW_NST = SW_NST_OVERRIDE;   // ring stages of u and crr: one being consumed, the others in flight
constexpr int SW_MAX_THREADS = 256;

struct SweepParams {
  // crr, css: the coefficient fields pre-multiplied by the norm weights of the *other* direction,
  //   crr'(i,j) = crr(i,j) * Hs[j] / hr,  css'(i,j) = css(i,j) * Hr[i] / hs   (global_curved.jl:261-268, 313-322);
  // every stiffness coupling is linear in its coefficient, so the kernel needs no scaling at all.
  const double *crr, *css, *crs, *u;
  // face terms of M-tilde, prepared per face point by k_face_prep (k_generic.cuh), block-face layout:
  //   y(point at normal offset m from face point n) += BS[m] * fcn[n] + (m == 0) * fgm[n];   null: volume part only
  const double *fcn, *fgm;
  // r-end table made by k_edge_prep: [block][line][clr], see SweepCfg::clr
  const double *rtab;
  double *y;
  int Nr, Ns;         // uniform block size
  int pitch;          // ODD kernels: distance between lines of the volume fields (Nr + 2, even); the pad entry of every line is 0
  int e0;             // first block of this launch (a launch may cover a range of the blocks)
  int ncs;            // chunks per side (a block is 2*ncs CTAs)
  int K;              // lines [0, K) are marched upwards, lines [K, Ns] downwards
  int per_up, per_dn; // output lines per chunk on either side
  // optional: blocks whose flag active[e * active_stride] is 0 are skipped (converged blocks of a batched PCG; y is not written)
  const int *active;
  int active_stride;
  // optional: dot[e * 2 ncs + chunk] = sum of u * y over the output lines of the chunk that are final when they are written
  // (all but the first BM lines of either s-end, which collect read-modify-write contributions): the u . M-tilde u of a PCG
  // step without a second pass over the vectors; summed in a fixed order
  double *dot;
};

template <int P> struct SweepCfg {
  using T = SweepTab<P>;
  static constexpr int H = T::H, W = 2 * T::H + 1, LB = 2 * T::H;
  static constexpr int PAD = (T::H + 1) & ~1;               // halo of a shared line, even (16-byte vectors)
  static constexpr int NB = (P == 2 ? 3 : (P == 4 ? 4 : 5)); // points of the boundary derivative BS (diagonal_sbp.jl:507,511,591)
  static constexpr int MCX = T::MC > NB ? T::MC : NB;       // rows of an r-end that take table values
  // r-end table of one line (k_edge_prep), laid out so that the lanes that own the end points pick their values up
  // with 16-byte loads in point order:
  //   [ near: rr(0 .. MCXP-1) | near: qr(0 .. MCXP-1) | far: rr(.. 1 0) | far: qr(.. 1 0) | 0 0 ]
  // rr(m) = row m of Hs/hr M(crr) u  +  face terms of faces 1, 2  +  row m of Qr^T w, w = crs o (Qs u)   (replaces the lane's value)
  // qr(m) = row m of Qr u (closure rows m < BM and the interior rows up to MCX, so that one mask serves both)
  // rows MCX .. MCXP-1 and the last pair are zero (a lane whose pair sticks out adds them).
  // ODD (lines with an odd number of points, stored with an even pitch: the last lane owns the last point and a phantom):
  // the far sections are two entries longer and the rows sit one position lower, so that the pairs stay aligned.
  static constexpr int MCXP = (MCX + 1) & ~1;
  template <bool ODD> static constexpr int lf() { return MCXP + (ODD ? 2 : 0); }          // length of a far section
  template <bool ODD> static constexpr int clr() { return 2 * MCXP + 2 * lf<ODD>() + 2; } // one line: clr * 8 bytes is a multiple of 16
  template <bool ODD> static constexpr int farpos(int m) { return lf<ODD>() - 1 - (ODD ? 1 : 0) - m; }   // row m inside a far section
  static constexpr int NKX = T::NK >= NB ? T::NK : ((NB + 1) & ~1);   // points normal to an r-face that k_edge_prep looks at
  static constexpr int RIMW = NKX + T::BN + 1;               // rim table entries per (line, end): crr' (NKX), crs (BN), tau Hf
  // ring depths (lines) of the four fields.  u and crr are only looked at on the newest line.  DEEP: css and crs
  // stay in shared memory for as long as the s-direction stencil needs them (css: lines j-2H+1 .. j for the
  // couplings of the pairs (j-H, j-H+O); crs: line j-H for w) instead of travelling through register windows
  template <bool DEEP> static constexpr int nsb() { return DEEP ? 2 * T::H + (SW_NST - 1) : SW_NST; }
  template <bool DEEP> static constexpr int nsc() { return DEEP ? T::H + 1 + (SW_NST - 1) : SW_NST; }
  template <bool DEEP> static constexpr int nlines() { return 2 * SW_NST + nsb<DEEP>() + nsc<DEEP>() + SW_NW; }
  __device__ static const double *bs() { return Sbp<P>::bs(); }
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
// s-end closure row `row` of M(css) u for one column, straight from memory (pb / pu: the column's point on
// marching line 0, lstride: distance between marching lines)
template <int P>
__device__ __noinline__ double sweep_sclosure_row(int row, const double *__restrict__ pb, const double *__restrict__ pu,
                                                  int64_t lstride, int64_t ulstride) {
  using T = SweepTab<P>;
  double b[T::NK], uu[T::NK];
#pragma unroll
  for (int k = 0; k < T::NK; ++k) { b[k] = __ldg(pb + k * lstride); uu[k] = __ldg(pu + k * ulstride); }
  return d2_closure_row<P>(row, b, uu);
}

// compile-time loop over lane numbers T0 .. T1-1
template <int T0, int T1, class F> __device__ __forceinline__ void for_lanes(F &&f) {
  if constexpr (T0 < T1) {
    f(std::integral_constant<int, T0>{});
    for_lanes<T0 + 1, T1>(f);
  }
}

// NT: upper bound of the CTA size; MINB: CTAs per SM the register allocation is sized for.
// DEEP: the steady state reads css / crs of older lines from deeper shared-memory rings instead of register
// windows (SweepCfg::nsb, nsc): 2H+1 lines of u and 2H+1 accumulators remain as per-point register state.
template <int P, int R, bool DEEP, bool ODD = false, bool DOT = false>
__device__ __forceinline__ void sweep_body(const SweepParams &prm) {
  using T = SweepTab<P>;
  using C = SweepCfg<P>;
  constexpr int H = C::H, W = C::W, PAD = C::PAD, MCXP = C::MCXP, NST = SW_NST;
  constexpr int LF = C::template lf<ODD>(), CLR = C::template clr<ODD>(), SODD = ODD ? 1 : 0;
  static_assert(!ODD || R == 2, "odd line lengths: two points per thread");
  constexpr int NSB = C::template nsb<DEEP>(), NSC = C::template nsc<DEEP>(), NLINES = C::template nlines<DEEP>();
  constexpr int MC = T::MC, BM = T::BM, BN = T::BN, MCX = C::MCX, NB = C::NB;
  constexpr int NV = R + 2 * PAD;                         // values of a line a thread looks at
  constexpr int DOFF = PAD;
  static_assert(R % 2 == 0 && PAD >= H, "layout");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthreads = blockDim.x;
  // Nrp: length of a stored line (ODD: the pitch, one phantom point behind the last one), Nrt: points of a line
  const int Nr = prm.Nr, Ns = prm.Ns, Nrt = Nr + 1, Nrp = ODD ? prm.pitch : Nrt, Nsp = Ns + 1;
  // A "slot" holds one line of one field: PAD doubles of zeros (left halo), the Nrp values, PAD zeros (right halo)
  const int LW = Nrp + 2 * PAD;
  double *ring_u = reinterpret_cast<double *>(smem_raw);  // [NST][LW]
  double *ring_rr = ring_u + (size_t)NST * LW;            // [NST][LW]  crr'
  double *ring_ss = ring_rr + (size_t)NST * LW;           // [NSB][LW]  css'
  double *ring_rs = ring_ss + (size_t)NSB * LW;           // [NSC][LW]  crs
  double *wbuf = ring_rs + (size_t)NSC * LW;              // [SW_NW][LW]
  double *clring = wbuf + SW_NW * LW;                     // [NST][CLR]: r-end table rows of the staged lines
  uint64_t *full = reinterpret_cast<uint64_t *>(clring + (size_t)NST * CLR);   // [NST] + 1 (split CTA barrier)

  // ---- which chunk --------------------------------------------------------------------------
  const int nch = 2 * prm.ncs;
  const int64_t el = blockIdx.x / nch;
  const int c = (int)(blockIdx.x - el * nch);
  const int64_t e = prm.e0 + el;
  if (prm.active != nullptr && prm.active[e * prm.active_stride] == 0) return;
  const bool up = c < prm.ncs;
  const int cc = up ? c : c - prm.ncs;
  const int nside = up ? prm.K : Nsp - prm.K;
  const int per = up ? prm.per_up : prm.per_dn;
  const int o0 = cc * per, o1 = min(nside, o0 + per);     // output lines [o0, o1), marching coordinates
  if (o0 >= o1) {
    if (DOT && tid == 0) prm.dot[e * nch + c] = 0.0;
    return;
  }
  const bool prologue = (o0 == 0);
  const int jstart = prologue ? 0 : o0 - H, jend = o1 - 1 + H;
  const int nlines = jend - jstart + 1;
  const double sig = up ? 1.0 : -1.0;
  const int64_t lstride = up ? (int64_t)Nrp : -(int64_t)Nrp;
  const int64_t base = e * (int64_t)Nrp * Nsp + (up ? 0 : (int64_t)Ns * Nrp);   // marching line j starts at base + j*lstride
  const uint32_t line_bytes = (uint32_t)Nrp * 8u;
  const int64_t foff = e * (2 * (int64_t)Nrt + 2 * (int64_t)Nsp);    // block e in the block-face layout

  // ---- one-time setup: zero the halos of the shared lines, barriers --------------------------
  for (int idx = tid; idx < NLINES * 2 * PAD; idx += nthreads) {
    const int line = idx / (2 * PAD), k = idx - line * (2 * PAD);
    ring_u[(size_t)line * LW + (k < PAD ? k : Nrp + k)] = 0.0;
  }
  if constexpr (ODD) {            // the phantom entry of the u slots: no copy ever writes it
    if (tid < NST) ring_u[(size_t)tid * LW + PAD + Nrt] = 0.0;
  }
  if constexpr (DEEP) {
    // lines in front of the first staged one are looked at by the first steps of a chunk (their pairs only
    // touch rows that are not output); keep them finite
    for (int idx = tid; idx < (NSB + NSC) * LW; idx += nthreads) ring_ss[idx] = 0.0;
  }
  if (tid == 0) {
    // ODD: u is not bulk-copied (its lines are not 16-byte aligned) -- every owning thread copies its own points with 8-byte
    // cp.async and arrives on the line's barrier when they have landed
    for (int s = 0; s < NST; ++s) mbar_init(&full[s], ODD ? 1u + (uint32_t)(Nrp / R) : 1u);
    fence_mbar_init();
  }
  __syncthreads();
  // programmatic dependent launch: this grid may start while k_edge_prep (which writes the r-end table and the face terms) is
  // still draining; everything above overlaps with its tail, nothing below may run before it has completed
  asm volatile("griddepcontrol.wait;" ::: "memory");
  auto issue = [&](int n) {      // marching line jstart + n into its slots  (thread 0 only)
    const int st = n % NST;
    const int64_t g = base + (int64_t)(jstart + n) * lstride;
    mbar_expect_tx(&full[st], (ODD ? 3u : 4u) * line_bytes + (uint32_t)(CLR * 8));
    if constexpr (!ODD) bulk_g2s(ring_u + (size_t)st * LW + DOFF, prm.u + g, line_bytes, &full[st]);
    bulk_g2s(ring_rr + (size_t)st * LW + DOFF, prm.crr + g, line_bytes, &full[st]);
    bulk_g2s(ring_ss + (size_t)(n % NSB) * LW + DOFF, prm.css + g, line_bytes, &full[st]);
    bulk_g2s(ring_rs + (size_t)(n % NSC) * LW + DOFF, prm.crs + g, line_bytes, &full[st]);
    const int64_t jl = up ? (int64_t)(jstart + n) : (int64_t)Ns - (jstart + n);       // actual line index
    bulk_g2s(clring + (size_t)st * CLR, prm.rtab + ((e * Nsp + jl) * CLR), (uint32_t)(CLR * 8), &full[st]);
  };
  if (tid == 0) {
    fence_proxy_async();
    for (int n = 0; n < NST && n < nlines; ++n) issue(n);
  }

  const int i0 = tid * R;
  const bool own = i0 < Nrp;
  const int nown = Nrp / R;                               // threads that own points (Nrp % R == 0)
  const double *qc = C::Qc();
  // u and y of ODD kernels: the caller's vectors (lines Nrt apart); everything else is read from pitched copies
  const int64_t ylstride = ODD ? (up ? (int64_t)Nrt : -(int64_t)Nrt) : lstride;
  const int64_t ybase = ODD ? e * (int64_t)Nrt * Nsp + (up ? 0 : (int64_t)Ns * Nrt) : base;
  const int nq = ODD ? min(R, Nrt - i0) : R;              // points of this thread that exist (own threads: >= 1)
  const double *gu = prm.u + ybase + i0;                  // + j*ylstride: this thread's points on marching line j
  const double *gss = prm.css + base + i0;
  auto issue_u = [&](int n) {    // ODD: this thread's points of marching line jstart + n
    if constexpr (ODD) {
      if (own) {
        const int st = n % NST;
        const double *src = gu + (int64_t)(jstart + n) * ylstride;
        const uint32_t dst = smem_u32(ring_u + (size_t)st * LW + DOFF + i0);
        cp_async8(dst, src);
        if (nq > 1) cp_async8(dst + 8u, src + 1);
        cp_async_arrive_noinc(smem_u32(&full[st]));
      }
    }
  };
  for (int n = 0; n < NST && n < nlines; ++n) issue_u(n);
  double *gy = prm.y + ybase + i0;                         // (ODD: 8-byte stores, the phantom point is left out)

  // s-direction windows.  Logical index k <-> marching line j-(W-1)+k (u, scaled css, crs) or j-H+k
  // (accumulators); the physical slot of logical k in a step with rotation PH is (PH+1+k) % W, so the
  // W-fold unrolled steady-state loop never moves a register.  bw uses k >= 1, cw uses k >= H.
  // (DEEP: no bw / cw, css and crs of older lines are read from the rings.)
  double uw[W][R], acc[W][R];
  [[maybe_unused]] double bw[DEEP ? 1 : W][R], cw[DEEP ? 1 : W][R];
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int k = 0; k < W; ++k) {
      uw[k][q] = 0.0; acc[k][q] = 0.0;
      if constexpr (!DEEP) { bw[k][q] = 0.0; cw[k][q] = 0.0; }
    }
  if (prologue && own) {                                  // lines that collect read-modify-write contributions
    for (int l = 0; l < BM; ++l)
#pragma unroll
      for (int q = 0; q < R; ++q)
        if (q < nq) gy[l * ylstride + q] = 0.0;
  }

  int j = jstart, n = 0, st = 0;
  uint32_t parity = 0;
  double dotp = 0.0;                                      // this thread's part of u . y (SweepParams::dot)
  auto finish = [&]() {                                   // chunk sum of u . y in a fixed order (every thread of the CTA gets here)
    if constexpr (!DOT) return;
    double v = dotp;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                                      // the shared lines are no longer needed
    if ((tid & 31) == 0) wbuf[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      double sum = 0.0;
      for (int w = 0; w < (nthreads >> 5); ++w) sum += wbuf[w];
      prm.dot[e * nch + c] = sum;
    }
  };

  // one marching step: line j arrives, line j-H is completed.  PH: rotation of the register windows
  // (compile time); FAST: steady state, no s-end closure logic.
  auto step = [&](auto PHc, auto FASTc) {
    constexpr int PH = decltype(PHc)::value;
    constexpr bool FAST = decltype(FASTc)::value;
    constexpr auto SL = [](int k) constexpr { return (PH + 1 + k) % W; };
    const bool pro = FAST ? false : prologue;

    mbar_wait(&full[st], parity);
    const double *sbu = ring_u + (size_t)st * LW + DOFF + i0 - PAD;           // this thread's NV values of line j
    const double *sbr = ring_rr + (size_t)st * LW + DOFF + i0 - PAD;
    const double *sbs = ring_ss + (size_t)(n % NSB) * LW + DOFF + i0;         // its own R values
    const double *sbc = ring_rs + (size_t)(n % NSC) * LW + DOFF + i0;
    const int jo = j - H;
    const bool outp = (jo >= o0) && (jo < o1);
    double *wb = wbuf + (size_t)(n & 1) * LW;

    if (own) {
      if constexpr (!FAST) {                              // canonical slot order: shift by one line
#pragma unroll
        for (int q = 0; q < R; ++q)
#pragma unroll
          for (int k = 0; k < W - 1; ++k) {               // only the slots that are read later (bw: k >= 1, cw: k >= H)
            uw[k][q] = uw[k + 1][q]; acc[k][q] = acc[k + 1][q];
            if constexpr (!DEEP) {
              if (k >= 1) bw[k][q] = bw[k + 1][q];
              if (k >= H) cw[k][q] = cw[k + 1][q];
            }
          }
      }
      // ---- take in line j -------------------------------------------------------------------
      double U[NV], Bq[NV];
#pragma unroll
      for (int k = 0; k < NV / 2; ++k) {
        const double2 a = *reinterpret_cast<const double2 *>(sbu + 2 * k);
        const double2 b2 = *reinterpret_cast<const double2 *>(sbr + 2 * k);
        U[2 * k] = a.x; U[2 * k + 1] = a.y; Bq[2 * k] = b2.x; Bq[2 * k + 1] = b2.y;
      }
#pragma unroll
      for (int q = 0; q < R; ++q) { uw[SL(W - 1)][q] = U[PAD + q]; acc[SL(W - 1)][q] = 0.0; }
      double cnew[R];                                     // crs on line j
#pragma unroll
      for (int k = 0; k < R / 2; ++k) {
        const double2 b2 = *reinterpret_cast<const double2 *>(sbc + 2 * k);
        cnew[2 * k] = b2.x; cnew[2 * k + 1] = b2.y;
        if constexpr (!DEEP) {
          const double2 a = *reinterpret_cast<const double2 *>(sbs + 2 * k);
          bw[SL(W - 1)][2 * k] = a.x; bw[SL(W - 1)][2 * k + 1] = a.y;
          cw[SL(W - 1)][2 * k] = b2.x; cw[SL(W - 1)][2 * k + 1] = b2.y;
        }
      }
      // css' on line j-H+s / crs on line j-H of point q: register windows, or (DEEP) the shared-memory rings
      auto css_at = [&](int s2, int q) -> double {
        if constexpr (DEEP) {
          int sl = (n - H + s2) % NSB;
          if (sl < 0) sl += NSB;
          return ring_ss[(size_t)sl * LW + DOFF + i0 + q];
        } else {
          return bw[SL(H + s2)][q];
        }
      };
      auto crs_lag = [&](int q) -> double {
        if constexpr (DEEP) {
          int sl = (n - H) % NSC;
          if (sl < 0) sl += NSC;
          return ring_rs[(size_t)sl * LW + DOFF + i0 + q];
        } else {
          return cw[SL(H)][q];
        }
      };

      // ---- r-direction on line j: rr = M(crr') u (pair form) goes straight into the accumulator of
      //      line j, qr = Qr u ----------------------------------------------------------------------
      double rr[R], qr[R];
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
          if constexpr (O == 1) {
            rr[q] = f[q + O] - f[q];
            qr[q] = C::template D<O>() * (U[PAD + q + O] - U[PAD + q - O]);
          } else {
            rr[q] += f[q + O] - f[q];
            qr[q] = fma(C::template D<O>(), U[PAD + q + O] - U[PAD + q - O], qr[q]);
          }
        }
      });
      // rows at the r-ends come from the table (SweepCfg::CLR: M(crr') u with the face terms and Qr^T w, and Qr u)
      {
        const double *cl = clring + (size_t)st * CLR;
#pragma unroll
        for (int q = 0; q < R; ++q) {
          const int i = i0 + q, m = Nr - i;
          if (i < MCX) { rr[q] = cl[i]; qr[q] = cl[MCXP + i]; }
          else if (m >= 0 && m < MCX) { rr[q] = cl[2 * MCXP + LF - 1 - SODD - m]; qr[q] = cl[2 * MCXP + LF + LF - 1 - SODD - m]; }
        }
      }
      double t[R];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        acc[SL(H)][q] += rr[q];
        t[q] = cnew[q] * qr[q];                           // t = crs o (Qr u) on line j
      }
      // ---- Qs^T t, pushed from row j: (Qs^T t)(l) += Qs[j][l] t(j) ------------------------------
      if (FAST || j >= BM) {
        for_offsets<1, H>([&](auto Oc) {
          constexpr int O = decltype(Oc)::value;
          const double d = sig * C::template D<O>();
#pragma unroll
          for (int q = 0; q < R; ++q) {
            acc[SL(H + O)][q] = fma(d, t[q], acc[SL(H + O)][q]);
            acc[SL(H - O)][q] = fma(-d, t[q], acc[SL(H - O)][q]);
          }
        });
      } else {
        for (int l = 0; l < BM; ++l) {                    // dense closure block: straight into y
          const double d = sig * qc[j * BN + l];
          if (d != 0.0) {
            double *yl = gy + l * ylstride;
#pragma unroll
            for (int q = 0; q < R; ++q)
              if (q < nq) yl[q] = fma(d, t[q], yl[q]);
          }
        }
        for_offsets<1, H>([&](auto Oc) {
          constexpr int O = decltype(Oc)::value;
          const int l = j + O;
          if (l >= BM && l < BN) {
            const double d = sig * qc[j * BN + l];
#pragma unroll
            for (int q = 0; q < R; ++q) acc[SL(H + O)][q] = fma(d, t[q], acc[SL(H + O)][q]);
          }
        });
      }
      // ---- s-direction stiffness: pairs (a, a+O), a = j-H; rows a <-> acc[0], a+O <-> acc[O] ------
      {
        const int a = j - H;
        const bool row_a_closure = pro && (a < MC);
        for_offsets<1, H>([&](auto Oc) {
          constexpr int O = decltype(Oc)::value;
          if (!(pro && (a + O < MC))) {
#pragma unroll
            for (int q = 0; q < R; ++q) {
              const double cf = pair_coef<P, O>([&](int s) { return css_at(s, q); });        // line a+s = j-H+s
              const double f = cf * (uw[SL(H + O)][q] - uw[SL(H)][q]);
              if (!row_a_closure) acc[SL(0)][q] += f;
              acc[SL(O)][q] -= f;
            }
          }
        });
      }
      // ---- w = crs o (Qs u) on the output line, shared with the r-neighbours ------------------
      if (outp) {
        double qs[R];
        if (FAST || jo >= BM) {
          for_offsets<1, H>([&](auto Oc) {
            constexpr int O = decltype(Oc)::value;
            const double d = sig * C::template D<O>();
#pragma unroll
            for (int q = 0; q < R; ++q) {
              if constexpr (O == 1) qs[q] = d * (uw[SL(H + O)][q] - uw[SL(H - O)][q]);
              else qs[q] = fma(d, uw[SL(H + O)][q] - uw[SL(H - O)][q], qs[q]);
            }
          });
        } else {
#pragma unroll
          for (int q = 0; q < R; ++q) qs[q] = 0.0;
          for (int l = 0; l < BN; ++l) {
            const double d = sig * qc[jo * BN + l];
            if (d != 0.0) {
              const double *ul = gu + l * ylstride;
#pragma unroll
              for (int q = 0; q < R; ++q) qs[q] = fma(d, q < nq ? ul[q] : 0.0, qs[q]);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < R / 2; ++k)
          *reinterpret_cast<double2 *>(wb + DOFF + i0 + 2 * k) =
              make_double2(crs_lag(2 * k) * qs[2 * k], crs_lag(2 * k + 1) * qs[2 * k + 1]);
      }
    }
    __syncthreads();
    if (tid == 0 && n + NST < nlines) { fence_proxy_async(); issue(n + NST); }    // the slots of line j are free again
    if (n + NST < nlines) issue_u(n + NST);
    if (own && outp) {
      // ---- rs = Qr^T w, then the output line ------------------------------------------------
      double Wv[NV], val[R];
#pragma unroll
      for (int k = 0; k < NV / 2; ++k) {
        const double2 a = *reinterpret_cast<const double2 *>(wb + DOFF + i0 - PAD + 2 * k);
        Wv[2 * k] = a.x; Wv[2 * k + 1] = a.y;
      }
#pragma unroll
      for (int q = 0; q < R; ++q) val[q] = acc[SL(0)][q];
      for_offsets<1, H>([&](auto Oc) {
        constexpr int O = decltype(Oc)::value;
#pragma unroll
        for (int q = 0; q < R; ++q) val[q] = fma(-C::template D<O>(), Wv[PAD + q + O] - Wv[PAD + q - O], val[q]);
      });
#pragma unroll
      for (int q = 0; q < R; ++q)                         // closure rows of Qr^T at the two r-ends came with the r-end table
        if (i0 + q < BM || Nr - (i0 + q) < BM) val[q] = acc[SL(0)][q];
      double *yl = gy + jo * ylstride;
      if constexpr (!FAST) {
        if (pro && jo < MC) {                             // closure row jo of M(css) u, straight from memory
#pragma unroll
          for (int q = 0; q < R; ++q)
            if (q < nq) val[q] += sweep_sclosure_row<P>(jo, gss + q, gu + q, lstride, ylstride);
        }
      }
      if constexpr (!FAST) {
        if (pro && jo < NB && prm.fcn != nullptr) {       // s-face terms: face 3 (line 0 side) / 4 (line Ns side)
          const int64_t fi = foff + 2 * Nsp + (up ? 0 : Nrt) + i0;
          const double bsm = C::bs()[jo];
#pragma unroll
          for (int q = 0; q < R; ++q) {
            if (q < nq) {
              val[q] = fma(bsm, prm.fcn[fi + q], val[q]);
              if (jo == 0) val[q] += prm.fgm[fi + q];
            }
          }
        }
      }
      if (!FAST && pro && jo < BM) {
#pragma unroll
        for (int q = 0; q < R; ++q)
          if (q < nq) yl[q] += val[q];
      } else {
#pragma unroll
        for (int q = 0; q < R; ++q)
          if (DOT && q < nq) dotp = fma(uw[SL(H)][q], val[q], dotp);
        if constexpr (ODD) {
#pragma unroll
          for (int q = 0; q < R; ++q)
            if (q < nq) yl[q] = val[q];
        } else {
#pragma unroll
          for (int k = 0; k < R / 2; ++k)
            *reinterpret_cast<double2 *>(yl + 2 * k) = make_double2(val[2 * k], val[2 * k + 1]);
        }
      }
    }
    ++j; ++n;
    if (++st == NST) { st = 0; parity ^= 1u; }
  };

  // ---- steady state --------------------------------------------------------------------------
  // Same step without any s-end closure logic, arranged so that the instruction footprint stays
  // inside the instruction cache: everything that does not touch the register windows (waiting for
  // the line, the r-direction work, the w exchange and the output) exists once; only the short
  // window section is specialised for the W rotations and selected by a uniform branch.
  // All shared-memory traffic of the loop goes through explicit 32-bit shared-window addresses that are advanced
  // incrementally, and the loop invariants are made opaque to the compiler: left alone it re-derives them from the
  // kernel parameters in every step (about a quarter of the executed instructions, measured with ncu).
  using GenericPH = std::integral_constant<int, W - 1>;    // rotation W-1: slot(k) == k (canonical order)
  const int jfast = prologue ? MCX + H : jstart;           // first step without s-end closure / face logic
  while (j < jfast && j <= jend) step(GenericPH{}, std::false_type{});
  if (j > jend) { finish(); return; }

  constexpr int NAS = DEEP ? 2 * H : 1, NAC = DEEP ? H + 1 : 1;
  constexpr int ELANES = (MCX + R - 1) / R;                // lanes per r-end that patch closure rows
  const uint32_t LWB = opaque((uint32_t)LW * 8u);
  const uint32_t full_s = opaque(smem_u32(full));
  // thread-constant addresses: this thread's NV values (u, crr, w) / its own R values (css, crs) in slot 0 of a ring
  const uint32_t c_u0 = opaque(smem_u32(ring_u) + (uint32_t)(DOFF + i0 - PAD) * 8u);
  const uint32_t c_drr = opaque((uint32_t)NST * LWB);                                    // ring_rr - ring_u
  const uint32_t c_ss0 = opaque(smem_u32(ring_ss) + (uint32_t)(DOFF + i0) * 8u), c_ssE = opaque(c_ss0 + (uint32_t)NSB * LWB);
  const uint32_t c_rs0 = opaque(smem_u32(ring_rs) + (uint32_t)(DOFF + i0) * 8u), c_rsE = opaque(c_rs0 + (uint32_t)NSC * LWB);
  const uint32_t c_w0 = opaque(smem_u32(wbuf) + (uint32_t)(DOFF + i0 - PAD) * 8u);
  const uint32_t c_wsum = opaque(2u * c_w0 + LWB);
  const uint32_t c_cl0 = opaque(smem_u32(clring));
  const int nout0 = opaque(o0 + H - jstart);               // first step whose line j-H belongs to the chunk
  const int nrefill = opaque(nlines - NST);                // steps after which no line is left to fetch
  const int nwarps = opaque(nthreads >> 5);
  const int mywarp = opaque(tid >> 5);
  constexpr int ELF = (MCX + SODD + R - 1) / R;            // lanes at the far end (ODD: the phantom point shifts the rows by one)
  const int edge = opaque((own && (tid < ELANES || tid >= nown - ELF)) ? 1 : 0);
  // edge lanes: byte offsets of their pairs in a table row (SweepCfg::CLR) and the blend factors; closm: points that are closure
  // rows of Qr^T (their accumulator already holds the whole row)
  uint32_t eo_rr[R / 2], eo_qr[R / 2];
  double keep[R];
  int closm_ = 0;
  {
    const bool nearl = tid < ELANES;
    const int d = nearl ? tid : nown - 1 - tid;            // distance of the lane from its end, in lanes
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      const int pos = nearl ? d * R + 2 * k : LF - (d + 1) * R + 2 * k;        // first entry of the pair inside its section
      const bool ok = pos >= 0 && pos < (nearl ? MCXP : LF);
      eo_rr[k] = opaque(8u * (uint32_t)(ok ? (nearl ? pos : 2 * MCXP + pos) : 2 * MCXP + 2 * LF));
      eo_qr[k] = opaque(8u * (uint32_t)(ok ? (nearl ? MCXP + pos : 2 * MCXP + LF + pos) : 2 * MCXP + 2 * LF));
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int row = nearl ? d * R + q : (d + 1) * R - 1 - SODD - q;          // row of the closure this point is (-1: phantom)
      keep[q] = opaque((edge && row >= 0 && row < MCX) ? 0.0 : 1.0);
      if (edge && row >= 0 && row < BM) closm_ |= 1 << q;
    }
  }
  const int closm = opaque(closm_);
  const int ownf = opaque(own ? 1 : 0);
  const double sgn = opaque(sig);                          // marching direction (the compiler would re-derive it from blockIdx)
  const int64_t lstr = opaque(lstride);
  [[maybe_unused]] const int64_t ylstr = opaque(ylstride);
  // ODD: u of marching line j + NST relative to the output line j - H of y (both the caller's vectors, same layout)
  [[maybe_unused]] const int64_t udelta = opaque((int64_t)(prm.u - prm.y) + (int64_t)(NST + H) * ylstride);
  [[maybe_unused]] const int nq2 = opaque(nq > 1 ? 1 : 0);
  const int64_t g0 = opaque(base + (int64_t)jstart * lstride);          // marching line 0 of this chunk in the volume fields
  const int64_t t0 = opaque((e * Nsp + (up ? (int64_t)jstart : (int64_t)Ns - jstart)) * CLR);   // ... in the r-end table
  const int64_t tstr = opaque(up ? (int64_t)CLR : -(int64_t)CLR);

  uint32_t au = c_u0 + (uint32_t)st * LWB, acl = c_cl0 + (uint32_t)(st * CLR) * 8u, abar = full_s + 8u * (uint32_t)st;
  uint32_t aw = c_w0 + (uint32_t)(n & 1) * LWB;
  uint32_t as_[NAS], ac_[NAC];                             // css' on lines j, j-1, ..; crs on lines j, .., j-H
#pragma unroll
  for (int k = 0; k < NAS; ++k) as_[k] = c_ss0 + (uint32_t)((((n - k) % NSB) + NSB) % NSB) * LWB;
#pragma unroll
  for (int k = 0; k < NAC; ++k) ac_[k] = c_rs0 + (uint32_t)((((n - k) % NSC) + NSC) % NSC) * LWB;
  int ph = 0;                                              // rotation of the next step
  int rw = 0;                                              // warp that issues the next refill
  double *yout = gy + (int64_t)(j - H) * ylstride;         // output line of the next step

  // ---- B: rs = Qr^T w and the output line (w buffer awb, accumulators accv, output line yo) --------------------
  auto outputB = [&](uint32_t awb, const double (&accv)[R], const double (&ucv)[R], double *yo) {
      // ---- B: rs = Qr^T w and the output line ---------------------------------------------------
    double Wv[NV], val[R];
#pragma unroll
    for (int k = 0; k < NV / 2; ++k) {
      const double2 a = lds128(awb + 16u * k);
      Wv[2 * k] = a.x; Wv[2 * k + 1] = a.y;
    }
#pragma unroll
    for (int q = 0; q < R; ++q) val[q] = accv[q];
    for_offsets<1, H>([&](auto Oc) {
      constexpr int O = decltype(Oc)::value;
#pragma unroll
      for (int q = 0; q < R; ++q) val[q] = fma(-C::template D<O>(), Wv[PAD + q + O] - Wv[PAD + q - O], val[q]);
    });
    if (closm) {                                          // closure rows of Qr^T at the two r-ends came with the r-end table
#pragma unroll
      for (int q = 0; q < R; ++q)
        if ((closm >> q) & 1) val[q] = accv[q];
    }
    if constexpr (DOT) {
      dotp = fma(ucv[0], val[0], dotp);
#pragma unroll
      for (int q = 1; q < R; ++q)
        if (!ODD || nq2) dotp = fma(ucv[q], val[q], dotp);
    }
    if constexpr (ODD) {
      yo[0] = val[0];
      if (nq2) yo[1] = val[1];
    } else {
#pragma unroll
      for (int k = 0; k < R / 2; ++k)
        *reinterpret_cast<double2 *>(yo + 2 * k) = make_double2(val[2 * k], val[2 * k + 1]);
    }
  };
#pragma unroll 1
  for (int left = nlines - n; left > 0; --left) {
    mbar_wait_s(abar, parity);
    const bool outp = n >= nout0;                          // line j-H is an output line of this chunk
    double accout[R], ucen[R];
    if (ownf) {
      // ---- A: line j from shared memory, r-direction work (rotation independent) ---------------
      double U[NV], Bq[NV], cn[R], rr[R], qr[R], wout[R];
#pragma unroll
      for (int k = 0; k < NV / 2; ++k) {
        const double2 a = lds128(au + 16u * k);
        const double2 b2 = lds128(au + c_drr + 16u * k);
        U[2 * k] = a.x; U[2 * k + 1] = a.y; Bq[2 * k] = b2.x; Bq[2 * k + 1] = b2.y;
      }
#pragma unroll
      for (int k = 0; k < R / 2; ++k) {
        const double2 b2 = lds128(ac_[0] + 16u * k);
        cn[2 * k] = b2.x; cn[2 * k + 1] = b2.y;
      }
      // DEEP: couplings of the s-direction pairs (a, a+O), a = j-H, and crs on line j-H, from the rings
      // (rotation independent); otherwise the newest css line for the register window
      [[maybe_unused]] double bn[R], cfs[H][R], c2[R];
      if constexpr (DEEP) {
        double Bs[2 * H][R];                              // css' on lines j-2H+1 .. j
#pragma unroll
        for (int m = 0; m < 2 * H; ++m)
#pragma unroll
          for (int k = 0; k < R / 2; ++k) {
            const double2 a = lds128(as_[2 * H - 1 - m] + 16u * k);
            Bs[m][2 * k] = a.x; Bs[m][2 * k + 1] = a.y;
          }
        for_offsets<1, H>([&](auto Oc) {
          constexpr int O = decltype(Oc)::value;
#pragma unroll
          for (int q = 0; q < R; ++q) cfs[O - 1][q] = pair_coef<P, O>([&](int s2) { return Bs[s2 + H - 1][q]; });
        });
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
          const double2 a = lds128(ac_[H] + 16u * k);
          c2[2 * k] = a.x; c2[2 * k + 1] = a.y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
          const double2 a = lds128(as_[0] + 16u * k);
          bn[2 * k] = a.x; bn[2 * k + 1] = a.y;
        }
      }
      for_offsets<1, H>([&](auto Oc) {
        constexpr int O = decltype(Oc)::value;
        double f[R + O];
#pragma unroll
        for (int k = 0; k < R + O; ++k) {                 // pair (a, a+O), a = i0 - O + k
          const int ia = PAD - O + k;
          const double cf = pair_coef<P, O>([&](int s2) { return Bq[ia + s2]; });
          f[k] = cf * (U[ia + O] - U[ia]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) {
          if constexpr (O == 1) {
            rr[q] = f[q + O] - f[q];
            qr[q] = C::template D<O>() * (U[PAD + q + O] - U[PAD + q - O]);
          } else {
            rr[q] += f[q + O] - f[q];
            qr[q] = fma(C::template D<O>(), U[PAD + q + O] - U[PAD + q - O], qr[q]);
          }
        }
      });
      // rows at the r-ends come from the table: one 16-byte load per pair of points for rr and for qr, blended with the
      // lane's own value through keep (0: row of the table, 1: interior point whose pair partner sticks out, table holds 0)
      if (edge) {
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
          const double2 tr = lds128(acl + eo_rr[k]), tq = lds128(acl + eo_qr[k]);
          rr[2 * k] = fma(rr[2 * k], keep[2 * k], tr.x); rr[2 * k + 1] = fma(rr[2 * k + 1], keep[2 * k + 1], tr.y);
          qr[2 * k] = fma(qr[2 * k], keep[2 * k], tq.x); qr[2 * k + 1] = fma(qr[2 * k + 1], keep[2 * k + 1], tq.y);
        }
      }
      // ---- S: the register windows, one specialisation per rotation -----------------------------
      auto windows = [&](auto PHc) {
        constexpr int PH = decltype(PHc)::value;
        constexpr auto SL = [](int k) constexpr { return (PH + 1 + k) % W; };
#pragma unroll
        for (int q = 0; q < R; ++q) {
          uw[SL(W - 1)][q] = U[PAD + q];
          if constexpr (!DEEP) { bw[SL(W - 1)][q] = bn[q]; cw[SL(W - 1)][q] = cn[q]; }
          acc[SL(H)][q] += rr[q];
        }
        for_offsets<1, H>([&](auto Oc) {                  // Qs^T t pushed from row j (row j+H is touched for the first time)
          constexpr int O = decltype(Oc)::value;
          const double d = sgn * C::template D<O>();
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const double t = cn[q] * qr[q];
            if constexpr (O == H) acc[SL(H + O)][q] = d * t;
            else acc[SL(H + O)][q] = fma(d, t, acc[SL(H + O)][q]);
            acc[SL(H - O)][q] = fma(-d, t, acc[SL(H - O)][q]);
          }
        });
        for_offsets<1, H>([&](auto Oc) {                  // s-direction stiffness pairs (a, a+O), a = j-H
          constexpr int O = decltype(Oc)::value;
#pragma unroll
          for (int q = 0; q < R; ++q) {
            double cf;
            if constexpr (DEEP) cf = cfs[O - 1][q];
            else cf = pair_coef<P, O>([&](int s2) { return bw[SL(H + s2)][q]; });
            const double f = cf * (uw[SL(H + O)][q] - uw[SL(H)][q]);
            acc[SL(0)][q] += f;
            acc[SL(O)][q] -= f;
          }
        });
        for_offsets<1, H>([&](auto Oc) {                  // w = crs o (Qs u) on line jo
          constexpr int O = decltype(Oc)::value;
          const double d = sgn * C::template D<O>();
#pragma unroll
          for (int q = 0; q < R; ++q) {
            if constexpr (O == 1) wout[q] = d * (uw[SL(H + O)][q] - uw[SL(H - O)][q]);
            else wout[q] = fma(d, uw[SL(H + O)][q] - uw[SL(H - O)][q], wout[q]);
          }
        });
#pragma unroll
        for (int q = 0; q < R; ++q) {
          if constexpr (DEEP) wout[q] *= c2[q];
          else wout[q] *= cw[SL(H)][q];
          accout[q] = acc[SL(0)][q];
          if constexpr (DOT) ucen[q] = uw[SL(H)][q];
        }
      };
      [&]<int... PHs>(std::integer_sequence<int, PHs...>) {
        ((ph == PHs ? (windows(std::integral_constant<int, PHs>{}), 0) : 0), ...);
      }(std::make_integer_sequence<int, W>{});
      if (outp) {
#pragma unroll
        for (int k = 0; k < R / 2; ++k) sts128(aw + 8u * PAD + 16u * k, wout[2 * k], wout[2 * k + 1]);
      }
    }
    __syncthreads();
    if (mywarp == rw && n < nrefill) {                     // the slots of line j are free again: line j + NST goes into them
                                                          // (the warps take turns, so no warp is the slow one)
      if (elect_one()) {
        const int nn = n + NST;
        const int64_t g = g0 + nn * lstr;
        const uint32_t tofs8 = (uint32_t)(DOFF + i0) * 8u;                 // thread offset inside c_ss0 / c_rs0
        uint32_t d_ss = as_[0] + (uint32_t)NST * LWB; if (d_ss >= c_ssE) d_ss -= (uint32_t)NSB * LWB;
        uint32_t d_rs = ac_[0] + (uint32_t)NST * LWB; if (d_rs >= c_rsE) d_rs -= (uint32_t)NSC * LWB;
        const uint32_t d_u = au + (uint32_t)(PAD - i0) * 8u;               // data start of the slot of line j in ring_u
        mbar_expect_tx_s(abar, (ODD ? 3u : 4u) * line_bytes + (uint32_t)(CLR * 8));
        if constexpr (!ODD) bulk_g2s_s(d_u, prm.u + g, line_bytes, abar);
        bulk_g2s_s(d_u + c_drr, prm.crr + g, line_bytes, abar);
        bulk_g2s_s(d_ss - tofs8 + (uint32_t)DOFF * 8u, prm.css + g, line_bytes, abar);
        bulk_g2s_s(d_rs - tofs8 + (uint32_t)DOFF * 8u, prm.crs + g, line_bytes, abar);
        bulk_g2s_s(acl, prm.rtab + t0 + nn * tstr, (uint32_t)(CLR * 8), abar);
      }
    }
    if constexpr (ODD) {                                   // this thread's points of u on line j + NST, into the slot of line j
      if (ownf && n < nrefill) {
        const double *src = yout + udelta;
        cp_async8(au + 8u * PAD, src);
        if (nq2) cp_async8(au + 8u * PAD + 8u, src + 1);
        cp_async_arrive_noinc(abar);
      }
    }
    if (ownf && outp) outputB(aw, accout, ucen, yout);
    yout += ODD ? ylstr : lstr;
    ++n;
    if (++rw == nwarps) rw = 0;
    au += LWB; acl += (uint32_t)(CLR * 8); abar += 8u;
    if (++st == NST) { st = 0; parity ^= 1u; au = c_u0; acl = c_cl0; abar = full_s; }
#pragma unroll
    for (int k = NAS - 1; k > 0; --k) as_[k] = as_[k - 1];
    as_[0] += LWB;
    if (as_[0] == c_ssE) as_[0] = c_ss0;
#pragma unroll
    for (int k = NAC - 1; k > 0; --k) ac_[k] = ac_[k - 1];
    ac_[0] += LWB;
    if (ac_[0] == c_rsE) ac_[0] = c_rs0;
    aw = c_wsum - aw;
    if (++ph == W) ph = 0;
  }
  finish();
}

// register windows: the register allocation is sized through the CTAs per SM (MINB)
template <int P, int R, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_sweep(const SweepParams prm) { sweep_body<P, R, false>(prm); }
// deep rings: explicit register cap (shared memory, not registers, bounds the CTAs per SM)
template <int P, int R, int MAXREG, bool ODD = false, bool DOT = false>
__global__ void __maxnreg__(MAXREG)
k_sweep_deep(const SweepParams prm) { sweep_body<P, R, true, ODD, DOT>(prm); }

// ---- edge preparation ---------------------------------------------------------------------------
// One CTA per (block, face), launched before k_sweep.  Everything that lives on the rim of a block and
// would otherwise need a second pass over y:
//   * the face terms of M-tilde (reference locoperator, global_curved.jl:444-458, 478-486): from a = L u,
//     g = G u and the face's boundary condition (k_generic.cuh header) the per-face-point quantities
//       fcn[n] = (Hf/hn) c_nn alpha,   fgm[n] = sgn (Q_t^T (c_x o alpha))_n + beta
//     with  y(point m deep behind face point n) += BS[m] fcn[n] + (m == 0) fgm[n];
//   * for the r-faces (k = 1, 2) the table row of every line n (layout: SweepCfg::clr): the first MCX rows of
//     Hs[n]/hr M(crr) u at that end with the face terms added, the closure rows of Qr^T w, w = crs o (Qs u) (the s-direction
//     derivative of the end points goes through shared memory), and the first MCX rows of Qr u (mirrored and sign-flipped
//     at the far end) -- everything the lanes of k_sweep that own the end points would otherwise have to branch for.
// with_faces = 0 leaves the face terms out (volume operator A-tilde only).
// dynamic shared memory: (2 + BN) * (max face points) doubles
template <int P, bool ODD = false>
#ifndef SW_EDGE_MINB
#define SW_EDGE_MINB 4      // CTAs per SM the register allocation is sized for (measured: 2 -> 0.076, 3 -> 0.068, 4 -> 0.064, 5 -> 0.071 ms)
#endif
#ifndef SW_EDGE_MINB6
#define SW_EDGE_MINB6 SW_EDGE_MINB
#endif
__global__ void __launch_bounds__(256, P == 6 ? SW_EDGE_MINB6 : SW_EDGE_MINB)
k_edge_prep(const BlockDesc *__restrict__ desc, const double *__restrict__ crr, const double *__restrict__ css,
            const double *__restrict__ crs, const double *__restrict__ tau, const double *__restrict__ u,
            double *__restrict__ fcn, double *__restrict__ fgm, double *__restrict__ rtab, int with_faces, int e0,
            const double *__restrict__ rim, const int *__restrict__ active, int active_stride, int uNr, int uNs) {
  using S = Sbp<P>;
  using T = SweepTab<P>;
  using C = SweepCfg<P>;
  // points normal to the face that are looked at: the closure rows need T::NK, the boundary derivative NB
  constexpr int NK = C::NKX, BN = T::BN, BM = T::BM, MCX = C::MCX, MCXP = C::MCXP, H = C::H;
  static_assert(NK % 2 == 0 && NK >= S::NB && NK >= T::NK && NK >= BN && NK >= MCX + H, "normal extent");
  static_assert(BM <= T::MC, "the closure rows of Qr^T are rows the table replaces");
  extern __shared__ double sm_face[];
  const int e = e0 + (blockIdx.x >> 2), k = blockIdx.x & 3;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // k_sweep's CTAs may take the SMs this grid leaves free
  if (active != nullptr && active[(int64_t)e * active_stride] == 0) return;       // block skipped by the caller (see SweepParams)
  // uniform blocks (the only ones the sweep path serves): sizes and offsets come with the launch, so no address waits for the
  // descriptor; only the face's boundary-condition code is read from it, and that is needed late
  BlockDesc d;
  d.Nr = uNr; d.Ns = uNs;
  d.voff = (int64_t)e * (uNr + 1) * (uNs + 1);
  d.foff = (int64_t)e * (2 * (int64_t)(uNr + 1) + 2 * (int64_t)(uNs + 1));
  const int bck = desc[e].bc[k];
  const FaceGeom fg = face_geom(d, k);
  double *sa = sm_face, *sx = sm_face + fg.nf, *su = sm_face + 2 * fg.nf;      // su[kk][n]: u at the BN end points of line n (r-faces)
  const int Nrp = d.Nr + 1, Nsp = d.Ns + 1;
  const int64_t up = Nrp, uoff = d.voff;                  // lines of u (the caller's vector)
  constexpr int LF = C::template lf<ODD>(), CLR = C::template clr<ODD>();
  // One face point per thread and trip; every global load of a trip is issued before its first use.
  // Faces longer than the CTA take several trips: the tangential operators need the whole face in shared
  // memory, so alpha-dependent data is parked in fcn / fgm between the passes (STAGE 0, 1, 2); a face that
  // fits does everything in one go (STAGE -1).
  const bool single = fg.nf <= (int)blockDim.x;
  for (int stage = single ? -1 : 0; stage < (single ? 0 : 3); ++stage) {
    for (int n0 = 0; n0 < fg.nf; n0 += blockDim.x) {
      const int n = n0 + threadIdx.x;
      const bool act = n < fg.nf;
      double b[NK], uu[NK];                    // c_nn and u at normal offsets 0 .. NK-1 behind face point n
      double cxf = 0.0, tauf = 0.0;
      [[maybe_unused]] double ce[BN];          // crs at the BN end points of line n (r-faces)
      const int64_t fi = d.foff + fg.fstart + n;
      if (act) {
        if (k < 2) {                           // r-faces: u is strided in memory (NK contiguous points per line, 16-byte
                                               // aligned); the static data comes from the rim table, coalesced in n
          const int64_t g0 = uoff + up * n + (k == 0 ? 0 : Nrp - NK);
          const double *pr = rim + (((int64_t)e * 2 + k) * C::RIMW) * Nsp + n;
          if constexpr (ODD) {                 // an odd number of points per line: no 16-byte alignment
#pragma unroll
            for (int m = 0; m < NK; ++m) uu[k == 0 ? m : NK - 1 - m] = u[g0 + m];
          } else {
            const double2 *pu = reinterpret_cast<const double2 *>(u + g0);
#pragma unroll
            for (int m = 0; m < NK / 2; ++m) {
              const double2 vu = pu[m];
              if (k == 0) { uu[2 * m] = vu.x; uu[2 * m + 1] = vu.y; }
              else { uu[NK - 1 - 2 * m] = vu.x; uu[NK - 2 - 2 * m] = vu.y; }
            }
          }
#pragma unroll
          for (int m = 0; m < NK; ++m) b[m] = pr[(int64_t)m * Nsp];       // already scaled by Hs[n] / hr
#pragma unroll
          for (int m = 0; m < BN; ++m) ce[m] = pr[(int64_t)(NK + m) * Nsp];
          cxf = ce[0];
          tauf = pr[(int64_t)(NK + BN) * Nsp];                             // tau * Hf
        } else {                               // s-faces: lines 0 .. NB-1 (or Ns .. Ns-NB+1), coalesced along the face
          const int64_t g0 = d.voff + n + (k == 2 ? 0 : (int64_t)Nrp * d.Ns);
          const int64_t gu0 = uoff + n + (k == 2 ? 0 : up * d.Ns);
          const int64_t ls = k == 2 ? up : -up;
#pragma unroll
          for (int m = 0; m < S::NB; ++m) uu[m] = u[gu0 + ls * m];
          b[0] = css[g0];
          cxf = crs[g0];
        }
        if (with_faces && k >= 2) tauf = tau[fi] * (fg.ht * hweight<P>(n, fg.Nt));
      }
      double cn = 0.0, beta = 0.0, alpha = 0.0;
      const double Hf = fg.ht * hweight<P>(act ? n : 0, fg.Nt);
      if (stage <= 0) {                        // restriction a = L u of the whole face; u at the end points of every line
        if (act) {
          sa[n] = uu[0];
          if (k < 2) {
#pragma unroll
            for (int m = 0; m < BN; ++m) su[m * fg.nf + n] = uu[m];
          }
        }
        if (stage == 0) continue;
        __syncthreads();
      }
      if (with_faces) {
        if (stage < 0 || stage == 1) {         // g = G u, then alpha / beta of the boundary condition
          if (act) {
            double bsu = S::bs()[0] * uu[0];
#pragma unroll
            for (int m = 1; m < S::NB; ++m) bsu += S::bs()[m] * uu[m];
            const double qt = q_apply<P>(n, fg.Nt, [&](int l) { return sa[l]; });
            cn = k < 2 ? b[0] : (Hf / fg.hn) * b[0];
            const double g = cn * bsu + fg.sgn * cxf * qt;
            const double tH = tauf;
            if (bck == HSBP_BC_NEUMANN) { alpha = -g / tH; beta = 0.0; }
            else                            { alpha = -uu[0];  beta = tH * uu[0] - g; }
            sx[n] = cxf * alpha;
            cn *= alpha;
            if (stage == 1) { fcn[fi] = cn; fgm[fi] = beta; }
          }
          if (stage == 1) continue;
          __syncthreads();
        }
        if (act) {                             // tangential part of G^T alpha lands on the face point itself
          if (stage == 2) { cn = fcn[fi]; beta = fgm[fi]; }
          beta += fg.sgn * qt_apply<P>(n, fg.Nt, [&](int l) { return sx[l]; });
          if (k >= 2 || stage == 2) { fcn[fi] = cn; fgm[fi] = beta; }   // (r-faces: only needed by the table below)
        }
      } else if (stage == 1) {
        continue;
      }
      if (k < 2 && act) {
        // r-end table row of line n
        double rows[MCX], qq[MCX];
#pragma unroll
        for (int m = 0; m < MCX; ++m) rows[m] = 0.0;
        d2_closure_rows<P>(b, uu, rows);
#pragma unroll
        for (int m = T::MC; m < MCX; ++m)       // (p = 2) interior rows that carry face terms
          rows[m] = m_apply<P>(m, 2 * NK, [&](int i) { return b[i]; }, [&](int i) { return uu[i]; });
        if (with_faces) {
#pragma unroll
          for (int m = 0; m < C::NB; ++m) rows[m] = fma(S::bs()[m], cn, rows[m]);
          rows[0] += beta;
        }
        {                                       // closure rows of Qr^T w, w = crs o (Qs u) at the BN end points of this line
          double w[BN];
#pragma unroll
          for (int m = 0; m < BN; ++m) w[m] = ce[m] * q_apply<P>(n, fg.Nt, [&](int l) { return su[m * fg.nf + l]; });
#pragma unroll
          for (int m = 0; m < BM; ++m) rows[m] += (k == 0 ? 1.0 : -1.0) * qt_closure_row<P>(m, w);
        }
        q_closure_rows<P>(uu, qq);
#pragma unroll
        for (int m = BM; m < MCX; ++m) {        // interior rows of Qr u
          double a = 0.0;
#pragma unroll
          for (int o = 1; o <= H; ++o) a = fma(S::d()[H + o], uu[m + o] - uu[m - o], a);
          qq[m] = a;
        }
        double *out = rtab + ((int64_t)e * Nsp + n) * CLR;
        if (k == 0) {
#pragma unroll
          for (int m = 0; m < MCXP; ++m) { out[m] = m < MCX ? rows[m] : 0.0; out[MCXP + m] = m < MCX ? qq[m] : 0.0; }
        } else {                                // far end: point order (row m at farpos(m)), Q changes sign
#pragma unroll
          for (int pos = 0; pos < LF; ++pos) {
            const int m = LF - 1 - (ODD ? 1 : 0) - pos;
            out[2 * MCXP + pos] = (m >= 0 && m < MCX) ? rows[m >= 0 && m < MCX ? m : 0] : 0.0;
            out[2 * MCXP + LF + pos] = (m >= 0 && m < MCX) ? -qq[m >= 0 && m < MCX ? m : 0] : 0.0;
          }
          out[2 * MCXP + 2 * LF] = 0.0; out[2 * MCXP + 2 * LF + 1] = 0.0;
        }
      }
    }
    __syncthreads();                           // stage boundary: sa / sx of the whole face are complete
  }
}

// ---- host side ------------------------------------------------------------------------------
// Nrp: stored line length (the even pitch of blocks with an odd number of points per line)
template <int P> static size_t sweep_smem(int Nrp, bool deep, bool odd = false) {
  using C = SweepCfg<P>;
  const int LW = Nrp + 2 * C::PAD;
  const int nl = deep ? C::template nlines<true>() : C::template nlines<false>();
  const int clr = odd ? C::template clr<true>() : C::template clr<false>();
  return (size_t)nl * LW * sizeof(double) + (size_t)SW_NST * clr * sizeof(double) + (SW_NST + 1) * sizeof(uint64_t);
}

static int sweep_points_per_thread(const hsbp_blocks *b) {
  const int Nrp = b->max_Nr + 1;
  if (Nrp & 1) return 2;                                      // odd line length: the pitched two-point variant
  if (b->sweep_r_override == 2 || (b->sweep_r_override == 4 && Nrp % 4 == 0)) return b->sweep_r_override;
  const bool can4 = Nrp % 4 == 0, can2 = ((Nrp / 2 + 31) & ~31) <= SW_MAX_THREADS;
  if (b->sweep_deep) {
    // measured on B200 at 256-point lines (deep rings): p = 4: R = 2 (126 registers, 16 warps/SM) 0.562 ms, R = 4
    // (195 registers, 10 warps/SM, two-way bank conflicts of the 32-byte-per-lane reads) 0.724 ms; p = 2: 0.42 / 0.47 ms;
    // p = 6: R = 2 at 168 registers (12 warps/SM, no spills) 0.98 ms, at 128 registers (a few spills) 1.13 ms, R = 4 1.20 ms
    return can2 ? 2 : 4;
  }
  // register windows: R = 4 (252 registers, 8 warps/SM) 0.65 ms, R = 2 (128 registers, 16 warps/SM, spilling) 0.67 ms
  if (b->p == 6) return 2;                                    // R = 4 spills heavily with the 7-line windows of p = 6
  return (can4 && Nrp >= 128) ? 4 : 2;
}

template <int P> static bool sweep_eligible(const hsbp_blocks *b) {
  if (!b->uniform) return false;
  const int Nrt = b->max_Nr + 1, Nsp = b->max_Ns + 1;
  // lines with an odd number of points (N = 34, 68, 136, 200 of the reference's drivers) run on pitched copies: the bulk copies
  // need 16-byte aligned lines; deep-ring kernels only
  const bool odd = Nrt & 1;
  if (odd && !b->sweep_deep) return false;
  const int Nrp = Nrt + (odd ? 1 : 0);
  if (Nrt < 32 || Nsp < 32) return false;                      // room for both closures and a chunk per side
  const int R = sweep_points_per_thread(b);
  const int nthreads = ((Nrp / R) + 31) & ~31;
  if (nthreads > SW_MAX_THREADS) return false;
  return sweep_smem<P>(Nrp, true, odd) <= b->ctx->smem_optin;
}

// crr' = crr * Hs[j] / hr, css' = css * Hr[i] / hs for uniform blocks (see SweepParams)
template <int P>
__global__ void __launch_bounds__(256)
k_sweep_scale(const double *__restrict__ crr, const double *__restrict__ css, const double *__restrict__ crs,
              double *__restrict__ crr_s, double *__restrict__ css_s, double *__restrict__ crs_p, int Nr, int Ns, int pitch,
              int64_t total) {
  // output lines are `pitch` apart (= Nr + 1, or Nr + 2 with a zero pad entry for odd line lengths; then crs is copied too)
  using C = SweepCfg<P>;
  constexpr int BM = SweepTab<P>::BM;
  const int Nrp = Nr + 1;
  const int64_t np = (int64_t)Nrp * (Ns + 1), npp = (int64_t)pitch * (Ns + 1);
  const double hr = 2.0 / Nr, hs = 2.0 / Ns;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / npp, loc = idx - e * npp;
    const int jj = (int)(loc / pitch), i = (int)(loc - (int64_t)jj * pitch);
    if (i >= Nrp) {
      crr_s[idx] = 0.0; css_s[idx] = 0.0;
      if (crs_p) crs_p[idx] = 0.0;
      continue;
    }
    const int64_t src = e * np + (int64_t)jj * Nrp + i;
    const double hwi = i < BM ? C::hw()[i] : (i > Nr - BM ? C::hw()[Nr - i] : 1.0);
    const double hwj = jj < BM ? C::hw()[jj] : (jj > Ns - BM ? C::hw()[Ns - jj] : 1.0);
    crr_s[idx] = crr[src] * (hs * hwj / hr);
    css_s[idx] = css[src] * (hr * hwi / hs);
    if (crs_p) crs_p[idx] = crs[src];
  }
}

// static data of the r-faces for k_edge_prep, laid out [block][end][entry][line] so that threads (= lines) read it
// coalesced: entries 0..NK-1 = Hs[n]/hr * crr at the NK points behind the face point, NK..NK+BN-1 = crs at the BN points
// behind it, NK+BN = tau * Hf
template <int P>
__global__ void __launch_bounds__(256)
k_rim_build(const BlockDesc *__restrict__ desc, const double *__restrict__ crr, const double *__restrict__ crs,
            const double *__restrict__ tau, double *__restrict__ rim) {
  using S = Sbp<P>;
  using T = SweepTab<P>;
  using C = SweepCfg<P>;
  constexpr int NK = C::NKX, BN = T::BN;
  const int e = blockIdx.x >> 1, k = blockIdx.x & 1;
  const BlockDesc d = desc[e];
  const FaceGeom fg = face_geom(d, k);
  const int Nrp = d.Nr + 1, Nsp = d.Ns + 1;
  for (int n = threadIdx.x; n < Nsp; n += blockDim.x) {
    const double Hf = fg.ht * hweight<P>(n, fg.Nt);
    const int64_t g0 = d.voff + (int64_t)Nrp * n;
    double *pr = rim + (((int64_t)e * 2 + k) * C::RIMW) * Nsp + n;
    for (int m = 0; m < NK; ++m) pr[(int64_t)m * Nsp] = (Hf / fg.hn) * crr[g0 + (k == 0 ? m : d.Nr - m)];
    for (int m = 0; m < BN; ++m) pr[(int64_t)(NK + m) * Nsp] = crs[g0 + (k == 0 ? m : d.Nr - m)];
    pr[(int64_t)(NK + BN) * Nsp] = tau[d.foff + fg.fstart + n] * Hf;
  }
}

template <int P> static int sweep_prepare(hsbp_blocks *b) {
  hsbp_ctx *ctx = b->ctx;
  if (b->sweep_scaled_valid && b->rim_valid) return HSBP_OK;
  const bool odd = (b->max_Nr + 1) & 1;
  const int pitch = b->max_Nr + 1 + (odd ? 1 : 0);
  const int64_t VNpp = b->nblocks * (int64_t)pitch * (b->max_Ns + 1);      // points of a pitched volume vector
  const size_t vb = (size_t)VNpp * sizeof(double);
  if (!b->d_crr_s) {
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_crr_s, vb));
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_css_s, vb));
    if (odd) {                                                // pitched copy of crs
      HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_crs_p, vb));
    }
    const int clr = odd ? SweepCfg<P>::template clr<true>() : SweepCfg<P>::template clr<false>();
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_rtab, (size_t)b->nblocks * (b->max_Ns + 1) * clr * sizeof(double)));
    HSBP_CUDA(ctx, cudaMalloc((void **)&b->d_rim, (size_t)b->nblocks * 2 * SweepCfg<P>::RIMW * (b->max_Ns + 1) * sizeof(double)));
  }
  k_sweep_scale<P><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(b->d_crr, b->d_css, b->d_crs, b->d_crr_s, b->d_css_s, b->d_crs_p,
                                                              b->max_Nr, b->max_Ns, pitch, VNpp);
  cudaError_t e1 = cudaGetLastError();
  if (e1 != cudaSuccess) {
    ctx->err = std::string("k_sweep_scale: ") + cudaGetErrorString(e1);
    return HSBP_ERR_CUDA;
  }
  k_rim_build<P><<<(unsigned)(2 * b->nblocks), 256, 0, ctx->stream>>>(b->d_desc, b->d_crr, b->d_crs, b->d_tau, b->d_rim);
  if ((e1 = cudaGetLastError()) != cudaSuccess) {
    ctx->err = std::string("k_rim_build: ") + cudaGetErrorString(e1);
    return HSBP_ERR_CUDA;
  }
  b->sweep_scaled_valid = true;
  b->rim_valid = true;
  return HSBP_OK;
}

#ifndef SW_REGS2_P6
#define SW_REGS2_P6 168
#endif
#ifndef SW_DEEP_REGS4
#define SW_DEEP_REGS4 200     // register cap of the DEEP, R = 4 variant (5 CTAs of 64 threads per SM)
#endif
#ifndef SW_DEEP_REGS2
#define SW_DEEP_REGS2 128     // register cap of the DEEP, R = 2 variant
#endif
template <int P, int R, int NT, bool DEEP> static int sweep_launch(hsbp_blocks *b, const double *u, double *y, bool with_faces,
                                                                   int64_t e0, int64_t ne) {
  hsbp_ctx *ctx = b->ctx;
  const bool odd = (b->max_Nr + 1) & 1;
  // register budget per thread.  Register windows (DEEP = false): R = 4 takes all 255; R = 2 fits 128 (p = 2, 4) --
  // p = 6 has 7-line windows and needs more.  DEEP: only u and the accumulators are per-point state.
  constexpr int REGS2 = (P == 6) ? SW_REGS2_P6 : 128;
  constexpr int MINB = (R == 2 ? 65536 / REGS2 : 256) / NT;
  void (*kern)(const SweepParams);
  const bool dot = b->sweep_dot_out != nullptr;               // u . y per chunk (SweepParams::dot): two-point deep kernels only
  if (dot && (!DEEP || R != 2)) { ctx->err = "k_sweep: the fused dot product needs the deep-ring two-point kernel"; return HSBP_ERR_STATE; }
  if constexpr (DEEP) {
    if constexpr (P == 6 && R == 2) {    // 7-line windows: 128 registers spill a little, 168 cost a CTA per SM
      if (odd) kern = dot ? k_sweep_deep<P, R, 168, true, true> : k_sweep_deep<P, R, 168, true>;
      else if (dot) kern = k_sweep_deep<P, R, 168, false, true>;
      else if (b->sweep_p6_regs == 168) kern = k_sweep_deep<P, R, 168>;
      else kern = k_sweep_deep<P, R, SW_DEEP_REGS2>;
    } else if (R == 2 && (odd || dot)) {
      if constexpr (R == 2)
        kern = odd ? (dot ? k_sweep_deep<P, 2, SW_DEEP_REGS2, true, true> : k_sweep_deep<P, 2, SW_DEEP_REGS2, true>)
                   : k_sweep_deep<P, 2, SW_DEEP_REGS2, false, true>;
      else kern = nullptr;
    } else {
      kern = k_sweep_deep<P, R, (R == 2 ? SW_DEEP_REGS2 : (P == 6 ? 255 : SW_DEEP_REGS4))>;
    }
  } else {
    kern = k_sweep<P, R, NT, MINB>;
  }
  if (odd && (!DEEP || R != 2)) { ctx->err = "k_sweep: odd line lengths need the deep-ring two-point kernel"; return HSBP_ERR_STATE; }
  const int Nrp = b->max_Nr + 1 + (odd ? 1 : 0), Nsp = b->max_Ns + 1;
  const int nthreads = ((Nrp / R) + 31) & ~31;
  const size_t sm = sweep_smem<P>(Nrp, DEEP, odd);
  int ctas_per_sm = 1;
  if (hsbp_smem_optin(ctx, kern, ctx->smem_optin) != cudaSuccess) {
    ctx->err = "cudaFuncSetAttribute(max dynamic shared memory) failed";
    return HSBP_ERR_CUDA;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, nthreads, sm) != cudaSuccess || ctas_per_sm < 1)
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
    const int64_t ctas = ne * 2 * ncs;
    const double waves = (double)ctas / (double)slots;
    const double eff = waves / std::ceil(waves) * ((double)per / (per + 2 * SweepCfg<P>::H));
    if (eff > best_eff + 1e-9) { best_eff = eff; best = ncs; }
  }
  if (b->sweep_ncs_override > 0) best = std::min(b->sweep_ncs_override, std::max(1, K / 16));
  SweepParams prm;
  prm.crr = b->d_crr_s; prm.css = b->d_css_s; prm.crs = odd ? b->d_crs_p : b->d_crs; prm.u = u; prm.y = y; prm.pitch = Nrp;
  prm.fcn = with_faces ? b->d_fa : nullptr; prm.fgm = with_faces ? b->d_fb : nullptr;
  prm.rtab = b->d_rtab;
  prm.active = b->skip_flags; prm.active_stride = b->skip_stride;
  prm.dot = b->sweep_dot_out;
  b->sweep_nch = 2 * best;
  prm.Nr = b->max_Nr; prm.Ns = b->max_Ns; prm.ncs = best; prm.K = K; prm.e0 = (int)e0;
  prm.per_up = (K + best - 1) / best;
  prm.per_dn = (Nsp - K + best - 1) / best;
  b->last_sweep_ctas_per_sm = ctas_per_sm;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ne * 2 * best)); cfg.blockDim = dim3((unsigned)nthreads); cfg.dynamicSmemBytes = sm; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = b->sweep_no_pdl ? 0 : 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, prm);
  }
  cudaError_t e1 = cudaGetLastError();
  if (e1 != cudaSuccess) {
    ctx->err = std::string("k_sweep: ") + cudaGetErrorString(e1);
    return HSBP_ERR_CUDA;
  }
  return HSBP_OK;
}

template <int P, int R> static int sweep_launch_nt(hsbp_blocks *b, const double *u, double *y, bool with_faces,
                                                   int64_t e0, int64_t ne) {
  const int Nrp = b->max_Nr + 1;
  const int nthreads = ((Nrp / R) + 31) & ~31;
  if (b->sweep_deep) return sweep_launch<P, R, 256, true>(b, u, y, with_faces, e0, ne);
  if (nthreads <= 64) return sweep_launch<P, R, 64, false>(b, u, y, with_faces, e0, ne);
  if (nthreads <= 128) return sweep_launch<P, R, 128, false>(b, u, y, with_faces, e0, ne);
  return sweep_launch<P, R, 256, false>(b, u, y, with_faces, e0, ne);
}

// y = A-tilde u (with_faces = false) or y = M-tilde u with the face terms prepared in d_fa / d_fb by k_face_prep
// for the blocks [e0, e0 + ne); u and y are the full concatenated vectors
template <int P> static int vol_sweep(hsbp_blocks *b, const double *u, double *y, bool with_faces,
                                     cudaEvent_t ev_between = nullptr, int64_t e0 = 0, int64_t ne = -1) {
  hsbp_ctx *ctx = b->ctx;
  if (!((b->max_Nr + 1) & 1) && (((uintptr_t)u & 15) || ((uintptr_t)y & 15))) {
    ctx->err = "hsbp_apply: u / y must be 16-byte aligned for the line-marching kernel";
    return HSBP_ERR_ARG;
  }
  if (ne < 0) ne = b->nblocks - e0;
  int rc = sweep_prepare<P>(b);
  if (rc) return rc;
  const size_t fsm = (2 + SweepTab<P>::BN) * (size_t)(std::max(b->max_Nr, b->max_Ns) + 1) * sizeof(double);
  const bool odd = (b->max_Nr + 1) & 1;
  if (odd) {                       // odd line lengths: pitched copies of the coefficient fields (sweep_prepare); u and y are the
                                   // caller's vectors (u by 8-byte cp.async, y by 8-byte stores)
    k_edge_prep<P, true><<<(unsigned)(4 * ne), 256, fsm, ctx->stream>>>(
        b->d_desc, b->d_crr, b->d_css, b->d_crs, b->d_tau, u, b->d_fa, b->d_fb, b->d_rtab, with_faces ? 1 : 0, (int)e0, b->d_rim,
        b->skip_flags, b->skip_stride, b->max_Nr, b->max_Ns);
  } else {
    k_edge_prep<P><<<(unsigned)(4 * ne), 256, fsm, ctx->stream>>>(
        b->d_desc, b->d_crr, b->d_css, b->d_crs, b->d_tau, u, b->d_fa, b->d_fb, b->d_rtab, with_faces ? 1 : 0, (int)e0, b->d_rim,
        b->skip_flags, b->skip_stride, b->max_Nr, b->max_Ns);
  }
  if (ev_between) cudaEventRecord(ev_between, ctx->stream);
  rc = sweep_points_per_thread(b) == 4 ? sweep_launch_nt<P, 4>(b, u, y, with_faces, e0, ne)
                                       : sweep_launch_nt<P, 2>(b, u, y, with_faces, e0, ne);
  return rc;
}

}  // namespace hsbp
