// K5: rate-and-state fault stage of the SEAS BP1 ODE right-hand side, one thread per fault node.
//
// Reference: seas/BP1/odefun.jl:59-108 (traction -> bracketed Newton -> state evolution),
// global_curved.jl:1031-1039 (rateandstate), :1041-1075 (newtbndv), :627-634 (computetraction_mod).
// The whole stage -- shear traction from the traction operator HfI_FT u, the root find for the slip
// rate and d(psi)/dt -- is one kernel; failures (no bracket, NaN, iteration limit, non-finite dpsi) are
// counted so that the host integrator can reject the step exactly where the reference sets reject_step.
#pragma once
#include "hsbp_internal.h"

namespace hsbp {

struct Bp1Dev {
  double mu_shear, sigma_n, eta, V0, tau_z0, Dc, f0, b;
  double ftol, atolx, rtolx;
  int maxiter;
};

// g(V) = sigma_n a asinh(V Y) + eta V - tau,  Y = exp(psi / a) / (2 V0)      global_curved.jl:1031-1039
__device__ __forceinline__ void rateandstate(double V, double psi, double sigma_n, double phi, double eta, double a,
                                             double V0, double &g, double &dg) {
  const double Y = (1.0 / (2.0 * V0)) * exp(psi / a);
  const double f = a * asinh(V * Y);
  const double dfdV = a * (1.0 / sqrt(1.0 + (V * Y) * (V * Y))) * Y;
  g = sigma_n * f + eta * V - phi;
  dg = sigma_n * dfdV + eta;
}

// status bits written per launch (d_flags[0] |= ...), d_flags[1] = max Newton iterations
enum { BP1_TAU_NAN = 1, BP1_V_FAIL = 2, BP1_PSI_FAIL = 4 };

// One fault node: shear traction -> bracketed Newton for the slip rate -> state evolution.
// dtau: change of the shear stress on the fault at this node (odefun.jl:59); state = [psi; delta] (2 nf), out = [dpsi; V] (2 nf)
__device__ __forceinline__ void bp1_fault_node_dtau(int n, int nf, double dtau, double a, const double *__restrict__ state,
                                                    double *__restrict__ out, const Bp1Dev &prm, int *__restrict__ flags) {
  const double psi = state[n];
  const double taun = dtau + prm.tau_z0;
  double Vout = 0.0, dpsi = 0.0;
  int fl = 0, iters = 0;
  if (isnan(taun)) {
    fl = BP1_TAU_NAN;                                                     // odefun.jl:73-78
  } else {
    double xR = fabs(taun / prm.eta), xL = -xR;                         // odefun.jl:80-81
    double x = 0.0;                                                       // initial guess V[n] = 0 (odefun.jl:51, 82)
    double fL, fR, f, df, tmp;
    rateandstate(xL, psi, prm.sigma_n, taun, prm.eta, a, prm.V0, fL, tmp);
    rateandstate(xR, psi, prm.sigma_n, taun, prm.eta, a, prm.V0, fR, tmp);
    bool ok = false;
    if (fL * fR > 0.0) {                                                  // newtbndv: no bracket -> NaN, iter < 0
      x = nan("");
    } else {
      rateandstate(x, psi, prm.sigma_n, taun, prm.eta, a, prm.V0, f, df);
      for (int it = 1; it <= prm.maxiter; ++it) {
        double dx = -f / df;
        x = x + dx;
        if (x < xL || x > xR) {                                           // minchange = 0: |dx|/dxlr < 0 never holds
          x = (xR + xL) / 2.0;
          dx = (xR - xL) / 2.0;
        }
        rateandstate(x, psi, prm.sigma_n, taun, prm.eta, a, prm.V0, f, df);
        if (f * fL > 0.0) { fL = f; xL = x; } else { fR = f; xR = x; }
        iters = it;
        if (fabs(f) < prm.ftol && fabs(dx) < prm.atolx + prm.rtolx * (fabs(dx) + fabs(x))) { ok = true; break; }
      }
    }
    if (!ok || isnan(x)) {
      fl = BP1_V_FAIL;                                                    // odefun.jl:91-96
    } else {
      Vout = x;
      dpsi = (prm.b * prm.V0 / prm.Dc) * (exp((prm.f0 - psi) / prm.b) - fabs(x) / prm.V0);     // odefun.jl:101
      if (!isfinite(dpsi)) { dpsi = 0.0; fl = BP1_PSI_FAIL; }             // odefun.jl:102-107
    }
  }
  out[n] = dpsi;
  out[nf + n] = Vout;
  if (fl) atomicOr(&flags[0], fl);
  if (fl) atomicAdd(&flags[2], 1);
  atomicMax(&flags[1], iters);
}

// trn: (HfI_FT_k u)_n on the fault face
__device__ __forceinline__ void bp1_fault_node(int n, int nf, double trn, double taupen, double sJn, double a,
                                               const double *__restrict__ state, double *__restrict__ out,
                                               const Bp1Dev &prm, int *__restrict__ flags) {
  const double delta = state[nf + n];
  // computetraction_mod: (HfI_FT u + tau (delta - delta/2)) / sJ ; odefun.jl:59
  const double T = (trn + taupen * (delta - delta / 2.0)) / sJn;
  bp1_fault_node_dtau(n, nf, -prm.mu_shear * T, a, state, out, prm, flags);
}

// tr: HfI_FT_k u on the fault face (k_face_gather, FACE_TRACTION); tau: penalty on that face
__global__ void __launch_bounds__(128)
k_bp1_fault(int nf, const double *__restrict__ tr, const double *__restrict__ tau, const double *__restrict__ sJ,
            const double *__restrict__ rsa, const double *__restrict__ state, double *__restrict__ out,
            Bp1Dev prm, int *__restrict__ flags) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nf) return;
  bp1_fault_node(n, nf, tr[n], tau[n], sJ[n], rsa[n], state, out, prm, flags);
}

// The same stage with the local solve condensed onto the fault (hsbp_bp1_condense): the displacement enters odefun only
// through the traction on the fault face, and u = M̃^-1 ge is linear in the boundary data (odefun.jl:36-43), so
//   HfI_FT_1 u = -1/2 Tf delta - (t Vp / 2) tl,   Tf = HfI_FT_1 M̃^-1 F_1 (nf x nf),  tl = HfI_FT_1 M̃^-1 F_2 1
// (the products assembleλmatrix forms for interface faces, global_curved.jl:759-790, here for the two Dirichlet faces
// of the BP1 block).  One kernel: a small dense matrix-vector product per node, then the root find.  state / out live
// in mapped host memory: no copies around the launch.
__global__ void __launch_bounds__(64)
k_bp1_fault_condensed(int nf, const double *__restrict__ Tf, const double *__restrict__ tl, double load,
                      const double *__restrict__ tau, const double *__restrict__ sJ, const double *__restrict__ rsa,
                      const double *__restrict__ state, double *__restrict__ out, Bp1Dev prm, int *__restrict__ flags) {
  extern __shared__ double st[];                 // [psi; delta]
  for (int i = threadIdx.x; i < 2 * nf; i += blockDim.x) st[i] = state[i];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nf) return;
  double s0 = 0.0, s1 = 0.0;
  int m = 0;
  for (; m + 1 < nf; m += 2) { s0 += Tf[n + (int64_t)nf * m] * st[nf + m]; s1 += Tf[n + (int64_t)nf * (m + 1)] * st[nf + m + 1]; }
  if (m < nf) s0 += Tf[n + (int64_t)nf * m] * st[nf + m];
  const double trn = -0.5 * (s0 + s1) - load * tl[n];
  bp1_fault_node(n, nf, trn, tau[n], sJ[n], rsa[n], st, out, prm, flags);
}

// Fault nodes of a multiblock mesh (hsbp_fault_*): the stress change is a given linear function of slip and time,
// dtau = A delta + t b (condensed trace solve), or comes precomputed (dtau_in, one trace solve per call).
__global__ void __launch_bounds__(64)
k_fault_linear(int nf, const double *__restrict__ A, const double *__restrict__ b, double t, const double *__restrict__ dtau_in,
               const double *__restrict__ rsa, const double *__restrict__ state, double *__restrict__ out, Bp1Dev prm,
               int *__restrict__ flags) {
  extern __shared__ double st[];                 // [psi; delta]
  for (int i = threadIdx.x; i < 2 * nf; i += blockDim.x) st[i] = state[i];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nf) return;
  double dtau;
  if (dtau_in) {
    dtau = dtau_in[n];
  } else {
    double s0 = 0.0, s1 = 0.0;
    int m = 0;
    for (; m + 1 < nf; m += 2) { s0 += A[n + (int64_t)nf * m] * st[nf + m]; s1 += A[n + (int64_t)nf * (m + 1)] * st[nf + m + 1]; }
    if (m < nf) s0 += A[n + (int64_t)nf * m] * st[nf + m];
    dtau = (s0 + s1) + t * b[n];
  }
  bp1_fault_node_dtau(n, nf, dtau, rsa[n], st, out, prm, flags);
}

// Dirichlet data of the ODE stage on the block-face vector v (zero elsewhere): fault face <- delta / 2,
// loading face <- t Vp / 2   (odefun.jl:36)
__global__ void k_bp1_bc(int nf_fault, int64_t off_fault, const double *__restrict__ state, int nf_load, int64_t off_load,
                         double load_value, double *__restrict__ v) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < nf_fault) v[off_fault + n] = state[nf_fault + n] / 2.0;
  if (n < nf_load) v[off_load + n] = load_value;
}

}  // namespace hsbp
