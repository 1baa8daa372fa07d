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

// tr: HfI_FT_k u on the fault face (k_face_gather, FACE_TRACTION); tau: penalty on that face;
// state = [psi; delta] (2 nf), out = [dpsi; V] (2 nf)
__global__ void __launch_bounds__(128)
k_bp1_fault(int nf, const double *__restrict__ tr, const double *__restrict__ tau, const double *__restrict__ sJ,
            const double *__restrict__ rsa, const double *__restrict__ state, double *__restrict__ out,
            Bp1Dev prm, int *__restrict__ flags) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nf) return;
  const double psi = state[n], delta = state[nf + n];
  // computetraction_mod: (HfI_FT u + tau (delta - delta/2)) / sJ ; odefun.jl:59
  const double T = (tr[n] + tau[n] * (delta - delta / 2.0)) / sJ[n];
  const double dtau = -prm.mu_shear * T;
  const double taun = dtau + prm.tau_z0;
  double Vout = 0.0, dpsi = 0.0;
  int fl = 0, iters = 0;
  if (isnan(taun)) {
    fl = BP1_TAU_NAN;                                                     // odefun.jl:73-78
  } else {
    const double a = rsa[n];
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

// Dirichlet data of the ODE stage on the block-face vector v (zero elsewhere): fault face <- delta / 2,
// loading face <- t Vp / 2   (odefun.jl:36)
__global__ void k_bp1_bc(int nf_fault, int64_t off_fault, const double *__restrict__ state, int nf_load, int64_t off_load,
                         double load_value, double *__restrict__ v) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < nf_fault) v[off_fault + n] = state[nf_fault + n] / 2.0;
  if (n < nf_load) v[off_load + n] = load_value;
}

}  // namespace hsbp
