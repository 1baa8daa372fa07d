// Line-marching TMA kernel for the volume part (fast path) -- see DESIGN.md section 4.
#pragma once
#include "hsbp_internal.h"
#include "sbp1d.cuh"

namespace hsbp {
template <int P> static bool march_eligible(const hsbp_blocks *b) { (void)b; return false; }
template <int P> static int vol_march(hsbp_blocks *b, const double *u, double *y) {
  (void)u; (void)y;
  b->ctx->err = "marching kernel not available";
  return HSBP_ERR_UNSUPP;
}
}  // namespace hsbp
