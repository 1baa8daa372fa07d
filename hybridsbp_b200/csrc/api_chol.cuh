// Batched dense fp64 Cholesky local solver (K2a) -- placeholder until the DMMA kernels land.
#pragma once
namespace {
int chol_setup(hsbp_blocks *b) { HSBP_FAIL(b->ctx, HSBP_ERR_UNSUPP, "dense Cholesky local solver not built yet"); }
int chol_solve(hsbp_blocks *b, const double *, double *, hsbp_local_stats *) {
  HSBP_FAIL(b->ctx, HSBP_ERR_UNSUPP, "dense Cholesky local solver not built yet");
}
}  // namespace
