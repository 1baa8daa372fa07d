/* ORACLE / CPU baseline -- test and benchmark infrastructure only.
 *
 * The reference's CPU "operator apply" is the sparse product  lop[e].M̃ * u  of the matrix that
 * locoperator assembles (global_curved.jl:470-492), executed by its host language's serial
 * SparseMatrixCSC mul!.  This is that product in plain C over the CSR form of the same
 * (symmetric) matrices, rows in parallel with OpenMP when more than one thread is allowed.
 * Built by __graft_entry__.build() into oracle/_build/libspmv.so; never linked into libhsbp.
 */
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* y = A x for nb independent CSR blocks stored back to back:
 *   rowptr: concatenated (rows_b + 1) entries per block, local numbering
 *   blk_rowptr_off[b], blk_nnz_off[b], blk_row_off[b]: where block b starts in rowptr / (col,val) / (x,y) */
void hsbp_oracle_spmv_blocks(int64_t nb, const int64_t *blk_rowptr_off, const int64_t *blk_nnz_off,
                             const int64_t *blk_row_off, const int64_t *blk_rows,
                             const int64_t *rowptr, const int32_t *col, const double *val,
                             const double *x, double *y, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  for (int64_t b = 0; b < nb; ++b) {
    const int64_t *rp = rowptr + blk_rowptr_off[b];
    const int32_t *c = col + blk_nnz_off[b];
    const double *v = val + blk_nnz_off[b];
    const double *xb = x + blk_row_off[b];
    double *yb = y + blk_row_off[b];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < blk_rows[b]; ++i) {
      double acc = 0.0;
      for (int64_t k = rp[i]; k < rp[i + 1]; ++k) acc += v[k] * xb[c[k]];
      yb[i] = acc;
    }
  }
}

int hsbp_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
