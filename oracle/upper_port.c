/* CPU ORACLE (test infrastructure, not product code): the reference's reductions restated in C
 * over the stored UPPER-triangle cells (row <= col, int32 triplets, any order unless stated), for
 * matrices too large for the numpy restatement (tadbit_oracle.py) -- BASELINE configs[1..4] at
 * full size.  Each function cites the reference lines it follows; each is checked against the
 * numpy restatement (itself pinned to the unmodified reference's golden vectors) in
 * tests/test_oracle_port.py.
 *
 * The reference holds both triangles in a dict (hic_data.py:54-118); a stored upper cell (i, j, v)
 * stands for the two entries (i, j) and (j, i), the diagonal for one.
 *
 * OpenMP is the only liberty taken (the reference is single-threaded): threads own contiguous
 * ranges of cells and private accumulators, which are then added in thread order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int nthreads_for(int threads) {
#ifdef _OPENMP
  /* 0 = every core of the host, whatever OMP_NUM_THREADS says (torchrun sets it to 1) */
  return threads > 0 ? threads : omp_get_num_procs();
#else
  (void)threads;
  return 1;
#endif
}

/* hic_filtering.py:187-190 (stored keys per row) and :198-200 (row sums) */
int up_row_stats(int64_t n, int64_t nnz, const int32_t *row, const int32_t *col, const int32_t *val,
                 int64_t *nnz_row, int64_t *row_sum, int threads) {
  const int T = nthreads_for(threads);
  int64_t *pn = (int64_t *)calloc((size_t)T * n, sizeof(int64_t));
  int64_t *ps = (int64_t *)calloc((size_t)T * n, sizeof(int64_t));
  if (!pn || !ps) { free(pn); free(ps); return -1; }
#pragma omp parallel num_threads(T)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    int64_t *mn = pn + (size_t)t * n, *ms = ps + (size_t)t * n;
    const int64_t b = nnz * t / T, e = nnz * (t + 1) / T;
    for (int64_t k = b; k < e; ++k) {
      const int32_t i = row[k], j = col[k];
      mn[i] += 1; ms[i] += val[k];
      if (i != j) { mn[j] += 1; ms[j] += val[k]; }
    }
  }
  for (int64_t i = 0; i < n; ++i) {
    int64_t a = 0, s = 0;
    for (int t = 0; t < T; ++t) { a += pn[(size_t)t * n + i]; s += ps[(size_t)t * n + i]; }
    nnz_row[i] = a; row_sum[i] = s;
  }
  free(pn); free(ps);
  return 0;
}

/* hic_filtering.py:36-39 / :141-145: sum of column j over the rows that are not bad */
int up_masked_col_sums(int64_t n, int64_t nnz, const int32_t *row, const int32_t *col, const int32_t *val,
                       const uint8_t *bad, int64_t *out, int threads) {
  const int T = nthreads_for(threads);
  int64_t *ps = (int64_t *)calloc((size_t)T * n, sizeof(int64_t));
  if (!ps) return -1;
#pragma omp parallel num_threads(T)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    int64_t *ms = ps + (size_t)t * n;
    const int64_t b = nnz * t / T, e = nnz * (t + 1) / T;
    for (int64_t k = b; k < e; ++k) {
      const int32_t i = row[k], j = col[k];
      if (!bad[i]) ms[j] += val[k];                 /* entry (i, j): row i, column j */
      if (i != j && !bad[j]) ms[i] += val[k];       /* entry (j, i): row j, column i */
    }
  }
  for (int64_t i = 0; i < n; ++i) {
    int64_t s = 0;
    for (int t = 0; t < T; ++t) s += ps[(size_t)t * n + i];
    out[i] = s;
  }
  free(ps);
  return 0;
}

/* normalize_hic.py:271-283: cells (i, i + d) of one section, only the row index i tested */
int up_diag_sums(int64_t n, int64_t nnz, const int32_t *row, const int32_t *col, const int32_t *val,
                 const uint8_t *bad, const int32_t *chrom_of, int64_t ndiag, int64_t *out, int threads) {
  const int T = nthreads_for(threads);
  int64_t *ps = (int64_t *)calloc((size_t)T * ndiag, sizeof(int64_t));
  if (!ps) return -1;
  (void)n;
#pragma omp parallel num_threads(T)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    int64_t *ms = ps + (size_t)t * ndiag;
    const int64_t b = nnz * t / T, e = nnz * (t + 1) / T;
    for (int64_t k = b; k < e; ++k) {
      const int32_t i = row[k], j = col[k];
      const int64_t d = (int64_t)j - i;
      if (d < ndiag && !bad[i] && chrom_of[i] == chrom_of[j]) ms[d] += val[k];
    }
  }
  for (int64_t d = 0; d < ndiag; ++d) {
    int64_t s = 0;
    for (int t = 0; t < T; ++t) s += ps[(size_t)t * ndiag + d];
    out[d] = s;
  }
  free(ps);
  return 0;
}

/* hic_data.py:350-375 with biases: sum of v / bias_i / bias_j over both triangles, bad rows and
 * columns skipped */
double up_norm_sum(int64_t nnz, const int32_t *row, const int32_t *col, const int32_t *val,
                   const uint8_t *bad, const double *bias, int threads) {
  const int T = nthreads_for(threads);
  double total = 0.0;
#pragma omp parallel for num_threads(T) reduction(+ : total) schedule(static)
  for (int64_t k = 0; k < nnz; ++k) {
    const int32_t i = row[k], j = col[k];
    if (bad[i] || bad[j]) continue;
    const double t = (double)val[k] / bias[i] / bias[j];
    total += i == j ? t : 2.0 * t;
  }
  return total;
}

/* normalize_hic.py:136-231 (copy_matrix, _update_S, _updateDB, _update_W, iterative) in the
 * non-mutating form S = x * (W0 x), x = 1 / B.  Returns the passes executed, -1 when no row is
 * present (:207-208), -2 when the mean row sum is 0 (_updateDB divides by it), -3 out of memory.
 * trace: (iterations + 1) x 4 doubles = Smin, meanS, Smax, dev (:219-220). */
int up_ice(int64_t n, int64_t nnz, const int32_t *row, const int32_t *col, const int32_t *val,
           const uint8_t *bad, int iterations, double max_dev, double *bias, double *trace, int threads) {
  const int T = nthreads_for(threads);
  double *b = (double *)malloc(n * sizeof(double));
  double *x = (double *)malloc(n * sizeof(double));
  double *s = (double *)malloc(n * sizeof(double));
  double *priv = (double *)malloc((size_t)T * n * sizeof(double));
  unsigned char *present = (unsigned char *)calloc(n, 1);
  int passes = 0;
  double mean_s = 0.0;
  int64_t npresent = 0;
  if (!b || !x || !s || !priv || !present) { passes = -3; goto done; }
  for (int64_t k = 0; k < nnz; ++k)              /* copy_matrix :165-177: cells with both ends good */
    if (!bad[row[k]] && !bad[col[k]]) present[row[k]] = present[col[k]] = 1;
  for (int64_t i = 0; i < n; ++i) {
    npresent += present[i];
    b[i] = 1.0;
    x[i] = present[i] ? 1.0 : 0.0;
  }
  if (npresent == 0) { passes = -1; goto done; }
  for (int it = 0; it <= iterations; ++it) {
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
      const int t = omp_get_thread_num();
#else
      const int t = 0;
#endif
      double *ms = priv + (size_t)t * n;
      memset(ms, 0, n * sizeof(double));
      const int64_t kb = nnz * t / T, ke = nnz * (t + 1) / T;
      for (int64_t k = kb; k < ke; ++k) {          /* x is 0 on rows that are not present */
        const int32_t i = row[k], j = col[k];
        const double v = (double)val[k];
        ms[i] += v * x[j];
        if (i != j) ms[j] += v * x[i];
      }
    }
    double sum = 0.0, smin = INFINITY, smax = -INFINITY;
#pragma omp parallel for num_threads(T) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      double a = 0.0;
      for (int t = 0; t < T; ++t) a += priv[(size_t)t * n + i];
      s[i] = a * x[i];                             /* _update_S on the corrected matrix */
    }
    for (int64_t i = 0; i < n; ++i)
      if (present[i]) {
        sum += s[i];
        if (s[i] < smin) smin = s[i];
        if (s[i] > smax) smax = s[i];
      }
    mean_s = sum / (double)npresent;
    if (mean_s == 0.0) { passes = -2; goto done; }
    for (int64_t i = 0; i < n; ++i)                /* _updateDB */
      if (present[i]) {
        b[i] *= s[i] / mean_s;
        x[i] = b[i] != 0.0 ? 1.0 / b[i] : 0.0;     /* _update_W, ZeroDivisionError -> skip */
      }
    ++passes;
    if (iterations == 0) break;
    {
      const double dev = fmax(fabs(smin / mean_s - 1.0), fabs(smax / mean_s - 1.0));
      trace[4 * it + 0] = smin; trace[4 * it + 1] = mean_s;
      trace[4 * it + 2] = smax; trace[4 * it + 3] = dev;
      if (dev < max_dev) break;
    }
  }
  for (int64_t i = 0; i < n; ++i)                  /* :223-230 */
    bias[i] = (present[i] && b[i] != 0.0) ? b[i] * sqrt(mean_s) : 1.0;
done:
  free(b); free(x); free(s); free(priv); free(present);
  return passes;
}
