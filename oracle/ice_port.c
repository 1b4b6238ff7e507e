/* CPU ORACLE (test infrastructure, not product code): C restatement of the ICE loop of
 * the reference, _pytadbit/utils/normalize_hic.py:136-231 (iterative, _update_S,
 * _updateDB, _update_W, copy_matrix), in the non-mutating form S = x * (W0 x), x = 1/B.
 *
 * The matrix arrives as the CSR of the FULL symmetric matrix with the cells touching a
 * bad bin already removed (copy_matrix, :165-177).  "present" rows are those with at
 * least one stored cell (:173-176).  OpenMP over rows is the only liberty taken: the
 * reference is single-threaded; bench.py reports the thread count next to the number.
 *
 * Built by oracle/Makefile into oracle/_build/libice_port.so, loaded with ctypes by
 * oracle/tadbit_oracle.py:iterative_fast, checked against the numpy restatement and
 * through it against the reference's golden vectors (tests/test_oracle_port.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int ice_port_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* returns the number of passes executed, -1 if no row is present (:207-208), -2 if the mean row
 * sum is 0 (ZeroDivisionError in _updateDB, :148).
 * trace: (iterations+1) x 4 doubles = Smin, meanS, Smax, dev per pass (:219-220). */
int ice_port(int64_t n, const int64_t *indptr, const int32_t *indices, const double *data,
             int iterations, double max_dev, double *bias, double *trace, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  double *b = (double *)malloc(n * sizeof(double));
  double *x = (double *)malloc(n * sizeof(double));
  double *s = (double *)malloc(n * sizeof(double));
  unsigned char *present = (unsigned char *)malloc(n);
  int64_t npresent = 0;
  for (int64_t i = 0; i < n; ++i) {
    present[i] = indptr[i + 1] > indptr[i];
    npresent += present[i];
    b[i] = 1.0;
    x[i] = present[i] ? 1.0 : 0.0;
  }
  int passes = 0;
  double mean_s = 0.0;
  if (npresent == 0) { passes = -1; goto done; }
  for (int it = 0; it <= iterations; ++it) {
    double sum = 0.0, smin = INFINITY, smax = -INFINITY;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : sum) reduction(min : smin) reduction(max : smax)
    for (int64_t i = 0; i < n; ++i) {
      if (!present[i]) { s[i] = 0.0; continue; }
      double a = 0.0;
      for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) a += data[k] * x[indices[k]];
      a *= x[i];                                   /* _update_S on the corrected matrix */
      s[i] = a;
      sum += a;
      if (a < smin) smin = a;
      if (a > smax) smax = a;
    }
    mean_s = sum / (double)npresent;
    if (mean_s == 0.0) { passes = -2; goto done; } /* float(S) / meanS raises in _updateDB (:148) */
    for (int64_t i = 0; i < n; ++i)                /* _updateDB */
      if (present[i]) {
        b[i] *= s[i] / mean_s;
        x[i] = b[i] != 0.0 ? 1.0 / b[i] : 0.0;     /* _update_W, ZeroDivisionError -> skip */
      }
    ++passes;
    if (iterations == 0) break;
    double dev = fmax(fabs(smin / mean_s - 1.0), fabs(smax / mean_s - 1.0));
    trace[4 * it + 0] = smin; trace[4 * it + 1] = mean_s;
    trace[4 * it + 2] = smax; trace[4 * it + 3] = dev;
    if (dev < max_dev) break;
  }
  for (int64_t i = 0; i < n; ++i)                  /* :223-230 */
    bias[i] = (present[i] && b[i] != 0.0) ? b[i] * sqrt(mean_s) : 1.0;
done:
  free(b); free(x); free(s); free(present);
  return passes;
}
