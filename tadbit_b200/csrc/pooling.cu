// Host-side tails of the distance reductions: the pooling walks that turn per-distance sums into the
// `expected` dictionary (_pytadbit/utils/normalize_hic.py:262-291, _meandiag) and into the per-
// chromosome `decay` of `tadbit normalize` (_pytadbit/tools/tadbit_normalize.py:1077-1106).  The sums
// come from the device (tb_diag_sums, tb_decay_sums); the walks are sequential by nature (a group of
// distances closes when it holds more than min_n reads) and cost more than the kernels when written as
// Python loops over 10^5..10^6 distances, so they are here.  Same order of additions and the same
// divisions as the reference's Python (doubles are Python floats; the integer sums stay below 2^53).
#include <cmath>
#include <vector>

#include "common.cuh"

namespace tb {

// expected(): keys 0 .. *nkeys-1 in ascending order, out_val[key] = value (out_val has size + 2 slots)
void pool_expected(int64_t size, int64_t ndiag, const int64_t *sums, const int64_t *cells, double min_n,
                   double *out_val, int64_t *nkeys) {
  TB_REQUIRE(size >= 0 && ndiag >= 0 && out_val && nkeys && (ndiag == 0 || (sums && cells)),
             "pool_expected: bad arguments");
  int64_t dist = 0, top = 0;
  while (dist < size) {
    int64_t total = 0, ncell = 0, stop = dist;
    double val;
    for (;;) {                                  // _meandiag's recursion, unrolled
      if (stop < ndiag) { total += sums[stop]; ncell += cells[stop]; }
      if (ncell == 0) { val = 0.0; break; }
      if ((double)total > min_n || stop >= size) { val = (double)total / (double)ncell; break; }
      ++stop;
    }
    for (int64_t d = dist; d < stop + 2; ++d) out_val[d] = val;
    if (stop + 2 > top) top = stop + 2;
    dist = stop + 1;                            // the reference's loop variable ends on the extra key
  }
  *nkeys = top;
}

// decay of one chromosome: out_val[k] and out_state[k] = 0 no key, 1 key the walk began with
// (a distance with reads), 2 key added by the walk; *empty = 1 when the reference ends with {}
void pool_decay(int64_t size, const double *nrm, const int64_t *raw, const int64_t *ndiag, double min_n,
                double *out_val, uint8_t *out_state, int32_t *empty) {
  TB_REQUIRE(size >= 0 && empty && (size == 0 || (nrm && raw && ndiag && out_val && out_state)),
             "pool_decay: bad arguments");
  *empty = 0;
  for (int64_t k = 0; k < size; ++k) {
    out_state[k] = raw[k] != 0;
    out_val[k] = raw[k] != 0 ? nrm[k] : 0.0;
  }
  double tmpdec = 0.0, val = 0.0;
  int64_t tmpsum = 0, ndg = 0, first = 0;       // the open group is [first, k]
  auto assign = [&](int64_t beg, int64_t end) {
    if (ndg == 0) return;                       // ZeroDivisionError: every column at these distances is bad
    const double ratio = val / (double)ndg;
    for (int64_t q = beg; q < end; ++q) {
      out_val[q] = ratio;
      if (!out_state[q]) out_state[q] = 2;
    }
  };
  for (int64_t k = 0; k < size; ++k) {
    if (raw[k] != 0) { tmpdec += nrm[k]; tmpsum += raw[k]; }
    if ((double)tmpsum > min_n) {
      ndg = 0;
      for (int64_t q = first; q <= k; ++q) ndg += ndiag[q];
      val = tmpdec;
      assign(first, k + 1);
      first = k + 1;
      tmpdec = 0.0;
      tmpsum = 0;
    }
  }
  if (first == 0 && size > 0) {                 // no group ever closed
    *empty = 1;
  } else if (first < size && (double)tmpsum < min_n) {
    for (int64_t q = first; q < size; ++q) ndg += ndiag[q];
    val += tmpdec;
    assign(first, size);
  }
}

}  // namespace tb
