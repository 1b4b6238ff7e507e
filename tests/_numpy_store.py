"""Test double for ``tadbit_b200.ContactStore`` (TESTS ONLY): the same methods the host classes
call on a store, answered on the CPU by the oracle.  It lets ``-m "not gpu"`` tests drive the
product's host layer (tadbit_b200/hic_data.py, hic_filtering.py, normalize_hic.py: thresholds,
numpy tail of filter_by_mean, pooling of expected, rescaling, printed lines, exceptions) next to
the unmodified reference.  The product itself never falls back to this: without the CUDA library
``ContactStore`` raises."""
import numpy as np

import tadbit_oracle as orc


class NumpyStore(object):
    def __init__(self, size, rows, cols, vals, chromosomes=None, symmetricized=False):
        self.size = int(size)
        self.m = orc.OracleMatrix(size, rows, cols, vals, chromosomes=chromosomes,
                                  symmetricized=symmetricized)

    @classmethod
    def from_coo(cls, size, rows, cols, vals, chromosomes=None):   # ContactStore.from_coo
        return cls(size, rows, cols, vals, chromosomes=chromosomes)

    @staticmethod
    def _bads(mask):
        if mask is None:
            return {}
        return dict((int(i), True) for i in np.flatnonzero(np.asarray(mask)))

    def row_stats(self):                                   # tb_row_stats
        return orc.row_stats(self.m)

    def col_sums_masked(self, bad=None):                   # tb_col_sums_masked
        mask = np.zeros(self.size, dtype=bool) if bad is None else np.asarray(bad).astype(bool)
        return orc.masked_col_sums(self.m, mask)

    def ice(self, bad=None, iterations=0, max_dev=1e-5):   # tb_ice
        bias, trace = orc.iterative(self.m, bads=self._bads(bad), iterations=iterations,
                                    max_dev=max_dev)
        arr = np.array([(t[0], t[1], t[2], t[4]) for t in trace], dtype=np.float64).reshape(-1, 4)
        return np.asarray(bias, dtype=np.float64), arr

    def norm_sum(self, bias=None, bad=None):               # tb_norm_sum
        return orc.matrix_sum(self.m, bias, self._bads(bad))

    def diag_sums(self, bad, ndiag):                       # tb_diag_sums
        return np.asarray(orc.diagonal_sums(self.m, self._bads(bad), ndiag)[0], dtype=np.int64)

    def pair_stats(self, bias=None, bad=None, center=0.0):  # tb_pair_stats
        r, c, v = orc.pair_values(self.m, bias, self._bads(bad))
        if v.size == 0:
            return 0, float("inf"), 0.0, 0.0
        logs = np.log10(v.astype(np.float64))
        return int(v.size), float(v.min()), float(logs.sum()), float(((logs - center) ** 2).sum())

    def pair_values(self, bias=None, bad=None, zscore=None):  # tb_pair_values
        r, c, v = orc.pair_values(self.m, bias, self._bads(bad))
        v = v.astype(np.float64)
        if zscore is not None:
            v = (np.log10(v) - zscore[0]) / zscore[1]
        return r.astype(np.int32), c.astype(np.int32), v

    def window(self, r0, r1, c0, c1, bias=None):           # tb_window
        full = self.m.full_csr(dtype=np.float64)[r0:r1, c0:c1].toarray()
        if bias is not None:
            b = np.asarray(bias, dtype=np.float64)
            full = full / b[r0:r1, None] / b[None, c0:c1]
        return full

    def corr_build(self, beg, end, bias, bad, expected, stage=1):   # tb_corr_build
        mat = orc.obs_exp_matrix(self.m, beg, end, bias, self._bads(bad), expected)
        self._dense = orc.correlation_matrix(mat) if stage >= 1 and mat.shape[0] else mat
        return self._dense.shape[0]

    def corr_get(self):                                    # tb_corr_get
        return np.array(self._dense, dtype=np.float64)

    def corr_set(self, matrix):                            # tb_corr_set
        self._dense = np.array(matrix, dtype=np.float64)

    def corr_eig(self, k):                                 # tb_corr_eig
        w, v = orc.largest_eigenpairs(self._dense, k)
        return w, v, dict(steps=0, restarts=0, converged=True)

    def corr_timing(self):                                 # tb_corr_timing
        return {}

    def region_cells(self, r0, r1, c0, c1, bad=None, bias=None, decay=None, half=False):  # tb_region_cells
        table = None
        if decay is not None:
            table = dict((crm, dict((d, float(decay[beg + d])) for d in range(end - beg)))
                         for crm, (beg, end) in self.m.section_pos.items())
        r, c, v, nrm, dec = orc.region_cells(self.m, r0, r1, c0, c1, self._bads(bad), bias, table, half)
        return r.astype(np.int32), c.astype(np.int32), v.astype(np.int32), nrm, dec

    def decay_sums(self, bias, bad=None):                  # tb_decay_sums
        return orc.decay_sums(self.m, bias, self._bads(bad))

    def export_upper(self):                                # tb_store_export_upper
        return (self.m.rows.astype(np.int32), self.m.cols.astype(np.int32), self.m.vals)

    def close(self):
        pass
