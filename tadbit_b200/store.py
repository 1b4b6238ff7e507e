"""Device-resident contact store: thin object wrapper over the C-ABI handle.

One ``ContactStore`` = one ``tb_store`` = the rows of the full symmetric matrix that one
GPU owns.  All numerics run in libtadbit_b200 (CUDA); this module only moves numpy
buffers across the boundary.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import SynthParams, check, lib, ptr


def _offsets(size, chromosomes):
    """chromosomes (OrderedDict name -> nbins) -> (n_chrom, int64 offsets or None)."""
    if not chromosomes:
        return 0, None
    sizes = np.array([int(v) for v in chromosomes.values()], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    if offs[-1] != size:
        raise ValueError("chromosome sizes sum to %d, matrix has %d bins" % (offs[-1], size))
    return len(sizes), offs


class CommGroup(object):
    """The GPUs that hold the row blocks of one matrix (one process per GPU).  Creating the group
    sets up an NCCL communicator (slow); create it once and ``attach`` every sharded store."""

    def __init__(self, rank, world, unique_id, device=0):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        self._h = C.c_void_p()
        self.rank, self.world = int(rank), int(world)
        check(lib().tb_comm_create(rank, world, ptr(uid, C.c_uint8), device, C.byref(self._h)))

    def close(self):
        if self._h is not None and self._h.value:
            lib().tb_comm_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ContactStore(object):
    def __init__(self, handle, size):
        self._h = handle
        self.size = int(size)

    # -- construction ---------------------------------------------------------
    @classmethod
    def from_coo(cls, size, rows, cols, vals, chromosomes=None, row_block=None, device=0):
        """Upper-triangle cells (row <= col, unique) as host arrays."""
        rows, cols, vals = _capi.as_i32(rows, "rows"), _capi.as_i32(cols, "cols"), _capi.as_i32(vals, "vals")
        if not (rows.shape == cols.shape == vals.shape and rows.ndim == 1):
            raise ValueError("rows, cols, vals must be 1-D arrays of equal length")
        n_chrom, offs = _offsets(size, chromosomes)
        rb, re = row_block if row_block is not None else (0, size)
        out = C.c_void_p()
        check(lib().tb_store_create_coo(
            size, rows.size, rows.ctypes.data, cols.ctypes.data, vals.ctypes.data,
            _capi.TB_MEM_HOST, n_chrom, ptr(offs, C.c_int64) if offs is not None else None,
            rb, re, device, C.byref(out)))
        return cls(out, size)

    @classmethod
    def from_csr(cls, size, indptr, cols, vals, form="upper", chromosomes=None, row_block=None,
                 device=0):
        """Cells in CSR form as host arrays (pinned memory makes the copy faster).

        ``form="upper"``: upper-triangle rows of the whole matrix, ``indptr`` has ``size + 1``
        entries.  ``form="full_rows"``: every stored cell of the rows ``row_block`` of the full
        symmetric matrix (both triangles, columns ascending) -- what each GPU of a row-sharded
        matrix is given."""
        forms = {"upper": _capi.TB_CELLS_UPPER, "full_rows": _capi.TB_CELLS_FULL_ROWS}
        if form not in forms:
            raise ValueError("form must be 'upper' or 'full_rows'")
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        cols, vals = _capi.as_i32(cols, "cols"), _capi.as_i32(vals, "vals")
        rb, re = row_block if row_block is not None else (0, size)
        nrows = size if form == "upper" else re - rb
        if indptr.shape != (nrows + 1,) or cols.shape != vals.shape or cols.ndim != 1:
            raise ValueError("indptr must have %d entries; cols and vals equal 1-D shapes" % (nrows + 1))
        if indptr[0] != 0 or indptr[-1] != cols.size:
            raise ValueError("indptr must run from 0 to the number of cells")
        n_chrom, offs = _offsets(size, chromosomes)
        out = C.c_void_p()
        check(lib().tb_store_create_csr(
            size, forms[form], indptr.ctypes.data, cols.ctypes.data, vals.ctypes.data,
            _capi.TB_MEM_HOST, n_chrom, ptr(offs, C.c_int64) if offs is not None else None,
            rb, re, device, C.byref(out)))
        return cls(out, size)

    @classmethod
    def from_device_pointers(cls, size, nnz, d_rows, d_cols, d_vals, chromosomes=None,
                             row_block=None, device=0):
        """Cells already resident on the device (int32 arrays, e.g. torch tensors'
        ``data_ptr()``); the caller keeps ownership of those buffers."""
        n_chrom, offs = _offsets(size, chromosomes)
        rb, re = row_block if row_block is not None else (0, size)
        out = C.c_void_p()
        check(lib().tb_store_create_coo(
            size, nnz, d_rows, d_cols, d_vals, _capi.TB_MEM_DEVICE, n_chrom,
            ptr(offs, C.c_int64) if offs is not None else None, rb, re, device, C.byref(out)))
        return cls(out, size)

    @classmethod
    def synthetic(cls, size, params, chromosomes=None, row_block=None, device=0):
        n_chrom, offs = _offsets(size, chromosomes)
        rb, re = row_block if row_block is not None else (0, size)
        out = C.c_void_p()
        check(lib().tb_store_create_synthetic(
            size, n_chrom, ptr(offs, C.c_int64) if offs is not None else None,
            C.byref(params), rb, re, device, C.byref(out)))
        return cls(out, size)

    @staticmethod
    def synthetic_row_counts(size, params, chromosomes=None, row_block=None, device=0):
        n_chrom, offs = _offsets(size, chromosomes)
        rb, re = row_block if row_block is not None else (0, size)
        counts = np.zeros(re - rb, dtype=np.int64)
        check(lib().tb_synthetic_row_counts(
            size, n_chrom, ptr(offs, C.c_int64) if offs is not None else None,
            C.byref(params), rb, re, device, ptr(counts, C.c_int64)))
        return counts

    def close(self):
        if self._h is not None and self._h.value:
            lib().tb_store_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- info -----------------------------------------------------------------
    def info(self):
        v = [C.c_int64() for _ in range(6)]
        check(lib().tb_store_info(self._h, *[C.byref(x) for x in v]))
        keys = ("nbins", "row_begin", "row_end", "nnz_upper", "nnz_full", "device_bytes")
        return dict(zip(keys, [x.value for x in v]))

    def set_option(self, name, value):
        """e.g. ``set_option("rowsum", "csr")``: ICE passes walk the plain CSR, not the compact stream."""
        check(lib().tb_store_set_option(self._h, name.encode(), value.encode()))

    def layout(self):
        """Compact row-sum stream: payload bytes and chunks per encoding."""
        out = np.zeros(16, dtype=np.int64)
        check(lib().tb_store_layout(self._h, ptr(out, C.c_int64)))
        return dict(stream_bytes=int(out[0] + out[13]), chunk_bytes=int(out[0]),
                    tile_bytes=int(out[13]), chunks=int(out[7]), csr_bytes=int(out[8]),
                    tile_cells=int(out[9]), tile_cells_padded=int(out[10]),
                    tile_units=int(out[11]), overflow_cells=int(out[12]),
                    overflow_cells_in_tiles=int(out[14]),
                    encodings=dict(zip(("S32", "S16", "D32", "D16", "D8", "T"),
                                       out[1:7].tolist())))

    def export_upper(self):
        """Owned upper-triangle cells, row-major sorted: (rows, cols, vals) int32."""
        n = C.c_int64()
        check(lib().tb_store_export_upper(self._h, C.byref(n), None, None, None))
        r = np.empty(n.value, dtype=np.int32)
        c = np.empty(n.value, dtype=np.int32)
        v = np.empty(n.value, dtype=np.int32)
        if n.value:
            check(lib().tb_store_export_upper(self._h, C.byref(n), ptr(r, C.c_int32),
                                              ptr(c, C.c_int32), ptr(v, C.c_int32)))
        return r, c, v

    def export_csr(self, form="upper", pinned=None):
        """Owned rows in CSR form: (indptr int64[nrows+1], cols int32, vals int32).  ``pinned``: a
        callable ``(count, dtype) -> array`` that supplies the output buffers (e.g. pinned memory)."""
        f = {"upper": _capi.TB_CELLS_UPPER, "full_rows": _capi.TB_CELLS_FULL_ROWS}[form]
        n = C.c_int64()
        check(lib().tb_store_export_csr(self._h, f, C.byref(n), None, None, None))
        i = self.info()
        alloc = pinned or (lambda count, dtype: np.empty(count, dtype=dtype))
        indptr = alloc(i["row_end"] - i["row_begin"] + 1, np.int64)
        c = alloc(n.value, np.int32)
        v = alloc(n.value, np.int32)
        check(lib().tb_store_export_csr(self._h, f, C.byref(n), ptr(indptr, C.c_int64),
                                        ptr(c, C.c_int32), ptr(v, C.c_int32)))
        return indptr, c, v

    # -- multi-GPU ------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = np.zeros(128, dtype=np.uint8)
        check(lib().tb_comm_unique_id(ptr(buf, C.c_uint8)))
        return buf

    def attach(self, group):
        """Join a :class:`CommGroup` (a collective call: every rank attaches its shard)."""
        check(lib().tb_store_attach(self._h, group._h))
        self._group = group                      # keep it alive as long as the store

    def comm_init(self, rank, world, unique_id):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        check(lib().tb_comm_init(self._h, rank, world, ptr(uid, C.c_uint8)))

    # -- kernels ----------------------------------------------------------------
    def _mask(self, bad):
        if bad is None:
            return None, None
        m = np.ascontiguousarray(bad, dtype=np.uint8)
        if m.shape != (self.size,):
            raise ValueError("bad mask must have %d entries" % self.size)
        return m, ptr(m, C.c_uint8)

    def row_stats(self):
        nnz = np.zeros(self.size, dtype=np.int64)
        rsum = np.zeros(self.size, dtype=np.int64)
        check(lib().tb_row_stats(self._h, ptr(nnz, C.c_int64), ptr(rsum, C.c_int64)))
        return nnz, rsum

    def col_sums_masked(self, bad=None):
        m, mp = self._mask(bad)
        out = np.zeros(self.size, dtype=np.int64)
        check(lib().tb_col_sums_masked(self._h, mp, ptr(out, C.c_int64)))
        return out

    def ice(self, bad=None, iterations=0, max_dev=1e-5):
        """-> (bias float64[N], trace float64[passes,4] = Smin, meanS, Smax, dev)."""
        m, mp = self._mask(bad)
        bias = np.empty(self.size, dtype=np.float64)
        trace = np.zeros((iterations + 1, 4), dtype=np.float64)
        npass = C.c_int32(0)
        check(lib().tb_ice(self._h, mp, int(iterations), float(max_dev), ptr(bias, C.c_double),
                           ptr(trace, C.c_double), C.byref(npass)))
        ntrace = npass.value if iterations > 0 else 0
        return bias, trace[:ntrace]

    def probe_rowsum(self, reps=4):
        """Mean device ms of one K1 evaluation of this store (also on an unconnected shard):
        (total, sparse-tile kernel, dense kernels)."""
        ms = np.zeros(3, dtype=np.float64)
        check(lib().tb_probe_rowsum(self._h, int(reps), ptr(ms, C.c_double)))
        return tuple(ms.tolist())

    def probe_k1(self, reps=8):
        """Mean device ms of one K1 evaluation as the ICE loop issues it: (kernels back to back,
        dense kernel on a second stream next to the tile kernel)."""
        ms = np.zeros(2, dtype=np.float64)
        check(lib().tb_probe_k1(self._h, int(reps), ptr(ms, C.c_double)))
        return tuple(ms.tolist())

    def row_work(self):
        """Per owned row: payload bytes of its dense chunks, cells in sparse tiles."""
        i = self.info()
        n = i["row_end"] - i["row_begin"]
        dense = np.zeros(n, dtype=np.int64)
        tiles = np.zeros(n, dtype=np.int64)
        if n:
            check(lib().tb_row_work(self._h, ptr(dense, C.c_int64), ptr(tiles, C.c_int64)))
        return dense, tiles

    def ice_timing(self):
        a, b = C.c_double(), C.c_double()
        n, t = C.c_int32(), C.c_int32()
        check(lib().tb_ice_last_timing(self._h, C.byref(a), C.byref(b), C.byref(n), C.byref(t)))
        return dict(loop_ms=a.value, rowsum_ms=b.value, rowsum_launches=n.value,
                    total_launches=t.value)

    def timing(self):
        """Device times (ms) of the last calls: ICE loop, its K1 launches, the exchange + statistics
        + update between them, the last one-pass reduction (masked column sums / distance sums)."""
        out = np.zeros(8, dtype=np.float64)
        check(lib().tb_last_timing(self._h, ptr(out, C.c_double)))
        return dict(loop_ms=out[0], rowsum_ms=out[1], exchange_ms=out[2], kernel_ms=out[3],
                    rowsum_launches=int(out[4]), total_launches=int(out[5]),
                    exchange={2: "peer-mapped", 1: "local/nccl", 0: "not set up"}[int(out[6])],
                    csr_resident=bool(out[7]))

    def norm_sum(self, bias=None, bad=None):
        m, mp = self._mask(bad)
        bp = None
        if bias is not None:
            bias = np.ascontiguousarray(bias, dtype=np.float64)
            if bias.shape != (self.size,):
                raise ValueError("bias must have %d entries" % self.size)
            bp = ptr(bias, C.c_double)
        out = C.c_double()
        check(lib().tb_norm_sum(self._h, bp, mp, C.byref(out)))
        return out.value

    def _bias(self, bias):
        if bias is None:
            return None, None
        b = np.ascontiguousarray(bias, dtype=np.float64)
        if b.shape != (self.size,):
            raise ValueError("bias must have %d entries" % self.size)
        return b, ptr(b, C.c_double)

    def pair_stats(self, bias=None, bad=None, center=0.0):
        """Over the stored pairs i < j between good bins: (pairs with a non-zero value, smallest
        value, sum of log10(value), sum of (log10(value) - center)^2); value = count, or
        count / bias[i] / bias[j] when ``bias`` is given."""
        m, mp = self._mask(bad)
        b, bp = self._bias(bias)
        out = np.zeros(4, dtype=np.float64)
        check(lib().tb_pair_stats(self._h, bp, mp, float(center), ptr(out, C.c_double)))
        return int(out[0]), float(out[1]), float(out[2]), float(out[3])

    def pair_values(self, bias=None, bad=None, zscore=None):
        """The same pairs, row-major: (rows, cols, values).  ``zscore=(mean, std)`` turns the values
        into (log10(value) - mean) / std."""
        m, mp = self._mask(bad)
        b, bp = self._bias(bias)
        mean, std = zscore if zscore is not None else (0.0, 1.0)
        n = C.c_int64()
        check(lib().tb_pair_values(self._h, bp, mp, int(zscore is not None), float(mean), float(std),
                                   C.byref(n), None, None, None))
        r = np.empty(n.value, dtype=np.int32)
        c = np.empty(n.value, dtype=np.int32)
        v = np.empty(n.value, dtype=np.float64)
        if n.value:
            check(lib().tb_pair_values(self._h, bp, mp, int(zscore is not None), float(mean), float(std),
                                       C.byref(n), ptr(r, C.c_int32), ptr(c, C.c_int32),
                                       ptr(v, C.c_double)))
        return r, c, v

    def window(self, r0, r1, c0, c1, bias=None):
        """Dense float64 window rows [r0, r1) x columns [c0, c1) of the full symmetric matrix: counts,
        or count / bias[i] / bias[j]."""
        b, bp = self._bias(bias)
        out = np.zeros((max(r1 - r0, 0), max(c1 - c0, 0)), dtype=np.float64)
        if out.size:
            check(lib().tb_window(self._h, int(r0), int(r1), int(c0), int(c1), bp, ptr(out, C.c_double)))
        return out

    def region_cells(self, r0, r1, c0, c1, bad=None, bias=None, decay=None, half=False):
        """The stored cells inside rows [r0, r1) x columns [c0, c1) of the full symmetric matrix between
        good bins, row-major: (rows, cols, raw int32, normalised float64 | None, over-decay float64 |
        None).  ``decay``: float64[N] indexed by first bin of the section + distance (the layout of
        ``decay_sums``); cells between two sections get NaN.  ``half`` keeps the cells whose offset in
        the row range is <= their offset in the column range."""
        m, mp = self._mask(bad)
        b, bp = self._bias(bias)
        d, dp = self._bias(decay)
        if d is not None and b is None:
            raise ValueError("the decay needs the biases")
        n = C.c_int64()
        args = (self._h, int(r0), int(r1), int(c0), int(c1), mp, bp, dp, int(bool(half)), C.byref(n))
        check(lib().tb_region_cells(*(args + (None, None, None, None, None))))
        r = np.empty(n.value, dtype=np.int32)
        c = np.empty(n.value, dtype=np.int32)
        v = np.empty(n.value, dtype=np.int32)
        nrm = np.empty(n.value, dtype=np.float64) if b is not None else None
        dec = np.empty(n.value, dtype=np.float64) if d is not None else None
        if n.value:
            check(lib().tb_region_cells(*(args + (
                ptr(r, C.c_int32), ptr(c, C.c_int32), ptr(v, C.c_int32),
                ptr(nrm, C.c_double) if nrm is not None else None,
                ptr(dec, C.c_double) if dec is not None else None))))
        return r, c, v, nrm, dec

    def decay_sums(self, bias, bad=None):
        """Per section and distance: (normalised sums float64[N], raw sums int64[N]) indexed by
        first bin of the section + distance, and (cis, total) normalised off-diagonal sums."""
        m, mp = self._mask(bad)
        b, bp = self._bias(bias)
        nrm = np.zeros(self.size, dtype=np.float64)
        raw = np.zeros(self.size, dtype=np.int64)
        ct = np.zeros(2, dtype=np.float64)
        check(lib().tb_decay_sums(self._h, bp, mp, ptr(nrm, C.c_double), ptr(raw, C.c_int64),
                                  ptr(ct, C.c_double)))
        return nrm, raw, (float(ct[0]), float(ct[1]))

    # -- N4: compartments (tb_corr_*) ---------------------------------------------------------------
    def corr_build(self, beg, end, bias, bad, expected, stage=1):
        """Observed / expected matrix (``stage=0``) or its correlation matrix (``stage=1``) of the
        section ``[beg, end)`` over its good bins, left resident on the device; ``expected[d]`` for
        ``d < end - beg``.  Returns the order of the matrix (number of good bins)."""
        m, mp = self._mask(bad)
        b, bp = self._bias(bias)
        e = np.ascontiguousarray(expected, dtype=np.float64)
        if e.shape != (max(int(end) - int(beg), 0),):
            raise ValueError("expected must have one entry per distance of the section")
        n = C.c_int64()
        check(lib().tb_corr_build(self._h, int(beg), int(end), bp, mp, ptr(e, C.c_double), int(stage),
                                  C.byref(n)))
        self._corr_n = n.value
        return n.value

    def corr_get(self):
        """The resident matrix as float64[n, n]."""
        n = getattr(self, "_corr_n", 0)
        out = np.zeros((n, n), dtype=np.float64)
        if n:
            check(lib().tb_corr_get(self._h, ptr(out, C.c_double)))
        return out

    def corr_set(self, matrix):
        """Replace the resident matrix by a symmetric float64[n, n] from the host."""
        a = np.ascontiguousarray(matrix, dtype=np.float64)
        if a.ndim != 2 or a.shape[0] != a.shape[1] or a.shape[0] < 1:
            raise ValueError("corr_set needs a square matrix")
        check(lib().tb_corr_set(self._h, a.shape[0], ptr(a, C.c_double)))
        self._corr_n = a.shape[0]

    def corr_eig(self, k):
        """The ``k`` eigenpairs of largest magnitude of the resident matrix (``k >= n``: all of
        them): (eigenvalues descending float64[k'], eigenvectors float64[k', n], info dict)."""
        n = getattr(self, "_corr_n", 0)
        kk = max(min(int(k), n), 1)
        evals = np.zeros(kk, dtype=np.float64)
        evecs = np.zeros((kk, max(n, 1)), dtype=np.float64)
        info = np.zeros(8, dtype=np.int32)
        check(lib().tb_corr_eig(self._h, int(k), ptr(evals, C.c_double), ptr(evecs, C.c_double),
                                ptr(info, C.c_int32)))
        got = int(info[0])
        return evals[:got], evecs[:got, :n], dict(steps=int(info[1]), restarts=int(info[2]),
                                                  converged=bool(info[3]), round_trips=int(info[4]))

    def corr_timing(self):
        """Device ms of the last corr_build / corr_eig: cell visitor, centring, X X^T, scaling, eigen-solver."""
        out = np.zeros(8, dtype=np.float64)
        check(lib().tb_corr_timing(self._h, ptr(out, C.c_double)))
        return dict(visit_ms=out[0], center_ms=out[1], gemm_ms=out[2], scale_ms=out[3], eig_ms=out[4],
                    n=int(out[5]), stage=int(out[6]), gemm="simple" if out[7] else "dmma")

    def diag_sums(self, bad, ndiag):
        m, mp = self._mask(bad)
        out = np.zeros(max(int(ndiag), 0), dtype=np.int64)
        if ndiag > 0:
            check(lib().tb_diag_sums(self._h, mp, int(ndiag), ptr(out, C.c_int64)))
        return out


__all__ = ["ContactStore", "CommGroup", "SynthParams"]
