"""CPU ORACLE -- test infrastructure, NOT product code.

A numpy/scipy restatement of the TADbit Hi-C balancing hot path, written from the
behaviour of the reference (citations are ``path:line`` under the reference root):

* ``filter_by_zero_count``  <- _pytadbit/utils/hic_filtering.py:173-222
* ``filter_by_mean``        <- _pytadbit/utils/hic_filtering.py:19-170
* ``iterative`` (ICE)       <- _pytadbit/utils/normalize_hic.py:136-231
* ``expected``/``_meandiag``<- _pytadbit/utils/normalize_hic.py:234-291
* ``matrix_sum``            <- _pytadbit/hic_data.py:350-375
* ``filter_columns``        <- _pytadbit/hic_data.py:318-348
* ``normalize_hic``         <- _pytadbit/hic_data.py:380-413
* ``pair_values`` / ``pair_zscores`` / ``decay_sums`` / ``rescale_and_decay`` (N1, N2)
                            <- _pytadbit/experiment.py:737-799, tools/tadbit_normalize.py:963-1159
* ``region_cells`` (N3, extraction side)
                            <- _pytadbit/parsers/hic_bam_parser.py:755-872, :892-1242
* ``obs_exp_matrix`` / ``correlation_matrix`` / ``largest_eigenpairs`` / ``compartment_breaks`` (N4)
                            <- _pytadbit/hic_data.py:845-947 (find_compartments)

Parity is PINNED: ``oracle/gen_golden.py`` runs the unmodified reference (through
``oracle/ref_shim.py``) on the reference's own 12 chrT fixtures and on seeded
synthetic matrices, and ``tests/test_oracle_golden.py`` checks this module against
those committed vectors (``tests/golden/*.json``), including the values pinned by
the reference's own tests (test/test_all.py:332,400-402 feed from mask {22: 48}).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
``tadbit_b200`` never does.

Storage model: the reference keeps BOTH triangles in a dict keyed ``row*N+col``
(hic_data.py:54-118).  Here a matrix is the set of stored upper-triangle cells
(i <= j) as three arrays; every reference sum is re-expressed over that store
(SURVEY.md section 8 a', trap 1).
"""
from collections import OrderedDict
from sys import stderr

import numpy as np
from scipy import sparse


class OracleMatrix(object):
    """Upper-triangle store of a symmetric Hi-C matrix (stored cells only).

    ``chromosomes``: OrderedDict name -> number of bins, or None.  When None the
    matrix behaves as one section spanning [0, N) (hic_data.py:70-78 with
    ``sections == {}``, and the no-section branch normalize_hic.py:279-283).
    """

    def __init__(self, size, rows, cols, vals, chromosomes=None,
                 symmetricized=False):
        rows = np.asarray(rows, dtype=np.int64)
        cols = np.asarray(cols, dtype=np.int64)
        vals = np.asarray(vals)
        if np.any(rows > cols):
            raise ValueError("upper-triangle store needs row <= col")
        key = rows * size + cols
        order = np.argsort(key, kind="stable")
        key = key[order]
        if key.size and np.any(key[1:] == key[:-1]):
            raise ValueError("duplicate cells")
        self.size = int(size)
        self.rows = rows[order]
        self.cols = cols[order]
        self.vals = vals[order]
        self.chromosomes = chromosomes
        self.symmetricized = symmetricized
        if chromosomes:
            pos, total = OrderedDict(), 0
            for crm in chromosomes:
                pos[crm] = (total, total + chromosomes[crm])
                total += chromosomes[crm]
            self.section_pos = pos
        else:
            self.section_pos = OrderedDict([(None, (0, self.size))])

    # -- constructors -----------------------------------------------------
    @classmethod
    def from_full_dict(cls, dico, size, **kw):
        """From a reference-style dict ``{row*N+col: v}`` holding both triangles."""
        keys = np.fromiter(dico.keys(), dtype=np.int64, count=len(dico))
        vals = np.array([dico[k] for k in keys.tolist()])
        i, j = np.divmod(keys, size)
        keep = i <= j
        return cls(size, i[keep], j[keep], vals[keep], **kw)

    @classmethod
    def from_dense(cls, mat, **kw):
        mat = np.asarray(mat)
        iu = np.triu_indices(mat.shape[0])
        v = mat[iu]
        keep = v != 0
        return cls(mat.shape[0], iu[0][keep], iu[1][keep], v[keep], **kw)

    # -- helpers ----------------------------------------------------------
    def offdiag(self):
        return self.rows != self.cols

    def full_csr(self, keep=None, dtype=np.float64):
        """Both triangles (diagonal once) as scipy CSR, optionally masked."""
        r, c, v = self.rows, self.cols, self.vals.astype(dtype)
        if keep is not None:
            r, c, v = r[keep], c[keep], v[keep]
        od = r != c
        rr = np.concatenate([r, c[od]])
        cc = np.concatenate([c, r[od]])
        vv = np.concatenate([v, v[od]])
        return sparse.csr_matrix((vv, (rr, cc)), shape=(self.size, self.size))


def _bad_mask(size, bads):
    mask = np.zeros(size, dtype=bool)
    if bads:
        idx = np.fromiter((int(b) for b in bads), dtype=np.int64, count=len(bads))
        idx = idx[(idx >= 0) & (idx < size)]
        mask[idx] = True
    return mask


def row_stats(m):
    """(stored cells per full row, full row sums) -- exact integers.

    hic_filtering.py:187-190 (``cols[k // size] -= 1`` for every stored key) and
    :198-200 (``cols[k // size] += v``) over the both-triangle dict."""
    od = m.offdiag()
    nnz = np.bincount(m.rows, minlength=m.size).astype(np.int64)
    nnz += np.bincount(m.cols[od], minlength=m.size)
    if np.issubdtype(m.vals.dtype, np.integer):
        v = m.vals.astype(np.int64)
        rsum = np.zeros(m.size, dtype=np.int64)
        np.add.at(rsum, m.rows, v)
        np.add.at(rsum, m.cols[od], v[od])
    else:
        rsum = np.bincount(m.rows, weights=m.vals, minlength=m.size)
        rsum += np.bincount(m.cols[od], weights=m.vals[od], minlength=m.size)
    return nnz, rsum


def masked_col_sums(m, bad):
    """Column sums over rows NOT in ``bad`` (exact int64), for every column."""
    od = m.offdiag()
    v = m.vals.astype(np.int64)
    out = np.zeros(m.size, dtype=np.int64)
    k1 = ~bad[m.rows]                      # cell (i,j) adds to column j if row i good
    np.add.at(out, m.cols[k1], v[k1])
    k2 = od & ~bad[m.cols]                 # mirror cell (j,i) adds to column i if row j good
    np.add.at(out, m.rows[k2], v[k2])
    return out


# --------------------------------------------------------------------------
# filters
# --------------------------------------------------------------------------
def filter_by_zero_count(m, perc_zero, min_count=None, silent=True):
    """hic_filtering.py:173-222."""
    size = m.size
    nnz, rsum = row_stats(m)
    if min_count is None:
        min_val = int(size * float(perc_zero) / 100)
        flagged = (size - nnz) > min_val                       # :189-191,203
    else:
        if m.symmetricized:                                    # :193-197
            if not silent:
                stderr.write('\nWARNING: Using twice min_count as the matrix was '
                             'symmetricized and contains twice as many '
                             'interactions as the original\n')
            min_count *= 2
        flagged = rsum < min_count                             # :205
    return dict((int(i), True) for i in np.nonzero(flagged)[0])


def _r2_swapped(poly, counts, edges):
    # get_r2(p, x, y) is called with the counts as X and the bin edges as Y
    # (hic_filtering.py:19-22 vs :94).  Sequential python sums, as there.
    mean_y = np.mean(edges)
    sstot = sum([(edges[i] - mean_y) ** 2 for i in range(len(edges))])
    sserr = sum([(edges[i] - poly(counts[i])) ** 2 for i in range(len(edges))])
    return 1 - sserr / sstot


def _hundred_bins(cols, nbins):
    edges = np.linspace(min(cols), max(cols), nbins)
    which = np.digitize(cols, edges)
    counts = [sum(which == i) for i in range(1, nbins + 1)]
    return edges, counts


def _hundred_bins_sorted(cols, nbins):
    """_hundred_bins for ASCENDING ``cols`` without the 100 passes: digitize(x) == b iff
    edges[b-1] <= x < edges[b] (last bin open-ended), so the counts are differences of
    searchsorted positions.  Checked equal to _hundred_bins in tests/test_oracle_port.py."""
    if len(cols) == 0:
        raise ValueError('min() arg is an empty sequence')     # as min(cols) in _hundred_bins
    edges = np.linspace(cols[0], cols[-1], nbins)
    first = np.searchsorted(cols, edges, side='left')
    return edges, list(np.diff(np.append(first, len(cols))))


def mean_filter_cutoff(cols, nbins=100, bins=None):
    """The <=100-point numpy tail of filter_by_mean (hic_filtering.py:43-99).

    ``cols``: ascending column sums of the good sub-matrix.  Returns the chosen
    root, or raises ValueError("few") / LookupError("none") for the two failure
    exits (:164-167 and :153-156).  ``bins``: histogram builder (_hundred_bins, the literal
    restatement, unless the full-size path passes _hundred_bins_sorted)."""
    _hundred_bins = bins or globals()['_hundred_bins']
    percentile = np.percentile(cols, 5)
    edges, counts = _hundred_bins(cols, nbins)
    dropped = 0
    while list(counts).count(0) > len(counts) / 2:
        dropped += 1
        cols = cols[:-1]
        edges, counts = _hundred_bins(cols, nbins)
        if dropped > 10000:
            raise ValueError("few")
    best_r2, best_root = float('-inf'), None
    for order in range(6, 18):
        z = np.polyfit(edges, counts, order)
        dz = np.polyder(z, m=1)
        roots = np.roots(np.polyder(z))
        slope = np.polyval(dz, abs(roots[-2] - roots[-1]) / 2 + roots[-1])
        root = roots[-1] if slope > 0 else roots[-2]
        if root <= 0:
            continue
        if root >= percentile:
            continue
        r2 = _r2_swapped(np.poly1d(z), counts, edges)
        if order > 13:
            r2 -= float(order) / 30
        if best_r2 < r2:
            best_r2, best_root = r2, root
    if best_root is None:
        raise LookupError("none")
    return best_root


def filter_by_mean(m, bads=None, silent=True):
    """hic_filtering.py:25-170 (draw_hist=False path)."""
    if not bads:
        bads = {}
    bad = _bad_mask(m.size, bads)
    good_sums = masked_col_sums(m, bad)[~bad]                  # :36-39
    cols = np.sort(good_sums)
    try:
        try:
            root = mean_filter_cutoff(cols)
        except IndexError:
            if cols.size:
                raise
            return bads                                        # :45-47
        except LookupError:
            if not silent:
                stderr.write('WARNING: Too many zeroes to filter columns.'
                             ' SKIPPING...\n')
            return bads
        all_sums = masked_col_sums(m, np.zeros(m.size, dtype=bool))
        for i in np.nonzero(all_sums < root)[0]:               # :141-145
            bads[int(i)] = int(all_sums[i])
    except ValueError:
        if not silent:
            stderr.write('WARNING: Too few data to filter columns based on '
                         'mean value.\n')
    return bads


def filter_columns(m, perc_zero=99, by_mean=True, min_count=None, silent=True):
    """hic_data.py:318-348; returns the new ``bads`` dict."""
    bads = filter_by_zero_count(m, perc_zero, min_count=min_count, silent=silent)
    if by_mean:
        bads.update(filter_by_mean(m, bads=bads, silent=silent))
    return bads


# --------------------------------------------------------------------------
# ICE
# --------------------------------------------------------------------------
def iterative(m, bads=None, iterations=0, max_dev=0.00001):
    """normalize_hic.py:180-231 in the non-mutating form S = x * (W0 x), x = 1/B.

    Returns (bias[N] float64, trace) where trace rows are
    (Smin, meanS, Smax, it, dev) exactly as printed at :219-220."""
    size = m.size
    bad = _bad_mask(size, bads)
    keep = ~(bad[m.rows] | bad[m.cols])                        # copy_matrix :165-177
    w0 = m.full_csr(keep)
    present = np.zeros(size, dtype=bool)
    present[m.rows[keep]] = True
    present[m.cols[keep]] = True
    npresent = int(present.sum())
    if npresent == 0:
        raise ZeroDivisionError('ERROR: normalization failed, all bad columns')
    b = np.ones(size)
    x = np.where(present, 1.0, 0.0)
    trace = []
    mean_s = 0.0
    for it in range(iterations + 1):
        s = x * (w0 @ x)                                       # _update_S :136-143
        mean_s = s[present].sum() / npresent
        if mean_s == 0:                                        # float(S) / meanS in _updateDB :148
            raise ZeroDivisionError('float division by zero')
        db = s / mean_s                                        # _updateDB :146-151
        b[present] *= db[present]
        if iterations == 0:
            break
        with np.errstate(divide='ignore'):
            x = np.where(present & (b != 0), 1.0 / b, 0.0)     # _update_W :154-162
        smin, smax = s[present].min(), s[present].max()
        dev = max(abs(smin / mean_s - 1), abs(smax / mean_s - 1))
        trace.append((smin, mean_s, smax, it, dev))
        if dev < max_dev:
            break
    bias = np.ones(size)                                       # :223-230
    ok = present & (b != 0)
    bias[ok] = b[ok] * mean_s ** .5
    return bias, trace


_PORT = None


def _load_port():
    """ctypes handle of oracle/_build/libice_port.so (built by oracle/Makefile), or None."""
    global _PORT
    if _PORT is None:
        import ctypes
        import os
        import subprocess
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "_build", "libice_port.so")
        if not os.path.exists(path):
            try:
                subprocess.check_call(["make", "-s", "-C", here])
            except Exception:
                _PORT = False
                return None
        try:
            lib = ctypes.CDLL(path)
        except OSError:
            _PORT = False
            return None
        lib.ice_port.restype = ctypes.c_int
        lib.ice_port.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_int]
        lib.ice_port_threads.restype = ctypes.c_int
        _PORT = lib
    return _PORT or None


def synthetic_block(nbins, params, chromosomes=None):
    """Upper-triangle cells (rows, cols, vals; int32, row-major sorted) of the leading
    ``nbins`` x ``nbins`` block of the synthetic Hi-C model of SURVEY.md 8d, generated on the
    CPU by oracle/synth_port.c.  ``params`` is any object with the fields seed, cis_scale,
    cis_alpha, trans_mean, vis_sigma, low_frac, low_scale, empty_frac; ``chromosomes`` an
    ordered mapping name -> bins (the block is cut out of their concatenation).  Used by
    bench.py's CPU legs so that they need neither a GPU nor any product code."""
    import ctypes

    class _Params(ctypes.Structure):
        _fields_ = [("seed", ctypes.c_uint64)] + [(k, ctypes.c_double) for k in (
            "cis_scale", "cis_alpha", "trans_mean", "vis_sigma", "low_frac", "low_scale", "empty_frac")]

    lib = _load_port()
    if lib is None or not hasattr(lib, "synth_block"):
        raise RuntimeError("oracle/_build/libice_port.so is missing or stale: run `make -C oracle`")
    cp = _Params(int(params.seed), *[float(getattr(params, k[0])) for k in _Params._fields_[1:]])
    sizes = [int(v) for v in chromosomes.values()] if chromosomes else [int(nbins)]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    fn = lib.synth_block
    fn.restype = ctypes.c_int64
    fn.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    n = fn(int(nbins), len(sizes), offs.ctypes.data, ctypes.addressof(cp), None, None, None)
    if n < 0:
        raise MemoryError("synth_block")
    rows = np.empty(n, dtype=np.int32)
    cols = np.empty(n, dtype=np.int32)
    vals = np.empty(n, dtype=np.int32)
    got = fn(int(nbins), len(sizes), offs.ctypes.data, ctypes.addressof(cp), rows.ctypes.data,
             cols.ctypes.data, vals.ctypes.data)
    assert got == n
    return rows, cols, vals


def iterative_fast(m, bads=None, iterations=0, max_dev=0.00001, threads=None):
    """Same as :func:`iterative` with the pass loop in C + OpenMP (oracle/ice_port.c).

    Returns (bias, trace, cores, seconds spent in the loop).  Falls back to the numpy loop
    (cores = 1) when the C library cannot be built."""
    import time
    lib = _load_port()
    if lib is None:
        t0 = time.perf_counter()
        bias, trace = iterative(m, bads, iterations, max_dev)
        return bias, trace, 1, time.perf_counter() - t0
    size = m.size
    bad = _bad_mask(size, bads)
    keep = ~(bad[m.rows] | bad[m.cols])
    w0 = m.full_csr(keep)
    w0.sort_indices()
    indptr = np.ascontiguousarray(w0.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(w0.indices, dtype=np.int32)
    data = np.ascontiguousarray(w0.data, dtype=np.float64)
    bias = np.empty(size)
    trace = np.zeros((iterations + 1, 4))
    cores = int(threads) if threads else lib.ice_port_threads()
    t0 = time.perf_counter()
    passes = lib.ice_port(size, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data,
                          int(iterations), float(max_dev), bias.ctypes.data, trace.ctypes.data,
                          cores)
    dt = time.perf_counter() - t0
    if passes == -2:
        raise ZeroDivisionError('float division by zero')
    if passes < 0:
        raise ZeroDivisionError('ERROR: normalization failed, all bad columns')
    ntrace = passes if iterations > 0 else 0
    return bias, [(t[0], t[1], t[2], it, t[3]) for it, t in enumerate(trace[:ntrace])], cores, dt


def matrix_sum(m, bias=None, bads=None):
    """hic_data.py:350-375 -- both triangles, diagonal once."""
    bad = _bad_mask(m.size, bads)
    keep = ~(bad[m.rows] | bad[m.cols])
    r, c, v = m.rows[keep], m.cols[keep], m.vals[keep]
    w = np.where(r == c, 1, 2)
    if bias is None:
        return (v * w).sum()
    bias = np.asarray(bias, dtype=np.float64)
    return float(((v / (bias[r] * bias[c])) * w).sum())


def normalize_hic(m, bads=None, iterations=0, max_dev=0.1, sqrt=False, factor=1):
    """hic_data.py:380-413.  Returns (bias[N], trace, norm_sum or None)."""
    bias, trace = iterative(m, bads=bads, iterations=iterations, max_dev=max_dev)
    if sqrt:
        bias = bias ** 0.5
    norm_sum = None
    if factor:
        norm_sum = matrix_sum(m, bias, bads)
        target = (norm_sum / float(m.size * m.size * factor)) ** 0.5
        bias = bias * target
    return bias, trace, norm_sum


# --------------------------------------------------------------------------
# expected
# --------------------------------------------------------------------------
def diagonal_sums(m, bads, ndiag):
    """Per-distance integer sums and valid-cell counts of _meandiag's cells
    (normalize_hic.py:271-283): cell (i, i+d), i and i+d in the same section, only
    the row index tested against bads.  Returns (sum_d, cnt_d) for d in [0, ndiag)."""
    size = m.size
    bad = _bad_mask(size, bads)
    chrom_of = np.empty(size, dtype=np.int64)
    for n, (beg, end) in enumerate(m.section_pos.values()):
        chrom_of[beg:end] = n
    d = m.cols - m.rows
    keep = (~bad[m.rows]) & (chrom_of[m.rows] == chrom_of[m.cols]) & (d < ndiag)
    if np.issubdtype(m.vals.dtype, np.integer):
        sum_d = np.zeros(ndiag, dtype=np.int64)
        np.add.at(sum_d, d[keep], m.vals[keep].astype(np.int64))
    else:
        sum_d = np.bincount(d[keep], weights=m.vals[keep], minlength=ndiag)
    good_before = np.concatenate([[0], np.cumsum(~bad)])       # good rows in [0, k)
    cnt_d = np.zeros(ndiag, dtype=np.int64)
    for beg, end in m.section_pos.values():
        span = min(end - beg, ndiag)
        dd = np.arange(span)
        cnt_d[:span] += good_before[end - dd] - good_before[beg]
    return sum_d, cnt_d


def pool_diagonals(sum_d, cnt_d, size, min_n):
    """The pooling recursion of expected/_meandiag (normalize_hic.py:260-291), run
    iteratively over precomputed per-distance sums/counts."""
    expc = {}
    dist = 0
    ndiag = len(sum_d)
    while dist < size:
        tot, cnt, d = 0, 0, dist
        while True:
            if d < ndiag:
                tot += sum_d[d]
                cnt += cnt_d[d]
            if not cnt:
                new_dist, val = d + 1, 0.
                break
            if tot > min_n or d >= size:
                new_dist, val = d + 1, float(tot) / int(cnt)
                break
            d += 1
        for dist in range(dist, new_dist + 1):
            expc[dist] = val
    return expc


def expected(m, bads=None, signal_to_noise=0.05, inter_chrom=False):
    """normalize_hic.py:234-268."""
    min_n = signal_to_noise ** -2.
    size = m.size
    if m.chromosomes and not inter_chrom:
        size = max(m.chromosomes.values())
    sum_d, cnt_d = diagonal_sums(m, bads or {}, size)
    sum_d = [int(s) for s in sum_d] if sum_d.dtype.kind == 'i' else list(sum_d)
    return pool_diagonals(sum_d, cnt_d, size, min_n)


# --------------------------------------------------------------------------
# Full-size matrices: the same restatement over raw upper-triangle int32 arrays,
# loops in C + OpenMP (oracle/upper_port.c).  No sorting, no int64 copies: BASELINE
# configs[1..3] (3e8 .. 2e9 stored cells) fit the host this way.  Every function is checked
# against the numpy restatement above in tests/test_oracle_port.py.
# --------------------------------------------------------------------------
class UpperCells(object):
    """Stored upper-triangle cells (row <= col, each once) as contiguous int32 arrays."""

    def __init__(self, size, rows, cols, vals, chromosomes=None, symmetricized=False):
        self.size = int(size)
        self.rows = np.ascontiguousarray(rows, dtype=np.int32)
        self.cols = np.ascontiguousarray(cols, dtype=np.int32)
        self.vals = np.ascontiguousarray(vals, dtype=np.int32)
        self.chromosomes = chromosomes
        self.symmetricized = symmetricized
        if chromosomes:
            pos, total = OrderedDict(), 0
            for crm in chromosomes:
                pos[crm] = (total, total + chromosomes[crm])
                total += chromosomes[crm]
            self.section_pos = pos
        else:
            self.section_pos = OrderedDict([(None, (0, self.size))])

    def _args(self):
        return (self.rows.size, self.rows.ctypes.data, self.cols.ctypes.data, self.vals.ctypes.data)


def _upper_lib():
    import ctypes as C
    lib = _load_port()
    if lib is None or not hasattr(lib, "up_ice"):
        raise RuntimeError("oracle/_build/libice_port.so is missing or stale: run `make -C oracle`")
    if not getattr(lib, "_upper_ready", False):
        vp, i64 = C.c_void_p, C.c_int64
        lib.up_row_stats.argtypes = [i64, i64, vp, vp, vp, vp, vp, C.c_int]
        lib.up_masked_col_sums.argtypes = [i64, i64, vp, vp, vp, vp, vp, C.c_int]
        lib.up_diag_sums.argtypes = [i64, i64, vp, vp, vp, vp, vp, i64, vp, C.c_int]
        lib.up_norm_sum.argtypes = [i64, vp, vp, vp, vp, vp, C.c_int]
        lib.up_norm_sum.restype = C.c_double
        lib.up_ice.argtypes = [i64, i64, vp, vp, vp, vp, C.c_int, C.c_double, vp, vp, C.c_int]
        for f in (lib.up_row_stats, lib.up_masked_col_sums, lib.up_diag_sums, lib.up_ice):
            f.restype = C.c_int
        lib._upper_ready = True
    return lib


def _bad_u8(size, bads):
    return np.ascontiguousarray(_bad_mask(size, bads), dtype=np.uint8)


def fast_row_stats(m, threads=0):
    nnz = np.zeros(m.size, dtype=np.int64)
    rsum = np.zeros(m.size, dtype=np.int64)
    if _upper_lib().up_row_stats(m.size, *m._args(), nnz.ctypes.data, rsum.ctypes.data, threads):
        raise MemoryError("up_row_stats")
    return nnz, rsum


def fast_masked_col_sums(m, bad_u8, threads=0):
    out = np.zeros(m.size, dtype=np.int64)
    if _upper_lib().up_masked_col_sums(m.size, *m._args(), bad_u8.ctypes.data, out.ctypes.data, threads):
        raise MemoryError("up_masked_col_sums")
    return out


def fast_filter_columns(m, perc_zero=99, by_mean=True, min_count=None, silent=True):
    """hic_data.py:318-348 = filter_by_zero_count (hic_filtering.py:173-222) then filter_by_mean
    (:25-170); the integer sums come from C, the <=100-point tail is mean_filter_cutoff above."""
    size = m.size
    nnz, rsum = fast_row_stats(m)
    if min_count is None:
        flagged = (size - nnz) > int(size * float(perc_zero) / 100)
    else:
        if m.symmetricized:
            min_count *= 2
        flagged = rsum < min_count
    bads = dict((int(i), True) for i in np.nonzero(flagged)[0])
    if not by_mean:
        return bads
    bad = _bad_u8(size, bads)
    cols = np.sort(fast_masked_col_sums(m, bad)[bad == 0])
    try:
        try:
            root = mean_filter_cutoff(cols, bins=_hundred_bins_sorted)
        except IndexError:
            if cols.size:
                raise
            return bads
        except LookupError:
            if not silent:
                stderr.write('WARNING: Too many zeroes to filter columns. SKIPPING...\n')
            return bads
        all_sums = fast_masked_col_sums(m, np.zeros(size, dtype=np.uint8))
        for i in np.nonzero(all_sums < root)[0]:
            bads[int(i)] = int(all_sums[i])
    except ValueError:
        if not silent:
            stderr.write('WARNING: Too few data to filter columns based on mean value.\n')
    return bads


def fast_iterative(m, bads=None, iterations=0, max_dev=0.00001, threads=0):
    """normalize_hic.py:180-231 -> (bias, trace rows (Smin, meanS, Smax, it, dev), seconds)."""
    import time
    bias = np.empty(m.size)
    trace = np.zeros((iterations + 1, 4))
    bad = _bad_u8(m.size, bads)
    t0 = time.perf_counter()
    passes = _upper_lib().up_ice(m.size, *m._args(), bad.ctypes.data, int(iterations), float(max_dev),
                                 bias.ctypes.data, trace.ctypes.data, threads)
    dt = time.perf_counter() - t0
    if passes == -1:
        raise ZeroDivisionError('ERROR: normalization failed, all bad columns')
    if passes == -2:
        raise ZeroDivisionError('float division by zero')
    if passes < 0:
        raise MemoryError("up_ice")
    ntrace = passes if iterations > 0 else 0
    return bias, [(t[0], t[1], t[2], it, t[3]) for it, t in enumerate(trace[:ntrace])], dt


def fast_norm_sum(m, bias, bads=None, threads=0):
    bias = np.ascontiguousarray(bias, dtype=np.float64)
    bad = _bad_u8(m.size, bads)
    return _upper_lib().up_norm_sum(*m._args(), bad.ctypes.data, bias.ctypes.data, threads)


def fast_normalize_hic(m, bads=None, iterations=0, max_dev=0.1, sqrt=False, factor=1):
    """hic_data.py:380-413 -> (bias, trace)."""
    bias, trace, _ = fast_iterative(m, bads, iterations, max_dev)
    if sqrt:
        bias = bias ** 0.5
    if factor:
        target = (fast_norm_sum(m, bias, bads) / float(m.size * m.size * factor)) ** 0.5
        bias = bias * target
    return bias, trace


def fast_expected(m, bads=None, signal_to_noise=0.05, inter_chrom=False, threads=0):
    """normalize_hic.py:234-291: integer per-distance sums in C, counts and pooling as above."""
    min_n = signal_to_noise ** -2.
    size = m.size
    if m.chromosomes and not inter_chrom:
        size = max(m.chromosomes.values())
    bad = _bad_u8(m.size, bads or {})
    chrom_of = np.empty(m.size, dtype=np.int32)
    for n, (beg, end) in enumerate(m.section_pos.values()):
        chrom_of[beg:end] = n
    sum_d = np.zeros(size, dtype=np.int64)
    if _upper_lib().up_diag_sums(m.size, *m._args(), bad.ctypes.data, chrom_of.ctypes.data, size,
                                 sum_d.ctypes.data, threads):
        raise MemoryError("up_diag_sums")
    good_before = np.concatenate([[0], np.cumsum(bad == 0)])
    cnt_d = np.zeros(size, dtype=np.int64)
    for beg, end in m.section_pos.values():
        span = min(end - beg, size)
        dd = np.arange(span)
        cnt_d[:span] += good_before[end - dd] - good_before[beg]
    return pool_diagonals([int(v) for v in sum_d], cnt_d, size, min_n)


# --------------------------------------------------------------------------
# N1: interaction pairs, normalised values and z-scores
# (_pytadbit/experiment.py:737-743, :751-799; _pytadbit/utils/tadmaths.py:94-172)
# --------------------------------------------------------------------------
def pair_values(m, bias=None, bads=None):
    """The stored pairs i < j between columns that are not bad, row-major:
    (rows, cols, values) with values = count, or float(count) / bias[i] / bias[j]
    (experiment.py:739-741) when ``bias`` is given.  Pairs whose value is 0 are left out
    (``remove_zeros``, experiment.py:774-777)."""
    bad = _bad_mask(m.size, bads)
    keep = (m.rows < m.cols) & ~bad[m.rows] & ~bad[m.cols] & (m.vals != 0)
    r, c, v = m.rows[keep], m.cols[keep], m.vals[keep]
    if bias is not None:
        bias = np.asarray(bias, dtype=np.float64)
        v = v.astype(np.float64) / bias[r] / bias[c]
    return r, c, v


def pair_zscores(m, bias=None, bads=None):
    """tadmaths.zscore over those pairs (:142-172): log10, then (x - mean) / std with numpy's
    population standard deviation.  Returns (rows, cols, z, (n, min, mean, std))."""
    r, c, v = pair_values(m, bias, bads)
    if v.size == 0:
        raise ValueError('min() arg is an empty sequence')            # nozero_log :96
    logs = np.log10(v.astype(np.float64))
    mean_v, std_v = np.mean(logs), np.std(logs)
    return r, c, (logs - mean_v) / std_v, (v.size, float(v.min()), float(mean_v), float(std_v))


# --------------------------------------------------------------------------
# N3 (extraction side): get_matrix / write_matrix of the BAM parser with the matrix in the place of
# the BAM file (_pytadbit/parsers/hic_bam_parser.py:755-872, :892-1242)
# --------------------------------------------------------------------------
def region_cells(m, r0, r1, c0, c1, bads=None, bias=None, decay=None, half=False):
    """The stored cells (both triangles) inside rows [r0, r1) x columns [c0, c1), row-major, between
    bins that are not bad (`if i not in bads1 and j not in bads2`, :860, :1205):
    (rows, cols, raw, nrm | None, dec | None) with global bin indices,
    nrm = v / bias[i] / bias[j] (transform_value_norm, :828) and
    dec = nrm / decay[crm][|i - j|] (transform_value_decay, :830-833; ``decay`` = {crm: {dist: x}});
    NaN where the cell lies between two chromosomes or the distance has no entry.
    ``half``: what _read_half_bam_frag keeps (:447-458), offset in rows <= offset in columns."""
    bad = _bad_mask(m.size, bads)
    od = m.rows != m.cols
    rr = np.concatenate([m.rows, m.cols[od]])
    cc = np.concatenate([m.cols, m.rows[od]])
    vv = np.concatenate([m.vals, m.vals[od]])
    keep = (rr >= r0) & (rr < r1) & (cc >= c0) & (cc < c1) & ~bad[rr] & ~bad[cc] & (vv != 0)
    if half:
        keep &= (rr - r0) <= (cc - c0)
    rr, cc, vv = rr[keep], cc[keep], vv[keep]
    order = np.lexsort((cc, rr))
    rr, cc, vv = rr[order], cc[order], vv[order]
    nrm = dec = None
    if bias is not None:
        b = np.asarray(bias, dtype=np.float64)
        nrm = vv.astype(np.float64) / b[rr] / b[cc]
    if decay is not None:
        dec = np.full(rr.size, np.nan)
        for crm, (beg, end) in m.section_pos.items():
            inside = (rr >= beg) & (rr < end) & (cc >= beg) & (cc < end)
            table = decay.get(crm, {})
            for k in np.flatnonzero(inside):
                d = abs(int(rr[k]) - int(cc[k]))
                if d in table:
                    dec[k] = nrm[k] / table[d]
    return rr, cc, vv, nrm, dec


# --------------------------------------------------------------------------
# N2: rescaled biases, cis percentage and per-chromosome decay of `tadbit normalize`
# (_pytadbit/tools/tadbit_normalize.py:963-1107, helpers :1110-1159)
# --------------------------------------------------------------------------
def decay_sums(m, bias, bads=None):
    """sum_dec_matrix (:1110-1136) and get_cis_perc (:1139-1152) over the upper store: per section
    and distance the sums of count / bias_i / bias_j and of count over cells between good bins
    (arrays indexed first bin of the section + distance), and (cis, total) normalised sums of the
    off-diagonal pairs."""
    size = m.size
    bias = np.asarray(bias, dtype=np.float64)
    bad = _bad_mask(size, bads) | np.isnan(bias)
    chrom_of = np.empty(size, dtype=np.int64)
    beg_of = np.empty(size, dtype=np.int64)
    for n, (beg, end) in enumerate(m.section_pos.values()):
        chrom_of[beg:end] = n
        beg_of[beg:end] = beg
    keep = ~bad[m.rows] & ~bad[m.cols]
    r, c, v = m.rows[keep], m.cols[keep], m.vals[keep]
    x = v.astype(np.float64) / bias[r] / bias[c]
    same = chrom_of[r] == chrom_of[c]
    off = r != c
    total = float(np.sum(x[off]))
    cis = float(np.sum(x[off & same]))
    nrm = np.zeros(size)
    raw = np.zeros(size, dtype=np.int64)
    idx = beg_of[r[same]] + (c[same] - r[same])
    np.add.at(nrm, idx, x[same])
    np.add.at(raw, idx, v[same].astype(np.int64))
    return nrm, raw, (cis, total)


def decay_cell_counts(section_pos, badcol):
    """ndiags (:1044-1072), literally: per chromosome and distance the cells of the diagonal minus
    the ones that touch a bad column."""
    out = OrderedDict()
    for crm, (beg_chr, end_chr) in section_pos.items():
        chr_size = end_chr - beg_chr
        nd = dict((k, 0) for k in range(chr_size))
        thesebads = [b for b in badcol if beg_chr <= b <= end_chr]
        for dist in range(1, chr_size):
            nd[dist] += chr_size - dist
            bad_diag = set()
            maxp = end_chr - dist
            minp = beg_chr + dist
            for b in thesebads:
                if b < maxp:
                    bad_diag.add(b)
                if b >= minp:
                    bad_diag.add(b - dist)
            nd[dist] -= len(bad_diag)
        if chr_size:
            nd[0] += chr_size - sum(beg_chr <= b < end_chr for b in thesebads)
        out[crm] = nd
    return out


def rescale_and_decay(m, biases, badcol, factor=1, signal_to_noise=0.05):
    """The tail of read_bam (:963-1107) -> (biases array, decay {crm: {dist: value}}, norm_cisprc)."""
    size = m.size
    b = np.asarray(biases, dtype=np.float64)
    nan = np.isnan(b)
    w = np.where(m.rows == m.cols, 1.0, 2.0)                           # both triangles, diagonal once
    with np.errstate(invalid='ignore'):
        sumnrm = np.nansum(w * (m.vals.astype(np.float64) / b[m.rows] / b[m.cols]))   # sum_nrm_matrix
    target = (sumnrm / float(size * size * factor)) ** 0.5
    b = b * target
    nrm, raw, (cis, total) = decay_sums(m, b, badcol)
    norm_cisprc = float(cis) / total
    ndiags = decay_cell_counts(m.section_pos, badcol)
    min_n = signal_to_noise ** -2.
    decay = {}
    for crm, (beg, end) in m.section_pos.items():
        dec = dict((k, float(nrm[beg + k])) for k in range(end - beg) if raw[beg + k] != 0)
        rawd = dict((k, int(raw[beg + k])) for k in range(end - beg) if raw[beg + k] != 0)
        tmpdec = 0
        tmpsum = 0
        ndiag = 0
        val = 0
        previous = []
        for k in ndiags[crm]:
            tmpdec += dec.get(k, 0.)
            tmpsum += rawd.get(k, 0.)
            previous.append(k)
            if tmpsum > min_n:
                ndiag = sum(ndiags[crm][q] for q in previous)
                val = tmpdec
                try:
                    ratio = val / ndiag
                    for q in previous:
                        dec[q] = ratio
                except ZeroDivisionError:
                    pass
                previous = []
                tmpdec = 0
                tmpsum = 0
        if len(previous) == len(ndiags[crm]):
            dec = {}
        elif tmpsum < min_n:
            ndiag += sum(ndiags[crm][q] for q in previous)
            val += tmpdec
            try:
                ratio = val / ndiag
                for q in previous:
                    dec[q] = ratio
            except ZeroDivisionError:
                pass
        decay[crm] = dec
    return b, decay, norm_cisprc


# ---------------------------------------------------------------------------------------------
# N4 -- compartments: observed / expected, correlation, leading eigenpairs
#        <- _pytadbit/hic_data.py:845-937 (HiC_data.find_compartments)
# ---------------------------------------------------------------------------------------------
def obs_exp_matrix(m, beg, end, bias, bads, expected_by_distance):
    """The nested lists of hic_data.py:851-871 as an array over the good bins of [beg, end):
    cell (a, b) = float(count) / expected[|j - i|] / bias[i] / bias[j] with i <= j the two bins
    (the reference computes every cell both ways and keeps the one of its lower triangle, i.e.
    the smaller bin divides first)."""
    bad = _bad_mask(m.size, bads)
    idx = np.array([i for i in range(beg, end) if not bad[i]], dtype=np.int64)
    n = idx.size
    if n == 0:
        return np.zeros((0, 0))
    full = m.full_csr(dtype=np.float64)[idx][:, idx].toarray()
    b = np.asarray(bias, dtype=np.float64)
    e = np.asarray(expected_by_distance, dtype=np.float64)
    lo = np.minimum(idx[:, None], idx[None, :])
    hi = np.maximum(idx[:, None], idx[None, :])
    with np.errstate(divide="ignore", invalid="ignore"):
        return full / e[hi - lo] / b[lo] / b[hi]


def correlation_matrix(matrix):
    """numpy.corrcoef of the rows, NaN -> 0 (hic_data.py:873, :881)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.corrcoef(matrix)
    return np.where(np.isnan(c), 0.0, c)


def largest_eigenpairs(matrix, k):
    """What the reference reads out of scipy's eigsh(matrix, k) (hic_data.py:926-941): the k
    eigenpairs of largest magnitude, taken back to front from ARPACK's ascending algebraic order
    -> descending eigenvalues; k >= n: every eigenpair (scipy switches to a dense solver).
    Computed with the dense LAPACK solver; eigenvectors as rows, sign arbitrary."""
    w, v = np.linalg.eigh(np.asarray(matrix, dtype=np.float64))
    n = w.size
    if k >= n:
        order = np.argsort(-w, kind="stable")
    else:
        order = np.argsort(-np.abs(w), kind="stable")[:k]
        order = order[np.argsort(-w[order], kind="stable")]
    return w[order], v[:, order].T.copy()


def compartment_breaks(vector):
    """Break points at the sign changes of an eigenvector over the good bins (hic_data.py:942-947)."""
    v = np.asarray(vector, dtype=np.float64)
    flips = np.flatnonzero(v[1:] * v[:-1] < 0).tolist() + [len(v) - 1]
    return [{"start": flips[i - 1] + 1 if i else 0, "end": b} for i, b in enumerate(flips)]
