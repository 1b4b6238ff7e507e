"""ICE and distance-decay with the reference's function signatures, run on the GPU.

Drop-in for ``pytadbit.utils.normalize_hic.iterative`` / ``expected``
(reference: _pytadbit/utils/normalize_hic.py:180-231 and :234-291).
"""
import ctypes as C

import numpy as np

from ._capi import bad_mask, check, lib, ptr
from .vectors import BinVector

TRACE_FORMAT = '   %15.3f %15.3f %15.3f %4s %9.5f'          # normalize_hic.py:220


def iterative(hic_data, bads=None, iterations=0, max_dev=0.00001,
              verbose=False, **kwargs):
    """Iterative correction (Imakaev 2012) on the device-resident contact store.

    :param hic_data: :class:`tadbit_b200.HiC_data`
    :param None bads: dictionary with columns not to be considered
    :param 0 iterations: at most ``iterations + 1`` passes; 0 is a single pass
    :param 0.00001 max_dev: stop once max(|Smin/meanS-1|, |Smax/meanS-1|) < max_dev
    :returns: biases, a mapping ``{bin: float}`` over all bins (array-backed)
    """
    if verbose:
        print('iterative correction')
        print('  - copying matrix')
    size = len(hic_data)
    try:
        bias, trace = hic_data.store.ice(bad_mask(size, bads), iterations=iterations,
                                         max_dev=max_dev)             # K1 + K2 (device)
    except ZeroDivisionError as exc:
        # "all bad columns" is raised before the third line is printed (normalize_hic.py:203-210),
        # a zero mean row sum after it (_updateDB, :148)
        if verbose and 'all bad columns' not in str(exc):
            print('  - computing biases')
        raise
    if verbose:
        print('  - computing biases')
        for it, (smin, mean_s, smax, dev) in enumerate(trace):
            print(TRACE_FORMAT % (smin, mean_s, smax, it, dev))
    hic_data.last_trace = trace
    return BinVector(bias)


def diagonal_cell_counts(section_pos, size, bad, ndiag):
    """Number of cells (i, i+d) that _meandiag visits for each d: i and i+d in the same
    section and i not bad (normalize_hic.py:272-278).  Closed form from a prefix sum."""
    good_before = np.concatenate([[0], np.cumsum(bad == 0)])
    counts = np.zeros(ndiag, dtype=np.int64)
    for beg, end in section_pos.values():
        span = min(end - beg, ndiag)
        if span > 0:
            d = np.arange(span)
            counts[:span] += good_before[end - d] - good_before[beg]
    return counts


def expected(hic_data, bads=None, signal_to_noise=0.05, inter_chrom=False, **kwargs):
    """Mean observed count per genomic distance, pooling consecutive distances until
    more than ``signal_to_noise ** -2`` reads are gathered.

    :returns: dict ``{distance: float}``
    """
    min_n = signal_to_noise ** -2.                     # 400 by default
    size = len(hic_data)
    if hic_data.chromosomes and not inter_chrom:
        size = max(hic_data.chromosomes.values())
    bad = bad_mask(len(hic_data), bads)
    ndiag = size
    sum_d = np.ascontiguousarray(hic_data.store.diag_sums(bad, ndiag), dtype=np.int64)   # K5 (device)
    sections = hic_data.section_pos or {None: (0, len(hic_data))}   # normalize_hic.py:279-283
    cell_d = np.ascontiguousarray(diagonal_cell_counts(sections, len(hic_data), bad, ndiag), dtype=np.int64)
    if sum_d.size and np.abs(sum_d).max() >= 2 ** 53:              # float(total) would round
        return _pool_expected_loop(sum_d.tolist(), cell_d.tolist(), size, ndiag, min_n)
    vals = np.empty(size + 2, dtype=np.float64)
    nkeys = C.c_int64()
    check(lib().tb_pool_expected(size, ndiag, ptr(sum_d, C.c_int64), ptr(cell_d, C.c_int64), float(min_n),
                                 ptr(vals, C.c_double), C.byref(nkeys)))
    return dict(zip(range(nkeys.value), vals[:nkeys.value].tolist()))


def _pool_expected_loop(sums, cells, size, ndiag, min_n):
    """The pooling of ``expected`` distance by distance in Python (normalize_hic.py:262-291): what
    tb_pool_expected does, kept for sums beyond 2^53 and as what the tests compare it with."""
    expc = {}
    dist = 0
    while dist < size:                                  # normalize_hic.py:262-267
        total, ncell, stop = 0, 0, dist
        while True:                                     # _meandiag recursion, unrolled
            if stop < ndiag:
                total += sums[stop]
                ncell += cells[stop]
            if ncell == 0:
                val = 0.
                break
            if total > min_n or stop >= size:
                val = float(total) / ncell
                break
            stop += 1
        for dist in range(dist, stop + 2):
            expc[dist] = val
    return expc
