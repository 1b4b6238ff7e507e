"""Bad-column filters with the reference's signatures, computed from GPU statistics.

Drop-in for ``pytadbit.utils.hic_filtering.filter_by_zero_count`` /
``filter_by_mean`` (reference: _pytadbit/utils/hic_filtering.py:173-222 and :25-170).

The O(nnz) / O(N^2) parts -- stored cells and sums per row, column sums with bad rows
masked -- are exact int64 reductions on the device (K4, ``tb_row_stats`` /
``tb_col_sums_masked``).  What is left on the host is the reference's <=100-point
numpy tail (percentile, linspace, digitize, polyfit, roots): it is fed exact integers
and calls the same numpy routines in the same order, so the mask is bit-identical
("Third-party arithmetic on the path").
"""
from sys import stderr
from warnings import warn

import numpy as np

from ._capi import bad_mask

_NBINS = 100
_MAX_TRIMS = 10000


def _listing(indices):
    """The 20-per-line, 1-based column listing of the reference's warnings."""
    return ' '.join(['%5s' % str(i + 1) + ('' if (j + 1) % 20 else '\n')
                     for j, i in enumerate(sorted(indices))])


def filter_by_zero_count(matrx, perc_zero, min_count=None, silent=True):
    """Columns with too many empty cells (or too few counts when ``min_count`` is set).

    :param matrx: :class:`tadbit_b200.HiC_data`
    :param perc_zero: maximum percentage of cells with no stored value
    :param None min_count: minimum row sum; overrides ``perc_zero``
    :returns: dict ``{column index: True}``
    """
    size = len(matrx)
    stored, sums = matrx.store.row_stats()                     # K4 (device)
    if min_count is None:
        min_val = int(size * float(perc_zero) / 100)           # hic_filtering.py:191
        flagged = (size - stored) > min_val
    else:
        if matrx.symmetricized:                                # :193-197
            stderr.write('\nWARNING: Using twice min_count as the matrix was '
                         'symmetricized and contains twice as many '
                         'interactions as the original\n')
            min_count *= 2
        min_val = size - min_count
        flagged = sums < min_count                             # :205
    bads = dict.fromkeys(np.flatnonzero(flagged).tolist(), True)
    if bads and not silent:
        if min_count is None:
            stderr.write(('\nWARNING: removing columns having more than %s '
                          'zeroes:\n %s\n') % (min_val, _listing(bads)))
        else:
            stderr.write(('\nWARNING: removing columns having less than %s '
                          'counts:\n %s\n') % (int(size - min_val), _listing(bads)))
    return bads


class _NoCutoff(Exception):
    """No polynomial order gave an admissible root (reference: unpack error -> SKIPPING)."""


def _histogram(sums):
    """100-bin histogram of the ASCENDING column sums, as the reference builds it with
    ``np.linspace`` / ``np.digitize`` / one ``sum(hist == i)`` per bin (hic_filtering.py:49-54).

    Same integers, but from the sorted order: ``digitize`` puts ``x`` in bin ``b`` when
    ``edges[b-1] <= x < edges[b]`` (the last bin is open-ended), i.e. the bin counts are
    differences of ``searchsorted(..., 'left')`` positions -- O(100 log N) instead of 100 passes
    over N values, which matters because the trimming loop below may rebuild the histogram
    thousands of times on a genome-wide matrix."""
    if len(sums) == 0:
        raise ValueError('min() arg is an empty sequence')     # what min(cols) raises there
    edges = np.linspace(sums[0], sums[-1], _NBINS)
    first = np.searchsorted(sums, edges, side='left')          # values < edges[b]
    counts = np.diff(np.append(first, len(sums)))
    return edges, [c for c in counts]


def _swapped_r2(poly, counts, edges):
    # the reference calls get_r2(p, x, y): X = counts, Y = bin edges (hic_filtering.py:19-22,94)
    centre = np.mean(edges)
    total = sum([(edges[k] - centre) ** 2 for k in range(len(edges))])
    # Horner's rule is elementwise: evaluating all bins at once gives the values of 100 separate
    # calls; the squares are still added left to right as the reference's sum([...]) adds them
    fitted = poly(np.asarray(counts))
    resid = sum([(edges[k] - fitted[k]) ** 2 for k in range(len(edges))])
    return 1 - resid / total


def _cutoff(sums):
    """First admissible root of the derivative of the best polynomial fit to the
    100-bin histogram of the ascending column sums (hic_filtering.py:43-99)."""
    ceiling = np.percentile(sums, 5)
    edges, counts = _histogram(sums)
    trims = 0
    while list(counts).count(0) > len(counts) / 2:             # need half of the bins filled
        trims += 1
        sums = sums[:-1]
        edges, counts = _histogram(sums)
        if trims > _MAX_TRIMS:
            raise ValueError('too few data')
    champion = None
    champion_score = float('-inf')
    for order in range(6, 18):
        coef = np.polyfit(edges, counts, order)
        deriv = np.polyder(coef, m=1)
        roots = np.roots(np.polyder(coef))
        probe = abs(roots[-2] - roots[-1]) / 2 + roots[-1]
        root = roots[-1] if np.polyval(deriv, probe) > 0 else roots[-2]
        if root <= 0 or root >= ceiling:
            continue
        score = _swapped_r2(np.poly1d(coef), counts, edges)
        if order > 13:
            score -= float(order) / 30
        if champion_score < score:
            champion_score, champion = score, root
    if champion is None:
        raise _NoCutoff()
    return champion


def filter_by_mean(matrx, draw_hist=False, silent=False, bads=None, savefig=None):
    """Columns whose total count lies below the first local minimum of the fitted
    distribution of column sums.

    :param matrx: :class:`tadbit_b200.HiC_data`
    :param None bads: columns already flagged; the dict is extended in place and returned
    :returns: dict ``{column index: column sum}`` merged into ``bads``
    """
    if draw_hist or savefig:
        warn('tadbit_b200: histogram plotting is outside the accelerated path; ignored')
    if not bads:
        bads = {}
    size = len(matrx)
    mask = bad_mask(size, bads)
    good_sums = matrx.store.col_sums_masked(mask)[mask == 0]   # K4 (device), :36-39
    sums = np.sort(good_sums)
    if sums.size == 0:
        warn('WARNING: no columns to filter out')              # :45-47
        return bads
    try:
        try:
            root = _cutoff(sums)
        except _NoCutoff:
            if not silent:
                stderr.write('WARNING: Too many zeroes to filter columns.'
                             ' SKIPPING...\n')
            return bads
        totals = matrx.store.col_sums_masked(None)             # all rows, :141-145
        low = np.flatnonzero(totals < root)
        bads.update(zip(low.tolist(), totals[low].tolist()))
        if bads and not silent:
            try:
                message = ('\nWARNING: removing columns having less than %s '
                           'counts:\n %s\n') % (round(root, 3), _listing(bads.keys()))
            except TypeError:
                # numpy's roots may be complex and a complex number cannot be rounded: in the
                # reference the columns are already flagged by then and its bare except
                # prints this instead (hic_filtering.py:141-156)
                message = 'WARNING: Too many zeroes to filter columns. SKIPPING...\n'
            stderr.write(message)
    except ValueError:
        if not silent:
            stderr.write('WARNING: Too few data to filter columns based on '
                         'mean value.\n')
    return bads
