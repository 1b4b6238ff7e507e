"""What ``tadbit normalize`` derives from the biases, on the GPU.

Mirrors the tail of ``read_bam`` in the reference CLI (reference:
_pytadbit/tools/tadbit_normalize.py:963-1107 and its helpers ``sum_nrm_matrix`` :1155-1159,
``get_cis_perc`` :1139-1152, ``sum_dec_matrix`` :1110-1136): rescale the biases so that the
normalised matrix averages ``factor`` per cell, the fraction of normalised interactions that stay
inside a chromosome, and the per-chromosome distance decay -- normalised counts per genomic
distance divided by the number of cells of that diagonal that do not touch a bad column, pooling
consecutive distances until more than 400 raw counts are gathered.  The ``decay`` dictionary is
what the CLI stores in the biases pickle next to the biases.

The reference streams per-region pickles of the BAM-derived dict through a process pool; here the
three reductions over the cells are ONE pass of a CUDA kernel over the resident contact store
(``tb_decay_sums``) plus the normalised sum (``tb_norm_sum``).  The cell counts of the diagonals
and the pooling loop are closed-form / O(N) host code with the reference's exact rules, including
its inclusive ``beg_chr <= b <= end_chr`` test.
"""
from collections import OrderedDict

import ctypes as C

import numpy as np

from ._capi import bad_mask, check, lib, ptr


def diagonal_cell_counts(section_pos, badcol):
    """``ndiags[crm][dist]`` of tadbit_normalize.py:1044-1072 without the loop over bad columns per
    distance: |A u B| = |A| + |B| - |A n B| with A = bad columns left of ``end - dist``, B = bad
    columns at or right of ``beg + dist`` shifted by ``dist``, and A n B = pairs of bad columns
    ``dist`` apart, i.e. the autocorrelation of the bad-column indicator of the chromosome."""
    try:
        bads = np.sort(np.fromiter(badcol, dtype=np.int64, count=len(badcol)))
    except (TypeError, ValueError):
        bads = np.array(sorted(int(b) for b in badcol), dtype=np.int64)
    out = OrderedDict()
    for crm, (beg, end) in section_pos.items():
        size = end - beg
        counts = np.zeros(max(size, 0), dtype=np.int64)
        if size <= 0:
            out[crm] = counts
            continue
        these = bads[(bads >= beg) & (bads <= end)]                # inclusive end, as the reference
        dist = np.arange(1, size)
        counts[1:] = size - dist
        if these.size and size > 1:
            n_a = np.searchsorted(these, end - dist, side='left')                  # b < maxp
            n_b = these.size - np.searchsorted(these, beg + dist, side='left')     # b >= minp
            inside = these[these < end] - beg
            ind = np.zeros(size, dtype=np.float64)
            ind[inside] = 1.0
            nfft = 1
            while nfft < 2 * size:
                nfft *= 2
            spec = np.fft.rfft(ind, nfft)
            auto = np.rint(np.fft.irfft(spec * np.conj(spec), nfft)[:size]).astype(np.int64)
            counts[1:] -= n_a + n_b - auto[1:]
        counts[0] = size - int(np.count_nonzero((these >= beg) & (these < end)))
        out[crm] = counts
    return out


def _pool_decay_loop(nrm, raw, ndiag, min_n):
    """The pooling loop of one chromosome (tadbit_normalize.py:1077-1106) over the per-distance
    normalised sums ``nrm``, raw sums ``raw`` (both indexed by distance, 0 where the reference's
    dict has no key) and cell counts ``ndiag``, distance by distance as the reference walks them:
    kept for inputs the group-wise form below does not take (negative or non-integer sums) and as
    what the tests compare it with."""
    size = len(ndiag)
    has = (np.asarray(raw) != 0)
    nrm_l, raw_l, nd_l = [float(v) for v in nrm], [int(v) for v in raw], [int(v) for v in ndiag]
    dec = dict((k, nrm_l[k]) for k in np.flatnonzero(has).tolist())
    tmpdec = 0
    tmpsum = 0
    ndg = 0
    val = 0
    previous = []
    for k in range(size):
        if has[k]:
            tmpdec += nrm_l[k]
            tmpsum += raw_l[k]
        else:
            tmpdec += 0.
            tmpsum += 0.
        previous.append(k)
        if tmpsum > min_n:
            ndg = sum(nd_l[q] for q in previous)
            val = tmpdec
            try:
                ratio = val / ndg
                for q in previous:
                    dec[q] = ratio
            except ZeroDivisionError:          # all columns at this distance are "bad"
                pass
            previous = []
            tmpdec = 0
            tmpsum = 0
    if len(previous) == size:
        dec = {}
    elif tmpsum < min_n:
        ndg += sum(nd_l[q] for q in previous)
        val += tmpdec
        try:
            ratio = val / ndg
            for q in previous:
                dec[q] = ratio
        except ZeroDivisionError:
            pass
    return dec


def pool_decay(nrm, raw, ndiag, min_n):
    """The pooling loop of one chromosome (tadbit_normalize.py:1077-1106) over the per-distance
    normalised sums ``nrm``, raw sums ``raw`` (both indexed by distance, 0 where the reference's
    dict has no key) and cell counts ``ndiag``.  Returns the ``{distance: value}`` dict.

    The walk itself runs in the library (``tb_pool_decay``, host C++: same additions in the same
    order, same divisions); the dictionary is put together here in the reference's key order -- the
    distances that hold reads first, then the ones the pooling added."""
    raw_a = np.asarray(raw)
    if raw_a.dtype.kind not in 'iu' or (raw_a.size and np.abs(raw_a).max() >= 2 ** 53):
        return _pool_decay_loop(nrm, raw, ndiag, min_n)            # never the case for counts
    size = len(ndiag)
    raw_a = np.ascontiguousarray(raw_a, dtype=np.int64)
    nrm_a = np.ascontiguousarray(nrm, dtype=np.float64)
    nd_a = np.ascontiguousarray(ndiag, dtype=np.int64)
    out = np.empty(size, dtype=np.float64)
    state = np.empty(size, dtype=np.uint8)
    empty = C.c_int32()
    check(lib().tb_pool_decay(size, ptr(nrm_a, C.c_double), ptr(raw_a, C.c_int64), ptr(nd_a, C.c_int64),
                              float(min_n), ptr(out, C.c_double), ptr(state, C.c_uint8), C.byref(empty)))
    if empty.value:
        return {}
    began, added = np.flatnonzero(state == 1), np.flatnonzero(state == 2)
    dec = dict(zip(began.tolist(), out[began].tolist()))
    dec.update(zip(added.tolist(), out[added].tolist()))
    return dec


def column_biases(hic_data, badcol, normalization='Vanilla'):
    """The biases ``tadbit normalize`` starts from (tadbit_normalize.py:897-925), before
    :func:`rescale_and_decay` rescales them: per bin the total number of interactions of its row
    (``cisprc[k][1]`` of the CLI, here ``tb_row_stats``; 1.0 for a bin without any), NaN on bad
    columns, then

    * ``'Vanilla'``: total / mean * sqrt(mean) (mean over the columns that are not NaN),
    * ``'SQRT'``: the same on the square roots of the totals,
    * ``'ICE'``: ``normalize_hic(iterations=100, max_dev=0.000001)`` with ``badcol`` as mask.

    Returns ``{bin: bias}`` (``'ICE'``: the ``bias`` mapping of the object, as the CLI copies it)."""
    size = len(hic_data)
    if normalization == 'ICE':
        hic_data.bads = badcol
        hic_data.normalize_hic(iterations=100, max_dev=0.000001)
        return hic_data.bias
    if normalization not in ('Vanilla', 'SQRT'):
        raise NotImplementedError('ERROR: method %s not implemented' % normalization)
    _, totals = hic_data.store.row_stats()
    biases = np.where(totals > 0, totals.astype(np.float64), 1.0)
    biases[bad_mask(size, badcol) != 0] = np.nan
    if normalization == 'SQRT':
        biases = biases ** 0.5
    mean_col = np.nanmean(biases)
    biases = biases / mean_col * mean_col ** 0.5
    return dict(enumerate(biases.tolist()))


def rescale_and_decay(hic_data, biases, badcol, factor=1, normalize_only=False,
                      signal_to_noise=0.05):
    """-> ``(biases, decay, badcol, norm_cisprc)`` as the tail of the reference's ``read_bam``
    returns them (its ``raw_cisprc`` comes from the read filters and is not part of this step).

    :param hic_data: :class:`tadbit_b200.HiC_data` with ``chromosomes`` set
    :param biases: mapping ``{bin: bias}`` over all bins (``nan`` marks columns without a bias), or
       an array
    :param badcol: dict of bad columns
    :param 1 factor: target mean of the normalised matrix
    """
    size = len(hic_data)
    store = hic_data.store
    if hasattr(biases, 'to_array'):
        b = biases.to_array().astype(np.float64)
    elif isinstance(biases, dict):
        b = np.fromiter(map(biases.__getitem__, range(size)), dtype=np.float64, count=size)
    else:
        b = np.asarray(biases, dtype=np.float64)
    nan = np.isnan(b)
    # sum_nrm_matrix: nansum over every cell of v / b_i / b_j -- only cells with a NaN bias drop out
    sumnrm = store.norm_sum(np.where(nan, 1.0, b), nan.astype(np.uint8))
    target = (sumnrm / float(size * size * factor)) ** 0.5
    b = b * target
    biases = dict(enumerate(b.tolist()))
    sections = hic_data.chromosomes or OrderedDict([(None, size)])
    section_pos = hic_data.section_pos or {None: (0, size)}
    mask = bad_mask(size, badcol)
    nrm, raw, (cis, total) = store.decay_sums(b, mask)              # one pass over the cells
    norm_cisprc = 0. if normalize_only else float(cis) / total
    ndiags = diagonal_cell_counts(OrderedDict((c, section_pos[c]) for c in sections), badcol)
    min_n = signal_to_noise ** -2.                                  # 400 with the default
    decay = {}
    for crm in sections:
        beg, end = section_pos[crm]
        decay[crm] = pool_decay(nrm[beg:end], raw[beg:end], ndiags[crm], min_n)
    return biases, decay, badcol, norm_cisprc


__all__ = ["rescale_and_decay", "column_biases", "diagonal_cell_counts", "pool_decay"]
