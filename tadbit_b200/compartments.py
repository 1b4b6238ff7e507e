"""A/B compartments: host layer of ``HiC_data.find_compartments`` and
``write_compartments`` (reference: _pytadbit/hic_data.py:736-1031, :1486-1511).

Per chromosome the reference builds the observed / expected matrix of the good bins as nested
Python lists (one dict lookup and three divisions per cell), calls ``numpy.corrcoef`` on it,
replaces NaN by 0 and asks ``scipy.sparse.linalg.eigsh`` for the ``max_ev`` leading
eigenvectors.  Here those three steps are device work on the resident contact store
(``tb_corr_build``: cell visitor + fp64 tensor-core ``X X^T``; ``tb_corr_eig``: Lanczos); this
module keeps what the reference does around them: the implicit filter / expected / one-pass ICE,
the break points at the sign changes of the chosen eigenvector, bad bins re-inserted as NaN, the
orientation by a "rich in A" track (Spearman), and the three output files.

Deliberate differences: plots (``savefig`` / ``show``) are not drawn; a
chromosome with fewer good bins than ``max_ev`` is skipped with the reference's "too small"
warning where the reference (with the scipy of this image) dies with an IndexError; the sign of
an eigenvector without ``rich_in_A`` is as arbitrary as ARPACK's (but reproducible).
"""
import gzip
import os
from warnings import warn

import numpy as np

from ._capi import bad_mask
from .normalize_hic import expected
from .vectors import BinVector


def read_bed_bins(fnam, resolution=1):
    """``{chromosome: {bin: summed value}}`` from a BED / BEDgraph / 3- or 2-column file, the
    structure ``pytadbit.parsers.bed_parser.parse_bed`` returns (bed_parser.py:41-104): column 5
    (BED) or 4 (BEDgraph) weighs an entry when it is a number, otherwise every entry counts 1; an
    entry goes to the bin of its END coordinate (``(beg + end - beg) // resolution``)."""
    opener = gzip.open if str(fnam).endswith('.gz') else open
    with opener(fnam, 'rt') as handle:
        lines = [l for l in handle
                 if not (l.startswith('#') or l.startswith('track') or l.startswith('browser'))]
    out = {}
    if not lines:
        return out
    first = lines[0]
    tabbed = first.split('\t', 5)
    if len(tabbed) == 6:                                  # classic BED: value in column 5 if numeric
        try:
            float(tabbed[4])
            kind = 'bed_value'
        except ValueError:
            kind = 'bed_one'
    elif len(tabbed) == 4:
        kind = 'bedgraph'
    elif len(first.split()) == 3:
        kind = 'three'
    else:
        kind = 'two'
    for line in lines:
        if kind == 'bed_value':
            crm, beg, end, _, val, _ = line.split('\t', 5)
            end, val = int(end), float(val)
        elif kind == 'bed_one':
            crm, beg, end, _ = line.split('\t', 3)
            end, val = int(end), 1
        elif kind == 'bedgraph':
            crm, beg, end, val = line.split()
            end, val = int(end), float(val)
        elif kind == 'three':
            crm, beg, end = line.split()
            end, val = int(end), 1
        else:
            crm, beg = line.split()
            end, val = int(beg), 1
        per = out.setdefault(crm, {})
        pos = end // resolution
        per[pos] = per.get(pos, 0) + val
    return out


def _section_expected(hic, sec, beg, end, good):
    """The expected count per distance the reference would use for this section, as an array
    over d in [0, end - beg): ``self.expected[sec][d]`` (per-chromosome decay of a biases pickle),
    falling back to the genome-wide ``self.expected[d]`` on KeyError (hic_data.py:851-867).
    Returns None when the reference ends with an empty matrix (empty per-chromosome entry)."""
    length = end - beg
    exp = hic.expected
    positions = np.flatnonzero(good)

    def needed_distances():
        """distances between good bins (what the nested loops would look up)"""
        if positions.size == 0:
            return np.zeros(0, dtype=np.int64)
        present = np.zeros(length, dtype=bool)
        # d is looked up iff two good bins lie d apart: autocorrelation support of the mask
        g = good.astype(np.float64)
        size = 1
        while size < 2 * length:
            size *= 2
        f = np.fft.rfft(g, size)
        ac = np.fft.irfft(f * np.conj(f), size)[:length]
        present = ac > 0.5
        return np.flatnonzero(present)

    def as_array(table):
        arr = np.full(length, np.nan)
        missing = []
        for d in needed_distances():
            try:
                arr[d] = table[int(d)]
            except (KeyError, IndexError):
                missing.append(int(d))
        return arr, missing

    try:
        per_section = exp[sec]
        use_section = True
    except (KeyError, TypeError, IndexError):
        use_section = False
    if use_section and not isinstance(per_section, (dict, list, tuple, np.ndarray, BinVector)):
        use_section = False                               # a flat table whose key happens to equal sec
    if use_section:
        arr, missing = as_array(per_section)
        if not missing:
            return arr
        if not per_section:                               # `sec in expected and not expected[sec]`
            return None
    arr, missing = as_array(exp)
    if missing:
        raise KeyError(missing[0])
    return arr


def find_compartments(hic, crms=None, savefig=None, savedata=None, savecorr=None, show=False,
                      suffix='', ev_index=None, rich_in_A=None, format='png', savedir=None,
                      max_ev=3, show_compartment_labels=False, smoothing_window=0, **kwargs):
    """Search for A/B compartments in each chromosome (hic_data.py:736-1031): the matrix is
    normalised by the expected count at each distance and by visibility (one pass of ICE unless
    biases are present), a correlation matrix is computed from it and the sign changes of its first
    eigenvector delimit the compartments.

    :param None crms: only these chromosomes
    :param None savedata: file for the compartment table (``write_compartments``)
    :param None savecorr: directory for one correlation matrix per chromosome
    :param None savedir: directory for one eigenvector table per chromosome
    :param None ev_index: per chromosome, the (1-based) eigenvector to cut by
    :param None rich_in_A: BED / BEDgraph path or ``{chromosome: {bin: value}}``: positive side of
       the eigenvector = bins rich in it (Spearman correlation)
    :param 3 max_ev: eigenvectors to compute
    :param 0 smoothing_window: ``size`` of a scipy median filter applied to the correlation matrix
    :returns: ``({chromosome: (eigenvalues, [eigenvectors])}, {chromosome: Spearman rho or None})``;
       ``hic.compartments`` holds the break points
    """
    verbose = kwargs.get('verbose', False)
    if not hic.bads:
        if verbose:
            print('Filtering bad columns %d' % 99)
        hic.filter_columns(perc_zero=kwargs.get('perc_zero', 99), by_mean=False, silent=True)
        if len(hic.bads) == len(hic):
            hic.bads = {}
            warn('WARNING: all columns would have been filtered out, '
                 'filtering disabled')
    if not hic.expected:
        if verbose:
            print('Normalizing by expected values')
        hic.expected = expected(hic, bads=hic.bads, **kwargs)
    if not hic.bias:
        if verbose:
            print('Normalizing by ICE (1 round)')
        hic.normalize_hic(iterations=0, silent=not verbose)
    if savefig or show:
        warn('WARNING: tadbit_b200 does not draw the compartment plots (savefig / show ignored)')
    for path in (savecorr, savedir):
        if path and not os.path.isdir(path):
            os.makedirs(path)
    if suffix != '':
        suffix = '_' + suffix
    if rich_in_A and isinstance(rich_in_A, str):
        rich_in_A = read_bed_bins(rich_in_A, resolution=hic.resolution)

    size = len(hic)
    mask = bad_mask(size, hic.bads)
    bias = hic._bias_array()
    cmprts, firsts, ev_nums = {}, {}, {}
    count = 0
    richA_stats = dict((sec, None) for sec in hic.section_pos)
    hic.compartment_timing = {}

    for sec in hic.section_pos:
        if crms and sec not in crms:
            continue
        if verbose:
            print('Processing chromosome', sec)
        beg, end = hic.section_pos[sec]
        good = mask[beg:end] == 0
        n = int(good.sum())
        exp_arr = _section_expected(hic, sec, beg, end, good) if n else None
        if exp_arr is None or n <= 1:
            # no good bin / empty decay entry, or numpy.corrcoef of a single row (a 0-d result the
            # reference cannot iterate over): hic_data.py:868-880
            warn('Chromosome %s is probably MT :)' % (sec))
            cmprts[sec] = []
            count += 1
            continue
        if np.any(exp_arr == 0):
            # float(x) / 0.0 in the reference's nested loops
            raise ZeroDivisionError('float division by zero')
        store = hic.store
        store.corr_build(beg, end, bias, mask, np.nan_to_num(exp_arr, nan=1.0), stage=1)
        if savecorr:
            _write_correlation(hic, sec, suffix, savecorr, store.corr_get())
        if smoothing_window:
            from scipy.ndimage import median_filter
            store.corr_set(median_filter(store.corr_get(), size=smoothing_window))
        k = max_ev if max_ev else (n - 1)
        wanted = max_ev if max_ev else (n - 1)
        if n < wanted or k < 1:
            warn('Chromosome %s too small to compute PC1' % (sec))
            cmprts[sec] = []                              # Y chromosome, or so...
            count += 1
            continue
        evals, evect, _ = store.corr_eig(k)
        hic.compartment_timing[sec] = store.corr_timing()
        n_first = [evect[i].tolist() for i in range(wanted)]
        ev_num = (ev_index[count] - 1) if ev_index else 0
        chosen = np.asarray(n_first[ev_num])
        flips = np.flatnonzero(chosen[1:] * chosen[:-1] < 0).tolist() + [len(chosen) - 1]
        breaks = [{'start': flips[i - 1] + 1 if i else 0, 'end': b} for i, b in enumerate(flips)]

        # bad columns back in as NaN, break points shifted accordingly (hic_data.py:949-966; the
        # inclusive upper bound is the reference's)
        bads = [b - beg for b in sorted(hic.bads) if beg <= b <= end]
        for vec in n_first:
            for b in bads:
                vec.insert(b, float('nan'))
        for b in bads:
            for brk in breaks:
                if brk['start'] >= b:
                    brk['start'] += 1
                    brk['end'] += 1
                else:
                    brk['end'] += brk['end'] > b
        bads = set(bads)

        richA_stats[sec] = None
        sign = 1
        if rich_in_A and sec in rich_in_A:
            from scipy.stats import spearmanr
            eves, gccs = [], []
            for i, v in enumerate(n_first[ev_num]):
                if i in bads:
                    continue
                try:
                    gc = rich_in_A[sec][i]
                except KeyError:
                    continue
                gccs.append(gc)
                eves.append(v)
            r_stat, richA_pval = spearmanr(eves, gccs)
            if verbose:
                print('  - Spearman correlation between "rich in A" and '
                      'Eigenvector:\n'
                      '      rho: %.7f p-val:%.7f' % (r_stat, richA_pval))
            richA_stats[sec] = r_stat
            sign = 1 if r_stat > 0 else -1
        for i in range(len(n_first)):
            n_first[i] = [sign * v for v in n_first[i]]
        ev_nums[sec] = ev_num + 1
        cmprts[sec] = breaks
        if rich_in_A:
            for cmprt in cmprts[sec]:
                try:
                    cmprt['dens'] = sum(rich_in_A.get(sec, {None: 0}).get(i, 0)
                                        for i in range(cmprt['start'], cmprt['end'] + 1)
                                        if i not in bads) / float(cmprt['end'] - cmprt['start'])
                except ZeroDivisionError:
                    cmprt['dens'] = float('nan')
                cmprt['type'] = 'A' if n_first[ev_num][cmprt['start']] > 0 else 'B'
        firsts[sec] = (evals, n_first)

    hic.compartments = cmprts
    if savedata:
        write_compartments(hic, savedata, chroms=list(hic.compartments.keys()), ev_nums=ev_nums)
    if savedir:
        ncrm = 0
        for sec in hic.section_pos:
            if crms and sec not in crms:
                continue
            if sec in firsts:
                with open(os.path.join(savedir, '%s_EigVect%d.tsv' % (
                        sec, ev_index[ncrm] if ev_index else 1)), 'w') as ev_file:
                    ev_file.write('# %s\n' % ('\t'.join(
                        'EV_%d (%.4f)' % (i, v) for i, v in enumerate(firsts[sec][0], 1))))
                    ev_file.write('\n'.join(['\t'.join([str(v) for v in vs])
                                             for vs in zip(*firsts[sec][1])]))
            ncrm += 1
    return firsts, richA_stats


def _write_correlation(hic, sec, suffix, savecorr, matrix):
    """One correlation matrix as text, filtered rows / columns as NaN (hic_data.py:883-913)."""
    start1, end1 = hic.section_pos[sec]
    rows = matrix.tolist()
    with open(os.path.join(savecorr, '%s_corr-matrix%s.tsv' % (sec, suffix)), 'w') as out:
        out.write('# MASKED %s\n' % (' '.join([str(k - start1) for k in hic.bads
                                               if start1 <= k <= end1])))
        rownam = ['%s\t%d-%d' % (k[0], k[1] * hic.resolution, (k[1] + 1) * hic.resolution)
                  for k in sorted(hic.sections, key=lambda x: hic.sections[x])
                  if k[0] == sec]
        length = end1 - start1
        empty = 'NaN\t' * (length - 1) + 'NaN\n'
        badrows = 0
        for row, posx in enumerate(range(start1, end1)):
            if posx in hic.bads:
                out.write(rownam.pop(0) + '\t' + empty)
                badrows += 1
                continue
            vals = []
            line = iter(rows[row - badrows])
            for posy in range(start1, end1):
                vals.append('NaN' if posy in hic.bads else str(next(line)))
            out.write(rownam.pop(0) + '\t' + '\t'.join(vals) + '\n')


def write_compartments(hic, savedata, chroms=None, ev_nums=None):
    """The compartment table (hic_data.py:1486-1511)."""
    with open(savedata, 'w') as out:
        sections = chroms if chroms else list(hic.compartments.keys())
        if ev_nums:
            for sec in sections:
                try:
                    out.write('## CHR %s\tEigenvector: %d\n' % (sec, ev_nums[sec]))
                except KeyError:
                    continue
        out.write('#%sstart\tend\trich in A\ttype\n' % (
            'CHR\t' if len(sections) > 1 else '\t'))
        for sec in sections:
            for c in hic.compartments[sec]:
                out.write('%s%d\t%d\t%.2f\t%s\n' % (
                    (str(sec) + '\t') if sections else '\t',
                    c['start'] + 1, c['end'] + 1,
                    c.get('dens', float('nan')), c.get('type', '')))
