"""N3, extraction side: the reference's ``get_matrix`` / ``write_matrix`` of the BAM parser
(_pytadbit/parsers/hic_bam_parser.py:755-872, :892-1242) -- extracting a region of the matrix with
the biases and the decay applied, and the ``.abc`` files ``tadbit bin`` writes -- over a matrix
that is resident on the GPU instead of a BAM file.

The first argument is a :class:`tadbit_b200.HiC_data` where the reference takes the path of a BAM
file; everything else keeps the reference's names, meaning, defaults, file names, file text and
exceptions (including the ones that are accidents of the reference under Python 3; each is noted
where it is reproduced).  The cells come from ``tb_region_cells`` (one pass over the store: the
cells of the region between good bins, with ``v / bias1[a] / bias2[b]`` and
``... / decay[c][|a - b|]`` computed on the device).

Differences, all in what a BAM file is and a matrix is not:

* cells are visited row-major; the reference visits them in the order the reads lie in the BAM
  file, so the *lines* of a file are the same set in another order, and a ``KeyError`` for a
  missing decay entry names the first offending cell in row-major order;
* ``filter_exclude``, ``tmpdir``, ``ncpus``, ``nchunks``, ``clean`` are accepted and ignored (they
  steer the BAM reading); ``chr_order`` must be the order of the matrix; ``append_to_tar`` is not
  supported; ``cooler`` output needs h5py, which the reference also demands (same message);
* a matrix knows its chromosomes in bins: ``lengths`` (base pairs per chromosome, for the
  ``# CRM`` header lines) defaults to ``(bins - 1) * resolution``;
* a file that the reference leaves half written when it raises in the middle holds only its header
  here.
* a row range that runs past the end of ``region1``'s chromosome holds nothing there.  (The
  reference agrees unless ``region2`` is also given: its list of row bins is then built over both
  regions with ``region1``'s offsets, hic_bam_parser.py:604-616, and reads of ``region2``'s
  chromosome that fall into the overrun come back as extra rows numbered after the range -- an
  accident that is not reproduced.)
"""
import os
from collections import OrderedDict
from pickle import load

import numpy as np

try:
    basestring
except NameError:
    basestring = str


def nicer(res, sep=' ', comma='', allowed_decimals=0):
    """Resolution for human beings (_pytadbit/utils/extraviews.py:78-96)."""
    def fmt(x):
        return '{:,g}'.format(x).replace(',', comma)
    for power, unit in ((9, 'Gb'), (6, 'Mb'), (3, 'kb')):
        if res and not res % 10 ** (power - allowed_decimals):
            return fmt(res / 10. ** power) + sep + unit
    return fmt(res) + sep + 'b'


def _generate_name(regions, starts, ends, resolution, chr_order=None):
    """File name stem (hic_bam_parser.py:874-889)."""
    if len(regions) not in (1, 2):
        return 'full'
    name = []
    for i, region in enumerate(regions):
        try:
            name.append('%s:%d-%d' % (region, starts[i] // resolution, ends[i] // resolution))
        except TypeError:                                    # all of the chromosome
            name.append('%s' % (region))
    return '_'.join(name)


def get_biases_region(biases, bin_coords, check_resolution=None):
    """Biases, decay and bad bins of a pickle (or its dict), re-indexed to the region
    (hic_bam_parser.py:719-752)."""
    start_bin1, end_bin1, start_bin2, end_bin2 = bin_coords
    if isinstance(biases, basestring):
        biases = load(open(biases, 'rb'))
    resolution = biases.get('resolution', float('NaN'))
    if check_resolution is not None and check_resolution != resolution:
        raise ValueError('ERROR: resolution not matching with wanted '
                         'resolution (wanted: {} vs found: {})'.format(check_resolution, resolution))
    decay = biases.get('decay', {})

    def window(name, beg, end):
        return dict((k - beg, v) for k, v in biases.get(name, {}).items() if beg <= k < end)

    bias1, bads1 = window('biases', start_bin1, end_bin1), window('badcol', start_bin1, end_bin1)
    if start_bin1 != start_bin2:
        bias2, bads2 = window('biases', start_bin2, end_bin2), window('badcol', start_bin2, end_bin2)
    else:
        bias2, bads2 = bias1, bads1
    return bias1, bias2, decay, bads1, bads2


def cooler_weights_region(biases, bin_coords):
    """The row and column weights write_matrix(cooler=True, normalizations=('norm',)) hands to the
    cooler writer (hic_bam_parser.py:1228-1232): ``1 / bias`` where the bias is a positive number,
    else 0, bins counted from the start of each region."""
    bias1, bias2, _, _, _ = get_biases_region(biases, bin_coords)

    def weights(bias):
        return [1. / bias[b] if not bias[b] != bias[b] and bias[b] > 0 else 0 for b in range(len(bias))]

    return weights(bias1), weights(bias2)


class _Region(object):
    """What read_bam (hic_bam_parser.py:496-686) derives from the region arguments."""

    def __init__(self, hic, resolution, region1, start1, end1, region2, start2, end2, max_size, chr_order,
                 nchunks=100):
        if hic.chromosomes:
            self.sections = OrderedDict((c, int(n)) for c, n in hic.chromosomes.items())
        else:
            raise ValueError('the matrix needs its chromosomes (name -> number of bins)')
        if resolution != hic.resolution:
            raise ValueError('the matrix is binned at %s, not at %s' % (hic.resolution, resolution))
        if chr_order and [c for c in chr_order if c in self.sections] != list(self.sections):
            raise NotImplementedError('chr_order other than the order of the matrix')
        sections = self.sections
        section_pos, total = {}, 0
        for crm in sections:
            section_pos[crm] = (total, total + sections[crm])
            total += sections[crm]
        self.section_pos = section_pos
        if not total:
            raise Exception('ERROR: Chromosome %s smaller than bin size\n' % (list(sections) or [None])[-1])
        self.regions = list(sections)
        if region1:
            self.regions = [region1]
            if region2:
                self.regions.append(region2)
        elif start1 is not None or end1:
            raise Exception('ERROR: Cannot use start/end1 without region')
        if start1 is not None:
            start_bin1 = section_pos[region1][0] + start1 // resolution
        else:
            start_bin1 = section_pos[region1][0] if region1 else 0
        if end1 is not None:
            end_bin1 = section_pos[region1][0] + end1 // resolution
        else:
            end_bin1 = section_pos[region1][1] if region1 else total
        # the reference cuts [start_bin1, end_bin1) into chunks and looks their first bins up in the
        # list of all bins: an empty range, or one that runs past the end of the genome, dies there
        span = end_bin1 - start_bin1
        step = span // (min(span, nchunks) + 1) + 1
        if span <= 0 or start_bin1 + (span - 1) // step * step >= total:
            raise IndexError('list index out of range')
        for crm in self.regions:                              # the reference indexes both before it checks
            section_pos[crm]
        if region2:
            if region2 not in section_pos:
                raise Exception('ERROR: chromosome %s not found' % region2)
            beg2 = section_pos[region2][0]
            start_bin2 = beg2 + (start2 // resolution if start2 is not None else 0)
            end_bin2 = beg2 + end2 // resolution if end2 is not None else section_pos[region2][1]
        else:
            start_bin2, end_bin2 = start_bin1, end_bin1
        size1, size2 = end_bin1 - start_bin1, end_bin2 - start_bin2
        if max_size and max_size < size1 * size2:
            raise Exception(('ERROR: matrix too large ({0}x{1}) should be at most '
                             '{2}x{2}').format(size1, size2, int(max_size ** 0.5)))
        self.bin_coords = start_bin1, end_bin1, start_bin2, end_bin2
        # the bins reads can fall into: a range that runs past the end of its chromosome holds
        # nothing there (the reads of the next chromosome carry another name)
        clip1 = section_pos[region1][1] if region1 else total
        clip2 = section_pos[region2][1] if region2 else clip1
        self.rows = (start_bin1, max(start_bin1, min(end_bin1, clip1)))
        self.cols = (start_bin2, max(start_bin2, min(end_bin2, clip2)))


class _Cells(object):
    """The cells of a region and the vectors the transforms need."""

    def __init__(self, hic, reg, biases, want_norm, want_decay, half, touch_bias=False):
        size = len(hic)
        start_bin1, _, start_bin2, _ = reg.bin_coords
        self.bads1 = self.bads2 = {}
        bad = bias = decay = None
        self.have_bias = self.decay_table = None
        if biases:
            self.bias1, self.bias2, self.decay_table, self.bads1, self.bads2 = get_biases_region(
                biases, reg.bin_coords)
            bad = np.zeros(size, dtype=np.uint8)
            for bads, beg in ((self.bads1, start_bin1), (self.bads2, start_bin2)):
                if bads:
                    idx = np.fromiter(bads.keys(), dtype=np.int64, count=len(bads)) + beg
                    bad[idx[idx < size]] = 1
            if want_norm or want_decay or touch_bias:
                bias = np.full(size, np.nan)
                have = np.zeros(size, dtype=bool)
                for table, beg in ((self.bias1, start_bin1), (self.bias2, start_bin2)):
                    if table:
                        idx = np.fromiter(table.keys(), dtype=np.int64, count=len(table)) + beg
                        val = np.fromiter(table.values(), dtype=np.float64, count=len(table))
                        ok = idx < size
                        bias[idx[ok]], have[idx[ok]] = val[ok], True
                self.have_bias = have
            if want_decay:
                decay = np.full(size, np.nan)
                self.have_decay = np.zeros(size, dtype=bool)
                for crm, (beg, end) in reg.section_pos.items():
                    if end <= reg.rows[0] or beg >= reg.rows[1]:   # no row of the region in it
                        continue
                    table = self.decay_table.get(crm, {})
                    if table:
                        d = np.fromiter(table.keys(), dtype=np.int64, count=len(table))
                        val = np.fromiter(table.values(), dtype=np.float64, count=len(table))
                        ok = (d >= 0) & (d < end - beg)
                        decay[beg + d[ok]], self.have_decay[beg + d[ok]] = val[ok], True
        self.rows, self.cols, self.raw, self.nrm, self.dec = hic.store.region_cells(
            reg.rows[0], reg.rows[1], reg.cols[0], reg.cols[1], bad=bad,
            bias=bias if (want_norm or want_decay) else None, decay=decay, half=half)
        # region-local indices, as the reference's bins_dict1 / bins_dict2 give them
        self.a = self.rows.astype(np.int64) - start_bin1
        self.b = self.cols.astype(np.int64) - start_bin2
        if self.have_bias is not None and self.rows.size:
            for loc, glob in ((self.a, self.rows), (self.b, self.cols)):
                missing = ~self.have_bias[glob]
                if missing.any():                            # bias1[a]: KeyError in the reference
                    raise KeyError(int(loc[np.flatnonzero(missing)[0]]))
        self.chrom_of = np.empty(size, dtype=np.int64)
        self.beg_of = np.empty(size, dtype=np.int64)
        self.names = list(reg.section_pos)
        for n, (beg, end) in enumerate(reg.section_pos.values()):
            self.chrom_of[beg:end] = n
            self.beg_of[beg:end] = beg

    def check_decay(self):
        """decay[c][|a - b|] of the writers without a KeyError guard: '' between two chromosomes,
        the chromosome or the distance when the table lacks it."""
        if not self.rows.size:
            return
        trans = self.chrom_of[self.rows] != self.chrom_of[self.cols]
        if trans.any():
            raise KeyError('')
        dist = np.abs(self.rows.astype(np.int64) - self.cols)
        missing = ~self.have_decay[self.beg_of[self.rows] + dist]
        if missing.any():
            k = int(np.flatnonzero(missing)[0])
            crm = self.names[self.chrom_of[self.rows[k]]]
            raise KeyError(crm if crm not in self.decay_table else int(dist[k]))


def get_matrix(hic_data, resolution, biases=None,
               filter_exclude=(1, 2, 3, 4, 6, 7, 8, 9, 10),
               region1=None, start1=None, end1=None,
               region2=None, start2=None, end2=None, dico=None, clean=False,
               return_headers=False, tmpdir='.', normalization='raw', ncpus=8,
               nchunks=100, verbose=False, max_size=None, chr_order=None, as_arrays=False):
    """
    The matrix at the intersection of two regions as a dictionary, with the wanted normalisation
    applied (hic_bam_parser.py:755-872).

    :param hic_data: :class:`tadbit_b200.HiC_data` (the reference: path to a BAM file)
    :param resolution: resolution of the matrix (must be the one ``hic_data`` is binned at)
    :param biases: path to a pickle with biases, or its dictionary
    :param 'raw' normalization: 'raw', 'norm' (count / bias1 / bias2) or 'decay' (that over the
       expected count at the distance of the two bins)
    :param None region1: chromosome name; ``start1`` / ``end1`` genomic coordinates inside it
    :param None region2: second chromosome (columns); defaults to the first region
    :param None dico: something with ``dico[i, j] = count`` to fill with the raw counts instead of
       returning a dictionary
    :param False return_headers: also return bads1, bads2, regions, name, bin_coords
    :param None max_size: largest number of cells accepted
    :param False as_arrays: (not in the reference) return ``(bin1, bin2, values)`` arrays, row-major,
       instead of building the dictionary -- which costs more than extracting the cells

    :returns: dictionary ``{(bin1, bin2): value}``, bins counted from the start of each region
    """
    reg = _Region(hic_data, resolution, region1, start1, end1, region2, start2, end2, max_size, chr_order,
                  nchunks)
    if not biases and normalization != 'raw':
        raise Exception('ERROR: should provide path to file with biases (pickle).')
    if normalization not in ('raw', 'norm', 'decay'):
        raise NotImplementedError(('ERROR: %s normalization not implemented '
                                   'here') % normalization)
    cells = _Cells(hic_data, reg, biases, normalization == 'norm', normalization == 'decay', half=False)
    if normalization == 'decay':
        cells.check_decay()
    a, b = cells.a.tolist(), cells.b.tolist()
    if dico is not None:                                     # probably an HiC data object
        for i, j, v in zip(a, b, cells.raw.tolist()):
            dico[i, j] = v
        return None
    values = {'raw': cells.raw, 'norm': cells.nrm, 'decay': cells.dec}[normalization]
    if as_arrays:
        dico = (cells.a, cells.b, values)
    else:
        dico = dict(zip(zip(a, b), values.tolist()))
    if return_headers:
        name = _generate_name(reg.regions, (start1, start2), (end1, end2), resolution, chr_order)
        return dico, cells.bads1, cells.bads2, reg.regions, name, reg.bin_coords
    return dico


def _write_lines(path, a, b, values):
    """``a <tab> b <tab> <tab> value`` per cell, appended to ``path`` by tb_write_abc_lines (integers
    as they are, floats as ``repr`` writes them)."""
    import ctypes as C
    from . import _capi
    a = np.ascontiguousarray(a, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.int32)
    ints = vals = None
    if values.dtype.kind in 'iu':
        values = np.ascontiguousarray(values, dtype=np.int32)
        ints = _capi.ptr(values, C.c_int32)
    else:
        values = np.ascontiguousarray(values, dtype=np.float64)
        vals = _capi.ptr(values, C.c_double)
    _capi.check(_capi.lib().tb_write_abc_lines(os.fsencode(path), a.size, _capi.ptr(a, C.c_int32),
                                               _capi.ptr(b, C.c_int32), ints, vals))


def write_matrix(hic_data, resolution, biases, outdir,
                 filter_exclude=(1, 2, 3, 4, 6, 7, 8, 9, 10),
                 normalizations=('decay',),
                 region1=None, start1=None, end1=None, clean=True,
                 region2=None, start2=None, end2=None, extra='',
                 half_matrix=True, nchunks=100, tmpdir='.', append_to_tar=None,
                 ncpus=8, cooler=False, cooler_name=None, row_names=False,
                 chr_order=None, verbose=True, lengths=None):
    """
    Writes the ``.abc`` matrix files of ``tadbit bin`` (hic_bam_parser.py:892-1242): one file per
    normalisation, ``raw_`` / ``nrm_`` / ``dec_`` + region + resolution, a header (``# CRM`` lines,
    region and resolution, masked bins) and one line ``bin1 <tab> bin2 <tab> <tab> value`` per cell.

    :param hic_data: :class:`tadbit_b200.HiC_data` (the reference: path to a BAM file)
    :param biases: path to a pickle with biases, or its dictionary (may be None for 'raw' alone)
    :param outdir: folder for the output files
    :param ('decay',) normalizations: any of 'raw', 'norm', 'decay', 'raw&decay'
    :param True half_matrix: only the cells with bin1 <= bin2
    :param False row_names: genomic coordinates instead of bin numbers
    :param None lengths: ``{chromosome: length in base pairs}`` for the header

    :returns: dictionary ``{'RAW' | 'NRM' | 'DEC' | 'RAW&DEC': path}``
    """
    if start1 is not None and end1:
        if end1 - start1 < resolution:
            raise Exception('ERROR: region1 should be at least as big as resolution')
    if start2 is not None and end2:
        if end2 - start2 < resolution:
            raise Exception('ERROR: region2 should be at least as big as resolution')
    if isinstance(normalizations, list):
        normalizations = tuple(normalizations)
    elif isinstance(normalizations, basestring):
        normalizations = tuple([normalizations])
    if append_to_tar:
        raise NotImplementedError('append_to_tar')
    reg = _Region(hic_data, resolution, region1, start1, end1, region2, start2, end2, None, chr_order,
                  nchunks)
    regions = reg.regions
    start_bin1, end_bin1, start_bin2, end_bin2 = reg.bin_coords
    if lengths is None:
        lengths = dict((c, (n - 1) * resolution) for c, n in reg.sections.items())
    sections = OrderedDict((c, lengths[c]) for c in reg.sections)

    section_pos1, section_pos2 = {}, {}
    totals, total_num = OrderedDict(), 0
    for c in sections:
        totals[c] = total_num
        total_num += sections[c] / resolution + 1            # a true division under Python 3
    if row_names:
        if len(regions) in [1, 2]:
            offset = start_bin1 - totals[regions[0]]
            section_pos1 = dict((i, (region1, resolution * (i + offset))) for i in range(end_bin1 - start_bin1))
            if region2:
                offset = start_bin2 - totals[regions[1]]
                section_pos2 = dict((i, (region2, resolution * (i + offset))) for i in range(end_bin2 - start_bin2))
        else:
            totals.iteritems()                               # AttributeError, as in the reference

    if not biases and normalizations != ('raw', ):
        raise Exception('ERROR: should provide path to file with biases (pickle).')
    want_norm = 'norm' in normalizations
    want_decay = 'decay' in normalizations
    # the raw&decay writer evaluates v / bias1[a] / bias2[b] for every cell before it prints the raw
    # count: a bin without a bias is a KeyError there too
    cells = _Cells(hic_data, reg, biases, want_norm and not cooler, want_decay and not cooler, half=half_matrix,
                   touch_bias='raw&decay' in normalizations and not cooler)
    bads1, bads2 = cells.bads1, cells.bads2

    name = _generate_name(regions, (start1, start2), (end1, end2), resolution, chr_order)
    if cooler:
        try:
            import h5py  # noqa: F401
        except ImportError:
            raise Exception('ERROR: cooler output is not available. Probably ' +
                            'you need to install h5py\n')
        raise NotImplementedError('cooler output')

    def header():
        text = ''.join('# CRM %s\t%d\n' % (r, sections[r]) for r in regions)
        text += '# %s resolution:%d\n' % (name, resolution)
        if region2:
            text += '# BADROWS %s\n' % (','.join([str(b) for b in bads1]))
            text += '# BADCOLS %s\n' % (','.join([str(b) for b in bads2]))
        else:
            text += '# MASKED %s\n' % (','.join([str(b) for b in bads1]))
        return text

    tail = '_%s%s.abc' % (nicer(resolution).replace(' ', ''), ('_' + extra) if extra else '')
    files = OrderedDict()                                    # path -> per-cell lists of lines, in writer order
    fnames = {}
    for key, prefix, wanted in (('RAW', 'raw_', 'raw' in normalizations), ('NRM', 'nrm_', want_norm),
                                ('DEC', 'dec_', want_decay or 'raw&decay' in normalizations)):
        if wanted:
            path = os.path.join(outdir, prefix + name + tail)
            files[key] = (path, [])
            with open(path, 'w') as out:
                out.write(header())
    if 'raw' in normalizations:
        fnames['RAW'] = files['RAW'][0]
    if want_norm:
        fnames['NRM'] = files['NRM'][0]
    if want_decay:
        fnames['DEC'] = files['DEC'][0]
    if 'raw&decay' in normalizations:
        fnames['RAW&DEC'] = files['DEC'][0]

    # per file the value columns its writers print, in the reference's writer order
    if 'raw' in normalizations:
        files['RAW'][1].append(cells.raw)
    if want_norm:
        files['NRM'][1].append(cells.nrm)
    if want_decay:
        if len(regions) in [1, 2]:                           # write_expc / write_expc_2reg: no guard
            cells.check_decay()
        files['DEC'][1].append(cells.dec)                    # write_expc_err: 'nan' where there is no decay
    if 'raw&decay' in normalizations:
        # the reference's format string has two fields for three values: the raw count is written
        files['DEC'][1].append(cells.raw)
    names = None
    if row_names:
        a, b = cells.a.tolist(), cells.b.tolist()
        for i, j in zip(a, b):                               # section_pos2 stays empty without region2
            if i not in section_pos1:
                raise KeyError(i)
            if j not in section_pos2:
                raise KeyError(j)
        names = ['{}\t{}\t{}\t{}\t'.format(*(section_pos1[i] + section_pos2[j])) for i, j in zip(a, b)]
    for path, columns in files.values():
        if names is None and len(columns) == 1:              # the usual case: formatted by the library
            _write_lines(path, cells.a, cells.b, columns[0])
            continue
        if names is None:
            names = ['%d\t%d\t' % (i, j) for i, j in zip(cells.a.tolist(), cells.b.tolist())]
        text = [['%s\t%r\n' % (n, v) for n, v in zip(names, col.tolist())] for col in columns]
        with open(path, 'a') as out:
            if len(text) == 1:
                out.write(''.join(text[0]))
            else:                                            # decay and raw&decay share a file, cell by cell
                out.write(''.join(''.join(lines) for lines in zip(*text)))
    return fnames
