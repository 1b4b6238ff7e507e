"""Normalised interaction pairs and their z-scores on the GPU.

Mirrors the part of ``pytadbit.Experiment`` that consumes the ICE biases
(reference: _pytadbit/experiment.py:704-748 ``normalize_hic``, :751-799 ``get_hic_zscores``,
:1078-1189 ``write_interaction_pairs``; _pytadbit/utils/tadmaths.py:94-172 ``nozero_log`` /
``zscore``).  The reference materialises ``hic[i, j] / bias[i] / bias[j]`` for all N^2 pairs in a
dict and then walks the pairs again in Python; here the stored cells are streamed once per
statistic by CUDA kernels (``tb_pair_stats``, ``tb_pair_values``) and pairs without contacts are
accounted for in closed form, so nothing of size N^2 is ever built unless the caller asks for the
dense dictionaries of the reference (``remove_zeros=False``).

Everything else of the reference class (TAD calling, 3D modelling, plots) is out of scope.
"""
import numpy as np

from ._capi import bad_mask
from .hic_data import HiC_data
from .vectors import BinVector, PairTable


class NormalizedMatrix(object):
    """``Experiment.norm[0]`` of the reference without the N^2 dict: ``norm[i * N + j]`` (or
    ``norm[i, j]``) is ``hic[i, j] / bias[i] / bias[j]``, computed on demand."""

    def __init__(self, hic, bias):
        self._hic, self._bias = hic, bias
        self.bads = hic.bads

    def __len__(self):
        return len(self._hic)

    def __getitem__(self, key):
        try:
            i, j = key
        except TypeError:
            i, j = divmod(int(key), len(self._hic))
        return float(self._hic[i, j]) / self._bias[i] / self._bias[j]

    def get(self, key, default=0):
        return self[key]


class Experiment(object):
    """The Hi-C side of ``pytadbit.Experiment``.

    :param name: name of the experiment
    :param resolution: bin size in base pairs
    :param None hic_data: a :class:`tadbit_b200.HiC_data`
    """

    def __init__(self, name, resolution, hic_data=None, conditions=None, identifier=None,
                 cell_type=None, enzyme=None, exp_type='Hi-C', **kw_descr):
        self.name = name
        self.resolution = resolution
        self.identifier = identifier
        self.cell_type = cell_type
        self.enzyme = enzyme
        self.exp_type = exp_type
        self.crm = None                                      # the chromosome object it belongs to (.name)
        self.conditions = sorted(conditions) if conditions else []
        self.description = kw_descr
        self.hic_data = None
        self.size = None
        self.norm = None
        self._normalization = None
        self._filtered_cols = False
        self._zeros = {}
        self._zscores = {}
        self._pairs = None
        if hic_data is not None:
            self.load_hic_data(hic_data)

    def load_hic_data(self, hic_data, **kwargs):
        """experiment.py:546-591 for data that are already a ``HiC_data``."""
        if isinstance(hic_data, HiC_data):
            hic_data = [hic_data]
        self.hic_data = list(hic_data)
        self.size = len(self.hic_data[0])
        if self.hic_data[0].bads:
            self._zeros = self.hic_data[0].bads
            self._filtered_cols = True

    def filter_columns(self, silent=False, draw_hist=False, savefig=None,
                       perc_zero=99, by_mean=True, min_count=None):
        """experiment.py:515-543: filtered columns end up in ``Experiment._zeros``."""
        data = self.hic_data[0]
        data.filter_columns(draw_hist=draw_hist, savefig=savefig, perc_zero=perc_zero,
                            by_mean=by_mean, min_count=min_count, silent=silent)
        self._zeros = data.bads
        self._filtered_cols = True

    def normalize_hic(self, factor=1, iterations=0, max_dev=0.1, silent=False):
        """experiment.py:704-748.  ``self.norm[0][i * size + j]`` gives the normalised value."""
        if not self.hic_data:
            raise Exception('ERROR: No Hi-C data loaded\n')
        if self.norm and not silent:
            from sys import stderr
            stderr.write('WARNING: removing previous weights\n')
        self.hic_data[0].normalize_hic(iterations=iterations, max_dev=max_dev,
                                       silent=silent, factor=factor)
        self.norm = [NormalizedMatrix(self.hic_data[0], self.hic_data[0].bias)]
        if factor:
            self._normalization = 'visibility_factor:' + str(factor)
        else:
            self._normalization = 'visibility'

    # -- the pairs -----------------------------------------------------------------------------
    def _bias_array(self):
        bias = self.hic_data[0].bias
        if isinstance(bias, BinVector):
            return bias.to_array()
        return np.array([bias[i] for i in range(self.size)], dtype=np.float64)

    def pair_zscores(self, normalized=True, zscored=True):
        """Arrays ``(rows, cols, values)`` of the pairs i < j with contacts between columns that are
        not filtered out -- what ``get_hic_zscores(remove_zeros=True)`` stores in
        ``Experiment._zscores``, without the dictionaries.  Also keeps the statistics in
        ``self._zscore_stats`` = (pairs, smallest value, mean, std of log10)."""
        store = self.hic_data[0].store
        mask = bad_mask(self.size, self._zeros)
        bias = self._bias_array() if normalized else None
        zs = None
        if zscored:
            n, vmin, slog, _ = store.pair_stats(bias, mask, 0.0)
            if n == 0:
                raise ValueError('min() arg is an empty sequence')          # tadmaths.py:96
            mean = slog / n
            _, _, _, dev2 = store.pair_stats(bias, mask, mean)              # np.std: two passes
            std = (dev2 / n) ** 0.5
            self._zscore_stats = (n, vmin, mean, std)
            zs = (mean, std)
        return store.pair_values(bias, mask, zscore=zs)

    def get_hic_zscores(self, normalized=True, zscored=True, remove_zeros=True):
        """experiment.py:751-799: fills ``Experiment._zscores`` (``{str(i): {str(j): value}}``)."""
        if normalized and not self.norm:
            raise TypeError("'NoneType' object is not subscriptable")       # self.norm[0] of the reference
        self._zscores = {}
        good = [i for i in range(self.size) if i not in self._zeros]
        dense = not (normalized and remove_zeros)
        if not dense:
            rows, cols, vals = self.pair_zscores(normalized=True, zscored=zscored)
        else:
            rows, cols, vals = self._dense_pairs(good, normalized, zscored)
        self._pairs = (rows, cols, vals)
        self._zscores = PairTable(rows, cols, vals)     # {str(i): {str(j): value}} over the arrays

    def _dense_pairs(self, good, normalized, zscored):
        """All pairs i < j between good columns, zeros included (``remove_zeros=False`` or raw
        counts): N^2 / 2 values, as in the reference -- for matrices small enough to hold them."""
        store = self.hic_data[0].store
        mask = bad_mask(self.size, self._zeros)
        bias = self._bias_array() if normalized else None
        r, c, v = store.pair_values(bias, mask, zscore=None)
        g = np.asarray(good, dtype=np.int64)
        pos = np.full(self.size, -1, dtype=np.int64)
        pos[g] = np.arange(g.size)
        mat = np.zeros((g.size, g.size), dtype=np.float64)
        mat[pos[r], pos[c]] = v
        iu = np.triu_indices(g.size, 1)
        vals = mat[iu]
        if not normalized:
            vals = vals.astype(np.int64)
        rows, cols = g[iu[0]], g[iu[1]]
        if zscored:
            vals = vals.astype(np.float64)
            nonzero = vals[vals != 0]
            if nonzero.size == 0:
                raise ValueError('min() arg is an empty sequence')
            # tadmaths.nozero_log: np.log10(0) is -inf, it does not raise, so the "virtual minimum"
            # is never used and zeros poison the mean (the docstring's "Dangerous")
            with np.errstate(divide='ignore', invalid='ignore'):
                logs = np.log10(vals)
                mean_v = np.mean(logs)
                std_v = np.std(logs)
                vals = (logs - mean_v) / std_v
        return rows.astype(np.int32), cols.astype(np.int32), vals

    def write_interaction_pairs(self, fname, normalized=True, zscored=True,
                                diagonal=False, cutoff=None, header=False,
                                true_position=False, uniq=True,
                                remove_zeros=False, focus=None, format='tsv'):
        """experiment.py:1078-1189 (tsv and json, same lines)."""
        cutoff = cutoff or float('-inf')
        if not self._zscores and zscored:
            self.get_hic_zscores()
        if not self.norm and normalized:
            raise Exception('Experiment not normalized.')
        out = open(fname, 'w') if isinstance(fname, str) else fname
        if header and format == 'tsv':
            out.write('elt1\telt2\t%s\n' % ('zscore' if zscored else
                                            'normalized hi-c' if normalized
                                            else 'raw hi-c'))
        elif header and format == 'json':                    # experiment.py:1118-1144, the same text
            out.write('''
{
    "metadata": {
                        "formatVersion" : 3,
                        %s
                        "species" : "%s",
                        "cellType" : "%s",
                        "experimentType" : "%s",
                        "identifier" : "%s",
                        "resolution" : %s,
                        "chromosome" : "%s",
                        "start" : %s,
                        "end" : %s
                },
    "interactions": [
                ''' % ('\n'.join(['"%s": "%s",' % (k, self.description[k])
                                  for k in self.description]),
                       self.description.get('species', ''),
                       self.cell_type,
                       self.exp_type,
                       self.identifier,
                       self.resolution,
                       self.crm.name,
                       focus[0] * self.resolution if focus else 1,
                       (focus[1] * self.resolution if focus else
                        self.resolution * self.size)))
        if focus:
            start, end = focus[0], focus[1] + 1
        else:
            start, end = 0, self.size
        sep = '\t' if format == 'tsv' else ','
        pairs = self._sparse_pairs(normalized, zscored, diagonal, cutoff, uniq, remove_zeros, start, end)
        if pairs is not None and format in ('tsv', 'json'):
            rows, cols, vals = pairs
            if true_position:
                a, b = self.resolution * (rows + 1), self.resolution * (cols + 1)
                tail = ',\n' if format == 'json' else '\n'
            else:
                a, b, tail = rows + 1 - start, cols + 1 - start, '\n'
            if isinstance(fname, str):
                out.close()
                _write_pair_lines(fname, a, b, vals, sep, tail)
                out = open(fname, 'a')
            else:
                out.write(''.join('%s%s%s%s%s%s' % (x, sep, y, sep, v, tail)
                                  for x, y, v in zip(a.tolist(), b.tolist(), vals.tolist())))
            start = end                                   # nothing left for the loop below
        for i in range(start, end):
            if i in self._zeros:
                continue
            newstart = i if uniq else 0
            for j in range(newstart, end):
                if j in self._zeros:
                    continue
                if not diagonal and i == j:
                    continue
                if zscored:
                    try:
                        if self._zscores[str(i)][str(j)] < cutoff:
                            continue
                        if self._zscores[str(i)][str(j)] == -99:
                            continue
                    except KeyError:
                        continue
                    val = self._zscores[str(i)][str(j)]
                elif normalized:
                    val = self.norm[0][self.size * i + j]
                else:
                    val = self.hic_data[0][self.size * i + j]
                if remove_zeros and not val:
                    continue
                if true_position:
                    a, b = self.resolution * (i + 1), self.resolution * (j + 1)
                    tail = ',\n' if format == 'json' else '\n'
                else:
                    a, b = i + 1 - start, j + 1 - start
                    tail = '\n'
                out.write('%s%s%s%s%s%s' % (a, sep, b, sep, val, tail))
        if format == 'json':
            out.write(']}\n')
        out.close()


    # -- the dense matrix ----------------------------------------------------------------------
    def _dense(self, start, end, normalized):
        store = self.hic_data[0].store
        if normalized:
            return store.window(start, end, start, end, self._bias_array()).tolist()
        return np.rint(store.window(start, end, start, end)).astype(np.int64).tolist()

    def get_hic_matrix(self, focus=None, diagonal=True, normalized=False):
        """
        Return the Hi-C matrix (experiment.py:1192-1226) -- a dense window from ``tb_window``.

        :param None focus: a tuple (start, end), both inclusive, bins counted from 1
        :param True diagonal: False replaces the values in the diagonal by one (zero where empty)
        :param False normalized: normalized data instead of raw Hi-C

        :returns: list of lists
        """
        if normalized and not self.norm:
            raise Exception('ERROR: experiment not normalized yet')
        if focus:
            start, end = focus
            start -= 1
        else:
            start, end = 0, self.size
        mtrx = self._dense(start, end, normalized)
        if not diagonal:
            for i in range(start, end):                       # absolute indices, as in the reference
                mtrx[i][i] = 1 if mtrx[i][i] else 0
        return mtrx

    def print_hic_matrix(self, print_it=True, normalized=False, zeros=False):
        """
        Return the Hi-C matrix as string (experiment.py:1229-1255)

        :param True print_it: Otherwise, returns the string
        :param False normalized: normalized data, instead of raw Hi-C
        :param False zeros: write 'nan' in the filtered columns
        """
        siz = self.size
        if normalized and not self.norm:
            raise TypeError("'NoneType' object is not subscriptable")     # self.norm[0] of the reference
        rows = self._dense(0, siz, normalized)
        if zeros:
            zero = [i in self._zeros for i in range(siz)]
            out = '\n'.join(['\t'.join(['nan' if (zero[i] or zero[j]) else str(rows[j][i])
                                        for i in range(siz)]) for j in range(siz)])
        else:
            out = '\n'.join(['\t'.join([str(x) for x in row]) for row in rows])
        if print_it:
            print(out)
        else:
            return out + '\n'

    def _sparse_pairs(self, normalized, zscored, diagonal, cutoff, uniq, remove_zeros, start, end):
        """The pairs the double loop of write_interaction_pairs would print, as arrays, when they
        can be had without walking N^2 pairs: z-scores kept as a PairTable (only stored pairs pass
        the loop's KeyError guard), or values without their zeros.  None otherwise."""
        if zscored:
            table = self._zscores
            if not isinstance(table, PairTable):
                return None
            rows, cols, vals = table.rows, table.cols, table.vals
            with np.errstate(invalid='ignore'):
                keep = ~(vals < cutoff) & (vals != -99)
        elif remove_zeros and uniq and not diagonal:
            store = self.hic_data[0].store
            bias = self._bias_array() if normalized else None
            rows, cols, vals = store.pair_values(bias, bad_mask(self.size, self._zeros), zscore=None)
            rows, cols = rows.astype(np.int64), cols.astype(np.int64)
            if not normalized:
                vals = vals.astype(np.int64)
            keep = np.ones(rows.size, dtype=bool)
        else:
            return None
        zero = bad_mask(self.size, self._zeros).astype(bool)
        keep &= (rows >= start) & (rows < end) & (cols >= start) & (cols < end) & ~zero[rows] & ~zero[cols]
        if remove_zeros:
            keep &= vals != 0
        return rows[keep], cols[keep], vals[keep]


def _write_pair_lines(path, a, b, vals, sep, tail):
    """``a <sep> b <sep> value <tail>`` per pair, appended to ``path`` by tb_write_pair_lines."""
    import ctypes as C
    import os
    from . import _capi
    a = np.ascontiguousarray(a, dtype=np.int64)
    b = np.ascontiguousarray(b, dtype=np.int64)
    ints = flts = None
    if vals.dtype.kind in 'iu':
        vals = np.ascontiguousarray(vals, dtype=np.int64)
        ints = _capi.ptr(vals, C.c_int64)
    else:
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        flts = _capi.ptr(vals, C.c_double)
    _capi.check(_capi.lib().tb_write_pair_lines(os.fsencode(path), a.size, _capi.ptr(a, C.c_int64),
                                                _capi.ptr(b, C.c_int64), ints, flts, sep.encode(),
                                                tail.encode()))


__all__ = ["Experiment", "NormalizedMatrix"]
