"""``HiC_data``: the reference's container surface over a GPU-resident contact store.

Mirrors ``pytadbit.HiC_data`` (reference: _pytadbit/hic_data.py:54-181, 318-444) for the
matrix-balancing path: same constructor arguments, ``len()``, ``hic[i, j]``,
``filter_columns``, ``sum``, ``normalize_expected``, ``normalize_hic``, ``save_biases`` /
``load_biases``, ``get_hic_data_as_csr`` and the attributes ``bias``, ``bads``,
``expected``, ``chromosomes``, ``sections``, ``section_pos``, ``resolution``,
``symmetricized``.  The cells live in HBM as a ``ContactStore``; all reductions are
CUDA kernels behind the C-ABI.  Dense windows (``get_matrix`` ...) and ``find_compartments``
are the consumers of the biases that SURVEY.md 8f ranks next; the rest of the reference class
(``find_compartments_beta``, cooler export, ...) is out of scope.
"""
from collections import OrderedDict
from pickle import HIGHEST_PROTOCOL, dump, load

import numpy as np

from .hic_filtering import filter_by_mean, filter_by_zero_count
from .normalize_hic import expected, iterative
from .store import ContactStore
from .vectors import BinVector

_INT32_MAX = 2147483647


def isclose(a, b, rel_tol=1e-09, abs_tol=0.0):
    return abs(a - b) <= max(rel_tol * max(abs(a), abs(b)), abs_tol)


def _as_key_value_arrays(items):
    """dict or iterable of (key, value) -> (keys int64, values) in insertion order,
    later duplicates overriding earlier ones as a dict would."""
    if not isinstance(items, dict):
        items = dict(items)
    n = len(items)
    keys = np.fromiter(items.keys(), dtype=np.int64, count=n)
    vals = np.array(list(items.values())) if n else np.zeros(0, dtype=np.int64)
    return keys, vals


def _symmetric_upper(keys, vals, size):
    """Upper-triangle cells of the matrix the reference would hold after
    ``_symmetricize`` (hic_data.py:83-118).

    The reference inspects the first non-zero off-diagonal cells (in dict order): if one
    differs from its mirror the matrix is completed -- mirrors copied when one side is
    empty, both sides summed when both hold counts -- otherwise it is taken as is.
    """
    order = np.argsort(keys, kind="stable")
    skeys, svals = keys[order], vals[order]

    def lookup(k):
        pos = np.searchsorted(skeys, k)
        pos = np.minimum(pos, max(skeys.size - 1, 0))
        found = (skeys[pos] == k) if skeys.size else np.zeros(np.shape(k), dtype=bool)
        return found, pos

    rows, cols = np.divmod(keys, size)
    mirror_found, mirror_pos = lookup(cols * size + rows)
    mirror_val = np.where(mirror_found, svals[mirror_pos] if skeys.size else 0, 0)

    # --- detection on the first cells, as the reference does -------------------------
    symmetric, to_sum, seen = True, False, 0
    for n in np.flatnonzero((rows != cols) & (vals != 0)):
        if not isclose(vals[n], mirror_val[n]):
            to_sum = bool(mirror_val[n] != 0)
            symmetric = False
            break
        if seen > 10:
            break
        seen += 1

    upper = rows <= cols
    lower = ~upper
    if symmetric:
        # cells only present below the diagonal would be invisible to an upper store
        if np.any(lower & ~mirror_found) or np.any(
                upper & mirror_found & ~np.isclose(vals, mirror_val, rtol=1e-9, atol=0)):
            raise ValueError('matrix is neither symmetric nor consistently half-filled')
        return rows[upper], cols[upper], vals[upper]
    if to_sum:
        # both mirrors summed; a pair present on both sides is visited twice (hic_data.py:106-111)
        tot = np.where(rows == cols, vals, np.where(mirror_found, 2 * (vals + mirror_val), vals))
    else:
        # mirrors copied: the cell visited first fixes the pair (hic_data.py:112-116)
        visit = np.arange(keys.size)
        mirror_visit = np.where(mirror_found, order[mirror_pos] if skeys.size else 0, keys.size)
        tot = np.where(visit <= mirror_visit, vals, mirror_val)
    keep = upper | (lower & ~mirror_found)
    r = np.where(upper, rows, cols)[keep]
    c = np.where(upper, cols, rows)[keep]
    return r, c, tot[keep]


class HiC_data(object):
    """Hi-C matrix whose cells are resident on a B200; see the module docstring.

    :param items: dict (or iterable of pairs) ``{row * size + col: count}`` as the
       reference takes it, either triangle or both
    :param size: number of bins
    :param None chromosomes: OrderedDict ``name -> number of bins``
    :param None dict_sec: ``{(chromosome, position): bin}``
    :param 0 device: CUDA device ordinal
    """

    def __init__(self, items, size, chromosomes=None, dict_sec=None,
                 resolution=1, masked=None, symmetricized=False, device=0):
        keys, vals = _as_key_value_arrays(items)
        if keys.size and (keys.min() < 0 or keys.max() >= size * size):
            raise IndexError('ERROR: position larger than %s^2' % size)
        rows, cols, vals = _symmetric_upper(keys, vals, size)
        self._init_common(size, chromosomes, dict_sec, resolution, masked, symmetricized)
        self._set_cells(rows, cols, vals, device)

    # -- alternative constructors ------------------------------------------------------
    @classmethod
    def from_coo(cls, size, rows, cols, vals, chromosomes=None, dict_sec=None, resolution=1,
                 masked=None, symmetricized=False, device=0):
        """From upper-triangle cells (``rows[k] <= cols[k]``, each cell once)."""
        self = cls.__new__(cls)
        self._init_common(size, chromosomes, dict_sec, resolution, masked, symmetricized)
        self._set_cells(np.asarray(rows), np.asarray(cols), np.asarray(vals), device)
        return self

    @classmethod
    def from_store(cls, store, chromosomes=None, dict_sec=None, resolution=1, masked=None,
                   symmetricized=False):
        """Around a store that is already resident (synthetic generator, device data)."""
        self = cls.__new__(cls)
        self._init_common(store.size, chromosomes, dict_sec, resolution, masked, symmetricized)
        self.store = store
        self._keys = self._vals = None
        if hasattr(type(store), 'from_coo'):               # cells set later rebuild a store of this kind
            self._new_store = type(store).from_coo
        return self

    def _init_common(self, size, chromosomes, dict_sec, resolution, masked, symmetricized):
        self.__size = int(size)
        self._size2 = self.__size ** 2
        self.bias = None
        self.bads = masked or {}
        self.chromosomes = chromosomes
        self.sections = dict_sec
        self.section_pos = {}
        self.resolution = resolution
        self.expected = None
        self.symmetricized = symmetricized
        self.compartments = {}
        self.last_trace = None
        self._pending = {}                                     # cells set since the store was built
        self._stale = False                                    # chromosomes changed since then
        self._store = None
        self._device = 0
        self._new_store = ContactStore.from_coo
        if self.chromosomes:                                   # hic_data.py:70-75
            total = 0
            for crm in self.chromosomes:
                self.section_pos[crm] = (total, total + self.chromosomes[crm])
                total += self.chromosomes[crm]
        if self.sections == {}:                                # hic_data.py:76-79
            self.section_pos = {None: (0, self.__size)}
            self.sections = dict([((None, i), i) for i in range(0, self.__size)])

    def _set_cells(self, rows, cols, vals, device):
        vals = np.asarray(vals)
        if vals.size:
            if vals.dtype.kind == 'f':
                if not np.all(vals == np.rint(vals)):
                    raise NotImplementedError(
                        'tadbit_b200 balances raw integer counts; normalised (float) '
                        'matrices are outside the accelerated path')
                vals = vals.astype(np.int64)
            if np.abs(vals).max() > _INT32_MAX:
                raise OverflowError('cell count does not fit int32')
        rows = np.asarray(rows, dtype=np.int64)
        cols = np.asarray(cols, dtype=np.int64)
        if rows.size and (rows.min() < 0 or cols.max() >= self.__size):
            raise IndexError('ERROR: row or column larger than %s' % self.__size)
        keys = rows * self.__size + cols
        order = np.argsort(keys, kind="stable")
        self._keys = keys[order]
        self._vals = vals[order].astype(np.int64)
        chrom = self.chromosomes if self.chromosomes else None
        self._device = device
        self.store = ContactStore.from_coo(self.__size, rows[order], cols[order],
                                           self._vals, chromosomes=chrom, device=device)

    # -- the resident store; cells set through ``hic[i, j] = v`` are folded in before it is used -----
    @property
    def store(self):
        if self._pending or self._stale:
            self._rebuild()
        return self._store

    @store.setter
    def store(self, value):
        self._store = value

    def _rebuild(self):
        """Fold the pending cells into the host copy of the upper triangle and build the store anew
        (a store is immutable: its compact stream is laid out once)."""
        if hasattr(self._store, "info"):
            block = self._store.info()
            if (block["row_begin"], block["row_end"]) != (0, self.__size):
                raise NotImplementedError('the cells / chromosomes of a row-sharded matrix cannot be '
                                          'changed: build the shards from the new data')
        if self._keys is None:
            r, c, v = self._store.export_upper()
            self._keys = r.astype(np.int64) * self.__size + c
            self._vals = v.astype(np.int64)
        self._stale = False
        pend_k = np.fromiter(self._pending.keys(), dtype=np.int64, count=len(self._pending))
        pend_v = np.fromiter(self._pending.values(), dtype=np.int64, count=len(self._pending))
        self._pending = {}
        order = np.argsort(pend_k)
        pend_k, pend_v = pend_k[order], pend_v[order]
        pos = np.searchsorted(self._keys, pend_k)
        found = (pos < self._keys.size) & (self._keys[np.minimum(pos, max(self._keys.size - 1, 0))] == pend_k) \
            if self._keys.size else np.zeros(pend_k.size, dtype=bool)
        vals = self._vals.copy()
        vals[pos[found]] = pend_v[found]
        keys = np.concatenate([self._keys, pend_k[~found]])
        vals = np.concatenate([vals, pend_v[~found]])
        order = np.argsort(keys, kind="stable")
        self._keys, self._vals = keys[order], vals[order]
        rows, cols = np.divmod(self._keys, self.__size)
        old = self._store
        chrom = self.chromosomes if self.chromosomes else None
        kw = dict(chromosomes=chrom)
        if self._new_store == ContactStore.from_coo:
            kw["device"] = self._device
        self._store = self._new_store(self.__size, rows, cols, self._vals, **kw)
        if old is not None and hasattr(old, "close"):
            old.close()

    # -- dict-like surface --------------------------------------------------------------
    def __len__(self):
        return self.__size

    def _host_cells(self):
        if self._pending or self._stale:
            self._rebuild()
        if self._keys is None:
            r, c, v = self.store.export_upper()
            self._keys = r.astype(np.int64) * self.__size + c
            self._vals = v.astype(np.int64)
        return self._keys, self._vals

    def get(self, pos, default=0):
        row, col = divmod(int(pos), self.__size)
        if row > col:
            row, col = col, row
        k = row * self.__size + col
        if k in self._pending:
            return self._pending[k]
        if self._keys is None:
            self._host_cells()
        keys, vals = self._keys, self._vals
        p = int(np.searchsorted(keys, k))
        if p < keys.size and keys[p] == k:
            return int(vals[p])
        return default

    def __getitem__(self, row_col):
        try:
            row, col = row_col
            pos = row * self.__size + col
            if pos > self._size2:
                raise IndexError('ERROR: row or column larger than %s' % self.__size)
            return self.get(pos, 0)
        except TypeError:
            if row_col > self._size2:
                raise IndexError('ERROR: position %d larger than %s^2' % (row_col, self.__size))
            return self.get(row_col, 0)

    def __setitem__(self, row_col, val):
        """``hic[row, col] = count`` (hic_data.py:146-164).  The matrix stays symmetric: the cell and
        its mirror are one stored value (the reference's dictionary would take them one by one; its
        parsers always set both).  Cells set this way are kept aside and folded into a new resident
        store the next time the matrix is used, so set many, then compute."""
        try:
            row, col = row_col
            pos = row * self.__size + col
            if pos > self._size2:
                print(row, col, pos)
                raise IndexError('ERROR: row or column larger than %s' % self.__size)
        except TypeError:
            if row_col > self._size2:
                raise IndexError('ERROR: position %d larger than %s^2' % (row_col, self.__size))
            row, col = divmod(int(row_col), self.__size)
        if val != int(val):
            raise NotImplementedError('tadbit_b200 balances raw integer counts; normalised (float) '
                                      'matrices are outside the accelerated path')
        if abs(int(val)) > _INT32_MAX:
            raise OverflowError('cell count does not fit int32')
        row, col = int(row), int(col)
        if not (0 <= row < self.__size and 0 <= col < self.__size):
            raise IndexError('ERROR: row or column larger than %s' % self.__size)
        if row > col:
            row, col = col, row
        self._pending[row * self.__size + col] = int(val)

    def items(self):
        """(key, value) over BOTH triangles, as the reference's dict holds them."""
        keys, vals = self._host_cells()
        n = self.__size
        for k, v in zip(keys.tolist(), vals.tolist()):
            yield k, v
            i, j = divmod(k, n)
            if i != j:
                yield j * n + i, v

    def keys(self):
        return (k for k, _ in self.items())

    def values(self):
        return (v for _, v in self.items())

    def __iter__(self):
        return self.keys()

    def get_hic_data_as_csr(self):
        """scipy CSR of the full symmetric matrix (hic_data.py:166-181)."""
        from scipy.sparse import csr_matrix
        keys, vals = self._host_cells()
        i, j = np.divmod(keys, self.__size)
        od = i != j
        rows = np.concatenate([i, j[od]])
        cols = np.concatenate([j, i[od]])
        data = np.concatenate([vals, vals[od]]).astype(float)
        return csr_matrix((data, (rows, cols)), shape=(self.__size, self.__size))

    def add_sections(self, lengths, chr_names=None, binned=False):
        """
        Add genomic coordinates to the object from a list of chromosome lengths; order matters
        (hic_data.py:216-250).  The per-chromosome reductions (expected, decay, compartments) follow
        the new chromosomes: the resident store is laid out again on the next use.

        :param lengths: list of chromosome lengths
        :param None chr_names: list of corresponding chromosome names
        :param False binned: if True, lengths will not be divided by resolution
        """
        from warnings import warn
        genome_seq = OrderedDict()
        size = 0
        resolution = 1 if binned else self.resolution
        for crm, length in enumerate(lengths):
            cnam = 'chr' + str(crm) if not chr_names else chr_names[crm]
            genome_seq[cnam] = int(length) // resolution + 1
            size += genome_seq[cnam]
        if size != self.__size:
            warn('WARNING: different sizes (%d, now:%d), ' % (self.__size, size)
                 + 'should adjust the resolution')
            # the reference goes on with a matrix whose size no longer matches its cells
            raise ValueError('chromosomes of %d bins in all do not fit a matrix of %d bins'
                             % (size, self.__size))
        sections = [(crm, i) for crm in genome_seq for i in range(genome_seq[crm])]
        self.chromosomes = genome_seq
        self.sections = dict([(j, i) for i, j in enumerate(sections)])
        # The nameless whole-matrix section of an object built without chromosomes goes: the
        # reference keeps it next to the new chromosomes, and its `expected` then walks every
        # diagonal once per entry (cells inside a chromosome are counted twice).
        self.section_pos.pop(None, None)
        total = 0
        for crm in self.chromosomes:
            self.section_pos[crm] = (total, total + self.chromosomes[crm])
            total += self.chromosomes[crm]
        self._stale = True

    # -- the accelerated path -------------------------------------------------------------
    def filter_columns(self, draw_hist=False, savefig=None, perc_zero=99,
                       by_mean=True, min_count=None, silent=False):
        """Flag artifactual columns (hic_data.py:318-348); result in ``self.bads``."""
        self.bads = filter_by_zero_count(self, perc_zero, min_count=min_count,
                                         silent=silent)
        if by_mean:
            self.bads.update(filter_by_mean(
                self, draw_hist=draw_hist, silent=silent,
                savefig=savefig, bads=self.bads))
        if not silent:
            print('Found %d of %d columns with poor signal' % (len(self.bads),
                                                               len(self)))

    def sum(self, bias=None, bads=None):
        """Sum of the matrix skipping bad columns, optionally divided by the biases
        (hic_data.py:350-375).  K3 on the device."""
        from ._capi import bad_mask
        bads = bads or self.bads
        mask = bad_mask(self.__size, bads)
        if bias:
            arr = bias.to_array() if isinstance(bias, BinVector) else np.array(
                [bias[i] for i in range(self.__size)], dtype=np.float64)
            return self.store.norm_sum(arr, mask)
        return int(self.store.norm_sum(None, mask))

    def cis_trans_ratio(self, normalized=False, exclude=None, diagonal=True, equals=None):
        """Interactions inside chromosomes over all interactions between good bins
        (hic_data.py:252-316).  One pass of the per-chromosome distance sums (``tb_decay_sums``): the
        cells of a chromosome at distance d > 0 count twice (both triangles), the diagonal once.

        :param False normalized: use count / bias[i] / bias[j]
        :param None exclude: chromosomes to leave out
        :param True diagonal: include the diagonal in the cis count
        :param None equals: function of two chromosome names that says whether they count as one
           (e.g. ``lambda x, y: x[:4] == y[:4]`` for chr2L / chr2R; consecutive chromosomes only, as
           in the reference).  Costs one pass of the normalised sum per group of chromosomes.
        """
        from ._capi import bad_mask
        if normalized and not self.bias:
            raise Exception('ERROR: experiment not normalized yet')
        if not self.chromosomes:
            return float('nan')
        bads = set(self.bads.keys())
        for c in (exclude or []):
            bads.update(range(*self.section_pos[c]))
        mask = bad_mask(self.__size, bads)
        weights = self._bias_array() if normalized else np.ones(self.__size)
        nrm, raw, _ = self.store.decay_sums(weights, mask)
        intra = 0.0 if normalized else 0
        if equals is None:
            for beg, end in self.section_pos.values():
                per_distance = nrm[beg:end] if normalized else raw[beg:end]
                if end > beg:
                    inside = 2 * per_distance[1:].sum() + (per_distance[0] if diagonal else 0)
                    intra += float(inside) if normalized else int(inside)
        else:
            # chromosomes merged with their successor (hic_data.py:281-290): a group of consecutive
            # chromosomes is one square block of the matrix; its sum is the normalised sum with
            # everything outside the block masked (K3), minus its diagonal when that is left out
            to_skip, c_prev = set(), ''
            for c in self.chromosomes:
                if equals(c, c_prev):
                    to_skip.add(c_prev)
                c_prev = c
            ends = sorted(self.section_pos[c][1] for c in self.section_pos if c not in to_skip)
            diag = dict((beg, (nrm if normalized else raw)[beg]) for beg, end in self.section_pos.values()
                        if end > beg)
            beg = 0
            for end in ends:
                block = mask.copy()
                block[:beg] = 1
                block[end:] = 1
                inside = self.store.norm_sum(weights if normalized else None, block)
                if not diagonal:
                    inside -= sum(v for b, v in diag.items() if beg <= b < end)
                intra += float(inside) if normalized else int(round(inside))
                beg = end
        try:
            return float(intra) / self.sum(bias=self.bias if normalized else None, bads=bads)
        except ZeroDivisionError:
            return 0.

    def normalize_expected(self, **kwargs):
        self.expected = expected(self, bads=self.bads, **kwargs)

    def normalize_hic(self, iterations=0, max_dev=0.1, silent=False,
                      sqrt=False, factor=1):
        """Compute the ICE biases (hic_data.py:380-413); result in ``self.bias``.

        :param 0 iterations: number of iterations
        :param 0.1 max_dev: stop when the maximum deviation between row sums drops below
        :param False sqrt: use the square root of the biases
        :param 1 factor: final mean number of normalised interactions per cell
        """
        bias = iterative(self, iterations=iterations,
                         max_dev=max_dev, bads=self.bads,
                         verbose=not silent)
        values = bias.to_array()
        if sqrt:
            values = values ** 0.5
            bias = BinVector(values)
        if factor:
            if not silent:
                print('rescaling to factor %d' % factor)
                print('  - getting the sum of the matrix')
            norm_sum = self.sum(bias)
            if not silent:
                print('    => %.3f' % norm_sum)
                print('  - rescaling biases')
            target = (norm_sum / float(len(self) * len(self) * factor)) ** 0.5
            bias = BinVector(values * target)
        self.bias = bias

    # -- dense windows (the HiC_data part) ---------------------------------
    def _bias_array(self):
        bias = self.bias
        if isinstance(bias, BinVector):
            return bias.to_array()
        return np.array([bias[i] for i in range(self.__size)], dtype=np.float64)

    def _window(self, r0, r1, c0, c1, normalized):
        """Rows [r0, r1) x columns [c0, c1) as nested lists: ints, or count / bias[i] / bias[j]
        (one kernel over the resident cells, ``tb_window``, instead of a dict lookup per cell)."""
        r0, r1, c0, c1 = max(r0, 0), min(r1, self.__size), max(c0, 0), min(c1, self.__size)
        if normalized:
            return self.store.window(r0, r1, c0, c1, self._bias_array()).tolist()
        return np.rint(self.store.window(r0, r1, c0, c1)).astype(np.int64).tolist()

    @staticmethod
    def _numeric_focus(focus):
        """(start, end) or (start1, end1, start2, end2), 1-based inclusive -> 0-based half-open."""
        if len(focus) == 2:
            return focus[0] - 1, focus[1], focus[0] - 1, focus[1]
        beg1, end1, beg2, end2 = focus
        return beg1 - 1, end1, beg2 - 1, end2

    def _focus_simple(self, focus):
        """Window of write_matrix / yield_matrix / write_coord_table (hic_data.py:590-610, :1532-1551)
        as (start1, end1, start2, end2): bin numbers, one chromosome name, or a pair of names."""
        if not focus:
            return 0, len(self), 0, len(self)
        if isinstance(focus, tuple) and isinstance(focus[0], int):
            return self._numeric_focus(focus)
        first, second = focus if isinstance(focus, tuple) else (focus, focus)
        return self.section_pos[first] + self.section_pos[second]

    def _bp_range(self, spec):
        """'chr3:10000-20000' -> bin offsets inside the chromosome (base pairs // resolution)."""
        try:
            beg, end = [int(p) // self.resolution for p in spec.split(':')[1].split('-')]
        except ValueError:
            raise Exception('ERROR: should be in format "chr3:10000:20000"')
        return beg, end

    def _focus_coords(self, focus):
        """Window of get_matrix (hic_data.py:688-738) as (start1, start2, end1, end2); chromosome
        names may carry a range in base pairs.  As in the reference the range of the SECOND name of
        a pair is read from the first name and offset from the first window."""
        if not focus:
            return 0, 0, len(self), len(self)
        if isinstance(focus, tuple) and isinstance(focus[0], int):
            beg1, end1, beg2, end2 = self._numeric_focus(focus)
            return beg1, beg2, end1, end2
        if isinstance(focus, tuple):
            beg1, end1 = self.section_pos[focus[0].split(':')[0]]
            beg2, end2 = self.section_pos[focus[1].split(':')[0]]
            if ':' in focus[0]:
                lo, hi = self._bp_range(focus[0])
                beg1, end1 = beg1 + lo, beg1 + hi
            if ':' in focus[1]:
                lo, hi = self._bp_range(focus[0])
                beg2, end2 = beg1 + lo, beg1 + hi
            return beg1, beg2, end1, end2
        beg1, end1 = self.section_pos[focus.split(':')[0]]
        if ':' in focus:
            lo, hi = self._bp_range(focus)
            beg1, end1 = beg1 + lo, beg1 + hi
        return beg1, beg1, end1, end1

    def get_matrix(self, focus=None, diagonal=True, normalized=False, masked=False):
        """A window of the matrix as a list of lists (hic_data.py:632-686).

        :param None focus: ``(start, end)`` (1-based, inclusive), a chromosome name, or a tuple of two
        :param True diagonal: if False the diagonal is replaced by ones, or zeroes if normalized
        :param False normalized: divide the counts by the biases
        :param False masked: return a masked array built from the bad columns
        """
        if normalized and not self.bias:
            raise Exception('ERROR: experiment not normalized yet')
        start1, start2, end1, end2 = self._focus_coords(focus)
        # matrix[j - start1][i - start2] = self[i, j] / bias[i] / bias[j]: the window is computed with
        # i as its row (the same order of the two divisions) and transposed
        matrix = [list(col) for col in zip(*self._window(start2, end2, start1, end1, normalized))]
        if not matrix:
            matrix = [[] for _ in range(start1, end1)]
        if not diagonal and start1 == start2:
            for i in range(len(matrix)):
                if normalized:
                    matrix[i][i] = 0
                else:
                    matrix[i][i] = 1 if matrix[i][i] else 0
        if masked:
            from numpy import ma, zeros_like
            bads1 = [b - start1 for b in self.bads if start1 <= b < end1]
            bads2 = [b - start2 for b in self.bads if start2 <= b < end2]
            m = zeros_like(matrix)
            for bad1 in bads1:
                m[:, bad1] = 1
                for bad2 in bads2:
                    m[bad2, :] = 1
            matrix = ma.masked_array(matrix, m)
        return matrix

    def yield_matrix(self, focus=None, diagonal=True, normalized=False):
        """The window line by line; bad rows come out as null rows (hic_data.py:1514-1582)."""
        if normalized and not self.bias:
            raise Exception('ERROR: experiment not normalized yet')
        start1, end1, start2, end2 = self._focus_simple(focus)
        step = max(1, (1 << 24) // max(end1 - start1, 1))           # 128 MB of doubles per slab
        for lo in range(start2, end2, step):
            hi = min(lo + step, end2)
            rows = self._window(lo, hi, start1, end1, normalized)
            for i, line in zip(range(lo, hi), rows):
                if i in self.bads:
                    yield [0.0 if normalized else 0 for _ in range(start1, end1)]
                    continue
                if not (diagonal or start1 != start2) and start1 <= i < end1:
                    line[i - start1] = 0.0 if normalized else 0
                yield line

    def cooler_weights(self):
        """The ``weight`` column of the bins table that ``write_cooler(normalized=True)`` stores
        (_pytadbit/hic_data.py:572-574): ``1 / bias`` for a good bin with a positive bias, else 0
        (cooler multiplies by the weights where TADbit divides by the biases)."""
        bias = self._bias_array()
        good = np.ones(len(self), dtype=bool)
        if self.bads:
            good[[int(k) for k in self.bads]] = False
        with np.errstate(divide='ignore', invalid='ignore'):
            return np.where(good & (bias > 0), 1. / bias, 0.)

    def write_cooler(self, fname, normalized=False):
        """
        writes the hic_data to a cooler file (_pytadbit/hic_data.py:543-576).  The HDF5 container
        needs h5py, as in the reference (same message without it); writing it is not part of this
        package: the values it would hold are ``items()`` (upper triangle) and
        :meth:`cooler_weights`.

        :param False normalized: get normalized data
        """
        try:
            import h5py  # noqa: F401
        except ImportError:
            raise Exception('ERROR: cooler output is not available. Probably ' +
                            'you need to install h5py\n')
        if normalized and not self.bias:
            raise Exception('ERROR: data not normalized yet')
        raise NotImplementedError('cooler container output')

    def write_matrix(self, fname, focus=None, diagonal=True, normalized=False):
        """hic_data.py:578-630."""
        start1, end1, start2, end2 = self._focus_simple(focus)
        out = open(fname, 'w')
        out.write('# MASKED %s\n' % (','.join([str(k - start1)
                                               for k in list(self.bads.keys())
                                               if start1 <= k <= end1])))
        rownam = ['%s\t%d-%d' % (k[0],
                                 k[1] * self.resolution + 1,
                                 (k[1] + 1) * self.resolution)
                  for k in sorted(self.sections,
                                  key=lambda x: self.sections[x])
                  if start2 <= self.sections[k] < end2]
        if rownam:
            for line in self.yield_matrix(focus=focus, diagonal=diagonal,
                                          normalized=normalized):
                out.write(rownam.pop(0) + '\t' +
                          '\t'.join([str(i) for i in line]) + '\n')
        else:
            for line in self.yield_matrix(focus=focus, diagonal=diagonal,
                                          normalized=normalized):
                out.write('\t'.join([str(i) for i in line]) + '\n')
        out.close()

    def get_as_tuple(self):
        """Every cell of the matrix, row after row, as one tuple of ints (hic_data.py:446-449)."""
        out = []
        step = max(1, (1 << 24) // max(self.__size, 1))
        for lo in range(0, self.__size, step):
            for line in self._window(lo, min(lo + step, self.__size), 0, self.__size, False):
                out.extend(line)
        return tuple(out)

    def write_coord_table(self, fname, focus=None, diagonal=True,
                          normalized=False, format='BED'):
        """Coordinate table of the non-zero cells of a window, 'BED' or 'long-range' format
        (hic_data.py:451-540), streamed from ``yield_matrix``.  As in the reference the values of a
        row are consumed from its first column on while the partner names start after the row's own
        bin (its iterator over the line is not advanced to the diagonal)."""
        start1, end1, start2, end2 = self._focus_simple(focus)
        if format not in ('long-range', 'BED'):
            open(fname, 'w').close()                       # the reference has opened the file by then
            raise Exception('ERROR: format "%s" not found\n' % format)
        names = sorted(self.sections, key=lambda x: self.sections[x])
        inside = [k for k in names if start2 <= self.sections[k] < end2]
        span = [(k[0], k[1] * self.resolution, (k[1] + 1) * self.resolution) for k in inside]
        colnam = ['%s:%d-%d' % t for t in span]
        if format == 'long-range':
            rownam = colnam
            missing = 'ERROR: HiC data object should have genomic coordinates'
            pair_string = '%s\t%s\t%f\n' if normalized else '%s\t%s\t%d\n'
        else:
            rownam = ['%s\t%d\t%d' % t for t in span]
            missing = 'ERROR: Hi-C data object should have genomic coordinates'
            pair_string = '%s\t%s,%f\t%d\t.\n' if normalized else '%s\t%s,%d\t%d\t.\n'
        with open(fname, 'w') as out:
            if not rownam:
                raise Exception(missing)
            iter_rows = self.yield_matrix(focus=focus, diagonal=diagonal, normalized=normalized)
            count = 1
            for nrow, row in enumerate(rownam, 1):
                line = next(iter_rows)
                for col, val in zip(colnam[nrow:], line):
                    if not val:
                        continue
                    if format == 'long-range':
                        out.write(pair_string % (row, col, val))
                    else:
                        out.write(pair_string % (row, col, val, count))
                        count += 1

    # -- compartments -------------------------------------------------------
    def find_compartments(self, crms=None, savefig=None, savedata=None, savecorr=None, show=False,
                          suffix='', ev_index=None, rich_in_A=None, format='png', savedir=None,
                          max_ev=3, show_compartment_labels=False, smoothing_window=0, **kwargs):
        """A/B compartments per chromosome (hic_data.py:736-1031); see tadbit_b200/compartments.py."""
        from .compartments import find_compartments
        return find_compartments(self, crms=crms, savefig=savefig, savedata=savedata,
                                 savecorr=savecorr, show=show, suffix=suffix, ev_index=ev_index,
                                 rich_in_A=rich_in_A, format=format, savedir=savedir, max_ev=max_ev,
                                 show_compartment_labels=show_compartment_labels,
                                 smoothing_window=smoothing_window, **kwargs)

    def write_compartments(self, savedata, chroms=None, ev_nums=None):
        """hic_data.py:1486-1511."""
        from .compartments import write_compartments
        return write_compartments(self, savedata, chroms=chroms, ev_nums=ev_nums)

    def save_biases(self, fnam, protocol=None):
        """Pickle biases, decay and bad columns in the reference's format
        (hic_data.py:415-429): plain dicts, readable by TADbit's loaders."""
        bias = self.bias.to_dict() if isinstance(self.bias, BinVector) else self.bias
        with open(fnam, 'wb') as out:
            dump({'biases': bias,
                  'decay': self.expected,
                  'badcol': self.bads,
                  'resolution': self.resolution}, out,
                 protocol if protocol else HIGHEST_PROTOCOL)

    def load_biases(self, fnam, protocol=None):
        """hic_data.py:431-444."""
        with open(fnam, 'rb') as fh:
            biases = load(fh)
        if biases['resolution'] != self.resolution:
            raise Exception(('Error: resolution in Pickle (%d) does not match '
                             'the one of this HiC_data object (%d)') % (
                                 biases['resolution'], self.resolution))
        self.bias = biases['biases']
        self.expected = biases['decay']
        self.bads = biases['badcol']


def from_reference(hic_data, device=0):
    """Build a :class:`HiC_data` from a TADbit ``pytadbit.HiC_data`` (a dict subclass) or
    any ``{row * N + col: count}`` mapping that has ``len() == N`` semantics."""
    size = len(hic_data)
    items = dict.items(hic_data) if isinstance(hic_data, dict) else hic_data.items()
    new = HiC_data(items, size,
                   chromosomes=getattr(hic_data, 'chromosomes', None),
                   dict_sec=getattr(hic_data, 'sections', None),
                   resolution=getattr(hic_data, 'resolution', 1),
                   masked=getattr(hic_data, 'bads', None),
                   symmetricized=getattr(hic_data, 'symmetricized', False), device=device)
    # an object built with dict_sec={} carries the filled-in sections; keep its section bounds
    if getattr(hic_data, 'section_pos', None):
        new.section_pos = dict(hic_data.section_pos)
    return new


__all__ = ["HiC_data", "from_reference", "OrderedDict"]
