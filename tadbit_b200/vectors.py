"""Array-backed stand-ins for the reference's ``{bin: float}`` and ``{str(i): {str(j): value}}`` dicts.

``HiC_data.bias`` in the reference is a dict with one float per bin
(_pytadbit/hic_data.py:413).  At 3.1 M bins building that dict costs more than the whole
GPU normalisation, so the values stay in one float64 array behind the Mapping protocol
(``bias[i]``, ``len``, iteration over keys, ``items()``, ``==`` against a dict).
``to_dict()`` gives the plain dict (used by ``save_biases`` so that pickles stay
readable by TADbit).
"""
from collections.abc import Mapping

import numpy as np


class BinVector(Mapping):
    __slots__ = ("_a",)

    def __init__(self, values):
        self._a = np.ascontiguousarray(values, dtype=np.float64)

    def __getitem__(self, key):
        try:
            k = int(key)
        except (TypeError, ValueError):
            raise KeyError(key)
        if k != key or k < 0 or k >= self._a.size:
            raise KeyError(key)
        return float(self._a[k])

    def __len__(self):
        return int(self._a.size)

    def __iter__(self):
        return iter(range(self._a.size))

    def __contains__(self, key):
        try:
            return 0 <= key < self._a.size and int(key) == key
        except TypeError:
            return False

    def values(self):
        return self._a.tolist()

    def items(self):
        return zip(range(self._a.size), self._a.tolist())

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def to_array(self):
        return self._a

    def to_dict(self):
        return dict(enumerate(self._a.tolist()))

    def __reduce__(self):
        return (BinVector, (self._a,))

    def __repr__(self):
        return "BinVector(%d bins)" % self._a.size


class PairTable(Mapping):
    """``Experiment._zscores`` of the reference, ``{str(i): {str(j): value}}`` for the pairs i < j
    (_pytadbit/experiment.py:751-799), over three arrays sorted row-major.  Building the dictionary
    of dictionaries costs about a microsecond per pair in Python -- more than computing the values on
    the GPU; here ``table[str(i)]`` builds the dictionary of one row when it is asked for, and
    everything that can work on the arrays (``write_interaction_pairs``) does."""
    __slots__ = ("rows", "cols", "vals", "_urows", "_starts", "_ends")

    def __init__(self, rows, cols, vals):
        self.rows = np.ascontiguousarray(rows, dtype=np.int64)
        self.cols = np.ascontiguousarray(cols, dtype=np.int64)
        self.vals = np.ascontiguousarray(vals)
        if self.rows.size and np.any(np.diff(self.rows) < 0):
            raise ValueError("pairs must be sorted by row")
        self._urows, self._starts = np.unique(self.rows, return_index=True)
        self._ends = np.append(self._starts[1:], self.rows.size)

    def _slot(self, key):
        if not isinstance(key, str):
            raise KeyError(key)
        try:
            i = int(key)
        except ValueError:
            raise KeyError(key)
        p = int(np.searchsorted(self._urows, i))
        if str(i) != key or p >= self._urows.size or self._urows[p] != i:
            raise KeyError(key)
        return p

    def __getitem__(self, key):
        p = self._slot(key)
        beg, end = int(self._starts[p]), int(self._ends[p])
        return dict(zip(map(str, self.cols[beg:end].tolist()), self.vals[beg:end].tolist()))

    def __contains__(self, key):
        try:
            self._slot(key)
        except KeyError:
            return False
        return True

    def __len__(self):
        return int(self._urows.size)

    def __iter__(self):
        return iter(map(str, self._urows.tolist()))

    def to_dict(self):
        return dict((k, self[k]) for k in self)

    def __repr__(self):
        return "PairTable(%d pairs in %d rows)" % (self.rows.size, self._urows.size)
