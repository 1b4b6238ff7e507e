"""Array-backed stand-in for the reference's ``{bin: float}`` dicts.

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
