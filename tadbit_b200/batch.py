"""Many matrices normalised at once (BASELINE configs[4]).

The reference normalises experiments one after the other
(``Experiment.normalize_hic``, _pytadbit/experiment.py:704-748).  Small matrices leave a
B200 idle between kernel launches, so here the resident stores are concatenated
(device-to-device) into one block-diagonal store and every ICE pass is one launch for all of
them; statistics, stop test and bias update stay per matrix (``tb_ice_batch``).
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import bad_mask, check, lib, ptr
from .normalize_hic import TRACE_FORMAT
from .store import ContactStore
from .vectors import BinVector


def batch_store(stores, device=0):
    """Block-diagonal store over whole (unsharded) ``ContactStore`` objects of one device."""
    handles = (C.c_void_p * len(stores))(*[s._h for s in stores])
    out = C.c_void_p()
    check(lib().tb_store_create_batch(len(stores), handles, device, C.byref(out)))
    return ContactStore(out, sum(s.size for s in stores))


def ice_batch(store, sizes, bad=None, iterations=0, max_dev=1e-5):
    """-> (bias float64[sum(sizes)], list of per-matrix traces, passes int32[len(sizes)])."""
    n = len(sizes)
    m, mp = store._mask(bad)
    bias = np.empty(store.size, dtype=np.float64)
    trace = np.zeros((n, iterations + 1, 4), dtype=np.float64)
    passes = np.zeros(n, dtype=np.int32)
    check(lib().tb_ice_batch(store._h, mp, int(iterations), float(max_dev), ptr(bias, C.c_double),
                             ptr(trace, C.c_double), ptr(passes, C.c_int32)))
    traces = [trace[k, :max(int(passes[k]), 0) if iterations > 0 else 0] for k in range(n)]
    return bias, traces, passes


def normalize_hic_batch(hics, iterations=0, max_dev=0.1, silent=False, sqrt=False, factor=1,
                        device=0):
    """Same result as ``for h in hics: h.normalize_hic(iterations, max_dev, silent, sqrt,
    factor)`` (each ``h.bias`` is set), with all ICE loops run as one batch on the GPU.

    As in the loop, a matrix that cannot be normalised raises ``ZeroDivisionError`` when its
    turn comes: the matrices before it have their ``bias`` set, the ones after it are left
    untouched (their loops did run on the device; the results are dropped)."""
    sizes = [len(h) for h in hics]
    store = batch_store([h.store for h in hics], device=device)
    try:
        bad = np.concatenate([bad_mask(len(h), h.bads) for h in hics])
        bias, traces, passes = ice_batch(store, sizes, bad, iterations=iterations, max_dev=max_dev)
    finally:
        store.close()
    off = 0
    for k, h in enumerate(hics):
        if not silent:
            print('iterative correction')
            print('  - copying matrix')
        if passes[k] == -1:                      # normalize_hic.py:207-208, before the third line
            raise ZeroDivisionError('ERROR: normalization failed, all bad columns')
        if not silent:
            print('  - computing biases')
        if passes[k] == -2:                      # _updateDB with meanS == 0 (normalize_hic.py:148)
            raise ZeroDivisionError('float division by zero')
        if not silent:
            for it, (smin, mean_s, smax, dev) in enumerate(traces[k]):
                print(TRACE_FORMAT % (smin, mean_s, smax, it, dev))
        values = bias[off:off + sizes[k]].copy()
        off += sizes[k]
        h.last_trace = traces[k]
        if sqrt:
            values = values ** 0.5
        vec = BinVector(values)
        if factor:
            if not silent:
                print('rescaling to factor %d' % factor)
                print('  - getting the sum of the matrix')
            norm_sum = h.sum(vec)
            if not silent:
                print('    => %.3f' % norm_sum)
                print('  - rescaling biases')
            target = (norm_sum / float(len(h) * len(h) * factor)) ** 0.5
            vec = BinVector(values * target)
        h.bias = vec
    return [h.bias for h in hics]
