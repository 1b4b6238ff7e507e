"""Parity at sizes beyond the golden fixtures.

* a 20k-bin multi-chromosome synthetic matrix against the oracle (numpy restatement + C port)
  on the exported cells: mask bit-exact, biases 1e-10, expected bit-exact;
* the dense chr1 @10 kb workload (BASELINE configs[1], 3e8 contacts) and the genome-wide 5 kb
  workload of the bench (configs[2], 1.97e9 contacts) through size-independent properties: the two independent K1 implementations (compact stream vs CSR walk) agree, integer
  reductions are consistent with each other, the rescaled matrix sums to N^2."""
from collections import OrderedDict

import numpy as np
import pytest

import tadbit_oracle as orc

pytestmark = pytest.mark.gpu


def _mid():
    import tadbit_b200
    chroms = OrderedDict([("c1", 9000), ("c2", 6500), ("c3", 4500)])
    n = sum(chroms.values())
    params = tadbit_b200.SynthParams(seed=11, cis_scale=900.0, cis_alpha=1.08, trans_mean=0.004,
                                     vis_sigma=0.35, low_frac=0.03, low_scale=0.02, empty_frac=0.01)
    store = tadbit_b200.ContactStore.synthetic(n, params, chromosomes=chroms)
    return tadbit_b200.HiC_data.from_store(store, chromosomes=chroms), chroms, n


def test_mid_size_against_oracle():
    hic, chroms, n = _mid()
    lay = hic.store.layout()
    assert lay["encodings"]["T"] > 0 and lay["encodings"]["D8"] > 0       # both K1 kernels in play
    r, c, v = hic.store.export_upper()
    m = orc.OracleMatrix(n, r, c, v, chromosomes=chroms)
    hic.filter_columns(silent=True)
    bads = orc.filter_columns(m, silent=True)
    assert hic.bads == bads
    hic.normalize_hic(iterations=100, max_dev=1e-6, silent=True, factor=None)
    bias, trace, _, _ = orc.iterative_fast(m, bads, 100, 1e-6)
    assert len(hic.last_trace) == len(trace)
    np.testing.assert_allclose(hic.bias.to_array(), bias, rtol=1e-10, atol=0)
    hic.normalize_expected()
    assert hic.expected == orc.expected(m, bads=bads)
    nnz, rsum = hic.store.row_stats()
    onnz, orsum = orc.row_stats(m)
    assert np.array_equal(nnz, onnz) and np.array_equal(rsum, orsum)
    assert hic.sum() == int(orc.matrix_sum(m, None, bads))


def test_full_size_properties_c2():
    import tadbit_b200
    from tadbit_b200 import workloads
    wl = workloads.workload("c2")
    n = wl["size"]
    store = tadbit_b200.ContactStore.synthetic(n, wl["params"], chromosomes=wl["chromosomes"])
    hic = tadbit_b200.HiC_data.from_store(store, chromosomes=wl["chromosomes"])
    assert store.info()["nnz_upper"] > 2.5e8
    hic.filter_columns(silent=True, by_mean=False)
    nnz, rsum = store.row_stats()
    assert int(nnz.sum()) == store.info()["nnz_full"]
    # zero-count filter re-derived from the exact per-row counts
    assert hic.bads == dict((int(i), True) for i in np.flatnonzero(n - nnz > int(n * 99.0 / 100)))
    # integer reductions agree with each other: raw sum, row sums, per-distance sums
    assert hic.sum(bads={-1: True}) == int(rsum.sum())
    diag = store.diag_sums(np.zeros(n, dtype=np.uint8), n)
    assert 2 * int(diag.sum()) - int(diag[0]) == int(rsum.sum())
    # K1 twice: compact stream vs plain CSR walk
    hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
    compact, passes = hic.bias.to_array().copy(), len(hic.last_trace)
    store.set_option("rowsum", "csr")
    hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
    assert len(hic.last_trace) == passes
    np.testing.assert_allclose(hic.bias.to_array(), compact, rtol=1e-11, atol=0)
    store.set_option("rowsum", "compact")
    # factor=1 rescales so that the normalised matrix sums to N^2 (hic_data.py:406-412)
    assert abs(hic.sum(hic.bias) / float(n * n) - 1.0) < 1e-9


def test_full_size_properties_c3():
    """The bench workload (BASELINE configs[2], genome-wide 5 kb, 1.97e9 contacts) at full size:
    exact integer identities, and the ICE loop over the compact stream (dense panels + two-section
    sparse tiles) against the same loop over the plain CSR."""
    import torch
    import tadbit_b200
    from tadbit_b200 import workloads
    from tadbit_b200._capi import trim_memory
    trim_memory()                                   # hand cached blocks of earlier tests back first
    if torch.cuda.mem_get_info(0)[0] < 90 * (1 << 30):
        pytest.skip("needs ~90 GB of free device memory")
    wl = workloads.workload("c3")
    n = wl["size"]
    store = tadbit_b200.ContactStore.synthetic(n, wl["params"], chromosomes=wl["chromosomes"])
    try:
        hic = tadbit_b200.HiC_data.from_store(store, chromosomes=wl["chromosomes"])
        info = store.info()
        assert info["nnz_upper"] > 1.9e9
        lay = store.layout()
        assert lay["tile_cells"] > 1e9 and lay["encodings"]["D8"] > 1e6
        hic.filter_columns(silent=True, by_mean=False)
        nnz, rsum = store.row_stats()
        assert int(nnz.sum()) == info["nnz_full"]
        assert hic.bads == dict((int(i), True) for i in np.flatnonzero(n - nnz > int(n * 99.0 / 100)))
        assert hic.sum(bads={-1: True}) == int(rsum.sum())
        hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
        compact, passes = hic.bias.to_array().copy(), len(hic.last_trace)
        assert np.isfinite(compact).all() and passes > 10
        store.set_option("rowsum", "csr")
        hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
        assert len(hic.last_trace) == passes
        np.testing.assert_allclose(hic.bias.to_array(), compact, rtol=1e-11, atol=0)
        store.set_option("rowsum", "compact")
        assert abs(hic.sum(hic.bias) / float(n * n) - 1.0) < 1e-9
    finally:
        store.close()
