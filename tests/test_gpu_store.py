"""Store construction: row-major sorted input (column-sort fast path), shuffled input (general
64-bit sort path) and the device generator's export give the same resident matrix."""
import numpy as np
import pytest

from _golden import load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["synth_3chrom", "chrT_20Kb_A", "stored_zero_60", "synth_sparse_600"])
def test_sorted_and_shuffled_input_agree(name):
    import tadbit_b200
    case = load(name)
    r, c, v = case.rows.astype(np.int32), case.cols.astype(np.int32), case.vals.astype(np.int32)
    a = tadbit_b200.ContactStore.from_coo(case.size, r, c, v)                 # sorted: fast path
    perm = np.random.default_rng(3).permutation(r.size)
    b = tadbit_b200.ContactStore.from_coo(case.size, r[perm], c[perm], v[perm])   # general path
    ea, eb = a.export_upper(), b.export_upper()
    for x, y in zip(ea, eb):
        assert np.array_equal(x, y)
    order = np.lexsort((c, r))
    assert np.array_equal(ea[0], r[order]) and np.array_equal(ea[1], c[order])
    assert np.array_equal(ea[2], v[order])
    na, sa = a.row_stats()
    nb, sb = b.row_stats()
    assert np.array_equal(na, nb) and np.array_equal(sa, sb)
    ba, _ = a.ice(None, 30, 1e-6)
    bb, _ = b.ice(None, 30, 1e-6)
    assert np.array_equal(ba, bb)


def test_invalid_input_is_rejected():
    import tadbit_b200
    with pytest.raises(ValueError, match="row > col"):
        tadbit_b200.ContactStore.from_coo(5, [3], [1], [1])
    with pytest.raises(ValueError, match="duplicate"):
        tadbit_b200.ContactStore.from_coo(5, [1, 0, 1], [2, 4, 2], [1, 1, 1])
    with pytest.raises(ValueError):
        tadbit_b200.ContactStore.from_coo(5, [1], [7], [1])


def test_large_roundtrip_export_rebuild():
    import tadbit_b200
    from tadbit_b200 import workloads
    wl = workloads.workload("small")
    s = tadbit_b200.ContactStore.synthetic(wl["size"], wl["params"], chromosomes=wl["chromosomes"])
    cells = s.export_upper()
    t = tadbit_b200.ContactStore.from_coo(wl["size"], *cells, chromosomes=wl["chromosomes"])
    for x, y in zip(cells, t.export_upper()):
        assert np.array_equal(x, y)
    b1, t1 = s.ice(None, 100, 1e-5)
    b2, t2 = t.ice(None, 100, 1e-5)
    assert np.array_equal(b1, b2) and len(t1) == len(t2)


@pytest.mark.gpu
def test_device_generator_agrees_with_cpu_generator():
    """Two independent implementations of the synthetic model (synth.cu on the device,
    oracle/synth_port.c on the host, the input of bench.py's CPU legs): same cells except where
    single-precision libm and the device round a Poisson threshold differently."""
    from collections import OrderedDict
    import tadbit_oracle as orc
    import tadbit_b200
    chroms = OrderedDict([("chrA", 1400), ("chrB", 900), ("chrC", 500)])
    n = 2800
    p = tadbit_b200.SynthParams(seed=20261019, cis_scale=600.0, cis_alpha=1.08, trans_mean=0.02,
                                vis_sigma=0.35, low_frac=0.03, low_scale=0.02, empty_frac=0.01)
    store = tadbit_b200.ContactStore.synthetic(n, p, chromosomes=chroms)
    r, c, v = store.export_upper()
    store.close()
    r2, c2, v2 = orc.synthetic_block(n, p, chroms)
    a = dict(zip((r.astype(np.int64) * n + c).tolist(), v.tolist()))
    b = dict(zip((r2.astype(np.int64) * n + c2).tolist(), v2.tolist()))
    differing = sum(1 for k in a.keys() | b.keys() if a.get(k, 0) != b.get(k, 0))
    assert differing <= 1e-4 * len(a), (differing, len(a))
    assert abs(int(v.sum()) - int(v2.sum())) <= 1e-4 * int(v.sum())
