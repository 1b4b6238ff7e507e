"""The drop-in surface of HiC_data beyond the numeric kernels: reference-style constructor
(dict keyed row*N+col, either triangle), from_reference, item access, iteration, CSR export,
save_biases / load_biases, error behaviour."""
import pickle
from collections import OrderedDict

import numpy as np
import pytest

from _golden import load

pytestmark = pytest.mark.gpu


def _full_dict(case):
    n = case.size
    d = {}
    for r, c, v in zip(case.rows.tolist(), case.cols.tolist(), case.vals.tolist()):
        d[r * n + c] = v
        d[c * n + r] = v
    return d


def test_reference_style_constructor_matches_from_coo():
    import tadbit_b200
    case = load("chrT_20Kb_A")
    n = case.size
    full = _full_dict(case)
    upper_only = dict((k, v) for k, v in full.items() if k // n <= k % n)
    want = case.out["ice"][1]["bias"]
    for items in (full, upper_only, list(full.items())):
        hic = tadbit_b200.HiC_data(items, n, dict_sec={})
        assert len(hic) == n
        hic.filter_columns(silent=True)
        assert hic.bads == {22: 48}
        hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
        np.testing.assert_allclose([hic.bias[i] for i in range(n)], want, rtol=1e-12, atol=0)
        assert hic.section_pos == {None: (0, n)}
    # item access follows the reference (symmetric lookups, 0 for absent cells, IndexError)
    r, c, v = int(case.rows[7]), int(case.cols[7]), int(case.vals[7])
    assert hic[r, c] == v and hic[c, r] == v and hic[r * n + c] == v
    assert hic.get(0 * n + 0, 0) == full.get(0, 0)
    with pytest.raises(IndexError):
        hic[n, n + 5]
    assert dict(hic.items()) == full
    csr = hic.get_hic_data_as_csr()
    assert csr.shape == (n, n) and csr.sum() == sum(full.values())
    # cells can be set (hic_data.py:146-164); the matrix stays symmetric and the resident store is
    # rebuilt the next time it is used
    old = hic[0, 1]
    hic[0, 1] = old + 3
    hic[n - 1, 2] = 7
    assert hic[0, 1] == old + 3 and hic[1, 0] == old + 3 and hic[2, n - 1] == 7
    changed = dict(full)
    for key in (1, n):
        changed[key] = old + 3
    for key in ((n - 1) * n + 2, 2 * n + n - 1):
        changed[key] = 7
    assert dict(hic.items()) == changed
    hic.bads = {}
    assert hic.sum() == sum(changed.values())              # K3 over the rebuilt store
    stored, sums = hic.store.row_stats()
    assert int(sums[0]) == sum(v for k, v in changed.items() if k // n == 0)
    with pytest.raises(NotImplementedError):
        hic[0, 1] = 3.5
    with pytest.raises(IndexError):
        hic[n, n + 5] = 1


def test_from_reference_like_object_and_bias_pickle(tmp_path):
    import tadbit_b200

    class RefLike(dict):                       # stands in for pytadbit.HiC_data (a dict subclass)
        def __init__(self, items, size, chromosomes):
            super(RefLike, self).__init__(items)
            self._n = size
            self.chromosomes = chromosomes
            self.sections = dict(((c, p), i) for i, (c, p) in enumerate(
                (c, p) for c in chromosomes for p in range(chromosomes[c])))
            self.resolution = 20000
            self.bads = {}
            self.symmetricized = False

        def __len__(self):
            return self._n

    case = load("synth_3chrom")
    ref = RefLike(_full_dict(case), case.size, OrderedDict(case.chromosomes))
    hic = tadbit_b200.from_reference(ref)
    assert hic.chromosomes == case.chromosomes and hic.resolution == 20000
    hic.filter_columns(silent=True)
    assert hic.bads == case.bads(case.out["filter_columns"])
    hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
    hic.normalize_expected()
    fnam = str(tmp_path / "biases.pickle")
    hic.save_biases(fnam)
    blob = pickle.load(open(fnam, "rb"))
    assert sorted(blob) == ["badcol", "biases", "decay", "resolution"]
    assert type(blob["biases"]) is dict and len(blob["biases"]) == case.size   # plain dict for TADbit
    np.testing.assert_allclose([blob["biases"][i] for i in range(case.size)],
                               case.out["ice"][1]["bias"], rtol=1e-12, atol=0)
    assert blob["decay"] == dict((d, v) for d, v in case.out["expected"][0]["values"])
    other = tadbit_b200.HiC_data.from_coo(case.size, case.rows, case.cols, case.vals,
                                          chromosomes=case.chromosomes, resolution=20000)
    other.load_biases(fnam)
    assert other.bads == hic.bads and other.expected == hic.expected
    wrong = tadbit_b200.HiC_data.from_coo(case.size, case.rows, case.cols, case.vals, resolution=1)
    with pytest.raises(Exception, match="resolution"):
        wrong.load_biases(fnam)


def test_float_matrices_are_refused_not_silently_truncated():
    import tadbit_b200
    with pytest.raises(NotImplementedError):
        tadbit_b200.HiC_data({0: 1.5, 1: 2.0, 2: 2.0, 3: 1.0}, 2)
    hic = tadbit_b200.HiC_data({0: 1.0, 1: 2.0, 2: 2.0, 3: 4.0}, 2)     # integral floats are counts
    assert hic[0, 1] == 2 and hic.sum() == 9


def test_cis_trans_ratio_through_the_kernels():
    """HiC_data.cis_trans_ratio (hic_data.py:252-316) from tb_decay_sums / tb_norm_sum against the same
    host code around the oracle (tests/test_host_vs_reference.py pins that to the live reference)."""
    import tadbit_b200
    from _numpy_store import NumpyStore
    case = load("synth_3chrom")
    gpu = tadbit_b200.HiC_data.from_coo(case.size, case.rows, case.cols, case.vals, chromosomes=case.chromosomes)
    cpu = tadbit_b200.HiC_data.from_store(NumpyStore(case.size, case.rows, case.cols, case.vals,
                                                      chromosomes=case.chromosomes),
                                          chromosomes=case.chromosomes)
    first = list(case.chromosomes)[0]
    for hic in (gpu, cpu):
        hic.filter_columns(silent=True)
    assert gpu.bads == cpu.bads
    for kw in (dict(), dict(diagonal=False), dict(exclude=[first])):
        assert gpu.cis_trans_ratio(**kw) == cpu.cis_trans_ratio(**kw), kw
    for hic in (gpu, cpu):
        hic.normalize_hic(iterations=20, max_dev=1e-5, silent=True)
    for kw in (dict(normalized=True), dict(normalized=True, diagonal=False, exclude=[first])):
        a, b = gpu.cis_trans_ratio(**kw), cpu.cis_trans_ratio(**kw)
        assert abs(a - b) <= 1e-12 * abs(b), (kw, a, b)
    plain = tadbit_b200.HiC_data.from_coo(case.size, case.rows, case.cols, case.vals)
    assert np.isnan(plain.cis_trans_ratio())                  # no chromosomes: NaN, as the reference
    # chromosomes merged with their successor (one K3 pass per group of chromosomes)
    names = list(case.chromosomes)
    merged = lambda x, y: {x, y} == set(names[:2])                        # noqa: E731
    for kw in (dict(equals=merged), dict(equals=merged, diagonal=False)):
        assert gpu.cis_trans_ratio(**kw) == cpu.cis_trans_ratio(**kw), kw
    a, b = gpu.cis_trans_ratio(normalized=True, equals=merged), cpu.cis_trans_ratio(normalized=True, equals=merged)
    assert abs(a - b) <= 1e-12 * abs(b)
    assert gpu.cis_trans_ratio(equals=lambda x, y: True) == 1.0              # everything one group
