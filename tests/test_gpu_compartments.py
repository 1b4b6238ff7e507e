"""N4 (compartments) through the CUDA kernels and the C-ABI (tb_corr_build / tb_corr_get /
tb_corr_set / tb_corr_eig): against the vectors of the UNMODIFIED reference's find_compartments
(tests/golden/compartments_*: its observed / expected matrix, its correlation matrix, its
eigenpairs and break points), with both forms of the X X^T kernel (fp64 tensor tiles and the plain
parity kernel), and at 2 600 bins against the oracle (numpy.corrcoef + LAPACK)."""
from collections import OrderedDict

import numpy as np
import pytest

import tadbit_oracle as orc
from _compartment_cases import compartment_case, sections_of
from test_oracle_compartments import (CASES, load_compartments, same_up_to_sign, section_bounds,
                                      with_bad_bins)

pytestmark = pytest.mark.gpu


def _hic(rec, z, chroms, bads, expected):
    import tadbit_b200
    from tadbit_b200.vectors import BinVector
    kw = dict(chromosomes=chroms, dict_sec=sections_of(chroms)) if chroms else dict(dict_sec={})
    hic = tadbit_b200.HiC_data.from_coo(rec["n"], z["rows"], z["cols"], z["vals"],
                                        resolution=rec["resolution"], **kw)
    return hic, BinVector(np.array(z["bias"])), bads, expected


@pytest.mark.parametrize("gemm", ["dmma", "simple"])
@pytest.mark.parametrize("name", CASES)
def test_kernels_against_the_reference_vectors(name, gemm):
    from tadbit_b200._capi import bad_mask
    rec, z, chroms, bads, expected = load_compartments(name)
    hic, bias, bads, expected = _hic(rec, z, chroms, bads, expected)
    store = hic.store
    store.set_option("corr_gemm", gemm)
    mask = bad_mask(rec["n"], bads)
    bounds = section_bounds(rec, chroms)
    for k, sec in enumerate(rec["sections"]):
        beg, end = bounds[sec]
        exp_arr = np.array([expected.get(d, 1.0) for d in range(end - beg)])
        n = store.corr_build(beg, end, bias.to_array(), mask, exp_arr, stage=0)
        want = z["obsexp_%d" % k]
        assert n == want.shape[0]
        got = store.corr_get()
        assert np.array_equal(got, want)                   # the reference's divisions, bit for bit
        assert store.corr_build(beg, end, bias.to_array(), mask, exp_arr, stage=1) == n
        corr = store.corr_get()
        np.testing.assert_allclose(corr, z["corr_%d" % k], rtol=0, atol=1e-12)
        assert np.array_equal(corr, corr.T)
        assert store.corr_timing()["gemm"] == gemm
        kk = rec["call"]["max_ev"]
        vals, vecs, info = store.corr_eig(kk)
        assert info["converged"]
        want_vals, want_vecs = z["evals_%d" % k], z["evecs_%d" % k]
        np.testing.assert_allclose(vals, want_vals, rtol=1e-10, atol=1e-12)
        for i in range(len(want_vecs)):
            if abs(want_vals[i]) < 1e-9 * abs(want_vals[0]):
                continue
            same_up_to_sign(with_bad_bins(vecs[i], beg, end, bads), want_vecs[i], 1e-8)


@pytest.mark.parametrize("name", CASES)
def test_find_compartments_against_the_reference(name, tmp_path):
    rec, z, chroms, bads, expected = load_compartments(name)
    hic, bias, bads, expected = _hic(rec, z, chroms, bads, expected)
    if name in ("chrT_20Kb_A", "synth_4chrom"):
        pass                                               # from scratch: the call filters / normalises itself
    else:
        hic.bads, hic.bias, hic.expected = bads, bias, expected
    with _collect() as caught:
        firsts, stats = hic.find_compartments(savedata=str(tmp_path / "c.tsv"),
                                              savecorr=str(tmp_path / "corr"),
                                              savedir=str(tmp_path / "ev"), **rec["call"])
    assert dict((int(k), v) for k, v in hic.bads.items()).keys() == bads.keys()
    np.testing.assert_allclose(hic.bias.to_array(), z["bias"], rtol=1e-12)
    assert dict(hic.expected) == expected
    assert dict((str(k), v) for k, v in hic.compartments.items()) == rec["compartments"]
    assert sorted(m for m in caught() if "Chromosome" in m) == rec["warnings"]
    bounds = section_bounds(rec, chroms)
    for k, sec in enumerate(rec["sections"]):
        key = None if sec == "None" else sec
        vals, vecs = firsts[key]
        np.testing.assert_allclose(vals, z["evals_%d" % k], rtol=1e-10, atol=1e-12)
        for i, want in enumerate(z["evecs_%d" % k]):
            if abs(z["evals_%d" % k][i]) < 1e-9 * abs(z["evals_%d" % k][0]):
                continue
            same_up_to_sign(vecs[i], want, 1e-8)
        assert (tmp_path / "corr" / ("%s_corr-matrix.tsv" % key)).exists()
        assert (tmp_path / "ev" / ("%s_EigVect1.tsv" % key)).exists()
        t = hic.compartment_timing[key]
        assert t["n"] == z["corr_%d" % k].shape[0] and t["gemm_ms"] > 0 and t["eig_ms"] > 0
    assert (tmp_path / "c.tsv").read_text().count("\n") >= 1


class _collect(object):
    """warnings raised inside the block, as strings"""

    def __enter__(self):
        import warnings
        self._cm = warnings.catch_warnings(record=True)
        self._log = self._cm.__enter__()
        warnings.simplefilter("always")
        return lambda: [str(w.message) for w in self._log]

    def __exit__(self, *exc):
        return self._cm.__exit__(*exc)


def test_one_chromosome_of_2600_bins_against_the_oracle():
    """21 x 21 tiles of 128 with a ragged edge, k not a multiple of the pipeline depth, 25 bad bins."""
    import tadbit_b200
    from tadbit_b200._capi import bad_mask
    case = compartment_case(4242, sizes=(2600,), n_dead=25, block=40, depth=400.0)
    n = case["n"]
    hic = tadbit_b200.HiC_data.from_coo(n, case["rows"], case["cols"], case["vals"],
                                        chromosomes=case["chromosomes"],
                                        dict_sec=sections_of(case["chromosomes"]), resolution=10000)
    hic.filter_columns(silent=True)
    hic.normalize_hic(iterations=20, max_dev=1e-5, silent=True)
    hic.normalize_expected()
    assert set(case["dead"].tolist()) <= set(hic.bads)
    m = orc.OracleMatrix(n, case["rows"], case["cols"], case["vals"], chromosomes=case["chromosomes"])
    exp_arr = np.array([hic.expected[d] for d in range(n)])
    bias = hic.bias.to_array()
    mask = bad_mask(n, hic.bads)
    want_oe = orc.obs_exp_matrix(m, 0, n, bias, hic.bads, exp_arr)
    want_corr = orc.correlation_matrix(want_oe)
    want_vals, want_vecs = orc.largest_eigenpairs(want_corr, 3)
    results = {}
    for gemm in ("dmma", "simple"):
        hic.store.set_option("corr_gemm", gemm)
        assert hic.store.corr_build(0, n, bias, mask, exp_arr, stage=0) == want_oe.shape[0]
        assert np.array_equal(hic.store.corr_get(), want_oe)
        hic.store.corr_build(0, n, bias, mask, exp_arr, stage=1)
        corr = hic.store.corr_get()
        np.testing.assert_allclose(corr, want_corr, rtol=0, atol=1e-11)
        results[gemm] = corr
        vals, vecs, info = hic.store.corr_eig(3)
        assert info["converged"] and info["steps"] < 400
        np.testing.assert_allclose(vals, want_vals, rtol=1e-10)
        for i in range(3):
            same_up_to_sign(vecs[i], want_vecs[i], 1e-8)
    np.testing.assert_allclose(results["dmma"], results["simple"], rtol=0, atol=1e-13)
    hic.store.set_option("corr_gemm", "dmma")
    firsts, _ = hic.find_compartments()
    got = hic.compartments["chr1"]
    full = with_bad_bins(want_vecs[0], 0, n, hic.bads)
    # the break points are where the oracle's first eigenvector changes sign
    good = np.flatnonzero(~np.isnan(full))
    flips = [int(good[i]) for i in range(len(good) - 1) if full[good[i]] * full[good[i + 1]] < 0]
    assert [c["end"] for c in got[:-1]] == flips
    # and they follow the planted checkerboard: compartments of ~40 bins, not noise
    assert 20 < len(got) < 140


def test_upload_smoothing_round_trip_and_errors():
    import tadbit_b200
    from tadbit_b200._capi import B200Error
    case = compartment_case(7, sizes=(90, 30), n_dead=2)
    hic = tadbit_b200.HiC_data.from_coo(case["n"], case["rows"], case["cols"], case["vals"],
                                        chromosomes=case["chromosomes"],
                                        dict_sec=sections_of(case["chromosomes"]))
    with pytest.raises((ValueError, B200Error)):
        hic.store._corr_n = 5
        hic.store.corr_eig(2)                              # nothing resident yet
    rng = np.random.default_rng(3)
    a = rng.normal(size=(150, 150))
    a = a + a.T
    hic.store.corr_set(a)
    assert np.array_equal(hic.store.corr_get(), a)
    vals, vecs, info = hic.store.corr_eig(4)
    want_vals, want_vecs = orc.largest_eigenpairs(a, 4)
    np.testing.assert_allclose(vals, want_vals, rtol=1e-10)
    for i in range(4):
        same_up_to_sign(vecs[i], want_vecs[i], 1e-8)
    # every eigenpair of a small matrix (k >= n)
    b = a[:5, :5]
    hic.store.corr_set(b)
    vals, vecs, info = hic.store.corr_eig(5)
    np.testing.assert_allclose(vals, np.sort(np.linalg.eigvalsh(b))[::-1], rtol=1e-10)
    # smoothing through the host, as find_compartments does it, against the oracle
    from scipy.ndimage import median_filter
    hic.find_compartments(smoothing_window=3)
    m = orc.OracleMatrix(case["n"], case["rows"], case["cols"], case["vals"], chromosomes=case["chromosomes"])
    exp_arr = np.array([hic.expected[d] for d in range(90)])
    corr = orc.correlation_matrix(orc.obs_exp_matrix(m, 0, 90, hic.bias.to_array(), hic.bads, exp_arr))
    _, want_vecs = orc.largest_eigenpairs(median_filter(corr, size=3), 3)
    got = hic.compartments["chr1"]
    assert [c["end"] for c in got] == [c["end"] for c in _shift(orc.compartment_breaks(want_vecs[0]), hic.bads, 0, 90)]


def _shift(breaks, bads, beg, end):
    for b in [k - beg for k in sorted(bads) if beg <= k <= end]:
        for brk in breaks:
            if brk["start"] >= b:
                brk["start"] += 1
                brk["end"] += 1
            else:
                brk["end"] += brk["end"] > b
    return breaks
