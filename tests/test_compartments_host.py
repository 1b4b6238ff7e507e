"""N4 (compartments), host layer on the CPU: ``tadbit_b200.HiC_data.find_compartments`` around the
store double (tests/_numpy_store.py: obs/exp, corrcoef and the eigenpairs answered by the oracle)
next to the UNMODIFIED reference's find_compartments on the same seeded matrices -- implicit
filter / expected / one-pass ICE, skipped chromosomes and their warnings, break points, NaN
re-insertion, ev_index, rich_in_A orientation (Spearman), crms, smoothing, and the three output
files.  The device side of the same calls is tests/test_gpu_compartments.py."""
import contextlib
import io
import os
import sys
import warnings

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ref_shim  # noqa: E402
from _compartment_cases import compartment_case, full_dict, rich_track, sections_of  # noqa: E402
from _numpy_store import NumpyStore  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def _pair(case, resolution=10000):
    import tadbit_b200
    ns = ref_shim.load()
    chroms = case["chromosomes"]
    sec = sections_of(chroms)
    ref = ns.HiC_data(full_dict(case), case["n"], chromosomes=chroms, dict_sec=dict(sec),
                      resolution=resolution)
    store = NumpyStore(case["n"], case["rows"], case["cols"], case["vals"], chromosomes=chroms)
    ours = tadbit_b200.HiC_data.from_store(store, chromosomes=chroms, dict_sec=dict(sec),
                                           resolution=resolution)
    return ref, ours


def _call(hic, **kw):
    out = io.StringIO()
    with warnings.catch_warnings(record=True) as caught, contextlib.redirect_stdout(out):
        warnings.simplefilter("always")
        res = hic.find_compartments(**kw)
    msgs = sorted(str(w.message) for w in caught
                  if "Chromosome" in str(w.message) or "filtered out" in str(w.message))
    return res, msgs, out.getvalue()


def _same_vectors(ours, ref, tol=1e-8):
    a, b = np.asarray(ours, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    assert min(np.abs(a[ok] - b[ok]).max(), np.abs(a[ok] + b[ok]).max()) < tol


def _compare(res_o, res_r, ours, ref, oriented=False):
    (f_o, s_o), (f_r, s_r) = res_o, res_r
    assert ours.compartments.keys() == ref.compartments.keys()
    assert list(f_o.keys()) == list(f_r.keys())
    for sec in f_r:
        ev_o, vec_o = f_o[sec]
        ev_r, vec_r = f_r[sec]
        np.testing.assert_allclose(ev_o, ev_r, rtol=1e-10, atol=1e-12)
        assert len(vec_o) == len(vec_r)
        for k, (a, b) in enumerate(zip(vec_o, vec_r)):
            if ev_r[k] < 1e-9 * ev_r[0]:
                continue                      # null space of a tiny chromosome: any basis will do
            if oriented and s_r[sec] is not None:
                a_, b_ = np.asarray(a), np.asarray(b)
                ok = ~np.isnan(b_)
                if k == 0:
                    np.testing.assert_allclose(a_[ok], b_[ok], rtol=0, atol=1e-8)
                    continue
            _same_vectors(a, b)
    for sec in ref.compartments:
        got, want = ours.compartments[sec], ref.compartments[sec]
        assert [(c["start"], c["end"]) for c in got] == [(c["start"], c["end"]) for c in want], sec
        for cg, cw in zip(got, want):
            assert cg.keys() == cw.keys()
            if "dens" in cw:
                np.testing.assert_allclose(cg["dens"], cw["dens"], rtol=1e-12, equal_nan=True)
            if "type" in cw and oriented and s_r[sec] is not None:
                assert cg["type"] == cw["type"]
    assert s_o.keys() == s_r.keys()
    for sec in s_r:
        if s_r[sec] is None:
            assert s_o[sec] is None
        else:
            assert abs(abs(s_o[sec]) - abs(s_r[sec])) < 1e-9


@pytest.mark.parametrize("seed", range(6))
def test_defaults_from_scratch(seed):
    """No bads, expected or biases yet: the call filters (zero counts only), computes expected and
    one ICE pass itself, as the reference does (hic_data.py:808-825)."""
    case = compartment_case(100 + seed)
    ref, ours = _pair(case)
    res_r, warn_r, _ = _call(ref)
    res_o, warn_o, _ = _call(ours)
    assert ours.bads == ref.bads and set(ours.bads) == set(case["dead"].tolist())
    assert dict(ours.expected) == dict(ref.expected)
    np.testing.assert_allclose(ours.bias.to_array(), [ref.bias[i] for i in range(case["n"])], rtol=1e-12)
    assert warn_o == warn_r and len(warn_r) == 1          # the 1-bin chromosome "is probably MT"
    _compare(res_o, res_r, ours, ref)
    # chr3 has exactly max_ev good bins: every eigenpair, by value (scipy's dense fall-back)
    assert len(res_o[0]["chr3"][0]) == 3


@pytest.mark.parametrize("seed", range(4))
def test_after_the_full_pipeline_with_tracks_and_files(seed, tmp_path):
    case = compartment_case(200 + seed, sizes=(80, 50, 30, 1), n_dead=4)
    ref, ours = _pair(case)
    track = rich_track(case, seed)
    results = []
    for tag, hic in (("ref", ref), ("our", ours)):
        hic.filter_columns(silent=True)
        hic.normalize_hic(iterations=30, max_dev=1e-6, silent=True)
        hic.normalize_expected()
        d = tmp_path / tag
        d.mkdir()
        kw = dict(rich_in_A=dict((c, dict(v)) for c, v in track.items()), ev_index=[1, 2, 1, 1],
                  savedata=str(d / "cmp.tsv"), savecorr=str(d / "corr"), suffix="x", verbose=True)
        if tag == "our":
            kw["savedir"] = str(d / "ev")     # the reference's savedir loop dies on the skipped chromosome
        results.append(_call(hic, **kw))
    (res_r, warn_r, out_r), (res_o, warn_o, out_o) = results
    assert ours.bads.keys() == ref.bads.keys()
    assert warn_o == warn_r
    _compare(res_o, res_r, ours, ref, oriented=True)
    # printed progress: same lines; the Spearman line up to the printed digits and the free sign
    lr, lo = out_r.splitlines(), out_o.splitlines()
    assert len(lr) == len(lo)
    for a, b in zip(lo, lr):
        if "rho:" in a:
            assert abs(abs(float(a.split()[1])) - abs(float(b.split()[1]))) < 2e-7
        else:
            assert a == b
    # compartment table: identical text unless a free sign flipped every type of a chromosome
    tr, to = (tmp_path / "ref" / "cmp.tsv").read_text(), (tmp_path / "our" / "cmp.tsv").read_text()
    assert to == tr
    # correlation matrices: same layout, values to 1e-10
    for name in sorted(os.listdir(tmp_path / "ref" / "corr")):
        a = (tmp_path / "our" / "corr" / name).read_text().splitlines()
        b = (tmp_path / "ref" / "corr" / name).read_text().splitlines()
        assert a[0] == b[0] and len(a) == len(b)
        for la, lb in zip(a[1:], b[1:]):
            fa, fb = la.split("\t"), lb.split("\t")
            assert fa[:2] == fb[:2] and len(fa) == len(fb)
            va = np.array([float(x) for x in fa[2:]])
            vb = np.array([float(x) for x in fb[2:]])
            assert np.array_equal(np.isnan(va), np.isnan(vb))
            np.testing.assert_allclose(va[~np.isnan(va)], vb[~np.isnan(vb)], rtol=0, atol=1e-10)
    # eigenvector tables (ours): header with the eigenvalues, one line per bin, NaN at the bad bins
    for sec in res_o[0]:
        idx = [1, 2, 1, 1][list(case["chromosomes"]).index(sec)]
        lines = (tmp_path / "our" / "ev" / ("%s_EigVect%d.tsv" % (sec, idx))).read_text().splitlines()
        assert lines[0].startswith("# EV_1 (%.4f)" % res_o[0][sec][0][0])
        assert len(lines) - 1 == len(res_o[0][sec][1][0])


def test_selected_chromosomes_smoothing_and_all_vectors():
    case = compartment_case(321, sizes=(40, 36, 12), n_dead=2)
    ref, ours = _pair(case)
    res_r, _, _ = _call(ref, crms=["chr2", "chr3"], smoothing_window=3)
    res_o, _, _ = _call(ours, crms=["chr2", "chr3"], smoothing_window=3)
    assert list(res_o[0].keys()) == ["chr2", "chr3"]
    _compare(res_o, res_r, ours, ref)
    # max_ev=None: all but one eigenpair, by magnitude
    ref2, ours2 = _pair(case)
    res_r, _, _ = _call(ref2, crms=["chr3"], max_ev=None)
    res_o, _, _ = _call(ours2, crms=["chr3"], max_ev=None)
    assert len(res_o[0]["chr3"][1]) == len(res_r[0]["chr3"][1]) == 12 - 1 - len(
        [b for b in ours2.bads if b >= 76])
    np.testing.assert_allclose(res_o[0]["chr3"][0], res_r[0]["chr3"][0], rtol=1e-9, atol=1e-12)


def test_bad_bin_at_the_first_bin_of_the_next_chromosome():
    """`beg <= k <= end` (hic_data.py:951): a bad first bin of the next chromosome appends a NaN."""
    case = compartment_case(77, sizes=(30, 30), n_dead=0)
    keep = (case["rows"] != 30) & (case["cols"] != 30) & (case["rows"] != 4) & (case["cols"] != 4)
    for key in ("rows", "cols", "vals"):
        case[key] = case[key][keep]
    ref, ours = _pair(case)
    res_r, _, _ = _call(ref)
    res_o, _, _ = _call(ours)
    assert set(ours.bads) == {4, 30}
    assert len(res_r[0]["chr1"][1][0]) == 31 and len(res_o[0]["chr1"][1][0]) == 31
    _compare(res_o, res_r, ours, ref)


def test_per_chromosome_decay_and_zero_expected():
    """`self.expected[sec][d]` first (the decay of a biases pickle), the flat table on KeyError, an
    empty entry -> skipped chromosome; a zero expected value raises as Python's float division."""
    case = compartment_case(55, sizes=(36, 30, 20), n_dead=2)
    ref, ours = _pair(case)
    for hic in (ref, ours):
        hic.filter_columns(silent=True, by_mean=False)
        hic.normalize_hic(iterations=10, max_dev=1e-4, silent=True)
        hic.normalize_expected()
    flat = dict(ours.expected)
    per = {"chr1": dict((d, v * 1.25) for d, v in flat.items()), "chr3": {}}
    for d, v in flat.items():
        per[d] = v                                        # flat keys next to the chromosome keys
    ref.expected, ours.expected = dict(per), dict(per)
    res_r, warn_r, _ = _call(ref)
    res_o, warn_o, _ = _call(ours)
    assert warn_o == warn_r and any("chr3" in w for w in warn_r)
    _compare(res_o, res_r, ours, ref)
    zeroed = dict(flat)
    zeroed[5] = 0.0
    ref.expected, ours.expected = dict(zeroed), dict(zeroed)
    for hic in (ref, ours):
        with pytest.raises(ZeroDivisionError):
            _call(hic)


def test_too_small_chromosome_is_skipped_where_the_reference_dies():
    case = compartment_case(91, sizes=(30, 2), n_dead=0)
    ref, ours = _pair(case)
    with pytest.raises(IndexError):
        _call(ref)
    res_o, warn_o, _ = _call(ours)
    assert ours.compartments["chr2"] == [] and any("too small" in w for w in warn_o)
    assert "chr1" in res_o[0]


def test_bed_reader_matches_the_reference_parser(tmp_path):
    """Same bins and sums as pytadbit.parsers.bed_parser.parse_bed.  That parser loses the first
    data line of a file without header (it seeks to the end of the line it sniffed,
    bed_parser.py:58-93) and mis-seeks after a header; ours reads every line, so the reference is
    given a sacrificial copy of the first line."""
    from tadbit_b200.compartments import read_bed_bins
    ref_shim.load()
    from pytadbit.parsers.bed_parser import parse_bed
    texts = {
        "bed6.bed": "chr1\t100\t250\tg1\t3.5\t+\nchr1\t900\t1200\tg2\t1\t-\nchr2\t10\t20\tg3\t2\t+\n",
        "bedname.bed": "chr1\t100\t250\tg1\tname\t+\nchr1\t130\t260\tg2\tname\t-\n",
        "graph.bedgraph": "chr1\t0\t1000\t0.5\nchr1\t1000\t2000\t1.5\nchr2\t0\t1000\t4\n",
        "three.txt": "chr1 100 250\nchr1 300 1250\n",
        "two.txt": "chr1 100\nchr2 5000\n",
    }
    for name, text in texts.items():
        p, q = tmp_path / name, tmp_path / ("ref_" + name)
        p.write_text(text)
        q.write_text(text.splitlines(True)[0] + text)
        for reso in (1, 100, 1000):
            assert read_bed_bins(str(p), reso) == parse_bed(str(q), reso), (name, reso)
    h = tmp_path / "hdr.bed"
    h.write_text("track name=x\n#c\nbrowser position\n" + texts["graph.bedgraph"])
    assert read_bed_bins(str(h), 1000) == read_bed_bins(str(tmp_path / "graph.bedgraph"), 1000)
