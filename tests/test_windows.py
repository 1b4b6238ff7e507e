"""N3 (HiC_data part): get_matrix / yield_matrix / write_matrix against the unmodified reference
(tests/golden/window_3chrom.*, oracle/gen_golden.py:reference_window_cases) -- on the CPU through
the store double (host logic: focus parsing, diagonal / bad-row rules, file text) and on the GPU
through tb_window."""
import json
import os
from collections import OrderedDict

import numpy as np
import pytest

from _golden import GOLDEN
from _numpy_store import NumpyStore


def _load():
    rec = json.load(open(os.path.join(GOLDEN, "window_3chrom.json")))
    z = np.load(os.path.join(GOLDEN, "window_3chrom.npz"))
    chroms = OrderedDict((k, int(v)) for k, v in rec["chromosomes"])
    return rec, z, chroms


def _sections(chroms):
    sec, idx = {}, 0
    for name in chroms:
        for p in range(chroms[name]):
            sec[(name, p)] = idx
            idx += 1
    return sec


def _focus(f):
    return tuple(f) if isinstance(f, list) else f


def _check(hic, rec, tmp_path):
    from tadbit_b200.vectors import BinVector
    hic.bads = dict((int(k), v) for k, v in rec["bads"].items())
    hic.bias = BinVector(np.array(rec["bias"]))
    for call in rec["calls"]:
        kw = dict(focus=_focus(call["focus"]), diagonal=call["diagonal"], normalized=call["normalized"])
        if call["fn"] == "get_matrix":
            got = hic.get_matrix(**kw)
        else:
            got = list(hic.yield_matrix(**kw))
        want = call["value"]
        assert len(got) == len(want), (call["fn"], kw)
        if not want:
            continue
        g, w = np.array(got, dtype=np.float64), np.array(want, dtype=np.float64)
        assert g.shape == w.shape, (call["fn"], kw)
        np.testing.assert_allclose(g, w, rtol=1e-13, atol=0, err_msg=str((call["fn"], kw)))
        if not call["normalized"]:
            assert all(isinstance(x, int) for x in got[0])
    mm = hic.get_matrix(focus=_focus(rec["masked"]["focus"]), normalized=True, masked=True)
    assert np.asarray(mm.mask).astype(int).tolist() == rec["masked"]["mask"]
    for tag, kw in (("raw", dict()), ("norm_focus", dict(focus="chrB", normalized=True, diagonal=False))):
        fname = str(tmp_path / ("m_%s.tsv" % tag))
        hic.write_matrix(fname, **kw)
        got, want = open(fname).read().splitlines(), rec["write_" + tag].splitlines()
        assert len(got) == len(want)
        for a, b in zip(got, want):
            fa, fb = a.split("\t"), b.split("\t")
            assert len(fa) == len(fb)
            for x, y in zip(fa, fb):
                if x != y:
                    assert abs(float(x) - float(y)) <= 1e-13 * abs(float(y)), (a[:60], b[:60])


def test_windows_host_layer_against_the_reference(tmp_path):
    import tadbit_b200
    rec, z, chroms = _load()
    store = NumpyStore(rec["size"], z["rows"], z["cols"], z["vals"], chromosomes=chroms)
    hic = tadbit_b200.HiC_data.from_store(store, chromosomes=chroms, dict_sec=_sections(chroms),
                                          resolution=rec["resolution"])
    _check(hic, rec, tmp_path)


@pytest.mark.gpu
def test_windows_through_the_kernel(tmp_path):
    import tadbit_b200
    rec, z, chroms = _load()
    hic = tadbit_b200.HiC_data.from_coo(rec["size"], z["rows"], z["cols"], z["vals"], chromosomes=chroms,
                                        dict_sec=_sections(chroms), resolution=rec["resolution"])
    _check(hic, rec, tmp_path)
    # the same on a store that kept its CSR, and a window of a 20k-bin matrix against the exported cells
    os.environ["TB_KEEP_CSR"] = "1"
    try:
        hic2 = tadbit_b200.HiC_data.from_coo(rec["size"], z["rows"], z["cols"], z["vals"], chromosomes=chroms,
                                             dict_sec=_sections(chroms), resolution=rec["resolution"])
    finally:
        del os.environ["TB_KEEP_CSR"]
    assert hic2.store.timing()["csr_resident"]
    _check(hic2, rec, tmp_path)
    big = OrderedDict([("c1", 9000), ("c2", 6500), ("c3", 4500)])
    n = sum(big.values())
    p = tadbit_b200.SynthParams(seed=31, cis_scale=900.0, cis_alpha=1.08, trans_mean=0.004, vis_sigma=0.35,
                                low_frac=0.03, low_scale=0.02, empty_frac=0.01)
    store = tadbit_b200.ContactStore.synthetic(n, p, chromosomes=big)
    r, c, v = store.export_upper()
    for (r0, r1, c0, c1) in ((0, 300, 0, 300), (8800, 9300, 100, 700), (15000, 15400, 15100, 15900)):
        want = np.zeros((r1 - r0, c1 - c0))
        for a, b in ((r, c), (c, r)):
            k = (a >= r0) & (a < r1) & (b >= c0) & (b < c1)
            want[a[k] - r0, b[k] - c0] = v[k]
        assert np.array_equal(store.window(r0, r1, c0, c1), want)


def test_coord_tables_against_the_live_reference(tmp_path):
    """write_coord_table (hic_data.py:451-540) in both formats, raw and normalised, next to the
    unmodified reference on the same matrix (skipped where the reference tree is absent)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    import tadbit_b200
    from tadbit_b200.vectors import BinVector
    ns = ref_shim.load()
    rec, z, chroms = _load()
    n = rec["size"]
    dico = {}
    for a, b, v in zip(z["rows"].tolist(), z["cols"].tolist(), z["vals"].tolist()):
        dico[a * n + b] = v
        dico[b * n + a] = v
    ref = ns.HiC_data(dico, n, chromosomes=chroms, dict_sec=_sections(chroms), resolution=rec["resolution"])
    store = NumpyStore(n, z["rows"], z["cols"], z["vals"], chromosomes=chroms)
    ours = tadbit_b200.HiC_data.from_store(store, chromosomes=chroms, dict_sec=_sections(chroms),
                                           resolution=rec["resolution"])
    bads = dict((int(k), v) for k, v in rec["bads"].items())
    ref.bads, ours.bads = dict(bads), dict(bads)
    ref.bias = dict(enumerate(rec["bias"]))
    ours.bias = BinVector(np.array(rec["bias"]))
    for fmt in ("BED", "long-range"):
        for kw in (dict(), dict(normalized=True), dict(focus="chrB", diagonal=False),
                   dict(focus=(3, 40), normalized=True)):
            fr, fo = str(tmp_path / "r.txt"), str(tmp_path / "o.txt")
            ref.write_coord_table(fr, format=fmt, **kw)
            ours.write_coord_table(fo, format=fmt, **kw)
            assert open(fo).read() == open(fr).read(), (fmt, kw)
    for hic in (ref, ours):
        with pytest.raises(Exception, match="not found"):
            hic.write_coord_table(str(tmp_path / "x.txt"), format="wig")


def test_focus_forms_against_the_live_reference():
    """Every focus form of get_matrix (hic_data.py:688-738), ranges in base pairs included, and the
    reference's quirk for a pair of ranged names, next to the unmodified reference."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    import tadbit_b200
    ns = ref_shim.load()
    rec, z, chroms = _load()
    n, reso = rec["size"], rec["resolution"]
    dico = {}
    for a, b, v in zip(z["rows"].tolist(), z["cols"].tolist(), z["vals"].tolist()):
        dico[a * n + b] = v
        dico[b * n + a] = v
    ref = ns.HiC_data(dico, n, chromosomes=chroms, dict_sec=_sections(chroms), resolution=reso)
    store = NumpyStore(n, z["rows"], z["cols"], z["vals"], chromosomes=chroms)
    ours = tadbit_b200.HiC_data.from_store(store, chromosomes=chroms, dict_sec=_sections(chroms),
                                           resolution=reso)
    names = list(chroms)
    a, b = names[0], names[1]
    rng = "%d-%d" % (2 * reso, 9 * reso)
    forms = [None, (3, 17), (2, 9, 20, 31), a, (a, b), "%s:%s" % (a, rng), ("%s:%s" % (a, rng), b),
             ("%s:%s" % (a, rng), "%s:%s" % (b, rng))]
    for focus in forms:
        assert ours._focus_coords(focus) == ref._focus_coords(focus), focus
        assert ours.get_matrix(focus=focus) == ref.get_matrix(focus=focus), focus
    assert ours.get_as_tuple() == ref.get_as_tuple()
    for hic in (ref, ours):
        with pytest.raises(Exception, match="should be in format"):
            hic.get_matrix(focus="%s:12" % a)
        with pytest.raises(IndexError):
            hic.get_matrix(focus=(a, "%s:%s" % (b, rng)))      # range of the 2nd name read from the 1st


def test_cooler_weights_and_the_missing_h5py_message():
    """write_cooler's weight column (_pytadbit/hic_data.py:572-574: 1 / bias on good bins with a
    positive bias, else 0) and the reference's error when h5py is absent (:549-551)."""
    import tadbit_b200
    from tadbit_b200.vectors import BinVector
    rec, z, chroms = _load()
    store = NumpyStore(rec["size"], z["rows"], z["cols"], z["vals"], chromosomes=chroms)
    hic = tadbit_b200.HiC_data.from_store(store, chromosomes=chroms, resolution=rec["resolution"])
    hic.bads = dict((int(k), v) for k, v in rec["bads"].items())
    bias = np.array(rec["bias"])
    bias[3], bias[4] = 0.0, -1.0
    hic.bias = BinVector(bias)
    want = [1. / hic.bias[i] if i not in hic.bads and hic.bias[i] > 0 else 0. for i in range(len(hic))]
    assert hic.cooler_weights().tolist() == want and want[3] == 0. and want[4] == 0.
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(Exception, match="cooler output is not available"):
            hic.write_cooler("x.mcool", normalized=True)
