"""N3, extraction side: get_matrix / write_matrix of the reference's BAM parser
(_pytadbit/parsers/hic_bam_parser.py:755-872, :892-1242) against goldens of the UNMODIFIED
reference (tests/golden/extract_3chrom.*, made by oracle/gen_golden_extract.py, which serves the
reference synthetic read pairs through a stand-in for pysam.AlignmentFile):

* the oracle's restatement of the cell extraction against the reference's dictionaries,
* the product's host layer (tadbit_b200/extraction.py: region arithmetic, bias / bad-column
  re-indexing, file names, headers, line text, exceptions) on the CPU through the store double,
* the same through tb_region_cells on the GPU.

A BAM file fixes the order of the cells, a matrix does not: file bodies are compared as sorted
lines."""
import json
import os
from collections import OrderedDict

import numpy as np
import pytest

import tadbit_oracle as orc
from _golden import GOLDEN
from _numpy_store import NumpyStore


def _load():
    rec = json.load(open(os.path.join(GOLDEN, "extract_3chrom.json")))
    z = np.load(os.path.join(GOLDEN, "extract_3chrom.npz"))
    return rec, z


def _biases(rec, which):
    if which is None:
        return None
    decay = OrderedDict((k, dict((int(d), x) for d, x in t)) for k, t in rec["decay"])
    if which == "partial":
        decay["chrC"] = dict((d, x) for d, x in decay["chrC"].items() if d <= 9)
    return dict(biases=dict((int(k), v) for k, v in rec["biases"]),
                badcol=OrderedDict((int(k), v) for k, v in rec["badcol"]),
                decay=decay, resolution=rec["resolution"])


def _chroms(rec, bam):
    pairs = rec["chromosomes"] if bam == "fake.bam" else rec["chromosomes"][:2]
    return OrderedDict((k, int(v)) for k, v in pairs)


def _matrix(z, bam):
    if bam == "fake.bam":
        return int(z["size"]), z["rows"], z["cols"], z["vals"]
    return int(z["size2"]), z["rows2"], z["cols2"], z["vals2"]


def _same_error(got, want):
    assert type(got).__name__ == want[0], (got, want)
    if want[0] != "KeyError" or not want[1].isdigit():      # which cell comes first depends on the read order
        assert str(got) == want[1]


def test_oracle_region_cells_against_the_reference():
    rec, z = _load()
    done = 0
    for item in rec["get"]:
        if "raises" in item:
            continue
        size, r, c, v = _matrix(z, item["bam"])
        m = orc.OracleMatrix(size, r, c, v, chromosomes=_chroms(rec, item["bam"]))
        b = _biases(rec, item["biases"])
        r0, r1, c0, c1 = item["bin_coords"]
        bias = decay = bads = None
        if b:
            bias = np.array([b["biases"][i] for i in range(size)])
            decay = b["decay"]
            bads = b["badcol"]
        # rows past the end of the chromosome of the region hold nothing
        pos = m.section_pos
        reg = item["regions_out"]
        if len(reg) in (1, 2) and reg != list(pos):
            r1 = min(r1, pos[reg[0]][1])
            c1 = min(c1, pos[reg[-1]][1])
        rr, cc, raw, nrm, dec = orc.region_cells(m, r0, r1, c0, c1, bads, bias, decay)
        val = {"raw": raw, "norm": nrm, "decay": dec}[item["normalization"]]
        got = sorted([int(i) - r0, int(j) - c0, x] for i, j, x in zip(rr, cc, val.tolist()))
        want = item["cells"]
        assert [g[:2] for g in got] == [w[:2] for w in want]
        np.testing.assert_allclose([g[2] for g in got], [w[2] for w in want], rtol=1e-15, atol=0)
        done += 1
    assert done >= 20


def _hic(rec, z, bam, make_store):
    import tadbit_b200
    size, r, c, v = _matrix(z, bam)
    chroms = _chroms(rec, bam)
    return tadbit_b200.HiC_data.from_store(make_store(size, r, c, v, chroms), chromosomes=chroms,
                                           resolution=rec["resolution"])


def _check_get(rec, z, make_store):
    from tadbit_b200.extraction import get_matrix
    hics = {}
    for item in rec["get"]:
        bam = item["bam"]
        if bam not in hics:
            hics[bam] = _hic(rec, z, bam, make_store)
        kw = dict(biases=_biases(rec, item["biases"]), normalization=item["normalization"],
                  return_headers=True, **item["regions"])
        if "raises" in item:
            with pytest.raises(BaseException) as err:
                get_matrix(hics[bam], rec["resolution"], **kw)
            _same_error(err.value, item["raises"])
            continue
        dico, bads1, bads2, regions, name, coords = get_matrix(hics[bam], rec["resolution"], **kw)
        assert list(regions) == item["regions_out"]
        assert name == item["name_out"] and list(coords) == item["bin_coords"]
        assert [[k, v] for k, v in bads1.items()] == item["bads1"]
        assert [[k, v] for k, v in bads2.items()] == item["bads2"]
        got = sorted([i, j, x] for (i, j), x in dico.items())
        want = item["cells"]
        assert [g[:2] for g in got] == [w[:2] for w in want], item["regions"]
        np.testing.assert_allclose([g[2] for g in got], [w[2] for w in want], rtol=1e-15, atol=0)
        if item["normalization"] == "raw":
            assert all(isinstance(g[2], int) for g in got)
    arr = get_matrix(hics["fake.bam"], rec["resolution"], biases=_biases(rec, "full"), normalization="norm",
                     region1="chrB", as_arrays=True)
    dic = get_matrix(hics["fake.bam"], rec["resolution"], biases=_biases(rec, "full"), normalization="norm",
                     region1="chrB")
    assert dict(zip(zip(arr[0].tolist(), arr[1].tolist()), arr[2].tolist())) == dic
    # a dict-like to fill instead of a dictionary to return
    filled = {}
    assert get_matrix(hics["fake.bam"], rec["resolution"], region1="chrB", dico=filled) is None
    want = [it for it in rec["get"] if it["regions"] == {"region1": "chrB"} and it["normalization"] == "raw"][0]
    assert sorted([i, j, x] for (i, j), x in filled.items()) == want["cells"]


def _check_write(rec, z, make_store, tmp_path):
    from tadbit_b200.extraction import write_matrix
    lengths = dict(zip([k for k, _ in rec["chromosomes"]], rec["lengths"]))
    hics = {}
    for n, item in enumerate(rec["write"]):
        bam = item["bam"]
        if bam not in hics:
            hics[bam] = _hic(rec, z, bam, make_store)
        outdir = tmp_path / ("w%d" % n)
        outdir.mkdir()
        kw = dict(item["kwargs"])
        if isinstance(kw.get("normalizations"), list) and n % 2:
            kw["normalizations"] = tuple(kw["normalizations"])
        args = (hics[bam], rec["resolution"], _biases(rec, item["biases"]), str(outdir))
        if "raises" in item:
            with pytest.raises(BaseException) as err:
                write_matrix(*args, lengths=lengths, verbose=False, **kw)
            _same_error(err.value, item["raises"])
            continue
        fnames = write_matrix(*args, lengths=lengths, verbose=False, **kw)
        assert dict((k, os.path.basename(p)) for k, p in fnames.items()) == item["fnames"]
        assert sorted(os.listdir(str(outdir))) == sorted(item["files"])
        for fname, text in item["files"].items():
            got = open(str(outdir / fname)).read().splitlines()
            want = text.splitlines()
            nhead = sum(1 for line in want if line.startswith("#"))
            assert got[:nhead] == want[:nhead], fname
            assert len(got) == len(want), fname
            body_got, body_want = sorted(got[nhead:]), sorted(want[nhead:])
            if body_got != body_want:                        # last-bit differences of a division
                for a, b in zip(body_got, body_want):
                    fa, fb = a.split("\t"), b.split("\t")
                    assert fa[:-1] == fb[:-1], (fname, a, b)
                    if fa[-1] != fb[-1]:
                        assert abs(float(fa[-1]) - float(fb[-1])) <= 1e-15 * abs(float(fb[-1])), (fname, a, b)
            if "raw&decay" in item["kwargs"]["normalizations"] and "decay" in item["kwargs"]["normalizations"]:
                # the two writers share the file: per cell the decay line, then the raw one
                cells = got[nhead:]
                assert all(cells[k].split("\t")[:2] == cells[k + 1].split("\t")[:2] for k in range(0, len(cells), 2))
                assert all("." not in cells[k + 1].split("\t")[-1] for k in range(0, len(cells), 2))


def _numpy_store(size, r, c, v, chroms):
    return NumpyStore(size, r, c, v, chromosomes=chroms)


def test_get_matrix_host_layer_against_the_reference():
    rec, z = _load()
    _check_get(rec, z, _numpy_store)


def test_write_matrix_host_layer_against_the_reference(tmp_path):
    rec, z = _load()
    _check_write(rec, z, _numpy_store, tmp_path)


def test_biases_pickle_path_and_default_lengths(tmp_path):
    """The biases as the path of a pickle (what the reference is normally given), the header built
    from the default chromosome lengths, and a region that lies below the diagonal."""
    import pickle
    from tadbit_b200.extraction import get_biases_region, get_matrix, nicer, write_matrix
    rec, z = _load()
    hic = _hic(rec, z, "fake.bam", _numpy_store)
    path = str(tmp_path / "biases.pickle")
    with open(path, "wb") as out:
        pickle.dump(_biases(rec, "full"), out)
    a = get_matrix(hic, rec["resolution"], biases=path, normalization="norm", region1="chrC", region2="chrA")
    b = get_matrix(hic, rec["resolution"], biases=_biases(rec, "full"), normalization="norm", region1="chrA",
                   region2="chrC")
    assert a and dict(((j, i), x) for (i, j), x in a.items()).keys() == b.keys()
    bias1, bias2, decay, bads1, bads2 = get_biases_region(path, (40, 65, 0, 40), check_resolution=rec["resolution"])
    assert len(bias1) == 25 and len(bias2) == 40 and set(decay) == {"chrA", "chrB", "chrC"}
    with pytest.raises(ValueError):
        get_biases_region(path, (0, 40, 0, 40), check_resolution=5000)
    out = write_matrix(hic, rec["resolution"], path, str(tmp_path), normalizations=("raw",), region1="chrB",
                       verbose=False)
    head = open(out["RAW"]).read().splitlines()[:3]
    assert head == ["# CRM chrB\t240000", "# chrB resolution:10000", "# MASKED "]
    assert [nicer(x) for x in (0, 999, 1000, 25000, 1000000, 2500000, 3000000000)] == \
        ["0 b", "999 b", "1 kb", "25 kb", "1 Mb", "2500 kb", "3 Gb"]
    from tadbit_b200.extraction import cooler_weights_region
    odd = dict(biases={0: 2.0, 1: float("nan"), 2: 0.0, 3: -1.0, 4: 0.5, 5: 4.0}, badcol={}, resolution=1)
    assert cooler_weights_region(odd, (0, 4, 2, 6)) == ([0.5, 0, 0, 0], [0, 0, 2.0, 0.25])   # :1228-1231
    with pytest.raises(ValueError):
        get_matrix(hic, 5000)
    with pytest.raises(NotImplementedError):
        write_matrix(hic, rec["resolution"], path, str(tmp_path), append_to_tar="x.tar")


def _gpu_store(size, r, c, v, chroms):
    import tadbit_b200
    return tadbit_b200.ContactStore.from_coo(size, r, c, v.astype(np.int32), chromosomes=chroms)


@pytest.mark.gpu
def test_get_matrix_through_the_kernel():
    rec, z = _load()
    _check_get(rec, z, _gpu_store)


@pytest.mark.gpu
def test_write_matrix_through_the_kernel(tmp_path):
    rec, z = _load()
    _check_write(rec, z, _gpu_store, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("keep_csr", [False, True])
def test_region_cells_against_the_oracle_on_a_larger_matrix(keep_csr):
    """tb_region_cells over every encoding of the compact stream (dense panels, per-cell chunks,
    sparse tiles) and over a resident CSR, against the oracle: windows on, above and below the
    diagonal, whole genome, half / full, with a mask, biases and a decay table."""
    import tadbit_b200
    chroms = OrderedDict([("c1", 5000), ("c2", 3000), ("c3", 1500)])
    n = sum(chroms.values())
    params = tadbit_b200.SynthParams(seed=11, cis_scale=300.0, cis_alpha=1.05, trans_mean=0.02,
                                     vis_sigma=0.35, low_frac=0.03, low_scale=0.02, empty_frac=0.01)
    store = tadbit_b200.ContactStore.synthetic(n, params, chromosomes=chroms)
    r, c, v = store.export_upper()
    if keep_csr:                                            # a stored zero keeps the CSR resident
        v = v.copy()
        v[::1000] = 0
        store.close()
        store = tadbit_b200.ContactStore.from_coo(n, r, c, v, chromosomes=chroms)
    m = orc.OracleMatrix(n, r, c, v, chromosomes=chroms)
    rng = np.random.default_rng(5)
    bias = rng.lognormal(0.0, 0.3, n)
    bad = (rng.random(n) < 0.05).astype(np.uint8)
    bads = dict((int(i), True) for i in np.flatnonzero(bad))
    decay = np.full(n, np.nan)
    table = {}
    for crm, (beg, end) in m.section_pos.items():
        decay[beg:end] = 80.0 * (np.arange(end - beg) + 1.0) ** -1.05
        table[crm] = dict((d, float(decay[beg + d])) for d in range(end - beg))
    windows = [(0, n, 0, n), (100, 900, 100, 900), (4000, 5600, 200, 1700), (300, 1500, 6000, 9500),
               (7990, 8010, 0, n), (0, 0, 0, n), (9000, 9500, 9000, 9500)]
    for half in (False, True):
        for (r0, r1, c0, c1) in windows[1:] if not half else windows:
            got = store.region_cells(r0, r1, c0, c1, bad=bad, bias=bias, decay=decay, half=half)
            want = orc.region_cells(m, r0, r1, c0, c1, bads, bias, table, half)
            for g, w in zip(got[:3], want[:3]):
                assert np.array_equal(g, w), (r0, r1, c0, c1, half)
            assert np.array_equal(got[3], want[3]), "normalised values differ"
            assert np.array_equal(got[4], want[4], equal_nan=True), "values over the decay differ"
    got = store.region_cells(0, n, 0, n)
    assert got[3] is None and got[4] is None and got[2].sum() == 2 * v.sum() - v[r == c].sum()
    store.close()


def test_bench_extraction_leg_on_the_store_double():
    """bench.py's N3 leg (pipeline.n3_extraction of the JSON line) driven on the CPU."""
    import os as _os
    import sys as _sys
    import tadbit_b200
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    import bench
    rec, z = _load()
    hic = _hic(rec, z, "fake.bam", _numpy_store)
    hic.filter_columns(silent=True)
    hic.normalize_hic(iterations=10, max_dev=1e-4, silent=True)
    hic.normalize_expected()
    wl = dict(chromosomes=_chroms(rec, "fake.bam"))
    out = bench.extraction_leg(hic, hic.store, wl, window=12)
    assert out["section"] == "chrA" and out["window_bins"] == 12 and out["files"] == 3
    assert out["cells_full"] > 0 and out["lines_per_file"] > 3 and out["bytes"] > 0
