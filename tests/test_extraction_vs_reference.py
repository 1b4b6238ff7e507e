"""extraction.get_matrix / write_matrix against the UNMODIFIED reference, live, on seeded random
genomes, matrices, region arguments, biases and normalisations (runs where the reference tree
exists).  The reference's parsers/hic_bam_parser.py reads synthetic read pairs through the stand-in
for pysam.AlignmentFile of oracle/gen_golden_extract.py; the product's host layer runs over the
store double.  Everything observable is compared: the returned dictionary and headers, the files'
names and text (bodies as sorted lines: a BAM file fixes an order, a matrix does not), the
exceptions."""
import contextlib
import io
import os
import sys
from collections import OrderedDict

import numpy as np
import pytest

from _numpy_store import NumpyStore

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")

NCASES = 60


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    reso = int(rng.choice([1000, 5000, 20000]))
    nchrom = int(rng.integers(1, 5))
    sizes = [int(x) for x in rng.integers(4, 26, nchrom)]
    lengths = [int((s - 1) * reso + rng.integers(0, reso)) for s in sizes]
    names = ["chr%s" % "ABCD"[k] for k in range(nchrom)]
    n = sum(sizes)
    chrom = np.repeat(np.arange(nchrom), sizes)
    i, j = np.triu_indices(n)
    lam = np.where(chrom[i] == chrom[j], 6.0 * (np.abs(i - j) + 1.0) ** -1.0, 0.15)
    v = rng.poisson(lam)
    v = np.where(i == j, 2 * ((v + 1) // 2), v)
    keep = v > 0
    r, c, v = i[keep], j[keep], v[keep].astype(np.int64)
    bias = dict((k, float(x)) for k, x in enumerate(rng.lognormal(0.0, 0.3, n)))
    nbad = int(rng.integers(0, max(n // 6, 1) + 1))
    badcol = OrderedDict((int(b), True) for b in sorted(rng.choice(n, nbad, replace=False).tolist()))
    decay = OrderedDict()
    for name, size in zip(names, sizes):
        top = size if rng.random() < 0.8 else max(size // 2, 1)          # sometimes a short table
        decay[name] = dict((d, float(9.0 * (d + 1.0) ** -1.1)) for d in range(top))
    biases = dict(biases=bias, badcol=badcol, decay=decay, resolution=reso)

    def coord(size, rng):
        return int(rng.integers(0, size + 3)) * reso + int(rng.choice([0, 0, reso // 3]))

    kw = {}
    if rng.random() < 0.8:
        k1 = int(rng.integers(0, nchrom))
        kw["region1"] = names[k1]
        if rng.random() < 0.5:
            a, b = sorted([coord(sizes[k1], rng), coord(sizes[k1], rng)])
            if rng.random() < 0.8:
                kw["start1"] = a
            if rng.random() < 0.8:
                kw["end1"] = max(b, a + reso)
        if rng.random() < 0.45:
            k2 = int(rng.integers(0, nchrom))
            kw["region2"] = names[k2]
            if rng.random() < 0.5:
                a, b = sorted([coord(sizes[k2], rng), coord(sizes[k2], rng)])
                if rng.random() < 0.8:
                    kw["start2"] = a
                if rng.random() < 0.8:
                    kw["end2"] = max(b, a + reso)
    elif rng.random() < 0.15:
        kw["start1"] = reso
    return dict(reso=reso, names=names, sizes=sizes, lengths=lengths, n=n, rows=r, cols=c, vals=v,
                biases=biases, regions=kw, rng=rng)


def _reads(case):
    from gen_golden_extract import reads_of
    return reads_of(case["rng"], case["sizes"], case["lengths"], case["reso"], case["rows"], case["cols"],
                    case["vals"])


@pytest.fixture(scope="module")
def parser():
    from gen_golden_extract import load_bam_parser
    bams = {}
    return load_bam_parser(bams), bams


def _call(fn, *a, **k):
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            return fn(*a, **k), None
    except BaseException as e:                                         # noqa: B036
        return None, e


def _spills(case):
    """Rows that run past the end of region1's chromosome while a second region is given: the
    reference then also accepts reads of region2's chromosome as ROWS, numbered after the row range
    (its all_bins loop runs over both regions with region1's offsets, hic_bam_parser.py:604-616) --
    an accident the product does not reproduce (it keeps the rows inside region1's chromosome)."""
    kw = case["regions"]
    if "region1" not in kw or "region2" not in kw:
        return False
    size = case["sizes"][case["names"].index(kw["region1"])]
    reso = case["reso"]
    return kw.get("start1", 0) // reso > size or kw.get("end1", 0) // reso > size


def _same_outcome(got, want, what):
    (g, ge), (w, we) = got, want
    if we is not None:
        assert ge is not None, (what, "the reference raises", repr(we))
        assert type(ge).__name__ == type(we).__name__, (what, repr(ge), repr(we))
        if not (isinstance(we, KeyError) and str(we).strip("'").lstrip("-").isdigit()):
            assert str(ge) == str(we), what
        return False
    assert ge is None, (what, repr(ge))
    return True


@pytest.mark.parametrize("seed", range(NCASES))
def test_random_extraction_against_the_live_reference(seed, parser, tmp_path):
    import tadbit_b200
    from tadbit_b200 import extraction
    mod, bams = parser
    case = _case(seed)
    bam = "case%d.bam" % seed
    bams[bam] = dict(references=case["names"], lengths=case["lengths"], reads=_reads(case))
    chroms = OrderedDict(zip(case["names"], case["sizes"]))
    store = NumpyStore(case["n"], case["rows"], case["cols"], case["vals"], chromosomes=chroms)
    hic = tadbit_b200.HiC_data.from_store(store, chromosomes=chroms, resolution=case["reso"])
    reso, regions = case["reso"], case["regions"]
    tmp = str(tmp_path)
    if _spills(case):
        pytest.skip("row range past the end of its chromosome with a second region (documented difference)")
    for normalization in ("raw", "norm", "decay"):
        b = case["biases"] if normalization != "raw" or seed % 2 else None
        want = _call(mod.get_matrix, bam, reso, biases=b, normalization=normalization, tmpdir=tmp, ncpus=1,
                     clean=True, verbose=False, return_headers=True, **regions)
        got = _call(extraction.get_matrix, hic, reso, biases=b, normalization=normalization,
                    return_headers=True, **regions)
        if not _same_outcome(got, want, (seed, normalization, regions)):
            continue
        (gd, gb1, gb2, greg, gname, gcoord), (wd, wb1, wb2, wreg, wname, wcoord) = got[0], want[0]
        assert (list(greg), gname, tuple(gcoord)) == (list(wreg), wname, tuple(wcoord))
        assert list(gb1.items()) == list(wb1.items()) and list(gb2.items()) == list(wb2.items())
        assert sorted(gd) == sorted(wd), (seed, normalization, regions)
        for key, x in wd.items():
            assert gd[key] == x or abs(gd[key] - x) <= 1e-15 * abs(x), (seed, key)
    lengths = dict(zip(case["names"], case["lengths"]))
    for k, kw in enumerate([dict(normalizations=("raw", "norm", "decay")),
                            dict(normalizations=("raw", "norm"), half_matrix=False, row_names=bool(seed % 3 == 0)),
                            dict(normalizations=("raw&decay",), extra="x")]):
        out_ref, out_new = tmp_path / ("ref%d" % k), tmp_path / ("new%d" % k)
        out_ref.mkdir(), out_new.mkdir()
        want = _call(mod.write_matrix, bam, reso, case["biases"], str(out_ref), tmpdir=tmp, ncpus=1, verbose=False,
                     **dict(kw, **regions))
        got = _call(extraction.write_matrix, hic, reso, case["biases"], str(out_new), verbose=False,
                    lengths=lengths, **dict(kw, **regions))
        if not _same_outcome(got, want, (seed, kw, regions)):
            continue
        assert dict((a, os.path.basename(p)) for a, p in got[0].items()) == \
            dict((a, os.path.basename(p)) for a, p in want[0].items())
        assert sorted(os.listdir(str(out_new))) == sorted(os.listdir(str(out_ref)))
        for fname in os.listdir(str(out_ref)):
            a = open(str(out_new / fname)).read().splitlines()
            b = open(str(out_ref / fname)).read().splitlines()
            nhead = sum(1 for line in b if line.startswith("#"))
            assert a[:nhead] == b[:nhead], (seed, fname)
            assert sorted(a[nhead:]) == sorted(b[nhead:]), (seed, fname)
