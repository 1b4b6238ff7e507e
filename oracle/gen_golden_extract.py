#!/usr/bin/env python
"""Generate tests/golden/extract_3chrom.* by running the UNMODIFIED reference's BAM extraction
(_pytadbit/parsers/hic_bam_parser.py: read_bam, get_biases_region, get_matrix, write_matrix) --
test infrastructure, build container only (needs /root/reference):

    python oracle/gen_golden_extract.py

pysam is not installed in this image, and the reference reads its input through
``pysam.AlignmentFile``.  The generator puts a stand-in for THAT CLASS ONLY into ``sys.modules``:
an object with ``references``, ``lengths`` and ``fetch(region, start, end)`` that serves synthetic
one-base read pairs, every pair twice (once from each end), which is how TADbit's BAM files hold
them.  Everything downstream -- chunking, binning, the per-chunk files, bias / bad-column
re-indexing, the value transforms and the text of the output files -- is the reference's own code,
untouched.  ``pytadbit.mapping.filter`` (imports scipy functions this image's scipy no longer has)
is replaced by a module holding only its ``MASKED`` table, read out of the reference's source with
``ast`` at run time.

Stored: the matrix the reads add up to (upper triangle, .npz), the biases pickle content, and for
each call its arguments and what the reference returned / wrote / raised.
"""
import ast
import contextlib
import importlib
import io
import json
import os
import pickle
import sys
import tempfile
from collections import OrderedDict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from gen_golden import synth_dense  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class _Read(object):
    __slots__ = ("flag", "reference_name", "reference_start", "mrnm", "mpos")

    def __init__(self, flag, name, start, mrnm, mpos):
        self.flag, self.reference_name, self.reference_start, self.mrnm, self.mpos = flag, name, start, mrnm, mpos


def load_bam_parser(bams):
    """The reference's hic_bam_parser module with ``AlignmentFile`` served from ``bams``
    ({path: dict(references, lengths, reads=[(flag, chrom index, 0-based pos, mate chrom, mate pos)])})."""
    ref_shim.load()
    src = open(os.path.join(ref_shim.REFERENCE_ROOT, "_pytadbit", "mapping", "filter.py")).read()
    masked = None
    for node in ast.parse(src).body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "MASKED":
            masked = ast.literal_eval(node.value)
    ref_shim._stub("pytadbit.mapping.filter", MASKED=masked)

    class AlignmentFile(object):
        def __init__(self, path, mode="rb"):
            spec = bams[path]
            self.references = tuple(spec["references"])
            self.lengths = tuple(spec["lengths"])
            self._reads = spec["reads"]

        def fetch(self, region=None, start=None, end=None, multiple_iterators=False):
            for flag, c1, p1, c2, p2 in self._reads:
                if self.references[c1] == region and start <= p1 < end:
                    yield _Read(flag, self.references[c1], p1, c2, p2)

    sys.modules["pysam"].AlignmentFile = AlignmentFile
    sys.modules.pop("pytadbit.parsers.hic_bam_parser", None)
    return importlib.import_module("pytadbit.parsers.hic_bam_parser")


def reads_of(rng, sizes, lengths, reso, r, c, v):
    """Read pairs that bin to the matrix: cell (i, j) = v pairs, each stored from both ends (the
    diagonal holds an even count, as a TADbit BAM gives)."""
    offs = np.concatenate([[0], np.cumsum(sizes)])
    chrom = np.repeat(np.arange(len(sizes)), sizes)
    reads = []
    for i, j, n in zip(r.tolist(), c.tolist(), v.tolist()):
        ci, cj = int(chrom[i]), int(chrom[j])
        for _ in range(n if i != j else n // 2):
            # 1-based position pos with pos // reso = bin; never past the end of the chromosome
            hi_i = min(reso - 1, lengths[ci] - (i - offs[ci]) * reso)
            hi_j = min(reso - 1, lengths[cj] - (j - offs[cj]) * reso)
            p1 = (i - offs[ci]) * reso + int(rng.integers(1, max(hi_i, 1) + 1))
            p2 = (j - offs[cj]) * reso + int(rng.integers(1, max(hi_j, 1) + 1))
            reads.append((0, ci, int(p1) - 1, cj, int(p2) - 1))
            reads.append((0, cj, int(p2) - 1, ci, int(p1) - 1))
    # a few pairs that a filter flag removes (filter_exclude's default mask covers flag bit 1)
    for _ in range(50):
        reads.append((1, 0, int(rng.integers(0, lengths[0])), 0, int(rng.integers(0, lengths[0]))))
    order = rng.permutation(len(reads))
    return [reads[k] for k in order]


def _jsonable(x):
    if isinstance(x, dict):
        return [[_jsonable(k), _jsonable(v)] for k, v in x.items()]
    if isinstance(x, (tuple, list)):
        return [_jsonable(k) for k in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    return x


def main():
    rng = np.random.default_rng(777)
    reso = 10000
    names = ["chrA", "chrB", "chrC"]
    sizes = [40, 25, 15]
    lengths = [395000, 240000 + 4321, 140000 + 9999]
    assert [l // reso + 1 for l in lengths] == sizes
    n, r, c, v = synth_dense(rng, sizes, a=60.0, trans=0.25, low_frac=0.1, empty_frac=0.05)
    v = np.where(r == c, 2 * ((v + 1) // 2), v)          # even diagonal
    bams = {"fake.bam": dict(references=names, lengths=lengths, reads=reads_of(rng, sizes, lengths, reso, r, c, v))}
    # a second genome of two chromosomes: the genome-wide decay writer has no KeyError guard there
    n2, r2, c2, v2 = synth_dense(rng, sizes[:2], a=40.0, trans=0.3, low_frac=0.0, empty_frac=0.0)
    v2 = np.where(r2 == c2, 2 * ((v2 + 1) // 2), v2)
    bams["two.bam"] = dict(references=names[:2], lengths=lengths[:2],
                           reads=reads_of(rng, sizes[:2], lengths[:2], reso, r2, c2, v2))
    mod = load_bam_parser(bams)

    bias = dict((i, float(x)) for i, x in enumerate(rng.lognormal(0.0, 0.3, n)))
    badcol = OrderedDict((int(b), True) for b in sorted(rng.choice(n, 7, replace=False).tolist()))
    decay = OrderedDict()
    for name, size in zip(names, sizes):
        decay[name] = dict((d, float(50.0 * (d + 1.0) ** -1.1 * rng.uniform(0.8, 1.2))) for d in range(size))
    biases = dict(biases=bias, badcol=badcol, decay=decay, resolution=reso)
    # chrC lacks the longest distances (a decay table from a shorter run)
    partial = dict(biases=bias, badcol=badcol, resolution=reso,
                   decay=OrderedDict((k, dict((d, x) for d, x in t.items() if not (k == "chrC" and d > 9)))
                                     for k, t in decay.items()))
    rec = OrderedDict(name="extract_3chrom", resolution=reso, chromosomes=list(zip(names, sizes)),
                      lengths=lengths, biases=_jsonable(bias), badcol=_jsonable(badcol),
                      decay=[[k, _jsonable(t)] for k, t in decay.items()], get=[], write=[])

    tmp = tempfile.mkdtemp()
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())          # noqa: E731

    def bias_arg(which):
        return {None: None, "full": biases, "partial": partial}[which]

    region_sets = [
        dict(),
        dict(region1="chrB"),
        dict(region1="chrA", start1=50000, end1=250000),
        dict(region1="chrA", start1=0, end1=120000, region2="chrC"),
        dict(region1="chrB", region2="chrA", start2=100000, end2=300000),
        dict(region1="chrC", start1=20000, end1=90000, region2="chrC", start2=60000, end2=140000),
        dict(region1="chrA", end1=100000),
        dict(region1="chrC", start1=100000),
    ]
    for bam, which, normalization, regions in (
            [("fake.bam", None, "raw", reg) for reg in region_sets] +
            [("fake.bam", "full", "norm", reg) for reg in region_sets] +
            [("fake.bam", "full", "decay", reg) for reg in region_sets] +
            [("fake.bam", "partial", "decay", dict(region1="chrC")),
             ("fake.bam", "full", "raw", dict(region1="chrA")),
             ("fake.bam", None, "norm", dict(region1="chrA")),
             ("fake.bam", "full", "zscore", dict(region1="chrA")),
             ("fake.bam", None, "raw", dict(start1=10000)),
             ("fake.bam", None, "raw", dict(region1="chrA", region2="chrZ")),
             ("fake.bam", None, "raw", dict(region1="chrZ")),
             ("fake.bam", None, "raw", dict(region1="chrA", max_size=100)),
             ("two.bam", None, "raw", dict())]):
        item = OrderedDict(bam=bam, biases=which, normalization=normalization, regions=regions)
        try:
            with quiet():
                out = mod.get_matrix(bam, reso, biases=bias_arg(which), normalization=normalization, tmpdir=tmp,
                                     ncpus=1, clean=True, verbose=False, return_headers=True, **regions)
            dico, bads1, bads2, regs, name, coords = out
            item["cells"] = sorted([int(i), int(j), x] for (i, j), x in dico.items())
            item["bads1"], item["bads2"] = _jsonable(bads1), _jsonable(bads2)
            item["regions_out"], item["name_out"], item["bin_coords"] = list(regs), name, [int(x) for x in coords]
        except BaseException as e:                                      # noqa: B036
            item["raises"] = [type(e).__name__, str(e)]
        rec["get"].append(item)

    write_sets = [
        dict(normalizations=("raw", "norm", "decay")),
        dict(normalizations=("raw", "norm", "decay"), half_matrix=False, extra="all"),
        dict(normalizations=("decay",), region1="chrB"),
        dict(normalizations="norm", region1="chrA", start1=50000, end1=250000),
        dict(normalizations=["raw", "decay"], region1="chrA", start1=0, end1=120000, region2="chrC"),
        dict(normalizations=("raw", "norm", "decay"), region1="chrB", region2="chrA", start2=100000, end2=300000,
             row_names=True),
        dict(normalizations=("raw", "norm"), region1="chrB", region2="chrA", start2=100000, end2=300000,
             row_names=True),
        dict(normalizations=("raw", "decay"), region1="chrC", start1=20000, end1=90000, region2="chrC",
             start2=60000, end2=140000),
        dict(normalizations=("norm",), region1="chrA", start1=300000, end1=600000),
        dict(normalizations=("raw&decay",), region1="chrC"),
        dict(normalizations=("raw&decay", "decay"), region1="chrA", half_matrix=False),
        dict(normalizations=("raw",), region1="chrA", row_names=True),
        dict(normalizations=("raw",), row_names=True),
        dict(normalizations=("raw",), region1="chrA", start1=50000, end1=55000),
        dict(normalizations=("norm",), cooler=True),
        dict(normalizations=("decay",), biases="partial", region1="chrC"),
        dict(normalizations=("decay",), biases="partial"),
        dict(normalizations=("raw",), biases=None),
        dict(normalizations=("norm",), biases=None),
        dict(normalizations=("decay",), bam="two.bam"),
        dict(normalizations=("raw", "norm"), bam="two.bam"),
    ]
    for kw in write_sets:
        kw = dict(kw)
        bam = kw.pop("bam", "fake.bam")
        which = kw.pop("biases", "full")
        item = OrderedDict(bam=bam, biases=which, kwargs=_jsonable_kwargs(kw))
        outdir = tempfile.mkdtemp()
        try:
            with quiet():
                fnames = mod.write_matrix(bam, reso, bias_arg(which), outdir, tmpdir=tmp, ncpus=1, verbose=False, **kw)
            item["fnames"] = dict((k, os.path.basename(p)) for k, p in fnames.items())
        except BaseException as e:                                      # noqa: B036
            item["raises"] = [type(e).__name__, str(e)]
        item["files"] = dict((f, open(os.path.join(outdir, f)).read()) for f in sorted(os.listdir(outdir)))
        rec["write"].append(item)

    np.savez_compressed(os.path.join(OUT, "extract_3chrom.npz"), size=n, rows=r.astype(np.int32),
                        cols=c.astype(np.int32), vals=v.astype(np.int64), size2=n2, rows2=r2.astype(np.int32),
                        cols2=c2.astype(np.int32), vals2=v2.astype(np.int64))
    with open(os.path.join(OUT, "extract_3chrom.json"), "w") as fh:
        json.dump(rec, fh)
    print("extract_3chrom: %d get_matrix calls, %d write_matrix calls" % (len(rec["get"]), len(rec["write"])))
    for item in rec["get"] + rec["write"]:
        if "raises" in item:
            print("  raises", item.get("regions", item.get("kwargs")), item["raises"])


def _jsonable_kwargs(kw):
    return dict((k, list(x) if isinstance(x, tuple) else x) for k, x in kw.items())


if __name__ == "__main__":
    main()
