#!/usr/bin/env python
"""Cost per cell of the UNMODIFIED reference's write_matrix / get_matrix (test infrastructure,
build container only): how long the part that tadbit_b200/extraction.py replaces takes in the
reference once the BAM file has been read.

    python oracle/time_reference_extract.py [bins]

read_bam (the pysam side, which is not replaced) is swapped for a function that writes the
per-chunk files the reference's readers would have written for a synthetic dense chromosome and
returns what read_bam returns; from there on -- get_biases_region, the header, _iter_matrix_frags
over the chunk files, the writer closures / the dictionary comprehension -- the reference's code
runs untouched and is timed.
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden_extract import load_bam_parser  # noqa: E402


def main():
    nbins = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
    reso = 10000
    rng = np.random.default_rng(4)
    bams = {"fake.bam": dict(references=["chr1"], lengths=[(nbins - 1) * reso + 5], reads=[])}
    mod = load_bam_parser(bams)
    i, j = np.triu_indices(nbins)
    v = rng.poisson(200.0 * (j - i + 1.0) ** -1.08 + 0.3)
    keep = v > 0
    i, j, v = i[keep], j[keep], v[keep]
    tmp = tempfile.mkdtemp()
    nchunk = 100
    edges = np.linspace(0, nbins, nchunk + 1).astype(int)

    def fake_read_bam(inbam, filter_exclude, resolution, half=False, tmpdir=".", **kw):
        rand_hash = "%016x" % int(rng.integers(0, 2 ** 62))
        os.mkdir(os.path.join(tmpdir, "_tmp_%s" % rand_hash))
        regs, begs, ends = [], [], []
        for k in range(nchunk):
            lo, hi = edges[k], edges[k + 1]
            sel = (i >= lo) & (i < hi)
            rows, cols, vals = i[sel], j[sel], v[sel]
            if not half:                                     # both triangles
                lower = (j >= lo) & (j < hi) & (i != j)
                rows = np.concatenate([rows, j[lower]])
                cols = np.concatenate([cols, i[lower]])
                vals = np.concatenate([vals, v[lower]])
            name = os.path.join(tmpdir, "_tmp_%s" % rand_hash, "chr1:%d-%d.tsv" % (lo * reso, hi * reso))
            with open(name, "w") as out:
                out.write("".join("chr1\t%d\t%d\t%d\n" % t for t in zip(rows.tolist(), cols.tolist(), vals.tolist())))
            regs.append("chr1"), begs.append(lo * reso), ends.append(hi * reso)
        return ["chr1"], rand_hash, (0, nbins, 0, nbins), (regs, begs, ends)

    mod.read_bam = fake_read_bam
    biases = dict(biases=dict(enumerate(rng.lognormal(0, 0.3, nbins).tolist())), badcol={5: True, 77: True},
                  decay={"chr1": dict(enumerate((50.0 * (np.arange(nbins) + 1.0) ** -1.1).tolist()))},
                  resolution=reso)
    out = {"bins": nbins, "stored_upper_cells": int(v.size), "runs": []}
    quiet = contextlib.redirect_stdout(io.StringIO())
    for norms in (("raw",), ("raw", "norm", "decay")):
        outdir = tempfile.mkdtemp()
        t0 = time.perf_counter()
        with quiet:
            mod.write_matrix("fake.bam", reso, biases, outdir, normalizations=norms, region1="chr1", tmpdir=tmp,
                             ncpus=1, verbose=False, clean=False)
        t1 = time.perf_counter()
        # the stand-in's own share: writing the chunk files
        t2 = time.perf_counter()
        fake_read_bam("fake.bam", 0, reso, half=True, tmpdir=tmp)
        t3 = time.perf_counter()
        lines = sum(1 for _ in open(os.path.join(outdir, os.listdir(outdir)[0]))) - 3
        out["runs"].append({"what": "write_matrix " + "+".join(norms), "cells": lines,
                            "s": (t1 - t0) - (t3 - t2), "us_per_cell": ((t1 - t0) - (t3 - t2)) / lines * 1e6})
    t0 = time.perf_counter()
    with quiet:
        d = mod.get_matrix("fake.bam", reso, biases=biases, normalization="decay", region1="chr1", tmpdir=tmp,
                           ncpus=1, verbose=False)
    t1 = time.perf_counter()
    fake_read_bam("fake.bam", 0, reso, half=False, tmpdir=tmp)
    t2 = time.perf_counter()
    out["runs"].append({"what": "get_matrix decay", "cells": len(d), "s": (t1 - t0) - (t2 - t1),
                        "us_per_cell": ((t1 - t0) - (t2 - t1)) / len(d) * 1e6})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
