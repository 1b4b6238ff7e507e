#!/usr/bin/env python
"""N3, extraction side, timed on one GPU: tb_region_cells and extraction.write_matrix on a synthetic
workload (wall clock of the API call: device passes + sort + copies to pageable host arrays).

    python tools/bench_extraction.py --workload c3
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps):
    best, out = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--window", type=int, default=2000, help="bins of the square window on the diagonal")
    args = ap.parse_args()
    import tadbit_b200
    from tadbit_b200 import extraction, workloads
    wl = workloads.workload(args.workload)
    reso = {"c2": 10000, "c3": 5000, "c3s": 5000, "c4": 1000}.get(args.workload, 10000)
    t0 = time.perf_counter()
    store, _ = workloads.build_sharded(wl, 0, 1, 0, None)
    build_s = time.perf_counter() - t0
    chroms = wl["chromosomes"]
    hic = tadbit_b200.HiC_data.from_store(store, chromosomes=chroms, resolution=reso)
    n = wl["size"]
    rng = np.random.default_rng(3)
    bias = rng.lognormal(0.0, 0.3, n)
    bad = (rng.random(n) < 0.02).astype(np.uint8)
    decay = np.empty(n)
    table = {}
    for crm, (beg, end) in hic.section_pos.items():
        decay[beg:end] = 900.0 * (np.arange(end - beg) + 1.0) ** -1.08
        table[crm] = dict(enumerate(decay[beg:end].tolist()))
    first = next(iter(chroms))
    beg, end = hic.section_pos[first]
    w0 = beg + (end - beg) // 3
    w1 = min(w0 + args.window, end)
    out = {"workload": wl["description"], "bins": n, "contacts_upper": int(store.info()["nnz_upper"]),
           "build_s": build_s, "runs": []}

    def run(tag, r0, r1, c0, c1, half, full=True):
        def count():
            return store.region_cells(r0, r1, c0, c1, bad=bad, half=half)
        def cells():
            return store.region_cells(r0, r1, c0, c1, bad=bad, bias=bias, decay=decay, half=half)
        t_raw, res = timed(count, args.reps)
        item = {"what": tag, "rows": r1 - r0, "cols": c1 - c0, "half": half, "cells": int(res[0].size),
                "raw_s": t_raw, "raw_cells_per_s": res[0].size / t_raw}
        del res
        if full:
            t_all, res = timed(cells, args.reps)
            item.update(norm_decay_s=t_all, norm_decay_cells_per_s=res[0].size / t_all,
                        host_bytes=int(sum(a.nbytes for a in res)))
            del res
        out["runs"].append(item)

    run("window on the diagonal", w0, w1, w0, w1, False)
    run("window rows x whole chromosome", w0, w1, beg, end, False)
    run("first chromosome, cis, half", beg, end, beg, end, True)
    run("first chromosome x genome, half", beg, end, 0, n, True, full=False)

    biases = dict(biases=dict(enumerate(bias.tolist())), badcol=dict((int(i), True) for i in np.flatnonzero(bad)),
                  decay=table, resolution=reso)
    with tempfile.TemporaryDirectory() as tmp:
        for norms in (("raw",), ("raw", "norm", "decay")):
            t0 = time.perf_counter()
            fn = extraction.write_matrix(hic, reso, biases, tmp, normalizations=norms, region1=first,
                                         start1=(w0 - beg) * reso, end1=(w1 - beg) * reso, verbose=False)
            dt = time.perf_counter() - t0
            lines = sum(1 for _ in open(list(fn.values())[0]))
            out["runs"].append({"what": "write_matrix %s, %d-bin window" % ("+".join(norms), w1 - w0), "s": dt,
                                "lines_per_file": lines, "files": len(fn),
                                "bytes": int(sum(os.path.getsize(p) for p in fn.values()))})
        t0 = time.perf_counter()
        d = extraction.get_matrix(hic, reso, biases=biases, region1=first, start1=(w0 - beg) * reso,
                                  end1=(w1 - beg) * reso, normalization="decay")
        out["runs"].append({"what": "get_matrix decay, %d-bin window (dictionary)" % (w1 - w0),
                            "s": time.perf_counter() - t0, "cells": len(d)})
        t0 = time.perf_counter()
        arr = extraction.get_matrix(hic, reso, biases=biases, region1=first, start1=(w0 - beg) * reso,
                                    end1=(w1 - beg) * reso, normalization="decay", as_arrays=True)
        out["runs"].append({"what": "get_matrix decay, %d-bin window (arrays)" % (w1 - w0),
                            "s": time.perf_counter() - t0, "cells": int(arr[0].size)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
