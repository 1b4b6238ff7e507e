#!/usr/bin/env python
"""Build a synthetic workload on cuda:0 and run the one-pass reductions once each (K4 masked column
sums, K5 distance sums, export): the target of `ncu -k regex:k_visit` captures.

    python tools/run_onepass.py c3s
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tadbit_b200 import ContactStore, workloads  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3s"
    wl = workloads.workload(name)
    store = ContactStore.synthetic(wl["size"], wl["params"], chromosomes=wl["chromosomes"])
    n = wl["size"]
    bad = np.zeros(n, dtype=np.uint8)
    ndiag = max(wl["chromosomes"].values())
    for rep in range(3):
        t0 = time.perf_counter()
        d = store.diag_sums(bad, ndiag)
        t1 = time.perf_counter()
        print("K5 diag sums: kernel %.3f ms, call %.1f ms, total %d" % (store.timing()["kernel_ms"], (t1 - t0) * 1e3, int(d.sum())))
    t0 = time.perf_counter()
    c = store.col_sums_masked(bad)
    print("K4 masked column sums: kernel %.3f ms, call %.1f ms, total %d" % (store.timing()["kernel_ms"], (time.perf_counter() - t0) * 1e3, int(c.sum())))
    lay = store.layout()
    print("stream bytes %d (chunks %d, tiles %d), upper cells %d" % (lay["stream_bytes"], lay["chunk_bytes"], lay["tile_bytes"], store.info()["nnz_upper"]))
    if os.environ.get("WITH_EXPORT"):
        t0 = time.perf_counter()
        r, cc, v = store.export_upper()
        print("export: %.1f ms for %d cells" % ((time.perf_counter() - t0) * 1e3, r.size))


if __name__ == "__main__":
    main()
