#!/usr/bin/env python
"""Build one synthetic workload on cuda:0 and print the K1 split (total, sparse tiles, dense
kernels) from tb_probe_rowsum, the stream layout and the TB_TIMING build stages.

    python tools/probe_k1.py c3 [reps]
    PROBE_SHARD=3/8 python tools/probe_k1.py c3      # the row block rank 3 of 8 would own (cell-balanced):
                                                     # K1 of a small shard measured on ONE GPU
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tadbit_b200 import ContactStore, HiC_data, workloads  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    wl = workloads.workload(name)
    if os.environ.get("PROBE_WARM"):          # a first build warms the memory pool
        ContactStore.synthetic(wl["size"], wl["params"], chromosomes=wl["chromosomes"]).close()
        print("---- second (warm) build ----", file=sys.stderr)
    block = None
    if os.environ.get("PROBE_SHARD"):
        import numpy as np
        r, w = [int(v) for v in os.environ["PROBE_SHARD"].split("/")]
        counts = ContactStore.synthetic_row_counts(wl["size"], wl["params"], chromosomes=wl["chromosomes"])
        block = workloads.balanced_row_blocks(counts, w)[r]
        print("shard %d/%d: rows %s, %.3g of %.3g cells" % (r, w, block, counts[block[0]:block[1]].sum(), counts.sum()))
    t0 = time.time()
    store = ContactStore.synthetic(wl["size"], wl["params"], chromosomes=wl["chromosomes"], row_block=block)
    t1 = time.time()
    lay = store.layout()
    print("workload %s: N=%d nnz_upper=%d build %.2f s" % (name, wl["size"], store.info()["nnz_upper"], t1 - t0))
    print("layout", lay)
    store.probe_rowsum(2)
    total, tiles, dense = store.probe_rowsum(reps)
    print("K1 ms: total %.3f  tiles %.3f  dense %.3f" % (total, tiles, dense))
    if os.environ.get("PROBE_ONLY") or block is not None:
        return
    hic = HiC_data.from_store(store, chromosomes=wl["chromosomes"])
    hic.filter_columns(silent=True)
    t0 = time.time()
    hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
    t1 = time.time()
    tm = store.timing()
    print("normalize_hic %.1f ms wall; loop %.1f ms, %d passes, K1 %.3f ms/pass"
          % ((t1 - t0) * 1e3, tm["loop_ms"], len(hic.last_trace),
             tm["rowsum_ms"] / max(tm["rowsum_launches"], 1)))


if __name__ == "__main__":
    main()
