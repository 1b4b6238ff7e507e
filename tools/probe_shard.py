#!/usr/bin/env python
"""K1 of ONE shard of a row-sharded bench workload, on one GPU: what a GPU of an N-GPU run executes
per pass (timing only; the shard is not connected to a group).  Prints tb_probe_rowsum (tile kernel,
dense kernel) and tb_probe_k1 (kernels back to back vs on two streams).

    python tools/probe_shard.py --workload c3 --world 8 --ranks 0,3,7
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--ranks", default="0,3,7")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    from tadbit_b200 import workloads
    from tadbit_b200.store import ContactStore
    wl = workloads.workload(args.workload)
    n = wl["size"]
    counts = ContactStore.synthetic_row_counts(n, wl["params"], chromosomes=wl["chromosomes"],
                                               row_block=(0, n), device=0)
    blocks = workloads.balanced_row_blocks(counts, args.world)
    out = {"workload": wl["description"], "world": args.world, "shards": []}
    for r in [int(x) for x in args.ranks.split(",")]:
        store = ContactStore.synthetic(n, wl["params"], chromosomes=wl["chromosomes"],
                                       row_block=blocks[r], device=0)
        total, tiles, dense = store.probe_rowsum(args.reps)
        serial, streams = store.probe_k1(args.reps)
        serial2, streams2 = store.probe_k1(args.reps)
        out["shards"].append({"rank": r, "rows": list(blocks[r]), "tile_ms": tiles, "dense_ms": dense,
                              "serial_ms": [serial, serial2], "two_stream_ms": [streams, streams2]})
        store.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
