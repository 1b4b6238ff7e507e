#!/usr/bin/env python
"""K1 with its two kernels back to back (k1_streams=1) or on two streams (2): ms per pass of the ICE
loop on a bench workload, alternating, and the biases compared bit for bit.

    python tools/probe_streams.py --workload c3 --passes 60 --rounds 3
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
    ap.add_argument("--passes", type=int, default=60)
    ap.add_argument("--rounds", type=int, default=3)
    args = ap.parse_args()
    import tadbit_b200
    from tadbit_b200 import workloads
    from tadbit_b200._capi import bad_mask
    wl = workloads.workload(args.workload)
    store, _ = workloads.build_sharded(wl, 0, 1, 0, None)
    hic = tadbit_b200.HiC_data.from_store(store, chromosomes=wl["chromosomes"])
    hic.filter_columns(silent=True, **wl["filter_kw"])
    mask = bad_mask(wl["size"], hic.bads)
    out = {"workload": wl["description"], "runs": []}
    ref = None
    store.ice(mask, iterations=5, max_dev=0.0)          # warm-up
    for r in range(args.rounds):
        for mode in ("1", "2"):
            store.set_option("k1_streams", mode)
            bias, trace = store.ice(mask, iterations=args.passes - 1, max_dev=0.0)
            tm = store.timing()
            n = max(len(trace), 1)
            same = None
            if ref is None:
                ref = bias.copy()
            else:
                same = bool(np.array_equal(ref, bias))
            out["runs"].append({"k1_streams": mode, "passes": n, "ms_per_pass": tm["loop_ms"] / n,
                                "k1_ms": tm["rowsum_ms"] / max(tm["rowsum_launches"], 1),
                                "exchange_ms": tm["exchange_ms"] / n, "bias_bits_equal_first_run": same})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
