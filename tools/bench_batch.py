#!/usr/bin/env python
"""BASELINE configs[4]: 64 synthetic chr1 @40 kb matrices (N = 6224, dense) normalised at once
on one B200 (tb_ice_batch) versus one after the other (tb_ice).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tadbit_b200                                        # noqa: E402
from tadbit_b200 import workloads                          # noqa: E402


def main():
    nmat = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    size = workloads.HG38[0] // 40000 + 1
    stores = []
    for seed in range(1, nmat + 1):
        p = tadbit_b200.SynthParams(seed=seed, cis_scale=6.0e4, cis_alpha=1.08, trans_mean=0.0,
                                    vis_sigma=0.35, low_frac=0.02, low_scale=0.02, empty_frac=0.005)
        stores.append(tadbit_b200.ContactStore.synthetic(size, p))
    contacts = sum(s.info()["nnz_upper"] for s in stores)
    its, dev = 100, 1e-5
    # one after the other
    for s in stores[:2]:
        s.ice(None, its, dev)
    t0 = time.perf_counter()
    loop_passes, loop_bias = 0, []
    for s in stores:
        b, tr = s.ice(None, its, dev)
        loop_passes += len(tr)
        loop_bias.append(b)
    t_loop = time.perf_counter() - t0
    # batch
    batch = tadbit_b200.batch_store(stores)
    sizes = [size] * nmat
    tadbit_b200.ice_batch(batch, sizes, None, its, dev)
    t0 = time.perf_counter()
    bias, traces, passes = tadbit_b200.ice_batch(batch, sizes, None, its, dev)
    t_batch = time.perf_counter() - t0
    dev_ms = batch.ice_timing()["loop_ms"]
    err = max(float(np.max(np.abs(bias[k * size:(k + 1) * size] - loop_bias[k]) / loop_bias[k]))
              for k in range(nmat))
    work = sum(s.info()["nnz_upper"] * int(p) for s, p in zip(stores, passes))
    print(json.dumps({"workload": "%d x synthetic chr1 @40kb (N=%d)" % (nmat, size),
                      "contacts_total": contacts, "passes": [int(p) for p in passes[:8]],
                      "loop_s": t_loop, "batch_s": t_batch, "batch_device_ms": dev_ms,
                      "loop_contacts_it_per_s": work / t_loop, "batch_contacts_it_per_s": work / t_batch,
                      "max_rel_diff_batch_vs_loop": err}))


if __name__ == "__main__":
    main()
