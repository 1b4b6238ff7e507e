#!/usr/bin/env python
"""Host-side share of the pipeline methods at a BASELINE size, without a GPU: the store is replaced
by an object that returns ready-made arrays of the right shapes at once, so what is timed is only
what tadbit_b200's Python layer (and its host-only library helpers) does around the kernels --
thresholds, the numpy tail of filter_by_mean, dictionaries of the reference's return types, the
pooling walks (tb_pool_expected / tb_pool_decay).

    python tools/profile_host_tails.py [c3|c4]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import warnings
    warnings.filterwarnings("ignore")
    import tadbit_b200
    from tadbit_b200 import decay, workloads
    wl = workloads.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
    chroms, n = wl["chromosomes"], wl["size"]
    rng = np.random.default_rng(0)
    vis = rng.lognormal(0, 0.35, n)
    vis[rng.random(n) < 0.02] *= 0.02
    vis[rng.random(n) < 0.005] = 0
    sums = (vis * 6400).astype(np.int64)
    stored = np.where(rng.random(n) < 0.3, 10, n // 50)    # ~30 % of the bins nearly empty
    bias = rng.lognormal(0, 0.3, n)
    raw = np.zeros(n, dtype=np.int64)
    beg = 0
    for c in chroms:
        raw[beg:beg + chroms[c]] = rng.poisson(4e6 * (np.arange(chroms[c]) + 1.0) ** -1.08)
        beg += chroms[c]
    nrm = raw * rng.random(n)

    class Arrays(object):                                   # what the kernels would hand back
        size = n

        def row_stats(self):
            return stored, sums

        def col_sums_masked(self, mask):
            return sums.copy()

        def ice(self, bad=None, iterations=0, max_dev=1e-5):
            return bias.copy(), np.ones((389, 4))

        def norm_sum(self, bias=None, bad=None):
            return 3.3e11

        def diag_sums(self, bad, ndiag):
            return raw[:ndiag]

        def decay_sums(self, bias, bad=None):
            return nrm, raw, (1e9, 2e9)

    hic = tadbit_b200.HiC_data.from_store(Arrays(), chromosomes=chroms)
    out = {"workload": wl["description"], "bins": n, "runs": []}
    for rep in range(3):
        t = [time.perf_counter()]
        hic.filter_columns(silent=True, **wl["filter_kw"])
        t.append(time.perf_counter())
        hic.normalize_hic(iterations=500, max_dev=1e-5, silent=True)
        t.append(time.perf_counter())
        hic.normalize_expected()
        t.append(time.perf_counter())
        decay.rescale_and_decay(hic, hic.bias, hic.bads)
        t.append(time.perf_counter())
        out["runs"].append(dict(zip(("filter_columns_ms", "normalize_hic_ms", "normalize_expected_ms",
                                     "rescale_and_decay_ms"), [(b - a) * 1e3 for a, b in zip(t, t[1:])])))
    out["bad_columns"] = len(hic.bads)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
