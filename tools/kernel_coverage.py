"""Small end-to-end pass over every kernel of the library on ONE GPU (a few seconds), written
as the target for ``compute-sanitizer`` (that tool is closed on the round-1 GPU pool, so the
run is only self-checked: compact stream vs plain CSR, export round trip, batch vs loop).
``tests/test_gpu_coverage.py`` runs it.

Covers: synthetic generator, COO build (sorted fast path and general sort), chunk plan,
every stream encoding (dense D8/D16/D32 with overflow cells, S16/S32, sparse tiles),
both row-sum modes, filter statistics, ICE, batch ICE, normalised sum, diagonal sums, export.
Results are cross-checked between the two row-sum modes (no oracle: nothing here imports
``oracle/``)."""
import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import tadbit_b200  # noqa: E402
from tadbit_b200 import ContactStore, HiC_data, SynthParams  # noqa: E402
from tadbit_b200.batch import normalize_hic_batch  # noqa: E402


def run_path(store, chroms, tag):
    hic = HiC_data.from_store(store, chromosomes=chroms)
    hic.filter_columns(silent=True)
    hic.normalize_hic(iterations=20, max_dev=1e-5, silent=True)
    hic.normalize_expected()
    bias = hic.bias.to_array().copy()
    store.set_option("rowsum", "csr")
    hic.normalize_hic(iterations=20, max_dev=1e-5, silent=True)
    store.set_option("rowsum", "compact")
    rel = float(np.max(np.abs(hic.bias.to_array() - bias) / np.abs(bias)))
    lay = store.layout()
    print("%-10s bads=%d passes=%d compact-vs-csr=%.1e enc=%s overflow=%d tile_cells=%d"
          % (tag, len(hic.bads), len(hic.last_trace), rel, lay["encodings"],
             lay["overflow_cells"], lay["tile_cells"]))
    assert rel < 1e-12, rel
    return hic


def main():
    if tadbit_b200.device_count() < 1:
        raise SystemExit("needs a CUDA device")
    # sparse + dense mix: three chromosomes, 6000 bins (two column panels, several tiles)
    chroms = OrderedDict([("chrA", 3000), ("chrB", 2000), ("chrC", 1000)])
    n = sum(chroms.values())
    p = SynthParams(seed=11, cis_scale=900.0, cis_alpha=1.08, trans_mean=0.03, vis_sigma=0.35,
                    low_frac=0.03, low_scale=0.02, empty_frac=0.01)
    syn = ContactStore.synthetic(n, p, chromosomes=chroms)
    run_path(syn, chroms, "synthetic")

    # the same cells through the COO constructor: sorted input, then shuffled input with a few
    # huge values (S32 / S16 chunks, overflow cells of dense panels) and negative ones
    rows, cols, vals = syn.export_upper()
    coo = ContactStore.from_coo(n, rows, cols, vals, chromosomes=chroms)
    a = run_path(coo, chroms, "coo-sorted")
    rng = np.random.RandomState(5)
    perm = rng.permutation(rows.size)
    v2 = vals.copy()
    big = rng.choice(rows.size, 200, replace=False)
    v2[big[:100]] = 70000 + rng.randint(0, 1000, 100)
    v2[big[100:150]] = 300000 + rng.randint(0, 1000, 50)
    v2[big[150:190]] = 1 << 20
    v2[big[190:195]] = -3
    v2[big[195:]] = 40000 + rng.randint(0, 1000, 5)
    shuf = ContactStore.from_coo(n, rows[perm], cols[perm], v2[perm], chromosomes=chroms)
    run_path(shuf, chroms, "coo-shuf")
    r2, c2, w2 = shuf.export_upper()
    assert np.array_equal(r2, rows) and np.array_equal(c2, cols) and np.array_equal(w2, v2)

    # value boundaries of the encodings (D8 byte, tile cell 2^15, D16, int32) placed on dense rows,
    # and a few very long sparse rows among short ones (tiles split by their busiest warp)
    edge = np.array([255, 256, 32767, 32768, 65535, 65536, 2147483647, -1, -3, 0], dtype=np.int64)
    v3 = vals.copy()
    pick = rng.choice(rows.size, 50 * edge.size, replace=False)
    v3[pick] = np.tile(edge, 50).astype(np.int32)
    long_rows = np.array([7, 1033, 2050, 4100, 5999])
    extra_r, extra_c = [], []
    have = set(zip(rows.tolist(), cols.tolist()))
    for r in long_rows:
        for c in range(0, n, 2):
            i, j = (r, c) if r <= c else (c, r)
            if (i, j) not in have:
                have.add((i, j))
                extra_r.append(i)
                extra_c.append(j)
    r3 = np.concatenate([rows, np.array(extra_r, dtype=np.int32)])
    c3 = np.concatenate([cols, np.array(extra_c, dtype=np.int32)])
    w3 = np.concatenate([v3, np.ones(len(extra_r), dtype=np.int32)])
    skew = ContactStore.from_coo(n, r3, c3, w3, chromosomes=chroms)
    run_path(skew, chroms, "edges+skew")
    r4, c4, w4 = skew.export_upper()
    order = np.lexsort((c3, r3))
    assert np.array_equal(r4, r3[order]) and np.array_equal(c4, c3[order]) and np.array_equal(w4, w3[order])

    # dense single chromosome (D8 / D16 panels)
    chrom1 = OrderedDict([("chr1", 5000)])
    pd = SynthParams(seed=3, cis_scale=3.0e4, cis_alpha=1.08, trans_mean=0.0, vis_sigma=0.35,
                     low_frac=0.02, low_scale=0.02, empty_frac=0.005)
    dense = ContactStore.synthetic(5000, pd, chromosomes=chrom1)
    run_path(dense, chrom1, "dense")

    # batch of small matrices
    hics = []
    for k in range(5):
        ck = OrderedDict([("c", 300 + 40 * k)])
        pk = SynthParams(seed=100 + k, cis_scale=200.0, cis_alpha=1.08, trans_mean=0.0,
                         vis_sigma=0.35, low_frac=0.03, low_scale=0.02, empty_frac=0.01)
        h = HiC_data.from_store(ContactStore.synthetic(300 + 40 * k, pk, chromosomes=ck),
                                chromosomes=ck)
        h.filter_columns(silent=True)
        hics.append(h)
    normalize_hic_batch(hics, iterations=20, max_dev=1e-5, silent=True)
    for h in hics:
        b = h.bias.to_array().copy()
        h.normalize_hic(iterations=20, max_dev=1e-5, silent=True)
        assert np.allclose(b, h.bias.to_array(), rtol=1e-12, atol=0)
    print("batch ok; raw sum %d, normalised sum %.6f" % (a.sum(), a.sum(a.bias)))
    for s in (syn, coo, shuf, skew, dense):
        s.close()
    from tadbit_b200._capi import trim_memory
    trim_memory()
    print("KERNEL_COVERAGE_OK")


if __name__ == "__main__":
    main()
