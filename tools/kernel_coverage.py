"""Small end-to-end pass over every kernel of the library on ONE GPU (a few seconds), written
as the target for ``compute-sanitizer`` (that tool is closed on the GPU pool, so the run is
self-checked: every store is built twice -- CSR released / CSR kept -- and the two must agree on
every result; export round trips, CSR-form constructors, batch vs loop).
``tests/test_gpu_coverage.py`` runs it.

Covers: synthetic generator, COO build (sorted fast path and general sort), chunk plan,
every stream encoding (dense D8/D16/D32 with overflow cells, S16/S32, sparse tiles),
both row-sum modes, filter statistics, ICE, batch ICE, normalised sum, diagonal sums, export,
region extraction (tb_region_cells against tb_window),
the compartment kernels (observed / expected visitor, both X X^T kernels, eigen-solver).
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


def keep_csr(make):
    """Build a store with TB_KEEP_CSR=1 (the CSR is not released after the build)."""
    old = os.environ.get("TB_KEEP_CSR")
    os.environ["TB_KEEP_CSR"] = "1"
    try:
        return make()
    finally:
        if old is None:
            del os.environ["TB_KEEP_CSR"]
        else:
            os.environ["TB_KEEP_CSR"] = old


def compartments_check(hic, hic2, chroms):
    """N4 on the first chromosome: the observed / expected visitor over the compact stream against
    the one over the kept CSR (same bits), X X^T as tensor-core tiles against the plain kernel,
    eigenpairs by their residuals."""
    from tadbit_b200._capi import bad_mask
    n = len(hic)
    beg, end = hic.section_pos[list(chroms)[0]]
    mask = bad_mask(n, hic.bads)
    bias = hic.bias.to_array()
    exp_arr = np.array([hic.expected.get(d, 1.0) or 1.0 for d in range(end - beg)])
    mats = []
    for h in (hic, hic2):
        h.store.corr_build(beg, end, bias, mask, exp_arr, stage=0)
        mats.append(h.store.corr_get())
    assert np.array_equal(mats[0], mats[1], equal_nan=True) and np.array_equal(mats[0], mats[0].T, equal_nan=True)
    if not np.all(np.isfinite(mats[0])) or mats[0].shape[0] < 8:
        return
    corr = {}
    for gemm in ("dmma", "simple"):
        hic.store.set_option("corr_gemm", gemm)
        hic.store.corr_build(beg, end, bias, mask, exp_arr, stage=1)
        corr[gemm] = hic.store.corr_get()
    hic.store.set_option("corr_gemm", "dmma")
    assert np.array_equal(corr["dmma"], corr["dmma"].T) and np.abs(corr["dmma"]).max() <= 1.0
    assert np.allclose(corr["dmma"], corr["simple"], rtol=0, atol=1e-12)
    vals, vecs, info = hic.store.corr_eig(3)
    assert info["converged"] and len(vals) == 3
    scale = max(abs(vals[0]), 1.0)
    for lam, v in zip(vals, vecs):
        assert np.linalg.norm(corr["dmma"] @ v - lam * v) <= 1e-10 * scale
        assert abs(np.linalg.norm(v) - 1.0) < 1e-12


def region_cells_check(hic, hic2):
    """N3, extraction side: tb_region_cells over the compact stream against the one over the kept
    CSR (same cells, same bits), against the dense window of tb_window, half vs full."""
    from tadbit_b200._capi import bad_mask
    n = len(hic)
    mask = bad_mask(n, hic.bads)
    bias = hic.bias.to_array()
    decay = 1.0 + np.arange(n, dtype=np.float64) % 97
    for (r0, r1, c0, c1) in ((0, n, 0, n), (n // 3, n // 2, 0, n // 4), (10, 700, 300, 1200), (n - 300, n, n - 300, n)):
        for half in (False, True):
            a = hic.store.region_cells(r0, r1, c0, c1, bad=mask, bias=bias, decay=decay, half=half)
            b = hic2.store.region_cells(r0, r1, c0, c1, bad=mask, bias=bias, decay=decay, half=half)
            for x, y in zip(a, b):
                assert np.array_equal(x, y, equal_nan=True)
            key = a[0].astype(np.int64) * n + a[1]
            assert np.all(np.diff(key) > 0) and not mask[a[0]].any() and not mask[a[1]].any()
            if half:
                assert np.all(a[0] - r0 <= a[1] - c0)
        if (r1 - r0) * (c1 - c0) <= 1 << 22:
            dense = hic.store.window(r0, r1, c0, c1, bias=bias)
            keep = ~mask[r0:r1, None].astype(bool) & ~mask[None, c0:c1].astype(bool)
            rr, cc, _, nrm, _ = hic.store.region_cells(r0, r1, c0, c1, bad=mask, bias=bias)
            got = np.zeros_like(dense)
            got[rr - r0, cc - c0] = nrm
            assert np.array_equal(got, np.where(keep, dense, 0.0))


def run_path(make, chroms, tag):
    """The pipeline on the store `make()` builds (CSR released when every value is positive: every
    kernel decodes the compact stream) and on its twin that kept the CSR; the two must agree --
    integers exactly, ICE to 1e-12 (also against the CSR-walk row sum)."""
    store, twin = make(), keep_csr(make)
    assert twin.timing()["csr_resident"]
    res = []
    for st in (store, twin):
        hic = HiC_data.from_store(st, chromosomes=chroms)
        hic.filter_columns(silent=True)
        hic.normalize_hic(iterations=20, max_dev=1e-5, silent=True)
        hic.normalize_expected()
        res.append((hic, hic.bias.to_array().copy()))
    (hic, bias), (hic2, bias2) = res
    assert hic.bads == hic2.bads and hic.expected == hic2.expected
    assert len(hic.last_trace) == len(hic2.last_trace)
    for x, y in zip(store.export_upper(), twin.export_upper()):
        assert np.array_equal(x, y)
    for x, y in zip(store.export_csr("full_rows"), twin.export_csr("full_rows")):
        assert np.array_equal(x, y)
    assert hic.sum() == hic2.sum()
    region_cells_check(hic, hic2)
    compartments_check(hic, hic2, chroms)
    twin.set_option("rowsum", "csr")
    hic2.normalize_hic(iterations=20, max_dev=1e-5, silent=True)
    twin.set_option("rowsum", "compact")
    rel = max(float(np.max(np.abs(bias2 - bias) / np.abs(bias))),
              float(np.max(np.abs(hic2.bias.to_array() - bias) / np.abs(bias))))
    lay = store.layout()
    print("%-10s csr_kept=%d bads=%d passes=%d compact-vs-csr=%.1e enc=%s overflow=%d tile_cells=%d"
          % (tag, store.timing()["csr_resident"], len(hic.bads), len(hic.last_trace), rel, lay["encodings"],
             lay["overflow_cells"], lay["tile_cells"]))
    assert rel < 1e-12, rel
    twin.close()
    return hic, store


def main():
    if tadbit_b200.device_count() < 1:
        raise SystemExit("needs a CUDA device")
    # sparse + dense mix: three chromosomes, 6000 bins (two column panels, several tiles)
    chroms = OrderedDict([("chrA", 3000), ("chrB", 2000), ("chrC", 1000)])
    n = sum(chroms.values())
    p = SynthParams(seed=11, cis_scale=900.0, cis_alpha=1.08, trans_mean=0.03, vis_sigma=0.35,
                    low_frac=0.03, low_scale=0.02, empty_frac=0.01)
    _, syn = run_path(lambda: ContactStore.synthetic(n, p, chromosomes=chroms), chroms, "synthetic")
    assert not syn.timing()["csr_resident"]

    # the same cells through the COO constructor: sorted input, then shuffled input with a few
    # huge values (S32 / S16 chunks, overflow cells of dense panels) and negative ones
    rows, cols, vals = syn.export_upper()
    a, coo = run_path(lambda: ContactStore.from_coo(n, rows, cols, vals, chromosomes=chroms), chroms,
                      "coo-sorted")
    # the same cells in CSR form (upper triangle, 8 bytes per cell) and as whole rows
    ptr, cc, vv = syn.export_csr("upper")
    up = ContactStore.from_csr(n, ptr, cc, vv, form="upper", chromosomes=chroms)
    for x, y in zip(up.export_upper(), (rows, cols, vals)):
        assert np.array_equal(x, y)
    fptr, fc, fv = syn.export_csr("full_rows")
    assert fptr[-1] == syn.info()["nnz_full"]
    full = ContactStore.from_csr(n, fptr, fc, fv, form="full_rows", chromosomes=chroms)
    for x, y in zip(full.export_upper(), (rows, cols, vals)):
        assert np.array_equal(x, y)
    assert np.array_equal(full.ice(None, 20, 1e-5)[0], coo.ice(None, 20, 1e-5)[0])
    up.close()
    full.close()
    rng = np.random.RandomState(5)
    perm = rng.permutation(rows.size)
    v2 = vals.copy()
    big = rng.choice(rows.size, 200, replace=False)
    v2[big[:100]] = 70000 + rng.randint(0, 1000, 100)
    v2[big[100:150]] = 300000 + rng.randint(0, 1000, 50)
    v2[big[150:190]] = 1 << 20
    v2[big[190:195]] = -3
    v2[big[195:]] = 40000 + rng.randint(0, 1000, 5)
    _, shuf = run_path(lambda: ContactStore.from_coo(n, rows[perm], cols[perm], v2[perm], chromosomes=chroms),
                       chroms, "coo-shuf")
    assert shuf.timing()["csr_resident"]                 # negative values: the CSR stays
    r2, c2, w2 = shuf.export_upper()
    assert np.array_equal(r2, rows) and np.array_equal(c2, cols) and np.array_equal(w2, v2)
    # the same with positive values only: CSR released, the visitors decode S16 / S32 chunks and the
    # overflow lists of dense panels
    v2p = np.abs(v2)
    _, big = run_path(lambda: ContactStore.from_coo(n, rows[perm], cols[perm], v2p[perm], chromosomes=chroms),
                      chroms, "coo-big")
    assert not big.timing()["csr_resident"]
    r2, c2, w2 = big.export_upper()
    assert np.array_equal(r2, rows) and np.array_equal(c2, cols) and np.array_equal(w2, v2p)
    big.close()

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
    _, skew = run_path(lambda: ContactStore.from_coo(n, r3, c3, w3, chromosomes=chroms), chroms, "edges+skew")
    r4, c4, w4 = skew.export_upper()
    order = np.lexsort((c3, r3))
    assert np.array_equal(r4, r3[order]) and np.array_equal(c4, c3[order]) and np.array_equal(w4, w3[order])
    w3p = np.where(w3 <= 0, 1, w3)                        # same boundaries without zeros / negatives
    _, skewp = run_path(lambda: ContactStore.from_coo(n, r3, c3, w3p, chromosomes=chroms), chroms, "edges-pos")
    assert not skewp.timing()["csr_resident"]
    r4, c4, w4 = skewp.export_upper()
    assert np.array_equal(r4, r3[order]) and np.array_equal(c4, c3[order]) and np.array_equal(w4, w3p[order])
    skewp.close()

    # dense single chromosome (D8 / D16 panels)
    chrom1 = OrderedDict([("chr1", 5000)])
    pd = SynthParams(seed=3, cis_scale=3.0e4, cis_alpha=1.08, trans_mean=0.0, vis_sigma=0.35,
                     low_frac=0.02, low_scale=0.02, empty_frac=0.005)
    _, dense = run_path(lambda: ContactStore.synthetic(5000, pd, chromosomes=chrom1), chrom1, "dense")

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
