"""Parity at BASELINE sizes, against the oracle on the cells the store exports.

* a 20k-bin multi-chromosome synthetic matrix against the numpy restatement + C port;
* the dense chr1 @10 kb workload (BASELINE configs[1], 3e8 contacts) and the genome-wide 5 kb
  workload of the bench (configs[2], 1.97e9 contacts) at FULL size against oracle/upper_port.c
  (the restatement over raw upper-triangle arrays, pinned to the reference's golden vectors in
  tests/test_oracle_port.py): filter mask bit-exact (dict values included), biases within 1e-10
  relative with the same number of passes, expected dict bit-exact, integer reductions exact;
* 64 x chr1 @40 kb (configs[4]) normalised as one batch against the oracle, matrix by matrix;
* the two copies of the matrix a store can walk (compact stream vs CSR) against each other for
  every kernel that has both paths.

Tolerance: north_star's 1e-10 relative for fp64 biases; everything integer is compared exactly.
"""
import os
from collections import OrderedDict

import numpy as np
import pytest

import tadbit_oracle as orc

pytestmark = pytest.mark.gpu


def _host_gb():
    try:
        with open("/proc/meminfo") as fh:
            for line in fh:
                if line.startswith("MemAvailable"):
                    return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def _mid(keep_csr=False, seed=11):
    import tadbit_b200
    chroms = OrderedDict([("c1", 9000), ("c2", 6500), ("c3", 4500)])
    n = sum(chroms.values())
    params = tadbit_b200.SynthParams(seed=seed, cis_scale=900.0, cis_alpha=1.08, trans_mean=0.004,
                                     vis_sigma=0.35, low_frac=0.03, low_scale=0.02, empty_frac=0.01)
    old = os.environ.get("TB_KEEP_CSR")
    if keep_csr:
        os.environ["TB_KEEP_CSR"] = "1"
    try:
        store = tadbit_b200.ContactStore.synthetic(n, params, chromosomes=chroms)
    finally:
        if keep_csr:
            if old is None:
                del os.environ["TB_KEEP_CSR"]
            else:
                os.environ["TB_KEEP_CSR"] = old
    return tadbit_b200.HiC_data.from_store(store, chromosomes=chroms), chroms, n


def test_mid_size_against_oracle():
    hic, chroms, n = _mid()
    lay = hic.store.layout()
    assert lay["encodings"]["T"] > 0 and lay["encodings"]["D8"] > 0       # both K1 kernels in play
    assert not hic.store.timing()["csr_resident"]                         # positive counts: CSR released
    assert lay["csr_bytes"] == 0
    r, c, v = hic.store.export_upper()
    key = r.astype(np.int64) * n + c
    assert np.all(np.diff(key) > 0) and np.all(r <= c) and np.all(v > 0)   # sorted, unique, upper
    m = orc.OracleMatrix(n, r, c, v, chromosomes=chroms)
    hic.filter_columns(silent=True)
    bads = orc.filter_columns(m, silent=True)
    assert hic.bads == bads
    hic.normalize_hic(iterations=100, max_dev=1e-6, silent=True, factor=None)
    bias, trace, _, _ = orc.iterative_fast(m, bads, 100, 1e-6)
    assert len(hic.last_trace) == len(trace)
    np.testing.assert_allclose(hic.bias.to_array(), bias, rtol=1e-10, atol=0)
    hic.normalize_expected()
    assert hic.expected == orc.expected(m, bads=bads)
    nnz, rsum = hic.store.row_stats()
    onnz, orsum = orc.row_stats(m)
    assert np.array_equal(nnz, onnz) and np.array_equal(rsum, orsum)
    assert hic.sum() == int(orc.matrix_sum(m, None, bads))


def test_compact_stream_and_csr_agree():
    """Same matrix twice: CSR released (every kernel decodes the compact stream) and CSR kept
    (TB_KEEP_CSR=1: the CSR walks).  Integer results identical, ICE within 1e-12."""
    a, chroms, n = _mid()
    b, _, _ = _mid(keep_csr=True)
    assert b.store.timing()["csr_resident"] and not a.store.timing()["csr_resident"]
    for x, y in zip(a.store.export_upper(), b.store.export_upper()):
        assert np.array_equal(x, y)
    rng = np.random.default_rng(3)
    bad = (rng.random(n) < 0.05).astype(np.uint8)
    assert np.array_equal(a.store.col_sums_masked(bad), b.store.col_sums_masked(bad))
    assert np.array_equal(a.store.diag_sums(bad, 9000), b.store.diag_sums(bad, 9000))
    assert a.store.norm_sum(None, bad) == b.store.norm_sum(None, bad)
    for x, y in zip(a.store.row_stats(), b.store.row_stats()):
        assert np.array_equal(x, y)
    ba, ta = a.store.ice(bad, 60, 1e-6)
    bb, tb = b.store.ice(bad, 60, 1e-6)
    b.store.set_option("rowsum", "csr")
    bc, tc = b.store.ice(bad, 60, 1e-6)
    assert len(ta) == len(tb) == len(tc)
    assert np.array_equal(ba, bb)                                          # same kernels, same order
    np.testing.assert_allclose(bc, ba, rtol=1e-12, atol=0)
    with pytest.raises(Exception):
        a.store.set_option("rowsum", "csr")                                # no CSR left to walk


def test_fused_pass_equals_separate_steps():
    """k_ice_exchange (one launch per pass) against k_ice_reduce + k_ice_update."""
    hic, chroms, n = _mid(seed=12)
    rng = np.random.default_rng(4)
    bad = (rng.random(n) < 0.03).astype(np.uint8)
    b1, t1 = hic.store.ice(bad, 100, 1e-7)
    assert hic.store.timing()["exchange"] == "local/nccl"                  # one GPU: local buffers
    launches_fused = hic.store.timing()["total_launches"]
    hic.store.set_option("exchange", "nccl")
    b2, t2 = hic.store.ice(bad, 100, 1e-7)
    assert hic.store.timing()["total_launches"] > launches_fused
    assert len(t1) == len(t2) and len(t1) > 5
    np.testing.assert_allclose(b2, b1, rtol=1e-13, atol=0)
    np.testing.assert_allclose(t2, t1, rtol=1e-12, atol=0)
    for its in (0, 1, 3):                                                  # pass-count edge cases
        hic.store.set_option("exchange", "fused")
        f, tf = hic.store.ice(bad, its, 1e-9)
        hic.store.set_option("exchange", "nccl")
        g, tg = hic.store.ice(bad, its, 1e-9)
        assert len(tf) == len(tg) == (0 if its == 0 else its + 1)
        np.testing.assert_allclose(g, f, rtol=1e-13, atol=0)


def _full_size_case(name):
    import torch
    import tadbit_b200
    from tadbit_b200 import workloads
    from tadbit_b200._capi import trim_memory
    trim_memory()                                   # hand cached blocks of earlier tests back first
    wl = workloads.workload(name)
    need_gpu, need_host = (20, 24) if name == "c2" else (110, 120)
    if torch.cuda.mem_get_info(0)[0] < need_gpu * (1 << 30):
        pytest.skip("needs ~%d GB of free device memory" % need_gpu)
    if _host_gb() < need_host:
        pytest.skip("needs ~%d GB of free host memory" % need_host)
    store = tadbit_b200.ContactStore.synthetic(wl["size"], wl["params"], chromosomes=wl["chromosomes"])
    return wl, store, tadbit_b200.HiC_data.from_store(store, chromosomes=wl["chromosomes"])


def _check_against_upper_port(wl, store, hic, iterations, max_dev):
    n = wl["size"]
    info = store.info()
    r, c, v = store.export_upper()
    assert r.size == info["nnz_upper"]
    m = orc.UpperCells(n, r, c, v, chromosomes=wl["chromosomes"])
    # integer reductions, exact
    nnz, rsum = store.row_stats()
    onnz, orsum = orc.fast_row_stats(m)
    assert np.array_equal(nnz, onnz) and np.array_equal(rsum, orsum)
    assert int(nnz.sum()) == info["nnz_full"]
    # filter_columns: zero-count + by-mean, mask and dict values bit-exact
    hic.filter_columns(silent=True)
    bads = orc.fast_filter_columns(m, silent=True)
    assert hic.bads == bads
    assert hic.sum() == int(orc.fast_norm_sum(m, np.ones(n), bads))        # raw sum over good cells
    # ICE + factor rescale
    hic.normalize_hic(iterations=iterations, max_dev=max_dev, silent=True)
    bias, trace = orc.fast_normalize_hic(m, bads, iterations, max_dev)
    assert len(hic.last_trace) == len(trace)
    got = hic.bias.to_array()
    rel = np.max(np.abs(got - bias) / np.abs(bias))
    assert rel < 1e-10, rel
    assert abs(hic.sum(hic.bias) / float(n * n) - 1.0) < 1e-9             # hic_data.py:406-412
    # expected: dict bit-exact
    hic.normalize_expected()
    assert hic.expected == orc.fast_expected(m, bads)
    return len(trace), float(hic.last_trace[-1][3]), rel, len(bads)


def test_full_size_c2_against_oracle():
    """BASELINE configs[1]: chr1 @10 kb, 2.99e8 stored contacts."""
    wl, store, hic = _full_size_case("c2")
    try:
        assert store.info()["nnz_upper"] > 2.5e8
        passes, dev, rel, nbad = _check_against_upper_port(wl, store, hic, 100, 1e-5)
        assert passes <= 101 and nbad > 0
        print("c2: passes %d final dev %.3g max rel bias err %.2e bads %d" % (passes, dev, rel, nbad))
    finally:
        store.close()


def test_full_size_c2_compartments_against_numpy_and_arpack():
    """N4 at BASELINE configs[1] size: chr1 @10 kb, ~24 300 good bins.  Observed / expected rows
    against the exported cells (same divisions: bit for bit), rows of the correlation matrix against
    numpy on the host, and the eigenpairs against the reference's own call, scipy's
    eigsh(matrix, k=3) (ARPACK), on the correlation matrix the device produced."""
    from scipy.sparse.linalg import eigsh
    from tadbit_b200._capi import bad_mask
    wl, store, hic = _full_size_case("c2")
    if _host_gb() < 48:
        store.close()
        pytest.skip("needs ~48 GB of free host memory")
    try:
        n = wl["size"]
        hic.filter_columns(silent=True)
        hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
        hic.normalize_expected()
        bias = hic.bias.to_array()
        mask = bad_mask(n, hic.bads)
        exp_arr = np.array([hic.expected[d] or 1.0 for d in range(n)])
        good = np.flatnonzero(mask == 0)
        pos = np.full(n, -1, dtype=np.int64)
        pos[good] = np.arange(good.size)
        ng = store.corr_build(0, n, bias, mask, exp_arr, stage=0)
        assert ng == good.size and ng > 24000
        oe = store.corr_get()
        assert np.array_equal(oe, oe.T)
        r, c, v = store.export_upper()
        rng = np.random.default_rng(5)
        for i in rng.choice(good, size=12, replace=False).tolist():
            want = np.zeros(ng)
            for a, b in ((r, c), (c, r)):                  # row i of the full symmetric matrix
                k = np.flatnonzero(a == i)
                j = b[k].astype(np.int64)
                keep = mask[j] == 0
                j, vv = j[keep], v[k][keep].astype(np.float64)
                lo, hi = np.minimum(i, j), np.maximum(i, j)
                want[pos[j]] = vv / exp_arr[hi - lo] / bias[lo] / bias[hi]
            assert np.array_equal(oe[pos[i]], want), i
        store.corr_build(0, n, bias, mask, exp_arr, stage=1)
        corr = store.corr_get()
        assert np.array_equal(corr, corr.T) and np.abs(corr).max() <= 1.0
        rows = np.sort(rng.choice(ng, size=256, replace=False))
        oe -= oe.mean(axis=1, keepdims=True)              # numpy.cov, in place
        sd = np.sqrt(np.einsum("ij,ij->i", oe, oe) / (ng - 1))
        cov = oe[rows] @ oe.T / (ng - 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            want = np.clip(cov / sd[rows, None] / sd[None, :], -1, 1)
        want = np.where(np.isnan(want), 0.0, want)
        np.testing.assert_allclose(corr[rows], want, rtol=0, atol=1e-12)
        del oe, cov, want
        vals, vecs, info = store.corr_eig(3)
        assert info["converged"]
        w, z = eigsh(corr, k=3)                            # hic_data.py:926-930
        np.testing.assert_allclose(vals, w[::-1], rtol=1e-10)
        for i in range(3):
            ref = z[:, -(i + 1)]
            assert min(np.abs(vecs[i] - ref).max(), np.abs(vecs[i] + ref).max()) < 1e-8, i
        t = store.corr_timing()
        print("c2 compartments: n %d, X X^T %.1f ms, eigen %.1f ms, %d Lanczos vectors" % (
            ng, t["gemm_ms"], t["eig_ms"], info["steps"]))
    finally:
        store.close()


def test_full_size_c3_against_oracle():
    """BASELINE configs[2], the bench workload: genome-wide 5 kb, 1.97e9 stored contacts."""
    wl, store, hic = _full_size_case("c3")
    try:
        assert store.info()["nnz_upper"] > 1.9e9
        lay = store.layout()
        assert lay["tile_cells"] > 1e9 and lay["encodings"]["D8"] > 1e6 and lay["csr_bytes"] == 0
        passes, dev, rel, nbad = _check_against_upper_port(wl, store, hic, 100, 1e-5)
        print("c3: passes %d final dev %.3g max rel bias err %.2e bads %d" % (passes, dev, rel, nbad))
    finally:
        store.close()


def test_batch_of_64_at_size_against_oracle():
    """BASELINE configs[4]: 64 synthetic chr1 @40 kb matrices (N = 6224) normalised as ONE batch;
    every matrix against the oracle on its exported cells."""
    import tadbit_b200
    from tadbit_b200 import workloads
    from tadbit_b200._capi import trim_memory
    trim_memory()
    size = workloads.HG38[0] // 40000 + 1
    nmat = 64
    hics = []
    for seed in range(1, nmat + 1):
        p = tadbit_b200.SynthParams(seed=seed, cis_scale=6.0e4, cis_alpha=1.08, trans_mean=0.0,
                                    vis_sigma=0.35, low_frac=0.02, low_scale=0.02, empty_frac=0.005)
        hics.append(tadbit_b200.HiC_data.from_store(tadbit_b200.ContactStore.synthetic(size, p)))
    for h in hics[::9]:
        h.filter_columns(silent=True)
    for h in hics:
        if not h.bads:
            h.filter_columns(silent=True, by_mean=False)
    tadbit_b200.normalize_hic_batch(hics, iterations=100, max_dev=1e-5, silent=True)
    worst = 0.0
    for k, h in enumerate(hics):
        r, c, v = h.store.export_upper()
        m = orc.UpperCells(size, r, c, v)
        if k % 9 == 0:
            assert h.bads == orc.fast_filter_columns(m, silent=True)
        bias, trace = orc.fast_normalize_hic(m, h.bads, 100, 1e-5)
        assert len(h.last_trace) == len(trace), k
        worst = max(worst, float(np.max(np.abs(h.bias.to_array() - bias) / np.abs(bias))))
    assert worst < 1e-10, worst
    for h in hics:
        h.store.close()
