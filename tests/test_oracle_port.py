"""The oracle's C + OpenMP pass loop (oracle/ice_port.c) against the reference's golden
vectors and the numpy restatement.  Tolerance 1e-12 relative (summation order differs)."""
import numpy as np
import pytest

import tadbit_oracle as orc
from _golden import CASES, load


@pytest.mark.parametrize("name", CASES)
def test_port_matches_reference_golden(name):
    case = load(name)
    m = orc.OracleMatrix(case.size, case.rows, case.cols, case.vals,
                         chromosomes=case.chromosomes)
    bads = case.bads(case.out["filter_columns"])
    for rec in case.out["ice"]:
        if rec["sqrt"] or rec["factor"]:
            continue                      # the port covers iterative() only
        if "raises" in rec:
            with pytest.raises(ZeroDivisionError):
                orc.iterative_fast(m, bads, rec["iterations"], rec["max_dev"], 2)
            continue
        bias, trace, cores, _ = orc.iterative_fast(m, bads, rec["iterations"], rec["max_dev"], 2)
        np.testing.assert_allclose(bias, np.array(rec["bias"]), rtol=1e-12, atol=0)
        assert len(trace) == len(case.trace_lines(rec["stdout"]))


def test_port_matches_numpy_oracle_and_uses_c():
    case = load("synth_1chrom_400")
    m = orc.OracleMatrix(case.size, case.rows, case.cols, case.vals)
    assert orc._load_port() is not None, "oracle/_build/libice_port.so did not build"
    b1, t1 = orc.iterative(m, {3: True, 7: True}, 30, 1e-7)
    b2, t2, cores, dt = orc.iterative_fast(m, {3: True, 7: True}, 30, 1e-7, 3)
    assert cores == 3 and len(t1) == len(t2)
    np.testing.assert_allclose(b2, b1, rtol=1e-13, atol=0)


def test_cpu_synthetic_block_is_a_valid_upper_triangle():
    """oracle/synth_port.c (input of bench.py's CPU legs): sorted unique upper-triangle cells,
    reproducible, empty bins stay empty, and a block is a prefix-consistent cut of a larger one."""
    import types
    from collections import OrderedDict
    p = types.SimpleNamespace(seed=20261019, cis_scale=600.0, cis_alpha=1.08, trans_mean=0.02,
                              vis_sigma=0.35, low_frac=0.03, low_scale=0.02, empty_frac=0.01)
    chroms = OrderedDict([("chrA", 700), ("chrB", 500)])
    r, c, v = orc.synthetic_block(1200, p, chroms)
    assert r.size > 10000 and r.dtype == np.int32
    assert np.all(r <= c) and np.all(v > 0) and c.max() < 1200
    key = r.astype(np.int64) * 1200 + c
    assert np.all(np.diff(key) > 0)                       # row-major sorted, no duplicates
    r2, c2, v2 = orc.synthetic_block(1200, p, chroms)
    assert np.array_equal(r, r2) and np.array_equal(c, c2) and np.array_equal(v, v2)
    # cells are a function of (seed, i, j) only: the 600-bin block is the corner of the 1200-bin one
    r3, c3, v3 = orc.synthetic_block(600, p, chroms)
    corner = c < 600
    assert np.array_equal(r[corner], r3) and np.array_equal(c[corner], c3) and np.array_equal(v[corner], v3)
    # cis contacts decay with distance, trans contacts are rarer than near-diagonal ones
    near = (c - r) < 20
    trans = (r < 700) & (c >= 700)
    assert v[near].mean() > 5 * v[trans].mean()
    # an OracleMatrix built from it balances
    m = orc.OracleMatrix(1200, r, c, v, chromosomes=chroms)
    bads = orc.filter_columns(m, silent=True)
    bias, trace, _ = orc.normalize_hic(m, bads=bads, iterations=50, max_dev=1e-4)
    assert np.isfinite(bias).all() and len(trace) > 3
