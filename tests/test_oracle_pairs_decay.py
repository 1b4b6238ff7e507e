"""N1 / N2 on the CPU: the oracle's restatements and the product's host logic (around the test
double of the store) against vectors produced by the UNMODIFIED reference -- the Experiment class
for the pairs and z-scores, the reference's own lines of tools/tadbit_normalize.py for the decay
(oracle/gen_golden.py: reference_pair_cases, reference_decay_cases)."""
import io
import json
import os
from collections import OrderedDict

import numpy as np
import pytest

import tadbit_oracle as orc
from _golden import GOLDEN
from _numpy_store import NumpyStore

INDEX = json.load(open(os.path.join(GOLDEN, "INDEX.json")))
PAIR_CASES, DECAY_CASES = INDEX["pair_cases"], INDEX["decay_cases"]


def load_pairs(name):
    rec = json.load(open(os.path.join(GOLDEN, "pairs_%s.json" % name)))
    z = np.load(os.path.join(GOLDEN, "pairs_%s.npz" % name))
    return rec, z


def load_decay(name):
    rec = json.load(open(os.path.join(GOLDEN, "%s.json" % name)))
    z = np.load(os.path.join(GOLDEN, "%s.npz" % name))
    chroms = OrderedDict((k, int(v)) for k, v in rec["chromosomes"])
    nan = float("nan")
    rec["biases_in"] = np.array([nan if b is None else b for b in rec["biases_in"]])
    rec["biases_out"] = np.array([nan if b is None else b for b in rec["biases_out"]])
    rec["decay"] = dict((c, dict((int(k), v) for k, v in items)) for c, items in rec["decay"].items())
    return rec, z, chroms


def check_pairs(got, z, tag, rtol=1e-10):
    r, c, v = got
    assert np.array_equal(r, z[tag + "_i"]) and np.array_equal(c, z[tag + "_j"])
    want = z[tag + "_v"]
    assert np.array_equal(np.isnan(v), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_allclose(np.asarray(v, dtype=np.float64)[ok], want[ok], rtol=rtol, atol=1e-12)


@pytest.mark.parametrize("name", PAIR_CASES)
def test_oracle_pairs_against_the_reference(name):
    rec, z = load_pairs(name)
    m = orc.OracleMatrix(rec["size"], z["rows"], z["cols"], z["vals"])
    bads = dict((k, True) for k in rec["zeros"])
    bias = np.array(rec["bias"])
    check_pairs(orc.pair_values(m, bias, bads), z, "norm")
    check_pairs(orc.pair_zscores(m, bias, bads)[:3], z, "z")
    r, c, v = orc.pair_values(m, None, bads)                       # raw, zeros removed: a subset of "raw"
    full = dict(((int(a), int(b)), x) for a, b, x in zip(z["raw_i"], z["raw_j"], z["raw_v"]))
    assert all(full[(int(a), int(b))] == x for a, b, x in zip(r, c, v))
    assert sum(1 for x in full.values() if x) == r.size


def _experiment(rec, z):
    import tadbit_b200
    from tadbit_b200.experiment import Experiment
    store = NumpyStore(rec["size"], z["rows"], z["cols"], z["vals"])
    hic = tadbit_b200.HiC_data.from_store(store, dict_sec={})
    return Experiment("exp", 20000, hic_data=hic)


@pytest.mark.parametrize("name", PAIR_CASES)
def test_experiment_host_layer_against_the_reference(name):
    """tadbit_b200.Experiment around the store double: every get_hic_zscores variant and the lines
    write_interaction_pairs produces."""
    rec, z = load_pairs(name)
    exp = _experiment(rec, z)
    exp.filter_columns(silent=True)
    assert sorted(exp._zeros) == rec["zeros"]
    exp.normalize_hic(silent=True, **rec["normalize"])
    np.testing.assert_allclose(exp.hic_data[0].bias.to_array(), np.array(rec["bias"]), rtol=1e-12)
    for tag, opts in (("z", dict()), ("norm", dict(zscored=False)),
                      ("raw", dict(normalized=False, zscored=False)), ("rawz", dict(normalized=False)),
                      ("z_keep0", dict(remove_zeros=False)),
                      ("norm_keep0", dict(zscored=False, remove_zeros=False))):
        exp.get_hic_zscores(**opts)
        flat = sorted((int(a), int(b), v) for a, d in exp._zscores.items() for b, v in d.items())
        got = (np.array([f[0] for f in flat], dtype=np.int32), np.array([f[1] for f in flat], dtype=np.int32),
               np.array([f[2] for f in flat], dtype=np.float64))
        check_pairs(got, z, tag)
    exp.get_hic_zscores(zscored=False)
    out = io.StringIO()
    out.close = lambda: None
    exp.write_interaction_pairs(out)
    lines = out.getvalue().splitlines()
    assert len(lines) == len(rec["pair_lines"])
    for a, b in zip(lines, rec["pair_lines"]):
        fa, fb = a.split("\t"), b.split("\t")
        assert fa[:2] == fb[:2] and abs(float(fa[2]) - float(fb[2])) <= 1e-10 * abs(float(fb[2]))


@pytest.mark.parametrize("name", DECAY_CASES)
def test_oracle_decay_against_the_reference_lines(name):
    rec, z, chroms = load_decay(name)
    m = orc.OracleMatrix(rec["size"], z["rows"], z["cols"], z["vals"], chromosomes=chroms)
    badcol = dict((k, True) for k in rec["badcol"])
    b, decay, cis = orc.rescale_and_decay(m, rec["biases_in"], badcol, factor=rec["factor"])
    ok = ~np.isnan(rec["biases_out"])
    assert np.array_equal(np.isnan(b), ~ok)
    np.testing.assert_allclose(b[ok], rec["biases_out"][ok], rtol=1e-12)
    assert abs(cis - rec["norm_cisprc"]) <= 1e-12 * rec["norm_cisprc"]
    for crm in chroms:
        assert sorted(decay[crm]) == sorted(rec["decay"][crm]), crm
        for k, v in rec["decay"][crm].items():
            assert abs(decay[crm][k] - v) <= 1e-11 * abs(v) + 1e-300, (crm, k)


@pytest.mark.parametrize("name", DECAY_CASES)
def test_decay_host_layer_against_the_reference_lines(name):
    """tadbit_b200.decay.rescale_and_decay around the store double (closed-form cell counts, pooling)."""
    import tadbit_b200
    from tadbit_b200.decay import rescale_and_decay
    rec, z, chroms = load_decay(name)
    store = NumpyStore(rec["size"], z["rows"], z["cols"], z["vals"], chromosomes=chroms)
    hic = tadbit_b200.HiC_data.from_store(store, chromosomes=chroms)
    badcol = dict((k, True) for k in rec["badcol"])
    biases, decay, bad_out, cis = rescale_and_decay(hic, dict(enumerate(rec["biases_in"].tolist())), badcol,
                                                    factor=rec["factor"])
    got = np.array([biases[k] for k in range(rec["size"])])
    ok = ~np.isnan(rec["biases_out"])
    assert np.array_equal(np.isnan(got), ~ok) and bad_out is badcol
    np.testing.assert_allclose(got[ok], rec["biases_out"][ok], rtol=1e-12)
    assert abs(cis - rec["norm_cisprc"]) <= 1e-12 * rec["norm_cisprc"]
    for crm in chroms:
        assert sorted(decay[crm]) == sorted(rec["decay"][crm]), crm
        for k, v in rec["decay"][crm].items():
            assert abs(decay[crm][k] - v) <= 1e-11 * abs(v) + 1e-300, (crm, k)


def test_closed_form_cell_counts_equal_the_literal_loop():
    from tadbit_b200.decay import diagonal_cell_counts
    rng = np.random.default_rng(9)
    for trial in range(30):
        sizes = [int(v) for v in rng.integers(1, 60, size=int(rng.integers(1, 5)))]
        pos, total = OrderedDict(), 0
        for k, sz in enumerate(sizes):
            pos["c%d" % k] = (total, total + sz)
            total += sz
        nbad = int(rng.integers(0, total + 1))
        badcol = dict((int(b), True) for b in rng.choice(total, size=nbad, replace=False))
        want = orc.decay_cell_counts(pos, badcol)
        got = diagonal_cell_counts(pos, badcol)
        for crm in pos:
            assert [int(v) for v in got[crm]] == [want[crm][k] for k in range(len(got[crm]))], (sizes, sorted(badcol))


def test_native_decay_pooling_equals_the_python_loop():
    """decay.pool_decay runs the walk in the library (tb_pool_decay, host C++); _pool_decay_loop is
    the reference's walk in Python, distance by distance.  Same dictionaries -- keys, their order,
    and the bits of the values -- on random chromosomes, incl. empty diagonals, diagonals without a
    good cell, thresholds an integer sum can hit exactly, negative counts."""
    from tadbit_b200 import decay as d
    rng = np.random.default_rng(2)
    for it in range(3000):
        size = int(rng.integers(0, 60)) if it % 50 else int(rng.integers(2000, 6000))
        scale = rng.choice([0, 1, 20, 300, 500, 5000, 1e5])
        raw = (scale * rng.random(size) * (np.arange(size) + 1.0) ** -rng.choice([0, 0.5, 1.5])).astype(np.int64)
        if rng.random() < 0.4:
            raw[rng.random(size) < 0.4] = 0
        if rng.random() < 0.05 and size:
            raw[rng.integers(0, size)] *= -1
        nrm = np.where(raw != 0, rng.random(size) * raw * rng.lognormal(0, 1, size), 0.0)
        if rng.random() < 0.1:
            nrm = rng.random(size)
        nd = np.maximum(size - np.arange(size) - rng.integers(0, 5, size), 0)
        if rng.random() < 0.2 and size:
            nd[rng.random(size) < 0.3] = 0
        min_n = float(rng.choice([0.05, 0.1, 0.5, 0.02, 0.25])) ** -2.
        got, want = d.pool_decay(nrm, raw, nd, min_n), d._pool_decay_loop(nrm, raw, nd, min_n)
        assert list(got.items()) == list(want.items()), (it, size, min_n)


def test_native_expected_pooling_equals_the_python_loop():
    """tb_pool_expected against the Python restatement of the loop around _meandiag
    (normalize_hic.py:262-291): same keys in the same order, same bits."""
    import ctypes as C
    from tadbit_b200 import _capi, normalize_hic as nh
    rng = np.random.default_rng(8)
    for it in range(3000):
        size = int(rng.integers(0, 70)) if it % 60 else int(rng.integers(3000, 9000))
        ndiag = size if rng.random() < 0.8 else int(rng.integers(0, size + 1))
        scale = rng.choice([0, 1, 50, 500, 5000, 1e7])
        sums = (scale * rng.random(ndiag) * (np.arange(ndiag) + 1.0) ** -rng.choice([0, 1, 2])).astype(np.int64)
        cells = np.maximum(ndiag - np.arange(ndiag) - rng.integers(0, 4, ndiag), 0).astype(np.int64)
        if rng.random() < 0.3 and ndiag:
            cells[rng.random(ndiag) < 0.3] = 0
        min_n = float(rng.choice([0.05, 0.1, 0.5, 0.25])) ** -2.
        vals = np.empty(size + 2, dtype=np.float64)
        nkeys = C.c_int64()
        _capi.check(_capi.lib().tb_pool_expected(size, ndiag, _capi.ptr(sums, C.c_int64),
                                                 _capi.ptr(cells, C.c_int64), min_n,
                                                 _capi.ptr(vals, C.c_double), C.byref(nkeys)))
        got = dict(zip(range(nkeys.value), vals[:nkeys.value].tolist()))
        want = nh._pool_expected_loop(sums.tolist(), cells.tolist(), size, ndiag, min_n)
        assert list(got.items()) == list(want.items()), (it, size, ndiag, min_n)
