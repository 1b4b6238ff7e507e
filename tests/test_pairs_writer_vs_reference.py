"""Experiment.write_interaction_pairs against the UNMODIFIED reference, live, over its options
(zscored / normalized / raw, cutoff, uniq, diagonal, remove_zeros, focus, true_position, tsv / json,
header) on seeded random matrices -- the array path (PairTable + tb_write_pair_lines) and the
pair-by-pair loop that remains for the dense variants.  Runs where the reference tree exists."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

from _numpy_store import NumpyStore

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def _matrix(seed):
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(25, 70))
    vis = rng.lognormal(0.0, 0.4, n)
    vis[rng.random(n) < 0.1] *= 0.02
    vis[rng.random(n) < 0.05] = 0.0
    i, j = np.triu_indices(n)
    v = rng.poisson(40.0 * (j - i + 1.0) ** -1.0 * vis[i] * vis[j])
    keep = v > 0
    return rng, n, i[keep], j[keep], v[keep].astype(np.int64)


def _reference_experiment(n, r, c, v):
    from gen_golden import build_reference
    ns = ref_shim.load()
    ref_shim._stub("pytadbit.modelling.structuralmodels", StructuralModels=object)
    from pytadbit.experiment import Experiment
    exp = Experiment("ref", 10000)
    exp.hic_data = [build_reference(ns, n, r, c, v)]
    exp.size = n
    return exp


def _ours(n, r, c, v):
    import tadbit_b200
    from tadbit_b200.experiment import Experiment
    hic = tadbit_b200.HiC_data.from_store(NumpyStore(n, r, c, v), dict_sec={})
    return Experiment("new", 10000, hic_data=hic)


@pytest.mark.parametrize("seed", range(12))
def test_written_pairs_against_the_live_reference(seed, tmp_path):
    rng, n, r, c, v = _matrix(seed)
    ref, new = _reference_experiment(n, r, c, v), _ours(n, r, c, v)
    quiet = lambda: contextlib.ExitStack()                            # noqa: E731
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()), \
            np.errstate(all="ignore"):
        import types
        for exp in (ref, new):
            exp.crm = types.SimpleNamespace(name="chr7")
            exp.cell_type, exp.identifier = "K562", "SRX000001"
            exp.description.update(species="Homo sapiens", assembly="hg38")
            exp.filter_columns(silent=True)
            exp.normalize_hic(silent=True, iterations=10, max_dev=1e-3)
        assert sorted(ref._zeros) == sorted(new._zeros)
        lo = int(rng.integers(0, n // 2))
        focus = (lo, int(rng.integers(lo + 3, n)))
        variants = [
            dict(),
            dict(cutoff=0.5), dict(cutoff=-0.3, uniq=False), dict(focus=focus), dict(focus=focus, true_position=True),
            dict(header=True, diagonal=True), dict(format="json"), dict(format="json", true_position=True),
            dict(remove_zeros=True, cutoff=1.0),
            dict(zscored=False), dict(zscored=False, remove_zeros=True), dict(zscored=False, remove_zeros=True, focus=focus),
            dict(zscored=False, remove_zeros=True, uniq=False), dict(zscored=False, remove_zeros=True, diagonal=True),
            dict(zscored=False, normalized=False), dict(zscored=False, normalized=False, remove_zeros=True, header=True),
            dict(zscored=False, normalized=False, remove_zeros=True, true_position=True, format="json"),
            dict(format="json", header=True), dict(format="json", header=True, focus=focus, cutoff=0.2),
        ]
        for zkw in (dict(), dict(zscored=False), dict(remove_zeros=False), dict(normalized=False)):
            for exp in (ref, new):
                exp.get_hic_zscores(**zkw)
            for k, kw in enumerate(variants):
                a, b = str(tmp_path / "ref.txt"), str(tmp_path / "new.txt")
                ref.write_interaction_pairs(a, **kw)
                new.write_interaction_pairs(b, **kw)
                ta, tb = open(a).read().splitlines(), open(b).read().splitlines()
                assert len(ta) == len(tb), (seed, zkw, kw)
                for x, y in zip(ta, tb):
                    if x == y:
                        continue
                    sep = "," if kw.get("format") == "json" else "\t"
                    fx, fy = x.rstrip(",").split(sep), y.rstrip(",").split(sep)
                    assert fx[:2] == fy[:2] and x.endswith(",") == y.endswith(","), (seed, zkw, kw, x, y)
                    assert abs(float(fx[2]) - float(fy[2])) <= 1e-9 * abs(float(fx[2])), (seed, zkw, kw, x, y)
    # a file object instead of a path
    out = io.StringIO()
    out.close = lambda: None
    new.get_hic_zscores()
    new.write_interaction_pairs(out, cutoff=0.0)
    new.write_interaction_pairs(str(tmp_path / "p.tsv"), cutoff=0.0)
    assert out.getvalue() == open(str(tmp_path / "p.tsv")).read()


@pytest.mark.parametrize("seed", range(6))
def test_dense_matrix_against_the_live_reference(seed):
    """Experiment.get_hic_matrix / print_hic_matrix (experiment.py:1192-1255) over tb_window."""
    rng, n, r, c, v = _matrix(100 + seed)
    ref, new = _reference_experiment(n, r, c, v), _ours(n, r, c, v)
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()), \
            np.errstate(all="ignore"):
        for exp in (ref, new):
            with pytest.raises(Exception, match="not normalized"):
                exp.get_hic_matrix(normalized=True)
            exp.filter_columns(silent=True)
            exp.normalize_hic(silent=True, iterations=8, max_dev=1e-3)
    hi = int(rng.integers(5, n))
    for kw in (dict(), dict(diagonal=False), dict(focus=(1, hi)), dict(focus=(1, hi), diagonal=False),
               dict(focus=(3, hi)), dict(normalized=True), dict(normalized=True, focus=(2, hi)),
               dict(normalized=True, diagonal=False)):
        a, b = ref.get_hic_matrix(**kw), new.get_hic_matrix(**kw)
        assert len(a) == len(b) and all(len(x) == len(y) for x, y in zip(a, b)), kw
        if kw.get("normalized"):
            np.testing.assert_allclose(np.array(b, dtype=float), np.array(a, dtype=float), rtol=1e-9, atol=0)
        else:
            assert a == b and all(isinstance(x, int) for x in b[0]), kw
    for kw in (dict(focus=(4, hi), diagonal=False),):               # absolute indices into a window
        outcome = []
        for exp in (ref, new):
            try:
                outcome.append(exp.get_hic_matrix(**kw))
            except IndexError as e:
                outcome.append(str(e))
        assert outcome[0] == outcome[1]
    for kw in (dict(), dict(zeros=True)):
        assert ref.print_hic_matrix(print_it=False, **kw) == new.print_hic_matrix(print_it=False, **kw)
    ta = ref.print_hic_matrix(print_it=False, normalized=True, zeros=True).split()
    tb = new.print_hic_matrix(print_it=False, normalized=True, zeros=True).split()
    assert len(ta) == len(tb)
    for x, y in zip(ta, tb):
        assert x == y or abs(float(x) - float(y)) <= 1e-9 * abs(float(x))


def test_pair_table_is_the_dictionary_of_dictionaries():
    from tadbit_b200.vectors import PairTable
    rows = np.array([0, 0, 2, 2, 2, 7])
    cols = np.array([1, 5, 3, 4, 9, 8])
    vals = np.array([0.5, -1.0, 2.0, 3.5, 0.0, 9.0])
    t = PairTable(rows, cols, vals)
    want = {"0": {"1": 0.5, "5": -1.0}, "2": {"3": 2.0, "4": 3.5, "9": 0.0}, "7": {"8": 9.0}}
    assert t == want and dict(t) == want and t.to_dict() == want
    assert list(t) == ["0", "2", "7"] and len(t) == 3 and bool(t)
    assert "2" in t and "1" not in t and 2 not in t and "02" not in t and "x" not in t
    assert t["2"]["4"] == 3.5 and t.get("3") is None
    with pytest.raises(KeyError):
        t["1"]
    with pytest.raises(KeyError):
        t[2]
    assert not PairTable([], [], []) and len(PairTable([], [], [])) == 0
    assert PairTable(rows, cols, np.arange(6))["7"] == {"8": 5}
    with pytest.raises(ValueError):
        PairTable([3, 1], [4, 2], [1.0, 2.0])
