"""GPU parity against the reference's own outputs (committed golden vectors) and the
oracle.  Everything goes through the public HiC_data API, i.e. through the C-ABI.

Bars (north_star): filter mask / indices / integer sums bit-exact; expected bit-exact
(integer sums + one fp64 divide); biases 1e-10 relative (we assert 1e-12)."""
import io
import contextlib

import re

import numpy as np
import pytest

from _golden import CASES, load

pytestmark = pytest.mark.gpu

TRACE_FMT = '   %15.3f %15.3f %15.3f %4s %9.5f'


def _hic(case):
    import tadbit_b200
    return tadbit_b200.HiC_data.from_coo(case.size, case.rows, case.cols, case.vals,
                                         chromosomes=case.chromosomes,
                                         dict_sec=None if case.chromosomes else {},
                                         symmetricized=case.symmetricized)


@pytest.fixture(scope="module")
def hics():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = _hic(load(name))
        return cache[name]
    return get


@pytest.mark.parametrize("name", CASES)
def test_zero_count_filter(name, hics):
    import tadbit_b200
    case, hic = load(name), hics(name)
    assert tadbit_b200.filter_by_zero_count(hic, 99) == case.bads(case.out["zero_count_99"])
    assert tadbit_b200.filter_by_zero_count(hic, 75) == case.bads(case.out["zero_count_75"])
    mc = case.out["min_count"]
    with contextlib.redirect_stderr(io.StringIO()):
        got = tadbit_b200.filter_by_zero_count(hic, 99, min_count=mc["min_count"])
    assert got == case.bads(mc["bads"])


@pytest.mark.parametrize("name", CASES)
def test_filter_columns(name, hics):
    case, hic = load(name), hics(name)
    hic.bads = {}
    hic.filter_columns(silent=True, **case.out["filter_kw"])
    want = case.bads(case.out["filter_columns"])
    assert hic.bads == want
    for k in want:
        assert (hic.bads[k] is True) == (want[k] is True)


@pytest.mark.parametrize("name", CASES)
def test_normalize_hic(name, hics):
    case, hic = load(name), hics(name)
    for rec in case.out["ice"]:
        hic.bads = case.bads(case.out["filter_columns"])
        kw = dict(iterations=rec["iterations"], max_dev=rec["max_dev"],
                  sqrt=rec["sqrt"], factor=rec["factor"])
        if "raises" in rec:
            with pytest.raises(ZeroDivisionError, match=re.escape(rec["message"])):
                hic.normalize_hic(silent=True, **kw)
            continue
        out = io.StringIO()
        with contextlib.redirect_stdout(out):
            hic.normalize_hic(silent=False, **kw)
        assert len(hic.bias) == case.size
        got = np.array([hic.bias[i] for i in range(case.size)])
        np.testing.assert_allclose(got, np.array(rec["bias"]), rtol=1e-12, atol=0)
        # the verbose trace is observable output (normalize_hic.py:219-220)
        assert case.trace_lines(out.getvalue()) == case.trace_lines(rec["stdout"])
        assert out.getvalue().splitlines()[:3] == rec["stdout"].splitlines()[:3]


@pytest.mark.parametrize("name", CASES)
def test_sum_and_expected(name, hics):
    case, hic = load(name), hics(name)
    hic.bads = case.bads(case.out["filter_columns"])
    assert hic.sum() == case.out["raw_sum"]
    for rec in case.out["expected"]:
        hic.normalize_expected(**rec["kw"])
        want = dict((d, v) for d, v in rec["values"])
        assert hic.expected == want


def test_reference_test_values(hics):
    """chrT 20Kb A: mask {22: 48} and the bias behind test/test_all.py:400-402."""
    case, hic = load("chrT_20Kb_A"), hics("chrT_20Kb_A")
    hic.filter_columns(silent=True)
    assert hic.bads == {22: 48}
    hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
    assert abs(hic.bias[0] - 3.977595148485566) < 1e-12 * 3.98
    hic.normalize_expected()
    assert hic.expected[0] == 378.2323232323232
    assert hic.expected[1] == 120.42857142857143
    assert len(hic.expected) == 102


# ---- the reference's own known-answer tests (test/test_all.py test_09, test_11) through the product
def test_reference_test_09_hic_normalization(hics):
    import _known_answers as ka
    case, hic = load(ka.KNOWN["fixture"]), hics(ka.KNOWN["fixture"])
    want = ka.KNOWN["test_09"]
    hic.bads = {}
    hic.normalize_hic(silent=True)                                 # defaults, as Experiment.normalize_hic
    bias = hic.bias.to_array()
    np.testing.assert_allclose(bias, np.array(want["bias"]), rtol=1e-12, atol=0)
    lines = ka.pair_lines(case, bias, {})
    assert len(lines) == want["pairs"]
    total = sum(v for _, _, v in lines)
    assert abs(total - want["sum_normalised"]) < 1e-8
    assert round(total, 4) == round(ka.TEST_09_SUM, 4)            # reference test/test_all.py:332


def test_reference_test_11_write_interaction_pairs(hics):
    import _known_answers as ka
    case, hic = load(ka.KNOWN["fixture"]), hics(ka.KNOWN["fixture"])
    want = ka.KNOWN["test_11"]
    hic.bads = {}
    hic.filter_columns(silent=True)
    assert sorted(hic.bads) == want["zeros"]
    hic.normalize_hic(factor=1, silent=True)
    bias = hic.bias.to_array()
    np.testing.assert_allclose(bias, np.array(want["bias"]), rtol=1e-12, atol=0)
    lines = ka.pair_lines(case, bias, hic.bads)
    assert len(lines) == ka.TEST_11_LINES                         # reference test/test_all.py:400
    assert tuple(k + 1 for k in lines[25][:2]) == tuple(int(x) for x in want["line_25"][:2])
    assert abs(lines[25][2] - ka.TEST_11_LINE_25) < 1e-7          # assertAlmostEqual, :401
    assert abs(lines[2000][2] - ka.TEST_11_LINE_2000) < 1e-7      # :402
    assert abs(lines[25][2] - float(want["line_25"][2])) < 1e-12
    assert abs(lines[2000][2] - float(want["line_2000"][2])) < 1e-12
