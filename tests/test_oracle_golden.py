"""The oracle (numpy restatement) against the reference's own outputs.

Golden vectors were produced by the unmodified reference (oracle/gen_golden.py);
this pins the oracle that the GPU parity tests then rely on.  Tolerances: masks,
indices, integer sums and expected values bit-exact; biases 1e-12 relative (the
oracle sums in a different order than the reference's dict loops)."""
import re

import numpy as np
import pytest

import tadbit_oracle as orc
from _golden import CASES, load

TRACE_FMT = '   %15.3f %15.3f %15.3f %4s %9.5f'


def _matrix(case):
    return orc.OracleMatrix(case.size, case.rows, case.cols, case.vals,
                            chromosomes=case.chromosomes,
                            symmetricized=case.symmetricized)


@pytest.mark.parametrize("name", CASES)
def test_zero_count_filter(name):
    case = load(name)
    m = _matrix(case)
    assert orc.filter_by_zero_count(m, 99) == case.bads(case.out["zero_count_99"])
    assert orc.filter_by_zero_count(m, 75) == case.bads(case.out["zero_count_75"])
    mc = case.out["min_count"]
    assert orc.filter_by_zero_count(m, 99, min_count=mc["min_count"]) == case.bads(mc["bads"])


@pytest.mark.parametrize("name", CASES)
def test_filter_columns(name):
    case = load(name)
    got = orc.filter_columns(_matrix(case), silent=True, **case.out["filter_kw"])
    want = case.bads(case.out["filter_columns"])
    assert got == want
    for k in want:   # True vs integer column sum must match too (hic_filtering.py:145)
        assert (got[k] is True) == (want[k] is True)


@pytest.mark.parametrize("name", CASES)
def test_ice(name):
    case = load(name)
    m = _matrix(case)
    bads = case.bads(case.out["filter_columns"])
    for rec in case.out["ice"]:
        kw = dict(iterations=rec["iterations"], max_dev=rec["max_dev"],
                  sqrt=rec["sqrt"], factor=rec["factor"])
        if "raises" in rec:
            with pytest.raises(ZeroDivisionError, match=re.escape(rec["message"])):
                orc.normalize_hic(m, bads=bads, **kw)
            continue
        bias, trace, _ = orc.normalize_hic(m, bads=bads, **kw)
        np.testing.assert_allclose(bias, np.array(rec["bias"]), rtol=1e-12, atol=0)
        want_lines = case.trace_lines(rec["stdout"])
        got_lines = [TRACE_FMT % t for t in trace]
        assert len(got_lines) == len(want_lines)
        assert got_lines == want_lines


@pytest.mark.parametrize("name", CASES)
def test_raw_sum_and_expected(name):
    case = load(name)
    m = _matrix(case)
    bads = case.bads(case.out["filter_columns"])
    assert int(orc.matrix_sum(m, None, bads)) == case.out["raw_sum"]
    for rec in case.out["expected"]:
        got = orc.expected(m, bads=bads, **rec["kw"])
        want = dict((d, v) for d, v in rec["values"])
        assert sorted(got) == sorted(want)
        assert got == want          # integer sums + one fp64 divide: bit-exact


def test_survey_pinned_values():
    """Values quoted in SURVEY.md 8(c) for chrT 20Kb A (mask behind test_all.py:400-402)."""
    case = load("chrT_20Kb_A")
    assert case.bads(case.out["filter_columns"]) == {22: 48}
    rec = case.out["ice"][1]
    assert rec["bias"][0] == 3.977595148485566
    exp = dict(case.out["expected"][0]["values"])
    assert exp[0] == 378.2323232323232 and exp[1] == 120.42857142857143
    assert case.out["expected"][0]["keys"] == 102


# ---- the reference's own known-answer tests (test/test_all.py test_09, test_11) ---------------
def test_reference_test_09_hic_normalization():
    import _known_answers as ka
    case = load(ka.KNOWN["fixture"])
    m = orc.OracleMatrix(case.size, case.rows, case.cols, case.vals, chromosomes=case.chromosomes)
    want = ka.KNOWN["test_09"]
    assert want["zeros"] == []
    bias, _, _ = orc.normalize_hic(m, bads={}, iterations=0, max_dev=0.1, factor=1)
    np.testing.assert_allclose(bias, np.array(want["bias"]), rtol=1e-12, atol=0)
    lines = ka.pair_lines(case, bias, {})
    assert len(lines) == want["pairs"]
    total = sum(v for _, _, v in lines)
    assert abs(total - want["sum_normalised"]) < 1e-8
    assert round(total, 4) == round(ka.TEST_09_SUM, 4)            # reference test/test_all.py:332


def test_reference_test_11_write_interaction_pairs():
    import _known_answers as ka
    case = load(ka.KNOWN["fixture"])
    m = orc.OracleMatrix(case.size, case.rows, case.cols, case.vals, chromosomes=case.chromosomes)
    want = ka.KNOWN["test_11"]
    bads = orc.filter_columns(m, silent=True)
    assert sorted(bads) == want["zeros"]
    bias, _, _ = orc.normalize_hic(m, bads=bads, iterations=0, max_dev=0.1, factor=1)
    np.testing.assert_allclose(bias, np.array(want["bias"]), rtol=1e-12, atol=0)
    lines = ka.pair_lines(case, bias, bads)
    assert len(lines) == ka.TEST_11_LINES                         # reference test/test_all.py:400
    # the pair file numbers bins from 1 (experiment.py write_interaction_pairs)
    assert tuple(k + 1 for k in lines[25][:2]) == tuple(int(x) for x in want["line_25"][:2])
    assert tuple(k + 1 for k in lines[2000][:2]) == tuple(int(x) for x in want["line_2000"][:2])
    assert abs(lines[25][2] - ka.TEST_11_LINE_25) < 1e-7          # assertAlmostEqual, :401
    assert abs(lines[2000][2] - ka.TEST_11_LINE_2000) < 1e-7      # :402
    assert abs(lines[25][2] - float(want["line_25"][2])) < 1e-12
    assert abs(lines[2000][2] - float(want["line_2000"][2])) < 1e-12
