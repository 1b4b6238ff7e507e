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
