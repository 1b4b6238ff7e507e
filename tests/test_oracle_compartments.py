"""N4 (compartments): the oracle's restatement (oracle/tadbit_oracle.py: obs_exp_matrix,
correlation_matrix, largest_eigenpairs, compartment_breaks) pinned to what the UNMODIFIED
reference's find_compartments produced (tests/golden/compartments_*, made by oracle/gen_golden.py
with numpy.corrcoef / scipy eigsh wrapped to record their inputs)."""
import json
import os
from collections import OrderedDict

import numpy as np
import pytest

import tadbit_oracle as orc
from _golden import GOLDEN

CASES = json.load(open(os.path.join(GOLDEN, "INDEX.json")))["compartment_cases"]


def load_compartments(name):
    rec = json.load(open(os.path.join(GOLDEN, "compartments_%s.json" % name)))
    z = np.load(os.path.join(GOLDEN, "compartments_%s.npz" % name))
    chroms = OrderedDict((k, int(v)) for k, v in rec["chromosomes"]) if rec["chromosomes"] else None
    bads = dict((int(k), v) for k, v in rec["bads"].items())
    expected = dict((int(k), v) for k, v in rec["expected"].items())
    return rec, z, chroms, bads, expected


def section_bounds(rec, chroms):
    """name -> (begin, end), as HiC_data.section_pos; one unnamed section without chromosomes."""
    if not chroms:
        return OrderedDict([("None", (0, rec["n"]))])
    out, tot = OrderedDict(), 0
    for name, size in chroms.items():
        out[name] = (tot, tot + size)
        tot += size
    return out


def with_bad_bins(vec, beg, end, bads):
    """bad bins back in as NaN (hic_data.py:949-953, inclusive upper bound)."""
    vec = list(vec)
    for b in [k - beg for k in sorted(bads) if beg <= k <= end]:
        vec.insert(b, float("nan"))
    return np.array(vec)


def same_up_to_sign(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    assert min(np.abs(a[ok] - b[ok]).max(), np.abs(a[ok] + b[ok]).max()) < tol


def shifted_breaks(vector, beg, end, bads):
    breaks = orc.compartment_breaks(vector)
    for b in [k - beg for k in sorted(bads) if beg <= k <= end]:
        for brk in breaks:
            if brk["start"] >= b:
                brk["start"] += 1
                brk["end"] += 1
            else:
                brk["end"] += brk["end"] > b
    return breaks


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_reference(name):
    rec, z, chroms, bads, expected = load_compartments(name)
    m = orc.OracleMatrix(rec["n"], z["rows"], z["cols"], z["vals"], chromosomes=chroms)
    bounds = section_bounds(rec, chroms)
    for k, sec in enumerate(rec["sections"]):
        beg, end = bounds[sec]
        exp_arr = np.array([expected.get(d, np.nan) for d in range(end - beg)])
        oe = orc.obs_exp_matrix(m, beg, end, z["bias"], bads, exp_arr)
        assert np.array_equal(oe, z["obsexp_%d" % k])                 # same divisions, same order: same bits
        corr = orc.correlation_matrix(oe)
        np.testing.assert_allclose(corr, z["corr_%d" % k], rtol=0, atol=1e-14)
        want_vals, want_vecs = z["evals_%d" % k], z["evecs_%d" % k]
        kk = rec["call"]["max_ev"]
        vals, vecs = orc.largest_eigenpairs(corr, kk)
        np.testing.assert_allclose(vals, want_vals, rtol=1e-10, atol=1e-12)
        for i in range(len(want_vecs)):
            if abs(want_vals[i]) < 1e-9 * abs(want_vals[0]):
                continue
            same_up_to_sign(with_bad_bins(vecs[i], beg, end, bads), want_vecs[i], 1e-8)
        assert shifted_breaks(vecs[0], beg, end, bads) == rec["compartments"][sec]
    for sec, cm in rec["compartments"].items():
        if sec not in rec["sections"]:
            assert cm == []
