"""Batch ICE (BASELINE configs[4]) == the one-by-one loop, per matrix, against the reference's
golden biases and against the single-matrix GPU path."""
import contextlib
import io

import numpy as np
import pytest

from _golden import CASES, load

pytestmark = pytest.mark.gpu


def _hic(case):
    import tadbit_b200
    return tadbit_b200.HiC_data.from_coo(case.size, case.rows, case.cols, case.vals,
                                         chromosomes=case.chromosomes,
                                         dict_sec=None if case.chromosomes else {})


def test_batch_matches_reference_golden():
    import tadbit_b200
    cases = [load(n) for n in CASES]
    cases = [c for c in cases if not any("raises" in rec for rec in c.out["ice"])]
    for idx in (1, 3):                                  # (100, 1e-5, factor 1) and (100, 1e-6, None)
        hics = [_hic(c) for c in cases]
        recs = []
        for h, c in zip(hics, cases):
            h.bads = c.bads(c.out["filter_columns"])
            rec = c.out["ice"][min(idx, len(c.out["ice"]) - 1)]
            recs.append(rec)
        groups = {}
        for h, c, rec in zip(hics, cases, recs):        # batch the matrices that share settings
            key = (rec["iterations"], rec["max_dev"], rec["sqrt"], rec["factor"])
            groups.setdefault(key, []).append((h, c, rec))
        for (iterations, max_dev, sqrt, factor), members in groups.items():
            out = io.StringIO()
            with contextlib.redirect_stdout(out):
                tadbit_b200.normalize_hic_batch([m[0] for m in members], iterations=iterations,
                                                max_dev=max_dev, silent=False, sqrt=sqrt, factor=factor)
            for h, c, rec in members:
                got = h.bias.to_array()
                np.testing.assert_allclose(got, np.array(rec["bias"]), rtol=1e-12, atol=0)
                assert len(h.last_trace) == len(c.trace_lines(rec["stdout"]))


def test_batch_equals_loop_and_reports_all_bad():
    import tadbit_b200
    a, b = load("synth_3chrom"), load("chrT_40Kb_A")
    ha, hb = _hic(a), _hic(b)
    ha.bads = a.bads(a.out["filter_columns"])
    tadbit_b200.normalize_hic_batch([ha, hb], iterations=20, max_dev=1e-4, silent=True)
    batch = [ha.bias.to_array().copy(), hb.bias.to_array().copy()]
    ha.normalize_hic(iterations=20, max_dev=1e-4, silent=True)
    hb.normalize_hic(iterations=20, max_dev=1e-4, silent=True)
    np.testing.assert_allclose(batch[0], ha.bias.to_array(), rtol=1e-13, atol=0)
    np.testing.assert_allclose(batch[1], hb.bias.to_array(), rtol=1e-13, atol=0)
    bad = load("all_bad_80")
    hbad = _hic(bad)
    hbad.bads = bad.bads(bad.out["filter_columns"])
    with pytest.raises(ZeroDivisionError, match="all bad columns"):
        tadbit_b200.normalize_hic_batch([hb, hbad], iterations=5, max_dev=1e-4, silent=True)
