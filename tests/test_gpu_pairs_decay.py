"""N1 / N2 through the CUDA kernels (tb_pair_stats, tb_pair_values, tb_decay_sums): against the
vectors of the unmodified reference (tests/golden, see test_oracle_pairs_decay.py), against the
reference's own known answers (test_09: sum of the normalised pairs 4059.2877; test_11: 4674
written lines) and, at 20 000 bins, against the oracle."""
import io
import json
import os
from collections import OrderedDict

import numpy as np
import pytest

import tadbit_oracle as orc
from _golden import GOLDEN
from test_oracle_pairs_decay import DECAY_CASES, PAIR_CASES, check_pairs, load_decay, load_pairs

pytestmark = pytest.mark.gpu


def _experiment(rec, z):
    import tadbit_b200
    hic = tadbit_b200.HiC_data.from_coo(rec["size"], z["rows"], z["cols"], z["vals"], dict_sec={})
    return tadbit_b200.Experiment("exp", 20000, hic_data=hic)


@pytest.mark.parametrize("name", PAIR_CASES)
def test_pairs_and_zscores(name):
    rec, z = load_pairs(name)
    exp = _experiment(rec, z)
    exp.filter_columns(silent=True)
    assert sorted(exp._zeros) == rec["zeros"]
    exp.normalize_hic(silent=True, **rec["normalize"])
    np.testing.assert_allclose(exp.hic_data[0].bias.to_array(), np.array(rec["bias"]), rtol=1e-12)
    # the arrays straight from the kernels
    check_pairs(exp.pair_zscores(zscored=True), z, "z")
    check_pairs(exp.pair_zscores(zscored=False), z, "norm")
    n, vmin, mean, std = exp._zscore_stats if hasattr(exp, "_zscore_stats") else (0, 0, 0, 0)
    assert n == z["z_v"].size
    # the reference's dictionaries in every variant
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
    for a, b in zip(lines[::7], rec["pair_lines"][::7]):
        fa, fb = a.split("\t"), b.split("\t")
        assert fa[:2] == fb[:2] and abs(float(fa[2]) - float(fb[2])) <= 1e-10 * abs(float(fb[2]))


def test_reference_known_answers_through_the_kernels():
    """test/test_all.py:312-334 (test_09) and :380-405 (test_11) of the reference, with the pairs
    coming out of tb_pair_values instead of the N^2 dictionary."""
    import tadbit_b200
    known = json.load(open(os.path.join(GOLDEN, "reference_known_answers.json")))
    z = np.load(os.path.join(GOLDEN, "chrT_20Kb_A.npz"))
    n = int(z["size"])
    hic = tadbit_b200.HiC_data.from_coo(n, z["rows"], z["cols"], z["vals"], dict_sec={})
    exp = tadbit_b200.Experiment("exp1", 20000, hic_data=hic)
    exp.normalize_hic(silent=True)                        # test_09: no filter, defaults
    exp.get_hic_zscores()
    exp.get_hic_zscores(zscored=False)
    total = sum(exp._zscores[a][b] for a in exp._zscores for b in exp._zscores[a])
    assert round(total, 4) == 4059.2877
    assert sum(len(v) for v in exp._zscores.values()) == known["test_09"]["pairs"]
    hic = tadbit_b200.HiC_data.from_coo(n, z["rows"], z["cols"], z["vals"], dict_sec={})
    exp = tadbit_b200.Experiment("exp1", 20000, hic_data=hic)
    exp.filter_columns(silent=True)                       # test_11
    exp.normalize_hic(factor=1, silent=True)
    exp.get_hic_zscores(zscored=False)
    out = io.StringIO()
    out.close = lambda: None
    exp.write_interaction_pairs(out)
    lines = out.getvalue().splitlines()
    assert len(lines) == 4674
    assert abs(float(lines[25].split("\t")[2]) - 0.5852295196345679) < 1e-7
    assert abs(float(lines[2000].split("\t")[2]) - 0.07060448846960976) < 1e-7


@pytest.mark.parametrize("name", DECAY_CASES)
def test_decay(name):
    import tadbit_b200
    from tadbit_b200.decay import rescale_and_decay
    rec, z, chroms = load_decay(name)
    hic = tadbit_b200.HiC_data.from_coo(rec["size"], z["rows"], z["cols"], z["vals"], chromosomes=chroms)
    badcol = dict((k, True) for k in rec["badcol"])
    biases, decay, _, cis = rescale_and_decay(hic, dict(enumerate(rec["biases_in"].tolist())), badcol,
                                              factor=rec["factor"])
    got = np.array([biases[k] for k in range(rec["size"])])
    ok = ~np.isnan(rec["biases_out"])
    assert np.array_equal(np.isnan(got), ~ok)
    np.testing.assert_allclose(got[ok], rec["biases_out"][ok], rtol=1e-12)
    assert abs(cis - rec["norm_cisprc"]) <= 1e-12 * rec["norm_cisprc"]
    for crm in chroms:
        assert sorted(decay[crm]) == sorted(rec["decay"][crm]), crm
        for k, v in rec["decay"][crm].items():
            assert abs(decay[crm][k] - v) <= 1e-10 * abs(v) + 1e-300, (crm, k)


def test_pairs_and_decay_at_20k_bins_against_the_oracle():
    import tadbit_b200
    from tadbit_b200.decay import rescale_and_decay
    chroms = OrderedDict([("c1", 9000), ("c2", 6500), ("c3", 4500)])
    n = sum(chroms.values())
    params = tadbit_b200.SynthParams(seed=21, cis_scale=900.0, cis_alpha=1.08, trans_mean=0.004,
                                     vis_sigma=0.35, low_frac=0.03, low_scale=0.02, empty_frac=0.01)
    store = tadbit_b200.ContactStore.synthetic(n, params, chromosomes=chroms)
    hic = tadbit_b200.HiC_data.from_store(store, chromosomes=chroms)
    hic.filter_columns(silent=True)
    hic.normalize_hic(iterations=50, max_dev=1e-5, silent=True)
    r, c, v = store.export_upper()
    m = orc.OracleMatrix(n, r, c, v, chromosomes=chroms)
    bias = hic.bias.to_array()
    exp = tadbit_b200.Experiment("e", 5000, hic_data=hic)
    exp.norm = [None]
    rr, cc, zz = exp.pair_zscores(zscored=True)
    orr, occ, ozz, (on, omin, omean, ostd) = orc.pair_zscores(m, bias, hic.bads)
    assert np.array_equal(rr, orr) and np.array_equal(cc, occ)
    np.testing.assert_allclose(zz, ozz, rtol=1e-9, atol=1e-11)
    gn, gmin, gmean, gstd = exp._zscore_stats
    assert gn == on and abs(gmin - omin) <= 1e-12 * omin
    assert abs(gmean - omean) < 1e-12 * abs(omean) + 1e-13 and abs(gstd - ostd) < 1e-12 * ostd
    biases, decay, _, cis = rescale_and_decay(hic, hic.bias, hic.bads)
    ob, odecay, ocis = orc.rescale_and_decay(m, bias, hic.bads)
    np.testing.assert_allclose(np.array([biases[k] for k in range(n)]), ob, rtol=1e-12)
    assert abs(cis - ocis) <= 1e-11 * ocis
    for crm in chroms:
        assert sorted(decay[crm]) == sorted(odecay[crm])
        a = np.array([decay[crm][k] for k in sorted(decay[crm])])
        b = np.array([odecay[crm][k] for k in sorted(odecay[crm])])
        np.testing.assert_allclose(a, b, rtol=1e-10)
