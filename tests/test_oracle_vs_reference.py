"""The oracle against the UNMODIFIED reference, live, on seeded random matrices.

Runs only where the reference tree exists (the build container; it is skipped on the GPU box):
the reference's HiC_data / filter_columns / normalize_hic / normalize_expected are imported
through oracle/ref_shim.py and called on small random matrices -- several chromosomes or none,
stored zeros, low-coverage and empty bins, the symmetricized flag, random ICE settings -- and the
oracle must give the same masks (bit-exact, dict values included), the same biases (1e-12: the
oracle sums in another order), the same printed trace lines and the same expected dict (bit-exact).
Complements the committed golden vectors, which freeze 20 fixed cases."""
import contextlib
import io
import os
import sys
from collections import OrderedDict

import numpy as np
import pytest

import tadbit_oracle as orc

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")

TRACE_FMT = '   %15.3f %15.3f %15.3f %4s %9.5f'


def _random_case(seed):
    rng = np.random.RandomState(seed)
    nchrom = rng.randint(0, 4)
    if nchrom:
        sizes = [int(x) for x in rng.randint(15, 60, nchrom)]
    else:
        sizes = [int(rng.randint(30, 120))]
    n = sum(sizes)
    vis = rng.lognormal(0.0, 0.4, n)
    vis[rng.rand(n) < 0.08] *= 0.03
    vis[rng.rand(n) < 0.04] = 0.0
    chrom = np.repeat(np.arange(len(sizes)), sizes)
    i, j = np.triu_indices(n)
    a = float(rng.choice([30.0, 300.0, 3000.0]))
    lam = np.where(chrom[i] == chrom[j], a * (np.abs(i - j) + 1.0) ** -1.0, a / 400.0) * vis[i] * vis[j]
    v = rng.poisson(lam)
    keep = v > 0
    if rng.rand() < 0.4:                       # a few stored zeros (HiC_data.__setitem__ allows them)
        zero = (~keep) & (rng.rand(keep.size) < 0.01)
        keep = keep | zero
    r, c, v = i[keep], j[keep], v[keep].astype(np.int64)
    names = ["c%d" % k for k in range(len(sizes))] if nchrom else None
    return dict(n=n, rows=r, cols=c, vals=v, sizes=sizes if nchrom else None, names=names,
                symmetricized=bool(rng.rand() < 0.3),
                perc_zero=int(rng.choice([99, 95, 80])), by_mean=bool(rng.rand() < 0.7),
                iterations=int(rng.choice([0, 3, 25, 100])),
                max_dev=float(rng.choice([0.1, 1e-3, 1e-6])),
                sqrt=bool(rng.rand() < 0.25), factor=rng.choice([1, 2, None]),
                inter_chrom=bool(rng.rand() < 0.3), s2n=float(rng.choice([0.05, 0.1, 0.02])))


def _reference(ns, k):
    n = k["n"]
    dico = {}
    for a, b, x in zip(k["rows"].tolist(), k["cols"].tolist(), k["vals"].tolist()):
        dico[a * n + b] = x
        dico[b * n + a] = x
    chromosomes, sections = None, {}
    if k["sizes"]:
        chromosomes = OrderedDict(zip(k["names"], k["sizes"]))
        idx = 0
        for name in chromosomes:
            for p in range(chromosomes[name]):
                sections[(name, p)] = idx
                idx += 1
    return ns.HiC_data(dico, n, chromosomes=chromosomes, dict_sec=sections,
                       symmetricized=k["symmetricized"])


@pytest.mark.parametrize("seed", range(40))
def test_random_matrix(seed):
    ns = ref_shim.load()
    k = _random_case(1000 + seed)
    n = k["n"]
    factor = None if k["factor"] is None else int(k["factor"])
    hic = _reference(ns, k)
    chroms = OrderedDict(zip(k["names"], k["sizes"])) if k["sizes"] else None
    m = orc.OracleMatrix(n, k["rows"], k["cols"], k["vals"], chromosomes=chroms,
                         symmetricized=k["symmetricized"])
    # filter_columns
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        hic.filter_columns(silent=True, perc_zero=k["perc_zero"], by_mean=k["by_mean"])
        bads = orc.filter_columns(m, silent=True, perc_zero=k["perc_zero"], by_mean=k["by_mean"])
    assert bads == hic.bads
    for key in bads:
        assert (bads[key] is True) == (hic.bads[key] is True)
    # normalize_hic
    buf = io.StringIO()
    raised = False
    try:
        with contextlib.redirect_stdout(buf):
            hic.normalize_hic(iterations=k["iterations"], max_dev=k["max_dev"], silent=False,
                              sqrt=k["sqrt"], factor=factor)
    except ZeroDivisionError:
        raised = True
    if raised:
        with pytest.raises(ZeroDivisionError):
            orc.normalize_hic(m, bads=bads, iterations=k["iterations"], max_dev=k["max_dev"],
                              sqrt=k["sqrt"], factor=factor)
    else:
        bias, trace, _ = orc.normalize_hic(m, bads=bads, iterations=k["iterations"],
                                           max_dev=k["max_dev"], sqrt=k["sqrt"], factor=factor)
        want = np.array([hic.bias[i] for i in range(n)])
        np.testing.assert_allclose(bias, want, rtol=1e-12, atol=0)
        ref_lines = [l for l in buf.getvalue().splitlines() if l.startswith("   ") and len(l.split()) == 5]
        assert len(trace) == len(ref_lines)
        for t, line in zip(trace, ref_lines):           # printed with 3 / 5 decimals
            assert TRACE_FMT % tuple(t) == line or \
                np.allclose([float(x) for x in line.split()], [float(x) for x in (TRACE_FMT % tuple(t)).split()],
                            rtol=0, atol=2e-3)
        assert orc.matrix_sum(m, None, bads) == hic.sum()
        # the C + OpenMP port of the pass loop (bench.py's CPU baseline) against iterative() itself
        raw = ns.iterative(hic, bads=hic.bads, iterations=k["iterations"], max_dev=k["max_dev"])
        fast, ftrace, _, _ = orc.iterative_fast(m, bads, k["iterations"], k["max_dev"], 2)
        np.testing.assert_allclose(fast, np.array([raw[i] for i in range(n)]), rtol=1e-12, atol=0)
        assert len(ftrace) == len(trace)
    # expected
    hic.normalize_expected(signal_to_noise=k["s2n"], inter_chrom=k["inter_chrom"])
    got = orc.expected(m, bads=bads, signal_to_noise=k["s2n"], inter_chrom=k["inter_chrom"])
    assert got == dict(hic.expected)
