"""The product's HOST layer next to the unmodified reference, on the CPU.

``tadbit_b200.HiC_data`` is built around a test double of the contact store (tests/_numpy_store.py:
the store's reductions answered by the oracle), so everything above the C-ABI runs as shipped --
filter thresholds and warnings, the numpy tail of filter_by_mean, the pooling loop of expected,
sqrt / factor rescaling, printed progress and trace lines, exception types -- and must reproduce
what the reference's own ``HiC_data`` prints and returns on the same seeded random matrices.
Skipped where the reference tree is absent."""
import contextlib
import io
import os
import sys
from collections import OrderedDict

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ref_shim  # noqa: E402
from _numpy_store import NumpyStore  # noqa: E402
from test_oracle_vs_reference import _random_case, _reference  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def _product(k):
    import tadbit_b200
    chroms = OrderedDict(zip(k["names"], k["sizes"])) if k["sizes"] else None
    store = NumpyStore(k["n"], k["rows"], k["cols"], k["vals"], chromosomes=chroms,
                       symmetricized=k["symmetricized"])
    sections = None if chroms else {}
    if chroms:
        sections, idx = {}, 0
        for name in chroms:
            for p in range(chroms[name]):
                sections[(name, p)] = idx
                idx += 1
    return tadbit_b200.HiC_data.from_store(store, chromosomes=chroms, dict_sec=sections,
                                           symmetricized=k["symmetricized"])


def _run(hic, k, factor):
    """filter_columns -> normalize_hic -> normalize_expected with everything printed."""
    out, err = io.StringIO(), io.StringIO()
    raised = None
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
        hic.filter_columns(silent=False, perc_zero=k["perc_zero"], by_mean=k["by_mean"])
        try:
            hic.normalize_hic(iterations=k["iterations"], max_dev=k["max_dev"], silent=False,
                              sqrt=k["sqrt"], factor=factor)
        except ZeroDivisionError as e:
            raised = str(e)
        hic.normalize_expected(signal_to_noise=k["s2n"], inter_chrom=k["inter_chrom"])
    return out.getvalue(), err.getvalue(), raised


@pytest.mark.parametrize("seed", range(80))
def test_host_layer_prints_and_returns_what_the_reference_does(seed):
    ns = ref_shim.load()
    k = _random_case(5000 + seed)
    n = k["n"]
    factor = None if k["factor"] is None else int(k["factor"])
    ref = _reference(ns, k)
    ours = _product(k)
    ref_out, ref_err, ref_raised = _run(ref, k, factor)
    our_out, our_err, our_raised = _run(ours, k, factor)
    assert ours.bads == ref.bads
    for key in ref.bads:
        assert (ours.bads[key] is True) == (ref.bads[key] is True)
    assert our_raised == ref_raised
    if ref_raised is None:
        got = ours.bias.to_array()
        want = np.array([ref.bias[i] for i in range(n)])
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
        assert len(ours.bias) == len(ref.bias) == n
    assert dict(ours.expected) == dict(ref.expected)
    # printed lines: same text, numbers equal to what 3 / 5 printed decimals can show
    ref_lines, our_lines = ref_out.splitlines(), our_out.splitlines()
    assert len(our_lines) == len(ref_lines)
    for a, b in zip(our_lines, ref_lines):
        if a == b:
            continue
        ta, tb = a.split(), b.split()
        assert len(ta) == len(tb)
        for x, y in zip(ta, tb):
            if x != y:
                assert abs(float(x) - float(y)) <= 2e-3 * max(1.0, abs(float(y)) * 1e-9 + 1.0)
    # warnings on stderr (columns removed, thresholds): identical text
    assert our_err == ref_err


@pytest.mark.parametrize("seed", range(20))
def test_dict_surface_min_count_and_pickles(seed, tmp_path):
    """The rest of the reference-facing surface: min_count filtering (with the symmetricized
    doubling and its warning), cell access, items(), CSR export, sum(), and biases pickles that
    the reference's own loader reads (and the other way round)."""
    ns = ref_shim.load()
    k = _random_case(9000 + seed)
    n = k["n"]
    ref, ours = _reference(ns, k), _product(k)
    rng = np.random.RandomState(seed)
    # cell access and iteration over BOTH triangles
    assert len(ours) == len(ref) == n
    for _ in range(50):
        i, j = int(rng.randint(n)), int(rng.randint(n))
        assert ours[i, j] == ref[i, j]
        assert ours[i * n + j] == ref[i * n + j]
    assert sorted(ours.items()) == sorted(dict.items(ref))
    assert np.array_equal(ours.get_hic_data_as_csr().toarray(), ref.get_hic_data_as_csr().toarray())
    # min_count mode
    row_sums = np.asarray(ref.get_hic_data_as_csr().sum(axis=1)).ravel()
    mc = int(np.percentile(row_sums, 25)) + 1
    outs = []
    for hic in (ref, ours):
        err = io.StringIO()
        with contextlib.redirect_stderr(err), contextlib.redirect_stdout(io.StringIO()):
            hic.filter_columns(silent=False, min_count=mc, by_mean=k["by_mean"])
        outs.append(err.getvalue())
    assert ours.bads == ref.bads
    assert outs[1] == outs[0]
    # sums and pickles
    assert ours.sum() == ref.sum()
    for hic in (ref, ours):
        with contextlib.redirect_stdout(io.StringIO()):
            try:
                hic.normalize_hic(iterations=10, max_dev=1e-4, silent=True)
            except ZeroDivisionError:
                return
        hic.normalize_expected()
    assert abs(ours.sum(ours.bias) - ref.sum(ref.bias)) <= 1e-9 * abs(ref.sum(ref.bias))
    a, b = str(tmp_path / "ours.pickle"), str(tmp_path / "ref.pickle")
    ours.save_biases(a)
    ref.save_biases(b)
    ref2, ours2 = _reference(ns, k), _product(k)
    ref2.load_biases(a)                      # the reference reads our file
    ours2.load_biases(b)                     # and we read the reference's
    assert ref2.bads == ours.bads and ours2.bads == ref.bads
    assert ref2.expected == ours.expected and ours2.expected == ref.expected
    np.testing.assert_allclose([ref2.bias[i] for i in range(n)], [ref.bias[i] for i in range(n)],
                               rtol=1e-12, atol=0)
    np.testing.assert_allclose([ours2.bias[i] for i in range(n)], [ref.bias[i] for i in range(n)],
                               rtol=0, atol=0)


@pytest.mark.parametrize("seed", range(20))
def test_module_level_functions_and_constructor(seed, monkeypatch):
    """Seam 2 of INTEGRATION.md -- the four module-level callables the reference binds at import
    time -- and the dict constructor (from_reference), with the store replaced by the double."""
    import tadbit_b200
    import tadbit_b200.hic_data as hd
    ns = ref_shim.load()
    k = _random_case(13000 + seed)
    n = k["n"]
    ref = _reference(ns, k)

    def fake_from_coo(size, rows, cols, vals, chromosomes=None, row_block=None, device=0):
        return NumpyStore(size, rows, cols, vals, chromosomes=chromosomes)
    monkeypatch.setattr(hd.ContactStore, "from_coo", staticmethod(fake_from_coo))
    # the reference object itself (a dict subclass, both triangles) goes through the constructor
    ours = tadbit_b200.from_reference(ref)
    assert len(ours) == n and ours.symmetricized == ref.symmetricized
    assert sorted(ours.items()) == sorted(dict.items(ref))
    assert ours.section_pos == ref.section_pos and ours.chromosomes == ref.chromosomes
    # filter_by_zero_count / filter_by_mean as functions; the passed dict is extended in place
    b_ref = ns.filter_by_zero_count(ref, k["perc_zero"])
    b_our = tadbit_b200.filter_by_zero_count(ours, k["perc_zero"])
    assert b_our == b_ref
    with contextlib.redirect_stderr(io.StringIO()):
        r_ref = ns.filter_by_mean(ref, silent=True, bads=b_ref)
        r_our = tadbit_b200.filter_by_mean(ours, silent=True, bads=b_our)
    assert r_our == r_ref
    assert (r_our is b_our) == (r_ref is b_ref)            # same aliasing rule (empty dict -> new one)
    # iterative / expected as functions, verbose output included
    outs, results = [], []
    for mod, hic, bads in ((ns, ref, r_ref), (tadbit_b200, ours, r_our)):
        buf = io.StringIO()
        try:
            with contextlib.redirect_stdout(buf):
                bias = mod.iterative(hic, bads=bads, iterations=k["iterations"], max_dev=k["max_dev"],
                                     verbose=True)
            results.append(np.array([bias[i] for i in range(n)]))
        except ZeroDivisionError as e:
            results.append(str(e))
        outs.append(buf.getvalue().splitlines())
    if isinstance(results[0], str):
        assert results[1] == results[0]
    else:
        np.testing.assert_allclose(results[1], results[0], rtol=1e-12, atol=0)
        assert len(outs[1]) == len(outs[0])
        assert outs[1][:3] == outs[0][:3]                  # 'iterative correction', '  - copying matrix', ...
    e_ref = ns.expected(ref, bads=r_ref, signal_to_noise=k["s2n"], inter_chrom=k["inter_chrom"])
    e_our = tadbit_b200.expected(ours, bads=r_our, signal_to_noise=k["s2n"], inter_chrom=k["inter_chrom"])
    assert dict(e_our) == dict(e_ref)


def test_batch_host_logic_matches_the_reference_loop(monkeypatch):
    """normalize_hic_batch (tadbit_b200/batch.py) against `for h in hics: h.normalize_hic(...)`
    of the reference: biases, printed lines.  The two C-ABI calls of the batch path are replaced
    by doubles (per-matrix oracle ICE), everything else runs as shipped."""
    import tadbit_b200.batch as batch
    ns = ref_shim.load()
    cases = [_random_case(17000 + s) for s in range(6)]
    for k in cases:                                        # one common set of settings
        k.update(perc_zero=99, by_mean=False)
    refs = [_reference(ns, k) for k in cases]
    ours = [_product(k) for k in cases]
    for r, o in zip(refs, ours):
        with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
            r.filter_columns(silent=True, by_mean=False)
            o.filter_columns(silent=True, by_mean=False)
        assert o.bads == r.bads

    class _Batch(object):
        def __init__(self, stores):
            self.stores = stores

        def close(self):
            pass

    def fake_batch_store(stores, device=0):
        return _Batch(stores)

    def fake_ice_batch(store, sizes, bad=None, iterations=0, max_dev=1e-5):
        bias, traces, passes, off = [], [], [], 0
        for st, n in zip(store.stores, sizes):
            mask = None if bad is None else bad[off:off + n]
            off += n
            try:
                b, t = st.ice(mask, iterations=iterations, max_dev=max_dev)
                passes.append(len(t) if iterations > 0 else 1)
            except ZeroDivisionError:
                b, t = np.ones(n), np.zeros((0, 4))
                passes.append(-1)
            bias.append(b)
            traces.append(t)
        return np.concatenate(bias), traces, np.array(passes, dtype=np.int32)

    monkeypatch.setattr(batch, "batch_store", fake_batch_store)
    monkeypatch.setattr(batch, "ice_batch", fake_ice_batch)
    for kw in (dict(iterations=0, max_dev=0.1), dict(iterations=30, max_dev=1e-4, factor=2),
               dict(iterations=5, max_dev=1e-9, sqrt=True, factor=None)):
        ref_out = io.StringIO()
        ok = True
        with contextlib.redirect_stdout(ref_out):
            try:
                for r in refs:
                    r.normalize_hic(silent=False, **kw)
            except ZeroDivisionError:
                ok = False
        our_out = io.StringIO()
        with contextlib.redirect_stdout(our_out):
            if ok:
                batch.normalize_hic_batch(ours, silent=False, **kw)
            else:
                with pytest.raises(ZeroDivisionError):
                    batch.normalize_hic_batch(ours, silent=False, **kw)
        if not ok:
            continue
        for r, o in zip(refs, ours):
            n = len(r)
            np.testing.assert_allclose(o.bias.to_array(), [r.bias[i] for i in range(n)], rtol=1e-12, atol=0)
        a, b = our_out.getvalue().splitlines(), ref_out.getvalue().splitlines()
        assert len(a) == len(b)
        for x, y in zip(a, b):
            if x != y:
                words = lambda line: [t for t in line.split()
                                      if not t.replace('.', '').replace('-', '').isdigit()]
                assert words(x) == words(y)


@pytest.mark.parametrize("kind", ["all_bad", "all_zero"])
def test_failing_normalisation_prints_and_raises_like_the_reference(kind):
    """Verbose output up to the exception: 'all bad columns' is raised before '  - computing
    biases' is printed (normalize_hic.py:203-210), a zero mean row sum right after it (:148)."""
    ns = ref_shim.load()
    n = 40
    i, j = np.triu_indices(n)
    keep = (j - i) < 3
    rows, cols = i[keep], j[keep]
    vals = np.zeros(rows.size, dtype=np.int64) if kind == "all_zero" else np.ones(rows.size, dtype=np.int64)
    k = dict(n=n, rows=rows, cols=cols, vals=vals, sizes=None, names=None, symmetricized=False)
    ref, ours = _reference(ns, k), _product(k)
    bads = dict((b, True) for b in range(n)) if kind == "all_bad" else {3: True}
    outs = []
    for h in (ref, ours):
        h.bads = dict(bads)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), pytest.raises(ZeroDivisionError) as exc:
            h.normalize_hic(iterations=5, max_dev=1e-5, silent=False)
        outs.append((buf.getvalue(), str(exc.value)))
    assert outs[0] == outs[1]
    assert ("computing biases" in outs[0][0]) == (kind == "all_zero")


@pytest.mark.parametrize("seed", range(12))
def test_cis_trans_ratio_against_the_reference(seed):
    """HiC_data.cis_trans_ratio (hic_data.py:252-316): raw (exact integers -> the same float),
    normalised, without the diagonal, with an excluded chromosome."""
    ns = ref_shim.load()
    k = _random_case(21000 + seed)
    if not k["sizes"]:
        k["sizes"], k["names"] = [k["n"]], ["chrX"]
    ref = _reference(ns, k)
    ours = _product(k)
    for hic in (ref, ours):
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            hic.filter_columns(silent=True, perc_zero=k["perc_zero"], by_mean=k["by_mean"])
    assert ours.bads == ref.bads
    assert ours.cis_trans_ratio() == ref.cis_trans_ratio()
    assert ours.cis_trans_ratio(diagonal=False) == ref.cis_trans_ratio(diagonal=False)
    if len(k["names"]) > 1:
        assert ours.cis_trans_ratio(exclude=[k["names"][0]]) == ref.cis_trans_ratio(exclude=[k["names"][0]])
    # chromosomes merged with their successor: pairs (c0, c1), (c2, c3) by name; all into one
    for equals in (lambda x, y: x[1:].isdigit() and y[1:].isdigit() and int(x[1:]) // 2 == int(y[1:]) // 2,
                   lambda x, y: bool(y)):
        for kw in (dict(), dict(diagonal=False)):
            assert ours.cis_trans_ratio(equals=equals, **kw) == ref.cis_trans_ratio(equals=equals, **kw)
    for hic in (ref, ours):
        with pytest.raises(Exception, match="not normalized"):
            hic.cis_trans_ratio(normalized=True)
    try:
        for hic in (ref, ours):
            with contextlib.redirect_stdout(io.StringIO()):
                hic.normalize_hic(iterations=10, max_dev=1e-4, silent=True)
    except ZeroDivisionError:
        return
    pairs = lambda x, y: x[1:].isdigit() and y[1:].isdigit() and int(x[1:]) // 2 == int(y[1:]) // 2   # noqa: E731
    for kw in (dict(normalized=True), dict(normalized=True, diagonal=False),
               dict(normalized=True, equals=pairs), dict(normalized=True, diagonal=False, equals=pairs)):
        a, b = ours.cis_trans_ratio(**kw), ref.cis_trans_ratio(**kw)
        assert abs(a - b) <= 1e-12 * abs(b), (kw, a, b)


def _reference_cli_biases():
    """Lines of _pytadbit/tools/tadbit_normalize.py from 'Rescaling sum of interactions per bins' to
    the end of the SQRT branch, compiled UNMODIFIED as a function of the variables they read (the
    module itself needs pysam / matplotlib).  The ICE branch (reads a BAM) is cut off."""
    src = open(os.path.join(ref_shim.REFERENCE_ROOT, "_pytadbit", "tools", "tadbit_normalize.py")).read().splitlines()
    first = next(i for i, l in enumerate(src) if "Rescaling sum of interactions per bins" in l)
    ice0 = next(i for i, l in enumerate(src) if l.strip() == "if normalization == 'ICE':")
    van0 = next(i for i, l in enumerate(src) if l.strip() == "elif normalization == 'Vanilla':")
    one0 = next(i for i, l in enumerate(src) if l.strip() == "elif normalization == 'oneD':")
    body = src[first:ice0] + [src[van0].replace("elif", "if", 1)] + src[van0 + 1:one0]
    code = "def _biases(bins, badcol, cisprc, normalization):\n" + "\n".join(body) + "\n    return biases\n"
    space = dict(printime=lambda msg: None, nanmean=np.nanmean)
    exec(compile(code, "tadbit_normalize.py[%d:%d]" % (first + 1, one0), "exec"), space)
    return space["_biases"]


@pytest.mark.parametrize("seed", range(8))
def test_cli_column_biases_against_the_reference_lines(seed):
    """tadbit_b200.column_biases ('Vanilla', 'SQRT') next to the reference CLI's own lines
    (tools/tadbit_normalize.py:896-925) fed with the same per-bin totals."""
    import tadbit_b200
    k = _random_case(31000 + seed)
    ours = _product(k)
    n = k["n"]
    _, totals = ours.store.row_stats()
    cisprc = dict((i, [0, int(t), 0, 0]) for i, t in enumerate(totals.tolist()) if t > 0)
    rng = np.random.default_rng(seed)
    badcol = dict((int(b), True) for b in rng.choice(n, size=max(1, n // 12), replace=False))
    ref_fn = _reference_cli_biases()
    for method in ("Vanilla", "SQRT"):
        want = ref_fn(list(range(n)), dict(badcol), dict(cisprc), method)
        got = tadbit_b200.column_biases(ours, dict(badcol), method)
        assert sorted(got) == sorted(want) == list(range(n))
        a = np.array([got[i] for i in range(n)])
        b = np.array([want[i] for i in range(n)])
        assert np.array_equal(np.isnan(a), np.isnan(b)) and set(np.flatnonzero(np.isnan(b)).tolist()) == set(badcol)
        np.testing.assert_allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=1e-14, atol=0)
    with pytest.raises(NotImplementedError):
        tadbit_b200.column_biases(ours, badcol, "oneD")
    # 'ICE': the object's own normalize_hic(100, 1e-6) with badcol as the mask
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            bias = tadbit_b200.column_biases(ours, dict(badcol), "ICE")
    except ZeroDivisionError:
        return
    assert ours.bads == badcol and bias is ours.bias and len(bias) == n


@pytest.mark.parametrize("seed", range(8))
def test_add_sections_against_the_reference(seed):
    """HiC_data.add_sections (hic_data.py:216-250) on a matrix built without chromosomes: the
    per-chromosome reductions follow the new chromosomes (expected, cis / trans ratio)."""
    ns = ref_shim.load()
    k = _random_case(31000 + seed)
    rng = np.random.RandomState(seed)
    n = k["n"]
    k["sizes"], k["names"] = None, None
    cuts = sorted(set(rng.randint(1, n, rng.randint(1, 4)).tolist()))
    sizes = np.diff([0] + cuts + [n]).tolist()
    names = ["c%d" % i for i in range(len(sizes))] if seed % 2 else None
    ref, ours = _reference(ns, k), _product(k)
    if seed % 3 == 0:                                   # objects that never had sections
        ref.sections, ref.section_pos = None, {}
        ours.sections, ours.section_pos = None, {}
    for hic in (ref, ours):
        hic.add_sections([s - 1 for s in sizes], chr_names=names, binned=True)    # bins = length // 1 + 1
    assert ours.chromosomes == ref.chromosomes and ours.sections == ref.sections
    # the reference keeps the nameless section of an object built without chromosomes (and then
    # counts the cells inside chromosomes twice in `expected`); the product drops it
    assert (None in ref.section_pos) == (seed % 3 != 0) and None not in ours.section_pos
    ref.section_pos.pop(None, None)
    assert ours.section_pos == ref.section_pos
    for hic in (ref, ours):
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            hic.filter_columns(silent=True, perc_zero=k["perc_zero"], by_mean=k["by_mean"])
            hic.normalize_expected()
    assert ours.bads == ref.bads and ours.expected == ref.expected
    assert ours.cis_trans_ratio() == ref.cis_trans_ratio()
    with pytest.warns(UserWarning, match="different sizes"), pytest.raises(ValueError):
        ours.add_sections([n + 5], binned=True)
