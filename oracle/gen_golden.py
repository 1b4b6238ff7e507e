#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference (test infrastructure).

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

For every case it stores the INPUT (upper-triangle stored cells, as .npz -- parsed
data, not reference source) and the reference's OUTPUTS (bads dicts, biases, printed
ICE trace, normalised sum, expected dict) as .json.  Cases:

* the reference's own 12 fixtures test/{20,40,80}Kb/chrT/chrT_{A,B,C,D}.tsv
  (the inputs behind the pinned values at test/test_all.py:332,400-402);
* seeded synthetic matrices that exercise the parity traps of SURVEY.md 8(a'):
  several chromosomes, low-coverage and empty bins, stored zeros, rows that touch
  only bad bins, min_count mode on a symmetricized matrix, all-bad failure.
"""
import contextlib
import io
import json
import os
import sys
import warnings
from collections import OrderedDict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

ICE_CONFIGS = [
    # (iterations, max_dev, sqrt, factor)
    (0, 0.1, False, 1),          # method defaults: "Vanilla" single pass
    (100, 1e-5, False, 1),       # SURVEY golden setting
    (3, 1e-9, False, 1),         # iteration-capped
    (100, 1e-6, False, None),    # CLI setting, no rescale
    (10, 1e-3, True, 2),         # sqrt before factor
]


def synth_dense(rng, sizes, a=4000.0, trans=0.6, low_frac=0.06, empty_frac=0.02,
                alpha=1.08):
    """Poisson counts with per-bin visibility, power-law cis decay, flat trans."""
    n = int(sum(sizes))
    vis = rng.lognormal(0.0, 0.35, n)
    low = rng.random(n) < low_frac
    vis[low] *= 0.02
    empty = rng.random(n) < empty_frac
    vis[empty] = 0.0
    chrom = np.repeat(np.arange(len(sizes)), sizes)
    i, j = np.triu_indices(n)
    lam = np.where(chrom[i] == chrom[j],
                   a * (np.abs(i - j) + 1.0) ** -alpha, trans) * vis[i] * vis[j]
    v = rng.poisson(lam)
    keep = v > 0
    return n, i[keep], j[keep], v[keep].astype(np.int64)


def synth_sparse(rng, n, per_row=12):
    """Few contacts per row so that the zero-count filter flags many bins."""
    rows, cols = [], []
    for i in range(n):
        k = rng.integers(0, 2 * per_row) if i % 7 else rng.integers(0, 3)
        off = np.unique(np.minimum(rng.geometric(0.02, k) - 1, n - 1 - i))
        rows.append(np.full(off.size, i))
        cols.append(i + off)
    r = np.concatenate(rows)
    c = np.concatenate(cols)
    v = 1 + rng.geometric(0.4, r.size)
    return n, r, c, v.astype(np.int64)


def build_reference(ns, n, r, c, v, chrom_sizes=None, names=None, symmetricized=False):
    dico = {}
    for a, b, x in zip(r.tolist(), c.tolist(), v.tolist()):
        dico[a * n + b] = x
        dico[b * n + a] = x
    chromosomes = sections = None
    if chrom_sizes is not None:
        chromosomes = OrderedDict(zip(names, [int(s) for s in chrom_sizes]))
        sections, idx = {}, 0
        for name in chromosomes:
            for p in range(chromosomes[name]):
                sections[(name, p)] = idx
                idx += 1
    else:
        sections = {}
    return ns.HiC_data(dico, n, chromosomes=chromosomes, dict_sec=sections,
                       symmetricized=symmetricized)


def upper_store(hic):
    n = len(hic)
    keys = np.fromiter(dict.keys(hic), dtype=np.int64, count=dict.__len__(hic))
    vals = np.array([dict.__getitem__(hic, int(k)) for k in keys])
    i, j = np.divmod(keys, n)
    keep = i <= j
    order = np.argsort(keys[keep])
    return i[keep][order], j[keep][order], vals[keep][order]


def jsonable_bads(bads):
    return OrderedDict((str(k), (True if bads[k] is True else int(bads[k])))
                       for k in sorted(bads))


def run_case(ns, name, hic, ice_configs=ICE_CONFIGS, filter_kw=None, extra=None):
    print("case", name, file=sys.stderr)
    n = len(hic)
    r, c, v = upper_store(hic)
    chroms = list(hic.chromosomes.items()) if hic.chromosomes else None
    np.savez_compressed(os.path.join(OUT, name + ".npz"), size=n, rows=r.astype(np.int32),
                        cols=c.astype(np.int32), vals=v)
    out = OrderedDict(name=name, size=n, nnz_upper=int(r.size), chromosomes=chroms,
                      symmetricized=bool(hic.symmetricized))
    # --- filters --------------------------------------------------------
    out["zero_count_99"] = jsonable_bads(ns.filter_by_zero_count(hic, 99))
    out["zero_count_75"] = jsonable_bads(ns.filter_by_zero_count(hic, 75))
    mc = int(np.percentile(np.bincount(np.concatenate([r, c[r != c]]),
                                       weights=np.concatenate([v, v[r != c]]),
                                       minlength=n), 20)) + 1
    with contextlib.redirect_stderr(io.StringIO()):
        out["min_count"] = dict(min_count=mc,
                                bads=jsonable_bads(ns.filter_by_zero_count(hic, 99, min_count=mc)))
    hic.bads = {}
    hic.filter_columns(silent=True, **(filter_kw or {}))
    out["filter_kw"] = filter_kw or {}
    out["filter_columns"] = jsonable_bads(hic.bads)
    bads = dict(hic.bads)
    # --- ICE ------------------------------------------------------------
    out["ice"] = []
    for iterations, max_dev, sqrt, factor in ice_configs:
        hic.bads = dict(bads)
        buf = io.StringIO()
        rec = OrderedDict(iterations=iterations, max_dev=max_dev, sqrt=sqrt, factor=factor)
        try:
            with contextlib.redirect_stdout(buf):
                hic.normalize_hic(iterations=iterations, max_dev=max_dev, silent=False,
                                  sqrt=sqrt, factor=factor)
            rec["bias"] = [float(hic.bias[i]) for i in range(n)]
            rec["stdout"] = buf.getvalue()
        except ZeroDivisionError as e:
            rec["raises"] = "ZeroDivisionError"
            rec["message"] = str(e)
        out["ice"].append(rec)
    # raw sum and normalised sum with the last successful bias
    hic.bads = dict(bads)
    out["raw_sum"] = int(hic.sum())
    # --- expected -------------------------------------------------------
    out["expected"] = []
    for kw in ({}, {"inter_chrom": True}, {"signal_to_noise": 0.02}):
        hic.bads = dict(bads)
        hic.normalize_expected(**kw)
        out["expected"].append(OrderedDict(
            kw=kw, keys=len(hic.expected),
            values=[[int(d), float(hic.expected[d])] for d in sorted(hic.expected)]))
    if extra:
        out.update(extra)
    with open(os.path.join(OUT, name + ".json"), "w") as fh:
        json.dump(out, fh)
    return out


def reference_known_answers(ns):
    """The two known-answer tests of the reference that run through the balancing path
    (test/test_all.py:312-334 test_09_hic_normalization, :380-405 test_11_write_interaction_pairs),
    replayed with the UNMODIFIED reference classes: Experiment.load_hic_data, filter_columns,
    normalize_hic, get_hic_zscores, write_interaction_pairs.  The asserted literals of those
    tests (4059.2877; 4674 lines, 0.5852295196345679, 0.07060448846960976) are reproduced here and
    stored next to the intermediate biases, so that the oracle and the CUDA path can be pinned to
    the reference's own expectations."""
    import tempfile
    ref_shim._stub("pytadbit.modelling.structuralmodels", StructuralModels=object)
    from pytadbit.experiment import Experiment
    base = os.path.join(ns.root, "test", "20Kb", "chrT")
    out = {"source": "reference test/test_all.py:312-334 and :380-405, replayed by oracle/gen_golden.py",
           "fixture": "chrT_20Kb_A"}
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        # test_09: no column filter, normalize_hic() defaults (one pass, factor=1)
        exp = Experiment("exp1", 20000, hic_data=os.path.join(base, "chrT_D.tsv"))
        exp.load_hic_data(os.path.join(base, "chrT_A.tsv"), silent=True)
        exp.normalize_hic(silent=True)
        exp.get_hic_zscores()
        exp.get_hic_zscores(zscored=False)
        sumz = sum(exp._zscores[k1][k2] for k1 in exp._zscores for k2 in exp._zscores[k1])
        bias = exp.hic_data[0].bias
        out["test_09"] = {"zeros": sorted(int(k) for k in exp._zeros),
                          "bias": [float(bias[i]) for i in range(exp.size)],
                          "sum_normalised": float(sumz), "asserted_rounded_4": 4059.2877,
                          "pairs": sum(len(v) for v in exp._zscores.values())}
        # test_11: filter_columns, normalize_hic(factor=1), normalised values written pair by pair
        exp = Experiment("exp1", 20000, hic_data=os.path.join(base, "chrT_D.tsv"))
        exp.load_hic_data(os.path.join(base, "chrT_A.tsv"), silent=True)
        exp.filter_columns(silent=True)
        exp.normalize_hic(factor=1, silent=True)
        exp.get_hic_zscores(zscored=False)
        with tempfile.TemporaryDirectory() as tmp:
            fname = os.path.join(tmp, "pairs.tsv")
            exp.write_interaction_pairs(fname)
            lines = open(fname).read().splitlines()
        bias = exp.hic_data[0].bias
        out["test_11"] = {"zeros": sorted(int(k) for k in exp._zeros),
                          "bias": [float(bias[i]) for i in range(exp.size)],
                          "lines": len(lines), "asserted_lines": 4674,
                          "line_25": lines[25].split("\t"), "asserted_line_25_value": 0.5852295196345679,
                          "line_2000": lines[2000].split("\t"),
                          "asserted_line_2000_value": 0.07060448846960976}
    assert round(out["test_09"]["sum_normalised"], 4) == 4059.2877
    assert out["test_11"]["lines"] == 4674
    assert abs(float(out["test_11"]["line_25"][2]) - 0.5852295196345679) < 1e-7
    assert abs(float(out["test_11"]["line_2000"][2]) - 0.07060448846960976) < 1e-7
    with open(os.path.join(OUT, "reference_known_answers.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    return out


def _zscore_arrays(exp):
    out = []
    for k1 in exp._zscores:
        for k2, v in exp._zscores[k1].items():
            out.append((int(k1), int(k2), float(v)))
    out.sort()
    a = np.array(out, dtype=np.float64).reshape(-1, 3)
    return a[:, 0].astype(np.int32), a[:, 1].astype(np.int32), a[:, 2]


def reference_pair_cases(ns):
    """N1: interaction pairs and z-scores from the UNMODIFIED ``Experiment`` class
    (_pytadbit/experiment.py:704-799, utils/tadmaths.py:94-172) on a reference fixture and on a
    seeded synthetic matrix: filter_columns -> normalize_hic -> get_hic_zscores in its variants."""
    ref_shim._stub("pytadbit.modelling.structuralmodels", StructuralModels=object)
    from pytadbit.experiment import Experiment
    rng = np.random.default_rng(77)
    cases = {}
    base = os.path.join(ns.root, "test", "20Kb", "chrT")
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        exp = Experiment("exp1", 20000, hic_data=os.path.join(base, "chrT_D.tsv"))
        exp.load_hic_data(os.path.join(base, "chrT_A.tsv"), silent=True)
        cases["chrT_20Kb_A"] = (exp, dict(iterations=0, max_dev=0.1, factor=1))
        n, r, c, v = synth_dense(rng, [70], a=500.0, low_frac=0.1, empty_frac=0.04)
        hic = build_reference(ns, n, r, c, v)
        exp2 = Experiment("synth", 10000)
        exp2.hic_data = [hic]
        exp2.size = n
        cases["pairs_synth_70"] = (exp2, dict(iterations=20, max_dev=1e-4, factor=1))
    index = []
    for name, (exp, kw) in cases.items():
        print("pairs", name, file=sys.stderr)
        hic = exp.hic_data[0]
        r, c, v = upper_store(hic)
        rec = OrderedDict(name=name, size=len(hic), normalize=kw)
        with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()), \
                np.errstate(all="ignore"):
            exp.filter_columns(silent=True)
            exp.normalize_hic(silent=True, **kw)
            rec["zeros"] = sorted(int(k) for k in exp._zeros)
            rec["bias"] = [float(exp.hic_data[0].bias[i]) for i in range(exp.size)]
            arrays = {}
            for tag, opts in (("z", dict()), ("norm", dict(zscored=False)),
                              ("raw", dict(normalized=False, zscored=False)),
                              ("rawz", dict(normalized=False)),
                              ("z_keep0", dict(remove_zeros=False)),
                              ("norm_keep0", dict(zscored=False, remove_zeros=False))):
                exp.get_hic_zscores(**opts)
                arrays[tag] = _zscore_arrays(exp)
            # written pairs, as test_11 does (normalised values, default arguments of the writer)
            import tempfile
            exp.get_hic_zscores(zscored=False)
            with tempfile.TemporaryDirectory() as tmp:
                fname = os.path.join(tmp, "pairs.tsv")
                exp.write_interaction_pairs(fname)
                rec["pair_lines"] = open(fname).read().splitlines()
        np.savez_compressed(os.path.join(OUT, "pairs_" + name + ".npz"), size=len(hic),
                            rows=r.astype(np.int32), cols=c.astype(np.int32), vals=v,
                            **dict(("%s_%s" % (t, k), a) for t, arr in arrays.items()
                                   for k, a in zip(("i", "j", "v"), arr)))
        with open(os.path.join(OUT, "pairs_" + name + ".json"), "w") as fh:
            json.dump(rec, fh)
        index.append(name)
    return index


def reference_window_cases(ns):
    """N3 (HiC_data part): get_matrix / yield_matrix / write_matrix of the UNMODIFIED reference
    (_pytadbit/hic_data.py:578-686, :1514-1582) on a seeded 3-chromosome matrix."""
    import tempfile
    rng = np.random.default_rng(606)
    sizes = [40, 25, 15]
    n, r, c, v = synth_dense(rng, sizes, a=300.0, trans=0.4, low_frac=0.1, empty_frac=0.05)
    names = ["chrA", "chrB", "chrC"]
    hic = build_reference(ns, n, r, c, v, sizes, names)
    hic.resolution = 10000
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        hic.filter_columns(silent=True)
        hic.normalize_hic(iterations=30, max_dev=1e-5, silent=True)
    rec = OrderedDict(name="window_3chrom", size=n, chromosomes=list(zip(names, sizes)), resolution=10000,
                      bads=jsonable_bads(hic.bads), bias=[float(hic.bias[i]) for i in range(n)], calls=[])
    focuses = [None, (5, 30), (3, 20, 41, 70), "chrB", ("chrA", "chrC"), "chrA:50000-250000"]
    for focus in focuses:
        for normalized in (False, True):
            for diagonal in (True, False):
                m = hic.get_matrix(focus=focus, diagonal=diagonal, normalized=normalized)
                item = OrderedDict(fn="get_matrix", focus=focus, normalized=normalized, diagonal=diagonal,
                                   value=[[float(x) for x in row] for row in m])
                rec["calls"].append(item)
                if isinstance(focus, str) and ":" in focus:
                    continue                       # yield_matrix / write_matrix do not parse ranges
                lines = list(hic.yield_matrix(focus=focus, diagonal=diagonal, normalized=normalized))
                rec["calls"].append(OrderedDict(fn="yield_matrix", focus=focus, normalized=normalized,
                                                diagonal=diagonal,
                                                value=[[float(x) for x in row] for row in lines]))
    mm = hic.get_matrix(focus=(2, 30), normalized=True, masked=True)
    rec["masked"] = OrderedDict(focus=(2, 30), mask=np.asarray(mm.mask).astype(int).tolist())
    with tempfile.TemporaryDirectory() as tmp:
        for tag, kw in (("raw", dict()), ("norm_focus", dict(focus="chrB", normalized=True, diagonal=False))):
            fname = os.path.join(tmp, "m.tsv")
            hic.write_matrix(fname, **kw)
            rec["write_" + tag] = open(fname).read()
    np.savez_compressed(os.path.join(OUT, "window_3chrom.npz"), size=n, rows=r.astype(np.int32),
                        cols=c.astype(np.int32), vals=v)
    with open(os.path.join(OUT, "window_3chrom.json"), "w") as fh:
        json.dump(rec, fh)
    return ["window_3chrom"]


class _NowPool(object):
    """Stand-in for multiprocessing.Pool: the reference's tail farms its three helper functions out
    to worker processes; here they run in place (same functions, same arguments)."""

    class _Res(object):
        def __init__(self, value):
            self._v = value

        def get(self):
            return self._v

        def ready(self):
            return True

    def __init__(self, *a, **k):
        pass

    def apply_async(self, fn, args=()):
        return self._Res(fn(*args))

    def close(self):
        pass

    def join(self):
        pass


def reference_decay_tail(root):
    """Compile, UNMODIFIED, the lines of _pytadbit/tools/tadbit_normalize.py that follow the
    computation of the biases inside read_bam (from 'Getting sum of normalized bins' to its return
    statement) together with sum_dec_matrix / get_cis_perc / sum_nrm_matrix, as a function of the
    variables those lines read.  The module itself cannot be imported here (pysam, matplotlib,
    sqlite helpers ...); only the process pool is substituted (_NowPool)."""
    import pickle
    import types
    src = open(os.path.join(root, "_pytadbit", "tools", "tadbit_normalize.py")).read().splitlines()
    first = next(i for i, l in enumerate(src) if "Getting sum of normalized bins" in l)
    last = next(i for i, l in enumerate(src) if l.strip() == "return biases, nrmdec, badcol, raw_cisprc, norm_cisprc")
    h0 = next(i for i, l in enumerate(src) if l.startswith("def sum_dec_matrix"))
    h1 = next(i for i, l in enumerate(src) if l.startswith("class SmartFormatter"))
    code = ("def _tail(outdir, regs, begs, ends, extra_out, biases, ncpus, size, factor, normalize_only, "
            "badcol, bins, section_pos, sections, raw_cisprc):\n" + "\n".join(src[first:last + 1]) +
            "\n\n" + "\n".join(src[h0:h1]) + "\n")
    mu = types.SimpleNamespace(Pool=_NowPool)
    space = dict(path=os.path, load=pickle.load, system=os.system, nansum=np.nansum, mu=mu,
                 printime=lambda msg: None, print_progress=lambda procs: None, print=lambda *a, **k: None)
    exec(compile(code, "tadbit_normalize.py[%d:%d]" % (first + 1, last + 1), "exec"), space)
    return space["_tail"], (first + 1, last + 1)


def reference_decay_cases(ns):
    """N2: the tail of `tadbit normalize` (rescaled biases, cis percentage, per-chromosome decay)
    from the reference's own lines on seeded multi-chromosome matrices."""
    import pickle
    import tempfile
    tail, lines = reference_decay_tail(ns.root)
    rng = np.random.default_rng(4242)
    index = []
    for name, sizes, kind, factor in (("decay_3chrom_vanilla", [90, 60, 35], "vanilla", 1),
                                      ("decay_2chrom_ice", [120, 45], "ice", 1),
                                      ("decay_sparse_factor2", [150, 80], "vanilla_sparse", 2),
                                      ("decay_tiny", [12, 9], "vanilla", 1)):
        print("decay", name, file=sys.stderr)
        if kind == "vanilla_sparse":
            n, r, c, v = synth_dense(rng, sizes, a=6.0, trans=0.004, low_frac=0.1, empty_frac=0.05)
        else:
            n, r, c, v = synth_dense(rng, sizes, a=400.0, trans=0.3, low_frac=0.08, empty_frac=0.04)
        names = ["chr%d" % (k + 1) for k in range(len(sizes))]
        hic = build_reference(ns, n, r, c, v, sizes, names)
        sections = OrderedDict(zip(names, sizes))
        section_pos, total = {}, 0
        for crm in sections:
            section_pos[crm] = (total, total + sections[crm])
            total += sections[crm]
        bins = [(crm, i) for crm in sections for i in range(sections[crm])]
        with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
            hic.filter_columns(silent=True)
            badcol = dict(hic.bads)
            # a bad column on a chromosome border exercises the inclusive test of the cell counts
            badcol[section_pos[names[1]][0]] = True
            if kind == "ice":
                hic.bads = dict(badcol)
                hic.normalize_hic(iterations=100, max_dev=1e-6, silent=True)
                biases = dict((k, float(hic.bias[k])) for k in range(n))
            else:                                       # Vanilla, as read_bam computes it (:901-906)
                tot = np.zeros(n)
                np.add.at(tot, r, v)
                np.add.at(tot, c[r != c], v[r != c])
                col = [float("nan") if k in badcol else float(tot[k]) for k in range(n)]
                mean_col = np.nanmean(col)
                biases = dict((k, b / mean_col * mean_col ** 0.5) for k, b in enumerate(col))
        dico = {}
        for a, b, x in zip(r.tolist(), c.tolist(), v.tolist()):
            dico[(a, b)] = x
            dico[(b, a)] = x
        with tempfile.TemporaryDirectory() as tmp:
            # two regions: the cells are split by their first bin, as read_bam's per-region fetches do
            cut = section_pos[names[1]][0]
            regs, begs, ends = [names[0], names[1]], [0, 0], [10, 20]
            parts = [dict((k, x) for k, x in dico.items() if k[0] < cut),
                     dict((k, x) for k, x in dico.items() if k[0] >= cut)]
            for reg, b0, e0, part in zip(regs, begs, ends, parts):
                with open(os.path.join(tmp, "tmp_%s:%d-%d_%s.pickle" % (reg, b0, e0, "x")), "wb") as fh:
                    pickle.dump(part, fh)
            with np.errstate(all="ignore"), contextlib.redirect_stdout(io.StringIO()):
                out_b, nrmdec, out_bad, raw_c, norm_c = tail(
                    tmp, regs, begs, ends, "x", dict(biases), 1, n, factor, False, dict(badcol), bins,
                    section_pos, sections, 0.5)
        rec = OrderedDict(
            name=name, size=n, chromosomes=list(sections.items()), factor=factor,
            reference_lines="_pytadbit/tools/tadbit_normalize.py:%d-%d" % lines,
            badcol=sorted(int(k) for k in badcol),
            biases_in=[None if np.isnan(biases[k]) else float(biases[k]) for k in range(n)],
            biases_out=[None if np.isnan(out_b[k]) else float(out_b[k]) for k in range(n)],
            norm_cisprc=float(norm_c),
            decay=OrderedDict((crm, [[int(k), float(x)] for k, x in sorted(nrmdec[crm].items())])
                              for crm in sections))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), size=n, rows=r.astype(np.int32),
                            cols=c.astype(np.int32), vals=v)
        with open(os.path.join(OUT, name + ".json"), "w") as fh:
            json.dump(rec, fh)
        index.append(name)
    return index


def reference_compartment_cases(ns):
    """N4: HiC_data.find_compartments of the UNMODIFIED reference (_pytadbit/hic_data.py:736-1031) on
    a reference fixture and on seeded matrices with an A/B checkerboard.  The matrices the
    reference hands to numpy.corrcoef (observed / expected over the good bins) and to scipy's
    eigsh (the correlation matrix, NaN -> 0) are recorded by wrapping those two names in its
    module namespace; nothing of the reference is changed."""
    import pytadbit.hic_data as ref_mod
    sys.path.insert(0, os.path.join(HERE, "..", "tests"))
    from _compartment_cases import compartment_case, full_dict, sections_of
    names = []

    def record(name, hic, prepare, call_kw):
        print("compartments", name, file=sys.stderr)
        seen = {"obsexp": [], "corr": []}
        real_corrcoef, real_eigsh = ref_mod.corrcoef, ref_mod.eigsh

        def spy_corrcoef(m, *a, **k):
            if len(m) > 1:                    # (a single row ends in the reference's "MT" branch)
                seen["obsexp"].append(np.array(m, dtype=np.float64))
            return real_corrcoef(m, *a, **k)

        def spy_eigsh(m, *a, **k):
            seen["corr"].append(np.array(m, dtype=np.float64))
            return real_eigsh(m, *a, **k)
        prepare(hic)
        ref_mod.corrcoef, ref_mod.eigsh = spy_corrcoef, spy_eigsh
        try:
            with warnings.catch_warnings(record=True) as caught:
                warnings.simplefilter("always")
                firsts, stats = hic.find_compartments(**call_kw)
        finally:
            ref_mod.corrcoef, ref_mod.eigsh = real_corrcoef, real_eigsh
        r, c, v = upper_store(hic)
        n = len(hic)
        arrays = dict(rows=r, cols=c, vals=v, bias=np.array([hic.bias[i] for i in range(n)]))
        secs = [sec for sec in hic.section_pos if sec in firsts]
        assert len(secs) == len(seen["obsexp"]) == len(seen["corr"])
        for k, sec in enumerate(secs):
            arrays["obsexp_%d" % k] = seen["obsexp"][k]
            arrays["corr_%d" % k] = seen["corr"][k]
            arrays["evals_%d" % k] = np.array(firsts[sec][0], dtype=np.float64)
            arrays["evecs_%d" % k] = np.array(firsts[sec][1], dtype=np.float64)
        np.savez_compressed(os.path.join(OUT, "compartments_%s.npz" % name), **arrays)
        rec = OrderedDict(
            name=name, n=n,
            chromosomes=[[str(k), int(x)] for k, x in hic.chromosomes.items()] if hic.chromosomes else None,
            resolution=hic.resolution, call=call_kw, sections=[str(s_) for s_ in secs],
            bads=jsonable_bads(hic.bads),
            expected=OrderedDict((str(k), float(x)) for k, x in sorted(hic.expected.items())),
            compartments=OrderedDict((str(k), x) for k, x in hic.compartments.items()),
            warnings=sorted(str(w.message) for w in caught if "Chromosome" in str(w.message)))
        with open(os.path.join(OUT, "compartments_%s.json" % name), "w") as fh:
            json.dump(rec, fh, indent=1)
        names.append(name)

    def nothing(hic):
        pass

    def pipeline(hic):
        hic.filter_columns(silent=True)
        hic.normalize_hic(iterations=50, max_dev=1e-6, silent=True)
        hic.normalize_expected()

    fix = ns.read_matrix(os.path.join(ns.root, "test", "20Kb", "chrT", "chrT_A.tsv"), one=True,
                         resolution=20000)
    record("chrT_20Kb_A", fix, nothing, dict(max_ev=3))
    for name, seed, sizes, prep, kw in (
            ("synth_4chrom", 11, (120, 90, 3, 1), nothing, dict(max_ev=3)),
            ("synth_pipeline", 12, (150, 64, 40), pipeline, dict(max_ev=2)),
            ("synth_300", 13, (300,), pipeline, dict(max_ev=3))):
        case = compartment_case(seed, sizes=sizes, n_dead=max(3, sum(sizes) // 60))
        hic = ns.HiC_data(full_dict(case), case["n"], chromosomes=case["chromosomes"],
                          dict_sec=sections_of(case["chromosomes"]), resolution=10000)
        record(name, hic, prep, kw)
    return names


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_shim.load()
    if "--only-compartments" in sys.argv:
        names = reference_compartment_cases(ns)
        path = os.path.join(OUT, "INDEX.json")
        index = json.load(open(path))
        index["compartment_cases"] = names
        with open(path, "w") as fh:
            json.dump(index, fh, indent=1)
        return
    index = []
    # 1. the reference's own fixtures
    for reso in ("20Kb", "40Kb", "80Kb"):
        for rep in "ABCD":
            path = os.path.join(ns.root, "test", reso, "chrT", "chrT_%s.tsv" % rep)
            with contextlib.redirect_stderr(io.StringIO()):
                hic = ns.read_matrix(path)
            hic = hic[0] if isinstance(hic, list) else hic
            name = "chrT_%s_%s" % (reso, rep)
            run_case(ns, name, hic)
            index.append(name)
    # 2. synthetic, seeded
    rng = np.random.default_rng(20261019)
    sizes = [120, 80, 50]
    n, r, c, v = synth_dense(rng, sizes)
    run_case(ns, "synth_3chrom", build_reference(ns, n, r, c, v, sizes, ["c1", "c2", "c3"]))
    index.append("synth_3chrom")

    n, r, c, v = synth_dense(rng, [400], a=900.0, low_frac=0.1, empty_frac=0.03)
    run_case(ns, "synth_1chrom_400", build_reference(ns, n, r, c, v))
    index.append("synth_1chrom_400")

    n, r, c, v = synth_sparse(rng, 600)
    run_case(ns, "synth_sparse_600", build_reference(ns, n, r, c, v, [350, 250], ["a", "b"]),
             filter_kw={"perc_zero": 97})
    index.append("synth_sparse_600")

    # a low-coverage tail that only the polynomial-fit filter can see
    n = 500
    vis = rng.lognormal(0.0, 0.25, n)
    tail = rng.random(n) < 0.08
    vis[tail] *= rng.uniform(0.05, 0.35, int(tail.sum()))
    i, j = np.triu_indices(n)
    v = rng.poisson(600.0 * (np.abs(i - j) + 1.0) ** -0.9 * vis[i] * vis[j])
    keep = v > 0
    run_case(ns, "synth_lowtail_500",
             build_reference(ns, n, i[keep], j[keep], v[keep].astype(np.int64)),
             ice_configs=[(0, 0.1, False, 1), (100, 1e-5, False, 1), (100, 1e-6, False, None)])
    index.append("synth_lowtail_500")

    # stored zeros: row 5 holds only explicit zero cells; row 9 only touches row 5
    n, r, c, v = synth_dense(rng, [60], a=300.0, low_frac=0.0, empty_frac=0.0)
    keep = (r != 5) & (c != 5)
    r, c, v = r[keep], c[keep], v[keep]
    r = np.concatenate([r, [5, 5, 3]])
    c = np.concatenate([c, [5, 11, 5]])
    v = np.concatenate([v, [0, 0, 0]])
    hic = build_reference(ns, n, r, c, v)
    run_case(ns, "stored_zero_60", hic, filter_kw={"by_mean": False},
             ice_configs=[(0, 0.1, False, 1), (6, 1e-5, False, 1), (4, 1e-5, True, None)])
    index.append("stored_zero_60")

    # symmetricized flag (min_count doubling) + rows touching only bad bins
    n, r, c, v = synth_dense(rng, [90, 40], a=500.0, low_frac=0.15, empty_frac=0.05)
    hic = build_reference(ns, n, r, c, v, [90, 40], ["x", "y"], symmetricized=True)
    run_case(ns, "synth_symmetricized", hic)
    index.append("synth_symmetricized")

    # all-bad failure: every bin flagged by a harsh zero-count filter
    n, r, c, v = synth_sparse(rng, 80, per_row=2)
    hic = build_reference(ns, n, r, c, v)
    run_case(ns, "all_bad_80", hic, filter_kw={"perc_zero": 1, "by_mean": False},
             ice_configs=[(0, 0.1, False, 1), (10, 1e-5, False, 1)])
    index.append("all_bad_80")

    # every stored cell is an explicit zero: rows are "present" (copy_matrix keeps the cells) but the
    # mean row sum is 0 -> ZeroDivisionError('float division by zero') in _updateDB
    n = 30
    i, j = np.triu_indices(n, 0)
    keep = (j - i) < 4
    hic = build_reference(ns, n, i[keep], j[keep], np.zeros(int(keep.sum()), dtype=np.int64))
    run_case(ns, "all_zero_30", hic, filter_kw={"perc_zero": 99, "by_mean": False},
             ice_configs=[(0, 0.1, False, 1), (5, 1e-5, False, None)])
    index.append("all_zero_30")

    reference_known_answers(ns)
    pair_cases = reference_pair_cases(ns)
    decay_cases = reference_decay_cases(ns)
    window_cases = reference_window_cases(ns)
    compartment_cases = reference_compartment_cases(ns)

    with open(os.path.join(OUT, "INDEX.json"), "w") as fh:
        json.dump(dict(cases=index, pair_cases=pair_cases, decay_cases=decay_cases, window_cases=window_cases, compartment_cases=compartment_cases,
                       numpy=np.__version__,
                       generator="oracle/gen_golden.py", reference="TADbit v1.1 (unmodified, via oracle/ref_shim.py)"), fh, indent=1)


if __name__ == "__main__":
    main()
