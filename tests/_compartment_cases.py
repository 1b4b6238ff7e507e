"""Seeded Hi-C matrices with an A/B checkerboard (tests of N4: compartments)."""
from collections import OrderedDict

import numpy as np


def compartment_case(seed, sizes=(70, 45, 3, 1), names=None, n_dead=3, block=6, depth=300.0):
    """Upper-triangle cells of a multi-chromosome matrix whose bins alternate between two
    compartments in blocks of ~``block`` bins; ``n_dead`` bins are emptied (they end up in bads)."""
    rng = np.random.default_rng(seed)
    names = names or ["chr%s" % (k + 1) for k in range(len(sizes))]
    n = int(sum(sizes))
    chrom_of = np.repeat(np.arange(len(sizes)), sizes)
    comp = np.zeros(n, dtype=np.int64)
    pos = 0
    for s in sizes:
        state, i = int(rng.integers(0, 2)), 0
        while i < s:
            run = int(rng.integers(max(block // 2, 1), block * 2))
            comp[pos + i:pos + min(i + run, s)] = state
            state ^= 1
            i += run
        pos += s
    vis = rng.lognormal(0.0, 0.25, n)
    i, j = np.triu_indices(n)
    same_chrom = chrom_of[i] == chrom_of[j]
    same_comp = comp[i] == comp[j]
    lam = np.where(same_chrom, depth / (j - i + 1.0) ** 0.85 * np.where(same_comp, 1.7, 0.6) + 2.0, 0.4)
    lam = lam * vis[i] * vis[j]
    v = rng.poisson(lam)
    dead = rng.choice(n - (sizes[-1] if len(sizes) > 1 else 0) - 1, size=n_dead, replace=False) if n_dead else np.zeros(0, dtype=np.int64)
    keep = (v > 0) & ~np.isin(i, dead) & ~np.isin(j, dead)
    chroms = OrderedDict(zip(names, [int(s) for s in sizes]))
    return dict(n=n, rows=i[keep].astype(np.int64), cols=j[keep].astype(np.int64),
                vals=v[keep].astype(np.int64), chromosomes=chroms, comp=comp, dead=np.sort(dead))


def sections_of(chroms):
    sec, idx = {}, 0
    for name in chroms:
        for p in range(chroms[name]):
            sec[(name, p)] = idx
            idx += 1
    return sec


def full_dict(case):
    n = case["n"]
    d = {}
    for r, c, v in zip(case["rows"].tolist(), case["cols"].tolist(), case["vals"].tolist()):
        d[r * n + c] = v
        d[c * n + r] = v
    return d


def rich_track(case, seed=0):
    """A 'rich in A' track {chromosome: {bin within the chromosome: value}} following compartment 1."""
    rng = np.random.default_rng(1000 + seed)
    out, pos = {}, 0
    for name, size in case["chromosomes"].items():
        per = {}
        for b in range(size):
            if rng.random() < 0.9:
                per[b] = float(case["comp"][pos + b] * 5 + rng.poisson(2))
        if size > 3:
            out[name] = per
        pos += size
    return out
