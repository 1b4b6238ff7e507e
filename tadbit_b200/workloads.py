"""Synthetic workloads of BASELINE.json (shapes and generator parameters) and the
nnz-balanced row sharding used when one matrix is spread over several GPUs."""
from collections import OrderedDict

import numpy as np

from ._capi import SynthParams
from .store import ContactStore

# hg38 primary assembly, chr1..22, X, Y
HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973,
        145138636, 138394717, 133797422, 135086622, 133275309, 114364328, 107043718,
        101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468,
        156040895, 57227415]
HG38_NAMES = ["chr%d" % i for i in range(1, 23)] + ["chrX", "chrY"]
SEED = 20261019


def genome_bins(resolution, chroms=None):
    """bins per chromosome = len_bp // resolution + 1 (reference: parsers/hic_parser.py:578)."""
    idx = range(len(HG38)) if chroms is None else chroms
    return OrderedDict((HG38_NAMES[i], HG38[i] // resolution + 1) for i in idx)


def workload(name):
    """-> dict(name, chromosomes, size, params, description)."""
    filter_kw = {}
    if name == "c2":        # chr1 at 10 kb, dense
        chroms = genome_bins(10000, [0])
        params = SynthParams(seed=SEED, cis_scale=2.8e5, cis_alpha=1.08, trans_mean=0.0,
                             vis_sigma=0.35, low_frac=0.02, low_scale=0.02, empty_frac=0.005)
        desc = "synthetic human chr1 @10kb, dense Poisson x power-law (BASELINE configs[1])"
    elif name == "c3":      # genome-wide 5 kb, sparse
        chroms = genome_bins(5000)
        params = SynthParams(seed=SEED, cis_scale=1900.0, cis_alpha=1.08, trans_mean=1.2e-3,
                             vis_sigma=0.35, low_frac=0.02, low_scale=0.02, empty_frac=0.005)
        desc = "synthetic genome-wide human @5kb (BASELINE configs[2])"
    elif name == "c4":      # genome-wide 1 kb, 8 GPUs
        chroms = genome_bins(1000)
        params = SynthParams(seed=SEED, cis_scale=2600.0, cis_alpha=1.08, trans_mean=2.4e-4,
                             vis_sigma=0.35, low_frac=0.02, low_scale=0.02, empty_frac=0.005)
        desc = "synthetic genome-wide human @1kb (BASELINE configs[3])"
        # at 1 kb a row holds ~11 000 of 3.1 M cells: the default perc_zero=99 would flag every bin
        filter_kw = dict(perc_zero=99.9)
    elif name == "c3s":     # 1/8 genome (chr1-3) at 5 kb: sparse profile at a size quick to build
        chroms = genome_bins(5000, [0, 1, 2])
        params = SynthParams(seed=SEED, cis_scale=1300.0, cis_alpha=1.08, trans_mean=1.1e-3,
                             vis_sigma=0.35, low_frac=0.02, low_scale=0.02, empty_frac=0.005)
        desc = "synthetic human chr1-3 @5kb (sparse profile of configs[2], reduced)"
    elif name == "small":   # parity-sized
        chroms = OrderedDict([("chrA", 1400), ("chrB", 900), ("chrC", 500)])
        params = SynthParams(seed=SEED, cis_scale=600.0, cis_alpha=1.08, trans_mean=0.02,
                             vis_sigma=0.35, low_frac=0.03, low_scale=0.02, empty_frac=0.01)
        desc = "small 3-chromosome synthetic"
    else:
        raise ValueError("unknown workload %r" % name)
    return dict(name=name, chromosomes=chroms, size=int(sum(chroms.values())), params=params,
                description=desc, filter_kw=filter_kw)


def balanced_row_blocks(row_counts, world):
    """Contiguous row blocks with (nearly) equal numbers of stored cells."""
    row_counts = np.asarray(row_counts)
    acc = np.float64 if row_counts.dtype.kind == 'f' else np.int64
    csum = np.concatenate([[0], np.cumsum(row_counts, dtype=acc)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(csum, total * r / world)))
    bounds.append(len(row_counts))
    for r in range(1, world + 1):
        bounds[r] = max(bounds[r], bounds[r - 1])
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def rebalance_by_time(blocks, probes, size):
    """New contiguous row blocks of equal estimated K1 time.

    ``probes[r]`` = (ms_tiles, ms_dense, dense_bytes_per_row, tile_cells_per_row) of the shard
    that currently owns ``blocks[r]``.  Each kernel's measured time is spread over the shard's
    rows in proportion to the row's share of that kernel's work, which gives a cost per row
    that follows the structure of the matrix inside a block."""
    cost = np.zeros(size)
    for (a, b), (ms_t, ms_d, dense, tiles) in zip(blocks, probes):
        if b <= a:
            continue
        row = np.zeros(b - a)
        dt, tt = float(np.sum(dense)), float(np.sum(tiles))
        row += ms_d * (dense / dt) if dt > 0 else ms_d / (b - a)
        row += ms_t * (tiles / tt) if tt > 0 else ms_t / (b - a)
        cost[a:b] = row
    return balanced_row_blocks(cost, len(blocks))


def build_sharded(wl, rank, world, device, gather=None, refine=4):
    """Generate this rank's block of the synthetic matrix.  ``gather`` is a callable that
    all-gathers a python object over ranks (host plumbing, e.g. torch.distributed).

    Blocks are first balanced by stored cells; a cell does not cost the same in the dense and
    in the sparse part of the row-sum, so the blocks are then re-cut ``refine`` times from the
    measured K1 time of every shard (``tb_probe_rowsum`` on the not yet connected shard)."""
    n = wl["size"]
    if world == 1:
        return ContactStore.synthetic(n, wl["params"], chromosomes=wl["chromosomes"],
                                      device=device), (0, n)
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    mine = ContactStore.synthetic_row_counts(n, wl["params"], chromosomes=wl["chromosomes"],
                                             row_block=(lo, hi), device=device)
    counts = np.concatenate(gather(mine))
    blocks = balanced_row_blocks(counts, world)
    store = ContactStore.synthetic(n, wl["params"], chromosomes=wl["chromosomes"],
                                   row_block=blocks[rank], device=device)
    for _ in range(refine):
        total, ms_t, ms_d = store.probe_rowsum(8)                    # shard alone, timing only
        dense, tiles = store.row_work()
        probes = gather((ms_t, ms_d, dense, tiles))
        times = [p[0] + p[1] for p in probes]
        if max(times) <= 1.006 * (sum(times) / world):
            break
        new = rebalance_by_time(blocks, probes, n)
        if new == blocks:
            break
        blocks = new
        store.close()
        store = ContactStore.synthetic(n, wl["params"], chromosomes=wl["chromosomes"],
                                       row_block=blocks[rank], device=device)
    return store, blocks[rank]
