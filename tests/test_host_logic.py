"""Host-side logic that needs no GPU: constructor symmetrisation (hic_data.py:83-118 of the
reference), the array-backed bias mapping, diagonal cell counts, shard balancing, and the
multi-rank exchange protocol over gloo (world_size 2)."""
import os
import pickle
import socket
import subprocess
import sys
from collections import OrderedDict

import numpy as np
import pytest

import tadbit_oracle as orc
from tadbit_b200 import hic_data as hd
from tadbit_b200 import workloads
from tadbit_b200.normalize_hic import diagonal_cell_counts
from tadbit_b200.vectors import BinVector

HERE = os.path.dirname(os.path.abspath(__file__))


def _reference_symmetricize(items, size):
    """Plain-python replay of the reference's dict behaviour (hic_data.py:83-118)."""
    d = dict(items)
    get = lambda i, j: d.get(i * size + j, 0)
    to_sum, symmetric, count = False, True, 0
    done = False
    for n in list(d):
        i, j = divmod(n, size)
        if i == j or get(i, j) == 0:
            continue
        if not hd.isclose(get(i, j), get(j, i)):
            if get(i, j) != 0 and get(j, i) != 0:
                to_sum = True
            symmetric = False
            break
        if count > 10:
            done = True
            break
        count += 1
    if symmetric or done:
        return d
    if to_sum:
        for n in list(d.keys()):
            i, j = divmod(n, size)
            if i != j:
                v = get(j, i) + get(i, j)
                d[j * size + i] = d[i * size + j] = v
    else:
        for n in list(d.keys()):
            i, j = divmod(n, size)
            d[j * size + i] = d[i * size + j] = d[n]
    return d


def _upper_dict(rows, cols, vals, size):
    return dict((int(r) * size + int(c), int(v)) for r, c, v in zip(rows, cols, vals))


@pytest.mark.parametrize("kind", ["symmetric", "upper_only", "lower_only", "asymmetric_sum",
                                  "mixed_half"])
def test_symmetric_upper_matches_reference_dict(kind):
    rng = np.random.default_rng(5)
    size = 23
    full = {}
    for i in range(size):
        for j in range(i, size):
            if rng.random() < 0.5:
                v = int(rng.integers(1, 50))
                full[i * size + j] = v
                full[j * size + i] = v
    keys = list(full)
    rng.shuffle(keys)
    if kind == "symmetric":
        items = [(k, full[k]) for k in keys]
    elif kind == "upper_only":
        items = [(k, full[k]) for k in keys if k // size <= k % size]
    elif kind == "lower_only":
        items = [(k, full[k]) for k in keys if k // size >= k % size]
    elif kind == "asymmetric_sum":
        items = [(k, full[k] + (3 if k // size < k % size else 0)) for k in keys]
    else:
        items = [(k, full[k]) for k in keys if (k // size <= k % size) or rng.random() < 0.3]
    want_full = _reference_symmetricize(items, size)
    want = dict((k, v) for k, v in want_full.items() if k // size <= k % size)
    ks = np.array([k for k, _ in items], dtype=np.int64)
    vs = np.array([v for _, v in items])
    r, c, v = hd._symmetric_upper(ks, vs, size)
    assert _upper_dict(r, c, v, size) == want
    # and the reference dict is symmetric afterwards, so the upper store loses nothing
    for k, val in want_full.items():
        i, j = divmod(k, size)
        assert want_full[j * size + i] == val


def test_binvector_mapping_protocol(tmp_path):
    b = BinVector([1.5, 2.0, 0.25])
    assert len(b) == 3 and b[1] == 2.0 and list(b) == [0, 1, 2]
    assert dict(b) == {0: 1.5, 1: 2.0, 2: 0.25} == b.to_dict()
    assert b == {0: 1.5, 1: 2.0, 2: 0.25}
    assert 2 in b and 3 not in b and "x" not in b
    with pytest.raises(KeyError):
        b[3]
    assert sorted(b.items()) == [(0, 1.5), (1, 2.0), (2, 0.25)]
    assert pickle.loads(pickle.dumps(b)) == b


def test_diagonal_cell_counts_match_oracle():
    rng = np.random.default_rng(2)
    chroms = OrderedDict([("a", 40), ("b", 25), ("c", 7)])
    n = sum(chroms.values())
    m = orc.OracleMatrix(n, [0], [1], [1], chromosomes=chroms)
    bads = dict((int(i), True) for i in rng.choice(n, 9, replace=False))
    bad = np.zeros(n, dtype=np.uint8)
    bad[list(bads)] = 1
    for ndiag in (40, 72, 5):
        _, want = orc.diagonal_sums(m, bads, ndiag)
        got = diagonal_cell_counts(m.section_pos, n, bad, ndiag)
        assert np.array_equal(got, want)


def test_balanced_row_blocks():
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 1000, 5000)
    for world in (1, 2, 3, 8):
        blocks = workloads.balanced_row_blocks(counts, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == len(counts)
        assert all(blocks[k][1] == blocks[k + 1][0] for k in range(world - 1))
        loads = [counts[a:b].sum() for a, b in blocks]
        assert max(loads) - min(loads) <= 2 * counts.max()
    assert workloads.balanced_row_blocks(np.zeros(10, dtype=np.int64), 4)[-1][1] == 10


def test_workload_shapes():
    assert workloads.workload("c2")["size"] == 24896           # chr1 @ 10 kb
    assert workloads.workload("c3")["size"] == 617665          # hg38 @ 5 kb
    assert sum(workloads.genome_bins(1000).values()) == 3088281


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_protocol_gloo_world2():
    """The multi-rank ICE protocol (own-row slice + zeros, summed; every rank then takes the
    same decision) emulated on CPU ranks with the oracle standing in for the kernels."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "gloo_protocol_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         timeout=300, env=env)
    assert "GLOO_PROTOCOL PASS" in out.stdout, out.stdout[-3000:]


def test_rebalance_by_time_equalises_estimated_time():
    """Shards re-cut from per-kernel probe times spread over per-row work (workloads.py): with
    a matrix whose rows cost more in one region, the new blocks have (nearly) equal estimated
    time and still tile [0, N) contiguously."""
    from tadbit_b200.workloads import balanced_row_blocks, rebalance_by_time
    rng = np.random.RandomState(3)
    n, world = 4000, 4
    dense = rng.randint(1000, 5000, n).astype(np.int64)
    tiles = rng.randint(10, 400, n).astype(np.int64)
    tiles[:700] *= 20                                        # an expensive sparse region
    per_row_ms = 2e-6 * dense + 9e-6 * tiles                 # the "true" cost model
    blocks = balanced_row_blocks(dense + tiles, world)
    probes = []
    for a, b in blocks:
        probes.append((float(9e-6 * tiles[a:b].sum()), float(2e-6 * dense[a:b].sum()),
                       dense[a:b], tiles[a:b]))
    before = [per_row_ms[a:b].sum() for a, b in blocks]
    new = rebalance_by_time(blocks, probes, n)
    assert new[0][0] == 0 and new[-1][1] == n
    assert all(new[r][1] == new[r + 1][0] for r in range(world - 1))
    after = [per_row_ms[a:b].sum() for a, b in new]
    assert max(after) / (sum(after) / world) < 1.02 < max(before) / (sum(before) / world)
    # an empty shard or a kernel without work does not break the apportioning
    probes[1] = (0.0, probes[1][1], probes[1][2], np.zeros_like(probes[1][3]))
    again = rebalance_by_time(blocks, probes, n)
    assert again[0][0] == 0 and again[-1][1] == n


def test_cells_can_be_set_and_the_store_is_rebuilt():
    """HiC_data.__setitem__ (reference: _pytadbit/hic_data.py:146-164): symmetric assignment, pending
    cells visible to item access at once and folded into a new store on the next use -- the host
    logic of tests/test_gpu_api_surface.py, here over the store double."""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import numpy as np
    import pytest
    import tadbit_b200
    from _numpy_store import NumpyStore
    rng = np.random.default_rng(3)
    n = 40
    i, j = np.triu_indices(n)
    keep = rng.random(i.size) < 0.4
    r, c, v = i[keep], j[keep], rng.integers(1, 50, int(keep.sum()))
    full = {}
    for a, b, x in zip(r.tolist(), c.tolist(), v.tolist()):
        full[a * n + b] = x
        full[b * n + a] = x
    hic = tadbit_b200.HiC_data.from_store(NumpyStore(n, r, c, v), dict_sec={})
    first_store = hic.store
    old = hic[0, 1]
    hic[0, 1] = old + 3
    hic[n - 1, 2] = 7                                        # given below the diagonal
    hic[5 * n + 5] = 0                                       # flat position; an explicit zero is a stored cell
    assert hic[0, 1] == old + 3 and hic[1, 0] == old + 3 and hic[2, n - 1] == 7 and hic[5, 5] == 0
    changed = dict(full)
    changed[1] = changed[n] = old + 3
    changed[(n - 1) * n + 2] = changed[2 * n + n - 1] = 7
    changed[5 * n + 5] = 0
    assert dict(hic.items()) == changed
    assert hic.store is not first_store and hic.store is hic.store           # rebuilt once
    assert hic.sum() == sum(changed.values())
    stored, sums = hic.store.row_stats()
    assert int(sums[0]) == sum(x for k, x in changed.items() if k // n == 0)
    assert int(stored[5]) == sum(1 for k in changed if k // n == 5)
    hic[0, 1] = old                                          # and back: same matrix as at the start
    hic[n - 1, 2] = full.get((n - 1) * n + 2, 0)
    got = dict((k, x) for k, x in hic.items() if x != 0 or k in full)
    assert dict((k, x) for k, x in got.items() if x) == dict((k, x) for k, x in full.items() if x)
    with pytest.raises(NotImplementedError):
        hic[0, 1] = 3.5
    with pytest.raises(IndexError):
        hic[n, n + 5] = 1
    with pytest.raises(OverflowError):
        hic[0, 0] = 2 ** 31
