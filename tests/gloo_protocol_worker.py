"""world_size-2 CPU emulation of the row-sharded ICE exchange (run under torchrun, gloo).

Each rank owns an nnz-balanced block of rows of the same matrix, computes the row sums of
its block only, writes them into a zero vector of length N and all-reduces (sum) -- exactly
what the library does with NCCL.  Then every rank runs the bias update redundantly.  Checks:
ranks agree bit-for-bit, and the result equals the single-process oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
sys.path.insert(0, HERE)

import tadbit_oracle as orc            # noqa: E402
from _golden import load               # noqa: E402
from tadbit_b200 import workloads      # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    case = load("synth_3chrom")
    m = orc.OracleMatrix(case.size, case.rows, case.cols, case.vals, chromosomes=case.chromosomes)
    bads = case.bads(case.out["filter_columns"])
    n = m.size
    bad = orc._bad_mask(n, bads)
    keep = ~(bad[m.rows] | bad[m.cols])
    w0 = m.full_csr(keep)
    row_counts = np.diff(w0.indptr)
    # every rank computes counts of an equal slice, gathers, balances -- as build_sharded does
    lo, hi = n * rank // world, n * (rank + 1) // world
    parts = [None] * world
    dist.all_gather_object(parts, row_counts[lo:hi])
    blocks = workloads.balanced_row_blocks(np.concatenate(parts), world)
    rb, re = blocks[rank]
    mine = w0[rb:re]
    present = np.zeros(n, dtype=bool)
    cnt = torch.zeros(n, dtype=torch.float64)
    cnt[rb:re] = torch.from_numpy(np.diff(mine.indptr).astype(np.float64))
    dist.all_reduce(cnt)
    present = cnt.numpy() > 0
    npresent = int(present.sum())
    b = np.ones(n)
    x = np.where(present, 1.0, 0.0)
    iterations, max_dev = 100, 1e-5
    trace = []
    for it in range(iterations + 1):
        part = torch.zeros(n, dtype=torch.float64)
        part[rb:re] = torch.from_numpy(mine @ x)
        dist.all_reduce(part)                              # the one exchange per pass
        s = x * part.numpy()
        mean_s = s[present].sum() / npresent
        b[present] *= s[present] / mean_s
        x = np.where(present & (b != 0), 1.0 / np.where(b != 0, b, 1.0), 0.0)
        dev = max(abs(s[present].min() / mean_s - 1), abs(s[present].max() / mean_s - 1))
        trace.append(dev)
        if dev < max_dev:
            break
    bias = np.where(present & (b != 0), b * mean_s ** .5, 1.0)
    gathered = [None] * world
    dist.all_gather_object(gathered, (bias, len(trace)))
    ok = all(np.array_equal(g[0], gathered[0][0]) and g[1] == gathered[0][1] for g in gathered)
    want, want_trace = orc.iterative(m, bads, iterations, max_dev)
    ok = ok and len(want_trace) == len(trace) and np.allclose(bias, want, rtol=1e-12, atol=0)
    ok = ok and blocks[0][0] == 0 and blocks[-1][1] == n
    if rank == 0:
        print("GLOO_PROTOCOL", "PASS" if ok else "FAIL", blocks)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
