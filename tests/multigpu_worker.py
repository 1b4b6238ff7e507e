"""Run under torchrun (one rank per GPU): row-sharded path vs the single-GPU path.

Each rank builds its nnz-balanced row block of the same synthetic matrix, joins the NCCL
communicator of the library, and runs filter statistics, ICE, normalised sum and diagonal
sums.  Rank 0 also builds the whole matrix on its GPU and compares: integers bit-exact,
biases bit-identical (same chunk partials, same reduction order; the all-reduce only adds
zeros)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import tadbit_b200
    from tadbit_b200 import workloads
    from tadbit_b200._capi import bad_mask

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    wl = workloads.workload(sys.argv[1] if len(sys.argv) > 1 else "small")

    def gather(arr):
        out = [None] * world
        dist.all_gather_object(out, arr)
        return out

    store, block = workloads.build_sharded(wl, rank, world, local, gather)
    uid = [tadbit_b200.ContactStore.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    store.comm_init(rank, world, uid[0])
    hic = tadbit_b200.HiC_data.from_store(store, chromosomes=wl["chromosomes"])
    hic.filter_columns(silent=True)
    hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True, factor=None)
    raw_bias = hic.bias.to_array().copy()
    hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
    hic.normalize_expected()
    nnz, rsum = store.row_stats()
    raw = hic.sum()
    exchange = store.timing()["exchange"]
    # the same ICE with the pass as separate steps around an NCCL all-reduce (the parity reference
    # of the fused exchange kernel)
    store.set_option("exchange", "nccl")
    hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True, factor=None)
    nccl_bias, nccl_passes = hic.bias.to_array().copy(), len(hic.last_trace)
    store.set_option("exchange", "fused")
    hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
    # K1 with its two kernels back to back instead of on two streams: the same bits
    store.set_option("k1_streams", "1")
    hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
    serial_bias = hic.bias.to_array().copy()
    store.set_option("k1_streams", "2")
    # N4: observed / expected and correlation matrix of the first chromosome from the row-sharded
    # store (every rank ends with the whole matrix), leading eigenpairs
    import hashlib
    sec = list(wl["chromosomes"])[0]
    c_beg, c_end = hic.section_pos[sec]
    c_end = min(c_end, c_beg + 3000)              # a 3000-bin stretch is enough (n x n doubles per rank)
    digest = lambda arr: hashlib.sha1(np.ascontiguousarray(arr).tobytes()).hexdigest()   # noqa: E731
    c_mask = bad_mask(wl["size"], hic.bads)
    c_exp = np.array([hic.expected[d] or 1.0 for d in range(c_end - c_beg)])
    store.corr_build(c_beg, c_end, hic.bias.to_array(), c_mask, c_exp, stage=0)
    cmp_oe = digest(store.corr_get())
    store.corr_build(c_beg, c_end, hic.bias.to_array(), c_mask, c_exp, stage=1)
    cmp_corr = digest(store.corr_get())
    cmp_vals, cmp_vecs, _ = store.corr_eig(3)
    # pre-partitioned input: this rank's rows exported in CSR form and a second store built from them
    ptr, cc, vv = store.export_csr("full_rows")
    again = tadbit_b200.ContactStore.from_csr(wl["size"], ptr, cc, vv, form="full_rows",
                                              chromosomes=wl["chromosomes"], row_block=block, device=local)
    same_rows = all(np.array_equal(a, b) for a, b in zip(again.export_csr("full_rows"), (ptr, cc, vv)))
    same_rows &= again.info()["nnz_upper"] == store.info()["nnz_upper"]
    again.close()
    result = dict(bads=hic.bads, bias=hic.bias.to_array(), raw_bias=raw_bias, expected=hic.expected, nnz=nnz,
                  rsum=rsum, raw=raw, passes=len(hic.last_trace), block=block,
                  nnz_upper=store.info()["nnz_upper"], exchange=exchange, nccl_bias=nccl_bias,
                  nccl_passes=nccl_passes, same_rows=bool(same_rows), serial_bias=serial_bias,
                  cmp_oe=cmp_oe, cmp_corr=cmp_corr, cmp_vals=cmp_vals, cmp_vecs=cmp_vecs)
    allres = gather(result)
    ok = True
    if rank == 0:
        # every rank must hold the same answer
        for r in range(1, world):
            ok &= allres[r]["bads"] == allres[0]["bads"]
            ok &= np.array_equal(allres[r]["bias"], allres[0]["bias"])
            ok &= allres[r]["expected"] == allres[0]["expected"]
        one = tadbit_b200.ContactStore.synthetic(wl["size"], wl["params"],
                                                 chromosomes=wl["chromosomes"], device=local)
        ref = tadbit_b200.HiC_data.from_store(one, chromosomes=wl["chromosomes"])
        ref.filter_columns(silent=True)
        ref.normalize_hic(iterations=100, max_dev=1e-5, silent=True, factor=None)
        ref_raw = ref.bias.to_array().copy()
        ref.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
        ref.normalize_expected()
        n1, s1 = one.row_stats()
        one.corr_build(c_beg, c_end, hic.bias.to_array(), c_mask, c_exp, stage=0)
        one_oe = digest(one.corr_get())
        one.corr_build(c_beg, c_end, hic.bias.to_array(), c_mask, c_exp, stage=1)
        one_corr = digest(one.corr_get())
        one_vals, one_vecs, _ = one.corr_eig(3)
        checks = {
            "blocks cover rows": [b["block"] for b in allres][0][0] == 0 and
                                 allres[-1]["block"][1] == wl["size"],
            "nnz_upper adds up": sum(b["nnz_upper"] for b in allres) == one.info()["nnz_upper"],
            "row nnz": np.array_equal(n1, nnz), "row sums": np.array_equal(s1, rsum),
            "bads": ref.bads == hic.bads, "raw sum": ref.sum() == raw,
            "passes": len(ref.last_trace) == len(hic.last_trace),
            # ICE itself: every rank holds identical bits (checked above); against ONE GPU the
            # sparse tiles of a shard are cut differently, so the summation order differs
            "ICE bias 1e-13": np.allclose(ref_raw, raw_bias, rtol=1e-13, atol=0),
            # factor step sums per-rank partial dot products: 1e-13 relative
            "rescaled bias 1e-13": np.allclose(ref.bias.to_array(), hic.bias.to_array(),
                                               rtol=1e-13, atol=0),
            "expected": ref.expected == hic.expected,
            # fused exchange kernel vs combine -> ncclAllReduce -> statistics -> update
            "fused vs nccl passes": all(b["nccl_passes"] == b["passes"] for b in allres),
            "fused vs nccl 1e-13": np.allclose(allres[0]["nccl_bias"], raw_bias, rtol=1e-13, atol=0),
            "nccl path ranks identical": all(np.array_equal(b["nccl_bias"], allres[0]["nccl_bias"])
                                             for b in allres),
            "rows round trip (CSR form)": all(b["same_rows"] for b in allres),
            "K1 one stream = two streams": all(np.array_equal(b["serial_bias"], b["bias"]) for b in allres),
            "obs/exp = one GPU (bits)": all(b["cmp_oe"] == one_oe for b in allres),
            "correlation = one GPU (bits)": all(b["cmp_corr"] == one_corr for b in allres),
            "eigenpairs = one GPU (bits)": all(np.array_equal(b["cmp_vals"], one_vals) and
                                               np.array_equal(b["cmp_vecs"], one_vecs) for b in allres),
        }
        print("exchange in use:", [b["exchange"] for b in allres])
        for k, v in checks.items():
            print("%-22s %s" % (k, "ok" if v else "FAIL"))
            ok &= bool(v)
        print("MULTIGPU_RESULT", "PASS" if ok else "FAIL", "world", world,
              "blocks", [b["block"] for b in allres])
    dist.barrier()
    store.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
