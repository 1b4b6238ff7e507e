#!/usr/bin/env python
"""N4 (compartments) on one chromosome of a synthetic workload: device times of the observed /
expected visitor, the fp64 X X^T (DMMA tiles), the scaling and the Lanczos eigen-solver, plus a
host check of the result against numpy on a leading block.

    python tools/bench_compartments.py --workload c2 [--simple] [--reps 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--simple", action="store_true", help="also time the plain parity kernel")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", type=int, default=1500, help="rows of the correlation matrix recomputed with numpy")
    args = ap.parse_args()
    import tadbit_b200
    from tadbit_b200 import workloads
    from tadbit_b200._capi import bad_mask
    wl = workloads.workload(args.workload)
    store, _ = workloads.build_sharded(wl, 0, 1, 0, None)
    hic = tadbit_b200.HiC_data.from_store(store, chromosomes=wl["chromosomes"])
    hic.filter_columns(silent=True, **wl["filter_kw"])
    hic.normalize_hic(iterations=100, max_dev=1e-5, silent=True)
    hic.normalize_expected()
    n = wl["size"]
    sec = max(wl["chromosomes"], key=lambda c: wl["chromosomes"][c])
    beg, end = hic.section_pos[sec]
    mask = bad_mask(n, hic.bads)
    bias = hic.bias.to_array()
    exp_arr = np.array([hic.expected.get(d, 1.0) or 1.0 for d in range(end - beg)])
    out = {"workload": wl["description"], "section": sec, "bins": end - beg, "runs": []}
    for gemm in ["dmma"] + (["simple"] if args.simple else []):
        store.set_option("corr_gemm", gemm)
        for rep in range(args.reps):
            t0 = time.perf_counter()
            n_good = store.corr_build(beg, end, bias, mask, exp_arr, stage=1)
            t1 = time.perf_counter()
            evals, evecs, info = store.corr_eig(3)
            t2 = time.perf_counter()
            ct = store.corr_timing()
            ld = -(-n_good // 16) * 16
            ntile = -(-n_good // 128)
            flop = (ntile * (ntile + 1) / 2 * 128 * 128 * ld * 2.0) if gemm == "dmma" else \
                (-(-n_good // 16) * 16) ** 2 * ld * 2.0
            out["runs"].append({"gemm": gemm, "rep": rep, "good_bins": n_good, "build_wall_ms": (t1 - t0) * 1e3,
                                "eig_wall_ms": (t2 - t1) * 1e3, "timing": ct, "flop": flop,
                                "tflops": flop / (ct["gemm_ms"] * 1e-3) / 1e12, "lanczos": info,
                                "eigenvalues": [float(x) for x in evals]})
    # host check: a block of rows of the correlation matrix and the eigenpairs' residuals
    store.set_option("corr_gemm", "dmma")
    store.corr_build(beg, end, bias, mask, exp_arr, stage=0)
    oe = store.corr_get()
    store.corr_build(beg, end, bias, mask, exp_arr, stage=1)
    corr = store.corr_get()
    evals, evecs, info = store.corr_eig(3)
    k = min(args.check, oe.shape[0])
    x = oe - oe.mean(axis=1, keepdims=True)
    c = x[:k] @ x.T / (oe.shape[0] - 1)
    sd = np.sqrt(np.einsum("ij,ij->i", x, x) / (oe.shape[0] - 1))
    with np.errstate(divide="ignore", invalid="ignore"):
        want = np.clip(c / sd[:k, None] / sd[None, :], -1, 1)
    want = np.where(np.isnan(want), 0.0, want)
    res = [float(np.linalg.norm(corr @ evecs[i] - evals[i] * evecs[i])) for i in range(len(evals))]
    out["check"] = {"rows": k, "corr_max_abs_err": float(np.abs(corr[:k] - want).max()),
                    "symmetric": bool(np.array_equal(corr, corr.T)),
                    "eig_residuals": res, "eigenvalues": [float(v) for v in evals]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
