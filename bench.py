#!/usr/bin/env python
"""ICE throughput bench (contract: see the task description; one JSON line on rank 0).

    python bench.py --gpus 1 --steps K --warmup W            # our arm, 1 GPU
    torchrun ... bench.py --gpus N --steps K --warmup W       # our arm, N GPUs (row-sharded)
    python bench.py --impl reference ...                     # CPU arm: the oracle port

A step = one complete ICE normalisation (tb_ice: iterations=100, max_dev=1e-5) of the
resident contact store.  metric = contacts * passes / s, contacts = stored upper-triangle
cells, passes = row-sum evaluations executed (= verbose trace lines).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ICE contacts*iterations/s"
UNIT = "contacts*iterations/s"
ITERATIONS, MAX_DEV = 100, 1e-5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-bins", type=int, default=32768,
                    help="bins of the leading sub-matrix used for the CPU baseline sample")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for k, nm in enumerate(names):
                if len(r) > 2 + k and r[2 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")      # host plumbing only; the data path is NCCL in the library
    return rank, local, world


def cpu_baseline(wl, nbins, device, threads=None):
    """The oracle port on a bounded sample: the leading nbins x nbins block of the same
    synthetic model, generated on the host (oracle/synth_port.c) and balanced on the host --
    no GPU and no product code on this leg."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tadbit_oracle as orc
    nbins = min(nbins, wl["size"])
    r, c, v = orc.synthetic_block(nbins, wl["params"], wl["chromosomes"])
    m = orc.OracleMatrix(nbins, r, c, v)
    # pass loop in C + OpenMP (oracle/ice_port.c); only the loop is timed, the CSR build is not
    bias, trace, cores, dt = orc.iterative_fast(m, {}, ITERATIONS, MAX_DEV, threads)
    return {"value": r.size * len(trace) / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "leading %d x %d block of the %s model (%d contacts, %d passes, %.2f s "
                      "in the pass loop)" % (nbins, nbins, wl["name"], r.size, len(trace), dt)}


def run_reference(args, rank, world):
    """CPU arm: the reference is pure Python and cannot travel to the GPU box, so this times
    the oracle port (validated against the unmodified reference's golden vectors)."""
    if rank != 0:
        return
    from tadbit_b200 import workloads                    # workload definitions only (no library call)
    name = args.workload if args.workload != "auto" else "c3"
    wl = workloads.workload(name)
    vals = []
    for _ in range(max(1, args.warmup > 0) + args.steps):
        vals.append(cpu_baseline(wl, args.cpu_bins, 0, threads=os.cpu_count()))
    cb = vals[-1]
    value = float(np.mean([v["value"] for v in vals[1:]])) if len(vals) > 1 else cb["value"]
    cb = dict(cb, value=value)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["description"], "sample": cb["sample"]},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse()
    rank, local, world = dist_setup(args)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import tadbit_b200
    from tadbit_b200 import workloads
    if tadbit_b200.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; tadbit_b200 has no CPU fallback")
    device = local
    torch.cuda.set_device(device)
    name = args.workload
    if name == "auto":
        # the metric is quoted "at 1/2/4/8 B200" on the genome-wide 5 kb matrix (configs[2]); it
        # fits one GPU (~41 GB), so every N runs the same matrix (strong scaling).  --workload c2
        # gives the dense chr1 @10 kb case (configs[1]).
        name = "c3"
    wl = workloads.workload(name)

    gather = None
    if world > 1:
        import torch.distributed as dist

        def gather(arr):
            out = [None] * world
            dist.all_gather_object(out, arr)
            return out
    e2e_result = None
    if rank == 0 and world == 1 and not args.no_e2e:
        # end-to-end leg first, on a clean device: the matrix is generated once only to obtain
        # the host buffers a user would start from (pinned upper-triangle cells)
        tmp = workloads.build_sharded(wl, 0, 1, device, None)[0]
        cells = [torch.from_numpy(a).pin_memory() for a in tmp.export_upper()]
        tmp.close()
        e2e_result = e2e(cells, wl, device, args)
        del cells, tmp
    t_build = time.perf_counter()
    store, block = workloads.build_sharded(wl, rank, world, device, gather)
    if world > 1:
        uid = [tadbit_b200.ContactStore.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        store.comm_init(rank, world, uid[0])
    info = store.info()
    t_build = time.perf_counter() - t_build
    nnz_upper, nnz_full = info["nnz_upper"], info["nnz_full"]
    if world > 1:
        t = torch.tensor([nnz_upper, nnz_full], dtype=torch.int64)
        dist.all_reduce(t)
        nnz_upper, nnz_full = int(t[0].item()), int(t[1].item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        bias, trace = store.ice(None, iterations=ITERATIONS, max_dev=MAX_DEV)
        return bias, trace, store.ice_timing()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, rowsum_ms, rowsum_n, launches, passes = 0.0, 0.0, 0, 0, 0
    for _ in range(args.steps):
        bias, trace, tm = step()
        dev_ms += tm["loop_ms"]
        rowsum_ms += tm["rowsum_ms"]
        rowsum_n += tm["rowsum_launches"]
        launches += tm["total_launches"]
        passes += len(trace)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = t.tolist()
    value = nnz_upper * passes / (dev_ms * 1e-3)

    # roofline of the dominant kernel (K1 row sum), this rank's launches
    peak, peak_src = measured_peaks()
    n = wl["size"]
    # algorithmic bytes of THIS rank's launch (12 B per contact + 24 B per bin): the
    # rank walks info["nnz_full"] of the nnz_full directed cells, i.e. that share of the contacts
    alg_bytes = 12.0 * nnz_upper * info["nnz_full"] / max(nnz_full, 1) + 24.0 * (block[1] - block[0])
    k1_ms = rowsum_ms / max(rowsum_n, 1)
    rank_k1 = [k1_ms]
    if world > 1:                          # every rank's mean K1 time: shows the shard balance
        rank_k1 = [None] * world
        dist.all_gather_object(rank_k1, k1_ms)
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
    lay = store.layout()
    csr_walk = os.environ.get("TB_ROWSUM") == "csr"
    layout_bytes = 8.0 * info["nnz_full"] if csr_walk else float(lay["stream_bytes"])
    traffic = None                       # ncu dram bytes of one K1 evaluation (1 GPU capture)
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(wl["name"])
        if rec and world == 1 and not csr_walk:
            traffic = rec["k1_dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "note": "achieved = SURVEY 8d algorithmic bytes (12 B/contact + 24 B/bin) / K1 time; "
                        "the compact stream moves layout_bytes_per_launch (< algorithmic), so frac "
                        "can exceed the fraction of HBM peak actually drawn (layout_gbs / peak)",
                "kernel": "k_chunk_reduce<F64> (CSR walk)" if csr_walk else
                          "K1 row sum = k_rowsum_tiles + k_rowsum_dense (+ k_rowsum_sparse for the few "
                          "per-cell chunks), timed together per pass",
                "kernel_ms": k1_ms, "stream_encodings": lay["encodings"],
                "tile_cells": lay["tile_cells"], "tile_cells_padded": lay["tile_cells_padded"],
                "kernel_share_of_step": rowsum_ms / dev_ms if dev_ms else None,
                "algorithmic_bytes_per_launch": alg_bytes,
                "layout_bytes_per_launch": layout_bytes,
                "layout_gbs": layout_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0,
                # the same two rates against the B200 spec sheet (SURVEY 8d asks for both peaks)
                "peak_spec": 8000.0, "frac_spec": achieved / 8000.0}
    roofline["layout_frac"] = roofline["layout_gbs"] / peak if peak else None
    roofline["traffic_gbs"] = traffic / (k1_ms * 1e-3) / 1e9 if traffic and k1_ms > 0 else None

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["description"], "bins": n, "contacts_upper": nnz_upper,
                       "passes_per_step": passes // max(args.steps, 1),
                       "ice": "iterations=%d max_dev=%g" % (ITERATIONS, MAX_DEV),
                       "l2": "inputs exceed L2 (store %.1f GB vs 126 MB)" % (info["device_bytes"] / 1e9),
                       "sharding": "rows, nnz-balanced, 1 NCCL all-reduce/pass" if world > 1 else "single GPU",
                       "build_s": t_build},
            "wall_ms_per_step": wall_ms / args.steps,
            "roofline": roofline, "rank_kernel_ms": rank_k1, "gpu_launches": launches,
            "clocks": clocks}

    if e2e_result is not None:
        line["e2e"] = e2e_result
    store.close()
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(wl, args.cpu_bins, device)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def e2e(cells, wl, device, args):
    """Same metric through the public API with HOST buffers: pinned upper-triangle cells ->
    tb_store_create_coo (H2D + sort + layout build) -> tb_ice -> biases back on the host."""
    import torch
    import tadbit_b200
    pr, pc, pv = cells
    h2d = 3 * pr.numel() * 4
    times, passes = [], 0
    nrep = max(1, min(args.steps, 2))
    for it in range(1 + nrep):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hic = tadbit_b200.HiC_data.from_store(
            tadbit_b200.ContactStore.from_coo(wl["size"], pr.numpy(), pc.numpy(), pv.numpy(),
                                              chromosomes=wl["chromosomes"], device=device),
            chromosomes=wl["chromosomes"])
        t1 = time.perf_counter()
        hic.normalize_hic(iterations=ITERATIONS, max_dev=MAX_DEV, silent=True, factor=None)
        bias = hic.bias.to_array()
        dt = time.perf_counter() - t0
        if os.environ.get("TB_TIMING") == "1":
            sys.stderr.write("[bench e2e] create %.3f s, normalize_hic %.3f s\n" % (t1 - t0, dt - (t1 - t0)))
        npass = len(hic.last_trace)
        hic.store.close()
        if it:
            times.append(dt)
            passes = npass
    dt = float(np.mean(times))
    return {"value": pr.numel() * passes / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": int(bias.nbytes + passes * 32), "s_per_step": dt,
            "what": "pinned host COO -> tb_store_create_coo -> tb_ice -> host biases"}


if __name__ == "__main__":
    main()
