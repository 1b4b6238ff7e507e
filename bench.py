#!/usr/bin/env python
"""ICE throughput bench (contract: see the task description; one JSON line on rank 0).

    python bench.py --gpus 1 --steps K --warmup W            # our arm, 1 GPU
    torchrun ... bench.py --gpus N --steps K --warmup W       # our arm, N GPUs (row-sharded)
    python bench.py --impl reference ...                     # CPU arm: the oracle port (+ the
                                                             #   unmodified reference on a small block)

Pipeline measured (north_star): filter_columns() -> normalize_hic(iterations=100, max_dev=1e-5)
-> normalize_expected() on the resident contact store.  A timed STEP is one complete ICE
normalisation with the mask of filter_columns (tb_ice: K1 row sums + exchange/statistics/update per
pass); the filter and the distance sums run once, outside the timed steps, and their device
times are reported next to it (`pipeline`).  metric = contacts * passes / s with contacts =
stored upper-triangle cells and passes = row-sum evaluations executed (= verbose trace lines).
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
# iterations is a cap: the loop stops when the deviation drops below max_dev (north_star: converged to
# 1e-5).  The genome-wide 5 kb model needs more than the 100 of the reference's examples.
ITERATIONS, MAX_DEV = 500, 1e-5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-compartments", action="store_true")
    ap.add_argument("--parity-workload", default="c3s")
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
        dist.init_process_group("gloo")      # host plumbing only; the data path is NVLink / NCCL in the library
    return rank, local, world


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tadbit_oracle as orc
    return orc


# ---------------------------------------------------------------------------------------------
# CPU legs
# ---------------------------------------------------------------------------------------------
def cpu_baseline(wl, nbins, threads=None):
    """The oracle port on a bounded sample: the leading nbins x nbins block of the same
    synthetic model, generated on the host (oracle/synth_port.c) and balanced on the host --
    no GPU and no product code on this leg."""
    orc = _oracle()
    nbins = min(nbins, wl["size"])
    r, c, v = orc.synthetic_block(nbins, wl["params"], wl["chromosomes"])
    m = orc.OracleMatrix(nbins, r, c, v)
    # pass loop in C + OpenMP (oracle/ice_port.c); only the loop is timed, the CSR build is not
    bias, trace, cores, dt = orc.iterative_fast(m, {}, ITERATIONS, MAX_DEV, threads)
    return {"value": r.size * len(trace) / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "same_config": False, "seconds": dt,
            "sample": "leading %d x %d block of the %s model (%d contacts, %d passes, %.2f s "
                      "in the pass loop)" % (nbins, nbins, wl["name"], r.size, len(trace), dt)}


def reference_rate(wl, nbins=1500):
    """The UNMODIFIED reference (pure Python, single-threaded) on the leading nbins x nbins block:
    pytadbit.utils.normalize_hic.iterative through oracle/ref_shim.py, from /root/reference or the
    copy build() leaves under baseline/_ref.  None when neither is present."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_shim
        if not ref_shim.available():
            return None
        import contextlib
        import io
        orc = _oracle()
        r, c, v = orc.synthetic_block(nbins, wl["params"], wl["chromosomes"])
        with contextlib.redirect_stderr(io.StringIO()):
            ns = ref_shim.load()
            dico = {}
            for a, b, x in zip(r.tolist(), c.tolist(), v.tolist()):
                dico[a * nbins + b] = x
                dico[b * nbins + a] = x
            hic = ns.HiC_data(dico, nbins, dict_sec={})
            buf = io.StringIO()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(buf):
                ns.iterative(hic, bads={}, iterations=10, max_dev=MAX_DEV, verbose=True)
            dt = time.perf_counter() - t0
        passes = sum(1 for line in buf.getvalue().splitlines() if line.startswith("   "))
        return {"value": r.size * passes / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                "same_config": False,
                "sample": "unmodified pytadbit iterative() on the leading %d x %d block (%d contacts, "
                          "%d passes, %.2f s incl. copy_matrix)" % (nbins, nbins, r.size, passes, dt)}
    except Exception as exc:                                      # noqa: BLE001
        return {"unavailable": "%s: %s" % (type(exc).__name__, exc)}


def reference_compartments(wl, nbins=1000):
    """N4 beside the GPU numbers: the UNMODIFIED reference's HiC_data.find_compartments (nested Python
    lists, numpy.corrcoef, scipy eigsh) on the leading nbins x nbins block of the model as one
    chromosome, single-threaded Python apart from numpy's BLAS.  None when the reference is absent."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_shim
        if not ref_shim.available():
            return None
        import contextlib
        import io
        import warnings
        from collections import OrderedDict
        orc = _oracle()
        r, c, v = orc.synthetic_block(nbins, wl["params"], wl["chromosomes"])
        with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()), \
                warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns = ref_shim.load()
            dico = {}
            for a, b, x in zip(r.tolist(), c.tolist(), v.tolist()):
                dico[a * nbins + b] = x
                dico[b * nbins + a] = x
            chroms = OrderedDict([("chr1", nbins)])
            hic = ns.HiC_data(dico, nbins, chromosomes=chroms,
                              dict_sec=dict((("chr1", p), p) for p in range(nbins)), resolution=10000)
            hic.filter_columns(silent=True, by_mean=False)
            hic.normalize_hic(iterations=10, max_dev=1e-4, silent=True)
            hic.normalize_expected()
            t0 = time.perf_counter()
            hic.find_compartments()
            dt = time.perf_counter() - t0
        n = nbins - len(hic.bads)
        return {"seconds": dt, "good_bins": n, "us_per_cell": dt / max(n * n, 1) * 1e6, "kind": "reference",
                "sample": "unmodified pytadbit HiC_data.find_compartments on one chromosome of %d bins "
                          "(leading block of the %s model)" % (nbins, wl["name"])}
    except Exception as exc:                                      # noqa: BLE001
        return {"unavailable": "%s: %s" % (type(exc).__name__, exc)}


def run_reference(args, rank, world):
    """CPU arm.  The reference's ICE is pure Python over a dict (1.6-2.0e6 contacts*it/s) and cannot
    hold the benchmarked matrix, so the line's value is the oracle port (C + OpenMP, all host
    threads; validated against the unmodified reference's golden vectors) on a bounded block of
    the same model; the unmodified reference itself is timed on a small block and reported in
    `reference_python`."""
    if rank != 0:
        return
    from tadbit_b200 import workloads                    # workload definitions only (no library call)
    name = args.workload if args.workload != "auto" else "c3"
    wl = workloads.workload(name)
    vals = []
    for _ in range(max(1, args.warmup > 0) + args.steps):
        vals.append(cpu_baseline(wl, args.cpu_bins, threads=os.cpu_count()))
    cb = vals[-1]
    value = float(np.mean([v["value"] for v in vals[1:]])) if len(vals) > 1 else cb["value"]
    cb = dict(cb, value=value)
    step_ms = float(np.mean([v["seconds"] for v in (vals[1:] if len(vals) > 1 else vals)])) * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["description"], "sample": cb["sample"], "same_config": False},
            "cpu_baseline": cb, "reference_python": reference_rate(wl),
            "reference_python_compartments": reference_compartments(wl),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
_GROUP = []


def connect(store, rank, world, dist, device=0):
    """Attach a shard to the process's GPU group.  The group (NCCL communicator) is created once,
    like the torch.distributed process group the launcher sets up; attaching a store maps its
    row-sum exchange buffers on the peers (CUDA IPC)."""
    if world > 1:
        import tadbit_b200
        if not _GROUP:
            uid = [tadbit_b200.ContactStore.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            _GROUP.append(tadbit_b200.CommGroup(rank, world, uid[0], device=device))
        store.attach(_GROUP[0])


PARITY_ITERATIONS = 100          # bounded: the oracle runs the same passes on the host cores


def pipeline(hic):
    hic.filter_columns(silent=True)
    hic.normalize_hic(iterations=PARITY_ITERATIONS, max_dev=MAX_DEV, silent=True)
    hic.normalize_expected()
    return dict(bads=hic.bads, bias=hic.bias.to_array().copy(), expected=hic.expected,
                passes=len(hic.last_trace), dev=float(hic.last_trace[-1][3]) if len(hic.last_trace) else None)


def parity_block(args, rank, local, world, gather, dist):
    """Before anything is timed: the sharded pipeline on a reduced genome-wide matrix against the
    same pipeline on ONE GPU and against the oracle (oracle/upper_port.c on the exported cells),
    rank 0 judging.  This is what makes the N > 1 numbers below credible."""
    import tadbit_b200
    from tadbit_b200 import workloads
    wl = workloads.workload(args.parity_workload)
    store, block = workloads.build_sharded(wl, rank, world, local, gather, refine=0)
    connect(store, rank, world, dist, local)
    hic = tadbit_b200.HiC_data.from_store(store, chromosomes=wl["chromosomes"])
    res = pipeline(hic)
    exch = store.timing()["exchange"]
    out = None
    allres = gather(res) if world > 1 else [res]
    if rank == 0:
        same = all(r["bads"] == res["bads"] and np.array_equal(r["bias"], res["bias"]) and
                   r["expected"] == res["expected"] and r["passes"] == res["passes"] for r in allres)
        out = {"workload": wl["description"], "bins": wl["size"], "exchange": exch,
               "ranks_identical": bool(same), "passes": res["passes"], "final_dev": res["dev"],
               "bad_columns": len(res["bads"])}
        if world > 1:
            one = tadbit_b200.ContactStore.synthetic(wl["size"], wl["params"],
                                                     chromosomes=wl["chromosomes"], device=local)
        else:
            one = store
        out["contacts"] = one.info()["nnz_upper"]
        if world > 1:
            ref = pipeline(tadbit_b200.HiC_data.from_store(one, chromosomes=wl["chromosomes"]))
            out["vs_single_gpu"] = {
                "mask_equal": ref["bads"] == res["bads"], "passes_equal": ref["passes"] == res["passes"],
                "expected_equal": ref["expected"] == res["expected"],
                "bias_max_rel": float(np.max(np.abs(ref["bias"] - res["bias"]) / np.abs(ref["bias"])))}
        orc = _oracle()
        r, c, v = one.export_upper()
        m = orc.UpperCells(wl["size"], r, c, v, chromosomes=wl["chromosomes"])
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())      # torchrun sets it to 1
        obads = orc.fast_filter_columns(m, silent=True)
        obias, otrace = orc.fast_normalize_hic(m, obads, PARITY_ITERATIONS, MAX_DEV)
        oexp = orc.fast_expected(m, obads)
        out["vs_oracle"] = {
            "mask_equal": obads == res["bads"], "passes_equal": len(otrace) == res["passes"],
            "expected_equal": oexp == res["expected"],
            "bias_max_rel": float(np.max(np.abs(obias - res["bias"]) / np.abs(obias)))}
        checks = [out["ranks_identical"], out["vs_oracle"]["mask_equal"], out["vs_oracle"]["passes_equal"],
                  out["vs_oracle"]["expected_equal"], out["vs_oracle"]["bias_max_rel"] < 1e-10]
        if world > 1:
            g = out["vs_single_gpu"]
            checks += [g["mask_equal"], g["passes_equal"], g["expected_equal"], g["bias_max_rel"] < 1e-10]
        out["ok"] = bool(all(checks))
        if world > 1:
            one.close()
    store.close()
    if world > 1:
        dist.barrier()
    return out


def extraction_leg(hic, store, wl, window=2000):
    """N3, extraction side, once: the cells of a `window`-bin square on the diagonal of the largest
    chromosome (tb_region_cells with the norm and decay transforms) and the three .abc files of
    extraction.write_matrix for it (wall clock of the API calls)."""
    import tempfile
    from tadbit_b200 import extraction
    from tadbit_b200._capi import bad_mask
    n = len(hic)
    reso = hic.resolution
    sec = max(wl["chromosomes"], key=lambda c: wl["chromosomes"][c])
    beg, end = hic.section_pos[sec]
    w0 = beg + (end - beg) // 3
    w1 = min(w0 + window, end)
    bias = hic.bias.to_array()
    decay = np.ones(n)
    table = {}
    for crm, (b0, b1) in hic.section_pos.items():
        vals = [hic.expected.get(d, 1.0) or 1.0 for d in range(b1 - b0)]
        decay[b0:b1] = vals
        table[crm] = dict(enumerate(vals))
    t0 = time.perf_counter()
    cells = store.region_cells(w0, w1, w0, w1, bad=bad_mask(n, hic.bads), bias=bias, decay=decay)
    t_cells = time.perf_counter() - t0
    biases = dict(biases=hic.bias.to_dict(), badcol=dict((int(k), v) for k, v in hic.bads.items()),
                  decay=table, resolution=reso)
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        fnames = extraction.write_matrix(hic, reso, biases, tmp, normalizations=("raw", "norm", "decay"),
                                         region1=sec, start1=(w0 - beg) * reso, end1=(w1 - beg) * reso,
                                         verbose=False)
        t_write = time.perf_counter() - t0
        nbytes = sum(os.path.getsize(f) for f in fnames.values())
        nlines = sum(1 for _ in open(fnames["RAW"]))
    return {"section": sec, "window_bins": w1 - w0, "cells_full": int(cells[0].size),
            "region_cells_ms": t_cells * 1e3, "cells_per_s": cells[0].size / t_cells if t_cells else None,
            "write_matrix_ms": t_write * 1e3, "files": len(fnames), "lines_per_file": nlines, "bytes": nbytes,
            "what": "tb_region_cells (count, scan, place, radix sort, v / b_i / b_j / decay on the device) and "
                    "the raw_ / nrm_ / dec_ .abc files of the window (half matrix), lines formatted by "
                    "tb_write_abc_lines; the unmodified reference spends 3.96 us per cell on the same three "
                    "files after its BAM file is read (profiles/r02_n3_reference_cost_per_cell.json)"}


def e2e_leg(args, wl, store, rank, local, world, dist, mask):
    """Same metric through the public API with HOST buffers, every rank at once: pinned CSR-form
    cells (N = 1: the upper triangle of the matrix, 8 B per cell; N > 1: the rows each rank owns,
    pre-partitioned) -> tb_store_create_csr (H2D + layout build) -> tb_store_attach to the process's
    GPU group (created once, outside: it is the counterpart of the launcher's process group) ->
    tb_ice with the filter mask -> biases back on the host.  Wall clock, max over ranks."""
    import torch
    import tadbit_b200

    def pinned(count, dtype):
        t = torch.empty(int(count), dtype={np.int64: torch.int64, np.int32: torch.int32}[dtype]).pin_memory()
        return t.numpy()

    form = "upper" if world == 1 else "full_rows"
    info = store.info()
    block = (info["row_begin"], info["row_end"])
    cells = store.export_csr(form, pinned=pinned)
    store.close()                          # the resident store of the timed steps is no longer needed
    tadbit_b200._capi.trim_memory(local)
    h2d = sum(int(a.nbytes) for a in cells) + int(mask.nbytes)
    times, passes, bias = [], 0, None
    nrep = max(1, min(args.steps, 2))
    for it in range(1 + nrep):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        st = tadbit_b200.ContactStore.from_csr(wl["size"], cells[0], cells[1], cells[2], form=form,
                                               chromosomes=wl["chromosomes"],
                                               row_block=None if world == 1 else block, device=local)
        t1 = time.perf_counter()
        connect(st, rank, world, dist, local)
        t2 = time.perf_counter()
        bias, trace = st.ice(mask, iterations=ITERATIONS, max_dev=MAX_DEV)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        if os.environ.get("TB_TIMING") == "1":
            sys.stderr.write("[bench e2e] rank %d create %.3f s, connect %.3f s, ice %.3f s\n"
                             % (rank, t1 - t0, t2 - t1, time.perf_counter() - t2))
        st.close()
        if it:
            times.append(dt)
            passes = len(trace)
    del cells
    nnz = torch.tensor([info["nnz_upper"], h2d], dtype=torch.int64)
    if world > 1:
        dist.all_reduce(nnz)
    dt = float(np.mean(times))
    return {"value": int(nnz[0]) * passes / dt, "unit": UNIT, "h2d_bytes_per_step": int(nnz[1]),
            "d2h_bytes_per_step": int(world * (bias.nbytes + passes * 32)), "s_per_step": dt,
            "passes": passes,
            "what": "pinned host CSR (%s) -> tb_store_create_csr -> %stb_ice(filter mask) -> host biases, "
                    "all ranks, wall clock max over ranks" % (
                        "upper triangle" if world == 1 else "each rank's rows",
                        "" if world == 1 else "tb_store_attach -> ")}


def main():
    args = parse()
    rank, local, world = dist_setup(args)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import tadbit_b200
    from tadbit_b200 import workloads
    from tadbit_b200._capi import bad_mask
    if tadbit_b200.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; tadbit_b200 has no CPU fallback")
    device = local
    torch.cuda.set_device(device)
    name = args.workload
    if name == "auto":
        # the metric is quoted "at 1/2/4/8 B200" on the genome-wide 5 kb matrix (configs[2]); it
        # fits one GPU, so every N runs the same matrix (strong scaling).  --workload c2
        # gives the dense chr1 @10 kb case (configs[1]).
        name = "c3"
    wl = workloads.workload(name)
    dist = None
    gather = None
    if world > 1:
        import torch.distributed as dist

        def gather(arr):
            out = [None] * world
            dist.all_gather_object(out, arr)
            return out

    parity = None if args.no_parity else parity_block(args, rank, local, world, gather, dist)
    tadbit_b200._capi.trim_memory(device)

    t_build = time.perf_counter()
    store, block = workloads.build_sharded(wl, rank, world, device, gather)
    connect(store, rank, world, dist, device)
    info = store.info()
    t_build = time.perf_counter() - t_build
    nnz_upper, nnz_full = info["nnz_upper"], info["nnz_full"]
    if world > 1:
        t = torch.tensor([nnz_upper, nnz_full], dtype=torch.int64)
        dist.all_reduce(t)
        nnz_upper, nnz_full = int(t[0].item()), int(t[1].item())
    n = wl["size"]
    hic = tadbit_b200.HiC_data.from_store(store, chromosomes=wl["chromosomes"])

    # ---- the filter, once (K4 + the <=100-point host tail) ------------------------------------
    t0 = time.perf_counter()
    hic.filter_columns(silent=True, **wl["filter_kw"])
    filter_ms = (time.perf_counter() - t0) * 1e3
    k4_ms = store.timing()["kernel_ms"]                # the last masked column-sum pass
    mask = bad_mask(n, hic.bads)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        bias, trace = store.ice(mask, iterations=ITERATIONS, max_dev=MAX_DEV)
        return bias, trace, store.timing()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, rowsum_ms, exch_ms, rowsum_n, launches, passes = 0.0, 0.0, 0.0, 0, 0, 0
    trace = []
    for _ in range(args.steps):
        bias, trace, tm = step()
        dev_ms += tm["loop_ms"]
        rowsum_ms += tm["rowsum_ms"]
        exch_ms += tm["exchange_ms"]
        rowsum_n += tm["rowsum_launches"]
        launches += tm["total_launches"]
        passes += len(trace)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    exchange_kind = tm["exchange"]
    if world > 1:
        t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = t.tolist()
    value = nnz_upper * passes / (dev_ms * 1e-3)
    final_dev = float(trace[-1][3]) if len(trace) else None
    passes_per_step = passes // max(args.steps, 1)

    # ---- the rest of the pipeline, once: K3 rescale and K5 distance sums ------------------------
    t0 = time.perf_counter()
    hic.normalize_hic(iterations=ITERATIONS, max_dev=MAX_DEV, silent=True)   # ICE + K3 + factor
    normalize_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    hic.normalize_expected()
    expected_ms = (time.perf_counter() - t0) * 1e3
    k5_ms = store.timing()["kernel_ms"]

    # ---- roofline of the dominant kernel (K1 row sum), this rank's launches ---------------------
    peak, peak_src = measured_peaks()
    own_rows = block[1] - block[0]
    # algorithmic bytes of THIS rank's launch (12 B per contact + 24 B per bin): the
    # rank walks info["nnz_full"] of the nnz_full directed cells, i.e. that share of the contacts
    share = info["nnz_full"] / max(nnz_full, 1)
    alg_bytes = 12.0 * nnz_upper * share + 24.0 * own_rows
    k1_ms = rowsum_ms / max(rowsum_n, 1)
    rank_k1 = [k1_ms]
    if world > 1:                          # every rank's mean K1 time: shows the shard balance
        rank_k1 = [None] * world
        dist.all_gather_object(rank_k1, k1_ms)
    lay = store.layout()
    # bytes one K1 evaluation has to move with this layout: the compact stream once, the x windows /
    # panels the kernels stage in shared memory (counted once per bin here), the partials written
    layout_bytes = float(lay["stream_bytes"]) + 16.0 * n
    gbs = lambda b, ms: b / (ms * 1e-3) / 1e9 if ms and ms > 0 else 0.0      # noqa: E731
    traffic, traffic_note = None, "no ncu capture matches this layout"
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get("%s_n%d" % (wl["name"], world))
        if rec and abs(rec["stream_bytes"] - lay["stream_bytes"]) <= 0.001 * lay["stream_bytes"]:
            traffic = rec["k1_dram_bytes_per_launch"]
            traffic_note = "ncu dram__bytes_read+write of one K1 evaluation, %s" % rec.get("source", "profiles/")
    except Exception:
        pass
    probe = store.probe_rowsum(4)          # (total, tile kernel, dense kernels) alone, after the timed region
    kernels = [
        {"kernel": "k_rowsum_tiles", "ms": probe[1], "bytes": float(lay["tile_bytes"]),
         "gbs": gbs(lay["tile_bytes"], probe[1]), "frac": gbs(lay["tile_bytes"], probe[1]) / peak},
        {"kernel": "k_rowsum_dense (+ k_rowsum_sparse)", "ms": probe[2], "bytes": float(lay["chunk_bytes"]),
         "gbs": gbs(lay["chunk_bytes"], probe[2]), "frac": gbs(lay["chunk_bytes"], probe[2]) / peak}]
    achieved = gbs(traffic if traffic else layout_bytes, k1_ms)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "basis": "bytes moved per K1 evaluation: " + ("ncu DRAM traffic" if traffic else
                         "compact stream + 16 B per bin (layout; no matching ncu capture)"),
                "traffic_note": traffic_note,
                "frac_algorithmic": gbs(alg_bytes, k1_ms) / peak,
                "achieved_algorithmic": gbs(alg_bytes, k1_ms),
                "algorithmic_bytes_per_launch": alg_bytes,
                "layout_bytes_per_launch": layout_bytes,
                "note": "frac = bytes actually moved / K1 time / peak (a fraction of the HBM rate). "
                        "frac_algorithmic uses SURVEY 8d's 12 B per contact + 24 B per bin: the compact "
                        "stream moves fewer bytes than that, so it can exceed 1.",
                "kernel": "K1 row sum = k_rowsum_tiles + k_rowsum_dense (+ k_rowsum_sparse for the few "
                          "per-cell chunks), timed together per pass",
                "kernel_ms": k1_ms, "kernels": kernels, "stream_encodings": lay["encodings"],
                "tile_cells": lay["tile_cells"], "tile_cells_padded": lay["tile_cells_padded"],
                "kernel_share_of_step": rowsum_ms / dev_ms if dev_ms else None,
                "peak_spec": 8000.0, "frac_spec": achieved / 8000.0}
    # one-pass kernels, algorithmic bytes of SURVEY 8d for this rank's share
    k4_alg = 12.0 * nnz_upper * share + 16.0 * own_rows
    k5_alg = 12.0 * nnz_upper * share + 4.0 * own_rows + 16.0 * max(wl["chromosomes"].values())
    pipeline_rec = {
        "filter_columns_ms": filter_ms, "bad_columns": len(hic.bads),
        "k4_masked_col_sums": {"ms": k4_ms, "algorithmic_bytes": k4_alg, "moved_bytes": layout_bytes,
                               "frac_algorithmic": gbs(k4_alg, k4_ms) / peak, "frac": gbs(layout_bytes, k4_ms) / peak,
                               "what": "K1 over the compact stream with x = the good mask (exact integers in fp64)"},
        "normalize_hic_ms": normalize_ms, "normalize_expected_ms": expected_ms,
        "k5_diag_sums": {"ms": k5_ms, "algorithmic_bytes": k5_alg, "moved_bytes": float(lay["stream_bytes"]),
                         "frac_algorithmic": gbs(k5_alg, k5_ms) / peak,
                         "frac": gbs(lay["stream_bytes"], k5_ms) / peak,
                         "what": "cell visitor over the compact stream, int64 histogram"}}

    # ---- N4 (compartments) on the largest chromosome, once: obs/exp scatter, fp64 X X^T on the
    # tensor pipe (DMMA), Lanczos eigen-solver.  A dense contraction: its roofline is flop, not bytes.
    if world == 1 and not args.no_compartments:
        try:
            sec = max(wl["chromosomes"], key=lambda c: wl["chromosomes"][c])
            beg, end = hic.section_pos[sec]
            exp_arr = np.array([hic.expected.get(d, 1.0) or 1.0 for d in range(end - beg)])
            t0 = time.perf_counter()
            n_good = store.corr_build(beg, end, hic.bias.to_array(), bad_mask(n, hic.bads), exp_arr, stage=1)
            evals, _, einfo = store.corr_eig(3)
            cmp_wall = (time.perf_counter() - t0) * 1e3
            ct = store.corr_timing()
            ld = -(-n_good // 16) * 16
            ntile = -(-n_good // 128)
            flop = ntile * (ntile + 1) / 2 * 128 * 128 * ld * 2.0
            pipeline_rec["n4_compartments"] = {
                "section": sec, "good_bins": n_good, "wall_ms": cmp_wall,
                "obs_exp_visitor_ms": ct["visit_ms"], "center_ms": ct["center_ms"],
                "syrk": {"kernel": "k_syrk_dmma", "ms": ct["gemm_ms"], "flop": flop,
                         "tflops": flop / (ct["gemm_ms"] * 1e-3) / 1e12 if ct["gemm_ms"] else None,
                         "peak_tflops_nominal_fp64": 40.0,
                         "frac": flop / (ct["gemm_ms"] * 1e-3) / 1e12 / 40.0 if ct["gemm_ms"] else None,
                         "bound": "fp64 tensor pipe (mma.sync m8n8k4 f64; tcgen05 has no fp64 kind)"},
                "scale_ms": ct["scale_ms"],
                "eigen": {"ms": ct["eig_ms"], "lanczos_vectors": einfo["steps"], "restarts": einfo["restarts"],
                          "eigenvalues": [float(x) for x in evals]}}
        except Exception as exc:                           # never lose the headline over the extra
            pipeline_rec["n4_compartments"] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if world == 1 and not args.no_compartments:
        try:
            pipeline_rec["n3_extraction"] = extraction_leg(hic, store, wl)
        except Exception as exc:                           # never lose the headline over the extra
            pipeline_rec["n3_extraction"] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["description"], "bins": n, "contacts_upper": nnz_upper,
                       "passes_per_step": passes_per_step,
                       "ice": "iterations=%d max_dev=%g, mask of filter_columns(%s)" % (
                           ITERATIONS, MAX_DEV, ", ".join("%s=%s" % kv for kv in wl["filter_kw"].items())),
                       "l2": "inputs exceed L2 (store %.1f GB vs 126 MB)" % (info["device_bytes"] / 1e9),
                       "sharding": ("rows, balanced by measured K1 time; per pass one fused kernel stores the "
                                    "row sums into every GPU (%s)" % exchange_kind) if world > 1 else "single GPU",
                       "build_s": t_build, "store_gb": info["device_bytes"] / 1e9},
            "converged": bool(final_dev is not None and final_dev < MAX_DEV), "final_dev": final_dev,
            "passes_per_step": passes_per_step,
            "wall_ms_per_step": wall_ms / args.steps,
            "ms_per_pass": {"total": dev_ms / max(passes, 1), "row_sum": rowsum_ms / max(passes, 1),
                            "exchange_statistics_update": exch_ms / max(passes, 1)},
            "roofline": roofline, "pipeline": pipeline_rec, "rank_kernel_ms": rank_k1,
            "gpu_launches": launches, "clocks": clocks, "parity": parity}

    if not args.no_e2e:
        e2e = e2e_leg(args, wl, store, rank, local, world, dist, mask)
        line["e2e"] = e2e
    store.close()
    for g in _GROUP:
        g.close()
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(wl, args.cpu_bins)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
