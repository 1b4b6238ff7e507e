#!/usr/bin/env python
"""Build a synthetic workload on cuda:0 and run one short ICE call (the target of ncu captures of
the per-pass kernels).   python tools/run_ice_once.py c3 [iterations]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tadbit_b200 import ContactStore, workloads  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    its = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    wl = workloads.workload(name)
    store = ContactStore.synthetic(wl["size"], wl["params"], chromosomes=wl["chromosomes"])
    bias, trace = store.ice(None, its, 1e-12)
    print("passes", len(trace), store.timing())


if __name__ == "__main__":
    main()
