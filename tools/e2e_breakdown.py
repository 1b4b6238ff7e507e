"""Wall time of `tb_store_create_coo` (stage by stage, TB_TIMING=1), `tb_ice` and destruction
for one synthetic workload whose cells sit in pinned host memory: python tools/e2e_breakdown.py c3"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tadbit_b200
from tadbit_b200 import workloads
wl = workloads.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
s = tadbit_b200.ContactStore.synthetic(wl["size"], wl["params"], chromosomes=wl["chromosomes"])
cells = [torch.from_numpy(a).pin_memory() for a in s.export_upper()]
s.close()
os.environ["TB_TIMING"] = "1"
for rep in range(3):
    t0 = time.perf_counter()
    st = tadbit_b200.ContactStore.from_coo(wl["size"], *[c.numpy() for c in cells], chromosomes=wl["chromosomes"])
    t1 = time.perf_counter()
    b, tr = st.ice(None, 100, 1e-5)
    t2 = time.perf_counter()
    st.close()
    t3 = time.perf_counter()
    print("create %.3f s  ice %.3f s  close %.3f" % (t1 - t0, t2 - t1, t3 - t2), file=sys.stderr)
