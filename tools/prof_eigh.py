"""Minimal driver for profiling the sp_eigh replacement at BASELINE config 2:
    python tools/prof_eigh.py [neigen] [ncv] [ns]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import edipack_b200 as E

neigen = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ncv = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 16
E.ed_init(0)
m = E.EDModel(Norb=1, Nbath=ns - 1, Uloc=(2.0,), hfmode=True)
E.build_Hv_sector_normal(m, ns // 2, ns // 2)
t0 = time.perf_counter()
ev, _, nconv, nmv = E.sp_eigh(neigen, ncv, 512, 1e-12, want_vectors=False)
print(json.dumps({"neigen": neigen, "ncv": ncv, "ns": ns, "evals": list(ev), "nconv": nconv, "hxv": nmv,
                  "seconds": time.perf_counter() - t0}))
E.delete_Hv_sector_normal()
