"""Diagnostic: repeated ground-state solves at cfg 2 with clock sampling and per-product stage times.
    python tools/diag_lanczos.py
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import edipack_b200 as E  # noqa: E402
from edipack_b200 import _abi  # noqa: E402

L = _abi.load()
E.ed_init(0)
m = E.EDModel(**bench.model_kwargs(16))
E.build_Hv_sector_normal(m, 8, 8)
out = []
for rep in range(4):
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.2)
    _abi.check(L.edgpu_profile_begin(400))
    t0 = time.perf_counter()
    e, _, nit = E.sp_lanc_eigh(300, 1e-12, want_vector=False)
    dt = time.perf_counter() - t0
    ms3 = (C.c_float * 3)()
    nrec = C.c_int()
    _abi.check(L.edgpu_profile_end(ms3, C.byref(nrec)))
    ck = s.stop()
    out.append({"rep": rep, "seconds": dt, "niter": nit, "hxv_recorded": nrec.value,
                "mean_stage_ms": [ms3[k] / max(nrec.value, 1) for k in range(3)], "clocks": ck,
                "info": E.lanczos_last_info()})
    print(json.dumps(out[-1]), flush=True)
    if rep == 1:
        time.sleep(3.0)  # let the board cool: does the next solve run faster?
E.delete_Hv_sector_normal()
