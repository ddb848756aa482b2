"""Minimal driver for profiling: opens the cfg sector and runs a few device-resident H x v.
    python tools/prof_hxv.py --ns 16 --variant 2 --steps 3
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import edipack_b200 as E
from edipack_b200 import _abi

ap = argparse.ArgumentParser()
ap.add_argument("--ns", type=int, default=16)
ap.add_argument("--norb", type=int, default=1)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--lanczos", type=int, default=0)
a = ap.parse_args()
L = _abi.load()
E.ed_init(0)
if a.norb == 1:
    kw = dict(Norb=1, Nbath=a.ns - 1, Uloc=(2.0,), hfmode=True)
else:
    import numpy as np
    hloc = np.zeros((2, 2, 2)); hloc[0] = hloc[1] = np.diag([0.5, -0.5])
    kw = dict(Norb=2, Nbath=a.ns // 2 - 1, Uloc=(2.0, 2.0), Ust=2.0, Jh=0.125, Jx=0.125, Jp=0.125,
              hfmode=True, hloc=hloc)
m = E.EDModel(**kw)
E.set_kernel_variant(a.variant)
E.build_Hv_sector_normal(m, a.ns // 2, a.ns // 2)
n = int(L.edgpu_vec_padded_len())
DimUp, DimDw, qdw, d0 = E.sector_dims()
v = torch.randn(n, dtype=torch.float64, device="cuda")
v.view(qdw, -1)[:, DimUp:] = 0
hv = torch.zeros_like(v)
torch.cuda.synchronize()
for _ in range(a.steps):
    _abi.check(L.edgpu_hxv_dev(v.data_ptr(), hv.data_ptr()))
ms = (C.c_float * 4)()
L.edgpu_last_hxv_stage_ms(ms)
print("stage ms", list(ms), "checksum", float(hv.sum()))
if a.lanczos:
    e, _, nit = E.sp_lanc_eigh(a.lanczos, 1e-12, want_vector=False)
    print("lanczos", e, nit)
E.delete_Hv_sector_normal()
