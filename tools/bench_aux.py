"""Secondary measurements (not the headline bench): BASELINE configs 3 and 5 on one B200.
  cfg3: Norb=2 Nbath=6 (Ns=14), U=U'=2, J=Jx=Jp=0.125 -> H x v incl. the non-local kernel
  cfg5-shaped stored path: complex Hermitian CSR with 705 432 rows and ~50 nnz/row (the shape of
        nonsu2 Norb=3, hybrid Nbath=8, sector N=11; synthetic entries) -> SpMV GB/s
    python tools/bench_aux.py
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import edipack_b200 as E
from edipack_b200 import _abi

L = _abi.load()
E.ed_init(0)
PEAK = 6544.7
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def time_hxv(v, hv, steps=20, flush=None):
    for _ in range(3):
        _abi.check(L.edgpu_hxv_dev(v.data_ptr(), hv.data_ptr()))
    torch.cuda.synchronize()
    st = torch.cuda.ExternalStream(L.edgpu_stream())
    tot = 0.0
    for _ in range(steps):
        if flush is not None:
            flush.zero_()  # evict L2 (256 MB write) between timed products
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        _abi.check(L.edgpu_hxv_dev(v.data_ptr(), hv.data_ptr()))
        e1.record(st)
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def cfg3():
    hloc = np.zeros((2, 2, 2))
    hloc[0] = hloc[1] = np.diag([0.5, -0.5])
    m = E.EDModel(Norb=2, Nbath=6, Uloc=(2.0, 2.0), Ust=2.0, Jh=0.125, Jx=0.125, Jp=0.125, hfmode=True,
                  hloc=hloc)
    E.build_Hv_sector_normal(m, 7, 7)
    n = int(L.edgpu_vec_padded_len())
    DimUp, DimDw, qdw, d0 = E.sector_dims()
    v = torch.randn(n, dtype=torch.float64, device="cuda")
    v.view(qdw, -1)[:, DimUp:] = 0
    hv = torch.zeros_like(v)
    flush = torch.empty(32 * 1024 * 1024, dtype=torch.float64, device="cuda")
    warm = time_hxv(v, hv)
    ms = (C.c_float * 4)()
    _abi.check(L.edgpu_hxv_dev(v.data_ptr(), hv.data_ptr()))
    L.edgpu_last_hxv_stage_ms(ms)
    cold = time_hxv(v, hv, steps=10, flush=flush)
    dim = DimUp * DimDw
    t0 = time.perf_counter()
    e, _, nit = E.sp_lanc_eigh(300, 1e-12, want_vector=False)
    tl = time.perf_counter() - t0
    E.delete_Hv_sector_normal()
    print(json.dumps({"config": "cfg3 Norb=2 Nbath=6 Ns=14 sector (7,7) with Jx/Jp", "dim": dim,
                      "ms_per_hxv_warm_L2": warm, "ms_per_hxv_L2_flushed": cold,
                      "stage_ms(k_fast,k_slow,k_nonlocal)": [ms[0], ms[1], ms[2]],
                      "GBps_at_16B_per_state_flushed": 16.0 * dim / cold / 1e6,
                      "frac_of_measured_hbm": 16.0 * dim / cold / 1e6 / PEAK,
                      "lanczos_gs": {"egs": e, "niter": nit, "seconds": tl}}))


def csr_cfg5_shape():
    n, per_row = 705432, 25  # 25 strictly-upper entries per row -> ~51 nnz/row after symmetrising
    rng = np.random.default_rng(1)
    i = np.repeat(np.arange(n, dtype=np.int64), per_row)
    # hop-like locality: targets within a window like the sorted-Fock-order neighbours
    j = (i + rng.integers(1, 60000, size=i.size)) % n
    val = rng.standard_normal(i.size) + 1j * rng.standard_normal(i.size)
    import scipy.sparse as sp

    A = sp.coo_matrix((val, (i, j)), shape=(n, n)).tocsr()
    H = (A + A.conj().T + sp.diags(rng.standard_normal(n))).tocsr()
    H.sum_duplicates()
    nnz = H.nnz
    E.build_Hv_sector_csr(H.indptr.astype(np.int64), (H.indices + 1).astype(np.int32), H.data)
    npad = int(L.edgpu_vec_padded_len())
    v = torch.randn(npad, dtype=torch.float64, device="cuda")
    v[2 * n:] = 0
    hv = torch.zeros_like(v)
    flush = torch.empty(32 * 1024 * 1024, dtype=torch.float64, device="cuda")
    warm = time_hxv(v, hv)
    cold = time_hxv(v, hv, steps=10, flush=flush)
    # parity spot check against scipy on the host
    x = (v[: 2 * n].cpu().numpy()).view(np.complex128)
    ref = H @ x
    got = hv[: 2 * n].cpu().numpy().view(np.complex128)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    E.delete_Hv_sector_csr()
    alg = nnz * (16 + 4) + n * (16 + 16) + (n + 1) * 8
    print(json.dumps({"config": "cfg5-shaped complex CSR SpMV (synthetic entries)", "rows": n, "nnz": int(nnz),
                      "ms_warm_L2": warm, "ms_L2_flushed": cold, "algorithmic_bytes": alg,
                      "GBps_flushed": alg / cold / 1e6, "frac_of_measured_hbm": alg / cold / 1e6 / PEAK,
                      "GBps_warm": alg / warm / 1e6, "rel_err_vs_scipy": float(err)}))


def cfg5_device_built():
    """BASELINE config 5 itself: ed_mode=nonsu2, Norb=3, hybrid bath Nbath=8 (Ns=11, 22 levels),
    on-site spin-orbit coupling lambda L.S + crystal field (complex Hermitian Hloc with spin-flip
    terms), U=2, U'=1.5, J=0.25, sector N=11 (705 432 states).  Sector map and stored H are
    generated on the device (edgpu_sector_open_nonsu2)."""
    lam, cf = 0.2, 0.1
    # t2g-like L=1 matrices in the (yz, zx, xy) basis and spin-1/2 matrices
    Lx = np.array([[0, 0, 0], [0, 0, -1j], [0, 1j, 0]])
    Ly = np.array([[0, 0, 1j], [0, 0, 0], [-1j, 0, 0]])
    Lz = np.array([[0, -1j, 0], [1j, 0, 0], [0, 0, 0]])
    sx = 0.5 * np.array([[0, 1], [1, 0]])
    sy = 0.5 * np.array([[0, -1j], [1j, 0]])
    sz = 0.5 * np.array([[1, 0], [0, -1]])
    hloc = np.zeros((2, 2, 3, 3), complex)
    for s in range(2):
        for t in range(2):
            hloc[s, t] = lam * (Lx * sx[s, t] + Ly * sy[s, t] + Lz * sz[s, t])
        hloc[s, s] += np.diag([cf, 0.0, -cf])
    m = E.EDModelNonsu2(Norb=3, Nbath=8, bath_type="hybrid", Uloc=(2.0, 2.0, 2.0), Ust=1.5, Jh=0.25,
                        Jx=0.25, Jp=0.25, hfmode=True, hloc=hloc)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    E.build_Hv_sector_nonsu2(m, 11)
    t_build = time.perf_counter() - t0
    n = E.vecDim_Hv_sector_normal()
    nnz = int(L.edgpu_csr_nnz())
    npad = int(L.edgpu_vec_padded_len())
    v = torch.randn(npad, dtype=torch.float64, device="cuda")
    v[2 * n:] = 0
    hv = torch.zeros_like(v)
    flush = torch.empty(32 * 1024 * 1024, dtype=torch.float64, device="cuda")
    warm = time_hxv(v, hv)
    cold = time_hxv(v, hv, steps=10, flush=flush)
    # Hermiticity as a size-independent property: <x|Hy> = conj(<y|Hx>)
    x = torch.randn(npad, dtype=torch.float64, device="cuda")
    x[2 * n:] = 0
    hx = torch.zeros_like(x)
    _abi.check(L.edgpu_hxv_dev(x.data_ptr(), hx.data_ptr()))
    _abi.check(L.edgpu_hxv_dev(v.data_ptr(), hv.data_ptr()))
    torch.cuda.synchronize()
    xc, vc = torch.view_as_complex(x[: 2 * n].view(-1, 2)), torch.view_as_complex(v[: 2 * n].view(-1, 2))
    hxc, hvc = torch.view_as_complex(hx[: 2 * n].view(-1, 2)), torch.view_as_complex(hv[: 2 * n].view(-1, 2))
    herm = abs(complex(torch.vdot(xc, hvc)) - complex(torch.vdot(vc, hxc)).conjugate()) / abs(complex(torch.vdot(xc, hvc)))
    t0 = time.perf_counter()
    e_l, _, nit = E.sp_lanc_eigh(512, 1e-12, want_vector=False)
    t_l = time.perf_counter() - t0
    t0 = time.perf_counter()
    ev, _, nconv, nmv = E.sp_eigh(2, 20, 512, 1e-12, want_vectors=False)
    t_e = time.perf_counter() - t0
    E.delete_Hv_sector_nonsu2()
    # the same sector with ED_SPARSE_H=F: direct on-the-fly product, nothing stored but the map
    E.set_sparse_H(False)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    t0 = time.perf_counter()
    E.build_Hv_sector_nonsu2(m, 11)
    t_build_d = time.perf_counter() - t0
    free1 = torch.cuda.mem_get_info()[0]
    hd = torch.zeros_like(v)
    warm_d = time_hxv(v, hd)
    cold_d = time_hxv(v, hd, steps=10, flush=flush)
    torch.cuda.synchronize()
    err_d = float((hd[: 2 * n] - hv[: 2 * n]).abs().max() / hv[: 2 * n].abs().max())
    E.delete_Hv_sector_nonsu2()
    E.set_sparse_H(True)
    direct = {"ms_warm_L2": warm_d, "ms_L2_flushed": cold_d, "build_seconds": t_build_d,
              "device_bytes_held": int(free0 - free1), "stored_form_bytes": int(nnz * 20 + (n + 1) * 8 + n * 4),
              "rel_diff_vs_stored_product": err_d}
    alg = nnz * (16 + 4) + n * (16 + 16) + (n + 1) * 8
    print(json.dumps({"config": "cfg5 nonsu2 Norb=3 hybrid Nbath=8 SOC, sector N=11, device-built spH0",
                      "direct_ED_SPARSE_H_F": direct,
                      "rows": n, "nnz": nnz, "nnz_per_row": nnz / n, "build_seconds(map+count+fill)": t_build,
                      "ms_warm_L2": warm, "ms_L2_flushed": cold, "algorithmic_bytes": alg,
                      "GBps_flushed": alg / cold / 1e6, "frac_of_measured_hbm": alg / cold / 1e6 / PEAK,
                      "hermiticity_defect": float(herm),
                      "lanczos_gs": {"egs": e_l, "niter": nit, "seconds": t_l},
                      "sp_eigh(neigen=2,ncv=20)": {"evals": list(ev), "nconv": nconv, "hxv": nmv, "seconds": t_e}}))


def cfg2_eigh():
    """sp_eigh (thick-restart Lanczos, ARPACK's role) at BASELINE config 2 next to the plain
    Lanczos ground state: Neigen=1 with the reference's default basis Nblock = 10*max(Neigen,2)."""
    m = E.EDModel(Norb=1, Nbath=15, Uloc=(2.0,), hfmode=True)
    E.build_Hv_sector_normal(m, 8, 8)
    out = {"config": "cfg2 Ns=16 sector (8,8), 165636900 states"}
    E.sp_lanc_eigh(300, 1e-12, want_vector=False)
    t0 = time.perf_counter()
    e_l, _, nit = E.sp_lanc_eigh(300, 1e-12, want_vector=False)
    out["sp_lanc_eigh"] = {"egs": e_l, "niter": nit, "seconds_warm_pool": time.perf_counter() - t0}
    E.release_cache()
    for neigen, ncv in ((1, 20), (2, 20)):
        t0 = time.perf_counter()
        ev, _, nconv, nmv = E.sp_eigh(neigen, ncv, 512, 1e-12, want_vectors=False)
        out[f"sp_eigh(neigen={neigen},ncv={ncv})"] = {"evals": list(ev), "nconv": nconv, "hxv": nmv,
                                                       "seconds": time.perf_counter() - t0}
    E.delete_Hv_sector_normal()
    print(json.dumps(out))


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg3", "csr", "cfg5", "eigh"]
    if "cfg3" in which:
        cfg3()
    if "csr" in which:
        csr_cfg5_shape()
    if "cfg5" in which:
        cfg5_device_built()
    if "eigh" in which:
        cfg2_eigh()
