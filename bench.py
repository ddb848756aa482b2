#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native EDIpack H x v engine.

Metric (BASELINE.json): FP64 H x v per second (+ achieved HBM GB/s) on the synthetic
single-band Anderson-impurity sector of BASELINE config 2: Norb=1, Nbath=15 (Ns=16),
half-filled sector (nup=ndw=8, 165 636 900 states, 1.325 GB per vector), direct H x v.

    python bench.py --gpus 1 --steps K --warmup W           # our arm
    python bench.py --impl reference --steps K --warmup W   # reference CPU algorithm (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one H x v on device-resident vectors (`value`); `e2e` is the same product through
the reference-facing `spHtimesV_p` C entry point (edgpu_hxv_d) with pinned HOST buffers, the
host<->device copies inside the timed region.  For N>1 the same sector is sharded along the
dw index like the reference's MPI layout (strong scaling); the Hdw term reads the few remote
columns it needs from a halo the owners push over NVLink (DESIGN.md "Multi-GPU"), and rank 0 also
checks one sharded product of a Ns=14 sector against the oracle (`parity_rel_err`).  With
--gpus 8 a second timed block runs BASELINE config 4 (Ns=18, key `cfg4`).  One JSON line is
printed by rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def model_kwargs(ns: int):
    """Synthetic Anderson impurity of SURVEY 8d: Norb=1, Nbath=ns-1, normal bath from
    init_dmft_bath (ed_hw_bath=2), Uloc=2, hfmode, xmu=0, Hloc=0."""
    return dict(Norb=1, Nbath=ns - 1, Uloc=(2.0,), hfmode=True, xmu=0.0)


def workload(ns: int, world: int = 1) -> dict:
    """`config` of both arms: BASELINE config 2 (or --ns): the half-filled sector of the
    single-band Anderson impurity with Nbath = ns-1, direct (on-the-fly) H x v."""
    dim = math.comb(ns, ns // 2) ** 2
    name = {16: "cfg2", 18: "cfg4"}.get(ns, "cfg2-family")
    return {"workload": f"{name}: Norb=1 Nbath={ns - 1} (Ns={ns}) half-filled sector "
                        f"({dim} states), direct HxV, dw-sharded over {world} GPU(s)",
            "ns": ns, "dim": dim, "vector_gb": 8 * dim / 1e9,
            "l2_policy": f"inputs larger than L2 ({8 * dim / world / 1e9:.2f} GB of vector per GPU and HxV)"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 "100", "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_sample(ns: int, cores: int, target_s: float = 15.0, repeats: int = 1):
    """The reference's NON-default direct path (ED_SPARSE_H=F; oracle port of
    directMatVec_MPI_normal_main: per-element c/cdg + binary search, dw split, transposes) on
    `cores` threads = emulated MPI ranks, timed on a BOUNDED sample: P emulated ranks of which
    the first `cores` run (one per thread), EXTRAPOLATED linearly to a whole product.
    Returns (hxv_per_s, description, seconds list, oracle module, model, vector)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import edipack_oracle as O

    m = O.Model(**model_kwargs(ns))
    nup = ndw = ns // 2
    du, dd = O.sector_dims(ns, nup, ndw)
    dim = du * dd
    per_core_states_per_s = 1.0 / (340e-9 * 8)  # measured on this image: ~340 ns/state on 8 cores
    want = per_core_states_per_s * target_s
    P = max(cores, int(math.ceil(dim / want)))
    P = min(P, dd, du)
    nrun = min(cores, P)
    v = O.start_vector(dim, 1234) - 0.5
    secs = []
    for _ in range(repeats):
        secs.append(O.direct_hxv_mpi_sample(m, nup, ndw, v, P, nrun, nrun))
    t = float(np.mean(secs))
    frac = nrun / P
    desc = (f"Ns={ns} sector ({nup},{ndw}), {P} emulated MPI ranks, {nrun} run "
            f"(fraction {frac:.4f} of one HxV), extrapolated linearly")
    return frac / t, desc, secs, O, m, v


def cpu_stored_products(ns: int, cores: int, ncalls: int):
    """The reference's DEFAULT path (ED_SPARSE_H=T, ED_INPUT_VARS.f90:664: stored H_up / H_dw
    tables + stored diagonal, spMatVec_mpi_normal_main) on all host cores as emulated MPI ranks:
    `ncalls` FULL products of the same sector; returns (seconds per product, oracle, model, v)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import edipack_oracle as O

    m = O.Model(**model_kwargs(ns))
    nup = ndw = ns // 2
    du, dd = O.sector_dims(ns, nup, ndw)
    v = O.start_vector(du * dd, 1234) - 0.5
    _, t = O.stored_hxv_mpi(m, nup, ndw, v, cores, cores, ncalls)
    return t, O, m, v


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm on this box's host cores.  `value` is its
    DEFAULT path (stored hop tables), every step one FULL product; the non-default direct path
    (what BASELINE config 2 names) is reported next to it from a bounded, extrapolated sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    ns = args.ns
    nup = ndw = ns // 2
    t0 = time.perf_counter()
    if args.warmup > 0:
        cpu_stored_products(ns, cores, args.warmup)
    t_step, O, m, v = cpu_stored_products(ns, cores, max(1, args.steps))
    value = 1.0 / t_step
    dval, ddesc, dsecs = None, None, None
    try:
        dval, ddesc, dsecs, _, _, _ = cpu_reference_sample(ns, cores, target_s=10.0)
    except Exception as ex:  # pragma: no cover
        ddesc = f"failed: {ex}"
    # GS Lanczos time-to-solution on the same operator: pass 1 of sp_lanc_eigh in C
    # (ora_stored_lanczos_gs), the whole solve with the same start vector and stopping rule as the
    # GPU arm's `lanczos_gs` (pass 2 -- the eigenvector -- would double it: SciFortran replays the
    # recurrence)
    lz = None
    try:
        egs, n_used, _, _, sec = O.stored_lanczos_gs(m, nup, ndw, O.start_vector(len(v), 4321),
                                                     args.lanczos_niter, 1e-12, P=cores, nthreads=cores)
        lz = {"egs": egs, "niter": n_used, "seconds": sec, "seconds_per_iteration": sec / max(n_used, 1),
              "threshold": 1e-12, "note": "pass 1 only (energies); stored-table operator on all cores"}
    except Exception as ex:  # pragma: no cover
        lz = {"error": str(ex)}
    wall = time.perf_counter() - t0
    sample = (f"Ns={ns} sector ({nup},{ndw}): {max(1, args.steps)} FULL products of the stored-table "
              f"path (ED_SPARSE_H=T, the reference default) on {cores} emulated MPI ranks / threads")
    out = {
        "impl": "reference", "metric": "hxv_per_s", "value": value, "unit": "Hxv/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload(ns, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "Hxv/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "Hxv/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "direct_variant": {"value": dval, "unit": "Hxv/s", "extrapolated": True, "sample": ddesc,
                           "seconds": dsecs,
                           "note": "ED_SPARSE_H=F (directMatVec_MPI_normal_main port): per-element "
                                   "c/cdg + binary search; not the reference's default"},
        "lanczos_gs": lz,
        "note": "Fortran reference cannot be built in this image (no gfortran/MPI/SciFortran); "
                "this is the C oracle port of its algorithms (oracle/ed_oracle.c)",
        "wall_s": wall,
    }
    print(json.dumps(out))
    return 0


def setup_engine():
    import torch
    import torch.distributed as dist

    import edipack_b200 as E
    from edipack_b200 import _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _abi.load()
    E.ed_init(local_rank)
    if world > 1:
        uid = [E.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        E.ed_set_comm(rank, world, uid[0])
    return torch, dist, E, _abi, L, world, rank, local_rank, dev


COMM_MODES = {0: "single rank", 1: "halo push over peer memory (NVLink stores + flags)",
              2: "chunk-pipelined peer-memory transposes", 3: "NCCL grouped send/recv transposes"}
NVLINK_GBS = 770.0  # measured peer copy per direction (B200_PROFILING.md)


def limiter_note(kernel: str):
    """What bounds `kernel`, from the committed ncu --set full summary (profiles/limiter.json,
    written by tools/ncu_summary.py), not a literal."""
    try:
        with open(os.path.join(ROOT, "profiles", "limiter.json")) as f:
            lj = json.load(f)
        return lj.get(kernel)
    except Exception:
        return None


def timed_products(ctx, ns, steps, warmup, variant=0):
    """Opens the half-filled sector of the Ns-site single-band model, times `steps` device-resident
    H x v (CUDA events on the launching stream, barrier + synchronize on both sides, max over
    ranks).  Leaves the sector open; returns a dict with the timing and the live vectors."""
    torch, dist, E, _abi, L, world, rank, local_rank, dev = ctx
    nup = ndw = ns // 2
    model = E.EDModel(**model_kwargs(ns))
    if variant:
        E.set_kernel_variant(variant)
    E.build_Hv_sector_normal(model, nup, ndw)
    DimUp, DimDw, qdw, d0 = E.sector_dims()
    dim = DimUp * DimDw
    nloc = DimUp * qdw
    plen = int(L.edgpu_vec_padded_len())
    ldu = plen // max(qdw, 1)
    # synthetic input: i.i.d. N(0,1), fixed seed, pads zero (device memory via torch = plumbing)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    v = torch.randn((qdw, ldu), dtype=torch.float64, device=dev, generator=gen)
    v[:, DimUp:] = 0.0
    v /= math.sqrt(dim)
    hv = torch.zeros_like(v)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(L.edgpu_stream(), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step():
        _abi.check(L.edgpu_hxv_dev(v.data_ptr(), hv.data_ptr()))

    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    L.edgpu_launch_count(1)
    _abi.check(L.edgpu_profile_begin(steps))
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    ev1.synchronize()
    torch.cuda.synchronize()
    barrier()
    ms3 = (C.c_float * 3)()
    nrec = C.c_int()
    _abi.check(L.edgpu_profile_end(ms3, C.byref(nrec)))
    launches = int(L.edgpu_launch_count(0))
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    stage = [float(ms3[k]) / max(nrec.value, 1) for k in range(3)]
    mode, halo_cols, send_cols, nchunks = E.sector_comm_info()
    return dict(model=model, nup=nup, ndw=ndw, DimUp=DimUp, DimDw=DimDw, qdw=qdw, d0=d0, dim=dim, nloc=nloc,
                ldu=ldu, v=v, hv=hv, ms_per_step=ms_total / steps, stage=stage, launches=launches,
                comm_mode=mode, halo_cols=halo_cols, send_cols=send_cols, nchunks=nchunks,
                barrier=barrier, max_over_ranks=max_over_ranks)


def rooflines(r, world, peak, peak_kind, ns):
    """roofline of the dominant kernel + of the whole product + the NVLink bound (N>1)."""
    stage, ldu, qdw, nloc = r["stage"], r["ldu"], r["qdw"], r["nloc"]
    halo_b = 8.0 * ldu * r["halo_cols"]
    if world == 1:
        names = ["k_fastc", "k_slow", "k_nonlocal"]
        desc = ["pass B: diagonal + up hops (TMA-staged tile)", "pass A: dw hops", "non-local terms"]
    else:
        names = ["k_fastc", "k_slow", "k_nonlocal"]
        desc = ["pass B: diagonal + up hops (the halo push runs beside it)",
                "wait for the halo flags + pass A: dw hops incl. halo gathers", "non-local terms"]
    # compulsory bytes per launch (DESIGN.md "Kernels"): pass B reads v and writes Hv (16 B/state);
    # pass A reads v, read-modify-writes Hv (24 B/state) and reads every halo column once
    alg_bytes = [16.0 * ldu * qdw, 24.0 * ldu * qdw + halo_b, 24.0 * ldu * qdw]
    # N=1: the slower of the two passes.  N>1: pass B, the one stage that is a kernel and nothing else
    # (stage 1 interleaves the waits for the halo flags with the row chunks of pass A: it is listed
    # under `kernels` and bounded by `nvlink`, not by HBM)
    dom = int(np.argmax(stage[:2])) if world == 1 else 0
    ach = alg_bytes[dom] / (stage[dom] * 1e-3) / 1e9 if stage[dom] > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if tj.get("ns") == ns and world == 1:
            traffic = tj["kernels"][names[dom]]["dram_bytes_per_launch"]
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": names[dom], "what": desc[dom], "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_kind": peak_kind,
                "algorithmic_bytes_per_launch": alg_bytes[dom], "ms_per_launch": stage[dom],
                "limiter": limiter_note(names[dom])}
    hxv_ach = 16.0 * nloc / (r["ms_per_step"] * 1e-3) / 1e9
    hxv_roofline = {"bound": "hbm", "achieved": hxv_ach, "peak": peak, "unit": "GB/s",
                    "frac": hxv_ach / peak, "per_gpu_states": nloc,
                    "algorithmic_bytes_per_state": 16, "note": "whole HxV, 16 B/state (SURVEY 8d)"}
    nvlink = None
    if world > 1:
        out_b = 8.0 * ldu * r["send_cols"]
        t_bound = max(out_b, halo_b) / (NVLINK_GBS * 1e9) * 1e3
        nvlink = {"bound": "nvlink", "mode": COMM_MODES.get(r["comm_mode"]), "bytes_out_per_rank": out_b,
                  "bytes_in_per_rank": halo_b, "peak": NVLINK_GBS, "unit": "GB/s",
                  "ms_at_peak": t_bound, "frac_of_step": t_bound / r["ms_per_step"],
                  "double_transpose_bytes_per_rank": 16.0 * ldu * qdw * (world - 1) / world,
                  "note": "rank 0's traffic; the reference's two vector_transpose_MPI would move "
                          "double_transpose_bytes_per_rank each way"}
    kernels = [{"name": names[k], "what": desc[k], "ms": stage[k]} for k in range(3) if stage[k] > 0]
    return roofline, hxv_roofline, nvlink, kernels


def cfg3_block(ctx, steps):
    """BASELINE config 3 (Norb=2, Nbath=6, Ns=14, sector (7,7), U=U'=2, J=Jx=Jp=0.125: the non-local
    spin-flip / pair-hopping terms) at the run's GPU count: device-resident H x v time (the 94 MB
    vector fits L2), parity of one product against the oracle, ground-state Lanczos."""
    torch, dist, E, _abi, L, world, rank, local_rank, dev = ctx
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import edipack_oracle as O

    hloc = np.zeros((2, 2, 2))
    hloc[0] = hloc[1] = np.diag([0.5, -0.5])
    kw = dict(Norb=2, Nbath=6, Uloc=(2.0, 2.0), Ust=2.0, Jh=0.125, Jx=0.125, Jp=0.125, hfmode=True, hloc=hloc)
    m, mo = E.EDModel(**kw), O.Model(**kw)
    nup = ndw = 7
    du, dd = O.sector_dims(14, nup, ndw)
    full = O.start_vector(du * dd, 3) - 0.5
    lo, hi = E.chunk_bounds(du, dd, world, rank)
    E.build_Hv_sector_normal(m, nup, ndw)
    try:
        hvh = E.spHtimesV_p(full[lo:hi].copy())
        plen = int(L.edgpu_vec_padded_len())
        v = torch.randn(plen, dtype=torch.float64, device=dev)
        _abi.check(L.edgpu_vec_upload(v.data_ptr(), np.ascontiguousarray(full[lo:hi]).ctypes.data))
        hv = torch.zeros_like(v)
        torch.cuda.synchronize()
        stream = torch.cuda.ExternalStream(L.edgpu_stream(), device=dev)
        for _ in range(3):
            _abi.check(L.edgpu_hxv_dev(v.data_ptr(), hv.data_ptr()))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            _abi.check(L.edgpu_hxv_dev(v.data_ptr(), hv.data_ptr()))
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / steps
        t0 = time.perf_counter()
        egs, _, nit = E.sp_lanc_eigh(300, 1e-12, want_vector=False)
        t_l = time.perf_counter() - t0
    finally:
        E.delete_Hv_sector_normal()
    nt = max(1, host_cores() // world)
    ref = O.stored_hxv_mpi(mo, nup, ndw, full, nt, nt)[0]
    err = float(np.abs(hvh - ref[lo:hi]).max() / np.abs(ref).max()) if hi > lo else 0.0
    t = torch.tensor([ms, err, t_l], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, err, t_l = (float(x) for x in t)
    peak, _ = peaks()
    return {"workload": "cfg3: Norb=2 Nbath=6 (Ns=14) sector (7,7), U=U'=2, J=Jx=Jp=0.125, "
                        f"{du * dd} states, dw-sharded over {world} GPU(s)",
            "ms_per_hxv": ms, "hxv_per_s": 1e3 / ms, "l2_policy": "94 MB vector: L2-resident, warm",
            "hbm_frac_at_16B_per_state": 16.0 * du * dd / world / (ms * 1e-3) / 1e9 / peak,
            "parity_rel_err": err, "lanczos_gs": {"egs": egs, "niter": nit, "seconds": t_l}}


def sharded_parity(ctx, ns=14):
    """N>1: one product of the dw-sharded Ns=14 half-filled sector against the oracle's stored
    product (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:236-375); max relative error over ranks."""
    torch, dist, E, _abi, L, world, rank, local_rank, dev = ctx
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import edipack_oracle as O

    kw = model_kwargs(ns)
    m, mo = E.EDModel(**kw), O.Model(**kw)
    nup = ndw = ns // 2
    du, dd = O.sector_dims(ns, nup, ndw)
    full = O.start_vector(du * dd, 99) - 0.5
    lo, hi = E.chunk_bounds(du, dd, world, rank)
    E.build_Hv_sector_normal(m, nup, ndw)
    try:
        hv = E.spHtimesV_p(full[lo:hi].copy())
    finally:
        E.delete_Hv_sector_normal()
    nt = max(1, host_cores() // world)
    ref = O.stored_hxv_mpi(mo, nup, ndw, full, nt, nt)[0]
    err = float(np.abs(hv - ref[lo:hi]).max() / np.abs(ref).max()) if hi > lo else 0.0
    t = torch.tensor([err], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_ours(args):
    ctx = setup_engine()
    torch, dist, E, _abi, L, world, rank, local_rank, dev = ctx
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    ns = args.ns
    peak, peak_kind = peaks()

    # ---- timed region: K H x v back to back ------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    r = timed_products(ctx, ns, args.steps, args.warmup, args.variant)
    clocks = sampler.stop() if rank == 0 else None
    barrier, max_over_ranks = r["barrier"], r["max_over_ranks"]
    ms_per_step = r["ms_per_step"]
    value = 1e3 / ms_per_step
    nloc, nup, ndw = r["nloc"], r["nup"], r["ndw"]
    roofline, hxv_roofline, nvlink, kernels = rooflines(r, world, peak, peak_kind, ns)
    launches = r["launches"]
    del r["v"]

    # ---- e2e: spHtimesV_p drop-in with pinned host buffers --------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    hin = torch.empty(nloc, dtype=torch.float64).pin_memory()
    hout = torch.empty(nloc, dtype=torch.float64).pin_memory()
    hin.copy_(torch.from_numpy(np.random.default_rng(1234 + rank).standard_normal(nloc)))
    n32 = C.c_int32(nloc)
    if nloc >= 2 ** 31:
        raise SystemExit("local chunk exceeds the reference's 32-bit Nloc")
    vin_p, hv_p = C.c_void_p(hin.data_ptr()), C.c_void_p(hout.data_ptr())
    L.edgpu_hxv_d(C.byref(n32), vin_p, hv_p)  # warm-up (allocates staging buffers)
    _abi.check(L.edgpu_status())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        L.edgpu_hxv_d(C.byref(n32), vin_p, hv_p)
    _abi.check(L.edgpu_status())
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e = {"value": 1.0 / e2e_s, "unit": "Hxv/s", "h2d_bytes_per_step": 8 * nloc,
           "d2h_bytes_per_step": 8 * nloc, "steps": e2e_steps, "api": "edgpu_hxv_d (spHtimesV_p)",
           "note": "one product per call: the download cannot start before the whole upload has "
                   "arrived (every output column reads input columns anywhere), so the call is "
                   "bound by 2 x 8 B/state over PCIe"}
    del hout
    # the same call on a plain (pageable) array a Fortran caller owns, page-locked once through
    # edgpu_host_register -- what INTEGRATION.md tells the shim to do per sector
    try:
        pin = np.random.default_rng(7 + rank).standard_normal(nloc)
        pout = np.empty(nloc)
        E.host_register(pin)
        E.host_register(pout)
        pin_p, pout_p = C.c_void_p(pin.ctypes.data), C.c_void_p(pout.ctypes.data)
        L.edgpu_hxv_d(C.byref(n32), pin_p, pout_p)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            L.edgpu_hxv_d(C.byref(n32), pin_p, pout_p)
        _abi.check(L.edgpu_status())
        barrier()
        e2e["registered_caller_arrays_hxv_per_s"] = 1.0 / max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        E.host_unregister(pin)
        E.host_unregister(pout)
        del pin, pout
    except Exception as ex:  # pragma: no cover
        e2e["registered_caller_arrays_error"] = str(ex)

    # ---- GS Lanczos time-to-solution --------------------------------------------------------
    lanczos = None
    e2e_lanczos = None
    if args.lanczos:
        del r["hv"]
        torch.cuda.empty_cache()
        # first solve: pays lazy module loading of the Lanczos-only kernel variants and the
        # allocation of the pooled Lanczos-vector buffers (reported, not the headline); the
        # second, identical solve is the warm time-to-solution every later solve of a DMFT run sees
        barrier()
        t0 = time.perf_counter()
        E.sp_lanc_eigh(args.lanczos_niter, 1e-12, want_vector=False)
        barrier()
        t_first = max_over_ranks(time.perf_counter() - t0)
        # three warm solves, the median is reported: single solves were seen to take up to 3x longer
        # with unchanged per-product kernel times (tools/diag_lanczos.py; DESIGN.md "Lanczos drivers")
        t_all = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            egs, _, nit = E.sp_lanc_eigh(args.lanczos_niter, 1e-12, want_vector=False)
            barrier()
            t_all.append(max_over_ranks(time.perf_counter() - t0))
        t_l = float(np.median(t_all))
        nstored, nhxv = E.lanczos_last_info()
        lanczos = {"egs": egs, "niter": nit, "seconds": t_l, "seconds_first_solve": t_first,
                   "hxv": nhxv, "hxv_per_s": nhxv / t_l, "seconds_all_warm_solves": t_all,
                   "vectors_kept_in_hbm": nstored, "threshold": 1e-12,
                   "nitermax": args.lanczos_niter}
        # the same solve through the reference-facing driver call with HOST buffers: start vector
        # in (pinned), E_gs + eigenvector out -- how sp_lanc_eigh is used behind ed_diag_d
        start = hin.numpy()
        barrier()
        t0 = time.perf_counter()
        egs2, vec, nit2 = E.sp_lanc_eigh(args.lanczos_niter, 1e-12, vect=start, inplace=True)
        barrier()
        t_e = max_over_ranks(time.perf_counter() - t0)
        _, nhxv2 = E.lanczos_last_info()
        e2e_lanczos = {"seconds": t_e, "egs": egs2, "niter": nit2, "hxv": nhxv2, "hxv_per_s": nhxv2 / t_e,
                       "h2d_bytes": 8 * nloc, "d2h_bytes": 8 * nloc,
                       "api": "edgpu_lanczos_gs (sp_lanc_eigh) with host start vector in, "
                              "E_gs + eigenvector out"}
        del vec
    del hin

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        try:
            t_st, O, m, vv = cpu_stored_products(ns, cores, 3)
            cpu = {"value": 1.0 / t_st, "unit": "Hxv/s", "cores": cores, "kind": "port",
                   "sample": f"Ns={ns} sector ({nup},{ndw}): 3 FULL products of the stored-table path "
                             f"(ED_SPARSE_H=T, the reference default) on {cores} emulated MPI ranks / threads"}
            dval, ddesc, dsecs, _, _, _ = cpu_reference_sample(ns, cores, target_s=10.0)
            cpu["direct_variant"] = {"value": dval, "extrapolated": True, "sample": ddesc, "seconds": dsecs}
        except Exception as ex:  # pragma: no cover
            cpu = {"value": None, "unit": "Hxv/s", "cores": cores, "kind": "port", "sample": f"failed: {ex}"}

    E.delete_Hv_sector_normal()

    # ---- N>1: sharded parity against the oracle ---------------------------------------------
    parity = None
    if world > 1 and not args.no_parity:
        parity = sharded_parity(ctx)

    # ---- BASELINE config 3 (Hund / spin-flip / pair-hopping) at this GPU count --------------
    cfg3 = None
    if not args.no_cfg3:
        try:
            cfg3 = cfg3_block(ctx, args.steps)
        except Exception as ex:  # pragma: no cover
            cfg3 = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- BASELINE config 4 (Ns=18 on 8 GPUs): second timed block ----------------------------
    cfg4 = None
    if (world == 8 and args.cfg4 != 0 and ns != 18) or args.cfg4 == 1:
        torch.cuda.empty_cache()
        E.release_cache()
        ns4 = args.cfg4_ns
        r4 = timed_products(ctx, ns4, max(5, args.steps // 2), args.warmup)
        rf4, hrf4, nv4, k4 = rooflines(r4, world, peak, peak_kind, ns4)
        # size-independent property where no oracle reaches: <x|Hy> = <Hx|y> on the sharded vectors
        x, hx = r4["v"], torch.zeros_like(r4["hv"])
        gen = torch.Generator(device=dev)
        gen.manual_seed(4321 + rank)
        y = torch.randn(x.shape, dtype=torch.float64, device=dev, generator=gen)
        y[:, r4["DimUp"]:] = 0.0
        y /= math.sqrt(r4["dim"])
        hy = torch.zeros_like(x)
        torch.cuda.synchronize()  # torch's fills run on torch's stream, the engine on its own
        _abi.check(L.edgpu_hxv_dev(x.data_ptr(), hx.data_ptr()))
        _abi.check(L.edgpu_hxv_dev(y.data_ptr(), hy.data_ptr()))
        torch.cuda.synchronize()
        d = torch.stack([(x * hy).sum(), (hx * y).sum(), (x * x).sum(), (hy * hy).sum()])
        if world > 1:
            dist.all_reduce(d)
        d = [float(t) for t in d]
        # |<x|Hy> - <Hx|y>| on the Cauchy-Schwarz scale |x| |Hy| (the dots of two random vectors
        # are themselves O(dim^-1/2) of it)
        sym = abs(d[0] - d[1]) / max(math.sqrt(d[2] * d[3]), 1e-300)
        del x, y, hx, hy, r4["v"], r4["hv"]
        torch.cuda.empty_cache()
        r4["barrier"]()
        t0 = time.perf_counter()
        egs4, _, nit4 = E.sp_lanc_eigh(args.lanczos_niter, 1e-12, want_vector=False)
        r4["barrier"]()
        t4 = r4["max_over_ranks"](time.perf_counter() - t0)
        nst4, nhxv4 = E.lanczos_last_info()
        E.delete_Hv_sector_normal()
        E.release_cache()
        cfg4 = {"workload": workload(ns4, world)["workload"], "dim": r4["dim"], "ms_per_hxv": r4["ms_per_step"],
                "hxv_per_s": 1e3 / r4["ms_per_step"],
                "aggregate_hbm_frac": hrf4["frac"], "hxv_roofline": hrf4, "roofline": rf4, "nvlink": nv4,
                "kernels": k4, "symmetry_defect": sym, "symmetry_dots": d[:2],
                "lanczos_gs": {"egs": egs4, "niter": nit4, "seconds": t4, "hxv": nhxv4,
                               "vectors_kept_in_hbm": nst4}}

    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        out = {
            "metric": "hxv_per_s", "value": value, "unit": "Hxv/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload(ns, world),
            "roofline": roofline, "hxv_roofline": hxv_roofline, "nvlink": nvlink, "kernels": kernels,
            "cpu_baseline": cpu, "e2e": e2e, "e2e_lanczos": e2e_lanczos, "gpu_launches": launches,
            "clocks": clocks, "lanczos_gs": lanczos, "parity_rel_err": parity, "cfg3": cfg3, "cfg4": cfg4,
            "notes": {"kernel_variant": args.variant or 2,
                      "kernels": "two tiled passes: k_fastb (up block x 4 columns) + k_slow (16 rows x "
                                 "dw range); N>1: halo push of remote dw columns beside pass B"},
        }
        print(json.dumps(out))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ns", type=int, default=16)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--lanczos", type=int, default=1)
    ap.add_argument("--lanczos-niter", type=int, default=300)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the N>1 sharded parity check")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the BASELINE config 3 block")
    ap.add_argument("--cfg4-ns", type=int, default=18, help="Ns of the cfg4 block (testing: smaller)")
    ap.add_argument("--cfg4", type=int, default=-1,
                    help="-1: run the Ns=18 block when --gpus 8; 0: never; 1: always")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
