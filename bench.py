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
dw index like the reference's MPI layout (strong scaling), the Hdw term going through the
NCCL tile transpose.  One JSON line is printed by rank 0.
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
    return {"workload": f"cfg2: Norb=1 Nbath={ns - 1} (Ns={ns}) half-filled sector "
                        f"({dim} states), direct HxV, dw-sharded over {world} GPU(s)",
            "ns": ns, "dim": dim, "vector_gb": 8 * dim / 1e9}


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
    """Reference-faithful CPU path (oracle port of directMatVec_MPI_normal_main: per-element
    c/cdg + binary search, dw split, transposes) on `cores` threads = emulated MPI ranks, timed
    on a BOUNDED sample: P emulated ranks of which the first `cores` run (one per thread).
    Returns (hxv_per_s, description, seconds list)."""
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    ns = args.ns
    # each step is a bounded sample; size it so steps+warmup stay within a few minutes
    total = max(1, args.steps + args.warmup)
    target = max(2.0, min(15.0, 150.0 / total))
    vals, desc = [], ""
    t0 = time.perf_counter()
    for i in range(args.warmup + args.steps):
        r = cpu_reference_sample(ns, cores, target_s=target, repeats=1)
        desc = r[1]
        if i >= args.warmup:
            vals.append(r[0])
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    out = {
        "impl": "reference", "metric": "hxv_per_s", "value": value, "unit": "Hxv/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload(ns, max(args.gpus, 1)),
                       arm="reference CPU algorithm (directMatVec_MPI_normal_main port) on the host cores"),
        "cpu_baseline": {"value": value, "unit": "Hxv/s", "cores": cores, "kind": "port",
                         "sample": desc},
        "e2e": {"value": value, "unit": "Hxv/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Fortran reference cannot be built in this image (no gfortran/MPI/SciFortran); "
                "this is the C oracle port of its direct_mpi algorithm",
        "wall_s": wall,
    }
    print(json.dumps(out))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    import edipack_b200 as E
    from edipack_b200 import _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
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

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ns = args.ns
    nup = ndw = ns // 2
    model = E.EDModel(**model_kwargs(ns))
    if args.variant:
        E.set_kernel_variant(args.variant)
    E.build_Hv_sector_normal(model, nup, ndw)
    DimUp, DimDw, qdw, d0 = E.sector_dims()
    dim = DimUp * DimDw
    nloc = DimUp * qdw
    plen = int(L.edgpu_vec_padded_len())
    ldu = plen // qdw
    # synthetic input: i.i.d. N(0,1), fixed seed, pads zero (device memory via torch = plumbing)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    v = torch.randn((qdw, ldu), dtype=torch.float64, device=dev, generator=gen)
    v[:, DimUp:] = 0.0
    v /= math.sqrt(dim)
    hv = torch.zeros_like(v)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(L.edgpu_stream(), device=dev)

    def step():
        _abi.check(L.edgpu_hxv_dev(v.data_ptr(), hv.data_ptr()))

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: K H x v back to back, CUDA events on the launching stream ----------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    L.edgpu_launch_count(1)
    _abi.check(L.edgpu_profile_begin(args.steps))
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    ev1.synchronize()
    torch.cuda.synchronize()
    barrier()
    ms3 = (C.c_float * 3)()
    nrec = C.c_int()
    _abi.check(L.edgpu_profile_end(ms3, C.byref(nrec)))
    launches = int(L.edgpu_launch_count(0))
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step

    # ---- roofline of the dominant kernel + of the whole product ---------------------------
    peak, peak_kind = peaks()
    stage = [float(ms3[k]) / max(nrec.value, 1) for k in range(3)]  # ms per launch
    if world == 1:
        names = ["k_fastb(diag+up hops)", "k_slow(dw hops)", "k_nonlocal"]
    else:
        names = ["k_fast(diag+up hops) overlapped with transpose(v)+k_fast(dw hops on v^T)",
                 "join of the communication stream", "transpose(Hv^T)+accumulate"]
    # compulsory bytes per local state and launch (DESIGN.md "Kernels"): pass B k_fast reads v and
    # writes Hv (16 B); pass A k_slow reads v and read-modify-writes Hv (24 B)
    alg_bytes = [16.0 * ldu * qdw, 24.0 * ldu * qdw, 24.0 * ldu * qdw]
    dom = int(np.argmax(stage[:2] if world == 1 else stage))
    ach = alg_bytes[dom] / (stage[dom] * 1e-3) / 1e9 if stage[dom] > 0 else 0.0
    # dram__bytes_read+write per launch of that kernel from the committed ncu --set full capture
    # of the same workload (profiles/traffic.json), else null
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if tj.get("ns") == ns and world == 1:
            traffic = tj["kernels"][("k_fastb", "k_slow")[dom]]["dram_bytes_per_launch"]
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "peak_kind": peak_kind,
                "algorithmic_bytes_per_launch": alg_bytes[dom], "ms_per_launch": stage[dom],
                "limiter": "L1TEX/shared-memory data pipe (ncu l1tex__throughput ~90%), see profiles/"}
    hxv_ach = 16.0 * nloc / (ms_per_step * 1e-3) / 1e9
    hxv_roofline = {"bound": "hbm", "achieved": hxv_ach, "peak": peak, "unit": "GB/s",
                    "frac": hxv_ach / peak, "per_gpu_states": nloc,
                    "algorithmic_bytes_per_state": 16, "note": "whole HxV, 16 B/state (SURVEY 8d)"}
    kernels = [{"name": names[k], "ms": stage[k]} for k in range(3) if stage[k] > 0]

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
           "d2h_bytes_per_step": 8 * nloc, "steps": e2e_steps, "api": "edgpu_hxv_d (spHtimesV_p)"}
    del hin, hout

    # ---- GS Lanczos time-to-solution (device-resident vectors) ----------------------------
    lanczos = None
    if args.lanczos:
        del hv
        torch.cuda.empty_cache()
        # first solve: pays lazy module loading of the Lanczos-only kernel variants and the
        # allocation of the pooled Lanczos-vector buffers (reported, not the headline); the
        # second, identical solve is the warm time-to-solution every later solve of a DMFT run sees
        barrier()
        t0 = time.perf_counter()
        E.sp_lanc_eigh(args.lanczos_niter, 1e-12, want_vector=False)
        barrier()
        t_first = max_over_ranks(time.perf_counter() - t0)
        barrier()
        t0 = time.perf_counter()
        egs, _, nit = E.sp_lanc_eigh(args.lanczos_niter, 1e-12, want_vector=False)
        barrier()
        t_l = max_over_ranks(time.perf_counter() - t0)
        nstored, nhxv = E.lanczos_last_info()
        lanczos = {"egs": egs, "niter": nit, "seconds": t_l, "seconds_first_solve": t_first,
                   "hxv": nhxv,
                   "vectors_kept_in_hbm": nstored, "threshold": 1e-12,
                   "nitermax": args.lanczos_niter}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        val, desc, secs, O, m, vv = cpu_reference_sample(ns, cores, target_s=15.0)
        cpu = {"value": val, "unit": "Hxv/s", "cores": cores, "kind": "port", "sample": desc,
               "seconds": secs}
        try:
            # optimised CPU variant (stored hop tables, what ED_SPARSE_H=T does): full product
            _, t_st = O.stored_hxv_mpi(m, nup, ndw, vv, cores, cores, 2)
            cpu["stored_variant_hxv_per_s"] = 1.0 / t_st
        except Exception as ex:  # pragma: no cover
            cpu["stored_variant_error"] = str(ex)

    E.delete_Hv_sector_normal()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        out = {
            "metric": "hxv_per_s", "value": value, "unit": "Hxv/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": dict(workload(ns, world),
                           l2_policy=f"inputs larger than L2 ({8 * nloc / 1e9:.2f} GB local vector per HxV)",
                           kernel_variant=args.variant or 2,
                           kernels="two tiled passes: k_fastb (up block x 4 columns) + "
                                   "k_slow (16 rows x dw range)"),
            "roofline": roofline, "hxv_roofline": hxv_roofline, "kernels": kernels,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "lanczos_gs": lanczos,
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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
