"""world_size-2 gloo tests (CPU) of the host-side N>1 logic: the dw-column split and the
scatter / gather / allgather vector movers that mirror ED_AUX_FUNX.f90:598,742,840, checked
against the oracle's emulated-MPI split on the same sector."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ns, nup, ndw, q):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import edipack_b200.host as H
        import edipack_oracle as O

        du, dd = O.sector_dims(ns, nup, ndw)
        full = O.start_vector(du * dd, 5) - 0.5
        mine = H.scatter_vector_MPI(full if rank == 0 else None, du, dd, root=0)
        lo, hi = H.chunk_bounds(du, dd, world, rank)
        ok = np.array_equal(mine, full[lo:hi])
        # the split is the reference's: first (DimDw mod P) ranks get one more column
        qd, d0 = H.mpi_split(dd, world, rank)
        ok = ok and (lo, hi) == (d0 * du, (d0 + qd) * du)
        ok = ok and qd == dd // world + (1 if rank < dd % world else 0)
        back = H.allgather_vector_MPI(2.0 * mine, du, dd)
        ok = ok and np.array_equal(back, 2.0 * full)
        g = H.gather_vector_MPI(mine, du, dd, root=1)
        ok = ok and ((g is None) if rank != 1 else np.array_equal(g, full))
        # phonons: DimPh slices, every rank's chunk = its dw columns of every slice
        # (i_el + (iph-1)*DimUp*mpiQdw, direct_mpi/HxV_eph.f90:3-4)
        nph = 3
        fullp = O.start_vector(du * dd * nph, 9) - 0.5
        minep = H.scatter_vector_MPI(fullp if rank == 0 else None, du, dd, root=0, DimPh=nph)
        expect = np.concatenate([fullp[k * du * dd + lo:k * du * dd + hi] for k in range(nph)])
        ok = ok and np.array_equal(minep, expect)
        ok = ok and np.array_equal(H.allgather_vector_MPI(minep, du, dd, DimPh=nph), fullp)
        gp = H.gather_vector_MPI(minep, du, dd, root=0, DimPh=nph)
        ok = ok and ((gp is None) if rank != 0 else np.array_equal(gp, fullp))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ns,nup,ndw", [(6, 3, 3), (7, 3, 2), (5, 2, 4)])
def test_scatter_gather_world2(ns, nup, ndw):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + ns * 7 + nup) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ns, nup, ndw, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_split_matches_reference_formula():
    import edipack_b200.host as H

    for n in (1, 7, 70, 12870):
        for P in (1, 2, 3, 8):
            tot, prev_end = 0, 0
            for r in range(P):
                q, s0 = H.mpi_split(n, P, r)
                assert s0 == prev_end
                prev_end = s0 + q
                tot += q
                assert q == n // P + (1 if r < n % P else 0)  # ED_HAMILTONIAN_NORMAL.f90:128-131
            assert tot == n


# ------------------------------------------------------------------------------------------------
# halo exchange plan of the sharded Hdw term (edgpu_halo_plan = the host logic of
# edgpu_sector_open_normal with nranks > 1, comm.cu): checked on CPU against the oracle's hop table
# ------------------------------------------------------------------------------------------------
def _halo_inputs(ns, ndw, world):
    """need maps of every rank from the oracle's H_dw rows: need[r][d] = 1 when a column of rank r's
    chunk has a hop to column d outside the chunk."""
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import edipack_b200.host as H
    import edipack_oracle as O
    from models import star_kwargs

    mo = O.Model(**star_kwargs(ns - 1))
    rp, cols, vals = O.hop_csr(mo, 1, ndw)
    dd = len(rp) - 1
    need = np.zeros((world, dd), np.uint8)
    for r in range(world):
        q, d0 = H.mpi_split(dd, world, r)
        for d in range(d0, d0 + q):
            for k in range(rp[d], rp[d + 1]):
                t = int(cols[k]) - 1
                if t < d0 or t >= d0 + q:
                    need[r, t] = 1
    return H, O, mo, rp, cols, vals, dd, need


def _plan(H, dd, world, rank, need):
    import ctypes as C

    from edipack_b200 import _abi

    L = _abi.load()
    halo = np.zeros(world, np.int64)
    cap = int(need.sum()) + 1
    trip = np.zeros(3 * cap, np.int32)
    ns_ = C.c_int64()
    _abi.check(L.edgpu_halo_plan(dd, world, rank, need.ctypes.data, halo.ctypes.data, trip.ctypes.data, cap,
                                 C.byref(ns_)))
    return halo, trip[:3 * ns_.value].reshape(-1, 3)


@pytest.mark.parametrize("ns,ndw,world", [(8, 4, 2), (8, 4, 3), (10, 5, 4), (10, 4, 8), (6, 1, 8), (12, 6, 8)])
def test_halo_plan_covers_every_remote_hop(ns, ndw, world):
    """Every rank's halo is exactly the set of remote columns its hops read (ascending slots), and
    the send lists of all owners fill every halo slot exactly once with the right column."""
    H, O, mo, rp, cols, vals, dd, need = _halo_inputs(ns, ndw, world)
    filled = [dict() for _ in range(world)]
    halos = None
    for me in range(world):
        halo, trip = _plan(H, dd, world, me, need)
        halos = halo if halos is None else halos
        assert np.array_equal(halo, halos)  # every rank derives the same halo sizes
        q, d0 = H.mpi_split(dd, world, me)
        for loc, dst, slot in trip:
            assert 0 <= loc < q and dst != me
            assert slot not in filled[dst]
            filled[dst][int(slot)] = d0 + int(loc)
    for r in range(world):
        want = [d for d in range(dd) if need[r, d]]
        assert halos[r] == len(want)
        assert [filled[r][k] for k in range(len(want))] == want
    # ranks beyond DimDw own nothing and need nothing (the reference's shrunk communicator)
    for r in range(world):
        q, _ = H.mpi_split(dd, world, r)
        if q == 0:
            assert halos[r] == 0


def _halo_worker(rank, world, port, ns, nup, ndw, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, O, mo, rp, cols, vals, dd, need = _halo_inputs(ns, ndw, world)
        du = O.sector_dims(ns, nup, ndw)[0]
        halo, trip = _plan(H, dd, world, rank, need)
        full = (O.start_vector(du * dd, 3) - 0.5).reshape(dd, du)  # [column][up row]
        qd, d0 = H.mpi_split(dd, world, rank)
        mine = full[d0:d0 + qd].copy()
        # "push": every owner sends its listed columns to the readers' halo slots
        outs = []
        for dst in range(world):
            t = torch.zeros((int(halo[dst]), du), dtype=torch.float64)
            for loc, d, slot in trip:
                if d == dst:
                    t[slot] = torch.from_numpy(mine[loc])
            outs.append(t)
        mybox = torch.zeros((int(halo[rank]), du), dtype=torch.float64)
        for dst in range(world):  # one reduce per reader: the owners' slots are disjoint
            t = outs[dst].clone()
            dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
            if dst == rank:
                mybox = t
        # pass A on the chunk: local columns or halo slots (colmap of the engine)
        slot_of = {d: k for k, d in enumerate([d for d in range(dd) if need[rank, d]])}
        hv = np.zeros_like(mine)
        for j in range(qd):
            d = d0 + j
            for k in range(rp[d], rp[d + 1]):
                t = int(cols[k]) - 1
                src = mine[t - d0] if d0 <= t < d0 + qd else mybox[slot_of[t]].numpy()
                hv[j] += vals[k] * src
        # oracle: H_dw applied to the full vector
        ref = np.zeros_like(full)
        for d in range(dd):
            for k in range(rp[d], rp[d + 1]):
                ref[d] += vals[k] * full[int(cols[k]) - 1]
        ok = np.abs(hv - ref[d0:d0 + qd]).max() < 1e-13 if qd else True
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ns,nup,ndw", [(8, 4, 4), (7, 3, 1)])
def test_halo_exchange_world2_gloo(ns, nup, ndw):
    """The sharded Hdw term driven by edgpu_halo_plan on 2 gloo ranks (host arrays standing in for
    device memory, reduce standing in for the NVLink stores): equals the oracle's H_dw v."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + ns * 11 + ndw) % 2000
    procs = [ctx.Process(target=_halo_worker, args=(r, 2, port, ns, nup, ndw, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
