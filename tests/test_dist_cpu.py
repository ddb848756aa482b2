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
