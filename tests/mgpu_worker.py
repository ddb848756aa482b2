"""Multi-GPU parity worker (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_worker.py

Each rank owns the reference's dw-column chunk of the sector vector
(ED_HAMILTONIAN_NORMAL.f90:128-142).  Checks, against the CPU oracle on the full vector:
  * H x v of the sharded engine (NCCL tile transposes for the Hdw term)       1e-12 relative
  * sp_lanc_tridiag alpha/beta and sp_lanc_eigh energy (NCCL all-reduced dots) 1e-9 / 1e-10
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

import edipack_b200 as E
import edipack_oracle as O
from models import normal_normal_kwargs, star_kwargs, two_orb_kwargs


def amax(x):
    """max |x| of a possibly empty chunk (ranks beyond DimDw own no column)."""
    return float(np.abs(x).max()) if x.size else 0.0


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E.ed_init(local)
    uid = [E.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    E.ed_set_comm(rank, world, uid[0])
    failures = []
    cases = [("star7", star_kwargs(7), (4, 4)), ("star9", star_kwargs(9), (5, 5)),
             ("star9_54", star_kwargs(9), (5, 4)), ("star11", star_kwargs(11), (6, 6)),
             ("two_orb_nb3_nojx", two_orb_kwargs(3, with_nd=False), (4, 4)),
             ("two_orb_nb3_jxjp", two_orb_kwargs(3), (4, 4)), ("two_orb_nb2_jxjp", two_orb_kwargs(2), (3, 2)),
             ("star13", star_kwargs(13), (7, 7)),
             # sectors smaller than the communicator (the reference shrinks it,
             # ED_HAMILTONIAN_NORMAL.f90:98-126; here the ranks beyond DimDw own no column)
             ("star7_dw1", star_kwargs(7), (4, 0)), ("star7_up1", star_kwargs(7), (8, 3)),
             ("star7_1x1", star_kwargs(7), (8, 0)), ("two_orb_dw1", two_orb_kwargs(2), (2, 6)),
             ("star7_dw8", star_kwargs(7), (3, 1))]
    for name, kw, (nup, ndw) in cases:
        m, mo = E.EDModel(**kw), O.Model(**kw)
        du, dd = O.sector_dims(m.Ns, nup, ndw)
        full = O.start_vector(du * dd, 17) - 0.5
        lo, hi = E.chunk_bounds(du, dd, world, rank)
        E.build_Hv_sector_normal(m, nup, ndw)
        try:
            assert E.vecDim_Hv_sector_normal() == hi - lo
            hv = E.spHtimesV_p(full[lo:hi].copy())
            if du * dd <= 400000:
                ref = O.direct_hxv(mo, nup, ndw, full)
            else:
                ref = O.stored_hxv_mpi(mo, nup, ndw, full, 8, 8)[0]
            err = amax(hv - ref[lo:hi]) / np.abs(ref).max()
            if not err < 1e-12:
                failures.append(f"{name}: HxV rel err {err:.3e} on rank {rank}")
            if du * dd <= 70000:
                seed = full / np.linalg.norm(full)
                a0, b0, n0 = O.lanc_tridiag(lambda x: O.direct_hxv(mo, nup, ndw, x), seed, 12)
                a1, b1, n1, n2 = E.sp_lanc_tridiag(full[lo:hi].copy(), 12)
                if n0 != n1 or np.abs(a1[:n1] - a0[:n0]).max() > 1e-9 or np.abs(b1[:n1] - b0[:n0]).max() > 1e-9:
                    failures.append(f"{name}: tridiag mismatch on rank {rank}")
                if abs(n2 - full @ full) > 1e-10 * (full @ full):
                    failures.append(f"{name}: norm2 {n2} vs {full @ full}")
                e, vec, nit = E.sp_lanc_eigh(min(du * dd, 300), 1e-14)
                e_ref, _, _ = O.lanc_eigh(lambda x: O.direct_hxv(mo, nup, ndw, x), du * dd, 300, 1e-14)
                if abs(e - e_ref) > 1e-10:
                    failures.append(f"{name}: E_gs {e} vs {e_ref}")
                # the returned chunk is this rank's part of a normalised vector
                n2loc = torch.tensor([float(vec @ vec)], dtype=torch.float64, device="cuda")
                dist.all_reduce(n2loc)
                if abs(float(n2loc.item()) - 1.0) > 1e-10:
                    failures.append(f"{name}: |gs|^2 = {float(n2loc.item())}")
            if 3 <= du * dd <= 5000:
                # sp_eigh (thick-restart Lanczos, all-reduced projection coefficients) on the shards
                ref_ev = np.linalg.eigvalsh(O.dense_H(mo, nup, ndw))[:3]
                ev, vecs, nconv, nmv = E.sp_eigh(3, 24, 300, 1e-14)
                if nconv < 3 or np.abs(ev - ref_ev).max() > 1e-10:
                    failures.append(f"{name}: sp_eigh {ev} vs {ref_ev} (nconv {nconv})")
                g = torch.tensor(vecs.T @ vecs, dtype=torch.float64, device="cuda")
                dist.all_reduce(g)
                if np.abs(g.cpu().numpy() - np.eye(3)).max() > 1e-10:
                    failures.append(f"{name}: sp_eigh vectors not orthonormal across ranks")
        finally:
            E.delete_Hv_sector_normal()
    # The reference's own fixture over ALL sectors on the shards (ed_diag_d visits the 1-state and
    # DimDw < nranks sectors too): test/src/NORMAL_NORMAL/{evals,dens,docc,Sigma_momenta}.check
    from models import golden
    g = golden("normal_normal")
    kw = normal_normal_kwargs()
    m, mo = E.EDModel(**kw), O.Model(**kw)
    m.lanc_tolerance = 1e-18
    try:
        states = E.ed_diag_d(m)
        if not (len(states) == 1 and (states[0].nup, states[0].ndw) == (3, 3)):
            failures.append(f"golden scan: states {[(s.nup, s.ndw, s.e) for s in states]}")
        elif abs(states[0].e - g["evals"][0]) > 1e-9:
            failures.append(f"golden scan: E_gs {states[0].e} vs {g['evals'][0]}")
        else:
            dens, docc = E.observables_normal(m, states)
            if np.abs(dens - np.array(g["dens"])).max() > 1e-8 or np.abs(docc - np.array(g["docc"])).max() > 1e-8:
                failures.append(f"golden scan: dens {dens} docc {docc}")
            gold = np.array(g["Sigma_momenta"]).reshape(m.Norb, 4)
            for iorb in range(m.Norb):
                pw = E.lanc_build_gf_normal_diag(m, states, iorb, 0)
                wm, sig = O.sigma_matsubara(mo, pw, iorb, 0, int(g["inputs"]["LMATS"]))
                if np.abs(O.momenta(wm, sig) / gold[iorb] - 1.0).max() > 1e-8:
                    failures.append(f"golden scan: Sigma moments of orbital {iorb}")
        for s_ in states:
            E.state_free(s_.slot)
    except Exception as ex:  # keep the ranks' collectives matched: report, do not raise
        failures.append(f"golden scan raised {type(ex).__name__}: {ex}")
    # Green's-function seeds on the sharded state: c / c^+ of both spins applied to the resident
    # ground state, read back through norm2 and alpha_1 of the tridiagonalisation
    # (apply_op_C/CDG, ED_SECTOR.f90:465/654; the dw operators need the gathered state)
    kw = normal_normal_kwargs()
    m, mo = E.EDModel(**kw), O.Model(**kw)
    nup, ndw = 3, 3
    du, dd = O.sector_dims(m.Ns, nup, ndw)
    if True:
        H = O.dense_H(mo, nup, ndw)
        ev, U = np.linalg.eigh(H)
        gs = U[:, 0]
        lo, hi = E.chunk_bounds(du, dd, world, rank)
        E.build_Hv_sector_normal(m, nup, ndw)
        try:
            e, vec, nit = E.sp_lanc_eigh(1, 1e-14, vect=gs[lo:hi].copy())  # current state := gs
            E.state_store(5)
        finally:
            E.delete_Hv_sector_normal()
        for spin in (0, 1):
            for iorb in range(m.Norb):
                for op in (+1, -1):
                    ref, jn = O.apply_op(mo, op, iorb, spin, nup, ndw, gs)
                    E.build_Hv_sector_normal(m, jn[0], jn[1])
                    try:
                        E.apply_op(5, op, iorb, spin)
                        a, b, nused, n2 = E.sp_lanc_tridiag(None, 1)
                    finally:
                        E.delete_Hv_sector_normal()
                    if abs(n2 - ref @ ref) > 1e-12:
                        failures.append(f"apply_op spin={spin} orb={iorb} op={op}: norm2 {n2} vs {ref @ ref}")
                    elif n2 > 0:
                        hv = O.direct_hxv(mo, jn[0], jn[1], ref)
                        if abs(a[0] - (ref @ hv) / n2) > 1e-10:
                            failures.append(f"apply_op spin={spin} orb={iorb} op={op}: alpha1 mismatch")
        E.state_free(5)
    # a10 on the shards: coulomb_sundry (gathers from the all-gathered vector) and phonon slices
    # (every rank holds DimPh consecutive electronic chunks, direct_mpi/HxV_eph.f90:3-4; the
    # off-diagonal g_ph dw hops read other ranks' columns)
    from test_oracle_sundry_phonons import SUNDRY
    ph = dict(Nph=2, w0=0.37, g=[[0.5, 0.2], [0.2, -0.3]])
    for name, kw, (nup, ndw), sundry, phon in [
            ("sundry", two_orb_kwargs(2), (3, 3), SUNDRY, None),
            ("phonons", two_orb_kwargs(2), (3, 2), [], ph),
            ("sundry+phonons", two_orb_kwargs(3, with_nd=True), (4, 4), SUNDRY, ph)]:
        m, mo = E.EDModel(**kw), O.Model(**kw)
        du, dd = O.sector_dims(m.Ns, nup, ndw)
        nph = (phon["Nph"] if phon else 0) + 1
        full = O.start_vector(du * dd * nph, 23) - 0.5
        lo, hi = E.chunk_bounds(du, dd, world, rank)
        pick = np.concatenate([np.arange(k * du * dd + lo, k * du * dd + hi) for k in range(nph)])
        E.set_coulomb_sundry(sundry)
        E.set_phonons(**(phon or dict(Nph=0)))
        E.build_Hv_sector_normal(m, nup, ndw)
        try:
            hv = E.spHtimesV_p(full[pick].copy())
            ref = O.direct_hxv_ext(mo, nup, ndw, full, sundry, phon)
            err = amax(hv - ref[pick]) / np.abs(ref).max()
            if not err < 1e-12:
                failures.append(f"a10 {name}: HxV rel err {err:.3e} on rank {rank}")
            if du * dd * nph <= 6000:
                n = du * dd * nph
                H = np.column_stack([O.direct_hxv_ext(mo, nup, ndw, e_, sundry, phon) for e_ in np.eye(n)])
                # (the SUNDRY test list is not hermitian: ground-state check only for symmetric H)
                sym = np.abs(H - H.T).max() < 1e-12
                e, vec, nit = E.sp_lanc_eigh(300, 1e-14) if sym else (0.0, None, 0)
                if sym and abs(e - np.linalg.eigvalsh(H)[0]) > 1e-10:
                    failures.append(f"a10 {name}: E_gs {e} vs {np.linalg.eigvalsh(H)[0]}")
        finally:
            E.delete_Hv_sector_normal()
            E.set_coulomb_sundry(())
            E.set_phonons(0)
    # ed_mode=nonsu2 built on the device: every rank generates its rows of the flat row split
    # (ED_HAMILTONIAN_NONSU2.f90:72-79), the product all-gathers the input vector
    import edipack_oracle_nonsu2 as N
    from models import soc_nonsu2_model

    mo2 = soc_nonsu2_model(N, nbath=4)
    m2 = E.EDModelNonsu2(**vars(mo2))
    smap, rp, cj, va = N.stored_H(mo2, 6)
    dim = len(smap)
    q = dim // world
    lo, hi = q * rank, (dim if rank == world - 1 else q * (rank + 1))
    rng = np.random.default_rng(31)
    vfull = rng.standard_normal(dim) + 1j * rng.standard_normal(dim)
    ref = N.csr_matvec(rp, cj, va, vfull)
    E.build_Hv_sector_nonsu2(m2, 6)
    try:
        if E.vecDim_Hv_sector_normal() != hi - lo:
            failures.append(f"nonsu2: vecDim {E.vecDim_Hv_sector_normal()} vs {hi - lo}")
        hv = E.spHtimesV_cc(vfull[lo:hi].copy())
        err = amax(hv - ref[lo:hi]) / np.abs(ref).max()
        if not err < 1e-12:
            failures.append(f"nonsu2: HxV rel err {err:.3e} on rank {rank}")
        ev, _, nconv, _ = E.sp_eigh(2, 24, 300, 1e-14, want_vectors=False)
        ev_ref = np.linalg.eigvalsh(N.to_dense(rp, cj, va))[:2]
        if np.abs(ev - ev_ref).max() > 1e-10:
            failures.append(f"nonsu2: sp_eigh {ev} vs {ev_ref}")
    finally:
        E.delete_Hv_sector_nonsu2()
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    for f in failures:
        print(f"[rank {rank}] FAIL {f}", flush=True)
    if rank == 0:
        print(f"mgpu_worker: world={world} failures={int(flag.item())}", flush=True)
    E.ed_finalize()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
