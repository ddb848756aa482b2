"""ed_total_ud = F: orbital-resolved sectors of the NORMAL mode.  The oracle
restates direct/Orbs/HxV_local.f90, HxV_up.f90, HxV_dw.f90 literally; its independent check is the
block structure: without inter-orbital one-body terms and with Jx = Jp = 0 the orbital-resolved
sector is an invariant block of the ed_total_ud = T sector (sum Nups, sum Ndws), up to the constant
Hartree shift the two fragments disagree on (0.25 vs 0.5 per pair, Orbs/HxV_local.f90:66-67)."""
import numpy as np
import pytest

from models import star_kwargs, two_orb_kwargs


def orbs_kwargs(nbath=2, norb=2, hf=True):
    rng = np.random.default_rng(19)
    if norb == 1:
        kw = star_kwargs(nbath)
    else:
        kw = two_orb_kwargs(nbath, with_nd=False)
    kw.update(Nspin=2, hfmode=hf, xmu=0.13, bath_e=rng.standard_normal((2, norb, nbath)),
              bath_v=0.3 + rng.random((2, norb, nbath)), spin_field_z=tuple(0.1 * rng.standard_normal(norb)))
    return kw


def reorder_sign(oracle, m, nups, ndws):
    """Jordan-Wigner gauge between the two bases: c / cdg count the occupied levels below the
    operated one inside ONE spin integer -- the Ns-bit one (global site order) for ed_total_ud=T,
    the (1+Nbath)-bit one of the orbital for ed_total_ud=F.  The two conventions differ by the sign
    of the permutation that sorts the occupied sites of a spin from orbital-block order into global
    site order: S = prod_spin (-1)^(# inversions)."""
    No, Nb = m.Norb, m.Nbath
    nso = Nb + 1
    dups, ddws = oracle.orbs_dims(m, nups, ndws)
    dims = dups + ddws
    maps = [oracle.build_map(nso, n) for n in list(nups) + list(ndws)]
    out = np.ones(int(np.prod(dims)))
    for i in range(out.size):
        c, pats = i, []
        for f, d in enumerate(dims):
            pats.append(int(maps[f][c % d]))
            c //= d
        inv = 0
        for spin in range(2):
            sites = []          # global site of every occupied level, listed in orbital-block order
            for a in range(No):
                pm = pats[a + spin * No]
                if pm & 1:
                    sites.append(a)
                for k in range(Nb):
                    if (pm >> (1 + k)) & 1:
                        sites.append(m.bath_stride(a, k) - 1)
            inv += sum(1 for x in range(len(sites)) for y in range(x + 1, len(sites)) if sites[x] > sites[y])
        out[i] = -1.0 if inv & 1 else 1.0
    return out


def hartree_shift(m):
    """total_ud=T minus total_ud=F constant: 0.25 * sum_{a<b} (2 Ust - Jh) in hfmode."""
    if not m.hfmode or m.Norb == 1:
        return 0.0
    um = m.umatrix()
    s = 0.0
    for a in range(m.Norb):
        for b in range(a + 1, m.Norb):
            s += 0.25 * um["Ust"][a, b] + 0.25 * (um["Ust"][a, b] - um["Jh"][a, b])
    return s


@pytest.mark.parametrize("hf", [True, False])
@pytest.mark.parametrize("sec", [((1, 2), (2, 1)), ((0, 1), (3, 2)), ((2, 2), (1, 1)), ((3, 0), (0, 3))])
def test_orbs_sector_is_a_block_of_the_total_sector(oracle, sec, hf):
    nups, ndws = sec
    m = oracle.Model(**orbs_kwargs(2, 2, hf))
    emb = oracle.orbs_embedding(m, nups, ndws)
    assert len(set(emb.tolist())) == len(emb)
    Ht = oracle.dense_H(m, sum(nups), sum(ndws))
    n = len(emb)
    Ho = np.column_stack([oracle.orbs_direct_hxv(m, nups, ndws, e) for e in np.eye(n)])
    assert np.abs(Ho - Ho.T).max() < 1e-14
    S = reorder_sign(oracle, m, nups, ndws)
    block = Ht[np.ix_(emb, emb)] * np.outer(S, S)
    assert np.abs(Ho + hartree_shift(m) * np.eye(n) - block).max() < 1e-13
    # invariant block: no matrix element leaves it
    rest = np.setdiff1d(np.arange(Ht.shape[0]), emb)
    assert np.abs(Ht[np.ix_(rest, emb)]).max() == 0.0


def test_orbs_single_orbital_equals_total(oracle):
    """Norb = 1: the two quantum-number schemes coincide (same maps, same index)."""
    m = oracle.Model(**orbs_kwargs(3, 1))
    du, dd = oracle.sector_dims(m.Ns, 2, 3)
    v = np.random.default_rng(0).standard_normal(du * dd)
    assert np.abs(oracle.orbs_direct_hxv(m, (2,), (3,), v) - oracle.direct_hxv(m, 2, 3, v)).max() < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(2, 2, ((1, 2), (2, 1))), (2, 2, ((0, 3), (3, 0))), (2, 3, ((2, 1), (2, 3))),
                                  (3, 1, ((1, 1, 0), (0, 2, 1))), (1, 4, ((2,), (3,))), (2, 2, ((0, 0), (0, 0)))])
def test_orbs_device_built_sector(engine, oracle, case):
    """edgpu_sector_open_normal_orbs: H x v 1e-12 against the oracle, stored elements against the
    dense oracle matrix, ground state 1e-10."""
    E = engine
    norb, nbath, (nups, ndws) = case
    kw = orbs_kwargs(nbath, norb) if norb <= 2 else None
    if kw is None:
        rng = np.random.default_rng(23)
        kw = dict(Norb=3, Nbath=nbath, Nspin=2, Uloc=(2.0, 1.5, 1.0), Ust=0.8, Jh=0.2, hfmode=True, xmu=0.1,
                  hloc=np.array([np.diag(rng.standard_normal(3)) for _ in range(2)]),
                  bath_e=rng.standard_normal((2, 3, nbath)), bath_v=0.3 + rng.random((2, 3, nbath)))
    m, mo = E.EDModel(**kw), oracle.Model(**kw)
    dups, ddws = oracle.orbs_dims(mo, nups, ndws)
    n = int(np.prod(dups + ddws))
    v = np.random.default_rng(2).standard_normal(n)
    ref = oracle.orbs_direct_hxv(mo, nups, ndws, v)
    E.build_Hv_sector_normal_orbs(m, nups, ndws)
    try:
        assert E.vecDim_Hv_sector_normal() == n
        got = E.spHtimesV_p(v)
        assert np.abs(got - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1.0)
        if n <= 600:
            H = np.column_stack([oracle.orbs_direct_hxv(mo, nups, ndws, e) for e in np.eye(n)])
            rp, cj, va = E.stored_csr()
            Hd = np.zeros((n, n))
            for i in range(n):
                for k in range(rp[i], rp[i + 1]):
                    Hd[i, cj[k] - 1] += va[k]
            assert np.abs(Hd - H).max() < 1e-14
            if n > 1:
                ev, vecs, nconv, _ = E.sp_eigh(1, min(n, 20), 300, 0.0)
                assert abs(ev[0] - np.linalg.eigvalsh(H)[0]) < 1e-10
    finally:
        E.delete_Hv_sector_csr()
    # ED_SPARSE_H = F: directMatVec_normal_orbs, nothing stored -- the same product
    E.set_sparse_H(False)
    try:
        E.build_Hv_sector_normal_orbs(m, nups, ndws)
        try:
            direct = E.spHtimesV_p(v)
            with pytest.raises(E.EdgpuError, match="direct"):
                E.stored_csr()
        finally:
            E.delete_Hv_sector_csr()
    finally:
        E.set_sparse_H(True)
    assert np.abs(direct - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1.0)
    assert np.abs(direct - got).max() <= 1e-12 * max(np.abs(ref).max(), 1.0)


@pytest.mark.gpu
def test_orbs_refuses_inter_orbital_terms(engine):
    E = engine
    kw = two_orb_kwargs(2)          # Jx, Jp != 0
    with pytest.raises(E.EdgpuError, match="ed_total_ud"):
        E.build_Hv_sector_normal_orbs(E.EDModel(**kw), (1, 1), (1, 1))
    kw = two_orb_kwargs(2, with_nd=False)
    kw["bath_type"] = "hybrid"
    with pytest.raises(E.EdgpuError):
        E.build_Hv_sector_normal_orbs(E.EDModel(**kw), (1, 1), (1, 1))
