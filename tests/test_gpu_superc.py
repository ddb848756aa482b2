"""ed_mode=superc built on the device (edgpu_sector_open_superc) against the numpy oracle's
restatement of build_sector / ed_buildH_superc_main (oracle/edipack_oracle_superc.py, pinned to the
NORMAL_SUPERC / HYBRID_SUPERC goldens).  Bars: sector map bit-exact, matrix elements 1e-14, H x v
1e-12, E_gs 1e-9 against the reference's evals.check, dens / docc / phisc 5e-8 (see
tests/test_oracle_golden_superc.py for why not 1e-8)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def to_engine_model(E, mo):
    return E.EDModelSuperc(**vars(mo))


def messy_superc(S):
    """Spin-dependent bath, complex anomalous Hloc, pair field, inter-orbital Hloc, hybrid bath."""
    rng = np.random.default_rng(13)
    norb, nbath = 2, 3
    hloc = np.zeros((2, norb, norb), complex)
    for s in range(2):
        a = rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb))
        hloc[s] = 0.3 * (a + a.conj().T)
    an = 0.1 * (rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb)))
    an = an + an.T  # spin-singlet pairing: symmetric in the orbitals
    m = S.ModelSuperc(Norb=norb, Nbath=nbath, bath_type="hybrid", Uloc=(-1.3, -2.1), Ust=-0.9, Jh=0.2,
                      Jx=0.15, Jp=0.1, xmu=0.23, hfmode=False, hloc=hloc, hloc_anomalous=an,
                      pair_field=(0.05, -0.02))
    m.bath_e = rng.standard_normal((2, 1, nbath))
    m.bath_d = 0.1 * rng.standard_normal((1, nbath))
    m.bath_v = 0.4 + rng.random((2, norb, nbath))
    return m


def messy_replica_superc(S):
    """Replica bath with random complex hermitian Nambu blocks."""
    from models import replica_superc_model
    rng = np.random.default_rng(17)
    m = replica_superc_model(S, "replica")
    no, nb = m.Norb, m.Nbath
    hb = np.zeros((2, 2, no, no, nb), complex)
    for k in range(nb):
        a = rng.standard_normal((2 * no, 2 * no)) + 1j * rng.standard_normal((2 * no, 2 * no))
        a = 0.3 * (a + a.conj().T)
        hb[..., k] = a.reshape(2, no, 2, no).transpose(0, 2, 1, 3)
    m.hbath = hb
    m.bath_v = 0.3 + rng.random((2, no, nb))
    m.hfmode = False
    return m


@pytest.mark.parametrize("name", ["normal_superc", "hybrid_superc", "messy", "replica_superc",
                                  "general_superc", "messy_replica"])
def test_device_built_superc_matches_oracle(engine, name):
    import edipack_oracle_nonsu2 as N
    import edipack_oracle_superc as S
    from models import replica_superc_model, superc_model

    E = engine
    if name == "messy":
        mo = messy_superc(S)
    elif name == "messy_replica":
        mo = messy_replica_superc(S)
    elif name.startswith(("replica", "general")):
        mo = replica_superc_model(S, name.split("_")[0])
    else:
        mo = superc_model(S, name)
    m = to_engine_model(E, mo)
    rng = np.random.default_rng(4)
    for sz in (0, 1, -2, mo.Ns, -mo.Ns):
        smap, rp, cj, va = S.stored_H(mo, sz)
        Href = N.to_dense(rp, cj, va)
        E.build_Hv_sector_superc(m, sz)
        try:
            assert np.array_equal(E.sector_map_nonsu2(), smap.astype(np.int32)), (name, sz)
            drp, dcj, dva = E.stored_csr()
            Hdev = N.to_dense(drp, dcj, dva)
            assert np.abs(Hdev - Href).max() < 1e-14, (name, sz)
            v = rng.standard_normal(len(smap)) + 1j * rng.standard_normal(len(smap))
            hv = E.spHtimesV_cc(v)
            ref = N.csr_matvec(rp, cj, va, v)
            assert np.abs(hv - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1e-300), (name, sz)
        finally:
            E.delete_Hv_sector_superc()


@pytest.mark.parametrize("name", ["normal_superc", "hybrid_superc", "replica_superc", "general_superc"])
def test_golden_superc_device_built(engine, name):
    """test/src/{NORMAL,HYBRID,REPLICA,GENERAL}_SUPERC/{evals,dens,docc,phisc}.check with the sector
    map, the stored H and the eigen-solver all on the device."""
    import edipack_oracle_superc as S
    from models import golden, replica_superc_model, superc_model

    E = engine
    g = golden(name)
    mo = (replica_superc_model(S, name.split("_")[0]) if name.startswith(("replica", "general"))
          else superc_model(S, name))
    m = to_engine_model(E, mo)
    best = None
    for sz in (-1, 0, 1):
        E.build_Hv_sector_superc(m, sz)
        try:
            ev, vec, nconv, _ = E.sp_eigh(1, 20, 512, 1e-16)
            smap = E.sector_map_nonsu2()
        finally:
            E.delete_Hv_sector_superc()
        if best is None or ev[0] < best[0]:
            best = (ev[0], vec[:, 0], smap, sz)
    e, vec, smap, sz = best
    assert sz == 0
    assert abs(e - g["evals"][0]) < 1e-9
    dens, docc, phi = S.observables(mo, sz, smap.astype(np.int64), vec)
    assert np.abs(dens - np.array(g["dens"])).max() < 5e-8
    assert np.abs(docc - np.array(g["docc"])).max() < 5e-8
    assert np.abs(phi.ravel() - np.array(g["phisc"])).max() < 5e-8
