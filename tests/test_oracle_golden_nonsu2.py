"""Pins the nonsu2 stored-H oracle (oracle/edipack_oracle_nonsu2.py) to the reference's golden
values test/src/HYBRID_NONSU2/{evals,dens,docc,magX}.check at the reference's 1e-9
(test/src/ASSERTING.f90:78)."""
import numpy as np
import pytest

from models import golden, hybrid_nonsu2_model, replica_nonsu2_model, soc_nonsu2_model


@pytest.fixture(scope="module")
def N():
    import edipack_oracle_nonsu2 as N

    return N


@pytest.fixture(scope="module")
def solved(N):
    m = hybrid_nonsu2_model(N)
    best = None
    for nt in range(0, 2 * m.Ns + 1):
        smap, rp, cj, va = N.stored_H(m, nt)
        H = N.to_dense(rp, cj, va)
        assert np.abs(H - H.conj().T).max() < 1e-12  # the stored matrix is Hermitian
        ev, U = np.linalg.eigh(H)
        if best is None or ev[0] < best[0]:
            best = (ev[0], nt, smap, U[:, 0], ev)
    return m, best


def test_ground_state_energy(solved):
    m, (e, nt, smap, vec, ev) = solved
    g = golden("hybrid_nonsu2")
    assert nt == 6 and len(smap) == 924
    assert abs(e - g["evals"][0]) < 1e-9
    assert ev[1] - ev[0] > 1e-3  # unique ground state: zeta_function = 1


def test_dens_docc_magx(N, solved):
    m, (e, nt, smap, vec, ev) = solved
    g = golden("hybrid_nonsu2")
    dens, docc, magx = N.observables(m, smap, vec)
    assert np.abs(dens - np.array(g["dens"])).max() < 1e-9
    assert np.abs(docc - np.array(g["docc"])).max() < 1e-9
    assert np.abs(magx - np.array(g["magX"])).max() < 1e-9


def test_sector_map_and_matvec(N):
    """Map = ascending m = iup + idw*2**Ns with popcount Ntot; CSR product = dense product."""
    m = soc_nonsu2_model(N, nbath=3)
    smap, rp, cj, va = N.stored_H(m, 4)
    assert np.all(np.diff(smap) > 0)
    assert all(bin(int(x)).count("1") == 4 for x in smap)
    H = N.to_dense(rp, cj, va)
    assert np.abs(H - H.conj().T).max() < 1e-12
    rng = np.random.default_rng(0)
    v = rng.standard_normal(len(smap)) + 1j * rng.standard_normal(len(smap))
    assert np.abs(N.csr_matvec(rp, cj, va, v) - H @ v).max() < 1e-12


def test_normal_nonsu2_golden():
    """test/src/NORMAL_NONSU2/{evals,dens,docc,magX}.check: the normal-bath branch of the nonsu2
    generators (one bath chain per orbital, spin-flip hybridisation u = v)."""
    import edipack_oracle_nonsu2 as N

    g = golden("normal_nonsu2")
    m = hybrid_nonsu2_model(N, "normal_nonsu2")
    assert m.bath_type == "normal" and m.Ns == 6
    best = None
    for nt in (5, 6, 7):
        smap, rp, cj, va = N.stored_H(m, nt)
        ev, U = np.linalg.eigh(N.to_dense(rp, cj, va))
        if best is None or ev[0] < best[0]:
            best = (ev[0], nt, smap, U[:, 0])
    e, nt, smap, v = best
    assert nt == 6 and abs(e - g["evals"][0]) < 1e-9
    dens, docc, magx = N.observables(m, smap, v)
    assert np.abs(dens - np.array(g["dens"])).max() < 1e-8
    assert np.abs(docc - np.array(g["docc"])).max() < 1e-8
    assert np.abs(magx - np.array(g["magX"])).max() < 1e-8


@pytest.mark.parametrize("kind", ["replica", "general"])
def test_replica_general_nonsu2_golden(kind):
    """test/src/{REPLICA,GENERAL}_NONSU2/{evals,dens,docc}.check: bath replicas with same-spin and
    spin-flip inter-orbital bath hops (stored/Hbath.f90:49-133)."""
    import edipack_oracle_nonsu2 as N

    g = golden(f"{kind}_nonsu2")
    m = replica_nonsu2_model(N, kind)
    assert m.Ns == 6
    best = None
    for nt in range(3, 10):
        smap, rp, cj, va = N.stored_H(m, nt)
        H = N.to_dense(rp, cj, va)
        assert np.abs(H - H.conj().T).max() < 1e-12
        ev, U = np.linalg.eigh(H)
        if best is None or ev[0] < best[0]:
            best = (ev[0], nt, smap, U[:, 0], ev)
    e, nt, smap, v, ev = best
    assert abs(e - g["evals"][0]) < 1e-9, (e, nt)
    dens, docc, _ = N.observables(m, smap, v)
    assert np.abs(dens - np.array(g["dens"])).max() < 1e-8
    assert np.abs(docc - np.array(g["docc"])).max() < 1e-8
    # exciton.check = [S0, Tx, Ty, Tz](1,2): norms of two-operator seeds (apply_Cops)
    assert np.abs(N.exciton(m, nt, smap, v) - np.array(g["exciton"])).max() < 1e-8


@pytest.mark.parametrize("name", ["hybrid_nonsu2", "normal_nonsu2", "replica_nonsu2", "general_nonsu2"])
def test_nonsu2_sigma_momenta(oracle, name):
    """Sigma11 / Sigma12_momenta.check (hybrid, normal bath) and the full Sigma_momenta.check
    (replica, general: all Nspin x Nspin x Norb x Norb components) at the reference's 1e-8 relative:
    the impurity Green's function matrix from the exact Lehmann representation (dense N+-1 sectors),
    the (spin,orbital) hybridisation matrix of every bath type and Sigma = G0^-1 - G^-1."""
    import edipack_oracle_nonsu2 as N
    from models import _f

    g = golden(name)
    kind = name.split("_")[0]
    m = replica_nonsu2_model(N, kind) if kind in ("replica", "general") else hybrid_nonsu2_model(N, name)
    best = None
    for nt in (5, 6, 7):
        smap, rp, cj, va = N.stored_H(m, nt)
        ev, U = np.linalg.eigh(N.to_dense(rp, cj, va))
        if best is None or ev[0] < best[0]:
            best = (ev[0], nt, smap, U[:, 0])
    e, nt, smap, v = best
    wm, S = N.sigma_matsubara(m, nt, smap, v, e, _f(g["inputs"]["BETA"]), int(g["inputs"]["LMATS"]))
    No = m.Norb
    if "Sigma11_momenta" in g:
        g11 = np.array(g["Sigma11_momenta"]).reshape(No, 4)
        g12 = np.array(g["Sigma12_momenta"]).reshape(No, 4)
        for a in range(No):
            assert np.abs(oracle.momenta(wm, S[a, a]) / g11[a] - 1.0).max() < 1e-8
            assert np.abs(oracle.momenta(wm, S[a, a + No]) / g12[a] - 1.0).max() < 1e-8
    else:
        gold = np.array(g["Sigma_momenta"]).reshape(2, 2, No, No, 4)   # [ispin][jspin][iorb][jorb][moment]
        for s1 in range(2):
            for s2 in range(2):
                for a in range(No):
                    for b in range(No):
                        mom = oracle.momenta(wm, S[a + s1 * No, b + s2 * No])
                        assert np.abs(mom / gold[s1, s2, a, b] - 1.0).max() < 1e-8, (s1, s2, a, b)
