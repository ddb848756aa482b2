"""Pins the CPU oracle's hybrid / replica / general bath terms (hop tables of 1 (x) Hup and
Hdw (x) 1 with inter-orbital bath hops, ED_NORMAL/direct/HxV_up.f90:33-59) to the reference's
golden vectors test/src/{HYBRID,REPLICA,GENERAL}_NORMAL/{evals,dens,docc}.check at the reference's
own tolerance (1e-9, test/src/ASSERTING.f90:78)."""
import numpy as np
import pytest

from models import golden, hybrid_normal_kwargs, replica_normal_kwargs

FIXTURES = {
    "hybrid_normal": hybrid_normal_kwargs,
    "replica_normal": lambda: replica_normal_kwargs("replica"),
    "general_normal": lambda: replica_normal_kwargs("general"),
}


@pytest.mark.parametrize("name", list(FIXTURES))
def test_bath_fixture(oracle, name):
    g = golden(name)
    m = oracle.Model(**FIXTURES[name]())
    states = oracle.diagonalize(m)
    assert len(states) == 1
    assert abs(states[0].e - g["evals"][0]) < 1e-9
    dens, docc = oracle.observables(m, states)
    assert np.abs(dens - np.array(g["dens"])).max() < 1e-9
    assert np.abs(docc - np.array(g["docc"])).max() < 1e-9
    # direct (on-the-fly) and stored element generators agree on this bath type
    v = oracle.start_vector(400, 3) - 0.5
    a = oracle.direct_hxv(m, 3, 3, v)
    b = oracle.stored_hxv(m, 3, 3, v)
    assert np.abs(a - b).max() < 1e-13 * np.abs(a).max()


@pytest.mark.parametrize("name", list(FIXTURES))
def test_bath_fixture_sigma_momenta(oracle, name):
    """Sigma_momenta.check (1e-8 relative, ed_hybrid_normal.f90:100,153): needs the off-diagonal
    impurity Green's function (lanc_build_gf_normal_mix, seeds (c_a + c_b)|gs>), the orbital-matrix
    inversion of G and the matrix hybridisation function of the shared / replica baths."""
    g = golden(name)
    m = oracle.Model(**FIXTURES[name]())
    states = oracle.diagonalize(m)
    lmats = int(g["inputs"]["LMATS"])
    wm, S = oracle.sigma_matrix_matsubara(m, states, 0, lmats)
    gold = np.array(g["Sigma_momenta"]).reshape(m.Norb, 4)
    for a in range(m.Norb):
        mom = oracle.momenta(wm, S[a, a])
        assert np.abs(mom / gold[a] - 1.0).max() < 1e-8, (name, a, mom, gold[a])


@pytest.mark.parametrize("name", list(FIXTURES))
def test_host_mirror_sigma_assembly(oracle, monkeypatch, name):
    """The host mirror's G-matrix assembly, hybridisation matrix and self-energy (get_impG_normal,
    delta_bath_array, get_Sigma_normal) on CPU: the two device-backed pole/weight builders are
    replaced by the oracle's, everything else is the product's host code."""
    import edipack_b200.host as H

    g = golden(name)
    kw = FIXTURES[name]()
    mo = oracle.Model(**kw)
    states = oracle.diagonalize(mo)
    monkeypatch.setattr(H, "lanc_build_gf_normal_diag",
                        lambda model, sts, iorb, ispin=0: oracle.gf_poles_weights(mo, states, iorb, ispin))
    monkeypatch.setattr(H, "lanc_build_gf_normal_mix",
                        lambda model, sts, a, b, ispin=0: oracle.gf_poles_weights_mix(mo, states, a, b, ispin))
    m = H.EDModel(**kw)
    wm, S = H.get_Sigma_normal(m, [H.EState(states[0].e, states[0].nup, states[0].ndw, 0)],
                               int(g["inputs"]["LMATS"]), 0)
    gold = np.array(g["Sigma_momenta"]).reshape(m.Norb, 4)
    for a in range(m.Norb):
        assert np.abs(oracle.momenta(wm, S[a, a]) / gold[a] - 1.0).max() < 1e-8
