"""Pins the CPU oracle to the reference's own golden vectors:
test/src/NORMAL_NORMAL/{evals,dens,docc,Sigma_momenta}.check at the reference's tolerances
(1e-9 absolute, Sigma moments 1e-8 relative; test/src/ASSERTING.f90:78,
ed_normal_normal.f90:160-171)."""
import numpy as np
import pytest

from models import golden, normal_normal_kwargs


@pytest.fixture(scope="module")
def solved(oracle):
    m = oracle.Model(**normal_normal_kwargs())
    states = oracle.diagonalize(m)
    return m, states


def test_ground_state_energy(oracle, solved):
    m, states = solved
    g = golden("normal_normal")
    assert len(states) == 1
    assert (states[0].nup, states[0].ndw) == (3, 3)
    assert abs(states[0].e - g["evals"][0]) < 1e-9


def test_dens_docc(oracle, solved):
    m, states = solved
    g = golden("normal_normal")
    dens, docc = oracle.observables(m, states)
    assert np.abs(dens - np.array(g["dens"])).max() < 1e-9
    assert np.abs(docc - np.array(g["docc"])).max() < 1e-9


@pytest.mark.parametrize("hxv_kind", ["direct", "stored"])
def test_sigma_momenta(oracle, solved, hxv_kind):
    """Sigma(iw) through apply_op_C/CDG + sp_lanc_tridiag + add_to_lanczos_gf_normal, both
    ED_SPARSE_H=T and =F paths like the reference test (run_test sparse=T/F)."""
    m, states = solved
    g = golden("normal_normal")
    lmats = int(g["inputs"]["LMATS"])
    gold = np.array(g["Sigma_momenta"]).reshape(m.Norb, 4)
    for iorb in range(m.Norb):
        pw = oracle.gf_poles_weights(m, states, iorb, 0, hxv_kind=hxv_kind)
        wm, sig = oracle.sigma_matsubara(m, pw, iorb, 0, lmats)
        mom = oracle.momenta(wm, sig)
        assert np.abs(mom / gold[iorb] - 1.0).max() < 1e-8


def test_lanczos_ground_state_matches_dense(oracle):
    """The plain-Lanczos driver on the sector (3,3) H x v reproduces the golden energy."""
    m = oracle.Model(**normal_normal_kwargs())
    g = golden("normal_normal")
    dim = 400
    e, v, nit = oracle.lanc_eigh(lambda x: oracle.direct_hxv(m, 3, 3, x), dim, 300)
    assert abs(e - g["evals"][0]) < 1e-10
    hv = oracle.direct_hxv(m, 3, 3, v)
    assert np.abs(hv - e * v).max() < 1e-6
