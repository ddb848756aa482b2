"""Pins the oracle to test/src/INEQ_NORMAL_NORMAL/{dens,docc,Sigma_momenta}.check (two
inequivalent impurities, Norb=1 Nbath=7 -> Ns=8, Nspin=2, spin-split baths; the geometry of
BASELINE config 1).  The ground-state sector (4,4) has 4900 states > LANC_DIM_THRESHOLD, i.e.
the reference goes through ARPACK + H x v here."""
import numpy as np
import pytest

from models import golden, ineq_site_kwargs


@pytest.mark.parametrize("ilat", [0, 1])
def test_ineq_site(oracle, ilat):
    g = golden("ineq_normal_normal")
    m = oracle.Model(**ineq_site_kwargs(ilat))
    # dense LAPACK in the GS sector (the Lanczos vector at 1e-12 only gives ~1e-8 observables)
    ev, U = np.linalg.eigh(oracle.dense_H(m, 4, 4))
    # ... and it is the global ground state: the neighbouring sectors lie higher
    for sec in ((3, 4), (4, 3), (5, 4), (4, 5), (3, 3), (5, 5), (3, 5), (5, 3)):
        e, _, _ = oracle.lanc_eigh(lambda x: oracle.direct_hxv(m, sec[0], sec[1], x),
                                   int(np.prod(oracle.sector_dims(m.Ns, *sec))), 300)
        assert e > ev[0] + 1e-6
    st = [oracle.GState(float(ev[0]), 4, 4, U[:, 0].copy())]
    dens, docc = oracle.observables(m, st)
    assert abs(dens[0] - g["dens"][ilat]) < 1e-9
    assert abs(docc[0] - g["docc"][ilat]) < 1e-9
    lmats = int(g["inputs"]["LMATS"])
    pw = oracle.gf_poles_weights(m, st, 0, 0)
    wm, sig = oracle.sigma_matsubara(m, pw, 0, 0, lmats)
    gold = np.array(g["Sigma_momenta"]).reshape(2, 4)[ilat]
    assert np.abs(oracle.momenta(wm, sig) / gold - 1.0).max() < 1e-8
