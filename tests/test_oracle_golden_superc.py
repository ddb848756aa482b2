"""Pins the SUPERC-mode oracle (oracle/edipack_oracle_superc.py) to the reference's golden vectors
test/src/NORMAL_SUPERC and HYBRID_SUPERC {evals,dens,docc,phisc}.check (reference tolerance 1e-9
on evals; its dens/docc/phisc were produced from an ARPACK vector and are themselves consistent
only to ~2e-8 -- e.g. HYBRID_SUPERC dens sums to 1.99999998693 -- hence 5e-8 there)."""
import numpy as np
import pytest

from models import golden, replica_superc_model, superc_model


@pytest.mark.parametrize("name", ["normal_superc", "hybrid_superc", "replica_superc", "general_superc"])
def test_superc_fixture(name):
    import edipack_oracle_nonsu2 as N
    import edipack_oracle_superc as S

    g = golden(name)
    m = (replica_superc_model(S, name.split("_")[0]) if name.startswith(("replica", "general"))
         else superc_model(S, name))
    best = None
    for sz in range(-2, 3):
        smap, rp, cj, va = S.stored_H(m, sz)
        H = N.to_dense(rp, cj, va)
        assert np.abs(H - H.conj().T).max() < 1e-13
        ev, U = np.linalg.eigh(H)
        if best is None or ev[0] < best[0]:
            best = (ev[0], sz, smap, U[:, 0])
    e, sz, smap, v = best
    assert sz == 0
    assert abs(e - g["evals"][0]) < 1e-9
    dens, docc, phi = S.observables(m, sz, smap, v)
    assert np.abs(dens - np.array(g["dens"])).max() < 5e-8
    assert np.abs(docc - np.array(g["docc"])).max() < 5e-8
    assert np.abs(phi.ravel() - np.array(g["phisc"])).max() < 5e-8


def test_superc_sector_map_order():
    import edipack_oracle_superc as S

    for ns, sz in ((3, 0), (4, 1), (4, -2), (5, 5), (3, -3)):
        smap = S.build_sector(ns, sz)
        assert np.all(np.diff(smap) > 0)
        iup, idw = smap & ((1 << ns) - 1), smap >> ns
        pc = lambda x: np.array([bin(int(t)).count("1") for t in x])
        assert np.all(pc(iup) - pc(idw) == sz)
        from math import comb
        assert len(smap) == sum(comb(ns, k) * comb(ns, k - sz) for k in range(ns + 1) if 0 <= k - sz <= ns)


@pytest.mark.parametrize("name", ["normal_superc", "hybrid_superc", "replica_superc", "general_superc"])
def test_superc_sigma_and_self_momenta(name):
    """Sigma_momenta.check (normal self-energy) and Self_momenta.check (anomalous one) at the
    reference's 1e-8 relative (2e-8 for the two fixtures produced from an ARPACK vector): exact
    Nambu Green's function from the Lehmann representation (dense Sz+-1 sectors), Delta / Fdelta of
    every bath type, Sigma = G0^-1 - [M^-1]_11, Self = F0^-1 - [M^-1]_12.
    HYBRID_SUPERC Self_momenta: with the formulas of the sources (F0^-1 = -impHloc_anomalous - Fdelta,
    fdelta_hybrid.f90) the exact result differs from the fixture by 1.3 % to 6.4 % (growing with the
    moment order) while Sigma_momenta of the same run agrees to 1e-8; the fixture is reproduced to
    3.5e-8 by |+Fdelta - [M^-1]_12|, i.e. with the opposite relative sign between Fdelta and the
    anomalous function than in the normal-bath fixture (which the same code reproduces to 4e-9 with
    the sources' sign).  compute_momentum only sees |Self|, so the sign itself cannot be read off the
    fixture; asserted in that form and flagged here as an open inconsistency of the reference's
    shared-bath anomalous channel, not as understood behaviour."""
    import edipack_oracle as O
    import edipack_oracle_nonsu2 as N
    import edipack_oracle_superc as S
    from models import _f

    kind = name.split("_")[0]
    g = golden(name)
    m = replica_superc_model(S, kind) if kind in ("replica", "general") else superc_model(S, name)
    best = None
    for sz in (-1, 0, 1):
        smap, rp, cj, va = S.stored_H(m, sz)
        ev, U = np.linalg.eigh(N.to_dense(rp, cj, va))
        if best is None or ev[0] < best[0]:
            best = (ev[0], sz, smap, U[:, 0])
    e0, sz, smap, v = best
    No = m.Norb
    wm, Sig, Slf = S.sigma_self_matsubara(m, sz, smap, v, e0, _f(g["inputs"]["BETA"]), int(g["inputs"]["LMATS"]))
    tol = 1e-8 if kind in ("replica", "general") else 2e-8
    gs = np.array(g["Sigma_momenta"]).reshape(No, 4)
    for a in range(No):
        assert np.abs(O.momenta(wm, Sig[a, a]) / gs[a] - 1.0).max() < tol
    ga = np.array(g["Self_momenta"])
    if kind == "hybrid":
        _, Fd = S.bath_nambu_functions(m, 1j * wm)
        ga = ga.reshape(No, 4)
        for a in range(No):
            alt = Slf[a, a] + 2.0 * Fd[a, a]          # = +Fdelta - [M^-1]_12 (hloc_anomalous = 0 here)
            assert np.abs(O.momenta(wm, alt) / ga[a] - 1.0).max() < 5e-8
            assert np.abs(O.momenta(wm, Slf[a, a]) / ga[a] - 1.0).max() > 1e-2   # the sources' sign does not
        return
    if ga.size == 4 * No:
        ga = ga.reshape(No, 4)
        for a in range(No):
            assert np.abs(O.momenta(wm, Slf[a, a]) / ga[a] - 1.0).max() < tol
    else:
        ga = ga.reshape(No, No, 4)      # ASmomAB(Norb,Norb,Nmomenta), last index fastest in the file
        for a in range(No):
            for b in range(No):
                assert np.abs(O.momenta(wm, Slf[a, b]) / ga[a, b] - 1.0).max() < tol
