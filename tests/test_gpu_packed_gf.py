"""Green's-function seeds and observables of the packed-state modes on the device: state hand-off
(es_add_state), apply_COps (ED_SECTOR.f90), dens / docc, the superc order parameter through
apply_COps norms (ED_OBSERVABLES_SUPERC.f90:204-248) and a nonsu2 impurity Green's function
(lanc_build_gf_nonsu2_diag, ED_GF_NONSU2.f90:159-205) against the oracle / exact Lehmann sums.
Bars: seeds 1e-12, observables and G(iw) 1e-8."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ground_state(E, build, delete, qns):
    """Lowest state over the given sectors through the device eigen-solver, kept as state 0."""
    best = None
    for q in qns:
        build(q)
        try:
            ev, _, _, _ = E.sp_eigh(1, 20, 512, 1e-16, want_vectors=False)
            if best is None or ev[0] < best[0]:
                best = (ev[0], q)
                E.eigh_state_store(0, 0)
        finally:
            delete()
    return best


def oracle_seed(c_fn, cdg_fn, smap, tmap, vec, terms, Ns):
    """sum_k coef_k O_k |vec> in the target sector, python loops (terms: (coef, op, iorb, spin))."""
    tindex = {int(m): i for i, m in enumerate(tmap)}
    out = np.zeros(len(tmap), complex)
    for i, m_ in enumerate(smap):
        m = int(m_)
        for coef, op, a, sp in terms:
            r = (cdg_fn if op > 0 else c_fn)(a + 1 + sp * Ns, m)
            if r is not None:
                out[tindex[r[0]]] += coef * r[1] * vec[i]
    return out


def test_superc_order_parameter_on_device(engine):
    """phisc of test/src/NORMAL_SUPERC through device seeds: (c_{a,dw} + c^+_{b,up})|gs>."""
    import edipack_oracle_superc as S
    from edipack_oracle_nonsu2 import _c, _cdg
    from models import golden, superc_model

    E = engine
    g = golden("normal_superc")
    mo = superc_model(S, "normal_superc")
    m = E.EDModelSuperc(**vars(mo))
    e, sz = ground_state(E, lambda q: E.build_Hv_sector_superc(m, q), E.delete_Hv_sector_superc, (-1, 0, 1))
    assert sz == 0 and abs(e - g["evals"][0]) < 1e-9
    E.build_Hv_sector_superc(m, 0)
    try:
        dens, docc = E.state_observables(0, m.Norb)
    finally:
        E.delete_Hv_sector_superc()
    assert np.abs(dens - np.array(g["dens"])).max() < 5e-8
    assert np.abs(docc - np.array(g["docc"])).max() < 5e-8
    # <n_dw>, <n_up> per orbital from the oracle's view of the same golden state are not needed:
    # dens_dw(a) = dens_up(a) = dens(a)/2 only for a spin-symmetric state, so take them from seeds
    phi = np.zeros((m.Norb, m.Norb))
    E.build_Hv_sector_superc(m, 1)
    try:
        ndw = np.zeros(m.Norb)
        nup_hole = np.zeros(m.Norb)
        for a in range(m.Norb):
            E.apply_Cops(0, [1.0], [-1], [a], [1])     # c_{a,dw}|gs>   -> <n_{a,dw}>
            ndw[a] = E.seed_norm2()
            E.apply_Cops(0, [1.0], [+1], [a], [0])     # c^+_{a,up}|gs> -> 1 - <n_{a,up}>
            nup_hole[a] = E.seed_norm2()
        for a in range(m.Norb):
            for b in range(m.Norb):
                E.apply_Cops(0, [1.0, 1.0], [-1, +1], [a, b], [1, 0])
                phi[a, b] = 0.5 * (E.seed_norm2() - ndw[a] - nup_hole[b])
    finally:
        E.delete_Hv_sector_superc()
    assert np.abs(phi.ravel() - np.array(g["phisc"])).max() < 5e-8
    E.state_free(0)


def test_nonsu2_seeds_and_gf(engine):
    """Device seeds with complex coefficients against python loops, and G_{a,up;a,up}(iw) of the
    SOC model from device tridiagonalisations against the exact Lehmann sum."""
    import edipack_oracle_nonsu2 as N
    from models import soc_nonsu2_model

    E = engine
    mo = soc_nonsu2_model(N, nbath=4)
    m = E.EDModelNonsu2(**vars(mo))
    Ns = mo.Ns
    nt = 6
    smap, rp, cj, va = N.stored_H(mo, nt)
    Hd = N.to_dense(rp, cj, va)
    ev, U = np.linalg.eigh(Hd)
    e0, gs = ev[0], U[:, 0]
    assert ev[1] - ev[0] > 1e-6
    E.build_Hv_sector_nonsu2(m, nt)
    try:
        e, vec, nit = E.sp_lanc_eigh(1, 1e-14, vect=gs.copy())  # current state := exact gs
        E.state_store(3)
        dens, docc = E.state_observables(3, m.Norb)
    finally:
        E.delete_Hv_sector_nonsu2()
    rd, ro, _ = N.observables(mo, smap, gs)
    assert np.abs(dens - rd).max() < 1e-10 and np.abs(docc - ro).max() < 1e-10
    wm = np.pi / 50.0 * (2 * np.arange(1, 41) - 1)
    for terms, tq in (([(1.0, +1, 0, 0)], nt + 1), ([(1.0, -1, 1, 1)], nt - 1),
                      ([(1.0, +1, 0, 0), (1j, +1, 1, 1)], nt + 1),
                      ([(1.0, -1, 0, 0), (-1j, -1, 1, 0)], nt - 1)):
        tmap, trp, tcj, tva = N.stored_H(mo, tq)
        ref = oracle_seed(N._c, N._cdg, smap, tmap, gs, terms, Ns)
        E.build_Hv_sector_nonsu2(m, tq)
        try:
            E.apply_Cops(3, [t[0] for t in terms], [t[1] for t in terms], [t[2] for t in terms],
                         [t[3] for t in terms])
            n2 = E.seed_norm2()
            a, b, nused, n2b = E.sp_lanc_tridiag(None, 300)
        finally:
            E.delete_Hv_sector_nonsu2()
        assert abs(n2 - np.vdot(ref, ref).real) < 1e-12 and abs(n2b - n2) < 1e-12
        # continued fraction of the seed against the exact resolvent <ref| (z - (H - e0))^-1 |ref>
        Ht = N.to_dense(trp, tcj, tva)
        evt, Ut = np.linalg.eigh(Ht)
        amp = np.abs(Ut.conj().T @ ref) ** 2
        z = 1j * wm
        exact = (amp[None, :] / (z[:, None] - (evt[None, :] - e0))).sum(axis=1)
        tev, Z = E.tridiag_eigh(a[:nused], b[1:nused])
        got = ((n2 * Z[0, :] ** 2)[None, :] / (z[:, None] - (tev[None, :] - e0))).sum(axis=1)
        assert np.abs(got - exact).max() < 1e-8
    E.state_free(3)


@pytest.mark.parametrize("name", ["normal_superc", "normal_nonsu2"])
def test_ed_diag_c_golden(engine, name):
    """Whole sector scan of ed_diag_c on the device (every Sz / Ntot sector built, solved with
    sp_eigh, state list by the gs_threshold rule) -> evals / dens / docc of the reference fixture."""
    from models import golden, hybrid_nonsu2_model, superc_model

    E = engine
    g = golden(name)
    if name.endswith("superc"):
        import edipack_oracle_superc as S

        m = E.EDModelSuperc(**vars(superc_model(S, name)))
        tol = 5e-8
    else:
        import edipack_oracle_nonsu2 as N

        m = E.EDModelNonsu2(**vars(hybrid_nonsu2_model(N, name)))
        tol = 1e-8
    states = E.ed_diag_c(m)
    assert len(states) == 1
    assert states[0].nup == (0 if name.endswith("superc") else 6)
    assert abs(states[0].e - g["evals"][0]) < 1e-9
    dens, docc = E.observables_packed(m, states)
    assert np.abs(dens - np.array(g["dens"])).max() < tol
    assert np.abs(docc - np.array(g["docc"])).max() < tol
    for s in states:
        E.state_free(s.slot)
