"""ED_TWIN: twin_mask, twin_sector_order and the twin branch of
es_return_dvector restated in the oracles; device re-ordering (edgpu_state_twin) against them."""
import numpy as np
import pytest

from models import normal_normal_kwargs, star_kwargs


def test_twin_order_is_the_transpose(oracle):
    """Sorting the flipped integers mdw + mup*2^Ns lists B = (ndw,nup) in its own index order:
    vector_B(iup_B, idw_B) = vec_A(iup_A = idw_B, idw_A = iup_B)."""
    for ns, nup, ndw, nph in [(4, 3, 1, 0), (5, 2, 3, 0), (6, 4, 2, 2), (4, 0, 2, 0), (3, 3, 0, 1)]:
        du, dd = oracle.sector_dims(ns, nup, ndw)
        rng = np.random.default_rng(1)
        v = rng.standard_normal(du * dd * (nph + 1))
        vb = v[oracle.twin_sector_order(ns, nup, ndw, nph)]
        for k in range(nph + 1):
            a = v[k * du * dd:(k + 1) * du * dd].reshape(dd, du)      # [idw_A, iup_A]
            b = vb[k * du * dd:(k + 1) * du * dd].reshape(du, dd)     # [idw_B, iup_B]
            assert np.array_equal(b, a.T)


def test_twin_mask_keeps_nup_ge_ndw(oracle):
    ns = 4
    mask = oracle.twin_mask(ns)
    for isec, on in mask.items():
        nup, ndw = oracle.sector_qn(ns, isec)
        assert on == (nup >= ndw)
    assert sum(mask.values()) == (ns + 1) * (ns + 2) // 2


def doublet_kwargs():
    """Odd electron number in the ground state: degenerate pair of sectors (nup,ndw), (ndw,nup)."""
    return star_kwargs(4)   # Ns = 5 levels, half filling = 5 electrons: doublet (3,2) / (2,3)


def test_ed_twin_reproduces_the_full_scan(oracle):
    m = oracle.Model(**doublet_kwargs())
    full = oracle.diagonalize(m, ed_twin=False)
    twin = oracle.diagonalize(m, ed_twin=True)
    key = lambda s: (s.nup, s.ndw)
    assert sorted(map(key, full)) == sorted(map(key, twin)) and len(full) == 2
    assert full[0].nup != full[0].ndw
    for a in full:
        b = next(s for s in twin if key(s) == key(a))
        assert abs(a.e - b.e) < 1e-12
        assert abs(abs(a.vec @ b.vec) - 1.0) < 1e-10     # same ray (the reference's twin has no sign)
    da, db = oracle.observables(m, full), oracle.observables(m, twin)
    assert np.abs(da[0] - db[0]).max() < 1e-12 and np.abs(da[1] - db[1]).max() < 1e-12


def test_packed_twin_orders():
    import edipack_oracle_nonsu2 as N
    import edipack_oracle_superc as S
    from models import hybrid_nonsu2_model, superc_model

    m = hybrid_nonsu2_model(N)
    ns = m.Ns
    for nt in (4, 5):
        sa, rp, cj, va = N.stored_H(m, nt)
        sb = N.build_sector(ns, 2 * ns - nt)
        order = N.twin_sector_order(ns, nt)
        assert np.array_equal((~sa[order]) & ((1 << (2 * ns)) - 1), sb)
    ms = superc_model(S, "normal_superc")
    ns = ms.Ns
    for sz in (-1, 2):
        sa = S.build_sector(ns, sz)
        sb = S.build_sector(ns, -sz)
        order = S.twin_sector_order(ns, sz)
        lo = (1 << ns) - 1
        assert np.array_equal((sa[order] >> ns) | ((sa[order] & lo) << ns), sb)
        # (the reference applies no fermionic sign when it exchanges the halves: with pairing terms the
        # re-ordered vector is not an eigenvector of the twin sector in general -- only the
        # re-ordering itself is restated and checked)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [("normal_normal", 4, 2), ("star5", 2, 4), ("star5", 6, 0), ("normal_normal", 3, 3)])
def test_state_twin_normal_on_device(engine, oracle, case):
    E = engine
    name, nup, ndw = case
    kw = normal_normal_kwargs() if name == "normal_normal" else star_kwargs(5)
    m = E.EDModel(**kw)
    ns = m.Ns
    du, dd = oracle.sector_dims(ns, nup, ndw)
    v = np.random.default_rng(5).standard_normal(du * dd)
    v /= np.linalg.norm(v)
    E.build_Hv_sector_normal(m, nup, ndw)
    try:
        E.sp_lanc_eigh(1, 1e-14, vect=v.copy())    # current state := v (one-dimensional Krylov space)
        E.state_store(3)
        v3 = E.es_return_vector(3)                 # = v up to the driver's own normalisation (1 ulp)
        assert np.abs(v3 - v).max() < 1e-15
    finally:
        E.delete_Hv_sector_normal()
    E.build_Hv_sector_normal(m, ndw, nup)
    try:
        E.state_twin(3, 4)
        got = E.es_return_vector(4)
    finally:
        E.delete_Hv_sector_normal()
        E.state_free(3)
        E.state_free(4)
    assert np.array_equal(got, v3[oracle.twin_sector_order(ns, nup, ndw)])   # a pure re-ordering: bit-exact


@pytest.mark.gpu
def test_state_twin_with_phonons_on_device(engine, oracle):
    E = engine
    m = E.EDModel(**star_kwargs(3))
    ns, nup, ndw, nph = m.Ns, 3, 1, 2
    du, dd = oracle.sector_dims(ns, nup, ndw)
    v = np.random.default_rng(6).standard_normal(du * dd * (nph + 1))
    v /= np.linalg.norm(v)
    E.set_phonons(Nph=nph, w0=0.3, g=[[0.2]])
    try:
        E.build_Hv_sector_normal(m, nup, ndw)
        E.sp_lanc_eigh(1, 1e-14, vect=v.copy())
        E.state_store(3)
        v = E.es_return_vector(3)
        E.delete_Hv_sector_normal()
        E.build_Hv_sector_normal(m, ndw, nup)
        E.state_twin(3, 4)
        got = E.es_return_vector(4)
        E.delete_Hv_sector_normal()
    finally:
        E.set_phonons(0)
        E.state_free(3)
        E.state_free(4)
    assert np.array_equal(got, v[oracle.twin_sector_order(ns, nup, ndw, nph)])


@pytest.mark.gpu
def test_ed_twin_state_list_on_device(engine, oracle):
    """ed_diag_d with ED_TWIN=T: half of the off-diagonal sectors are solved, the list and the
    observables equal the full scan's (1e-10 / 1e-8)."""
    E = engine
    kw = doublet_kwargs()
    mo = oracle.Model(**kw)
    ref = oracle.diagonalize(mo, ed_twin=False)
    dens_ref, docc_ref = oracle.observables(mo, ref)
    m = E.EDModel(**kw, ed_twin=True, lanc_nstates_sector=1)
    states = E.ed_diag_d(m)
    try:
        assert sorted((s.nup, s.ndw) for s in states) == sorted((s.nup, s.ndw) for s in ref)
        assert all(abs(s.e - ref[0].e) < 1e-10 for s in states)
        dens, docc = E.observables_normal(m, states)
        assert np.abs(dens - dens_ref).max() < 1e-8 and np.abs(docc - docc_ref).max() < 1e-8
        for st in states:       # every listed vector is an eigenvector of its own sector
            H = oracle.dense_H(mo, st.nup, st.ndw)
            E.build_Hv_sector_normal(m, st.nup, st.ndw)
            try:
                v = E.es_return_vector(st.slot)
            finally:
                E.delete_Hv_sector_normal()
            assert np.abs(H @ v - st.e * v).max() < 1e-7
    finally:
        for s in states:
            E.state_free(s.slot)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["nonsu2", "superc"])
def test_state_twin_packed_on_device(engine, mode):
    import edipack_oracle_nonsu2 as N
    import edipack_oracle_superc as S
    from models import hybrid_nonsu2_model, superc_model

    E = engine
    if mode == "nonsu2":
        mo = hybrid_nonsu2_model(N)
        m, q, qt = E.EDModelNonsu2(**vars(mo)), 4, 2 * mo.Ns - 4
        build, order = E.build_Hv_sector_nonsu2, N.twin_sector_order(mo.Ns, q)
    else:
        mo = superc_model(S, "normal_superc")
        m, q, qt = E.EDModelSuperc(**vars(mo)), -1, 1
        build, order = E.build_Hv_sector_superc, S.twin_sector_order(mo.Ns, q)
    build(m, q)
    try:
        ev, vecs, nconv, _ = E.sp_eigh(1, 20, 300, 0.0)
        E.eigh_state_store(0, 3)
        v = E.es_return_vector(3)
        assert np.abs(v - vecs[:, 0]).max() < 1e-15
    finally:
        E.delete_Hv_sector_csr()
    build(m, qt)
    try:
        E.state_twin(3, 4)
        got = E.es_return_vector(4)
    finally:
        E.delete_Hv_sector_csr()
        E.state_free(3)
        E.state_free(4)
    assert np.array_equal(got, v[order])
