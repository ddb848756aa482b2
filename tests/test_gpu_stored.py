"""Stored-H (CSR) path on the GPU, through the C ABI: real matrices from the oracle's NORMAL-mode
Hamiltonian (the content of spH0d + 1 (x) spH0ups + spH0dws (x) 1 + spH0nd), complex Hermitian
matrices with the reference's storage quirks (unsorted rows, duplicate entries that add up,
empty rows), and the Lanczos drivers on top.  Bars: H x v 1e-12 relative, E_gs 1e-10."""
import numpy as np
import pytest

from models import messy_kwargs, normal_normal_kwargs, star_kwargs, two_orb_kwargs

pytestmark = pytest.mark.gpu


def dense_to_ref_csr(H, rng, dup_frac=0.2):
    """Dense -> (rowptr, cols 1-based, vals) the way sp_insert_element leaves it: insertion order
    (shuffled), and some entries split into two inserts of the same (i,j) stored separately."""
    n = H.shape[0]
    rowptr, cols, vals = [0], [], []
    for i in range(n):
        js = np.nonzero(H[i])[0]
        ent = []
        for j in js:
            if rng.random() < dup_frac:
                a = H[i, j] * rng.random()
                ent += [(j + 1, a), (j + 1, H[i, j] - a)]
            else:
                ent.append((j + 1, H[i, j]))
        order = rng.permutation(len(ent))
        for k in order:
            cols.append(ent[k][0])
            vals.append(ent[k][1])
        rowptr.append(len(cols))
    return np.array(rowptr, np.int64), np.array(cols, np.int32), np.array(vals, H.dtype)


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("kw,sec", [(normal_normal_kwargs(), (3, 3)), (star_kwargs(7), (4, 4)),
                                    (messy_kwargs(), (3, 2)), (two_orb_kwargs(3), (4, 4)),
                                    (star_kwargs(5), (0, 2)), (star_kwargs(5), (6, 6))])
def test_real_stored_matches_oracle(engine, oracle, kw, sec):
    E = engine
    mo = oracle.Model(**kw)
    H = oracle.dense_H(mo, *sec)
    rng = np.random.default_rng(3)
    rp, cj, va = dense_to_ref_csr(H, rng)
    v = rng.standard_normal(H.shape[0])
    ref = oracle.direct_hxv(mo, sec[0], sec[1], v)
    E.build_Hv_sector_csr(rp, cj, va)
    try:
        assert E.vecDim_Hv_sector_normal() == H.shape[0]
        hv = E.spHtimesV_p(v)
        assert rel_err(hv, ref) < 1e-12
        if H.shape[0] > 1:
            e, vec, nit = E.sp_lanc_eigh(min(H.shape[0], 300), 1e-14)
            ev = np.linalg.eigvalsh(H)
            assert abs(e - ev[0]) < 1e-10
            assert abs(np.linalg.norm(vec) - 1.0) < 1e-12
            with pytest.raises(E.EdgpuError, match="real"):
                E.spHtimesV_cc(v.astype(complex))
    finally:
        E.delete_Hv_sector_csr()


@pytest.mark.parametrize("n,density", [(1, 1.0), (37, 0.3), (924, 0.05), (5000, 0.004)])
def test_complex_stored_hermitian(engine, n, density):
    E = engine
    rng = np.random.default_rng(n)
    A = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) * (rng.random((n, n)) < density)
    H = A + A.conj().T + np.diag(rng.standard_normal(n))
    if n > 10:
        H[5, :] = 0.0
        H[:, 5] = 0.0  # an empty row
    rp, cj, va = dense_to_ref_csr(H, rng)
    v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    E.build_Hv_sector_csr(rp, cj, va)
    try:
        hv = E.spHtimesV_cc(v)
        assert rel_err(hv, H @ v) < 1e-12
        # linearity with complex coefficients
        w = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        lhs = E.spHtimesV_cc((2 - 1j) * v + 0.5j * w)
        assert rel_err(lhs, (2 - 1j) * hv + 0.5j * E.spHtimesV_cc(w)) < 1e-12
        if n > 1:
            e, vec, nit = E.sp_lanc_eigh(min(n, 400), 1e-14)
            ev = np.linalg.eigvalsh(H)
            assert abs(e - ev[0]) < 1e-9 * max(1.0, abs(ev[0]))
            assert np.abs(H @ vec - e * vec).max() < 1e-5
            # tridiagonalisation: alpha_1 = <s|H|s>/<s|s>, beta_1 = |H s - alpha s| / |s|
            a, b, nused, n2 = E.sp_lanc_tridiag(v, 3)
            assert abs(n2 - np.vdot(v, v).real) < 1e-12 * n2
            s = v / np.sqrt(n2)
            a1 = np.vdot(s, H @ s).real
            assert abs(a[0] - a1) < 1e-11 * max(1.0, abs(a1))
            if nused > 1:
                assert abs(b[1] - np.linalg.norm(H @ s - a1 * s)) < 1e-10
    finally:
        E.delete_Hv_sector_csr()


def test_csr_error_behaviour(engine):
    E = engine
    rp = np.array([0, 1, 2], np.int64)
    with pytest.raises(E.EdgpuError, match="out of range"):
        E.build_Hv_sector_csr(rp, np.array([1, 3], np.int32), np.array([1.0, 2.0]))
    E.build_Hv_sector_csr(rp, np.array([1, 2], np.int32), np.array([1.0, 2.0]))
    try:
        with pytest.raises(E.EdgpuError, match="Nloc"):
            E.spHtimesV_p(np.zeros(3))
        assert np.allclose(E.spHtimesV_p(np.array([1.0, 1.0])), [1.0, 2.0])
    finally:
        E.delete_Hv_sector_csr()


def test_golden_hybrid_nonsu2_through_gpu(engine):
    """The reference's HYBRID_NONSU2 fixture: the host (numpy oracle standing in for the Fortran
    ed_buildH_nonsu2_main) builds spH0, the GPU does the complex stored-H Lanczos (the reference
    uses ARPACK + spMatVec_nonsu2_main here, LANC_DIM_THRESHOLD=256 < 924) -> evals 1e-9,
    dens / docc / magX 1e-8 against test/src/HYBRID_NONSU2/*.check."""
    import edipack_oracle_nonsu2 as N
    from models import golden, hybrid_nonsu2_model

    E = engine
    g = golden("hybrid_nonsu2")
    m = hybrid_nonsu2_model(N)
    smap, rp, cj, va = N.stored_H(m, 6)
    E.build_Hv_sector_csr(rp, cj, va)
    try:
        e, vec, nit = E.sp_lanc_eigh(512, 1e-14)
    finally:
        E.delete_Hv_sector_csr()
    assert abs(e - g["evals"][0]) < 1e-9
    dens, docc, magx = N.observables(m, smap, vec)
    assert np.abs(dens - np.array(g["dens"])).max() < 1e-8
    assert np.abs(docc - np.array(g["docc"])).max() < 1e-8
    assert np.abs(magx - np.array(g["magX"])).max() < 1e-8


def test_nonsu2_soc_stored_hxv(engine):
    """cfg5 family (complex Hloc with spin-flip terms, spin field, spin-flip hybridisation) at
    3432 states: H x v of the complex CSR kernel vs the oracle's sp_matvec, GS energy vs LAPACK."""
    import edipack_oracle_nonsu2 as N
    from models import soc_nonsu2_model

    E = engine
    m = soc_nonsu2_model(N, nbath=5)
    smap, rp, cj, va = N.stored_H(m, 7)
    assert len(smap) == 3432
    rng = np.random.default_rng(9)
    v = rng.standard_normal(len(smap)) + 1j * rng.standard_normal(len(smap))
    ref = N.csr_matvec(rp, cj, va, v)
    E.build_Hv_sector_csr(rp, cj, va)
    try:
        hv = E.spHtimesV_cc(v)
        assert rel_err(hv, ref) < 1e-12
        e, vec, nit = E.sp_lanc_eigh(600, 1e-14)
    finally:
        E.delete_Hv_sector_csr()
    ev = np.linalg.eigvalsh(N.to_dense(rp, cj, va))
    assert abs(e - ev[0]) < 1e-10
