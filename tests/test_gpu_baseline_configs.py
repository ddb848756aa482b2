"""Parity at the BASELINE configs' OWN sizes (BASELINE.json `configs`), through the C ABI:

  cfg2  Norb=1 Nbath=15 (Ns=16), sector (8,8), 165 636 900 states: H x v 1e-12 vs the oracle's stored
        product, ground state 1e-10 vs the oracle's Lanczos (recurrence of sp_lanc_eigh in C)
  cfg3  Norb=2 Nbath=6 (Ns=14), sector (7,7), U=U'=2, J=Jx=Jp=0.125: H x v 1e-12, E_gs 1e-10
  cfg5  nonsu2 Norb=3, hybrid bath Nbath=8 (22 levels), spin-orbit Hloc, N=11 (705 432 states):
        device-built spH0 rows vs the oracle's rows on a seeded sample of rows (the per-row Python
        loops of the oracle over all rows take tens of minutes), H x v 1e-12 on those rows,
        hermiticity on the whole vector
  exc_field != 0 (the excitonic fields of stored/H_up.f90:85-103) on a two-orbital model

Reference loops: ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:236-375, ED_HAMILTONIAN_NONSU2_STORED_HxV.f90.
"""
import json
import os

import numpy as np
import pytest

from models import star_kwargs, two_orb_kwargs

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def ncores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------
# cfg2
# ----------------------------------------------------------------------------------------------
def test_cfg2_hxv_and_ground_state(engine, oracle):
    E, O = engine, oracle
    kw = star_kwargs(15)
    m, mo = E.EDModel(**kw), O.Model(**kw)
    nup = ndw = 8
    du, dd = O.sector_dims(16, nup, ndw)
    assert du * dd == 165636900
    nt = ncores()
    v = O.start_vector(du * dd, 77) - 0.5
    ref = O.stored_hxv_mpi(mo, nup, ndw, v, nt, nt)[0]
    with open(os.path.join(HERE, "golden", "baseline_cfg2.json")) as f:
        g = json.load(f)
    E.build_Hv_sector_normal(m, nup, ndw)
    try:
        assert E.vecDim_Hv_sector_normal() == du * dd
        for variant in (2, 1):  # tiled kernels (k_fastb + k_slow), generic gather kernel
            E.set_kernel_variant(variant)
            hv = E.spHtimesV_p(v)
            assert rel_err(hv, ref) < 1e-12, variant
            del hv
        E.set_kernel_variant(0)
        del ref
        # Lanczos coefficients of the first steps and the ground-state energy against the oracle's
        # recurrence from the same seeded start vector (tests/golden/baseline_cfg2.json)
        a, b, nused, n2 = E.sp_lanc_tridiag(O.start_vector(du * dd, g["seed"]), 12)
        assert np.abs(a[:12] - np.array(g["alanc"][:12])).max() < 1e-9
        assert np.abs(b[1:12] - np.array(g["blanc"][1:12])).max() < 1e-9
        e, x, nit = E.sp_lanc_eigh(300, g["threshold"], ncheck=g["ncheck"], seed=g["seed"])
        assert abs(e - g["egs"]) < 1e-10, (e, g["egs"])
        assert nit == g["niter"]
        if os.environ.get("EDGPU_LIVE_ORACLE_LANCZOS"):
            e_live = O.stored_lanczos_gs(mo, nup, ndw, O.start_vector(du * dd, g["seed"]), 300,
                                         g["threshold"], g["ncheck"], P=nt, nthreads=nt)[0]
            assert abs(e - e_live) < 1e-10
    finally:
        E.set_kernel_variant(0)
        E.delete_Hv_sector_normal()
    # the returned vector is an eigenpair of the ORACLE's operator: Rayleigh quotient and residual
    hx = O.stored_hxv_mpi(mo, nup, ndw, x, nt, nt)[0]
    assert abs(x @ x - 1.0) < 1e-10
    assert abs(x @ hx - e) < 1e-10
    assert np.linalg.norm(hx - e * x) < 2e-5  # sp_lanc_eigh stops on the Ritz VALUE (1e-12)


# ----------------------------------------------------------------------------------------------
# cfg3
# ----------------------------------------------------------------------------------------------
def test_cfg3_hxv_and_ground_state(engine, oracle):
    E, O = engine, oracle
    kw = two_orb_kwargs(6)  # Norb=2, Nbath=6, U=U'=2, Jh=Jx=Jp=0.125, Hloc = 0.5 sigma_z
    assert kw["Jx"] == 0.125 and kw["Jp"] == 0.125
    m, mo = E.EDModel(**kw), O.Model(**kw)
    assert m.Ns == 14
    nup = ndw = 7
    du, dd = O.sector_dims(14, nup, ndw)
    assert du * dd == 11778624
    nt = ncores()
    v = O.start_vector(du * dd, 5) - 0.5
    ref = O.stored_hxv_mpi(mo, nup, ndw, v, nt, nt)[0]
    v0 = O.start_vector(du * dd, 4321)
    e_ref, nit_ref, a0, b0, _ = O.stored_lanczos_gs(mo, nup, ndw, v0, 300, 1e-12, P=nt, nthreads=nt)
    E.build_Hv_sector_normal(m, nup, ndw)
    try:
        for variant in (2, 1):
            E.set_kernel_variant(variant)
            assert rel_err(E.spHtimesV_p(v), ref) < 1e-12, variant
        E.set_kernel_variant(0)
        a, b, nused, n2 = E.sp_lanc_tridiag(v0, 12)
        assert np.abs(a[:12] - a0[:12]).max() < 1e-9 and np.abs(b[1:12] - b0[1:12]).max() < 1e-9
        e, x, nit = E.sp_lanc_eigh(300, 1e-12)
        assert abs(e - e_ref) < 1e-10, (e, e_ref)
    finally:
        E.set_kernel_variant(0)
        E.delete_Hv_sector_normal()
    hx = O.stored_hxv_mpi(mo, nup, ndw, x, nt, nt)[0]
    assert abs(x @ hx - e) < 1e-10


# ----------------------------------------------------------------------------------------------
# cfg5
# ----------------------------------------------------------------------------------------------
def cfg5_hloc():
    """On-site spin-orbit coupling lambda L.S + crystal field in the t2g (yz, zx, xy) basis."""
    lam, cf = 0.2, 0.1
    Lx = np.array([[0, 0, 0], [0, 0, -1j], [0, 1j, 0]])
    Ly = np.array([[0, 0, 1j], [0, 0, 0], [-1j, 0, 0]])
    Lz = np.array([[0, -1j, 0], [1j, 0, 0], [0, 0, 0]])
    sx = 0.5 * np.array([[0, 1], [1, 0]])
    sy = 0.5 * np.array([[0, -1j], [1j, 0]])
    sz = 0.5 * np.array([[1, 0], [0, -1]])
    hloc = np.zeros((2, 2, 3, 3), complex)
    for s in range(2):
        for t in range(2):
            hloc[s, t] = lam * (Lx * sx[s, t] + Ly * sy[s, t] + Lz * sz[s, t])
        hloc[s, s] += np.diag([cf, 0.0, -cf])
    return hloc


def test_cfg5_nonsu2_device_csr_and_spmv(engine):
    import edipack_oracle_nonsu2 as N
    from test_gpu_nonsu2 import to_engine_model

    E = engine
    mo = N.ModelNonsu2(Norb=3, Nbath=8, bath_type="hybrid", Uloc=(2.0, 2.0, 2.0), Ust=1.5, Jh=0.25,
                       Jx=0.25, Jp=0.25, hfmode=True, hloc=cfg5_hloc())
    mo.default_bath()
    rng = np.random.default_rng(9)
    mo.bath_u = 0.1 + 0.1 * rng.random(mo.bath_u.shape)  # spin-flip hybridisation switched on
    m = to_engine_model(E, mo)
    ntot = 11
    n = 705432
    rows = np.unique(np.concatenate([[0, 1, n // 2, n - 2, n - 1], rng.integers(0, n, 1500)]))
    smap, rp, cj, va = N.stored_H(mo, ntot, only_rows=rows)
    assert len(smap) == n
    v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    w = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    E.build_Hv_sector_nonsu2(m, ntot)
    try:
        assert np.array_equal(E.sector_map_nonsu2(), smap.astype(np.int32))
        drp, dcj, dva = E.stored_csr()
        hv = E.spHtimesV_cc(v)
        hw = E.spHtimesV_cc(w)
    finally:
        E.delete_Hv_sector_nonsu2()
    nnz_dev = int(drp[-1])
    assert 30 * n < nnz_dev < 80 * n
    scale = 0.0
    for k, r in enumerate(rows):
        oc, ov = cj[rp[k]:rp[k + 1]], va[rp[k]:rp[k + 1]]
        dc, dv = dcj[drp[r]:drp[r + 1]], dva[drp[r]:drp[r + 1]]
        # same columns in the same (insertion) order, values to FMA-contraction rounding
        assert np.array_equal(oc, dc), r
        assert np.abs(ov - dv).max() < 1e-14, r
        ref = np.sum(ov * v[oc - 1])
        scale = max(scale, abs(ref))
        assert abs(hv[r] - ref) < 1e-12 * max(np.abs(hv).max(), 1.0), r
    assert scale > 0.0
    # hermiticity on the whole 705 432-vector: <w|Hv> = conj(<v|Hw>)
    assert abs(np.vdot(w, hv) - np.conj(np.vdot(v, hw))) < 1e-11 * abs(np.vdot(w, hv))


# ----------------------------------------------------------------------------------------------
# exc_field
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [1, 2])
def test_exc_field_hxv(engine, oracle, variant):
    """Excitonic fields F_0, F_z (ED_INPUT_VARS exc_field(1), exc_field(4)): the stored path's loop
    bounds (stored/H_up.f90:85-103, H_dw.f90), which is what the product's term list follows."""
    E, O = engine, oracle
    rng = np.random.default_rng(21)
    for nb, secs in ((2, ((3, 3), (2, 4), (4, 1))), (4, ((5, 5), (4, 6)))):
        kw = two_orb_kwargs(nb)
        kw["exc_field"] = (0.3, 0.0, 0.0, -0.2)
        m, mo = E.EDModel(**kw), O.Model(**kw)
        E.set_kernel_variant(variant)
        try:
            for nup, ndw in secs:
                du, dd = O.sector_dims(m.Ns, nup, ndw)
                v = rng.standard_normal(du * dd)
                ref = O.stored_hxv(mo, nup, ndw, v)
                # the field really contributes
                kw0 = dict(kw, exc_field=(0.0, 0.0, 0.0, 0.0))
                assert rel_err(O.stored_hxv(O.Model(**kw0), nup, ndw, v), ref) > 1e-3
                E.build_Hv_sector_normal(m, nup, ndw)
                try:
                    hv = E.spHtimesV_p(v)
                finally:
                    E.delete_Hv_sector_normal()
                assert rel_err(hv, ref) < 1e-12, (nb, nup, ndw)
        finally:
            E.set_kernel_variant(0)
