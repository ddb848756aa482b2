"""ED_SPARSE_H = F for the packed-state modes: direct on-the-fly H x v of nonsu2 / superc sectors
(directMatVec_nonsu2_main / directMatVec_superc_main, ED_HAMILTONIAN_NONSU2_DIRECT_HxV.f90:22-252,
ED_HAMILTONIAN_SUPERC_DIRECT_HxV.f90:22-311).  Nothing but the sector map is stored on the device.
Bars: H x v 1e-12 against the oracle's stored rows AND against the device's own stored CSR of the
same sector; ground-state energies of the reference's fixtures 1e-9 through the direct path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("idx", [0, 1, 2, 3, 4, 5])
def test_nonsu2_direct_matches_oracle_and_stored(engine, idx):
    from test_gpu_nonsu2 import cases, to_engine_model

    E = engine
    N, cs = cases()
    name, mo, ntots = cs[idx]
    m = to_engine_model(E, mo)
    rng = np.random.default_rng(8)
    for nt in ntots:
        smap, rp, cj, va = N.stored_H(mo, nt)
        v = rng.standard_normal(len(smap)) + 1j * rng.standard_normal(len(smap))
        ref = N.csr_matvec(rp, cj, va, v)
        E.build_Hv_sector_nonsu2(m, nt)
        try:
            hv_stored = E.spHtimesV_cc(v)
        finally:
            E.delete_Hv_sector_nonsu2()
        E.set_sparse_H(False)
        try:
            E.build_Hv_sector_nonsu2(m, nt)
            try:
                assert np.array_equal(E.sector_map_nonsu2(), smap.astype(np.int32))
                hv = E.spHtimesV_cc(v)
                with pytest.raises(E.EdgpuError, match="direct"):
                    E.stored_csr()
            finally:
                E.delete_Hv_sector_nonsu2()
        finally:
            E.set_sparse_H(True)
        assert rel(hv, ref) < 1e-12, (name, nt)
        assert rel(hv, hv_stored) < 1e-12, (name, nt)


@pytest.mark.parametrize("name", ["normal_superc", "hybrid_superc", "messy", "replica_superc", "messy_replica"])
def test_superc_direct_matches_oracle(engine, name):
    import edipack_oracle_nonsu2 as N
    import edipack_oracle_superc as S
    from models import replica_superc_model, superc_model
    from test_gpu_superc import messy_replica_superc, messy_superc, to_engine_model

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
    rng = np.random.default_rng(6)
    E.set_sparse_H(False)
    try:
        for sz in (0, 1, -2, mo.Ns):
            smap, rp, cj, va = S.stored_H(mo, sz)
            v = rng.standard_normal(len(smap)) + 1j * rng.standard_normal(len(smap))
            ref = N.csr_matvec(rp, cj, va, v)
            E.build_Hv_sector_superc(m, sz)
            try:
                hv = E.spHtimesV_cc(v)
            finally:
                E.delete_Hv_sector_superc()
            assert rel(hv, ref) < 1e-12, (name, sz)
    finally:
        E.set_sparse_H(True)


@pytest.mark.parametrize("name", ["hybrid_nonsu2", "replica_nonsu2"])
def test_golden_nonsu2_through_direct_path(engine, name):
    """test/src/{HYBRID,REPLICA}_NONSU2/evals.check with ED_SPARSE_H=F: the reference runs every
    fixture with both settings (test/test.sh:74-78)."""
    import edipack_oracle_nonsu2 as N
    from models import golden, hybrid_nonsu2_model, replica_nonsu2_model
    from test_gpu_nonsu2 import to_engine_model

    E = engine
    g = golden(name)
    mo = replica_nonsu2_model(N, "replica") if name.startswith("replica") else hybrid_nonsu2_model(N, name)
    m = to_engine_model(E, mo)
    E.set_sparse_H(False)
    try:
        best = None
        for nt in (5, 6, 7):
            E.build_Hv_sector_nonsu2(m, nt)
            try:
                ev, _, _, _ = E.sp_eigh(1, 20, 512, 1e-16, want_vectors=False)
            finally:
                E.delete_Hv_sector_nonsu2()
            best = ev[0] if best is None else min(best, ev[0])
    finally:
        E.set_sparse_H(True)
    assert abs(best - g["evals"][0]) < 1e-9


def test_cfg5_direct_vs_stored(engine):
    """BASELINE config 5 itself (705 432 states): the direct product equals the stored one to
    1e-12 and keeps nothing but the 2.8 MB sector map on the device."""
    import edipack_oracle_nonsu2 as N
    from test_gpu_baseline_configs import cfg5_hloc
    from test_gpu_nonsu2 import to_engine_model

    E = engine
    mo = N.ModelNonsu2(Norb=3, Nbath=8, bath_type="hybrid", Uloc=(2.0, 2.0, 2.0), Ust=1.5, Jh=0.25,
                       Jx=0.25, Jp=0.25, hfmode=True, hloc=cfg5_hloc())
    mo.default_bath()
    m = to_engine_model(E, mo)
    n = 705432
    rng = np.random.default_rng(2)
    v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    E.build_Hv_sector_nonsu2(m, 11)
    try:
        hv_stored = E.spHtimesV_cc(v)
    finally:
        E.delete_Hv_sector_nonsu2()
    E.set_sparse_H(False)
    try:
        E.build_Hv_sector_nonsu2(m, 11)
        try:
            hv = E.spHtimesV_cc(v)
        finally:
            E.delete_Hv_sector_nonsu2()
    finally:
        E.set_sparse_H(True)
    assert rel(hv, hv_stored) < 1e-12
