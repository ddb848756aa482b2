"""Ground-state driver with the Lanczos vectors kept in HBM (pass 2 = linear combination) against
the two-vector replay mode of SciFortran's sp_lanc_eigh: same energy, same vector."""
import os

import numpy as np
import pytest

from models import star_kwargs, two_orb_kwargs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kw,sec", [(star_kwargs(9), (5, 5)), (two_orb_kwargs(3), (4, 4))])
def test_store_equals_replay(engine, oracle, kw, sec):
    E = engine
    m = E.EDModel(**kw)
    E.build_Hv_sector_normal(m, *sec)
    try:
        os.environ["EDGPU_LANCZOS_STORE"] = "0"
        e0, v0, n0 = E.sp_lanc_eigh(300, 1e-14)
        s0, h0 = E.lanczos_last_info()
        os.environ["EDGPU_LANCZOS_STORE"] = "1"
        e1, v1, n1 = E.sp_lanc_eigh(300, 1e-14)
        s1, h1 = E.lanczos_last_info()
    finally:
        os.environ.pop("EDGPU_LANCZOS_STORE", None)
        E.delete_Hv_sector_normal()
    assert n0 == n1 and e0 == e1
    assert s0 == 0 and h0 == 2 * n0 - 1
    assert s1 == n1 and h1 == n1          # every vector kept: no second run of the recurrence
    assert np.abs(v0 - v1).max() < 1e-13
    assert abs(np.linalg.norm(v1) - 1.0) < 1e-12


@pytest.mark.parametrize("keep", [2, 3, 7, 20])
def test_partial_store_equals_full_store(engine, keep):
    """Device memory running out mid-way (emulated: EDGPU_LANCZOS_MAXSTORE): the first `keep` vectors
    stay in their slots, the recurrence continues on the scratch pair and pass 2 replays only the
    unstored tail -- same energy, same vector, nlanc - keep extra products."""
    E = engine
    m = E.EDModel(**star_kwargs(9))
    E.build_Hv_sector_normal(m, 5, 5)
    try:
        e1, v1, n1 = E.sp_lanc_eigh(300, 1e-14)
        os.environ["EDGPU_LANCZOS_MAXSTORE"] = str(keep)
        e2, v2, n2 = E.sp_lanc_eigh(300, 1e-14)
        s2, h2 = E.lanczos_last_info()
    finally:
        os.environ.pop("EDGPU_LANCZOS_MAXSTORE", None)
        E.delete_Hv_sector_normal()
    assert n1 == n2 and abs(e1 - e2) < 1e-13
    assert s2 == keep and h2 == n2 + (n2 - keep)
    assert np.abs(v1 - v2).max() < 1e-12
