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
