"""Small host-facing additions of round 2 through the C ABI: page-locking of caller-owned arrays
(edgpu_host_register), the in-place driver call (the reference's intent(inout) vect), the sector's
communication info on one rank, ED_SPARSE_H toggling."""
import ctypes as C

import numpy as np
import pytest

from models import star_kwargs

pytestmark = pytest.mark.gpu


def test_host_register_and_inplace_driver(engine, oracle):
    E, O = engine, oracle
    from edipack_b200 import _abi

    L = _abi.load()
    kw = star_kwargs(9)
    m, mo = E.EDModel(**kw), O.Model(**kw)
    du, dd = O.sector_dims(10, 5, 5)
    n = du * dd
    v = O.start_vector(n, 3) - 0.5
    hv = np.empty(n)
    ref = O.direct_hxv(mo, 5, 5, v)
    E.build_Hv_sector_normal(m, 5, 5)
    try:
        assert E.sector_comm_info()[0] == 0  # single rank: no exchange
        E.host_register(v)
        E.host_register(hv)
        try:
            n32 = C.c_int32(n)
            L.edgpu_hxv_d(C.byref(n32), v.ctypes.data_as(C.c_void_p), hv.ctypes.data_as(C.c_void_p))
            _abi.check(L.edgpu_status())
        finally:
            E.host_unregister(v)
            E.host_unregister(hv)
        assert np.abs(hv - ref).max() < 1e-12 * np.abs(ref).max()
        # in-place driver call: the start vector is overwritten by the eigenvector
        e0, x0, n0 = E.sp_lanc_eigh(300, 1e-14, vect=v)          # copy semantics: v untouched
        assert np.array_equal(v, O.start_vector(n, 3) - 0.5)
        w = v.copy()
        e1, x1, n1 = E.sp_lanc_eigh(300, 1e-14, vect=w, inplace=True)
        assert x1 is w and abs(e1 - e0) < 1e-13 and np.abs(x1 - x0).max() < 1e-12
        assert abs(np.linalg.norm(w) - 1.0) < 1e-12
    finally:
        E.delete_Hv_sector_normal()
    with pytest.raises(E.EdgpuError):
        E.sector_comm_info()
