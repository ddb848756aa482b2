"""local_energy_normal through the device (host mirror: every component is alpha_1 = <v|H_variant|v>
of the stored ground state) against the reference's energy.check / doubles.check.  Sorted last on
purpose: it composes calls that are each covered by earlier files."""
import pytest

from models import hybrid_normal_kwargs, normal_normal_kwargs
from test_local_energy import check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,kwf", [("normal_normal", normal_normal_kwargs), ("hybrid_normal", hybrid_normal_kwargs)])
def test_local_energy_through_gpu(engine, name, kwf):
    E = engine
    m = E.EDModel(**kwf())
    m.lanc_tolerance = 1e-18  # the reference's LANC_TOLERANCE default, as in test_gpu_parity.py
    states = E.ed_diag_d(m)
    try:
        check(name, E.local_energy_normal(m, states))
    finally:
        for s in states:
            E.state_free(s.slot)
