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


@pytest.mark.parametrize("name", ["hybrid_normal", "replica_normal"])
def test_offdiagonal_gf_sigma_momenta_through_gpu(engine, oracle, name):
    """Sigma_momenta.check of the shared / replica bath fixtures: needs the off-diagonal impurity GF,
    i.e. the two-operator seeds (c_a + c_b)|gs> built on the device (edgpu_apply_ops_normal) and
    the orbital-matrix inversion in the host mirror.  1e-8 relative (ed_hybrid_normal.f90:153)."""
    import numpy as np
    from models import golden, replica_normal_kwargs

    E = engine
    g = golden(name)
    kw = hybrid_normal_kwargs() if name == "hybrid_normal" else replica_normal_kwargs("replica")
    m = E.EDModel(**kw)
    m.lanc_tolerance = 1e-18
    states = E.ed_diag_d(m)
    try:
        wm, S = E.get_Sigma_normal(m, states, int(g["inputs"]["LMATS"]), 0)
    finally:
        for s in states:
            E.state_free(s.slot)
    gold = np.array(g["Sigma_momenta"]).reshape(m.Norb, 4)
    for a in range(m.Norb):
        mom = oracle.momenta(wm, S[a, a])
        assert np.abs(mom / gold[a] - 1.0).max() < 1e-8, (name, a, mom, gold[a])


def test_exciton_order_parameters_nonsu2_through_gpu(engine):
    """exciton.check of REPLICA_NONSU2: sector map, stored H, eigen-solver and the two-operator
    seeds (edgpu_apply_ops_packed) all on the device; 1e-8."""
    import numpy as np
    import edipack_oracle_nonsu2 as N
    from models import golden, replica_nonsu2_model

    E = engine
    g = golden("replica_nonsu2")
    mo = replica_nonsu2_model(N, "replica")
    m = E.EDModelNonsu2(**vars(mo))
    states = E.ed_diag_c(m, qns=(5, 6, 7), tol=1e-16)
    try:
        assert len(states) == 1 and states[0].nup == 6
        assert abs(states[0].e - g["evals"][0]) < 1e-9
        got = E.exciton_nonsu2(m, states[0])
    finally:
        for s in states:
            E.state_free(s.slot)
    assert np.abs(got - np.array(g["exciton"])).max() < 1e-8
