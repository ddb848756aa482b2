"""Pins the CPU oracle's hybrid / replica / general bath terms (hop tables of 1 (x) Hup and
Hdw (x) 1 with inter-orbital bath hops, ED_NORMAL/direct/HxV_up.f90:33-59) to the reference's
golden vectors test/src/{HYBRID,REPLICA,GENERAL}_NORMAL/{evals,dens,docc}.check at the reference's
own tolerance (1e-9, test/src/ASSERTING.f90:78)."""
import numpy as np
import pytest

from models import golden, hybrid_normal_kwargs, replica_normal_kwargs

FIXTURES = {
    "hybrid_normal": hybrid_normal_kwargs,
    "replica_normal": lambda: replica_normal_kwargs("replica"),
    "general_normal": lambda: replica_normal_kwargs("general"),
}


@pytest.mark.parametrize("name", list(FIXTURES))
def test_bath_fixture(oracle, name):
    g = golden(name)
    m = oracle.Model(**FIXTURES[name]())
    states = oracle.diagonalize(m)
    assert len(states) == 1
    assert abs(states[0].e - g["evals"][0]) < 1e-9
    dens, docc = oracle.observables(m, states)
    assert np.abs(dens - np.array(g["dens"])).max() < 1e-9
    assert np.abs(docc - np.array(g["docc"])).max() < 1e-9
    # direct (on-the-fly) and stored element generators agree on this bath type
    v = oracle.start_vector(400, 3) - 0.5
    a = oracle.direct_hxv(m, 3, 3, v)
    b = oracle.stored_hxv(m, 3, 3, v)
    assert np.abs(a - b).max() < 1e-13 * np.abs(a).max()
