"""Unit-level known-answer tests of the oracle (the reference has none at this level,
SURVEY F7): map bijection, anticommutation signs, direct == stored == dense, MPI emulation
== serial, hermiticity, edge sectors."""
import ctypes as C

import numpy as np
import pytest

from models import messy_kwargs, normal_normal_kwargs, replica_kwargs, star_kwargs, two_orb_kwargs

MODELS = {
    "normal_normal": normal_normal_kwargs,
    "star5": lambda: star_kwargs(5),
    "two_orb": lambda: two_orb_kwargs(2),
    "messy": messy_kwargs,
    "messy_hybrid": lambda: messy_kwargs("hybrid"),
    "replica": replica_kwargs,
}


def test_map_sorted_bijection(oracle):
    for ns, nel in [(6, 3), (8, 0), (8, 8), (10, 4), (1, 1)]:
        m = oracle.build_map(ns, nel)
        assert len(m) == oracle.binomial(ns, nel)
        assert np.all(np.diff(m) > 0)
        assert all(bin(int(x)).count("1") == nel for x in m)
        L = oracle.lib()
        for r in (0, len(m) // 2, len(m) - 1):
            assert L.ora_binary_search(m, len(m), int(m[r])) == r + 1
        assert L.ora_binary_search(m, len(m), -1) == 0


def test_c_cdg_anticommutation(oracle):
    """{c_i, c^+_j} = delta_ij on every 6-bit state with the reference's sign convention."""
    L = oracle.lib()

    def apply(op, pos, state):  # returns (sign, newstate) or None
        out, sg = C.c_int32(), C.c_double()
        ok = (L.ora_c if op == "c" else L.ora_cdg)(pos, state, C.byref(out), C.byref(sg))
        return (sg.value, out.value) if ok else None

    n = 6
    for state in range(2 ** n):
        for i in range(1, n + 1):
            for j in range(1, n + 1):
                acc = {}
                for first, second in ((("cdg", j), ("c", i)), (("c", i), ("cdg", j))):
                    r1 = apply(first[0], first[1], state)
                    if r1 is None:
                        continue
                    r2 = apply(second[0], second[1], r1[1])
                    if r2 is None:
                        continue
                    acc[r2[1]] = acc.get(r2[1], 0.0) + r1[0] * r2[0]
                for k, val in acc.items():
                    assert val == (1.0 if (i == j and k == state) else 0.0)


@pytest.mark.parametrize("name", list(MODELS))
def test_direct_stored_dense_agree(oracle, name):
    m = oracle.Model(**MODELS[name]())
    ns = m.Ns
    rng = np.random.default_rng(3)
    for nup, ndw in [(ns // 2, ns // 2), (ns // 2 + 1, ns // 2 - 1), (1, ns - 1), (0, 2), (ns, 0)]:
        H = oracle.dense_H(m, nup, ndw)
        assert np.abs(H - H.T).max() == 0.0
        v = rng.standard_normal(H.shape[0])
        ref = H @ v
        scale = max(np.abs(ref).max(), 1.0)
        assert np.abs(oracle.direct_hxv(m, nup, ndw, v) - ref).max() < 1e-13 * scale
        assert np.abs(oracle.stored_hxv(m, nup, ndw, v) - ref).max() < 1e-13 * scale


@pytest.mark.parametrize("P", [1, 2, 3, 5])
def test_mpi_emulation_matches_serial(oracle, P):
    m = oracle.Model(**normal_normal_kwargs())
    rng = np.random.default_rng(5)
    for nup, ndw in [(3, 3), (2, 4), (4, 1)]:
        du, dd = oracle.sector_dims(m.Ns, nup, ndw)
        v = rng.standard_normal(du * dd)
        ref = oracle.direct_hxv(m, nup, ndw, v)
        assert np.abs(oracle.direct_hxv_mpi(m, nup, ndw, v, P, 2) - ref).max() < 1e-13
        assert np.abs(oracle.stored_hxv_mpi(m, nup, ndw, v, P, 2)[0] - ref).max() < 1e-13


def test_apply_op_norms(oracle):
    """<v|n_a|v> = |c_a v|^2 and |c_a v|^2 + |c^+_a v|^2 = |v|^2."""
    m = oracle.Model(**normal_normal_kwargs())
    rng = np.random.default_rng(9)
    nup, ndw = 3, 2
    du, dd = oracle.sector_dims(m.Ns, nup, ndw)
    v = rng.standard_normal(du * dd)
    v /= np.linalg.norm(v)
    for spin in (0, 1):
        for iorb in range(m.Norb):
            c, _ = oracle.apply_op(m, -1, iorb, spin, nup, ndw, v)
            cd, _ = oracle.apply_op(m, +1, iorb, spin, nup, ndw, v)
            assert abs(c @ c + cd @ cd - 1.0) < 1e-13


def test_lanczos_tridiag_reproduces_spectrum(oracle):
    """Full-length tridiagonalisation of a small sector has the dense spectrum."""
    m = oracle.Model(**star_kwargs(3))
    nup, ndw = 2, 2
    H = oracle.dense_H(m, nup, ndw)
    n = H.shape[0]
    v = oracle.start_vector(n)
    a, b, nused = oracle.lanc_tridiag(lambda x: oracle.direct_hxv(m, nup, ndw, x), v, n)
    ev, _ = oracle.tridiag_eigh(a[:nused], b[1:nused])
    dense = np.linalg.eigvalsh(H)
    assert abs(ev[0] - dense[0]) < 1e-10 and abs(ev[-1] - dense[-1]) < 1e-10


def test_c_lanczos_matches_python_recurrence(oracle):
    """ora_stored_lanczos_gs (C pass 1 of sp_lanc_eigh on the stored operator, used at the
    BASELINE configs' own sizes) == the numpy recurrence lanc_eigh on the direct operator."""
    from models import star_kwargs, two_orb_kwargs

    O = oracle
    for kw, sec in ((star_kwargs(7), (4, 4)), (two_orb_kwargs(3), (4, 4)), (star_kwargs(9), (5, 4))):
        mo = O.Model(**kw)
        du, dd = O.sector_dims(mo.Ns, *sec)
        v0 = O.start_vector(du * dd, 31) - 0.5
        e_py, _, n_py = O.lanc_eigh(lambda x: O.direct_hxv(mo, sec[0], sec[1], x), du * dd, 300, 1e-12,
                                    v0=v0)
        e_c, n_c, a, b, _ = O.stored_lanczos_gs(mo, sec[0], sec[1], v0, 300, 1e-12, P=3, nthreads=3)
        assert n_c == n_py
        assert abs(e_c - e_py) < 1e-11
        a0, b0, _ = O.lanc_tridiag(lambda x: O.direct_hxv(mo, sec[0], sec[1], x),
                                   v0 / np.linalg.norm(v0), 12)
        assert np.abs(a[:12] - a0[:12]).max() < 1e-10 and np.abs(b[:12] - b0[:12]).max() < 1e-10
