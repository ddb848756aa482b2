"""a10: user two-body terms (coulomb_sundry) and phonons of the direct H x v.  The
reference holds no fixture with Nph>0 or a umatrix file, so the oracle's restatement of
direct/HxV_sundry.f90, HxV_ph.f90 and HxV_eph.f90 is checked against an INDEPENDENT construction:
Jordan-Wigner fermion matrices on the full 2*Ns-mode Fock space times a truncated boson space."""
import numpy as np
import pytest

from models import two_orb_kwargs


def jw_ops(nmodes):
    """Annihilators c_p on the 2^nmodes Fock space, c_p|n> = (-1)^{sum_{q<p} n_q} |n - e_p>."""
    dim = 1 << nmodes
    ops = []
    for p in range(nmodes):
        c = np.zeros((dim, dim))
        for s in range(dim):
            if (s >> p) & 1:
                sign = -1.0 if bin(s & ((1 << p) - 1)).count("1") % 2 else 1.0
                c[s ^ (1 << p), s] = sign
        ops.append(c)
    return ops


def sector_states(oracle, Ns, nup, ndw):
    """Full-Fock-space indices (mup + mdw*2^Ns) of the sector in the reference's order."""
    mu, md = oracle.build_map(Ns, nup), oracle.build_map(Ns, ndw)
    return np.array([int(a) + (int(b) << Ns) for b in md for a in mu])


def extra_dense(oracle, m, nup, ndw, sundry, ph):
    """(H_sundry (x) 1 + 1 (x) [w0 b^+b + A(b+b^+)] + sum_ab g_ab c^+_a c_b (x) (b+b^+)) in the sector."""
    Ns = m.Ns
    c = jw_ops(2 * Ns)
    mode = lambda orb, spin: (orb - 1) + (spin - 1) * Ns
    st = sector_states(oracle, Ns, nup, ndw)
    nel = len(st)
    Hs = np.zeros((1 << (2 * Ns),) * 2)
    for ci, cj, ck, cl, U in sundry:
        O = c[mode(*ci)].T @ c[mode(*ck)] @ c[mode(*cj)].T @ c[mode(*cl)]  # applied right to left
        Hs += U * O.T  # gather form Hv(j) += U <i|O|j> v(i), direct/HxV_sundry.f90:100-104
    Hs = Hs[np.ix_(st, st)]
    nph = (ph["Nph"] if ph else 0) + 1
    H = np.kron(np.eye(nph), Hs)
    if ph:
        b = np.diag(np.sqrt(np.arange(1, nph)), 1)
        x = b + b.T
        g = np.atleast_2d(np.asarray(ph["g"], float))
        G = np.zeros_like(Hs)
        full = np.zeros((1 << (2 * Ns),) * 2)
        for a in range(m.Norb):
            for bb in range(m.Norb):
                for spin in (1, 2):
                    full += g[a, bb] * c[mode(a + 1, spin)].T @ c[mode(bb + 1, spin)]
        G = full[np.ix_(st, st)]
        # the reference scatters Hv(c^+_a c_b |i>) += g(a,b) v(i): H = sum g_ab c^+_a c_b as a matrix
        H += np.kron(ph["w0"] * np.diag(np.arange(nph)) + ph.get("A", 0.0) * x, np.eye(nel))
        H += np.kron(x, G)
    return H


SUNDRY = [
    ((1, 1), (2, 2), (2, 2), (1, 1), 0.7),    # density-density n_1up n_2dw
    ((1, 1), (2, 2), (1, 2), (2, 1), -0.3),   # spin-exchange-like
    ((1, 1), (1, 2), (2, 2), (2, 1), 0.45),   # pair-hopping-like
    ((2, 1), (1, 1), (1, 1), (2, 1), 0.2),    # same-spin exchange
    ((2, 2), (2, 2), (1, 2), (2, 2), 0.11),   # correlated hopping (dw)
]


@pytest.mark.parametrize("nup,ndw", [(2, 2), (1, 3), (3, 2), (0, 4), (4, 4)])
@pytest.mark.parametrize("case", ["sundry", "phonons", "both", "phonons_A"])
def test_ext_terms_against_jordan_wigner(oracle, case, nup, ndw):
    m = oracle.Model(**two_orb_kwargs(1))  # Norb=2, Nbath=1: Ns=4, full Fock space 256
    assert m.Ns == 4
    sundry = SUNDRY if case in ("sundry", "both") else []
    ph = None
    if case != "sundry":
        ph = dict(Nph=3, w0=0.37, g=[[0.5, 0.2], [0.2, -0.3]], A=0.25 if case == "phonons_A" else 0.0)
    du, dd = oracle.sector_dims(m.Ns, nup, ndw)
    nph = (ph["Nph"] if ph else 0) + 1
    rng = np.random.default_rng(11)
    v = rng.standard_normal(du * dd * nph)
    base = np.concatenate([oracle.direct_hxv(m, nup, ndw, v[k * du * dd:(k + 1) * du * dd]) for k in range(nph)])
    got = oracle.direct_hxv_ext(m, nup, ndw, v, sundry, ph) - base
    ref = extra_dense(oracle, m, nup, ndw, sundry, ph) @ v
    assert np.abs(got - ref).max() < 1e-13 * max(1.0, np.abs(ref).max())


def test_nonsymmetric_g_follows_the_reference_scatter(oracle):
    """g_ph(a,b) != g_ph(b,a): the reference's scatter Hv(c^+_a c_b|i>) += g(a,b) v(i) is H = sum g_ab c^+_a c_b."""
    m = oracle.Model(**two_orb_kwargs(1))
    ph = dict(Nph=2, w0=0.1, g=[[0.0, 0.4], [-0.15, 0.0]])
    du, dd = oracle.sector_dims(m.Ns, 2, 2)
    v = np.random.default_rng(2).standard_normal(du * dd * 3)
    base = np.concatenate([oracle.direct_hxv(m, 2, 2, v[k * du * dd:(k + 1) * du * dd]) for k in range(3)])
    got = oracle.direct_hxv_ext(m, 2, 2, v, [], ph) - base
    assert np.abs(got - extra_dense(oracle, m, 2, 2, [], ph) @ v).max() < 1e-13


def test_spin_unbalanced_sundry_term_is_refused(oracle):
    m = oracle.Model(**two_orb_kwargs(1))
    with pytest.raises(ValueError):
        oracle.direct_hxv_ext(m, 2, 2, np.zeros(36), [((1, 1), (2, 1), (1, 2), (2, 2), 1.0)], None)


def test_kanamori_terms_as_sundry_lines(oracle):
    """Jx / Jp written as coulomb_sundry lines give the built-in non-local terms
    (direct/HxV_non_local.f90:16-72): S-E c^+_{a up} c_{b up} c^+_{b dw} c_{a dw}, P-H c^+_{a up} c^+_{a dw} c_{b dw} c_{b up}."""
    kw = two_orb_kwargs(2)
    J = 0.125
    m_with = oracle.Model(**{**kw, "Jx": J, "Jp": J})
    m_without = oracle.Model(**{**kw, "Jx": 0.0, "Jp": 0.0})
    lines = []
    for a in (1, 2):
        for b in (1, 2):
            if a == b:
                continue
            # chain applied right to left: c_l, cd_j, c_k, cd_i
            lines.append(((a, 1), (b, 2), (b, 1), (a, 2), J))   # S-E: c_{a dw}, c^+_{b dw}, c_{b up}, c^+_{a up}
            lines.append(((a, 1), (a, 2), (b, 1), (b, 2), J))   # P-H: c_{b dw}, c^+_{a dw}, c_{b up}, c^+_{a up}
    ns = m_with.Ns
    rng = np.random.default_rng(4)
    for nup, ndw in [(ns // 2, ns // 2), (2, 3)]:
        du, dd = oracle.sector_dims(ns, nup, ndw)
        v = rng.standard_normal(du * dd)
        ref = oracle.direct_hxv(m_with, nup, ndw, v)
        got = oracle.direct_hxv_ext(m_without, nup, ndw, v, lines, None)
        assert np.abs(got - ref).max() < 1e-13
