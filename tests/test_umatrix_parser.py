"""ED_PARSE_UMATRIX (set_umatrix / parse_umatrix_line) restated on both sides of the boundary.
Pins: (i) every umatrix.restart and set_twobody_hk() operator list of the reference's test drivers
gives the Kanamori matrices of the same driver's inputED.in -- the reference asserts the SAME
goldens for ED_READ_UMATRIX=T, run-time operators and ED_USE_KANAMORI=T; (ii) the whole chain
operator list -> (Uloc, Ust, Jh, Jx, Jp, mfHloc, coulomb_sundry) -> H x v equals the physical
definition H_int = 1/2 sum U_ijkl c^+_i c^+_j c_l c_k built from Jordan-Wigner matrices."""
import numpy as np
import pytest

from models import _f, golden
from test_oracle_sundry_phonons import jw_ops, sector_states

FIXTURES = ["normal_normal", "hybrid_normal", "replica_normal", "general_normal", "hybrid_nonsu2",
            "normal_nonsu2", "replica_nonsu2", "general_nonsu2", "normal_superc", "hybrid_superc",
            "replica_superc", "general_superc"]


def kanamori_of(g):
    i = g["inputs"]
    no = int(i["NORB"])
    off = 1.0 - np.eye(no)
    U = np.array([_f(x) for x in i["ULOC"].split(",")][:no])
    return dict(Uloc=U, Ust=_f(i["UST"]) * off, Jh=_f(i["JH"]) * off, Jx=_f(i["JX"]) * off,
                Jp=_f(i["JP"]) * off)


def parsers(oracle):
    import edipack_b200 as E
    return [("oracle", oracle.parse_umatrix), ("host", E.parse_umatrix)]


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("key", ["umatrix", "twobody_hk"])
def test_reference_operator_lists_give_kanamori(oracle, name, key):
    g = golden(name)
    lines = [tuple(l) for l in g[key]]
    assert len(lines) == 24
    exp = kanamori_of(g)
    no = int(g["inputs"]["NORB"])
    for who, parse in parsers(oracle):
        um = parse(no, lines, use_kanamori=False)
        for k in ("Uloc", "Ust", "Jh", "Jx", "Jp"):
            assert np.abs(um[k] - exp[k]).max() < 1e-12, (who, k)
        assert um["sundry"] == [] and np.abs(um["mfHloc"]).max() == 0.0


def test_ineq_umatrix_files(oracle):
    g = golden("ineq_normal_normal")
    um = oracle.parse_umatrix(1, [tuple(l) for l in g["umatrix"]])
    assert abs(um["Uloc"][0] - _f(g["inputs"]["ULOC"].split(",")[0])) < 1e-12


def random_twobody(rng, norb, n):
    """n random spin-conserving operators U c^+_i c^+_j c_l c_k, each with its equivalent form
    (i<->j, k<->l), its hermitian conjugate and the global spin flip of all of these.  The last
    is required by the reference's density-density bookkeeping: Ust(a,b) multiplies
    nup_a ndw_b + nup_b ndw_a and Ust-Jh multiplies nup_a nup_b + ndw_a ndw_b
    (direct/HxV_local.f90:36-52), so an interaction that is not spin-flip symmetric cannot be
    represented (ED_PARSE_UMATRIX.f90:116-133 symmetrises it away)."""
    lines = []
    sp = "ud"
    flip = {"u": "d", "d": "u"}
    for _ in range(n):
        while True:
            oi, oj, ok, ol = (int(x) for x in rng.integers(1, norb + 1, 4))
            si, sj = sp[rng.integers(2)], sp[rng.integers(2)]
            sk, sl = (si, sj) if rng.integers(2) else (sj, si)   # c^+_i c^+_j c_l c_k keeps Nup, Ndw
            if (oi, si) != (oj, sj) and (ok, sk) != (ol, sl):
                break
        U = float(np.round(rng.standard_normal(), 3))
        base = [(oi, si, oj, sj, ok, sk, ol, sl), (oj, sj, oi, si, ol, sl, ok, sk)]
        base += [(k_, sk_, l_, sl_, i_, si_, j_, sj_) for (i_, si_, j_, sj_, k_, sk_, l_, sl_) in base]
        base += [(a, flip[b], c_, flip[d], e, flip[f], g_, flip[h]) for (a, b, c_, d, e, f, g_, h) in base]
        for b in dict.fromkeys(base):   # distinct forms only
            lines.append(b + (U,))
    return lines


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_operator_list_to_hxv_equals_physical_definition(oracle, seed):
    import edipack_b200 as E
    rng = np.random.default_rng(seed)
    norb, nbath = 2, 1
    lines = random_twobody(rng, norb, 8)
    # interaction only: no hybridisation, no bath energies, no Hartree shift
    zeros = dict(bath_e=np.zeros((2, norb, nbath)), bath_v=np.zeros((2, norb, nbath)))
    kw = dict(Norb=norb, Nbath=nbath, Uloc=(0.0, 0.0), hfmode=False, ed_use_kanamori=False,
              umatrix_lines=tuple(lines), **zeros)
    m = oracle.Model(**kw)
    assert bytes(E.EDModel(**kw).params()) == bytes(m.params())
    assert E.EDModel(**kw).coulomb_sundry == m.coulomb_sundry
    ns = m.Ns
    c = jw_ops(2 * ns)
    mode = lambda orb, s: (orb - 1) + (0 if s == "u" else ns)
    H = np.zeros((1 << (2 * ns),) * 2)
    for (oi, si, oj, sj, ok, sk, ol, sl, U) in lines:
        H += 0.5 * U * c[mode(oi, si)].T @ c[mode(oj, sj)].T @ c[mode(ol, sl)] @ c[mode(ok, sk)]
    assert np.abs(H - H.T).max() < 1e-14
    assert len(m.coulomb_sundry) > 0
    for nup, ndw in [(2, 2), (1, 2), (3, 1), (2, 3), (4, 2)]:
        st = sector_states(oracle, ns, nup, ndw)
        Hs = H[np.ix_(st, st)]
        v = rng.standard_normal(len(st))
        got = oracle.direct_hxv_ext(m, nup, ndw, v, m.coulomb_sundry, None)
        assert np.abs(got - Hs @ v).max() < 1e-13 * max(1.0, np.abs(Hs @ v).max())


def test_read_umatrix_file_roundtrip(oracle, tmp_path):
    import edipack_b200 as E
    g = golden("normal_normal")
    f = tmp_path / "umatrix.restart"
    with open(f, "w") as fh:
        fh.write("#Interaction two-body operators\n2 BANDS\n")
        for l in g["umatrix"]:
            fh.write("%d %s %d %s %d %s %d %s %21.12E\n" % tuple(l))
        fh.write("this line is not an operator\n")
    for rd in (oracle.read_umatrix_file, E.read_umatrix_file):
        norb, lines = rd(str(f))
        assert norb == 2 and [list(x) for x in lines] == g["umatrix"]
