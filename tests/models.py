"""Model definitions shared by the oracle-side and product-side tests: the same plain numbers
are fed to edipack_oracle.Model and to edipack_b200.EDModel."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden(name):
    with open(os.path.join(HERE, "golden", name + ".json")) as f:
        return json.load(f)


def _f(s):
    return float(s.replace("d", "e").replace("D", "e"))


def normal_normal_kwargs():
    """test/src/NORMAL_NORMAL: inputED.in + Hloc = Delta*sigma_z(orbital)
    (ed_normal_normal.f90:60-66), default bath from init_dmft_bath."""
    g = golden("normal_normal")["inputs"]
    norb = int(g["NORB"])
    delta = _f(g["DELTA"])
    hloc = np.zeros((2, norb, norb))
    for s in range(2):
        hloc[s] = np.diag([delta, -delta])
    return dict(Norb=norb, Nbath=int(g["NBATH"]), Nspin=int(g["NSPIN"]), bath_type=g["BATH_TYPE"],
                Uloc=tuple(_f(x) for x in g["ULOC"].split(",")), Ust=_f(g["UST"]), Jh=_f(g["JH"]),
                Jx=_f(g["JX"]), Jp=_f(g["JP"]), xmu=_f(g["XMU"]), hfmode=g["HFMODE"] == "T",
                beta=_f(g["BETA"]), ed_hw_bath=_f(g["ED_HW_BATH"]), hloc=hloc)


def star_kwargs(nbath, U=2.0):
    """Synthetic single-band Anderson impurity of SURVEY 8d (cfg1/2/4 family): Norb=1, normal
    bath from init_dmft_bath, hfmode, xmu=0, Hloc=0."""
    return dict(Norb=1, Nbath=nbath, Uloc=(U,), hfmode=True, xmu=0.0)


def two_orb_kwargs(nbath, with_nd=True):
    """cfg3 family: Norb=2, normal bath, U=U'=2, Jh=Jx=Jp=0.125, Hloc=0.5 sigma_z."""
    hloc = np.zeros((2, 2, 2))
    for s in range(2):
        hloc[s] = np.diag([0.5, -0.5])
    j = 0.125
    return dict(Norb=2, Nbath=nbath, Uloc=(2.0, 2.0), Ust=2.0, Jh=j, Jx=j if with_nd else 0.0,
                Jp=j if with_nd else 0.0, hfmode=True, hloc=hloc)


def messy_kwargs(bath_type="normal"):
    """Everything switched on with irrational-ish numbers: spin-dependent bath, spin field,
    inter-orbital Hloc, xmu != 0, no hfmode -- exercises every term of HxV_local/up/dw."""
    rng = np.random.default_rng(7)
    norb, nbath = 2, 2
    hloc = np.zeros((2, norb, norb))
    for s in range(2):
        a = rng.standard_normal((norb, norb))
        hloc[s] = 0.3 * (a + a.T)
    nfoo = 1 if bath_type == "hybrid" else norb
    kw = dict(Norb=norb, Nbath=nbath, Nspin=2, bath_type=bath_type, Uloc=(1.7, 2.3), Ust=1.1,
              Jh=0.21, Jx=0.13, Jp=0.17, xmu=0.37, hfmode=False, hloc=hloc,
              bath_e=rng.standard_normal((2, nfoo, nbath)),
              bath_v=0.5 + rng.random((2, norb, nbath)), spin_field_z=(0.11, -0.07))
    return kw


def replica_kwargs():
    rng = np.random.default_rng(11)
    norb, nbath = 2, 2
    hb = np.zeros((2, norb, norb, nbath))
    for k in range(nbath):
        a = rng.standard_normal((norb, norb))
        a = 0.4 * (a + a.T)
        hb[0, :, :, k] = a
        hb[1, :, :, k] = a
    kw = two_orb_kwargs(nbath)
    kw.update(bath_type="replica", hbath=hb, bath_e=np.zeros((2, norb, nbath)),
              bath_v=np.full((2, norb, nbath), 0.6))
    return kw


def hybrid_nonsu2_model(oracle_nonsu2, name="hybrid_nonsu2"):
    """test/src/HYBRID_NONSU2 (or NORMAL_NONSU2: same driver with a normal bath): inputED.in +
    Hloc = Mh * Gamma5 (= sigma_0 (x) tau_z in the test's spin-major so2j ordering,
    ed_hybrid_nonsu2.f90:50,76; COMMON.f90:81-122), default bath."""
    g = golden(name)["inputs"]
    norb = int(g["NORB"])
    mh = _f(g["MH"])
    hloc = np.zeros((2, 2, norb, norb), complex)
    for s in range(2):
        hloc[s, s] = np.diag([mh, -mh])
    m = oracle_nonsu2.ModelNonsu2(
        Norb=norb, Nbath=int(g["NBATH"]), bath_type=g["BATH_TYPE"],
        Uloc=tuple(_f(x) for x in g["ULOC"].split(",")), Ust=_f(g["UST"]), Jh=_f(g["JH"]),
        Jx=_f(g["JX"]), Jp=_f(g["JP"]), xmu=_f(g["XMU"]), hfmode=g["HFMODE"] == "T",
        ed_hw_bath=_f(g["ED_HW_BATH"]), hloc=hloc)
    return m.default_bath()


def soc_nonsu2_model(oracle_nonsu2, nbath=5):
    """cfg5 family at a size the numpy oracle builds in seconds: nonsu2, hybrid bath, complex
    spin-orbit-like Hloc (spin-flip, imaginary inter-orbital terms), spin field."""
    norb = 2
    hloc = np.zeros((2, 2, norb, norb), complex)
    lam = 0.3
    for s in range(2):
        hloc[s, s] = np.array([[0.2, 1j * lam * (1 - 2 * s)], [-1j * lam * (1 - 2 * s), -0.2]])
    hloc[0, 1] = np.array([[0.0, lam], [-lam, 0.0]])
    hloc[1, 0] = hloc[0, 1].conj().T
    rng = np.random.default_rng(5)
    m = oracle_nonsu2.ModelNonsu2(Norb=norb, Nbath=nbath, bath_type="hybrid", Uloc=(2.0, 1.7),
                                  Ust=1.5, Jh=0.25, Jx=0.25, Jp=0.25, xmu=0.1, hfmode=True, hloc=hloc,
                                  spin_field=np.array([[0.05, 0.02, -0.03], [0.0, 0.04, 0.01]]))
    m.default_bath()
    m.bath_u = 0.2 + 0.2 * rng.random(m.bath_u.shape)
    return m


def ineq_site_kwargs(ilat: int):
    """test/src/INEQ_NORMAL_NORMAL, site ilat (0 or 1): Norb=1, Nbath=7, Nspin=2, Hloc =
    (sigma_z (x) tau_0) reshaped over (lat,spin,orb) -> +1 on site 1 and -1 on site 2 for both
    spins (ed_normal_normal_afm2.f90:71-72); default bath, then ed_break_symmetry_bath with
    sign (-1)**(ilat+1): e_up += sign*SB_FIELD, e_dw -= sign*SB_FIELD (ED_BATH_USER.f90:166-167)."""
    g = golden("ineq_normal_normal")["inputs"]
    nbath = int(g["NBATH"])
    sign = 1.0 if ilat == 0 else -1.0
    sb = _f(g["SB_FIELD"])
    hw = _f(g["ED_HW_BATH"])
    e = np.zeros(nbath)  # init_dmft_bath, odd Nbath (ED_BATH_DMFT.f90:211-244)
    e[0], e[-1] = -hw, hw
    nh = nbath // 2
    de = hw / nh
    e[nh] = 0.0
    for i in range(2, nh + 1):
        e[i - 1] = -hw + (i - 1) * de
        e[nbath - i] = hw - (i - 1) * de
    bath_e = np.stack([e + sign * sb, e - sign * sb])[:, None, :]
    bath_v = np.full((2, 1, nbath), max(0.1, 1.0 / np.sqrt(nbath)))
    return dict(Norb=1, Nbath=nbath, Nspin=2, bath_type=g["BATH_TYPE"], Uloc=(_f(g["ULOC"]),),
                xmu=_f(g["XMU"]), hfmode=g["HFMODE"] == "T", beta=_f(g["BETA"]),
                ed_hw_bath=hw, hloc=np.full((2, 1, 1), sign), bath_e=bath_e, bath_v=bath_v)


def hybrid_normal_kwargs():
    """test/src/HYBRID_NORMAL: Norb=2, hybrid bath of 4 shared levels (default init_dmft_bath),
    Hloc = Delta*sigma_z(orbital) (ed_hybrid_normal.f90:57-63)."""
    g = golden("hybrid_normal")["inputs"]
    norb = int(g["NORB"])
    delta = _f(g["DELTA"])
    hloc = np.zeros((2, norb, norb))
    for s in range(2):
        hloc[s] = np.diag([delta, -delta])
    return dict(Norb=norb, Nbath=int(g["NBATH"]), Nspin=int(g["NSPIN"]), bath_type="hybrid",
                Uloc=tuple(_f(x) for x in g["ULOC"].split(",")), Ust=_f(g["UST"]), Jh=_f(g["JH"]),
                Jx=_f(g["JX"]), Jp=_f(g["JP"]), xmu=_f(g["XMU"]), hfmode=g["HFMODE"] == "T",
                beta=_f(g["BETA"]), ed_hw_bath=_f(g["ED_HW_BATH"]), hloc=hloc)


def replica_normal_kwargs(kind="replica"):
    """test/src/REPLICA_NORMAL and GENERAL_NORMAL: Norb=2, Nbath=2 replicas with
    H_k = lambda_k1 * 1 + 0.1 * tau_x, lambda_k1 = -1 + 2(k-1)/(Nbath-1)
    (ed_replica_normal.f90:71-86; the offsets of init_dmft_bath do not apply: the lambdas of the
    diagonal symmetry differ, ED_BATH_DMFT.f90:268-276), V = max(0.1, 1/sqrt(Nbath)) (:251-257)."""
    g = golden(f"{kind}_normal")["inputs"]
    norb, nb = int(g["NORB"]), int(g["NBATH"])
    delta = _f(g["DELTA"])
    hloc = np.zeros((2, norb, norb))
    for s in range(2):
        hloc[s] = np.diag([delta, -delta])
    hb = np.zeros((2, norb, norb, nb))
    for k in range(nb):
        lam = -1.0 + 2.0 * k / (nb - 1)
        for s in range(2):
            hb[s, :, :, k] = lam * np.eye(norb) + 0.1 * (np.ones((norb, norb)) - np.eye(norb))
    return dict(Norb=norb, Nbath=nb, Nspin=int(g["NSPIN"]), bath_type=kind,
                Uloc=tuple(_f(x) for x in g["ULOC"].split(",")), Ust=_f(g["UST"]), Jh=_f(g["JH"]),
                Jx=_f(g["JX"]), Jp=_f(g["JP"]), xmu=_f(g["XMU"]), hfmode=g["HFMODE"] == "T",
                beta=_f(g["BETA"]), hloc=hloc, hbath=hb, bath_e=np.zeros((2, norb, nb)),
                bath_v=np.full((2, norb, nb), max(0.1, 1.0 / np.sqrt(nb))))


def superc_model(oracle_superc, name):
    """test/src/NORMAL_SUPERC / HYBRID_SUPERC: attractive two-orbital model, default bath with
    d = DELTASC, Hloc = Delta*sigma_z(orbital) (ed_normal_superc.f90:57-63)."""
    g = golden(name)["inputs"]
    norb = int(g["NORB"])
    delta = _f(g["DELTA"])
    hloc = np.zeros((2, norb, norb), complex)
    for s in range(2):
        hloc[s] = np.diag([delta, -delta])
    m = oracle_superc.ModelSuperc(
        Norb=norb, Nbath=int(g["NBATH"]), bath_type=g["BATH_TYPE"],
        Uloc=tuple(_f(x) for x in g["ULOC"].split(",")), Ust=_f(g["UST"]), Jh=_f(g["JH"]),
        Jx=_f(g["JX"]), Jp=_f(g["JP"]), xmu=_f(g["XMU"]), hfmode=g["HFMODE"] == "T",
        ed_hw_bath=_f(g["ED_HW_BATH"]), deltasc=_f(g["DELTASC"]), hloc=hloc)
    return m.default_bath()


def replica_nonsu2_model(oracle_nonsu2, kind="replica"):
    """test/src/REPLICA_NONSU2 / GENERAL_NONSU2 (ed_replica_nonsu2.f90:47-88): Hloc = Mh*Gamma5,
    bath replicas H_k = lambda_1(k) Gamma5 + sb (GammaE0 + GammaEz) - sb GammaEx in the spin-major
    so2j ordering (COMMON.f90:81-122).  init_dmft_bath (ED_BATH_DMFT.f90:246-276): Gamma5 is diagonal
    and its lambdas are all equal, so they are spread by linspace(-ED_OFFSET_BATH, ED_OFFSET_BATH, Nbath);
    V = max(0.1, 1/sqrt(Nbath)) for every replica / spin-orbital."""
    g = golden(f"{kind}_nonsu2")["inputs"]
    norb, nb = int(g["NORB"]), int(g["NBATH"])
    mh, sb, off = _f(g["MH"]), _f(g["SB_FIELD"]), _f(g["ED_OFFSET_BATH"])
    s0, sx, sz = np.eye(2), np.array([[0, 1], [1, 0.0]]), np.diag([1.0, -1.0])
    tx, tz = sx, sz
    G5, GE0, GEz, GEx = np.kron(s0, tz), np.kron(s0, tx), np.kron(sz, tx), np.kron(sx, tx)

    def j2so(M):   # (ispin-1)*Nspin + iorb, spin-major
        return np.asarray(M, complex).reshape(2, norb, 2, norb).transpose(0, 2, 1, 3)

    hloc = j2so(mh * G5)
    hb = np.zeros((2, 2, norb, norb, nb), complex)
    offs = np.linspace(-off, off, nb) if nb > 1 else np.zeros(1)
    for k in range(nb):
        hb[..., k] = j2so((mh + offs[k]) * G5 + sb * GE0 + sb * GEz - sb * GEx)
    m = oracle_nonsu2.ModelNonsu2(
        Norb=norb, Nbath=nb, bath_type=kind,
        Uloc=tuple(_f(x) for x in g["ULOC"].split(",")), Ust=_f(g["UST"]), Jh=_f(g["JH"]),
        Jx=_f(g["JX"]), Jp=_f(g["JP"]), xmu=_f(g["XMU"]), hfmode=g["HFMODE"] == "T",
        ed_hw_bath=_f(g["ED_HW_BATH"]), hloc=hloc, hbath=hb)
    m.bath_e = np.zeros((2, norb, nb))
    m.bath_v = np.full((2, norb, nb), max(0.1, 1.0 / np.sqrt(nb)))
    m.bath_u = np.zeros((2, norb, nb))
    return m


def replica_superc_model(oracle_superc, kind="replica"):
    """test/src/REPLICA_SUPERC / GENERAL_SUPERC (ed_replica_superc.f90:63-96): Nspin=1, Nambu replicas
    H_k = lambda_k sigma_z(nambu) (x) 1 + 0.1 sigma_x (x) 1 + 0.2 sigma_x (x) tau_x in the nambu-major
    mso2j ordering (COMMON.f90:124-131), lambda_k = -1 + 2(k-1)/(Nbath-1) (all different: no offsets),
    V = max(0.1, 1/sqrt(Nbath)); Hloc = Delta*sigma_z(orbital)."""
    g = golden(f"{kind}_superc")["inputs"]
    norb, nb = int(g["NORB"]), int(g["NBATH"])
    delta = _f(g["DELTA"])
    hloc = np.zeros((2, norb, norb), complex)
    for s in range(2):
        hloc[s] = np.diag([delta, -delta])
    s0, sx, sz = np.eye(2), np.array([[0, 1], [1, 0.0]]), np.diag([1.0, -1.0])
    GN, GAA, GAB = np.kron(sz, s0), np.kron(sx, s0), np.kron(sx, sx)
    hb = np.zeros((2, 2, norb, norb, nb), complex)
    for k in range(nb):
        lam = -1.0 + 2.0 * k / (nb - 1)
        M = lam * GN + 0.1 * GAA + 0.2 * GAB
        hb[..., k] = M.reshape(2, norb, 2, norb).transpose(0, 2, 1, 3)
    m = oracle_superc.ModelSuperc(
        Norb=norb, Nbath=nb, bath_type=kind,
        Uloc=tuple(_f(x) for x in g["ULOC"].split(",")), Ust=_f(g["UST"]), Jh=_f(g["JH"]),
        Jx=_f(g["JX"]), Jp=_f(g["JP"]), xmu=_f(g["XMU"]), hfmode=g["HFMODE"] == "T",
        ed_hw_bath=_f(g["ED_HW_BATH"]), deltasc=_f(g["DELTASC"]), hloc=hloc, hbath=hb)
    m.bath_e = np.zeros((2, norb, nb))
    m.bath_d = np.zeros((norb, nb))
    m.bath_v = np.full((2, norb, nb), max(0.1, 1.0 / np.sqrt(nb)))
    return m
