"""CPU oracle for the EDIpack NORMAL-mode Lanczos H x v hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module, and
only as the checker / the timed CPU baseline -- never from the product package.

The heavy loops live in ``ed_oracle.c`` (C restatement of the reference's Fortran include
fragments, each function citing reference file:line); this module adds the host-side
contracts needed to reproduce the reference's golden ``*.check`` files end to end:

* default bath / interaction set-up  (ED_BATH_DMFT.f90:211-244, ED_PARSE_UMATRIX.f90:88-165)
* sector bookkeeping                 (ED_SECTOR.f90:1559-1718, ED_SETUP.f90:525-665)
* plain Lanczos drivers              (SciFortran SF_SP_LINALG sp_lanc_eigh / sp_lanc_tridiag
                                      -- NOT in the reference tree, unpinned dependency;
                                      restated from their published algorithm, parity with
                                      the reference pinned only through end-to-end goldens)
* solver / GF / observables contracts (ED_DIAG_NORMAL.f90:76-296, ED_GF_NORMAL.f90:131-177,
                                      363-427, 568-605, 698-739, ED_OBSERVABLES_NORMAL.f90:150-215)

All paths relative to /root/reference/.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAXORB, MAXBATH = 5, 32
BATH_CODES = {"normal": 0, "hybrid": 1, "replica": 2, "general": 3}


class OraParams(C.Structure):
    """Mirror of ``ora_params`` in ed_oracle.h (same field order and sizes)."""

    _fields_ = [
        ("Ns", C.c_int32), ("Norb", C.c_int32), ("Nbath", C.c_int32), ("bath_type", C.c_int32),
        ("hfmode", C.c_int32), ("Nfoo", C.c_int32), ("pad0", C.c_int32), ("pad1", C.c_int32),
        ("xmu", C.c_double),
        ("eloc", C.c_double * (2 * MAXORB * MAXORB)),
        ("spin_field_z", C.c_double * MAXORB),
        ("exc_field", C.c_double * 4),
        ("Uloc", C.c_double * MAXORB),
        ("Ust", C.c_double * (MAXORB * MAXORB)),
        ("Jh", C.c_double * (MAXORB * MAXORB)),
        ("Jx", C.c_double * (MAXORB * MAXORB)),
        ("Jp", C.c_double * (MAXORB * MAXORB)),
        ("diag_hybr", C.c_double * (2 * MAXORB * MAXBATH)),
        ("bath_diag", C.c_double * (2 * MAXORB * MAXBATH)),
        ("hbath", C.c_double * (2 * MAXORB * MAXORB * MAXBATH)),
        ("stride", C.c_int32 * (MAXORB * MAXBATH)),
    ]


class OraSundryTerm(C.Structure):
    """Mirror of ``ora_sundry_term``: U cd_i cd_j c_k c_l, each operator (orbital 1-based, spin 1|2)."""

    _fields_ = [("cd_i", C.c_int32 * 2), ("cd_j", C.c_int32 * 2), ("c_k", C.c_int32 * 2),
                ("c_l", C.c_int32 * 2), ("U", C.c_double)]


class OraPhonons(C.Structure):
    """Mirror of ``ora_phonons``."""

    _fields_ = [("Nph", C.c_int32), ("pad", C.c_int32), ("w0", C.c_double), ("A", C.c_double),
                ("g", C.c_double * (MAXORB * MAXORB))]


def build_library(force: bool = False) -> str:
    """Compile ed_oracle.c -> libed_oracle.so (gcc, reference release flags)."""
    so = os.path.join(_HERE, "libed_oracle.so")
    src = os.path.join(_HERE, "ed_oracle.c")
    hdr = os.path.join(_HERE, "ed_oracle.h")
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libed_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_library())
        L = _lib
        dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        ip32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        ip64 = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        PP = C.POINTER(OraParams)
        L.ora_build_map.restype = C.c_int64
        L.ora_build_map.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.ora_binary_search.restype = C.c_int64
        L.ora_binary_search.argtypes = [ip32, C.c_int64, C.c_int32]
        L.ora_binomial.restype = C.c_int64
        L.ora_binomial.argtypes = [C.c_int, C.c_int]
        L.ora_c.argtypes = [C.c_int, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.ora_cdg.argtypes = [C.c_int, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.ora_direct_hxv.argtypes = [PP, C.c_int, C.c_int, dp, dp]
        L.ora_direct_hxv_ext.argtypes = [PP, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, dp, dp]
        L.ora_direct_hxv_mpi.argtypes = [PP, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp]
        L.ora_direct_hxv_mpi_sample.argtypes = [PP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp,
                                                dp, C.POINTER(C.c_double)]
        L.ora_stored_hxv.argtypes = [PP, C.c_int, C.c_int, dp, dp]
        L.ora_stored_hxv_mpi.argtypes = [PP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp,
                                         C.POINTER(C.c_double)]
        L.ora_stored_lanczos_gs.argtypes = [PP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                            C.c_int, dp, C.POINTER(C.c_double), C.POINTER(C.c_int),
                                            dp, dp, C.POINTER(C.c_double)]
        L.ora_build_hop_csr.restype = C.c_int64
        L.ora_build_hop_csr.argtypes = [PP, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_build_diag.argtypes = [PP, C.c_int, C.c_int, dp]
        L.ora_build_nonlocal_csr.restype = C.c_int64
        L.ora_build_nonlocal_csr.argtypes = [PP, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_apply_op.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp]
    return _lib



# --------------------------------------------------------------------------------------
# two-body operators: ED_PARSE_UMATRIX.f90 (read_umatrix_file :353-449, parse_umatrix_line
# :452-634, set_umatrix :88-165)
# --------------------------------------------------------------------------------------
def read_umatrix_file(path: str):
    """read_umatrix_file: '#'/'!'/'%' comment preamble, '<Norb> BANDS', then one operator per line
    'oi si oj sj ok sk ol sl U' (orbitals 1-based, spins u|d); malformed lines are skipped.
    Returns (Norb, [(oi, si, oj, sj, ok, sk, ol, sl, U), ...])."""
    norb, lines = None, []
    with open(path) as f:
        for raw in f:
            tok = raw.split()
            if not tok:
                continue
            if norb is None:
                if tok[0][0] in "#!%":
                    continue
                norb = int(tok[0])
                continue
            try:
                oi, si, oj, sj, ok, sk, ol, sl = (int(tok[0]), tok[1], int(tok[2]), tok[3], int(tok[4]),
                                                  tok[5], int(tok[6]), tok[7])
                U = float(tok[8].replace("d", "e").replace("D", "e"))
            except (ValueError, IndexError):
                continue
            lines.append((oi, si, oj, sj, ok, sk, ol, sl, U))
    return norb, lines


def parse_umatrix(Norb: int, lines, use_kanamori=False, Uloc=(), Ust=0.0, Jh=0.0, Jx=0.0, Jp=0.0):
    """set_umatrix: every line goes through parse_umatrix_line (1/2 prefactor and sign of the
    w2dynamics convention, canonical ordering of the creation / annihilation pairs, mean-field
    term of the anticommutator into mfHloc, re-swap to the c->cd->c->cd application order,
    classification into Uloc / Ust / Ust-Jh / Jx / Jp, everything else into coulomb_sundry); then
    the symmetrisations of :124-133 and the ED_USE_KANAMORI additions of :139-146.
    Returns dict(Uloc[Norb], Ust, Jh, Jx, Jp [Norb,Norb], mfHloc[2,2,Norb,Norb],
    sundry=[((orb,spin) cd_i, cd_j, c_k, c_l, U)] with spin 1 = up / 2 = dw)."""
    No = Norb
    U_in = np.zeros(No)
    Ust_in, Jh_in, Jx_in, Jp_in = (np.zeros((No, No)) for _ in range(4))
    mf = np.zeros((2, 2, No, No))
    sundry = []
    for (oi, si, oj, sj, ok, sk, ol, sl, U) in lines:
        if max(oi, oj, ok, ol) > No or min(oi, oj, ok, ol) < 1:
            raise ValueError("two-body operator: orbital index outside 1..Norb")
        if any(s not in ("u", "d") for s in (si, sj, sk, sl)):
            raise ValueError("two-body operator: spin index malformed")
        ci, cj = [oi, 1 if si == "u" else 2], [oj, 1 if sj == "u" else 2]
        ck, cl = [ok, 1 if sk == "u" else 2], [ol, 1 if sl == "u" else 2]
        if abs(U) < 1e-10:
            continue
        U = -0.5 * U
        if ci[0] > cj[0]:   # creation pair: increasing orbital ...
            ci, cj, U = cj, ci, -U
        if ci[1] > cj[1]:   # ... overridden by increasing spin
            ci, cj, U = cj, ci, -U
        if ck[0] > cl[0]:   # annihilation pair likewise
            ck, cl, U = cl, ck, -U
        if ck[1] > cl[1]:
            ck, cl, U = cl, ck, -U
        if cj == ck:        # anticommutator {c^+_j, c_k} = 1 leaves a one-body term
            mf[ci[1] - 1, ck[1] - 1, ci[0] - 1, ck[0] - 1] += U
        U = -U              # second and third operator are swapped at application time
        if ci[0] == ck[0] and cj[0] == cl[0]:
            if ci[1] != cj[1]:
                if ci[0] == cj[0]:
                    U_in[ci[0] - 1] += U
                else:
                    Ust_in[ci[0] - 1, cj[0] - 1] += U
                continue
            if ci[0] != cj[0]:
                Jh_in[ci[0] - 1, cj[0] - 1] += U   # holds Ust-Jh until the end
                continue
        if (ci[0] != cj[0] and ci[1] != cj[1] and ci[0] == cl[0] and ci[1] == ck[1]
                and cj[0] == ck[0] and cj[1] == cl[1]):
            Jx_in[ci[0] - 1, ck[0] - 1] += U
            continue
        if (ci[0] == cj[0] and ci[1] != cj[1] and ci[0] != ck[0] and ci[1] == ck[1]
                and cj[0] != cl[0] and cj[1] == cl[1]):
            Jp_in[ci[0] - 1, ck[0] - 1] += U
            continue
        sundry.append((tuple(ci), tuple(cj), tuple(ck), tuple(cl), U))
    Ust_in = (Ust_in + Ust_in.T) / 2.0
    Jh_in = (Jh_in + Jh_in.T) / 2.0
    Jh_in = Ust_in - Jh_in
    if use_kanamori:
        off = 1.0 - np.eye(No)
        U_in = U_in + np.asarray(Uloc, float)[:No]
        Ust_in = Ust_in + Ust * off
        Jh_in = Jh_in + Jh * off
        Jx_in = Jx_in + Jx * off
        Jp_in = Jp_in + Jp * off
    return dict(Uloc=U_in, Ust=Ust_in, Jh=Jh_in, Jx=Jx_in, Jp=Jp_in, mfHloc=mf, sundry=sundry)


# --------------------------------------------------------------------------------------
# model description (what ed_read_input + ed_init_solver + ed_set_Hloc + set_umatrix leave
# in the reference's module globals)
# --------------------------------------------------------------------------------------
@dataclass
class Model:
    Norb: int = 1
    Nbath: int = 1
    Nspin: int = 1
    bath_type: str = "normal"
    Uloc: tuple = (2.0,)
    Ust: float = 0.0
    Jh: float = 0.0
    Jx: float = 0.0
    Jp: float = 0.0
    xmu: float = 0.0
    hfmode: bool = True
    beta: float = 1000.0
    ed_hw_bath: float = 2.0
    hloc: np.ndarray | None = None      # [2(spin), Norb, Norb] real, impHloc(s,s,a,b)
    bath_e: np.ndarray | None = None    # [2, Nfoo, Nbath]
    bath_v: np.ndarray | None = None    # [2, Norb, Nbath]
    spin_field_z: tuple = ()
    exc_field: tuple = (0.0, 0.0, 0.0, 0.0)
    hbath: np.ndarray | None = None     # replica/general [2, Norb, Norb, Nbath]
    lanc_ngfiter: int = 200
    lanc_niter: int = 512
    lanc_tolerance: float = 1e-18
    lanc_dim_threshold: int = 1024
    gs_threshold: float = 1e-9
    ed_twin: bool = False               # ED_TWIN
    ed_use_kanamori: bool = True        # ED_USE_KANAMORI
    umatrix_lines: tuple = ()           # umatrix file lines + ed_add_twobody_operator calls
    _params: OraParams | None = field(default=None, repr=False)

    def umatrix(self):
        """set_umatrix (ED_PARSE_UMATRIX.f90:88-165) for this model."""
        return parse_umatrix(self.Norb, self.umatrix_lines, self.ed_use_kanamori, self.Uloc, self.Ust,
                             self.Jh, self.Jx, self.Jp)

    @property
    def coulomb_sundry(self):
        return self.umatrix()["sundry"]

    @property
    def Ns(self) -> int:
        # ED_SETUP.f90:118-126
        if self.bath_type == "hybrid":
            return self.Nbath + self.Norb
        return (self.Nbath + 1) * self.Norb

    @property
    def Nfoo(self) -> int:
        return 1 if self.bath_type == "hybrid" else self.Norb

    def bath_stride(self, a: int, k: int) -> int:
        """getBathStride(a+1,k+1) (1-based site), ED_SETUP.f90:605-622; a,k 0-based."""
        if self.bath_type == "normal":
            return self.Norb + a * self.Nbath + (k + 1)
        if self.bath_type == "hybrid":
            return self.Norb + (k + 1)
        return (a + 1) + (k + 1) * self.Norb

    def default_bath(self):
        """init_dmft_bath for normal/hybrid baths, ED_BATH_DMFT.f90:211-244."""
        Nb, hw = self.Nbath, self.ed_hw_bath
        e = np.zeros(Nb)
        e[0] = -hw
        e[Nb - 1] = hw
        Nh = Nb // 2
        if Nb % 2 == 0 and Nb >= 4:
            de = hw / max(Nh - 1, 1)
            e[Nh - 1] = -0.1
            e[Nh] = 0.1
            for i in range(2, Nh):  # i=2..Nh-1 (1-based)
                e[i - 1] = -hw + (i - 1) * de
                e[Nb - i] = hw - (i - 1) * de
        elif Nb % 2 != 0 and Nb >= 3:
            de = hw / Nh
            e[Nh] = 0.0
            for i in range(2, Nh + 1):
                e[i - 1] = -hw + (i - 1) * de
                e[Nb - i] = hw - (i - 1) * de
        self.bath_e = np.broadcast_to(e, (2, self.Nfoo, Nb)).copy()
        v = max(0.1, 1.0 / math.sqrt(Nb))
        self.bath_v = np.full((2, self.Norb, Nb), v)
        self._params = None
        return self

    def params(self) -> OraParams:
        if self._params is not None:
            return self._params
        p = OraParams()
        No, Nb = self.Norb, self.Nbath
        assert No <= MAXORB and Nb <= MAXBATH and self.Ns <= 31
        p.Ns, p.Norb, p.Nbath = self.Ns, No, Nb
        p.bath_type = BATH_CODES[self.bath_type]
        p.hfmode = int(self.hfmode)
        p.Nfoo = self.Nfoo
        p.xmu = self.xmu
        if self.bath_e is None:
            self.default_bath()
        hloc = np.zeros((2, No, No)) if self.hloc is None else np.asarray(self.hloc, float)
        um = self.umatrix()
        eloc = np.zeros((2, MAXORB, MAXORB))
        eloc[:, :No, :No] = hloc
        for s_ in range(2):  # impHloc + mfHloc, spin-diagonal blocks (direct/HxV_local.f90:17-20)
            eloc[s_, :No, :No] += um["mfHloc"][s_, s_]
        p.eloc[:] = eloc.ravel().tolist()
        sf = np.zeros(MAXORB)
        sf[: len(self.spin_field_z)] = self.spin_field_z
        p.spin_field_z[:] = sf.tolist()
        p.exc_field[:] = list(self.exc_field)
        # internal interaction matrices left by set_umatrix (ED_PARSE_UMATRIX.f90:88-165)
        U = np.zeros(MAXORB)
        U[:No] = um["Uloc"]
        p.Uloc[:] = U.tolist()
        for name in ("Ust", "Jh", "Jx", "Jp"):
            mat = np.zeros((MAXORB, MAXORB))
            mat[:No, :No] = um[name]
            getattr(p, name)[:] = mat.ravel().tolist()
        dh = np.zeros((2, MAXORB, MAXBATH))
        dh[:, :No, :Nb] = self.bath_v
        bd = np.zeros((2, MAXORB, MAXBATH))
        bd[:, : self.Nfoo, :Nb] = self.bath_e
        p.diag_hybr[:] = dh.ravel().tolist()
        p.bath_diag[:] = bd.ravel().tolist()
        hb = np.zeros((2, MAXORB, MAXORB, MAXBATH))
        if self.hbath is not None:
            hb[:, :No, :No, :Nb] = self.hbath
            # replica/general: bath_diag = diagonal of Hbath_tmp (DIRECT_HxV.f90:71-92)
            for s in range(2):
                for a in range(No):
                    bd[s, a, :Nb] = self.hbath[s, a, a, :]
            p.bath_diag[:] = bd.ravel().tolist()
        p.hbath[:] = hb.ravel().tolist()
        st = np.zeros((MAXORB, MAXBATH), np.int32)
        for a in range(No):
            for k in range(Nb):
                st[a, k] = self.bath_stride(a, k)
        p.stride[:] = st.ravel().tolist()
        self._params = p
        return p


# --------------------------------------------------------------------------------------
# sector bookkeeping
# --------------------------------------------------------------------------------------
def binomial(n: int, k: int) -> int:
    return int(lib().ora_binomial(n, k))


def sector_index(Ns: int, nup: int, ndw: int) -> int:
    """get_Sector_normal (ED_SECTOR.f90:1559-1570), QN=[nup,ndw], 1-based."""
    return 1 + nup * (Ns + 1) + ndw


def sector_qn(Ns: int, isector: int):
    """get_Nup / get_Ndw (ED_SECTOR.f90:1618-1640)."""
    c = isector - 1
    ndw = c % (Ns + 1)
    nup = c // (Ns + 1)
    return nup, ndw


def build_map(Ns: int, nel: int) -> np.ndarray:
    dim = binomial(Ns, nel)
    m = np.empty(dim, np.int32)
    got = lib().ora_build_map(Ns, nel, m.ctypes.data)
    assert got == dim
    return m


def sector_dims(Ns, nup, ndw):
    return binomial(Ns, nup), binomial(Ns, ndw)


def direct_hxv(model: Model, nup, ndw, v):
    v = np.ascontiguousarray(v, np.float64)
    hv = np.empty_like(v)
    rc = lib().ora_direct_hxv(C.byref(model.params()), nup, ndw, v, hv)
    assert rc == 0
    return hv


def sundry_array(terms):
    """terms: iterable of ((orb,spin) cd_i, (orb,spin) cd_j, (orb,spin) c_k, (orb,spin) c_l, U) with
    1-based orbitals and spin 1 = up / 2 = dw, like one line of coulomb_sundry."""
    arr = (OraSundryTerm * max(len(terms), 1))()
    for t, (ci, cj, ck, cl, U) in enumerate(terms):
        arr[t].cd_i[:] = ci
        arr[t].cd_j[:] = cj
        arr[t].c_k[:] = ck
        arr[t].c_l[:] = cl
        arr[t].U = U
    return arr


def phonon_struct(Nph, w0, g, A=0.0):
    ph = OraPhonons()
    ph.Nph, ph.w0, ph.A = int(Nph), float(w0), float(A)
    gg = np.zeros((MAXORB, MAXORB))
    g = np.atleast_2d(np.asarray(g, float))
    gg[: g.shape[0], : g.shape[1]] = g
    ph.g[:] = gg.ravel().tolist()
    return ph


def direct_hxv_ext(model: Model, nup, ndw, v, sundry=(), phonons=None):
    """directMatVec_normal_main with coulomb_sundry and DimPh = Nph+1 phonon slices
    (direct/HxV_sundry.f90, HxV_ph.f90, HxV_eph.f90).  phonons = dict(Nph=, w0=, g=, A=0)."""
    v = np.ascontiguousarray(v, np.float64)
    hv = np.empty_like(v)
    sundry = list(sundry)
    arr = sundry_array(sundry)
    ph = phonon_struct(**phonons) if phonons else None
    rc = lib().ora_direct_hxv_ext(C.byref(model.params()), nup, ndw, len(sundry),
                                  C.cast(arr, C.c_void_p), C.byref(ph) if ph else None, v, hv)
    if rc == -2:
        raise ValueError("In NORMAL mode, operators that change the total spin are forbidden")
    assert rc == 0
    return hv


# --------------------------------------------------------------------------------------
# ed_total_ud = F: orbital-resolved sectors (Ns_Ud = Norb factors of Ns_Orb = 1+Nbath levels per spin)
# --------------------------------------------------------------------------------------
def orbs_dims(model: Model, nups, ndws):
    """DimUps / DimDws of the sector (Nups(1:Norb), Ndws(1:Norb)) (ED_SETUP.f90:998-1033)."""
    nso = model.Nbath + 1
    return [binomial(nso, n) for n in nups], [binomial(nso, n) for n in ndws]


def orbs_direct_hxv(model: Model, nups, ndws, v):
    """directMatVec_normal_orbs (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:133-226; direct/Orbs/HxV_local.f90,
    HxV_up.f90, HxV_dw.f90), DimPh=1.  State index i = mixed radix over [DimUps, DimDws], first
    factor fastest (state2indices, ED_SECTOR.f90:1691-1702); every factor is a sector of the
    Ns_Orb = 1+Nbath levels {impurity, bath 1..Nbath} of one orbital and spin (build_sector
    :217-242); breorder puts the per-orbital bits at the global sites (ED_AUX_FUNX.f90:425-441).
    Only bath_type=normal (hybrid stops, ED_SETUP.f90:124; no inter-orbital one-body terms).
    Pure-python loops: small cases only."""
    assert model.bath_type == "normal"
    p = model.params()
    No, Nb, Ns = model.Norb, model.Nbath, model.Ns
    nso = Nb + 1
    dups, ddws = orbs_dims(model, nups, ndws)
    dims = dups + ddws
    maps = [build_map(nso, n) for n in list(nups) + list(ndws)]
    index = [{int(m): k for k, m in enumerate(mp)} for mp in maps]
    strides = np.cumprod([1] + dims[:-1]).tolist()
    Dim = int(np.prod(dims))
    v = np.ascontiguousarray(v, np.float64)
    assert v.size == Dim
    Hv = np.zeros(Dim)
    eloc = np.array(p.eloc[:]).reshape(2, MAXORB, MAXORB)
    Uloc = np.array(p.Uloc[:])
    Ust = np.array(p.Ust[:]).reshape(MAXORB, MAXORB)
    Jh = np.array(p.Jh[:]).reshape(MAXORB, MAXORB)
    sf = np.array(p.spin_field_z[:])
    bd = np.array(p.bath_diag[:]).reshape(2, MAXORB, MAXBATH)
    dh = np.array(p.diag_hybr[:]).reshape(2, MAXORB, MAXBATH)
    for i in range(Dim):
        idx, c = [], i
        for d in dims:
            idx.append(c % d)
            c //= d
        pats = [int(maps[f][idx[f]]) for f in range(2 * No)]
        Nup, Ndw = [0] * Ns, [0] * Ns
        for a in range(No):                       # breorder
            Nup[a], Ndw[a] = pats[a] & 1, pats[a + No] & 1
            for k in range(Nb):
                site = model.bath_stride(a, k) - 1
                Nup[site] = (pats[a] >> (1 + k)) & 1
                Ndw[site] = (pats[a + No] >> (1 + k)) & 1
        # ---- Orbs/HxV_local.f90
        h = 0.0
        for a in range(No):
            h += eloc[0, a, a] * Nup[a] + eloc[1, a, a] * Ndw[a] - model.xmu * (Nup[a] + Ndw[a])
        if np.any(sf != 0):
            for a in range(No):
                h += sf[a] * (Nup[a] - Ndw[a])
        for a in range(No):
            h += Uloc[a] * Nup[a] * Ndw[a]
        if No > 1:
            for a in range(No):
                for b in range(a + 1, No):
                    h += Ust[a, b] * (Nup[a] * Ndw[b] + Nup[b] * Ndw[a])
                    h += (Ust[a, b] - Jh[a, b]) * (Nup[a] * Nup[b] + Ndw[a] * Ndw[b])
        if model.hfmode:
            for a in range(No):
                h += -0.5 * Uloc[a] * (Nup[a] + Ndw[a]) + 0.25 * Uloc[a]
            if No > 1:
                for a in range(No):
                    for b in range(a + 1, No):
                        nn = Nup[a] + Ndw[a] + Nup[b] + Ndw[b]
                        # NB 0.25 here, 0.5 in the ed_total_ud=T fragment (direct/HxV_local.f90:66-67)
                        h += -0.5 * Ust[a, b] * nn + 0.25 * Ust[a, b]
                        h += -0.5 * (Ust[a, b] - Jh[a, b]) * nn + 0.25 * (Ust[a, b] - Jh[a, b])
        for a in range(model.Nfoo):
            for k in range(Nb):
                site = model.bath_stride(a, k) - 1
                h += bd[0, a, k] * Nup[site] + bd[1, a, k] * Ndw[site]
        Hv[i] += h * v[i]
        # ---- Orbs/HxV_up.f90 / HxV_dw.f90: scatter Hv(target) += h * vin(j)
        for spin in range(2):
            for a in range(No):
                f = a + spin * No
                m = pats[f]
                for k in range(Nb):
                    ialfa = 2 + k                      # 1-based position of bath level k+1
                    amp = dh[spin, a, k]
                    if amp == 0.0:
                        continue
                    imp, bth = m & 1, (m >> (1 + k)) & 1
                    out, s1, s2 = C.c_int32(), C.c_double(), C.c_double()
                    if imp == 1 and bth == 0:
                        lib().ora_c(1, m, C.byref(out), C.byref(s1))
                        lib().ora_cdg(ialfa, out.value, C.byref(out), C.byref(s2))
                    elif imp == 0 and bth == 1:
                        lib().ora_c(ialfa, m, C.byref(out), C.byref(s1))
                        lib().ora_cdg(1, out.value, C.byref(out), C.byref(s2))
                    else:
                        continue
                    t = i + (index[f][out.value] - idx[f]) * strides[f]
                    Hv[t] += amp * s1.value * s2.value * v[i]
    return Hv


def orbs_embedding(model: Model, nups, ndws):
    """Index (0-based) in the ed_total_ud=T sector (sum nups, sum ndws) of every state of the orbital-
    resolved sector, through breorder (ED_AUX_FUNX.f90:425-441)."""
    No, Nb, Ns = model.Norb, model.Nbath, model.Ns
    nso = Nb + 1
    dups, ddws = orbs_dims(model, nups, ndws)
    dims = dups + ddws
    maps = [build_map(nso, n) for n in list(nups) + list(ndws)]
    mu_t, md_t = build_map(Ns, sum(nups)), build_map(Ns, sum(ndws))
    iu = {int(m): k for k, m in enumerate(mu_t)}
    idd = {int(m): k for k, m in enumerate(md_t)}
    out = np.zeros(int(np.prod(dims)), np.int64)
    for i in range(out.size):
        c, pats = i, []
        for f, d in enumerate(dims):
            pats.append(int(maps[f][c % d]))
            c //= d
        g = [0, 0]
        for spin in range(2):
            for a in range(No):
                m = pats[a + spin * No]
                g[spin] |= (m & 1) << a
                for k in range(Nb):
                    g[spin] |= ((m >> (1 + k)) & 1) << (model.bath_stride(a, k) - 1)
        out[i] = iu[g[0]] + idd[g[1]] * len(mu_t)
    return out


def direct_hxv_mpi(model: Model, nup, ndw, v, P, nthreads=1):
    v = np.ascontiguousarray(v, np.float64)
    hv = np.empty_like(v)
    rc = lib().ora_direct_hxv_mpi(C.byref(model.params()), nup, ndw, P, nthreads, v, hv)
    assert rc == 0
    return hv


def direct_hxv_mpi_sample(model: Model, nup, ndw, v, P, nrun, nthreads):
    """Timing sample: run ranks [0,nrun) of P emulated ranks; returns seconds (product only)."""
    v = np.ascontiguousarray(v, np.float64)
    hv = np.empty_like(v)
    sec = C.c_double(0.0)
    rc = lib().ora_direct_hxv_mpi_sample(C.byref(model.params()), nup, ndw, P, nrun, nthreads, v,
                                         hv, C.byref(sec))
    assert rc == 0
    return sec.value


def stored_hxv(model: Model, nup, ndw, v):
    v = np.ascontiguousarray(v, np.float64)
    hv = np.empty_like(v)
    rc = lib().ora_stored_hxv(C.byref(model.params()), nup, ndw, v, hv)
    assert rc == 0
    return hv


def stored_hxv_mpi(model: Model, nup, ndw, v, P, nthreads=1, ncalls=1):
    v = np.ascontiguousarray(v, np.float64)
    hv = np.empty_like(v)
    sec = C.c_double(0.0)
    rc = lib().ora_stored_hxv_mpi(C.byref(model.params()), nup, ndw, P, nthreads, ncalls, v, hv,
                                  C.byref(sec))
    assert rc == 0
    return hv, sec.value


def stored_lanczos_gs(model: Model, nup, ndw, v0, nitermax=300, threshold=1e-12, ncheck=10, P=8,
                      nthreads=8):
    """Pass 1 of sp_lanc_eigh (recurrence + stopping rule of lanc_eigh below) run in C on the
    stored operator of stored_hxv_mpi: (egs, niter, alanc, blanc, seconds).  For sectors of the
    BASELINE configs' own size, where the numpy recurrence takes minutes."""
    v0 = np.ascontiguousarray(v0, np.float64)
    a = np.zeros(nitermax)
    b = np.zeros(nitermax)
    egs, nit, sec = C.c_double(0.0), C.c_int(0), C.c_double(0.0)
    rc = lib().ora_stored_lanczos_gs(C.byref(model.params()), nup, ndw, P, nthreads, nitermax,
                                     threshold, ncheck, v0, C.byref(egs), C.byref(nit), a, b,
                                     C.byref(sec))
    assert rc == 0
    return egs.value, nit.value, a[:nit.value], b[:nit.value], sec.value


def hop_csr(model: Model, spin: int, nel: int):
    """spH0ups(1)/spH0dws(1) as CSR in the reference's insertion order; cols 1-based."""
    p = C.byref(model.params())
    dim = binomial(model.Ns, nel)
    nnz = lib().ora_build_hop_csr(p, spin, nel, None, None, None)
    rp = np.zeros(dim + 1, np.int64)
    cols = np.zeros(max(nnz, 1), np.int32)
    vals = np.zeros(max(nnz, 1), np.float64)
    lib().ora_build_hop_csr(p, spin, nel, rp.ctypes.data, cols.ctypes.data, vals.ctypes.data)
    return rp, cols[:nnz], vals[:nnz]


def diag_vector(model: Model, nup, ndw):
    DimUp, DimDw = sector_dims(model.Ns, nup, ndw)
    d = np.empty(DimUp * DimDw)
    lib().ora_build_diag(C.byref(model.params()), nup, ndw, d)
    return d


def nonlocal_csr(model: Model, nup, ndw):
    p = C.byref(model.params())
    DimUp, DimDw = sector_dims(model.Ns, nup, ndw)
    nnz = lib().ora_build_nonlocal_csr(p, nup, ndw, None, None, None)
    rp = np.zeros(DimUp * DimDw + 1, np.int64)
    cols = np.zeros(max(nnz, 1), np.int64)
    vals = np.zeros(max(nnz, 1), np.float64)
    lib().ora_build_nonlocal_csr(p, nup, ndw, rp.ctypes.data, cols.ctypes.data, vals.ctypes.data)
    return rp, cols[:nnz], vals[:nnz]


def dense_H(model: Model, nup, ndw) -> np.ndarray:
    """Hmat = Hd + Hnd + Hdw (x) 1 + 1 (x) Hup (STORED_HxV.f90:199-262)."""
    DimUp, DimDw = sector_dims(model.Ns, nup, ndw)
    H = np.diag(diag_vector(model, nup, ndw))
    rp, cols, vals = nonlocal_csr(model, nup, ndw)
    for i in range(DimUp * DimDw):
        for k in range(rp[i], rp[i + 1]):
            H[i, cols[k] - 1] += vals[k]

    def small(spin, nel, dim):
        rp, cols, vals = hop_csr(model, spin, nel)
        M = np.zeros((dim, dim))
        for i in range(dim):
            for k in range(rp[i], rp[i + 1]):
                M[i, cols[k] - 1] += vals[k]
        return M

    Hup = small(0, nup, DimUp)
    Hdw = small(1, ndw, DimDw)
    H += np.kron(Hdw, np.eye(DimUp)) + np.kron(np.eye(DimDw), Hup)
    return H


def apply_op(model: Model, op: int, iorb: int, spin: int, nup, ndw, v):
    """apply_op_C (op=-1) / apply_op_CDG (op=+1); returns (vector, (nup',ndw'))."""
    Ns = model.Ns
    jn = (nup + (op if spin == 0 else 0), ndw + (op if spin == 1 else 0))
    if min(jn) < 0 or max(jn) > Ns:
        return None, None
    ov = np.zeros(binomial(Ns, jn[0]) * binomial(Ns, jn[1]))
    rc = lib().ora_apply_op(Ns, op, iorb, spin, nup, ndw, np.ascontiguousarray(v, np.float64), ov)
    assert rc == 0
    return ov, jn


# --------------------------------------------------------------------------------------
# Lanczos drivers (SciFortran SF_SP_LINALG semantics).  "hxv" is any callable
# v -> H v, i.e. the role of the spHtimesV_p procedure pointer.
# --------------------------------------------------------------------------------------
def start_vector(n: int, seed: int = 4321) -> np.ndarray:
    """Deterministic pseudo-random start vector shared by the oracle and the product
    (splitmix64 -> uniform(0,1)); the reference uses the compiler's random_number."""
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (idx + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def lanczos_iteration(hxv, it, vin, vout, beta):
    """One step of the three-term recurrence in the sp_lanc_* drivers:
    iter==1: vin/=|vin|; else (vin,vout) <- (vout/beta, -beta*vin);
    vout += H vin; alfa = <vin,vout>; vout -= alfa*vin; beta = |vout|."""
    if it == 1:
        nrm = math.sqrt(float(vin @ vin))
        vin = vin / nrm
        vout = np.zeros_like(vin)
    else:
        vin, vout = vout / beta, -beta * vin
    vout = vout + hxv(vin)
    alfa = float(vin @ vout)
    vout = vout - alfa * vin
    beta = math.sqrt(float(vout @ vout))
    return vin, vout, alfa, beta


def lanc_tridiag(hxv, vin, nlanc, threshold=1e-12):
    """sp_lanc_tridiag(MatVec, vin, alanc(N), blanc(N)): alanc(iter)=alfa, blanc(iter+1)=beta,
    early exit when |beta| < threshold (call site ED_HAMILTONIAN_NORMAL.f90:360-365)."""
    a = np.zeros(nlanc)
    b = np.zeros(nlanc)
    vin = np.array(vin, np.float64)
    vout = np.zeros_like(vin)
    beta = 0.0
    nused = 0
    for it in range(1, nlanc + 1):
        vin, vout, alfa, beta = lanczos_iteration(hxv, it, vin, vout, beta)
        a[it - 1] = alfa
        nused = it
        if abs(beta) < threshold:
            break
        if it < nlanc:
            b[it] = beta
    return a, b, nused


def tridiag_eigh(a, b_sub):
    from scipy.linalg import eigh_tridiagonal

    if len(a) == 1:
        return np.array(a, float), np.ones((1, 1))
    return eigh_tridiagonal(np.asarray(a), np.asarray(b_sub))


def lanc_eigh(hxv, n, nitermax, threshold=1e-12, ncheck=10, v0=None, seed=4321):
    """sp_lanc_eigh(MatVec, egs, vect, Nitermax, threshold): two-pass plain Lanczos GS
    (call site ED_DIAG_NORMAL.f90:206-213).  Pass 1 builds T, diagonalising it each step and
    stopping when the lowest Ritz value moved less than threshold over one check window or
    beta -> 0; pass 2 replays the recurrence accumulating vect += Z(iter,1) v_iter."""
    v0 = start_vector(n, seed) if v0 is None else np.array(v0, np.float64)
    v0 = v0 / math.sqrt(float(v0 @ v0))
    a, b = [], [0.0]
    vin, vout, beta = v0.copy(), np.zeros(n), 0.0
    esave = []
    nlanc = 0
    for it in range(1, min(nitermax, n) + 1):
        vin, vout, alfa, beta = lanczos_iteration(hxv, it, vin, vout, beta)
        if abs(beta) < threshold and it > 1:
            a.append(alfa)
            nlanc = it
            break
        a.append(alfa)
        b.append(beta)
        nlanc = it
        ev, _ = tridiag_eigh(a, b[1:nlanc])
        if nlanc >= ncheck:
            esave.append(ev[0])
            if len(esave) >= 2 and abs(esave[-1] - esave[-2]) <= threshold:
                break
    ev, Z = tridiag_eigh(a[:nlanc], b[1:nlanc])
    egs = float(ev[0])
    vect = np.zeros(n)
    vin, vout, beta = v0.copy(), np.zeros(n), 0.0
    for it in range(1, nlanc + 1):
        vin, vout, _, beta = lanczos_iteration(hxv, it, vin, vout, beta)
        vect += Z[it - 1, 0] * vin
    vect /= math.sqrt(float(vect @ vect))
    return egs, vect, nlanc


# --------------------------------------------------------------------------------------
# ed_solve contract for NORMAL mode at T=0 (enough to reproduce the golden files)
# --------------------------------------------------------------------------------------
@dataclass
class GState:
    e: float
    nup: int
    ndw: int
    vec: np.ndarray


def twin_mask(Ns: int):
    """twin_mask of setup_global (ED_SETUP.f90:592-602), NORMAL mode with ed_twin=T: sectors are
    scanned in index order and a sector with nup /= ndw is switched off when its twin (ndw,nup) is
    still on -- the sectors that stay on are those with nup >= ndw.  Returns {isector: bool}."""
    mask = {i: True for i in range(1, (Ns + 1) ** 2 + 1)}
    for isector in range(1, (Ns + 1) ** 2 + 1):
        nup, ndw = sector_qn(Ns, isector)
        if nup != ndw and mask[sector_index(Ns, ndw, nup)]:
            mask[isector] = False
    return mask


def twin_sector_order(Ns: int, nup: int, ndw: int, nph: int = 0):
    """twin_sector_order(isector, Order) for the NORMAL sector A = (nup,ndw) (ED_SECTOR.f90:1747-1776):
    Order(i) = flip_state([mup, mdw]) + (iph-1)*2**(2Ns) = mdw + mup*2**Ns + ... (flip_state_normal,
    :1787-1795), then sort_array REPLACES the array by its sorting permutation (:1866-1879), so that
    the twin state reads vector_B(i) = vec_A(Order(i)) (es_return_dvector, ED_EIGENSPACE.f90:640-660).
    Returns the 0-based permutation."""
    mu, md = build_map(Ns, nup).astype(np.int64), build_map(Ns, ndw).astype(np.int64)
    flipped = (md[:, None] + (mu[None, :] << Ns)).ravel()       # i = iup + idw*DimUp
    full = np.concatenate([flipped + (k << (2 * Ns)) for k in range(nph + 1)])
    return np.argsort(full, kind="stable")


def diagonalize(model: Model, use_lanczos_above: int | None = None, hxv_kind="stored", ed_twin=None):
    """ed_diag_d (ED_DIAG_NORMAL.f90:76-296): scan the (nup,ndw) sectors (with ed_twin only those
    left on by twin_mask, :110), dense LAPACK when dim <= lanc_dim_threshold else plain Lanczos;
    T=0 state list with the gs_threshold degeneracy rule (:262-278).  With ed_twin a state of a
    sector with nup /= ndw enters the list together with its twin (es_insert_state,
    ED_EIGENSPACE.f90:344-350), whose vector is the re-ordered one (twin_sector_order)."""
    Ns = model.Ns
    thr = model.lanc_dim_threshold if use_lanczos_above is None else use_lanczos_above
    states: list[GState] = []
    oldzero = 1000.0
    fn = stored_hxv if hxv_kind == "stored" else direct_hxv
    ed_twin = model.ed_twin if ed_twin is None else ed_twin
    mask = twin_mask(Ns) if ed_twin else None
    for isector in range(1, (Ns + 1) ** 2 + 1):
        if mask is not None and not mask[isector]:
            continue
        nup, ndw = sector_qn(Ns, isector)
        DimUp, DimDw = sector_dims(Ns, nup, ndw)
        dim = DimUp * DimDw
        if dim <= thr:
            ev, evec = np.linalg.eigh(dense_H(model, nup, ndw))
            e0, v0 = float(ev[0]), evec[:, 0].copy()
        else:
            e0, v0, _ = lanc_eigh(lambda x: fn(model, nup, ndw, x), dim,
                                  min(dim, model.lanc_niter), threshold=1e-12)
        new = [GState(e0, nup, ndw, v0)]
        if ed_twin and nup != ndw:
            new.append(GState(e0, ndw, nup, v0[twin_sector_order(Ns, nup, ndw)]))
        if e0 < oldzero - 10.0 * model.gs_threshold:
            oldzero = e0
            states = new
        elif abs(e0 - oldzero) <= model.gs_threshold:
            oldzero = min(oldzero, e0)
            states.extend(new)
    return states


def diagonalize_finite_t(model: Model, beta: float, nstates_sector: int, nstates_total: int,
                         cutoff: float):
    """ed_diag_d with ed_finite_temp=T (ED_DIAG_NORMAL.f90:262-266): every sector contributes its
    Neigen = min(dim, lanc_nstates_sector) lowest eigenpairs (dense LAPACK branch, :222-250), the
    list keeps the lanc_nstates_total lowest states (es_add_state with size=,
    ED_EIGENSPACE.f90:265-274; both counts rounded up to even, ED_SETUP.f90:279-287), then
    ed_post_diag trims the states with exp(-beta (E-Egs)) <= cutoff (:489-501).
    Returns (states sorted by energy, Boltzmann weights exp(-beta(E-Egs))/zeta, :405-411)."""
    Ns = model.Ns
    nstates_sector += nstates_sector % 2
    nstates_total += nstates_total % 2
    states: list[GState] = []
    for isector in range(1, (Ns + 1) ** 2 + 1):
        nup, ndw = sector_qn(Ns, isector)
        ev, evec = np.linalg.eigh(dense_H(model, nup, ndw))
        for i in range(min(len(ev), nstates_sector)):
            e = float(ev[i])
            if len(states) >= nstates_total:
                worst = max(range(len(states)), key=lambda k: states[k].e)
                if e >= states[worst].e:
                    continue
                states.pop(worst)
            states.append(GState(e, nup, ndw, evec[:, i].copy()))
    states.sort(key=lambda s: s.e)
    egs = states[0].e
    while len(states) > 1 and math.exp(-beta * (states[-1].e - egs)) <= cutoff:
        states.pop()
    w = np.array([math.exp(-beta * (s.e - egs)) for s in states])
    return states, w / w.sum()


def observables(model: Model, states, weights=None):
    """dens / docc of ED_OBSERVABLES_NORMAL.f90:150-215: peso = 1/zeta at T=0, the Boltzmann
    weights of diagonalize_finite_t otherwise."""
    Ns, No = model.Ns, model.Norb
    dens = np.zeros(No)
    docc = np.zeros(No)
    if weights is None:
        weights = [1.0 / len(states)] * len(states)
    for st, peso in zip(states, weights):
        mu, md = build_map(Ns, st.nup), build_map(Ns, st.ndw)
        w = (st.vec ** 2).reshape(len(md), len(mu)) * peso  # [idw, iup]
        for a in range(No):
            nu = ((mu >> a) & 1).astype(float)[None, :]
            nd = ((md >> a) & 1).astype(float)[:, None]
            dens[a] += float((w * (nu + nd)).sum())
            docc[a] += float((w * (nu * nd)).sum())
    return dens, docc



def _energy_variants(model):
    """Model variants whose ground-state expectation values give the energy components of
    local_energy_normal (ED_OBSERVABLES_NORMAL.f90:491-930).  Every variant keeps the bath and the
    hybridisation of the model (so that its hop tables have the model's structure) and differs from
    the `base` variant by one group of impurity terms: <O> = <H_variant> - <H_base>."""
    import dataclasses

    No = model.Norb
    zero_int = dict(Uloc=(0.0,) * No, Ust=0.0, Jh=0.0, Jx=0.0, Jp=0.0, umatrix_lines=(),
                    ed_use_kanamori=True)
    common = dict(xmu=0.0, hfmode=False, spin_field_z=(), exc_field=(0.0, 0.0, 0.0, 0.0), _params=None)
    rep = dataclasses.replace
    return {
        "base": rep(model, hloc=None, **zero_int, **common),
        # <sum impHloc(s,s,a,b) c^+ c> (:556-611): impHloc only, mfHloc belongs to Epot
        "eknot": rep(model, **zero_int, **common),
        # Jx, Jp, mfHloc, sundry, Uloc, Ust, Ust-Jh terms (:613-822), no Hartree shift
        "eint": rep(model, hloc=None, **common),
        # <sum_{a<b} nup_a ndw_b + nup_b ndw_a> (:806): Ust = 1 with Ust - Jh = 0
        "dust": rep(model, hloc=None, **{**zero_int, "Ust": 1.0, "Jh": 1.0}, **common),
        # <sum_{a<b} nup_a nup_b + ndw_a ndw_b> (:819): Ust - Jh = 1 with Ust = 0
        "dund": rep(model, hloc=None, **{**zero_int, "Jh": -1.0}, **common),
        "dse": rep(model, hloc=None, **{**zero_int, "Jx": 1.0}, **common),   # (:633)
        "dph": rep(model, hloc=None, **{**zero_int, "Jp": 1.0}, **common),   # (:660)
    }


def _hartree_energy(model, dens):
    """ed_Ehartree (:825-838) is diagonal and linear in the occupations: from dens(a)."""
    if not model.hfmode:
        return 0.0
    um = model.umatrix()
    No = model.Norb
    e = 0.0
    for a in range(No):
        e += -0.5 * um["Uloc"][a] * dens[a] + 0.25 * um["Uloc"][a]
    for a in range(No):
        for b in range(a + 1, No):
            for c in (um["Ust"][a, b], um["Ust"][a, b] - um["Jh"][a, b]):
                e += -0.5 * c * (dens[a] + dens[b]) + 0.5 * c
    return e


def local_energy(model: Model, states, weights=None):
    """local_energy_normal (ED_OBSERVABLES_NORMAL.f90:491-930), DimPh = 1: returns
    dict(Epot, Eint, Ehartree, Eknot, Dust, Dund, Dse, Dph) = ed_get_eimp / ed_get_doubles
    (ED_IO/get_energy.f90:7, get_doubles.f90:7).  Every component is the state-list average of
    <v|O|v> for an operator O that is the model Hamiltonian with one group of terms switched on."""
    if weights is None:
        weights = [1.0 / len(states)] * len(states)
    variants = _energy_variants(model)
    acc = {k: 0.0 for k in variants}
    for st, peso in zip(states, weights):
        for k, mv in variants.items():
            hv = direct_hxv_ext(mv, st.nup, st.ndw, st.vec, mv.coulomb_sundry, None)
            acc[k] += peso * float(st.vec @ hv)
    dens, _ = observables(model, states, weights)
    out = {k.capitalize() if k.startswith("d") else "E" + k[1:]: acc[k] - acc["base"]
           for k in variants if k != "base"}
    out["Ehartree"] = _hartree_energy(model, dens)
    out["Epot"] = out["Eint"] + out["Ehartree"]
    return out


def imp_info(model: Model, states, weights=None):
    """ed_imp_info = [s2tot, egs] (ED_OBSERVABLES_NORMAL.f90:180, 452): s2tot = <(sum_a S^z_a)^2>
    with S^z_a = (nup_a - ndw_a)/2 (:160-164), egs = lowest energy of the state list."""
    Ns, No = model.Ns, model.Norb
    if weights is None:
        weights = [1.0 / len(states)] * len(states)
    s2 = 0.0
    for st, peso in zip(states, weights):
        mu, md = build_map(Ns, st.nup), build_map(Ns, st.ndw)
        w = (st.vec ** 2).reshape(len(md), len(mu)) * peso
        sz = np.zeros((len(md), len(mu)))
        for a in range(No):
            sz += 0.5 * (((mu >> a) & 1).astype(float)[None, :] - ((md >> a) & 1).astype(float)[:, None])
        s2 += float((w * sz ** 2).sum())
    return np.array([s2, min(s.e for s in states)])


def gf_poles_weights(model: Model, states, iorb: int, spin: int = 0, hxv_kind="direct",
                     weights=None):
    """lanc_build_gf_normal_diag + add_to_lanczos_gf_normal (ED_GF_NORMAL.f90:131-177,
    363-427): list of (weight, pole); peso = exp(-beta(Ei-Egs))/zeta at finite T (:385-389)."""
    fn = stored_hxv if hxv_kind == "stored" else direct_hxv
    if weights is None:
        weights = [1.0 / len(states)] * len(states)
    out = []
    for st, peso in zip(states, weights):
        for op, isign in ((+1, 1), (-1, -1)):
            seed, jn = apply_op(model, op, iorb, spin, st.nup, st.ndw, st.vec)
            if seed is None:
                continue
            norm2 = float(seed @ seed)
            if norm2 == 0.0:
                continue
            seed = seed / math.sqrt(norm2)
            nlanc = min(len(seed), model.lanc_ngfiter)
            a, b, nused = lanc_tridiag(lambda x: fn(model, jn[0], jn[1], x), seed, nlanc)
            ev, Z = tridiag_eigh(a[:nused], b[1:nused])
            for j in range(nused):
                out.append((norm2 * peso * Z[0, j] ** 2, isign * (ev[j] - st.e)))
    return out


def gf_poles_weights_mix(model: Model, states, iorb: int, jorb: int, spin: int = 0,
                         hxv_kind="direct", weights=None):
    """lanc_build_gf_normal_mix (ED_GF_NORMAL.f90:182-262, real case): poles / weights of the
    auxiliary function built from the seeds (c^+_a + c^+_b)|gs> and (c_a + c_b)|gs>
    (apply_Cops with coefficients [1,1]); G_ab follows from get_impG_normal (:545-560)."""
    fn = stored_hxv if hxv_kind == "stored" else direct_hxv
    if weights is None:
        weights = [1.0 / len(states)] * len(states)
    out = []
    for st, peso in zip(states, weights):
        for op, isign in ((+1, 1), (-1, -1)):
            sa, jn = apply_op(model, op, iorb, spin, st.nup, st.ndw, st.vec)
            sb, _ = apply_op(model, op, jorb, spin, st.nup, st.ndw, st.vec)
            if sa is None:
                continue
            seed = sa + sb
            norm2 = float(seed @ seed)
            if norm2 == 0.0:
                continue
            seed = seed / math.sqrt(norm2)
            nlanc = min(len(seed), model.lanc_ngfiter)
            a, b, nused = lanc_tridiag(lambda x: fn(model, jn[0], jn[1], x), seed, nlanc)
            ev, Z = tridiag_eigh(a[:nused], b[1:nused])
            for j in range(nused):
                out.append((norm2 * peso * Z[0, j] ** 2, isign * (ev[j] - st.e)))
    return out


def impG_matrix(model: Model, states, spin: int, z, weights=None):
    """get_impG_normal (ED_GF_NORMAL.f90:495-575) with offdiag_gf_flag: G_aa from the diagonal
    builder, G_ab = (G_{a+b} - G_aa - G_bb)/2 = G_ba (real case, :553-560).  Returns [Norb,Norb,len(z)]."""
    No = model.Norb
    G = np.zeros((No, No, len(z)), complex)
    for a in range(No):
        G[a, a] = gf_eval(gf_poles_weights(model, states, a, spin, weights=weights), z)
    for a in range(No):
        for b in range(a + 1, No):
            mix = gf_eval(gf_poles_weights_mix(model, states, a, b, spin, weights=weights), z)
            G[a, b] = G[b, a] = 0.5 * (mix - G[a, a] - G[b, b])
    return G


def delta_matrix(model: Model, spin: int, z):
    """delta_bath_array (ED_BATH/delta_functions, ed_mode=normal): normal -> diagonal
    sum_k V_ak^2/(z-e_ak) (delta_normal.f90:33-42); hybrid -> sum_k V_ak V_bk/(z-e_k)
    (delta_hybrid.f90:30-41); replica / general -> sum_k V_k (z - H_k)^-1 V_k (delta_replica.f90:27-38,
    delta_general.f90, V_k = diag(vg))."""
    No, Nb = model.Norb, model.Nbath
    if model.bath_e is None:
        model.default_bath()
    D = np.zeros((No, No, len(z)), complex)
    if model.bath_type == "normal":
        for a in range(No):
            D[a, a] = (model.bath_v[spin, a][None, :] ** 2 / (z[:, None] - model.bath_e[spin, a][None, :])).sum(1)
    elif model.bath_type == "hybrid":
        e = model.bath_e[spin, 0]
        for a in range(No):
            for b in range(No):
                D[a, b] = (model.bath_v[spin, a][None, :] * model.bath_v[spin, b][None, :] / (z[:, None] - e[None, :])).sum(1)
    else:
        for k in range(Nb):
            Hk = model.hbath[spin, :, :, k]
            V = np.diag(model.bath_v[spin, :, k])
            for i, zi in enumerate(z):
                D[:, :, i] += V @ np.linalg.inv(zi * np.eye(No) - Hk) @ V
    return D


def sigma_matrix_matsubara(model: Model, states, spin: int, Lmats: int, weights=None):
    """get_Sigma_normal (ED_GF_NORMAL.f90:698-739): Sigma = G0^-1 - G^-1 with
    G0^-1 = (z+xmu) 1 - impHloc - Delta (invg0_hyrege.f90:22-30), G inverted as an orbital matrix for
    bath_type /= normal (:726-729).  Returns (wm, Sigma[Norb,Norb,Lmats])."""
    wm = math.pi / model.beta * (2 * np.arange(1, Lmats + 1) - 1)
    z = 1j * wm
    No = model.Norb
    G = impG_matrix(model, states, spin, z, weights)
    D = delta_matrix(model, spin, z)
    hl = np.zeros((No, No)) if model.hloc is None else np.asarray(model.hloc[spin], float)
    S = np.zeros_like(G)
    for i, zi in enumerate(z):
        invg0 = (zi + model.xmu) * np.eye(No) - hl - D[:, :, i]
        if model.bath_type == "normal":
            invg = np.diag(1.0 / np.diag(G[:, :, i]))
        else:
            invg = np.linalg.inv(G[:, :, i])
        S[:, :, i] = invg0 - invg
    return wm, S


def gf_eval(pw, z):
    g = np.zeros_like(z, dtype=complex)
    for w, p in pw:
        g += w / (z - p)
    return g


def sigma_matsubara(model: Model, pw, iorb: int, spin: int, Lmats: int):
    """Sigma = G0^-1 - G^-1 for the normal bath (ED_GF_NORMAL.f90:698-739,
    invg0_normal.f90:22-28, delta_normal.f90:33-42)."""
    wm = math.pi / model.beta * (2 * np.arange(1, Lmats + 1) - 1)
    z = 1j * wm
    if model.bath_e is None:
        model.default_bath()
    e = model.bath_e[spin, iorb if model.bath_type == "normal" else 0]
    v = model.bath_v[spin, iorb]
    delta = (v[None, :] ** 2 / (z[:, None] - e[None, :])).sum(axis=1)
    hl = 0.0 if model.hloc is None else model.hloc[spin, iorb, iorb]
    invg0 = z + model.xmu - hl - delta
    return wm, invg0 - 1.0 / gf_eval(pw, z)


def momenta(wm, F, nmom=4):
    """compute_momentum of test/src/COMMON.f90:178-192."""
    den = np.abs(F).sum()
    return np.array([(np.abs(F) * wm ** n).sum() / den for n in range(1, nmom + 1)])
