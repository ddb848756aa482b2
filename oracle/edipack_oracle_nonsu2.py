"""CPU restatement of EDIpack's NONSU2-mode stored Hamiltonian (ED_SPARSE_H=T), numpy.

TEST INFRASTRUCTURE ONLY (see oracle/ed_oracle.h): plays the role of the Fortran host that
builds ``spH0`` for the stored-H path; the product never imports it.

Follows, per state of the sector:
  build_sector (nonsu2, Jz_basis=F)   src/singlesite/ED_SECTOR.f90:335-368  (m = iup + idw*2**Ns,
                                       idw outer / iup inner -> ascending m, popcount = Ntot)
  c / cdg on the 2*Ns-bit state       src/singlesite/ED_AUX_FUNX.f90:334-384 (sign counts ALL lower bits)
  stored/Himp.f90                     diagonal :15-21, same-spin hops :38-80, spin-flip :85-110,
                                       spin_field :243-300 (z on the diagonal, x/y off-diagonal)
  stored/Hint.f90                     density-density + Hartree shifts :13-52, S-E :63-90, P-H :96-124
  stored/Hbath.f90                    normal/hybrid diagonal :12-27; replica/general diagonal :30-45,
                                       same-spin bath hops :49-100, spin-flip bath hops :103-133
  stored/Himp_bath.f90                spin-conserving hybridisation :10-67, spin-flip u :72-136
                                       (normal/hybrid only)
Matrix convention (sp_insert_element(spH0,htmp,i,j)): row i = the state the operators act on,
column j = the resulting state, value = conjg(amplitude)*sign; duplicates accumulate
(ED_SPARSE_MATRIX.f90:346-357).

Parity status: pinned to test/src/{HYBRID,NORMAL,REPLICA,GENERAL}_NONSU2/{evals,dens,docc}.check
(+ magX / exciton) (tests/test_oracle_golden_nonsu2.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class ModelNonsu2:
    Norb: int = 2
    Nbath: int = 4
    bath_type: str = "hybrid"          # normal | hybrid | replica | general
    Uloc: tuple = (1.0, 1.0)
    Ust: float = 0.0
    Jh: float = 0.0
    Jx: float = 0.0
    Jp: float = 0.0
    xmu: float = 0.0
    hfmode: bool = True
    ed_hw_bath: float = 2.0
    hloc: np.ndarray | None = None     # complex [2,2,Norb,Norb]  impHloc(ispin,jspin,iorb,jorb)
    bath_e: np.ndarray | None = None   # [2, Nfoo, Nbath]
    bath_v: np.ndarray | None = None   # [2, Norb, Nbath]
    bath_u: np.ndarray | None = None   # [2, Norb, Nbath]  spin-flip hybridisation
    spin_field: np.ndarray | None = None  # [Norb, 3]
    hbath: np.ndarray | None = None    # replica/general: complex [2,2,Norb,Norb,Nbath] Hbath_tmp

    @property
    def Ns(self):  # ED_SETUP.f90:118-126
        return self.Nbath + self.Norb if self.bath_type == "hybrid" else (self.Nbath + 1) * self.Norb

    @property
    def Nfoo(self):
        return 1 if self.bath_type == "hybrid" else self.Norb

    def stride(self, a, k):  # getBathStride(a+1,k+1), 1-based site (ED_SETUP.f90:605-622)
        if self.bath_type == "hybrid":
            return self.Norb + k + 1
        if self.bath_type in ("replica", "general"):
            return (a + 1) + (k + 1) * self.Norb
        return self.Norb + a * self.Nbath + k + 1

    def default_bath(self):
        """init_dmft_bath, ED_BATH_DMFT.f90:211-244 (nonsu2: u = v)."""
        Nb, hw = self.Nbath, self.ed_hw_bath
        e = np.zeros(Nb)
        e[0], e[-1] = -hw, hw
        Nh = Nb // 2
        if Nb % 2 == 0 and Nb >= 4:
            de = hw / max(Nh - 1, 1)
            e[Nh - 1], e[Nh] = -0.1, 0.1
            for i in range(2, Nh):
                e[i - 1] = -hw + (i - 1) * de
                e[Nb - i] = hw - (i - 1) * de
        elif Nb % 2 != 0 and Nb >= 3:
            de = hw / Nh
            e[Nh] = 0.0
            for i in range(2, Nh + 1):
                e[i - 1] = -hw + (i - 1) * de
                e[Nb - i] = hw - (i - 1) * de
        v = max(0.1, 1.0 / math.sqrt(Nb))
        self.bath_e = np.broadcast_to(e, (2, self.Nfoo, Nb)).copy()
        self.bath_v = np.full((2, self.Norb, Nb), v)
        self.bath_u = np.full((2, self.Norb, Nb), v)
        return self


def build_sector(Ns: int, Ntot: int) -> np.ndarray:
    """Ascending list of m = iup + idw*2**Ns with popcount(m) = Ntot (ED_SECTOR.f90:351-368)."""
    allm = np.arange(1 << (2 * Ns), dtype=np.int64)
    pc = np.zeros_like(allm)
    x = allm.copy()
    while x.any():
        pc += x & 1
        x >>= 1
    return allm[pc == Ntot]


def _c(pos, m):
    """c(pos,in,out,fsgn) with pos 1-based; returns (out, sign) or None (ED_AUX_FUNX.f90:334-357)."""
    b = pos - 1
    if not (m >> b) & 1:
        return None
    sgn = -1.0 if bin(m & ((1 << b) - 1)).count("1") & 1 else 1.0
    return m & ~(1 << b), sgn


def _cdg(pos, m):
    b = pos - 1
    if (m >> b) & 1:
        return None
    sgn = -1.0 if bin(m & ((1 << b) - 1)).count("1") & 1 else 1.0
    return m | (1 << b), sgn


def stored_H(model: ModelNonsu2, Ntot: int, only_rows=None):
    """ed_buildH_nonsu2_main (ED_HAMILTONIAN_NONSU2_STORED_HxV.f90:29-190) for one sector:
    returns (map, rowptr, cols [1-based], vals [complex]) in list-of-rows insertion order.
    only_rows (0-based sector indices, optional): generate just those rows (rowptr then runs over
    the subset, columns stay global) -- the bounded sample used at the BASELINE cfg 5 size, where
    the per-row Python loops over all 705 432 rows would take tens of minutes."""
    if model.bath_e is None:
        model.default_bath()
    Ns, No, Nb = model.Ns, model.Norb, model.Nbath
    smap = build_sector(Ns, Ntot)
    index = {int(m): i + 1 for i, m in enumerate(smap)}  # binary_search -> 1-based index
    hloc = np.zeros((2, 2, No, No), complex) if model.hloc is None else np.asarray(model.hloc, complex)
    U = np.asarray(model.Uloc, float)
    Ust = np.full((No, No), model.Ust) - np.diag(np.full(No, model.Ust))
    Jh = np.full((No, No), model.Jh) - np.diag(np.full(No, model.Jh))
    Jx = np.full((No, No), model.Jx) - np.diag(np.full(No, model.Jx))
    Jp = np.full((No, No), model.Jp) - np.diag(np.full(No, model.Jp))
    sf = np.zeros((No, 3)) if model.spin_field is None else np.asarray(model.spin_field, float)
    rows = []
    for m_ in (smap if only_rows is None else smap[np.asarray(only_rows, np.int64)]):
        m = int(m_)
        row = {}  # column -> value, insertion-ordered; duplicates accumulate

        def ins(val, j):
            row[j] = row.get(j, 0.0) + val

        ib = [(m >> k) & 1 for k in range(2 * Ns)]
        nup = [float(ib[a]) for a in range(No)]
        ndw = [float(ib[a + Ns]) for a in range(No)]
        i = index[m]

        def hop(alfa, beta, amp):
            """c(beta) then cdg(alfa); inserts conjg(amp)*sg1*sg2 at (i, j)."""
            r1 = _c(beta, m)
            if r1 is None:
                return
            r2 = _cdg(alfa, r1[0])
            if r2 is None:
                return
            ins(np.conj(amp) * r1[1] * r2[1], index[r2[0]])

        # ---- Himp.f90 : local part
        h = 0.0
        for a in range(No):
            h += hloc[0, 0, a, a] * nup[a] + hloc[1, 1, a, a] * ndw[a] - model.xmu * (nup[a] + ndw[a])
        ins(h, i)
        for a in range(No):
            for b in range(No):
                if hloc[0, 0, a, b] != 0 and ib[b] == 1 and ib[a] == 0:
                    hop(a + 1, b + 1, hloc[0, 0, a, b])
                if hloc[1, 1, a, b] != 0 and ib[b + Ns] == 1 and ib[a + Ns] == 0:
                    hop(a + 1 + Ns, b + 1 + Ns, hloc[1, 1, a, b])
        for isp in range(2):
            jsp = 1 - isp
            for a in range(No):
                for b in range(No):
                    ialfa, ibeta = a + 1 + isp * Ns, b + 1 + jsp * Ns
                    if hloc[isp, jsp, a, b] != 0 and ib[ibeta - 1] == 1 and ib[ialfa - 1] == 0:
                        hop(ialfa, ibeta, hloc[isp, jsp, a, b])
        if np.any(sf != 0):
            # NB the reference reuses the running htmp here (Himp.f90:247-250); it is zero unless a
            # hop was inserted just before -- restated as the intended F_z term only
            ins(sum(sf[a, 2] * (nup[a] - ndw[a]) for a in range(No)), i)
            for a in range(No):
                for (src, dst, sy) in ((a + 1, a + 1 + Ns, -1j), (a + 1 + Ns, a + 1, 1j)):
                    r1 = _c(src, m)
                    if r1 is None:
                        continue
                    r2 = _cdg(dst, r1[0])
                    if r2 is None:
                        continue
                    j = index[r2[0]]
                    ins(sf[a, 0] * r1[1] * r2[1], j)
                    ins(sy * sf[a, 1] * r1[1] * r2[1], j)
        # ---- Hint.f90
        h = 0.0
        for a in range(No):
            h += U[a] * nup[a] * ndw[a]
        for a in range(No):
            for b in range(a + 1, No):
                h += Ust[a, b] * (nup[a] * ndw[b] + nup[b] * ndw[a])
                h += (Ust[a, b] - Jh[a, b]) * (nup[a] * nup[b] + ndw[a] * ndw[b])
        if model.hfmode:
            for a in range(No):
                h += -0.5 * U[a] * (nup[a] + ndw[a]) + 0.25 * U[a]
            for a in range(No):
                for b in range(a + 1, No):
                    nn = nup[a] + ndw[a] + nup[b] + ndw[b]
                    h += -0.5 * Ust[a, b] * nn + 0.5 * Ust[a, b]
                    h += -0.5 * (Ust[a, b] - Jh[a, b]) * nn + 0.5 * (Ust[a, b] - Jh[a, b])
        ins(h, i)
        if No > 1 and np.any(Jx != 0):
            for a in range(No):
                for b in range(No):
                    if a != b and ib[b] == 1 and ib[a + Ns] == 1 and ib[b + Ns] == 0 and ib[a] == 0:
                        k, s = m, 1.0
                        for f, pos in ((_c, b + 1), (_c, a + 1 + Ns), (_cdg, b + 1 + Ns), (_cdg, a + 1)):
                            k, sg = f(pos, k)
                            s *= sg
                        ins(Jx[a, b] * s, index[k])
        if No > 1 and np.any(Jp != 0):
            for a in range(No):
                for b in range(No):
                    if a != b and ib[b] == 1 and ib[b + Ns] == 1 and ib[a + Ns] == 0 and ib[a] == 0:
                        k, s = m, 1.0
                        for f, pos in ((_c, b + 1), (_c, b + 1 + Ns), (_cdg, a + 1 + Ns), (_cdg, a + 1)):
                            k, sg = f(pos, k)
                            s *= sg
                        ins(Jp[a, b] * s, index[k])
        replica = model.bath_type in ("replica", "general")
        if not replica:
            # ---- Hbath.f90 (normal / hybrid)
            h = 0.0
            for a in range(model.Nfoo):
                for k in range(Nb):
                    s = model.stride(a, k)
                    h += model.bath_e[0, a, k] * ib[s - 1] + model.bath_e[1, a, k] * ib[s - 1 + Ns]
            ins(h, i)
        else:
            # ---- Hbath.f90 (replica / general): bath_diag(s,a,k) = Hbath_tmp(s,s,a,a,k)
            hb = np.asarray(model.hbath, complex)
            h = 0.0
            for k in range(Nb):
                for a in range(No):
                    s = model.stride(a, k)
                    h += hb[0, 0, a, a, k] * ib[s - 1] + hb[1, 1, a, a, k] * ib[s - 1 + Ns]
            ins(h, i)
            for k in range(Nb):          # same-spin bath hops
                for a in range(No):
                    for b in range(No):
                        ia, ibt = model.stride(a, k), model.stride(b, k)
                        if hb[0, 0, a, b, k] != 0 and ib[ibt - 1] == 1 and ib[ia - 1] == 0:
                            hop(ia, ibt, hb[0, 0, a, b, k])
                        if hb[1, 1, a, b, k] != 0 and ib[ibt - 1 + Ns] == 1 and ib[ia - 1 + Ns] == 0:
                            hop(ia + Ns, ibt + Ns, hb[1, 1, a, b, k])
            for k in range(Nb):          # spin-flip bath hops
                for isp in range(2):
                    jsp = 1 - isp
                    for a in range(No):
                        for b in range(No):
                            ia, ibt = model.stride(a, k) + isp * Ns, model.stride(b, k) + jsp * Ns
                            if hb[isp, jsp, a, b, k] != 0 and ib[ibt - 1] == 1 and ib[ia - 1] == 0:
                                hop(ia, ibt, hb[isp, jsp, a, b, k])
        # ---- Himp_bath.f90
        for a in range(No):
            for k in range(Nb):
                ms = model.stride(a, k)
                for sp in range(2):
                    v = model.bath_v[sp, a, k]
                    if v != 0:
                        hop(ms + sp * Ns, a + 1 + sp * Ns, v)   # imp -> bath
                        hop(a + 1 + sp * Ns, ms + sp * Ns, v)   # bath -> imp
        for a in range(No):
            if replica:                  # no spin-flip hybridisation (Himp_bath.f90:72)
                break
            for k in range(Nb):
                ms = model.stride(a, k)
                u1, u2 = model.bath_u[0, a, k], model.bath_u[1, a, k]
                if u1 != 0:  # IMP UP <--> BATH DW (amplitude inserted without conjg, it is real)
                    hop(ms + Ns, a + 1, u1)
                    hop(a + 1, ms + Ns, u1)
                if u2 != 0:  # IMP DW <--> BATH UP
                    hop(ms, a + 1 + Ns, u2)
                    hop(a + 1 + Ns, ms, u2)
        rows.append(row)
    rowptr = np.zeros(len(rows) + 1, np.int64)
    cols, vals = [], []
    for r, row in enumerate(rows):
        for j, val in row.items():
            cols.append(j)
            vals.append(val)
        rowptr[r + 1] = len(cols)
    return smap, rowptr, np.array(cols, np.int32), np.array(vals, complex)


def to_dense(rowptr, cols, vals):
    n = len(rowptr) - 1
    H = np.zeros((n, n), complex)
    for i in range(n):
        for k in range(rowptr[i], rowptr[i + 1]):
            H[i, cols[k] - 1] += vals[k]
    return H


def csr_matvec(rowptr, cols, vals, v):
    """sp_matvec_matrix_csr_c (ED_SPARSE_MATRIX.f90:778): Hv(i) = sum_k vals(k) * v(cols(k))."""
    out = np.zeros(len(rowptr) - 1, complex)
    for i in range(len(out)):
        s = slice(rowptr[i], rowptr[i + 1])
        out[i] = np.dot(vals[s], v[cols[s] - 1])
    return out


def observables(model: ModelNonsu2, smap, vec):
    """dens, docc, magX of one state (ED_OBSERVABLES_NONSU2.f90: n_up+n_dw, n_up*n_dw,
    <c+_up c_dw + c+_dw c_up> = |(c_up + c_dw)|gs>|^2 - n_up - n_dw, :265-294)."""
    Ns, No = model.Ns, model.Norb
    w = np.abs(vec) ** 2
    dens, docc, magx = np.zeros(No), np.zeros(No), np.zeros(No)
    index = {int(m): i for i, m in enumerate(smap)}
    for a in range(No):
        nu = ((smap >> a) & 1).astype(float)
        nd = ((smap >> (a + Ns)) & 1).astype(float)
        dens[a] = float((w * (nu + nd)).sum())
        docc[a] = float((w * nu * nd).sum())
        acc = 0.0 + 0.0j
        for i, m_ in enumerate(smap):  # <gs| c+_up c_dw |gs> + h.c.
            m = int(m_)
            r1 = _c(a + 1 + Ns, m)
            if r1 is None:
                continue
            r2 = _cdg(a + 1, r1[0])
            if r2 is None:
                continue
            acc += np.conj(vec[index[r2[0]]]) * r1[1] * r2[1] * vec[i]
        magx[a] = 2.0 * acc.real
    return dens, docc, magx


def twin_sector_order(Ns: int, Ntot: int):
    """twin_sector_order for the nonsu2 sector A = Ntot (ED_SECTOR.f90:1747-1776, flip_state_other
    :1797-1817: every bit complemented, Ntot -> 2Ns - Ntot); sort_array replaces the array by its
    sorting permutation: vector_B(i) = vec_A(Order(i)).  0-based."""
    smap = build_sector(Ns, Ntot)
    flipped = (~smap) & ((1 << (2 * Ns)) - 1)
    return np.argsort(flipped, kind="stable")


def _energy_variants(model):
    """Variants of the model whose ground-state expectation values give the components of
    local_energy_{mode} (ED_OBSERVABLES_NONSU2.f90:630-850 / ED_OBSERVABLES_SUPERC.f90): each keeps the
    bath (and, for superc, the pairing terms) and differs from `base` by one group of impurity terms,
    <O> = <H_variant> - <H_base>."""
    import dataclasses

    No = model.Norb
    zero_int = dict(Uloc=(0.0,) * No, Ust=0.0, Jh=0.0, Jx=0.0, Jp=0.0)
    common = dict(xmu=0.0, hfmode=False)
    if hasattr(model, "spin_field"):
        common["spin_field"] = None
    rep = dataclasses.replace
    return {
        "base": rep(model, hloc=None, **zero_int, **common),
        "eknot": rep(model, **zero_int, **common),
        "eint": rep(model, hloc=None, **common),
        "dust": rep(model, hloc=None, **{**zero_int, "Ust": 1.0, "Jh": 1.0}, **common),
        "dund": rep(model, hloc=None, **{**zero_int, "Jh": -1.0}, **common),
        "dse": rep(model, hloc=None, **{**zero_int, "Jx": 1.0}, **common),
        "dph": rep(model, hloc=None, **{**zero_int, "Jp": 1.0}, **common),
    }


def local_energy(model, qn: int, vec, dens):
    """ed_Epot, ed_Eint, ed_Ehartree, ed_Eknot, ed_Dust, ed_Dund, ed_Dse, ed_Dph of one state with
    weight 1 (local_energy of this mode), every component as <v|O|v> of a term-restricted stored H."""
    acc = {}
    for k, mv in _energy_variants(model).items():
        smap, rp, cj, va = stored_H(mv, qn)
        acc[k] = float(np.vdot(vec, csr_matvec(rp, cj, va, vec)).real)
    out = {("D" + k[1:]) if k.startswith("d") else ("E" + k[1:]): acc[k] - acc["base"] for k in acc if k != "base"}
    eh = 0.0
    if model.hfmode:
        No = model.Norb
        U = np.asarray(model.Uloc, float)
        for a in range(No):
            eh += -0.5 * U[a] * dens[a] + 0.25 * U[a]
        for a in range(No):
            for b in range(a + 1, No):
                for c in (model.Ust, model.Ust - model.Jh):
                    eh += -0.5 * c * (dens[a] + dens[b]) + 0.5 * c
    out["Ehartree"] = eh
    out["Epot"] = out["Eint"] + eh
    return out


def exciton(model: ModelNonsu2, Ntot: int, smap, vec, iorb: int = 0, jorb: int = 1):
    """Excitonic order parameters [S0, Tx, Ty, Tz](iorb,jorb) of one state with weight 1
    (ED_OBSERVABLES_NONSU2.f90:300-425): norms of the seeds apply_Cops(v,[1,c],[-1,-1],[a,b],[s,s'])
    in the sector Ntot-1, combined with dens and magZ (:420-423)."""
    Ns = model.Ns
    tmap = build_sector(Ns, Ntot - 1)
    tindex = {int(m): i for i, m in enumerate(tmap)}

    def seed_norm2(ca, sa, cb, sb):
        eta = np.zeros(len(tmap), complex)
        for i, m_ in enumerate(smap):
            m = int(m_)
            for coef, orb, spin in ((ca, iorb, sa), (cb, jorb, sb)):
                r = _c(orb + 1 + spin * Ns, m)
                if r is not None:
                    eta[tindex[r[0]]] += coef * r[1] * vec[i]
        return float(np.vdot(eta, eta).real)

    w = np.abs(vec) ** 2
    nu = [((smap >> a) & 1).astype(float) for a in range(model.Norb)]
    nd = [((smap >> (a + Ns)) & 1).astype(float) for a in range(model.Norb)]
    dens = [float((w * (nu[a] + nd[a])).sum()) for a in range(model.Norb)]
    magz = [float((w * (nu[a] - nd[a])).sum()) for a in range(model.Norb)]
    th_uu, th_dd = seed_norm2(1, 0, 1, 0), seed_norm2(1, 1, 1, 1)
    th_ud, th_du = seed_norm2(1, 0, 1, 1), seed_norm2(1, 1, 1, 0)
    om_ud, om_du = seed_norm2(1, 0, -1j, 1), seed_norm2(1, 1, -1j, 0)
    a, b = iorb, jorb
    return np.array([th_uu + th_dd - dens[a] - dens[b],      # S0
                     th_ud + th_du - dens[a] - dens[b],      # Tx
                     om_ud - om_du - magz[a] + magz[b],      # Ty
                     th_uu - th_dd - magz[a] - magz[b]])     # Tz


def lehmann_G(model: ModelNonsu2, Ntot: int, smap, vec, e0: float, z):
    """Exact impurity Green's function matrix of one state with weight 1 from the Lehmann
    representation (dense diagonalisation of the Ntot+1 and Ntot-1 sectors):
      G_ab(z) = <0|c_a (z - H + E0)^-1 c^+_b|0> + <0|c^+_b (z + H - E0)^-1 c_a|0>,
    a = (spin, orbital) -> index orbital + spin*Norb.  What build_impG_nonsu2 (ED_GF_NONSU2.f90)
    obtains from Lanczos continued fractions and their algebraic combinations.  [2No, 2No, len(z)]."""
    Ns, No = model.Ns, model.Norb
    z = np.asarray(z, complex)
    G = np.zeros((2 * No, 2 * No, len(z)), complex)
    bits = [a + sp * Ns for sp in range(2) for a in range(No)]
    for create in (True, False):
        nt = Ntot + (1 if create else -1)
        if nt < 0 or nt > 2 * Ns:
            continue
        tmap, rp, cj, va = stored_H(model, nt)
        ev, U = np.linalg.eigh(to_dense(rp, cj, va))
        tindex = {int(m): i for i, m in enumerate(tmap)}
        amps = []
        for b in bits:
            seed = np.zeros(len(tmap), complex)
            for i, m_ in enumerate(smap):
                r = (_cdg if create else _c)(b + 1, int(m_))
                if r is not None:
                    seed[tindex[r[0]]] += r[1] * vec[i]
            amps.append(U.conj().T @ seed)          # <n| O_b |0>
        A = np.array(amps)                          # [2No, n]
        w = ev - e0
        for i, zi in enumerate(z):
            if create:   # sum_n conj(P_a^n) P_b^n / (z - w_n),  P_b^n = <n|c^+_b|0>
                G[:, :, i] += (A.conj() / (zi - w)[None, :]) @ A.T
            else:        # sum_m conj(Q_b^m) Q_a^m / (z + w_m),  Q_a^m = <m|c_a|0>
                G[:, :, i] += (A / (zi + w)[None, :]) @ A.conj().T
    return G


def delta_matrix(model: ModelNonsu2, z):
    """delta_bath_array for ed_mode=nonsu2 (delta_normal.f90:62-77, delta_hybrid.f90:74-91):
    Delta(s,s',a,b) = sum_{h,k} W(s,h,a,k) W(s',h,b,k) / (z - e(h,k)), W(s,s)=v(s), W(up,dw)=u(up),
    W(dw,up)=u(dw) (get_Whyb_matrix, ED_BATH_AUX.f90:75-87); normal bath: diagonal in the orbitals.
    Index orbital + spin*Norb.  [2No, 2No, len(z)]."""
    No, Nb = model.Norb, model.Nbath
    z = np.asarray(z, complex)
    if model.bath_type in ("replica", "general"):
        # delta_replica.f90:27-38 / delta_general.f90: sum_k V_k (z - H_k)^-1 V_k on the (spin,orbital)
        # space, H_k = nn2so(Hbath_tmp(:,:,:,:,k)), V_k = v_k (replica) or diag(vg_k) (general)
        D = np.zeros((2 * No, 2 * No, len(z)), complex)
        hb = np.asarray(model.hbath, complex)
        for k in range(Nb):
            Hk = hb[..., k].transpose(0, 2, 1, 3).reshape(2 * No, 2 * No)
            V = np.diag(np.concatenate([model.bath_v[0, :, k], model.bath_v[1, :, k]]))
            for i, zi in enumerate(z):
                D[:, :, i] += V @ np.linalg.inv(zi * np.eye(2 * No) - Hk) @ V
        return D
    W = np.zeros((2, 2, No, Nb))
    for sp in range(2):
        W[sp, sp] = model.bath_v[sp]
    W[0, 1], W[1, 0] = model.bath_u[0], model.bath_u[1]
    D = np.zeros((2 * No, 2 * No, len(z)), complex)
    for a in range(No):
        for b in range(No):
            if model.bath_type == "normal" and a != b:
                continue
            for s1 in range(2):
                for s2 in range(2):
                    for h in range(2):
                        e = model.bath_e[h, 0 if model.bath_type == "hybrid" else a]
                        D[a + s1 * No, b + s2 * No] += (W[s1, h, a][None, :] * W[s2, h, b][None, :]
                                                        / (z[:, None] - e[None, :])).sum(1)
    return D


def sigma_matsubara(model: ModelNonsu2, Ntot: int, smap, vec, e0: float, beta: float, Lmats: int):
    """Sigma = G0^-1 - G^-1 as (spin,orbital) matrices (get_Sigma_nonsu2; invg0_hyrege.f90 nonsu2
    branch: G0^-1 = (z+xmu) 1 - impHloc - Delta).  Returns (wm, Sigma[2No,2No,Lmats])."""
    No = model.Norb
    wm = math.pi / beta * (2 * np.arange(1, Lmats + 1) - 1)
    z = 1j * wm
    G = lehmann_G(model, Ntot, smap, vec, e0, z)
    D = delta_matrix(model, z)
    hl = np.zeros((2 * No, 2 * No), complex)
    if model.hloc is not None:
        h = np.asarray(model.hloc, complex)
        for s1 in range(2):
            for s2 in range(2):
                hl[s1 * No:(s1 + 1) * No, s2 * No:(s2 + 1) * No] = h[s1, s2]
    S = np.zeros_like(G)
    for i, zi in enumerate(z):
        S[:, :, i] = (zi + model.xmu) * np.eye(2 * No) - hl - D[:, :, i] - np.linalg.inv(G[:, :, i])
    return wm, S
