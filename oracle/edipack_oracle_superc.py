"""CPU restatement of EDIpack's SUPERC-mode stored Hamiltonian (ED_SPARSE_H=T), numpy.

TEST INFRASTRUCTURE ONLY (see oracle/ed_oracle.h): the checker for the device-side builder of
``edgpu_sector_open_superc``; the product never imports it.

Follows, per state of the sector (normal / hybrid bath, Nspin=1 or 2):
  build_sector (superc)        src/singlesite/ED_SECTOR.f90:244-281  m = iup + idw*2**Ns, idw outer /
                               iup inner, popcnt(iup) - popcnt(idw) = Sz
  c / cdg on the 2*Ns-bit state  src/singlesite/ED_AUX_FUNX.f90:334-384
  ED_SUPERC/stored/Himp.f90    diagonal :11-27, same-spin hops :31-80, anomalous local pairing
                               impHloc_anomalous + pair_field :86-125
  ED_SUPERC/stored/Hint.f90    = ED_NONSU2/stored/Hint.f90 (density-density, Hartree shifts, S-E, P-H)
  ED_SUPERC/stored/Hbath.f90   normal/hybrid diagonal :12-27, bath pairing d :97-133; replica/general:
                               Nambu diagonal :29-45, particle (1,1) / hole (2,2) block hops :48-92,
                               anomalous (1,2) / (2,1) blocks :135-177
  ED_SUPERC/stored/Himp_bath.f90  spin-conserving hybridisation :10-67
  ED_OBSERVABLES_SUPERC.f90    dens/docc :150-165, phisc :204-248 through
                               apply_Cops(v,[1,1],[-1,+1],[a,b],[dw,up]) (ED_SECTOR.f90)
Matrix convention as in the reference: sp_insert_element(spH0,htmp,i,j), row i = the state the
operators act on, column j = the resulting state; duplicates accumulate.

Parity status: pinned to test/src/NORMAL_SUPERC/{evals,dens,docc,phisc}.check and
test/src/HYBRID_SUPERC/{evals,dens,docc}.check, and test/src/{REPLICA,GENERAL}_SUPERC
{evals,dens,docc,phisc}.check (tests/test_oracle_golden_superc.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from edipack_oracle_nonsu2 import _c, _cdg, csr_matvec, to_dense


@dataclass
class ModelSuperc:
    Norb: int = 2
    Nbath: int = 2
    bath_type: str = "normal"          # normal | hybrid | replica | general
    Uloc: tuple = (-2.0, -2.0)
    Ust: float = 0.0
    Jh: float = 0.0
    Jx: float = 0.0
    Jp: float = 0.0
    xmu: float = 0.0
    hfmode: bool = True
    ed_hw_bath: float = 2.0
    deltasc: float = 0.02
    hloc: np.ndarray | None = None       # complex [2, Norb, Norb]  impHloc(s,s,a,b)
    hloc_anomalous: np.ndarray | None = None  # complex [Norb, Norb]  impHloc_anomalous(1,1,a,b)
    pair_field: tuple = ()               # [Norb]
    bath_e: np.ndarray | None = None     # [2, Nfoo, Nbath]
    bath_d: np.ndarray | None = None     # [Nfoo, Nbath]   dmft_bath%d(1,:,:)
    bath_v: np.ndarray | None = None     # [2, Norb, Nbath]
    hbath: np.ndarray | None = None      # replica/general: complex [2(nambu),2,Norb,Norb,Nbath] Hbath_tmp

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
        """init_dmft_bath, ED_BATH_DMFT.f90:211-244 (superc: d = deltasc)."""
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
        self.bath_e = np.broadcast_to(e, (2, self.Nfoo, Nb)).copy()
        self.bath_d = np.full((self.Nfoo, Nb), self.deltasc)
        self.bath_v = np.full((2, self.Norb, Nb), max(0.1, 1.0 / math.sqrt(Nb)))
        return self


def build_sector(Ns: int, Sz: int) -> np.ndarray:
    """List of m = iup + idw*2**Ns with popcnt(iup) - popcnt(idw) = Sz, idw outer / iup inner
    (ED_SECTOR.f90:262-281): ascending in m."""
    one = np.arange(1 << Ns, dtype=np.int64)
    pc = np.array([bin(int(x)).count("1") for x in one])
    out = []
    for idw in one:
        ups = one[pc - pc[idw] == Sz]
        out.append(ups + (idw << Ns))
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def stored_H(model: ModelSuperc, Sz: int):
    """ed_buildH_superc_main (ED_HAMILTONIAN_SUPERC_STORED_HxV.f90:29-260) for one sector:
    returns (map, rowptr, cols [1-based], vals [complex]) in list-of-rows insertion order."""
    if model.bath_e is None:
        model.default_bath()
    Ns, No, Nb = model.Ns, model.Norb, model.Nbath
    smap = build_sector(Ns, Sz)
    index = {int(m): i + 1 for i, m in enumerate(smap)}
    hloc = np.zeros((2, No, No), complex) if model.hloc is None else np.asarray(model.hloc, complex)
    anom = (np.zeros((No, No), complex) if model.hloc_anomalous is None
            else np.asarray(model.hloc_anomalous, complex))
    pf = np.zeros(No)
    pf[: len(model.pair_field)] = model.pair_field
    U = np.asarray(model.Uloc, float)
    offd = 1.0 - np.eye(No)
    Ust, Jh, Jx, Jp = model.Ust * offd, model.Jh * offd, model.Jx * offd, model.Jp * offd
    rows = []
    for m_ in smap:
        m = int(m_)
        row = {}

        def ins(val, j):
            row[j] = row.get(j, 0.0) + val

        ib = [(m >> k) & 1 for k in range(2 * Ns)]
        nup = [float(ib[a]) for a in range(No)]
        ndw = [float(ib[a + Ns]) for a in range(No)]
        i = index[m]

        def chain(ops, amp):
            """ops applied left to right: list of (fn, pos 1-based); inserts amp*signs at (i, j)."""
            k, s = m, 1.0
            for f, pos in ops:
                r = f(pos, k)
                if r is None:
                    return
                k, sg = r
                s *= sg
            ins(amp * s, index[k])

        # ---- Himp.f90
        h = 0.0
        for a in range(No):
            h += hloc[0, a, a] * nup[a] + hloc[1, a, a] * ndw[a] - model.xmu * (nup[a] + ndw[a])
        ins(h, i)
        for a in range(No):
            for b in range(No):
                if hloc[0, a, b] != 0 and ib[b] == 1 and ib[a] == 0:
                    chain([(_c, b + 1), (_cdg, a + 1)], np.conj(hloc[0, a, b]))
                if hloc[1, a, b] != 0 and ib[b + Ns] == 1 and ib[a + Ns] == 0:
                    chain([(_c, b + 1 + Ns), (_cdg, a + 1 + Ns)], np.conj(hloc[1, a, b]))
        if np.any(pf != 0) or np.any(np.abs(anom) != 0):
            for a in range(No):
                for b in range(No):
                    if ib[a] == 1 and ib[b + Ns] == 1:
                        chain([(_c, a + 1), (_c, b + 1 + Ns)], anom[a, b] + (pf[a] if a == b else 0.0))
                    if ib[a] == 0 and ib[b + Ns] == 0:
                        chain([(_cdg, b + 1 + Ns), (_cdg, a + 1)],
                              np.conj(anom[a, b]) + (pf[a] if a == b else 0.0))
        # ---- Hint.f90 (identical to nonsu2)
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
                        chain([(_c, b + 1), (_c, a + 1 + Ns), (_cdg, b + 1 + Ns), (_cdg, a + 1)], Jx[a, b])
        if No > 1 and np.any(Jp != 0):
            for a in range(No):
                for b in range(No):
                    if a != b and ib[b] == 1 and ib[b + Ns] == 1 and ib[a + Ns] == 0 and ib[a] == 0:
                        chain([(_c, b + 1), (_c, b + 1 + Ns), (_cdg, a + 1 + Ns), (_cdg, a + 1)], Jp[a, b])
        if model.bath_type not in ("replica", "general"):
            # ---- Hbath.f90 (normal / hybrid)
            h = 0.0
            for a in range(model.Nfoo):
                for k in range(Nb):
                    s = model.stride(a, k)
                    h += model.bath_e[0, a, k] * ib[s - 1] + model.bath_e[1, a, k] * ib[s - 1 + Ns]
            ins(h, i)
            for a in range(model.Nfoo):
                for k in range(Nb):
                    ms = model.stride(a, k)
                    d = model.bath_d[a, k]
                    if d != 0 and ib[ms - 1] == 1 and ib[ms - 1 + Ns] == 1:
                        chain([(_c, ms), (_c, ms + Ns)], d)
                    if d != 0 and ib[ms - 1] == 0 and ib[ms - 1 + Ns] == 0:
                        chain([(_cdg, ms + Ns), (_cdg, ms)], d)
        else:
            # ---- Hbath.f90 (replica / general), Hbath_tmp in Nambu blocks: (1,1) particles (up),
            # (2,2) holes (dw, enters with -), (1,2)/(2,1) pairing
            hb = np.asarray(model.hbath, complex)
            h = 0.0
            for k in range(Nb):
                for a in range(No):
                    s = model.stride(a, k)
                    h += hb[0, 0, a, a, k] * ib[s - 1] - hb[1, 1, a, a, k] * ib[s - 1 + Ns]
            ins(h, i)
            for k in range(Nb):
                for a in range(No):
                    for b in range(No):
                        ia, ibt = model.stride(a, k), model.stride(b, k)
                        if hb[0, 0, a, b, k] != 0 and ib[ibt - 1] == 1 and ib[ia - 1] == 0:
                            chain([(_c, ibt), (_cdg, ia)], np.conj(hb[0, 0, a, b, k]))
                        ia, ibt = ia + Ns, ibt + Ns
                        if hb[1, 1, a, b, k] != 0 and ib[ibt - 1] == 0 and ib[ia - 1] == 1:
                            chain([(_cdg, ibt), (_c, ia)], np.conj(hb[1, 1, a, b, k]))
            for k in range(Nb):
                for a in range(No):
                    for b in range(No):
                        ia, ibt = model.stride(a, k), model.stride(b, k) + Ns
                        if hb[0, 1, a, b, k] != 0 and ib[ibt - 1] == 0 and ib[ia - 1] == 0:
                            chain([(_cdg, ibt), (_cdg, ia)], np.conj(hb[0, 1, a, b, k]))
                        ia, ibt = model.stride(a, k) + Ns, model.stride(b, k)
                        if hb[1, 0, a, b, k] != 0 and ib[ibt - 1] == 1 and ib[ia - 1] == 1:
                            chain([(_c, ibt), (_c, ia)], np.conj(hb[1, 0, a, b, k]))
        # ---- Himp_bath.f90
        for a in range(No):
            for k in range(Nb):
                ms = model.stride(a, k)
                for sp in range(2):
                    v = model.bath_v[sp, a, k]
                    if v != 0:
                        chain([(_c, a + 1 + sp * Ns), (_cdg, ms + sp * Ns)], np.conj(v))  # imp -> bath
                        chain([(_c, ms + sp * Ns), (_cdg, a + 1 + sp * Ns)], np.conj(v))  # bath -> imp
        rows.append(row)
    rowptr = np.zeros(len(rows) + 1, np.int64)
    cols, vals = [], []
    for r, row in enumerate(rows):
        for j, val in row.items():
            cols.append(j)
            vals.append(val)
        rowptr[r + 1] = len(cols)
    return smap, rowptr, np.array(cols, np.int32), np.array(vals, complex)


def observables(model: ModelSuperc, Sz: int, smap, vec):
    """dens, docc, phisc of one state with weight 1 (ED_OBSERVABLES_SUPERC.f90:150-165, 204-248):
    RePhi(a,b) = ( |(c_{a,dw} + c^+_{b,up})|gs>|^2 - <n_{a,dw}> - (1 - <n_{b,up}>) ) / 2."""
    Ns, No = model.Ns, model.Norb
    w = np.abs(vec) ** 2
    dens, docc = np.zeros(No), np.zeros(No)
    dup, ddw = np.zeros(No), np.zeros(No)
    for a in range(No):
        nu = ((smap >> a) & 1).astype(float)
        nd = ((smap >> (a + Ns)) & 1).astype(float)
        dens[a] = float((w * (nu + nd)).sum())
        docc[a] = float((w * nu * nd).sum())
        dup[a], ddw[a] = float((w * nu).sum()), float((w * nd).sum())
    phi = np.zeros((No, No))
    if Sz < Ns:
        tmap = build_sector(Ns, Sz + 1)
        tindex = {int(m): i for i, m in enumerate(tmap)}
        for a in range(No):
            for b in range(No):
                eta = np.zeros(len(tmap), complex)
                for i, m_ in enumerate(smap):
                    m = int(m_)
                    r = _c(a + 1 + Ns, m)
                    if r is not None:
                        eta[tindex[r[0]]] += r[1] * vec[i]
                    r = _cdg(b + 1, m)
                    if r is not None:
                        eta[tindex[r[0]]] += r[1] * vec[i]
                phi[a, b] = 0.5 * (float(np.vdot(eta, eta).real) - ddw[a] - (1.0 - dup[b]))
    return dens, docc, phi


def twin_sector_order(Ns: int, Sz: int):
    """twin_sector_order for the superc sector A = Sz (ED_SECTOR.f90:1747-1776, flip_state_other
    :1797-1817: up and dw halves exchanged, Sz -> -Sz); vector_B(i) = vec_A(Order(i)).  0-based."""
    smap = build_sector(Ns, Sz)
    lo = (1 << Ns) - 1
    flipped = (smap >> Ns) | ((smap & lo) << Ns)
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


def lehmann_nambu_G(model: ModelSuperc, Sz: int, smap, vec, e0: float, z):
    """Exact Nambu Green's function of one state (weight 1) from the Lehmann representation, spinor
    Psi = (c_{a up}, c^+_{a dw}): blocks G = <<c_up; c^+_up>>, F12 = <<c_up; c_dw>>,
    F21 = <<c^+_dw; c^+_up>>, G22 = <<c^+_dw; c_dw>> (what build_impG_superc / get_impF_superc,
    ED_GF_SUPERC.f90, obtain from Lanczos continued fractions).  Psi^+|0> lives in Sz+1, Psi|0> in
    Sz-1.  Returns M[2No, 2No, len(z)]."""
    Ns, No = model.Ns, model.Norb
    z = np.asarray(z, complex)

    def sector(q):
        tmap, rp, cj, va = stored_H(model, q)
        ev, U = np.linalg.eigh(to_dense(rp, cj, va))
        return tmap, ev - e0, U, {int(x): i for i, x in enumerate(tmap)}

    def amps(sec, fn, bit):
        tmap, _, U, tidx = sec
        seed = np.zeros(len(tmap), complex)
        for i, m_ in enumerate(smap):
            r = fn(bit + 1, int(m_))
            if r is not None:
                seed[tidx[r[0]]] += r[1] * vec[i]
        return U.conj().T @ seed

    sp, sm = sector(Sz + 1), sector(Sz - 1)
    # Psi^+_b |0> : c^+_{b up}, c_{b dw}   (sector Sz+1);   Psi_a |0> : c_{a up}, c^+_{a dw}   (Sz-1)
    A = np.array([amps(sp, _cdg, a) for a in range(No)] + [amps(sp, _c, a + Ns) for a in range(No)])
    B = np.array([amps(sm, _c, a) for a in range(No)] + [amps(sm, _cdg, a + Ns) for a in range(No)])
    M = np.zeros((2 * No, 2 * No, len(z)), complex)
    for i, zi in enumerate(z):
        M[:, :, i] = (A.conj() / (zi - sp[1])[None, :]) @ A.T + (B / (zi + sm[1])[None, :]) @ B.conj().T
    return M


def bath_nambu_functions(model: ModelSuperc, z):
    """Delta and Fdelta on the Matsubara axis for ed_mode=superc: normal / hybrid
    (delta_normal.f90:44-60, fdelta_normal.f90; delta_hybrid.f90:50-72, fdelta_hybrid.f90) and
    replica / general (delta_replica.f90:40-52, fdelta_replica.f90: V_k = sigma_z (x) diag(v),
    blocks (1,1) and (1,2) of V_k (z - H_k)^-1 V_k).  Returns (Delta, Fdelta) [No, No, len(z)]."""
    No, Nb = model.Norb, model.Nbath
    z = np.asarray(z, complex)
    D = np.zeros((No, No, len(z)), complex)
    Fd = np.zeros((No, No, len(z)), complex)
    w2 = z.imag ** 2
    if model.bath_type in ("normal", "hybrid"):
        for a in range(No):
            for b in range(No):
                if model.bath_type == "normal" and a != b:
                    continue
                f = 0 if model.bath_type == "hybrid" else a
                e, d = model.bath_e[0, f], model.bath_d[f]
                den = w2[:, None] + e[None, :] ** 2 + d[None, :] ** 2
                vv = model.bath_v[0, a][None, :] * model.bath_v[0, b][None, :]
                D[a, b] = -(vv * (z[:, None] + e[None, :]) / den).sum(1)
                Fd[a, b] = (d[None, :] * vv / den).sum(1)
    else:
        hb = np.asarray(model.hbath, complex)
        for k in range(Nb):
            Hk = hb[..., k].transpose(0, 2, 1, 3).reshape(2 * No, 2 * No)
            Vk = np.kron(np.diag([1.0, -1.0]), np.diag(model.bath_v[0, :, k]))
            for i, zi in enumerate(z):
                X = Vk @ np.linalg.inv(zi * np.eye(2 * No) - Hk) @ Vk
                D[:, :, i] += X[:No, :No]
                Fd[:, :, i] += X[:No, No:]
    return D, Fd


def sigma_self_matsubara(model: ModelSuperc, Sz: int, smap, vec, e0: float, beta: float, Lmats: int):
    """get_Sigma_superc / get_Self_superc (ED_GF_SUPERC.f90:938-1102) on the Matsubara axis:
    Sigma = G0^-1 - [M^-1]_(1,1), Self = F0^-1 - [M^-1]_(1,2) with M the Nambu Green's function,
    G0^-1 = (z+xmu) 1 - impHloc - Delta, F0^-1 = -impHloc_anomalous - Fdelta (invg0_*.f90, invf0_*.f90).
    Returns (wm, Sigma[No,No,L], Self[No,No,L])."""
    No = model.Norb
    wm = math.pi / beta * (2 * np.arange(1, Lmats + 1) - 1)
    z = 1j * wm
    M = lehmann_nambu_G(model, Sz, smap, vec, e0, z)
    D, Fd = bath_nambu_functions(model, z)
    hl = np.zeros((No, No), complex) if model.hloc is None else np.asarray(model.hloc[0], complex)
    an = (np.zeros((No, No), complex) if model.hloc_anomalous is None
          else np.asarray(model.hloc_anomalous, complex))
    Sig = np.zeros((No, No, Lmats), complex)
    Slf = np.zeros((No, No, Lmats), complex)
    for i, zi in enumerate(z):
        inv = np.linalg.inv(M[:, :, i])
        Sig[:, :, i] = (zi + model.xmu) * np.eye(No) - hl - D[:, :, i] - inv[:No, :No]
        Slf[:, :, i] = -an - Fd[:, :, i] - inv[:No, No:]
    return wm, Sig, Slf
