"""Host-side mirror of the reference's interface for the NORMAL-mode hot path, on top of the
C ABI (``include/edgpu.h``).  Names and argument meaning follow the reference so that the
parity tests read like the reference's own call sequence:

=============================  ==========================================================
reference                      here
=============================  ==========================================================
ed_init_solver / set_umatrix   :class:`EDModel` (plain numbers -> ``edgpu_normal_params``)
build_Hv_sector_normal         :func:`build_Hv_sector_normal`   ED_HAMILTONIAN_NORMAL.f90:31
delete_Hv_sector_normal        :func:`delete_Hv_sector_normal`  :212
vecDim_Hv_sector_normal        :func:`vecDim_Hv_sector_normal`  :286
spHtimesV_p(Nloc,v,Hv)         :func:`spHtimesV_p`              ED_VARS_GLOBAL.f90:111-120,196
sp_lanc_eigh                   :func:`sp_lanc_eigh`             call site ED_DIAG_NORMAL.f90:206
sp_lanc_tridiag                :func:`sp_lanc_tridiag`          ED_HAMILTONIAN_NORMAL.f90:360
sp_eigh (ARPACK)               :func:`sp_eigh`                  call site ED_DIAG_NORMAL.f90:179
build_Hv_sector_nonsu2         :func:`build_Hv_sector_nonsu2`   ED_HAMILTONIAN_NONSU2.f90:31
build_Hv_sector_superc         :func:`build_Hv_sector_superc`   ED_HAMILTONIAN_SUPERC.f90:31
tridiag_Hv_sector_normal       :func:`tridiag_Hv_sector_normal` :321
ed_diag_d                      :func:`ed_diag_d`                ED_DIAG_NORMAL.f90:76
lanc_build_gf_normal_diag      :func:`lanc_build_gf_normal_diag` ED_GF_NORMAL.f90:131
observables_normal             :func:`observables_normal`       ED_OBSERVABLES_NORMAL.f90:78
=============================  ==========================================================

Everything heavy runs in hand-written CUDA behind the ABI; this module only sequences calls.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _abi
from ._abi import (EdgpuError, MAXBATH, MAXORB, NormalParams, Nonsu2Params, SupercParams, check,
                   ptr)

BATH_CODES = {"normal": 0, "hybrid": 1, "replica": 2, "general": 3}


def binomial(n: int, k: int) -> int:
    return math.comb(n, k) if 0 <= k <= n else 0



# ------------------------------------------------------------------------------------------
# two-body operators: ED_PARSE_UMATRIX.f90 (read_umatrix_file :353-449, parse_umatrix_line
# :452-634, set_umatrix :88-165)
# ------------------------------------------------------------------------------------------
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


@dataclass
class EDModel:
    """What ``ed_read_input`` + ``ed_init_solver`` + ``ed_set_Hloc`` + ``set_umatrix`` leave in
    the reference's module globals for the NORMAL-mode H x v (SURVEY 8b "Semantics")."""

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
    hloc: np.ndarray | None = None     # [2, Norb, Norb]   impHloc(s,s,a,b)
    bath_e: np.ndarray | None = None   # [2, Nfoo, Nbath]  dmft_bath%e
    bath_v: np.ndarray | None = None   # [2, Norb, Nbath]  dmft_bath%v
    spin_field_z: tuple = ()
    exc_field: tuple = (0.0, 0.0, 0.0, 0.0)
    hbath: np.ndarray | None = None    # replica/general Hbath_tmp [2, Norb, Norb, Nbath]
    lanc_niter: int = 512              # LANC_NITER      (ED_INPUT_VARS.f90:723-732)
    lanc_ngfiter: int = 200            # LANC_NGFITER
    lanc_tolerance: float = 1e-12      # LANC_TOLERANCE (reference default 1e-18)
    lanc_dim_threshold: int = 1024     # LANC_DIM_THRESHOLD
    gs_threshold: float = 1e-9         # GS_THRESHOLD
    lanc_method: str = "arpack"        # LANC_METHOD: "arpack" (default) | "lanczos"
    lanc_ncv_factor: int = 10          # LANC_NCV_FACTOR
    lanc_ncv_add: int = 0              # LANC_NCV_ADD
    ed_finite_temp: bool = False       # ED_FINITE_TEMP
    lanc_nstates_sector: int = 2       # LANC_NSTATES_SECTOR
    lanc_nstates_total: int = 2        # LANC_NSTATES_TOTAL
    cutoff: float = 1e-9               # CUTOFF (spectrum cut-off exp(-beta(E-Egs)) at finite T)
    ed_twin: bool = False              # ED_TWIN: solve one sector of each (nup,ndw) / (ndw,nup) pair
    ed_use_kanamori: bool = True       # ED_USE_KANAMORI
    umatrix_lines: tuple = ()          # ED_READ_UMATRIX file lines + ed_add_twobody_operator calls
    _params: NormalParams | None = field(default=None, repr=False)

    def umatrix(self):
        """set_umatrix (ED_PARSE_UMATRIX.f90:88-165) for this model."""
        return parse_umatrix(self.Norb, self.umatrix_lines, self.ed_use_kanamori, self.Uloc, self.Ust,
                             self.Jh, self.Jx, self.Jp)

    @property
    def coulomb_sundry(self):
        return self.umatrix()["sundry"]

    def add_twobody_operator(self, oi, si, oj, sj, ok, sk, ol, sl, Uijkl):
        """ed_add_twobody_operator (ED_PARSE_UMATRIX.f90:44-86)."""
        self.umatrix_lines = tuple(self.umatrix_lines) + ((oi, si, oj, sj, ok, sk, ol, sl, Uijkl),)
        self._params = None

    @property
    def Ns(self) -> int:  # ED_SETUP.f90:118-126
        return self.Nbath + self.Norb if self.bath_type == "hybrid" else (self.Nbath + 1) * self.Norb

    @property
    def Nfoo(self) -> int:
        return 1 if self.bath_type == "hybrid" else self.Norb

    def getBathStride(self, a: int, k: int) -> int:
        """ED_SETUP.f90:605-622 (a, k 0-based; returns the 1-based site)."""
        if self.bath_type == "normal":
            return self.Norb + a * self.Nbath + k + 1
        if self.bath_type == "hybrid":
            return self.Norb + k + 1
        return (a + 1) + (k + 1) * self.Norb

    def init_dmft_bath(self):
        """Default bath of ``init_dmft_bath`` (ED_BATH_DMFT.f90:211-244), normal/hybrid."""
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
        self.bath_v = np.full((2, self.Norb, Nb), max(0.1, 1.0 / math.sqrt(Nb)))
        self._params = None
        return self

    def params(self) -> NormalParams:
        if self._params is not None:
            return self._params
        if self.bath_e is None:
            self.init_dmft_bath()
        No, Nb = self.Norb, self.Nbath
        if No > MAXORB or Nb > MAXBATH or self.Ns > 31:
            raise EdgpuError("model too large for edgpu_normal_params")
        p = NormalParams()
        p.Ns, p.Norb, p.Nbath = self.Ns, No, Nb
        p.bath_type = BATH_CODES[self.bath_type]
        p.hfmode, p.Nfoo, p.xmu = int(self.hfmode), self.Nfoo, self.xmu
        um = self.umatrix()
        eloc = np.zeros((2, MAXORB, MAXORB))
        if self.hloc is not None:
            eloc[:, :No, :No] = self.hloc
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
        hb = np.zeros((2, MAXORB, MAXORB, MAXBATH))
        if self.hbath is not None:
            hb[:, :No, :No, :Nb] = self.hbath
            for s in range(2):
                for a in range(No):
                    bd[s, a, :Nb] = self.hbath[s, a, a, :]
        p.diag_hybr[:] = dh.ravel().tolist()
        p.bath_diag[:] = bd.ravel().tolist()
        p.hbath[:] = hb.ravel().tolist()
        st = np.zeros((MAXORB, MAXBATH), np.int32)
        for a in range(No):
            for k in range(Nb):
                st[a, k] = self.getBathStride(a, k)
        p.stride[:] = st.ravel().tolist()
        self._params = p
        return p


# ------------------------------------------------------------------------------------------
# engine / communicator
# ------------------------------------------------------------------------------------------
def ed_init(device: int = 0):
    check(_abi.load().edgpu_init(device))


def ed_finalize():
    check(_abi.load().edgpu_finalize())


def ed_set_comm(rank: int, nranks: int, uid: bytes | None):
    """ed_set_MpiComm (ED_VARS_GLOBAL.f90:341): uid from :func:`comm_unique_id` on rank 0."""
    buf = C.create_string_buffer(uid, _abi.UID_BYTES) if uid is not None else None
    check(_abi.load().edgpu_comm_init(rank, nranks, buf))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(_abi.UID_BYTES)
    check(_abi.load().edgpu_comm_unique_id(buf))
    return buf.raw


# ------------------------------------------------------------------------------------------
# dw-split bookkeeping and the rank <-> root vector movers (host side; any torch.distributed
# backend: NCCL on the GPUs, gloo in the CPU tests)
# ------------------------------------------------------------------------------------------
def mpi_split(n: int, nranks: int, rank: int):
    """Block split with the first (n mod P) ranks one element longer: the dw-column split of
    build_Hv_sector_normal (ED_HAMILTONIAN_NORMAL.f90:128-142) and the row split of
    vector_transpose_MPI (ED_HAMILTONIAN_NORMAL_COMMON.f90:104-112).  Returns (count, start)."""
    base, rem = divmod(n, nranks)
    return base + (1 if rank < rem else 0), rank * base + min(rank, rem)


def chunk_bounds(DimUp: int, DimDw: int, nranks: int, rank: int):
    """[istart, iend) of this rank's chunk of the sector vector, i = iup + idw*DimUp:
    mpiIshift / mpiQ of ED_HAMILTONIAN_NORMAL.f90:136-142."""
    q, d0 = mpi_split(DimDw, nranks, rank)
    return d0 * DimUp, (d0 + q) * DimUp


def _dev_of(group):
    import torch
    import torch.distributed as dist

    return (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl"
            else torch.device("cpu"))


def scatter_vector_MPI(v_full, DimUp: int, DimDw: int, root: int = 0, group=None, DimPh: int = 1):
    """d_scatter_vector_MPI (ED_AUX_FUNX.f90:598-652): root's full vector -> every rank's chunk.
    With phonons the full vector is DimPh electronic slices of DimUp*DimDw and every rank's chunk
    is DimPh slices of its dw columns (one MPI_Scatterv per slice, :637-649)."""
    import torch
    import torch.distributed as dist

    P, r = dist.get_world_size(group), dist.get_rank(group)
    bounds = [chunk_bounds(DimUp, DimDw, P, k) for k in range(P)]
    nmax = max(hi - lo for lo, hi in bounds)  # collectives want equal pieces: pad to the longest
    lo, hi = bounds[r]
    dev = _dev_of(group)
    nel = DimUp * DimDw
    t = torch.as_tensor(np.ascontiguousarray(v_full, np.float64)) if r == root else None
    if r == root and t.numel() != nel * DimPh:
        raise EdgpuError("scatter_vector_MPI error: size(V) != Mpi_Allreduce(Nloc)")
    res = []
    for iph in range(DimPh):
        out = torch.empty(nmax, dtype=torch.float64, device=dev)
        pieces = None
        if r == root:
            pieces = []
            for a, b in bounds:
                x = torch.zeros(nmax, dtype=torch.float64)
                x[: b - a] = t[iph * nel + a: iph * nel + b]
                pieces.append(x.to(dev))
        dist.scatter(out, pieces, src=root, group=group)
        res.append(out[: hi - lo].cpu().numpy())
    return np.concatenate(res)


def allgather_vector_MPI(chunk, DimUp: int, DimDw: int, group=None, DimPh: int = 1):
    """d_allgather_vector_MPI (ED_AUX_FUNX.f90:840-890): every rank's chunk -> full vector, slice
    by slice when DimPh > 1."""
    import torch
    import torch.distributed as dist

    P = dist.get_world_size(group)
    dev = _dev_of(group)
    mine = torch.as_tensor(np.ascontiguousarray(chunk, np.float64)).to(dev)
    sizes = [hi - lo for lo, hi in (chunk_bounds(DimUp, DimDw, P, k) for k in range(P))]
    nmax = max(sizes)
    nloc = sizes[dist.get_rank(group)]
    if mine.numel() != nloc * DimPh:
        raise EdgpuError("allgather_vector_MPI error: chunk length is not DimPh * mpiQ")
    res = []
    for iph in range(DimPh):
        pad = torch.zeros(nmax, dtype=torch.float64, device=dev)
        pad[:nloc] = mine[iph * nloc:(iph + 1) * nloc]
        outs = [torch.empty(nmax, dtype=torch.float64, device=dev) for _ in range(P)]
        dist.all_gather(outs, pad, group=group)
        res.append(torch.cat([o[:n] for o, n in zip(outs, sizes)]))
    return torch.cat(res).cpu().numpy()


def gather_vector_MPI(chunk, DimUp: int, DimDw: int, root: int = 0, group=None, DimPh: int = 1):
    """d_gather_vector_MPI (ED_AUX_FUNX.f90:742-795); the full vector is returned on root only."""
    import torch.distributed as dist

    full = allgather_vector_MPI(chunk, DimUp, DimDw, group, DimPh)
    return full if dist.get_rank(group) == root else None


# ------------------------------------------------------------------------------------------
# sector + H x v
# ------------------------------------------------------------------------------------------
def build_Hv_sector_normal(model: EDModel, nup: int, ndw: int):
    check(_abi.load().edgpu_sector_open_normal(C.byref(model.params()), nup, ndw))


def build_Hv_sector_normal_orbs(model: "EDModel", nups, ndws):
    """build_Hv_sector_normal with ed_total_ud = F: the orbital-resolved sector (Nups(1:Norb),
    Ndws(1:Norb)), map and Hamiltonian generated on the device (directMatVec_normal_orbs /
    ed_buildh_normal_orbs); close it with delete_Hv_sector_csr."""
    global _open_is_complex
    nu = np.ascontiguousarray(nups, np.int32)
    nd = np.ascontiguousarray(ndws, np.int32)
    if nu.size != model.Norb or nd.size != model.Norb:
        raise EdgpuError("nups / ndws must have Norb entries")
    check(_abi.load().edgpu_sector_open_normal_orbs(C.byref(model.params()), ptr(nu), ptr(nd)))
    _open_is_complex = False


def set_coulomb_sundry(terms=()):
    """The module global ``coulomb_sundry`` (user two-body terms of a umatrix file, applied by
    direct/HxV_sundry.f90): terms = iterable of (cd_i, cd_j, c_k, c_l, U), every operator an
    (orbital 1-based, spin 1|2) pair.  Applies to NORMAL sectors opened afterwards; () clears."""
    terms = list(terms)
    arr = (_abi.SundryTerm * max(len(terms), 1))()
    for t, (ci, cj, ck, cl, U) in enumerate(terms):
        arr[t].cd_i[:] = ci
        arr[t].cd_j[:] = cj
        arr[t].c_k[:] = ck
        arr[t].c_l[:] = cl
        arr[t].U = U
    check(_abi.load().edgpu_set_coulomb_sundry(len(terms), C.cast(arr, C.c_void_p)))


def set_umatrix(model: "EDModel"):
    """The engine-side half of set_umatrix (ED_PARSE_UMATRIX.f90:88-165, called from
    ed_init_solver): hands the model's coulomb_sundry list to the engine (the Kanamori matrices and
    mfHloc travel in edgpu_normal_params)."""
    set_coulomb_sundry(model.coulomb_sundry)


def set_phonons(Nph: int = 0, w0: float = 0.0, g=None, A: float = 0.0):
    """Nph, w0_ph, g_ph(Norb,Norb), A_ph (ED_INPUT_VARS.f90:184-198): NORMAL sectors opened
    afterwards carry DimPh = Nph+1 phonon slices (direct/HxV_ph.f90, HxV_eph.f90)."""
    g = np.ascontiguousarray(np.atleast_2d(np.zeros((1, 1)) if g is None else np.asarray(g, float)))
    check(_abi.load().edgpu_set_phonons(int(Nph), float(w0), float(A), ptr(g), g.shape[0]))


def delete_Hv_sector_normal():
    global _open_is_complex
    _open_is_complex = False
    check(_abi.load().edgpu_sector_close())


def vecDim_Hv_sector_normal() -> int:
    return int(_abi.load().edgpu_sector_vecdim())


def host_register(a: np.ndarray):
    """Page-locks a caller-owned host array (``edgpu_host_register``) for full-speed copies in
    ``spHtimesV_p``; call :func:`host_unregister` before it is freed."""
    check(_abi.load().edgpu_host_register(ptr(a), a.nbytes))


def host_unregister(a: np.ndarray):
    check(_abi.load().edgpu_host_unregister(ptr(a)))


def sector_comm_info():
    """(mode, halo columns received, columns sent, chunks) of the open NORMAL sector's Hdw exchange:
    mode 0 single rank, 1 halo, 2 peer-memory transposes, 3 NCCL transposes."""
    L = _abi.load()
    mode, nch = C.c_int(), C.c_int()
    a, b = C.c_int64(), C.c_int64()
    check(L.edgpu_sector_comm_info(C.byref(mode), C.byref(a), C.byref(b), C.byref(nch)))
    return mode.value, a.value, b.value, nch.value


def sector_dims():
    L = _abi.load()
    a, b, c, d = (C.c_int64() for _ in range(4))
    check(L.edgpu_sector_dims(C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
    return a.value, b.value, c.value, d.value


def sector_map(spin: int) -> np.ndarray:
    DimUp, DimDw, _, _ = sector_dims()
    m = np.empty(DimUp if spin == 0 else DimDw, np.int32)
    check(_abi.load().edgpu_sector_get_map(spin, ptr(m)))
    return m


def sector_hops(spin: int):
    """Hop table as CSR (rowptr, 1-based targets, signed values), one row per source state."""
    L = _abi.load()
    DimUp, DimDw, _, _ = sector_dims()
    dim = DimUp if spin == 0 else DimDw
    n = int(L.edgpu_sector_hop_count(spin))
    if n < 0:
        check(1)
    rp = np.zeros(dim + 1, np.int64)
    tg = np.zeros(max(n, 1), np.int32)
    va = np.zeros(max(n, 1), np.float64)
    check(L.edgpu_sector_get_hops(spin, ptr(rp), ptr(tg), ptr(va)))
    return rp, tg[:n], va[:n]


def spHtimesV_p(v: np.ndarray) -> np.ndarray:
    """``call spHtimesV_p(Nloc,v,Hv)`` with host arrays (the drop-in procedure pointer)."""
    L = _abi.load()
    v = np.ascontiguousarray(v, np.float64)
    hv = np.empty_like(v)
    n = C.c_int32(v.size)
    L.edgpu_hxv_d(C.byref(n), ptr(v), ptr(hv))
    check(L.edgpu_status())
    return hv


# ------------------------------------------------------------------------------------------
# stored-H sectors (ED_SPARSE_H=T): sparse_matrix_csr -> device CSR
# ------------------------------------------------------------------------------------------
_open_is_complex = False


def build_Hv_sector_csr(rowptr, cols, vals, nglobal=None, row_offset: int = 0):
    """Hands the host-built stored Hamiltonian (the reference's spH0, ED_SPARSE_MATRIX.f90:16-41,
    filled by ed_buildh_*_main) to the device: rowptr (0-based offsets, nloc+1), cols (1-based
    global columns, any order, duplicates add up), vals real or complex.  Afterwards
    spHtimesV_p / spHtimesV_cc and the sp_lanc_* drivers act on this matrix."""
    global _open_is_complex
    L = _abi.load()
    rowptr = np.ascontiguousarray(rowptr, np.int64)
    cols = np.ascontiguousarray(cols, np.int32)
    nloc = rowptr.size - 1
    nglobal = nloc if nglobal is None else int(nglobal)
    if np.iscomplexobj(vals):
        v = np.ascontiguousarray(vals, np.complex128)
        check(L.edgpu_csr_open_z(nloc, nglobal, row_offset, ptr(rowptr), ptr(cols), ptr(v)))
        _open_is_complex = True
    else:
        v = np.ascontiguousarray(vals, np.float64)
        check(L.edgpu_csr_open_d(nloc, nglobal, row_offset, ptr(rowptr), ptr(cols), ptr(v)))
        _open_is_complex = False


def delete_Hv_sector_csr():
    global _open_is_complex
    _open_is_complex = False
    check(_abi.load().edgpu_sector_close())


def spHtimesV_cc(v: np.ndarray) -> np.ndarray:
    """``call spHtimesV_cc(Nloc,v,Hv)`` (cc_sparse_HxV, ED_VARS_GLOBAL.f90:121-132), complex(8)."""
    L = _abi.load()
    v = np.ascontiguousarray(v, np.complex128)
    hv = np.empty_like(v)
    n = C.c_int32(v.size)
    L.edgpu_hxv_z(C.byref(n), ptr(v), ptr(hv))
    check(L.edgpu_status())
    return hv


def _packed_stride(bath_type: str, No: int, Nb: int, a: int, k: int) -> int:
    """getBathStride(a+1,k+1) (ED_SETUP.f90:605-622), 1-based site."""
    if bath_type == "hybrid":
        return No + k + 1
    if bath_type in ("replica", "general"):
        return (a + 1) + (k + 1) * No
    return No + a * Nb + k + 1


def _push_hbath(model):
    """Hbath_tmp of a replica / general bath -> engine (edgpu_set_hbath_packed)."""
    if model.bath_type not in ("replica", "general"):
        return
    if model.hbath is None:
        raise EdgpuError("replica / general bath: the model has no hbath (Hbath_tmp)")
    hb = np.ascontiguousarray(np.asarray(model.hbath, np.complex128))
    if hb.shape != (2, 2, model.Norb, model.Norb, model.Nbath):
        raise EdgpuError("hbath must be [2, 2, Norb, Norb, Nbath]")
    check(_abi.load().edgpu_set_hbath_packed(ptr(hb), model.Norb, model.Nbath))


@dataclass
class EDModelNonsu2:
    """Module globals read by ``ed_buildH_nonsu2_main`` (ed_mode=nonsu2, normal / hybrid bath)."""

    Norb: int = 2
    Nbath: int = 4
    bath_type: str = "hybrid"
    Uloc: tuple = (1.0, 1.0)
    Ust: float = 0.0
    Jh: float = 0.0
    Jx: float = 0.0
    Jp: float = 0.0
    xmu: float = 0.0
    hfmode: bool = True
    ed_hw_bath: float = 2.0
    hloc: np.ndarray | None = None        # complex [2, 2, Norb, Norb]  impHloc(is, js, a, b)
    bath_e: np.ndarray | None = None      # [2, Nfoo, Nbath]
    bath_v: np.ndarray | None = None      # [2, Norb, Nbath]
    bath_u: np.ndarray | None = None      # [2, Norb, Nbath]
    spin_field: np.ndarray | None = None  # [Norb, 3]
    hbath: np.ndarray | None = None       # replica / general: complex [2, 2, Norb, Norb, Nbath] Hbath_tmp

    @property
    def Ns(self) -> int:
        return self.Nbath + self.Norb if self.bath_type == "hybrid" else (self.Nbath + 1) * self.Norb

    @property
    def Nfoo(self) -> int:
        return 1 if self.bath_type == "hybrid" else self.Norb

    def init_dmft_bath(self):
        """init_dmft_bath for nonsu2 (ED_BATH_DMFT.f90:211-244): e and v as NORMAL mode, u = v."""
        tmp = EDModel(Norb=self.Norb, Nbath=self.Nbath, bath_type=self.bath_type,
                      ed_hw_bath=self.ed_hw_bath).init_dmft_bath()
        self.bath_e, self.bath_v, self.bath_u = tmp.bath_e, tmp.bath_v, tmp.bath_v.copy()
        return self

    def params(self) -> Nonsu2Params:
        if self.bath_e is None:
            self.init_dmft_bath()
        No, Nb = self.Norb, self.Nbath
        if No > MAXORB or Nb > MAXBATH or 2 * self.Ns > 31:
            raise EdgpuError("model too large for edgpu_nonsu2_params")
        p = Nonsu2Params()
        p.Ns, p.Norb, p.Nbath = self.Ns, No, Nb
        p.bath_type = BATH_CODES[self.bath_type]
        p.hfmode, p.Nfoo, p.xmu = int(self.hfmode), self.Nfoo, self.xmu
        hl = np.zeros((2, 2, MAXORB, MAXORB, 2))
        if self.hloc is not None:
            h = np.asarray(self.hloc, complex)
            hl[:, :, :No, :No, 0], hl[:, :, :No, :No, 1] = h.real, h.imag
        p.hloc[:] = hl.ravel().tolist()
        sf = np.zeros((MAXORB, 3))
        if self.spin_field is not None:
            sf[:No] = self.spin_field
        p.spin_field[:] = sf.ravel().tolist()
        U = np.zeros(MAXORB)
        U[:No] = np.asarray(self.Uloc, float)[:No]
        p.Uloc[:] = U.tolist()
        off = np.zeros((MAXORB, MAXORB))
        off[:No, :No] = 1.0 - np.eye(No)
        for name, val in (("Ust", self.Ust), ("Jh", self.Jh), ("Jx", self.Jx), ("Jp", self.Jp)):
            getattr(p, name)[:] = (val * off).ravel().tolist()
        for name, arr, n0 in (("bath_e", self.bath_e, self.Nfoo), ("bath_v", self.bath_v, No),
                              ("bath_u", self.bath_u, No)):
            buf = np.zeros((2, MAXORB, MAXBATH))
            buf[:, :n0, :Nb] = arr
            getattr(p, name)[:] = buf.ravel().tolist()
        st = np.zeros((MAXORB, MAXBATH), np.int32)
        for a in range(No):
            for k in range(Nb):
                st[a, k] = _packed_stride(self.bath_type, No, Nb, a, k)
        p.stride[:] = st.ravel().tolist()
        return p


@dataclass
class EDModelSuperc:
    """Module globals read by ``ed_buildH_superc_main`` (ed_mode=superc, normal / hybrid bath)."""

    Norb: int = 2
    Nbath: int = 2
    bath_type: str = "normal"
    Uloc: tuple = (-2.0, -2.0)
    Ust: float = 0.0
    Jh: float = 0.0
    Jx: float = 0.0
    Jp: float = 0.0
    xmu: float = 0.0
    hfmode: bool = True
    ed_hw_bath: float = 2.0
    deltasc: float = 0.02
    hloc: np.ndarray | None = None            # complex [2, Norb, Norb]  impHloc(s, s, a, b)
    hloc_anomalous: np.ndarray | None = None  # complex [Norb, Norb]
    pair_field: tuple = ()
    bath_e: np.ndarray | None = None          # [2, Nfoo, Nbath]
    bath_d: np.ndarray | None = None          # [Nfoo, Nbath]
    bath_v: np.ndarray | None = None          # [2, Norb, Nbath]
    hbath: np.ndarray | None = None           # replica / general: complex [2(nambu), 2, Norb, Norb, Nbath]

    @property
    def Ns(self) -> int:
        return self.Nbath + self.Norb if self.bath_type == "hybrid" else (self.Nbath + 1) * self.Norb

    @property
    def Nfoo(self) -> int:
        return 1 if self.bath_type == "hybrid" else self.Norb

    def init_dmft_bath(self):
        """init_dmft_bath for superc (ED_BATH_DMFT.f90:211-244): e and v as NORMAL mode, d = deltasc."""
        tmp = EDModel(Norb=self.Norb, Nbath=self.Nbath, bath_type=self.bath_type,
                      ed_hw_bath=self.ed_hw_bath).init_dmft_bath()
        self.bath_e, self.bath_v = tmp.bath_e, tmp.bath_v
        self.bath_d = np.full((self.Nfoo, self.Nbath), self.deltasc)
        return self

    def params(self) -> SupercParams:
        if self.bath_e is None:
            self.init_dmft_bath()
        No, Nb = self.Norb, self.Nbath
        if No > MAXORB or Nb > MAXBATH or 2 * self.Ns > 31:
            raise EdgpuError("model too large for edgpu_superc_params")
        p = SupercParams()
        p.Ns, p.Norb, p.Nbath = self.Ns, No, Nb
        p.bath_type = BATH_CODES[self.bath_type]
        p.hfmode, p.Nfoo, p.xmu = int(self.hfmode), self.Nfoo, self.xmu
        hl = np.zeros((2, MAXORB, MAXORB, 2))
        if self.hloc is not None:
            h = np.asarray(self.hloc, complex)
            hl[:, :No, :No, 0], hl[:, :No, :No, 1] = h.real, h.imag
        p.hloc[:] = hl.ravel().tolist()
        an = np.zeros((MAXORB, MAXORB, 2))
        if self.hloc_anomalous is not None:
            h = np.asarray(self.hloc_anomalous, complex)
            an[:No, :No, 0], an[:No, :No, 1] = h.real, h.imag
        p.hloc_anomalous[:] = an.ravel().tolist()
        pf = np.zeros(MAXORB)
        pf[: len(self.pair_field)] = self.pair_field
        p.pair_field[:] = pf.tolist()
        U = np.zeros(MAXORB)
        U[:No] = np.asarray(self.Uloc, float)[:No]
        p.Uloc[:] = U.tolist()
        off = np.zeros((MAXORB, MAXORB))
        off[:No, :No] = 1.0 - np.eye(No)
        for name, val in (("Ust", self.Ust), ("Jh", self.Jh), ("Jx", self.Jx), ("Jp", self.Jp)):
            getattr(p, name)[:] = (val * off).ravel().tolist()
        for name, arr, n0 in (("bath_e", self.bath_e, self.Nfoo), ("bath_v", self.bath_v, No)):
            buf = np.zeros((2, MAXORB, MAXBATH))
            buf[:, :n0, :Nb] = arr
            getattr(p, name)[:] = buf.ravel().tolist()
        bd = np.zeros((MAXORB, MAXBATH))
        bd[: self.Nfoo, :Nb] = self.bath_d
        p.bath_d[:] = bd.ravel().tolist()
        st = np.zeros((MAXORB, MAXBATH), np.int32)
        for a in range(No):
            for k in range(Nb):
                st[a, k] = _packed_stride(self.bath_type, No, Nb, a, k)
        p.stride[:] = st.ravel().tolist()
        return p


def build_Hv_sector_superc(model: EDModelSuperc, sz: int):
    """build_Hv_sector_superc + ed_buildH_superc_main on the device for the sector Sz = Nup - Ndw."""
    global _open_is_complex
    _push_hbath(model)
    check(_abi.load().edgpu_sector_open_superc(C.byref(model.params()), sz))
    _open_is_complex = True


def delete_Hv_sector_superc():
    delete_Hv_sector_csr()


def set_sparse_H(flag: bool):
    """``ED_SPARSE_H`` for the packed-state modes: True (default) stores spH0 on the device, False
    applies the matrix elements on the fly (directMatVec_nonsu2_main / _superc_main)."""
    check(_abi.load().edgpu_set_sparse_h(int(bool(flag))))


def build_Hv_sector_nonsu2(model: EDModelNonsu2, ntot: int):
    """build_Hv_sector_nonsu2 + ed_buildH_nonsu2_main on the device (sector map and complex spH0
    generated by kernels); afterwards spHtimesV_cc / sp_lanc_* / sp_eigh act on it."""
    global _open_is_complex
    _push_hbath(model)
    check(_abi.load().edgpu_sector_open_nonsu2(C.byref(model.params()), ntot))
    _open_is_complex = True


def delete_Hv_sector_nonsu2():
    delete_Hv_sector_csr()


def sector_map_nonsu2() -> np.ndarray:
    """H(1)%map(DimEl) of the open device-built nonsu2 sector (packed states iup + idw*2^Ns)."""
    L = _abi.load()
    m = np.empty(int(L.edgpu_sector_dim()), np.int32)
    check(L.edgpu_sector_get_map(0, ptr(m)))
    return m


def stored_csr():
    """Downloads the device CSR of the open stored-H sector: (rowptr, cols 1-based, vals)."""
    L = _abi.load()
    nnz = int(L.edgpu_csr_nnz())
    if nnz < 0:
        raise EdgpuError("no stored-H sector open")
    nloc = vecDim_Hv_sector_normal()
    rp = np.zeros(nloc + 1, np.int64)
    cj = np.zeros(max(nnz, 1), np.int32)
    va = np.zeros(max(nnz, 1), np.complex128 if _open_is_complex else np.float64)
    check(L.edgpu_csr_get(ptr(rp), ptr(cj), ptr(va)))
    return rp, cj[:nnz], va[:nnz]


def set_kernel_variant(variant: int):
    check(_abi.load().edgpu_set_kernel_variant(variant))


# ------------------------------------------------------------------------------------------
# Lanczos drivers
# ------------------------------------------------------------------------------------------
def sp_lanc_eigh(nitermax: int, threshold: float = 1e-12, ncheck: int = 10, vect=None,
                 seed: int = 4321, want_vector: bool = True, inplace: bool = False):
    """sp_lanc_eigh(MatVec, egs, vect, Nitermax, threshold=): returns (egs, vect, niter).
    ``inplace=True`` overwrites the caller's ``vect`` with the eigenvector like the reference's
    ``intent(inout)`` argument does (no host copy; a page-locked array is then used as it is)."""
    L = _abi.load()
    n = vecDim_Hv_sector_normal()
    dt = np.complex128 if _open_is_complex else np.float64  # complex stored-H sector
    use_start = vect is not None
    buf = None
    if use_start:
        buf = np.ascontiguousarray(vect, dt)
        if buf is vect and not inplace:
            buf = buf.copy()
    elif want_vector:
        buf = np.zeros(n, dt)
    egs, nit = C.c_double(), C.c_int()
    check(L.edgpu_lanczos_gs(nitermax, threshold, ncheck, int(use_start), seed, C.byref(egs),
                             ptr(buf) if buf is not None else None, C.byref(nit)))
    return egs.value, buf, nit.value


def lanczos_last_info():
    """(Lanczos vectors kept in HBM, H x v products) of the last :func:`sp_lanc_eigh`."""
    a, b = C.c_int(), C.c_int()
    check(_abi.load().edgpu_lanczos_last_info(C.byref(a), C.byref(b)))
    return a.value, b.value


def release_cache():
    """Returns the pooled Lanczos-vector buffers to the device allocator."""
    check(_abi.load().edgpu_release_cache())


def sp_lanc_tridiag(vin, nlanc: int, threshold: float = 1e-12):
    """sp_lanc_tridiag(MatVec, vin, alanc, blanc); vin=None uses the device-resident seed
    left by :func:`apply_op`.  Returns (alanc, blanc, nused, norm2)."""
    L = _abi.load()
    a, b = np.zeros(nlanc), np.zeros(nlanc)
    nused, n2 = C.c_int(), C.c_double()
    seed = None if vin is None else np.ascontiguousarray(
        vin, np.complex128 if _open_is_complex else np.float64)
    check(L.edgpu_lanczos_tridiag(ptr(seed) if seed is not None else None, nlanc, threshold,
                                  ptr(a), ptr(b), C.byref(nused), C.byref(n2)))
    return a, b, nused.value, n2.value


def sp_eigh(neigen: int, nblock: int, nitermax: int, tol: float = 0.0, seed: int = 4321,
            want_vectors: bool = True):
    """sp_eigh(MatVec, eval(Neigen), evec(Nloc,Neigen), Nblock, Nitermax, tol=): the `neigen`
    lowest eigenpairs of the open sector (thick-restart Lanczos on the device in place of
    (P-)ARPACK).  Returns (evals[neigen], evecs[Nloc, neigen] or None, nconv, nmatvec); the
    eigenvectors also stay on the device, see :func:`eigh_state_store`."""
    L = _abi.load()
    n = vecDim_Hv_sector_normal()
    dim = int(L.edgpu_sector_dim())
    neigen = max(1, min(neigen, dim))
    dt = np.complex128 if _open_is_complex else np.float64
    ev = np.zeros(neigen)
    vec = np.zeros((neigen, n), dt) if want_vectors else None
    nconv, nmv = C.c_int(), C.c_int()
    check(L.edgpu_eigh(neigen, nblock, nitermax, tol, seed, ptr(ev),
                       ptr(vec) if vec is not None else None, C.byref(nconv), C.byref(nmv)))
    return ev, (vec.T if vec is not None else None), nconv.value, nmv.value


def eigh_state_store(k: int, slot: int):
    """es_add_state of eigenvector `k` (0-based) of the last :func:`sp_eigh` (device copy)."""
    check(_abi.load().edgpu_eigh_state_store(k, slot))


def tridiag_Hv_sector_normal(model: EDModel, nup: int, ndw: int, vvinit, nlanc=None):
    """ED_HAMILTONIAN_NORMAL.f90:321-369: norm2, normalise, build sector, tridiagonalise, delete."""
    build_Hv_sector_normal(model, nup, ndw)
    try:
        DimUp, DimDw, _, _ = sector_dims()
        nl = min(DimUp * DimDw, model.lanc_ngfiter) if nlanc is None else nlanc
        a, b, nused, n2 = sp_lanc_tridiag(vvinit, nl)
    finally:
        delete_Hv_sector_normal()
    return a, b, nused, n2


def state_store(slot: int):
    check(_abi.load().edgpu_state_store(slot))


def state_free(slot: int):
    check(_abi.load().edgpu_state_free(slot))


def es_return_vector(slot: int) -> np.ndarray:
    """es_return_dvector / es_return_cvector (ED_EIGENSPACE.f90:620-793): this rank's chunk of the
    stored state; its own sector must be open."""
    n = vecDim_Hv_sector_normal()
    v = np.zeros(n, np.complex128 if _open_is_complex else np.float64)
    check(_abi.load().edgpu_state_download(slot, ptr(v)))
    return v


def state_twin(src_slot: int, dst_slot: int):
    """es_return_dvector / es_return_cvector of a twin state (ED_EIGENSPACE.f90:640-660, 723-793;
    twin_sector_order, ED_SECTOR.f90:1747-1776): state ``dst_slot`` := state ``src_slot`` re-ordered
    into its twin sector, which must be the open one."""
    check(_abi.load().edgpu_state_twin(src_slot, dst_slot))


def apply_op(slot: int, op: int, iorb: int, spin: int):
    """apply_op_CDG (op=+1) / apply_op_C (op=-1) on stored state -> device-resident seed."""
    check(_abi.load().edgpu_apply_op(slot, op, iorb, spin))


def apply_Cops_normal(slot: int, coefs, op: int, iorbs, spin: int):
    """apply_Cops in NORMAL mode (lanc_build_gf_normal_mix, ED_GF_NORMAL.f90:211,227): device seed
    sum_k coefs[k] O_k |state(slot)>, all O_k = c^+ (op=+1) / c (op=-1) of one spin."""
    cf = np.ascontiguousarray(np.asarray(coefs, np.float64))
    a = np.ascontiguousarray(iorbs, np.int32)
    check(_abi.load().edgpu_apply_ops_normal(slot, len(a), ptr(cf), op, ptr(a), spin))


def apply_Cops(slot: int, coefs, ops, iorbs, spins):
    """apply_COps (ED_SECTOR.f90): device seed sum_k coefs[k] * O_k |state(slot)> in the open
    device-built nonsu2 / superc sector; ops[k] = +1 (c^+) / -1 (c), spins[k] = 0 up / 1 dw."""
    cf = np.ascontiguousarray(np.asarray(coefs, np.complex128))
    o = np.ascontiguousarray(ops, np.int32)
    a = np.ascontiguousarray(iorbs, np.int32)
    sp = np.ascontiguousarray(spins, np.int32)
    check(_abi.load().edgpu_apply_ops_packed(slot, len(o), ptr(cf), ptr(o), ptr(a), ptr(sp)))


def seed_norm2() -> float:
    n2 = C.c_double()
    check(_abi.load().edgpu_seed_norm2(C.byref(n2)))
    return n2.value


def state_observables(slot: int, Norb: int):
    dens, docc = np.zeros(Norb), np.zeros(Norb)
    check(_abi.load().edgpu_state_observables(slot, ptr(dens), ptr(docc)))
    return dens, docc


# ------------------------------------------------------------------------------------------
# solver-level sequencing (T = 0)
# ------------------------------------------------------------------------------------------
@dataclass
class EState:
    e: float
    nup: int
    ndw: int
    slot: int


def ed_diag_d(model: EDModel, sectors=None):
    """Sector loop of ``ed_diag_d`` (ED_DIAG_NORMAL.f90:76-296).  Per sector: Neigen / Nblock /
    Nitermax as at :119-128, then ``sp_eigh`` (LANC_METHOD=arpack, the default) or
    ``sp_lanc_eigh`` (LANC_METHOD=lanczos) on the device.  State list: T=0 keeps the states within
    ``gs_threshold`` of the minimum (:267-277); ``ed_finite_temp`` keeps the ``lanc_nstates_total``
    lowest ones (es_add_state(size=), ED_EIGENSPACE.f90:265-274) and trims those below ``cutoff``
    (ed_post_diag :489-501).  Eigenvectors never leave HBM (state slots).  Returns the list sorted
    by energy."""
    Ns = model.Ns
    if sectors is None:
        sectors = [(nu, nd) for nu in range(Ns + 1) for nd in range(Ns + 1)]
    finiteT = model.ed_finite_temp
    if model.ed_twin:
        # twin_mask (ED_SETUP.f90:592-602): of each pair (nup,ndw) / (ndw,nup) the sector with
        # nup > ndw stays on; its states enter the list together with their twins
        # (es_insert_state, ED_EIGENSPACE.f90:344-350)
        if finiteT:
            raise EdgpuError("ed_twin with ed_finite_temp is not mirrored (es_add_state(size=) counts twin pairs)")
        sectors = [(nu, nd) for (nu, nd) in sectors if nu >= nd]
    nst_sector, nst_total = model.lanc_nstates_sector, model.lanc_nstates_total
    if finiteT:  # ED_SETUP.f90:279-287
        nst_sector += nst_sector % 2
        nst_total += nst_total % 2
    states: list[EState] = []
    oldzero = 1000.0
    next_slot = 0

    def drop(st):
        state_free(st.slot)
        states.remove(st)

    for nup, ndw in sectors:
        dim = binomial(Ns, nup) * binomial(Ns, ndw)
        nitermax = min(dim, model.lanc_niter)
        if model.lanc_method == "lanczos":
            neigen, nblock = 1, 1
        else:
            neigen = min(dim, nst_sector)  # neigen_sector, ED_SETUP.f90:554
            nblock = min(dim, model.lanc_ncv_factor * max(neigen, nst_sector) + model.lanc_ncv_add)
        build_Hv_sector_normal(model, nup, ndw)
        try:
            if model.lanc_method == "lanczos":
                e0, _, _ = sp_lanc_eigh(nitermax, model.lanc_tolerance, want_vector=False)
                evals = [e0]
            else:
                ev, _, _, _ = sp_eigh(neigen, nblock, nitermax, model.lanc_tolerance,
                                      want_vectors=False)
                evals = list(ev)
            for i, e in enumerate(evals):
                if finiteT:
                    if len(states) >= nst_total:
                        worst = max(states, key=lambda s: s.e)
                        if e >= worst.e:
                            continue
                        drop(worst)  # es_pop_state
                elif e < oldzero - 10.0 * model.gs_threshold:
                    oldzero = e
                    for st in list(states):
                        drop(st)
                elif abs(e - oldzero) <= model.gs_threshold:
                    oldzero = min(oldzero, e)
                else:
                    continue
                if model.lanc_method == "lanczos":
                    state_store(next_slot)
                else:
                    eigh_state_store(i, next_slot)
                states.append(EState(float(e), nup, ndw, next_slot))
                next_slot += 1
        finally:
            delete_Hv_sector_normal()
    states.sort(key=lambda s: s.e)
    if finiteT and states:
        egs = states[0].e
        while len(states) > 1 and math.exp(-model.beta * (states[-1].e - egs)) <= model.cutoff:
            drop(states[-1])
    if model.ed_twin:
        # the twin states' vectors: re-ordered on the device (the reference re-orders on the master
        # every time es_return_dvector is called; here once, the copy stays in HBM)
        for st in [s for s in states if s.nup != s.ndw]:
            build_Hv_sector_normal(model, st.ndw, st.nup)
            try:
                state_twin(st.slot, next_slot)
            finally:
                delete_Hv_sector_normal()
            states.append(EState(st.e, st.ndw, st.nup, next_slot))
            next_slot += 1
        states.sort(key=lambda s: s.e)
    return states


def ed_diag_c(model, qns=None, neigen: int = 2, ncv_factor: int = 10, ncv_add: int = 0,
              nitermax: int = 512, tol: float = 1e-18, gs_threshold: float = 1e-9, ed_twin: bool = False):
    """Sector loop of ``ed_diag_c`` (ED_DIAG_NONSU2.f90:72-296, ED_DIAG_SUPERC.f90:73-277) at T=0
    for the packed-state modes: sectors are Ntot = 0..2Ns (:class:`EDModelNonsu2`) or
    Sz = -Ns..Ns (:class:`EDModelSuperc`), built on the device; ``sp_eigh`` with Neigen / Nblock as
    at ED_DIAG_NONSU2.f90:119-124; state list with the ``gs_threshold`` rule (:262-278).  The
    returned :class:`EState` carries the quantum number in ``nup`` (``ndw`` = 0)."""
    Ns = model.Ns
    superc = isinstance(model, EDModelSuperc)
    if qns is None:
        qns = range(-Ns, Ns + 1) if superc else range(0, 2 * Ns + 1)
    if ed_twin:  # twin_mask: superc keeps Sz <= 0 (ED_SETUP.f90:735-741), nonsu2 Ntot <= Ns (:896-902)
        qns = [q for q in qns if (q <= 0 if superc else q <= Ns)]
    twin_of = (lambda q: -q) if superc else (lambda q: 2 * Ns - q)
    build = build_Hv_sector_superc if superc else build_Hv_sector_nonsu2
    states: list[EState] = []
    oldzero = 1000.0
    next_slot = 0
    for q in qns:
        build(model, q)
        try:
            dim = int(_abi.load().edgpu_sector_dim())
            ne = min(dim, neigen)
            nblock = min(dim, ncv_factor * max(ne, neigen) + ncv_add)
            ev, _, _, _ = sp_eigh(ne, nblock, min(dim, nitermax), tol, want_vectors=False)
            for i, e in enumerate(ev):
                if e < oldzero - 10.0 * gs_threshold:
                    oldzero = e
                    for st in states:
                        state_free(st.slot)
                    states = []
                elif abs(e - oldzero) <= gs_threshold:
                    oldzero = min(oldzero, e)
                else:
                    continue
                eigh_state_store(i, next_slot)
                states.append(EState(float(e), q, 0, next_slot))
                next_slot += 1
        finally:
            delete_Hv_sector_csr()
    states.sort(key=lambda s: s.e)
    if ed_twin:
        for st in [s for s in states if twin_of(s.nup) != s.nup]:
            build(model, twin_of(st.nup))
            try:
                state_twin(st.slot, next_slot)
            finally:
                delete_Hv_sector_csr()
            states.append(EState(st.e, twin_of(st.nup), 0, next_slot))
            next_slot += 1
        states.sort(key=lambda s: s.e)
    return states


def observables_packed(model, states):
    """dens / docc averaged over the T=0 state list (ED_OBSERVABLES_NONSU2 / _SUPERC :150-165)."""
    superc = isinstance(model, EDModelSuperc)
    build = build_Hv_sector_superc if superc else build_Hv_sector_nonsu2
    dens, docc = np.zeros(model.Norb), np.zeros(model.Norb)
    for st in states:
        build(model, st.nup)
        try:
            d, o = state_observables(st.slot, model.Norb)
        finally:
            delete_Hv_sector_csr()
        dens += d / len(states)
        docc += o / len(states)
    return dens, docc


def exciton_nonsu2(model, st, iorb: int = 0, jorb: int = 1):
    """Excitonic order parameters [S0, Tx, Ty, Tz](iorb,jorb) of one stored nonsu2 state (weight 1),
    ED_OBSERVABLES_NONSU2.f90:300-425: every term is the norm of a device-built seed
    apply_Cops(v,[1,c],[-1,-1],[a,b],[s,s']) in the sector Ntot-1 (<n_{a s}> = |c_{a s} v|^2 gives
    dens and magZ the same way)."""
    build_Hv_sector_nonsu2(model, st.nup - 1)
    try:
        def n2(coefs, orbs, spins):
            apply_Cops(st.slot, coefs, [-1] * len(orbs), orbs, spins)
            return seed_norm2()

        occ = {(a, s): n2([1.0], [a], [s]) for a in (iorb, jorb) for s in (0, 1)}
        th_uu, th_dd = n2([1, 1], [iorb, jorb], [0, 0]), n2([1, 1], [iorb, jorb], [1, 1])
        th_ud, th_du = n2([1, 1], [iorb, jorb], [0, 1]), n2([1, 1], [iorb, jorb], [1, 0])
        om_ud, om_du = n2([1, -1j], [iorb, jorb], [0, 1]), n2([1, -1j], [iorb, jorb], [1, 0])
    finally:
        delete_Hv_sector_csr()
    dens = {a: occ[a, 0] + occ[a, 1] for a in (iorb, jorb)}
    magz = {a: occ[a, 0] - occ[a, 1] for a in (iorb, jorb)}
    return np.array([th_uu + th_dd - dens[iorb] - dens[jorb],
                     th_ud + th_du - dens[iorb] - dens[jorb],
                     om_ud - om_du - magz[iorb] + magz[jorb],
                     th_uu - th_dd - magz[iorb] - magz[jorb]])


def boltzmann_weights(model: EDModel, states):
    """Weights of the state list: 1/zeta at T=0 (zeta = number of kept states,
    ED_DIAG_NORMAL.f90:405-414), exp(-beta (E_i - E_gs)) / zeta at finite temperature."""
    if not states:
        return []
    if not model.ed_finite_temp:
        return [1.0 / len(states)] * len(states)
    egs = min(s.e for s in states)
    w = [math.exp(-model.beta * (s.e - egs)) for s in states]
    z = sum(w)
    return [x / z for x in w]


def observables_normal(model: EDModel, states):
    """dens/docc of ED_OBSERVABLES_NORMAL.f90:150-215 (Boltzmann-weighted over the state list)."""
    dens, docc = np.zeros(model.Norb), np.zeros(model.Norb)
    for st, w in zip(states, boltzmann_weights(model, states)):
        build_Hv_sector_normal(model, st.nup, st.ndw)
        try:
            d, o = state_observables(st.slot, model.Norb)
        finally:
            delete_Hv_sector_normal()
        dens += d * w
        docc += o * w
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


def local_energy_normal(model: EDModel, states):
    """local_energy_normal (ED_OBSERVABLES_NORMAL.f90:491-930), DimPh = 1: dict(Epot, Eint, Ehartree,
    Eknot, Dust, Dund, Dse, Dph) = ed_get_eimp / ed_get_doubles.  Every component is <v|O|v> with O
    the model Hamiltonian restricted to one group of terms: evaluated on the device as the first
    Lanczos coefficient alpha_1 = <v|H_variant|v> of the stored state (one H x v + one dot per
    variant; the reference gathers the state to the master and loops over Dim there)."""
    variants = _energy_variants(model)
    acc = {k: 0.0 for k in variants}
    for st, peso in zip(states, boltzmann_weights(model, states)):
        # the stored state is laid out in the enumeration order of ITS model's sector plan (which
        # depends on the one-body term list): fetch it under that sector, hand it to the variants
        # in the reference layout
        build_Hv_sector_normal(model, st.nup, st.ndw)
        try:
            vec = es_return_vector(st.slot)
        finally:
            delete_Hv_sector_normal()
        for k, mv in variants.items():
            build_Hv_sector_normal(mv, st.nup, st.ndw)
            try:
                a, _, nused, n2 = sp_lanc_tridiag(vec, 1)
            finally:
                delete_Hv_sector_normal()
            acc[k] += peso * float(a[0]) * n2
    dens, _ = observables_normal(model, states)
    out = {k.capitalize() if k.startswith("d") else "E" + k[1:]: acc[k] - acc["base"]
           for k in variants if k != "base"}
    out["Ehartree"] = _hartree_energy(model, dens)
    out["Epot"] = out["Eint"] + out["Ehartree"]
    return out


def imp_info(model: EDModel, states, energies=None, dens=None, docc=None):
    """ed_imp_info = [s2tot, egs] (ED_OBSERVABLES_NORMAL.f90:180, 452).  s2tot = <(sum_a S^z_a)^2>
    is a combination of quantities already evaluated on the device:
    (sum_a sz_a)^2 = 1/4 [ sum_a (n_a - 2 nup_a ndw_a) + 2 sum_{a<b} (nup_a nup_b + ndw_a ndw_b)
                                                         - 2 sum_{a<b} (nup_a ndw_b + nup_b ndw_a) ]
                   = 1/4 [ sum_a (dens_a - 2 docc_a) + 2 (Dund - Dust) ]."""
    if energies is None:
        energies = local_energy_normal(model, states)
    if dens is None or docc is None:
        dens, docc = observables_normal(model, states)
    s2 = 0.25 * (float(np.sum(dens - 2.0 * docc)) + 2.0 * (energies["Dund"] - energies["Dust"]))
    return np.array([s2, min(s.e for s in states)])


def tridiag_eigh(a, b_sub):
    """eigh(diag,subdiag,Ev=Z) of ED_GF_NORMAL.f90:416 (host, tiny)."""
    n = len(a)
    T = np.diag(np.asarray(a, float))
    for i in range(n - 1):
        T[i, i + 1] = T[i + 1, i] = b_sub[i]
    return np.linalg.eigh(T)


def lanc_build_gf_normal_diag(model: EDModel, states, iorb: int, ispin: int = 0):
    """ED_GF_NORMAL.f90:131-177 + add_to_lanczos_gf_normal :363-427: list of (weight, pole).
    Seeds c^+|gs>, c|gs> are built on the device from the resident ground states."""
    out = []
    Ns = model.Ns
    for st, peso in zip(states, boltzmann_weights(model, states)):
        for op, isign in ((+1, 1), (-1, -1)):
            jn = (st.nup + (op if ispin == 0 else 0), st.ndw + (op if ispin == 1 else 0))
            if min(jn) < 0 or max(jn) > Ns:
                continue
            build_Hv_sector_normal(model, jn[0], jn[1])
            try:
                apply_op(st.slot, op, iorb, ispin)
                dim = binomial(Ns, jn[0]) * binomial(Ns, jn[1])
                a, b, nused, norm2 = sp_lanc_tridiag(None, min(dim, model.lanc_ngfiter))
            finally:
                delete_Hv_sector_normal()
            if norm2 == 0.0 or nused == 0:
                continue
            ev, Z = tridiag_eigh(a[:nused], b[1:nused])
            for j in range(nused):
                out.append((norm2 * peso * Z[0, j] ** 2, isign * (ev[j] - st.e)))
    return out


def lanc_build_gf_normal_mix(model: EDModel, states, iorb: int, jorb: int, ispin: int = 0):
    """ED_GF_NORMAL.f90:182-262 (real case): poles / weights of the auxiliary function from the
    seeds (c^+_a + c^+_b)|gs> and (c_a + c_b)|gs>, built on the device (apply_Cops_normal)."""
    out = []
    Ns = model.Ns
    for st, peso in zip(states, boltzmann_weights(model, states)):
        for op, isign in ((+1, 1), (-1, -1)):
            jn = (st.nup + (op if ispin == 0 else 0), st.ndw + (op if ispin == 1 else 0))
            if min(jn) < 0 or max(jn) > Ns:
                continue
            build_Hv_sector_normal(model, jn[0], jn[1])
            try:
                apply_Cops_normal(st.slot, [1.0, 1.0], op, [iorb, jorb], ispin)
                dim = binomial(Ns, jn[0]) * binomial(Ns, jn[1])
                a, b, nused, norm2 = sp_lanc_tridiag(None, min(dim, model.lanc_ngfiter))
            finally:
                delete_Hv_sector_normal()
            if norm2 == 0.0 or nused == 0:
                continue
            ev, Z = tridiag_eigh(a[:nused], b[1:nused])
            for j in range(nused):
                out.append((norm2 * peso * Z[0, j] ** 2, isign * (ev[j] - st.e)))
    return out


def _gf_eval(pw, z):
    g = np.zeros(len(z), complex)
    for w, p in pw:
        g += w / (z - p)
    return g


def get_impG_normal(model: EDModel, states, z, ispin: int = 0):
    """get_impG_normal (ED_GF_NORMAL.f90:495-575) with offdiag_gf_flag: G_aa from the diagonal
    builder, G_ab = G_ba = (G_{a+b} - G_aa - G_bb)/2 (:553-560).  [Norb, Norb, len(z)]."""
    No = model.Norb
    z = np.asarray(z, complex)
    G = np.zeros((No, No, len(z)), complex)
    for a in range(No):
        G[a, a] = _gf_eval(lanc_build_gf_normal_diag(model, states, a, ispin), z)
    offdiag = model.bath_type != "normal"   # offdiag_gf_flag defaults to T for bath_type /= normal
    for a in range(No):
        for b in range(a + 1, No):
            if not offdiag:
                continue
            mix = _gf_eval(lanc_build_gf_normal_mix(model, states, a, b, ispin), z)
            G[a, b] = G[b, a] = 0.5 * (mix - G[a, a] - G[b, b])
    return G


def delta_bath_array(model: EDModel, z, ispin: int = 0):
    """delta_bath_array, ed_mode=normal (ED_BATH/delta_functions: delta_normal.f90:33-42,
    delta_hybrid.f90:30-41, delta_replica.f90:27-38 / delta_general.f90).  [Norb, Norb, len(z)]."""
    No, Nb = model.Norb, model.Nbath
    if model.bath_e is None:
        model.init_dmft_bath()
    z = np.asarray(z, complex)
    D = np.zeros((No, No, len(z)), complex)
    if model.bath_type == "normal":
        for a in range(No):
            D[a, a] = (model.bath_v[ispin, a][None, :] ** 2 / (z[:, None] - model.bath_e[ispin, a][None, :])).sum(1)
    elif model.bath_type == "hybrid":
        e = model.bath_e[ispin, 0]
        for a in range(No):
            for b in range(No):
                D[a, b] = (model.bath_v[ispin, a][None, :] * model.bath_v[ispin, b][None, :]
                           / (z[:, None] - e[None, :])).sum(1)
    else:
        for k in range(Nb):
            Hk = np.asarray(model.hbath)[ispin, :, :, k]
            V = np.diag(model.bath_v[ispin, :, k])
            for i, zi in enumerate(z):
                D[:, :, i] += V @ np.linalg.inv(zi * np.eye(No) - Hk) @ V
    return D


def get_Sigma_normal(model: EDModel, states, Lmats: int, ispin: int = 0):
    """get_Sigma_normal (ED_GF_NORMAL.f90:698-739) on the Matsubara axis: Sigma = G0^-1 - G^-1 with
    G0^-1 = (z+xmu) 1 - impHloc - Delta (invg0_normal.f90 / invg0_hyrege.f90:22-30); G is inverted
    as an orbital matrix when bath_type /= normal (:726-729).  Returns (wm, Sigma[Norb,Norb,Lmats])."""
    wm = math.pi / model.beta * (2 * np.arange(1, Lmats + 1) - 1)
    z = 1j * wm
    No = model.Norb
    G = get_impG_normal(model, states, z, ispin)
    D = delta_bath_array(model, z, ispin)
    hl = np.zeros((No, No)) if model.hloc is None else np.asarray(model.hloc[ispin], float)
    S = np.zeros_like(G)
    for i, zi in enumerate(z):
        invg0 = (zi + model.xmu) * np.eye(No) - hl - D[:, :, i]
        invg = np.diag(1.0 / np.diag(G[:, :, i])) if model.bath_type == "normal" else np.linalg.inv(G[:, :, i])
        S[:, :, i] = invg0 - invg
    return wm, S
