// ed_total_ud = F: orbital-resolved sectors of the NORMAL mode (Ns_Ud = Norb, Ns_Orb = 1+Nbath,
// ED_SETUP.f90:128-135).  A sector is labelled by (Nups(1:Norb), Ndws(1:Norb)); its states are
// the mixed-radix tuples [iup_1..iup_Norb, idw_1..idw_Norb] (first fastest, state2indices,
// ED_SECTOR.f90:1691-1702) of per-orbital, per-spin sectors of the 1+Nbath levels {impurity, bath
// 1..Nbath} (build_sector :217-242).  Replaces build_sector + directMatVec_normal_orbs
// (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:133-226 with direct/Orbs/HxV_local.f90, HxV_up.f90,
// HxV_dw.f90) = ed_buildh_normal_orbs + spMatVec_normal_orbs of the stored path: one thread per
// row enumerates the diagonal and the imp <-> bath hops of every factor (count -> prefix -> fill)
// into a real CSR in HBM; targets are ranked (combinadic), not searched.  The product, the Lanczos
// drivers and the eigen-solver are those of csr.cu.  With nranks>1 the rows are split along the
// LAST factor like the reference's MPI layout (mpiQdw of the last dw factor).
#include <algorithm>
#include <vector>

#include "edgpu_internal.cuh"

namespace edgpu {

struct OrbsDev {
  int32_t Norb, Nbath, hfmode, Nfoo, any_sf, pad0;
  double xmu;
  double eloc[2][EDGPU_MAXORB];
  double sfz[EDGPU_MAXORB];
  double Uloc[EDGPU_MAXORB];
  double Ust[EDGPU_MAXORB][EDGPU_MAXORB];
  double Jh[EDGPU_MAXORB][EDGPU_MAXORB];
  double bath_diag[2][EDGPU_MAXORB][EDGPU_MAXBATH];
  double diag_hybr[2][EDGPU_MAXORB][EDGPU_MAXBATH];
  int32_t nel[2 * EDGPU_MAXORB];     // electrons of factor f (f < Norb: up, else dw)
  int32_t dims[2 * EDGPU_MAXORB];
  int64_t strides[2 * EDGPU_MAXORB];
  int32_t binom[34][34];
};
__constant__ OrbsDev c_ob;

__device__ __forceinline__ int ob_rank(uint32_t m) {
  int r = 0, k = 0;
  while (m) {
    const int p = __ffs(m) - 1;
    m &= m - 1;
    k++;
    r += c_ob.binom[p][k];
  }
  return r;
}
__device__ __forceinline__ uint32_t ob_unrank(int r, int nbits, int k) {
  uint32_t m = 0;
  int p = nbits - 1;
  for (; k >= 1; k--) {
    while (c_ob.binom[p][k] > r) p--;
    m |= 1u << p;
    r -= c_ob.binom[p][k];
    p--;
  }
  return m;
}

struct ObCount {
  int n = 0;
  __device__ void emit(int64_t, double) { n++; }
};
struct ObFill {
  int32_t *cols;
  double *vals;
  int64_t k;
  __device__ void emit(int64_t col, double v) {
    cols[k] = (int32_t)col;
    vals[k] = v;
    k++;
  }
};

// direct (ED_SPARSE_H = F) product: the element is applied to the input vector instead of stored
struct ObApply {
  const double *v;
  double acc = 0.0;
  __device__ void emit(int64_t col, double val) { acc += val * v[col]; }
};

template <class Sink>
__device__ void ob_row(int64_t i, Sink &s) {
  const OrbsDev &P = c_ob;
  const int No = P.Norb, Nb = P.Nbath, nso = Nb + 1;
  uint32_t pat[2 * EDGPU_MAXORB];
  int idx[2 * EDGPU_MAXORB];
  {
    int64_t c = i;
    for (int f = 0; f < 2 * No; f++) {
      idx[f] = (int)(c % P.dims[f]);
      c /= P.dims[f];
      pat[f] = ob_unrank(idx[f], nso, P.nel[f]);
    }
  }
  // ---- direct/Orbs/HxV_local.f90
  double nup[EDGPU_MAXORB], ndw[EDGPU_MAXORB];
  for (int a = 0; a < No; a++) {
    nup[a] = (double)(pat[a] & 1u);
    ndw[a] = (double)(pat[a + No] & 1u);
  }
  double h = 0.0;
  for (int a = 0; a < No; a++) h += P.eloc[0][a] * nup[a] + P.eloc[1][a] * ndw[a] - P.xmu * (nup[a] + ndw[a]);
  if (P.any_sf)
    for (int a = 0; a < No; a++) h += P.sfz[a] * (nup[a] - ndw[a]);
  for (int a = 0; a < No; a++) h += P.Uloc[a] * nup[a] * ndw[a];
  if (No > 1) {
    for (int a = 0; a < No; a++)
      for (int b = a + 1; b < No; b++) h += P.Ust[a][b] * (nup[a] * ndw[b] + nup[b] * ndw[a]);
    for (int a = 0; a < No; a++)
      for (int b = a + 1; b < No; b++) h += (P.Ust[a][b] - P.Jh[a][b]) * (nup[a] * nup[b] + ndw[a] * ndw[b]);
  }
  if (P.hfmode) {
    for (int a = 0; a < No; a++) h += -0.5 * P.Uloc[a] * (nup[a] + ndw[a]) + 0.25 * P.Uloc[a];
    if (No > 1)
      for (int a = 0; a < No; a++)
        for (int b = a + 1; b < No; b++) {
          const double nn = nup[a] + ndw[a] + nup[b] + ndw[b];
          // 0.25 (not the 0.5 of the ed_total_ud=T fragment): Orbs/HxV_local.f90:66-67
          h += -0.5 * P.Ust[a][b] * nn + 0.25 * P.Ust[a][b];
          h += -0.5 * (P.Ust[a][b] - P.Jh[a][b]) * nn + 0.25 * (P.Ust[a][b] - P.Jh[a][b]);
        }
  }
  for (int a = 0; a < P.Nfoo; a++)
    for (int k = 0; k < Nb; k++)
      h += P.bath_diag[0][a][k] * (double)((pat[a] >> (1 + k)) & 1u) +
           P.bath_diag[1][a][k] * (double)((pat[a + No] >> (1 + k)) & 1u);
  s.emit(i, h);
  // ---- direct/Orbs/HxV_up.f90, HxV_dw.f90: imp <-> bath level of the same orbital and spin.
  // sign = occupied bath levels below the one operated on (c / cdg count inside the factor's own
  // (1+Nbath)-bit integer; the impurity bit is empty when it is passed over)
  for (int spin = 0; spin < 2; spin++)
    for (int a = 0; a < No; a++) {
      const int f = a + spin * No;
      const uint32_t m = pat[f];
      for (int k = 0; k < Nb; k++) {
        const double amp = P.diag_hybr[spin][a][k];
        const uint32_t bb = 1u << (1 + k);
        if (amp == 0.0 || ((m & 1u) != 0u) == ((m & bb) != 0u)) continue;
        const int par = __popc(m & (bb - 1u) & ~1u) & 1;
        const uint32_t m2 = m ^ 1u ^ bb;
        s.emit(i + (int64_t)(ob_rank(m2) - idx[f]) * P.strides[f], par ? -amp : amp);
      }
    }
}

__global__ void __launch_bounds__(128) k_ob_count(int64_t row0, int64_t nloc, int32_t *__restrict__ cnt) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  ObCount s;
  ob_row(row0 + r, s);
  cnt[r] = s.n;
}
__global__ void __launch_bounds__(128)
k_ob_fill(int64_t row0, int64_t nloc, const int64_t *__restrict__ rowptr, int32_t *__restrict__ cols,
          double *__restrict__ vals) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  ObFill s{cols, vals, rowptr[r]};
  ob_row(row0 + r, s);
}

// directMatVec_normal_orbs (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:134-227, direct/Orbs/HxV_local.f90,
// HxV_up.f90, HxV_dw.f90) without a stored matrix: the row generator applied to the (all-gathered)
// input vector; nothing is kept in HBM (the sector is a mixed-radix product, rows are unranked)
__global__ void __launch_bounds__(128)
k_ob_direct(int64_t row0, int64_t nloc, const double *__restrict__ vin, double *__restrict__ hv, int accum,
            double s_acc, double s_old) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  ObApply s{vin};
  ob_row(row0 + r, s);
  hv[r] = accum ? s_acc * s.acc + s_old * hv[r] : s_acc * s.acc;
}

int orbs_direct_hxv(Engine &E, const double *d_vin_full, double *d_hv, bool accum, double s_acc, double s_old) {
  const CsrSector &C = E.csr;
  if (C.nloc <= 0) return 0;
  k_ob_direct<<<(unsigned)((C.nloc + 127) / 128), 128, 0, E.stream>>>(C.row0, C.nloc, d_vin_full, d_hv,
                                                                     (int)accum, s_acc, s_old);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

static OrbsDev g_ob_host;

int orbs_open(Engine &E, const edgpu_normal_params *p, const int32_t *nups, const int32_t *ndws) {
  if (!E.inited) return set_error("edgpu_init was not called");
  if (E.sec.open) sector_close(E);
  if (E.csr.open) csr_close(E);
  const int No = p->Norb, Nb = p->Nbath;
  if (No < 1 || No > EDGPU_MAXORB || Nb < 0 || Nb > 31) return set_error("orbs: Norb/Nbath out of range");
  // "bath_type==hybrid AND .NOT.ed_total_ud" stops (ED_SETUP.f90:124); replica/general baths and any
  // other inter-orbital term do not conserve the per-orbital occupations
  if (p->bath_type != EDGPU_BATH_NORMAL) return set_error("ed_total_ud=F needs bath_type=normal");
  if (E.Nph > 0) return set_error("ed_total_ud=F sectors with phonons are not built on the device");
  for (int a = 0; a < No; a++)
    for (int b = 0; b < No; b++) {
      if (a == b) continue;
      if (p->Jx[a][b] != 0.0 || p->Jp[a][b] != 0.0)
        return set_error("ED ERROR: ed_total_ud=F cannot be used if non-density-density interaction terms are present");
      if (p->eloc[0][a][b] != 0.0 || p->eloc[1][a][b] != 0.0)
        return set_error("ed_total_ud=F: inter-orbital impHloc terms do not conserve the orbital occupations");
    }
  if (No > 1 && !E.sundry_terms.empty())
    return set_error("ED ERROR: ed_total_ud=F cannot be used if non-density-density interaction terms are present");
  if (p->exc_field[0] != 0.0 || p->exc_field[1] != 0.0 || p->exc_field[2] != 0.0 || p->exc_field[3] != 0.0)
    return set_error("ed_total_ud=F: exc_field couples different orbitals");
  OrbsDev &h = g_ob_host;
  h = OrbsDev();
  h.Norb = No;
  h.Nbath = Nb;
  h.hfmode = p->hfmode;
  h.Nfoo = p->Nfoo;
  h.xmu = p->xmu;
  for (int a = 0; a < No; a++) {
    h.eloc[0][a] = p->eloc[0][a][a];
    h.eloc[1][a] = p->eloc[1][a][a];
    h.sfz[a] = p->spin_field_z[a];
    if (h.sfz[a] != 0.0) h.any_sf = 1;
    h.Uloc[a] = p->Uloc[a];
    for (int b = 0; b < No; b++) {
      h.Ust[a][b] = p->Ust[a][b];
      h.Jh[a][b] = p->Jh[a][b];
    }
    for (int k = 0; k < Nb; k++)
      for (int s = 0; s < 2; s++) {
        h.bath_diag[s][a][k] = p->bath_diag[s][a][k];
        h.diag_hybr[s][a][k] = p->diag_hybr[s][a][k];
      }
  }
  for (int n = 0; n < 34; n++)
    for (int k = 0; k < 34; k++) {
      const int64_t c = k > n ? 0 : (k == 0 || k == n ? 1 : (int64_t)h.binom[n - 1][k - 1] + h.binom[n - 1][k]);
      h.binom[n][k] = (int32_t)std::min<int64_t>(c, INT32_MAX);
    }
  int64_t dim = 1;
  for (int f = 0; f < 2 * No; f++) {
    const int n = f < No ? nups[f] : ndws[f - No];
    if (n < 0 || n > Nb + 1) return set_error("orbs: occupation %d of factor %d outside [0,%d]", n, f, Nb + 1);
    h.nel[f] = n;
    h.dims[f] = (int32_t)host_binomial(Nb + 1, n);
    h.strides[f] = dim;
    dim *= h.dims[f];
    if (dim > INT32_MAX) return set_error("orbs: sector dimension exceeds 32-bit columns");
  }
  // rows of this rank: the last factor is split like mpiQdw (first ranks get the remainder)
  const int last = 2 * No - 1;
  int64_t q = h.dims[last], d0 = 0;
  std::vector<int64_t> counts, offs;
  if (E.nranks > 1) {
    if (h.dims[last] < E.nranks) return set_error("orbs: last factor smaller than the communicator");
    counts.resize(E.nranks);
    offs.resize(E.nranks);
    for (int r = 0; r < E.nranks; r++) {
      int64_t qq, dd;
      block_split(h.dims[last], E.nranks, r, &qq, &dd);
      counts[r] = qq * h.strides[last];
      offs[r] = dd * h.strides[last];
    }
    block_split(h.dims[last], E.nranks, E.rank, &q, &d0);
  }
  const int64_t row0 = d0 * h.strides[last], nloc = q * h.strides[last];
  EDGPU_CUDA(cudaMemcpyToSymbolAsync(c_ob, &h, sizeof(h), 0, cudaMemcpyHostToDevice, E.stream));
  int32_t *d_cnt = nullptr, *d_cols = nullptr;
  int64_t *d_rowptr = nullptr;
  double *d_vals = nullptr;
  auto fail = [&](int rc) {
    cudaFree(d_cnt);
    cudaFree(d_cols);
    cudaFree(d_rowptr);
    cudaFree(d_vals);
    return rc;
  };
#define OB_CUDA(call)                                                                                     \
  do {                                                                                                    \
    cudaError_t _e = (call);                                                                              \
    if (_e != cudaSuccess)                                                                                \
      return fail(set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__)); \
  } while (0)
  if (!E.sparse_h) {
    // ED_SPARSE_H = F: directMatVec_normal_orbs, nothing stored
    OB_CUDA(cudaStreamSynchronize(E.stream));
    int rc = csr_adopt_device(E, false, nloc, dim, row0, nullptr, nullptr, nullptr, 0, nullptr,
                              E.nranks > 1 ? &counts : nullptr, E.nranks > 1 ? &offs : nullptr);
    if (rc) return fail(rc);
    E.csr.direct = true;
    E.csr.direct_orbs = true;
    return 0;
  }
  OB_CUDA(cudaMalloc(&d_cnt, sizeof(int32_t) * std::max<int64_t>(nloc, 1)));
  const unsigned grid = (unsigned)std::max<int64_t>(1, (nloc + 127) / 128);
  k_ob_count<<<grid, 128, 0, E.stream>>>(row0, nloc, d_cnt);
  EDGPU_COUNT_LAUNCH();
  OB_CUDA(cudaGetLastError());
  std::vector<int32_t> cnt((size_t)std::max<int64_t>(nloc, 1));
  OB_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int32_t) * nloc, cudaMemcpyDeviceToHost, E.stream));
  OB_CUDA(cudaStreamSynchronize(E.stream));
  std::vector<int64_t> rowptr((size_t)nloc + 1, 0);
  for (int64_t r = 0; r < nloc; r++) rowptr[(size_t)r + 1] = rowptr[(size_t)r] + cnt[(size_t)r];
  const int64_t nnz = rowptr[(size_t)nloc];
  OB_CUDA(cudaMalloc(&d_rowptr, sizeof(int64_t) * (nloc + 1)));
  OB_CUDA(cudaMalloc(&d_cols, sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
  OB_CUDA(cudaMalloc(&d_vals, sizeof(double) * std::max<int64_t>(nnz, 1)));
  OB_CUDA(cudaMemcpyAsync(d_rowptr, rowptr.data(), sizeof(int64_t) * (nloc + 1), cudaMemcpyHostToDevice, E.stream));
  k_ob_fill<<<grid, 128, 0, E.stream>>>(row0, nloc, d_rowptr, d_cols, d_vals);
  EDGPU_COUNT_LAUNCH();
  OB_CUDA(cudaGetLastError());
  OB_CUDA(cudaStreamSynchronize(E.stream));
#undef OB_CUDA
  cudaFree(d_cnt);
  d_cnt = nullptr;
  int rc = csr_adopt_device(E, false, nloc, dim, row0, d_rowptr, d_cols, d_vals, nnz, nullptr,
                            E.nranks > 1 ? &counts : nullptr, E.nranks > 1 ? &offs : nullptr);
  if (rc) return fail(rc);
  return 0;
}

}  // namespace edgpu
