// the terms of directMatVec_normal_main that no BASELINE config switches on --
// user two-body operators (coulomb_sundry, direct/HxV_sundry.f90:1-109) and phonons
// (direct/HxV_ph.f90:1-6, direct/HxV_eph.f90:1-81; DimPh = Nph+1 slices, ED_SETUP.f90:137).
// Both are applied after the two tiled electronic passes as thread-per-state gather kernels:
// they touch O(1) extra elements per state, so they stay bound by the 16 B/state stream of hv.
#include <algorithm>
#include <vector>

#include "edgpu_internal.cuh"

namespace edgpu {

// ---------------------------------------------------------------------------------------
// coulomb_sundry: Hv(j) += U sg1 sg2 sg3 sg4 vin(i), |i> = cd_i c_k cd_j c_l |j> (operators applied
// right to left on the ROW state, HxV_sundry.f90:41-104).  c / cdg of ED_AUX_FUNX.f90:334-384:
// sign = parity of the occupied bits below the operated one IN THE SAME spin integer.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_sundry(const double *__restrict__ vfull, double *__restrict__ hv, int64_t nrow, int64_t ldv,
         int64_t col_offset, int64_t slice, int64_t slice_full, const int32_t *__restrict__ mapu,
         const int32_t *__restrict__ mapd, RankView Ru, RankView Rd,
         const SundryDev *__restrict__ T, int nT, double s_acc) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  const int64_t iph = blockIdx.z;
  if (i >= nrow) return;
  const uint32_t mu0 = (uint32_t)mapu[i], md0 = (uint32_t)mapd[c + col_offset];
  const double *__restrict__ vs = vfull + iph * slice_full;
  double acc = 0.0;
  for (int t = 0; t < nT; t++) {
    const SundryDev op = T[t];
    uint32_t mu = mu0, md = md0;
    int par = 0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t m = op.dw[k] ? md : mu;
      const uint32_t b = 1u << op.bit[k];
      const bool occ = (m & b) != 0u;
      if (occ == (op.create[k] != 0)) {
        ok = false;
        break;
      }
      par ^= __popc(m & (b - 1u)) & 1;
      if (op.dw[k])
        md ^= b;
      else
        mu ^= b;
    }
    if (!ok) continue;
    const int64_t ru = rank_of(mu, Ru), rd = rank_of(md, Rd);
    acc += (par ? -op.U : op.U) * vs[rd * ldv + ru];
  }
  if (acc != 0.0) hv[iph * slice + c * ldv + i] += s_acc * acc;
}

// ---------------------------------------------------------------------------------------
// phonons, gather form of the reference's scatter loops: row (T, n) receives
//   w0 n v(T,n)                                                        HxV_ph.f90:2-5
//   sqrt(max(n,n')) [A + sum_a g(a,a)(nup_a+ndw_a)] v(T,n')             HxV_eph.f90:15-27 (+A: stored/H_ph.f90)
//   sqrt(max(n,n')) g(a,b) sg v(S,n'),  T = c^+_a c_b S (up, then dw)   HxV_eph.f90:32-79
// for n' = n-1, n+1.  upT / dwT are the impurity hop tables of sector.cu (k_imphop_fill): entry
// (b*Norb+a) of row T is S = c^+_b c_a T with the sign of the pair (the bits between a and b are
// the same in S and T).  The dw hop changes the column: vfull holds every column of the slice.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_phonon(const double *__restrict__ v, const double *__restrict__ vfull, double *__restrict__ hv,
         int64_t nrow, int64_t ld, int64_t col_offset, int64_t slice, int64_t slice_full, int DimPh,
         double w0, double A, const uint8_t *__restrict__ impu, const uint8_t *__restrict__ impd,
         const int32_t *__restrict__ upT, int64_t ldu, const int32_t *__restrict__ dwT, int64_t ldd,
         int Norb, const double *__restrict__ g, int offdiag, double s_acc) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  const int n = blockIdx.z;
  if (i >= nrow) return;
  const int64_t cg = c + col_offset;
  const int mu = impu[i], md = impd[cg];
  double gd = A;
  for (int a = 0; a < Norb; a++) gd += g[a * Norb + a] * (double)(((mu >> a) & 1) + ((md >> a) & 1));
  double acc = w0 * (double)n * v[n * slice + c * ld + i];
  for (int dn = -1; dn <= 1; dn += 2) {
    const int n2 = n + dn;
    if (n2 < 0 || n2 >= DimPh) continue;
    const double *__restrict__ vs = v + n2 * slice;
    const double *__restrict__ vf = vfull + n2 * slice_full;
    double t = gd * vs[c * ld + i];
    if (offdiag) {
      for (int a = 0; a < Norb; a++)
        for (int b = 0; b < Norb; b++) {
          const double gab = g[a * Norb + b];
          if (a == b || gab == 0.0) continue;
          const int32_t eu = upT[(int64_t)(b * Norb + a) * ldu + i];
          if (eu != -1) t += (eu < 0 ? -gab : gab) * vs[c * ld + (eu & 0x7FFFFFFF)];
          const int32_t ed = dwT[(int64_t)(b * Norb + a) * ldd + cg];
          if (ed != -1) t += (ed < 0 ? -gab : gab) * vf[(int64_t)(ed & 0x7FFFFFFF) * ld + i];
        }
    }
    acc += sqrt((double)max(n, n2)) * t;
  }
  hv[n * slice + c * ld + i] += s_acc * acc;
}

int extra_setup(Engine &E) {
  Sector &S = E.sec;
  const int No = S.Norb;
  S.nsundry = 0;
  S.sundry = nullptr;
  S.gph = nullptr;
  S.DimPh = E.Nph + 1;
  S.w0_ph = E.w0_ph;
  S.A_ph = E.A_ph;
  S.eph_offdiag = false;
  if (!E.sundry_terms.empty()) {
    std::vector<SundryDev> dev;
    for (const edgpu_sundry_term &t : E.sundry_terms) {
      // application order c_l, cd_j, c_k, cd_i (HxV_sundry.f90:45-91)
      const int32_t *ops[4] = {t.c_l, t.cd_j, t.c_k, t.cd_i};
      SundryDev d = {};
      for (int k = 0; k < 4; k++) {
        if (ops[k][0] < 1 || ops[k][0] > No)
          return set_error("coulomb_sundry: orbital %d outside 1..Norb=%d", ops[k][0], No);
        d.bit[k] = (int8_t)(ops[k][0] - 1);
        d.dw[k] = (int8_t)(ops[k][1] == 2);
        d.create[k] = (int8_t)(k & 1);
      }
      d.U = t.U;
      dev.push_back(d);
    }
    S.nsundry = (int)dev.size();
    EDGPU_CUDA(cudaMalloc(&S.sundry, sizeof(SundryDev) * dev.size()));
    EDGPU_CUDA(cudaMemcpy(S.sundry, dev.data(), sizeof(SundryDev) * dev.size(), cudaMemcpyHostToDevice));
  }
  if (S.DimPh > 1) {
    std::vector<double> gg((size_t)No * No);
    for (int a = 0; a < No; a++)
      for (int b = 0; b < No; b++) {
        gg[(size_t)a * No + b] = E.g_ph[a][b];
        if (a != b && E.g_ph[a][b] != 0.0) S.eph_offdiag = true;
      }
    EDGPU_CUDA(cudaMalloc(&S.gph, sizeof(double) * gg.size()));
    EDGPU_CUDA(cudaMemcpy(S.gph, gg.data(), sizeof(double) * gg.size(), cudaMemcpyHostToDevice));
  }
  return 0;
}

int extra_hxv(Engine &E, const double *d_v, const double *d_vfull, double *d_hv, double s_acc) {
  Sector &S = E.sec;
  const int64_t slice = S.slice_len(), slice_full = S.up.ld * S.dw.dim;
  if (S.qdw <= 0) return 0;  // a rank without columns (DimDw < nranks)
  dim3 grid((unsigned)((S.up.dim + 127) / 128), (unsigned)S.qdw, (unsigned)S.DimPh);
  if (S.DimPh > 1) {
    k_phonon<<<grid, 128, 0, E.stream>>>(d_v, d_vfull, d_hv, S.up.dim, S.up.ld, S.d0, slice, slice_full,
                                         S.DimPh, S.w0_ph, S.A_ph, S.up.imp, S.dw.imp, S.up.imphop,
                                         S.up.ld, S.dw.imphop, S.dw.ld, S.Norb, S.gph,
                                         (int)S.eph_offdiag, s_acc);
    EDGPU_COUNT_LAUNCH();
    EDGPU_CUDA(cudaGetLastError());
  }
  if (S.nsundry) {
    k_sundry<<<grid, 128, 0, E.stream>>>(d_vfull, d_hv, S.up.dim, S.up.ld, S.d0, slice, slice_full,
                                         S.up.map, S.dw.map, rank_view(S.up.lin, S.up.ord),
                                         rank_view(S.dw.lin, S.dw.ord), S.sundry, S.nsundry, s_acc);
    EDGPU_COUNT_LAUNCH();
    EDGPU_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace edgpu
