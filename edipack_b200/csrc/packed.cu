// ed_mode=nonsu2 and ed_mode=superc on the device: sector map and stored Hamiltonian built by
// kernels, replacing
//   build_sector (nonsu2 branch)      ED_SECTOR.f90:335-368   m = iup + idw*2^Ns, popcount = Ntot
//   build_sector (superc branch)      ED_SECTOR.f90:244-281   m = iup + idw*2^Ns, popcnt(iup)-popcnt(idw) = Sz
//                                     (both: idw outer / iup inner -> ascending m)
//   build_Hv_sector_nonsu2 / _superc  ED_HAMILTONIAN_NONSU2.f90:31-130, ED_HAMILTONIAN_SUPERC.f90:31-140
//                                     (flat row split, remainder to the last rank)
//   ed_buildH_nonsu2_main             ED_HAMILTONIAN_NONSU2_STORED_HxV.f90:29-190 with the element
//                                     generators ED_NONSU2/stored/Himp.f90, Hint.f90, Hbath.f90,
//                                     Himp_bath.f90 (all four bath types; replica / general: same-spin
//                                     and spin-flip bath hops of Hbath.f90:49-133)
//   ed_buildH_superc_main             ED_HAMILTONIAN_SUPERC_STORED_HxV.f90:29-260 with
//                                     ED_SUPERC/stored/Himp.f90 (incl. anomalous local pairing :86-125),
//                                     Hint.f90, Hbath.f90 (bath pairing d :97-133; replica / general:
//                                     Nambu blocks of Hbath_tmp :29-92, :135-177), Himp_bath.f90
// The reference inserts element by element into a list of rows (`sp_insert_element`, O(row) search
// + realloc per element, ED_SPARSE_MATRIX.f90:346-357) after a recursive binary search of the
// target state.  Here one thread per row enumerates the same terms in the same order twice
// (count, then fill) straight into flat CSR arrays in HBM.  The target index is computed, not
// searched: nonsu2 sectors are all Ntot-subsets of 2*Ns bits in ascending integer order
// (combinadic rank); superc sectors are, for each dw integer in ascending order, the up integers
// with popcnt(idw)+Sz electrons (a prefix table over idw + the combinadic rank of iup).
// Fermionic signs are popc of a bit window over ALL lower bits (up and dw), as c/cdg do on the
// packed state (ED_AUX_FUNX.f90:334-384 called with pos+Ns, Himp.f90:62).
// The diagonal contributions are summed in the reference's order into one entry; off-diagonal
// entries that hit the same column stay separate entries (they add up in the product, like
// sp_insert_element's accumulation).  The product itself is the complex CSR SpMV of csr.cu.
#include <algorithm>
#include <cstring>

#include "edgpu_internal.cuh"

namespace edgpu {

enum { MODE_NONSU2 = 0, MODE_SUPERC = 1 };

struct Nonsu2Dev {
  edgpu_nonsu2_params p;  // superc sectors reuse the fields they share (hloc spin-diagonal)
  // superc only
  double anom[EDGPU_MAXORB][EDGPU_MAXORB][2];  // impHloc_anomalous(1,1,a,b)
  double pair_field[EDGPU_MAXORB];
  double bath_d[EDGPU_MAXORB][EDGPU_MAXBATH];  // dmft_bath%d(1,a,k)
  const int32_t *off;                          // [2^Ns + 1] first sector index of each idw
  // replica / general baths: Hbath_tmp(is,js,a,b,k) as [2][2][Norb][Norb][Nbath] (re,im) in global
  // memory (51 KB at the maximal sizes: does not fit beside the rest in constant memory)
  const double *hb;
  int32_t replica;
  int32_t binom[33][33];  // C(n,k), n,k <= 32 (entries that overflow int32 are never used)
  int32_t mode;
  int32_t ntot;   // nonsu2: electrons; superc: Sz
  int32_t nbits;  // 2*Ns
};
__constant__ Nonsu2Dev c_n2;

// combinadic rank of a bit pattern among the patterns with the same popcount (ascending order)
__device__ __forceinline__ int64_t comb_rank(uint32_t m) {
  int64_t r = 0;
  int k = 0;
  while (m) {
    const int p = __ffs(m) - 1;
    m &= m - 1;
    k++;
    r += c_n2.binom[p][k];
  }
  return r;
}
// pattern number r (ascending) among those of `nbits` bits with k ones
__device__ __forceinline__ uint32_t comb_unrank(int64_t r, int nbits, int k) {
  uint32_t m = 0;
  int p = nbits - 1;
  for (; k >= 1; k--) {
    while (c_n2.binom[p][k] > r) p--;
    m |= 1u << p;
    r -= c_n2.binom[p][k];
    p--;
  }
  return m;
}

// sector index (0-based) of the packed state m
__device__ __forceinline__ int64_t n2_rank(uint32_t m) {
  if (c_n2.mode == MODE_NONSU2) return comb_rank(m);
  const int Ns = c_n2.nbits >> 1;
  return (int64_t)c_n2.off[m >> Ns] + comb_rank(m & ((1u << Ns) - 1u));
}

__device__ __forceinline__ uint32_t n2_unrank(int64_t r) {
  if (c_n2.mode == MODE_NONSU2) return comb_unrank(r, c_n2.nbits, c_n2.ntot);
  const int Ns = c_n2.nbits >> 1;
  int lo = 0, hi = 1 << Ns;  // largest idw with off[idw] <= r among non-empty slots
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (c_n2.off[mid] <= r) lo = mid; else hi = mid;
  }
  const uint32_t iup = comb_unrank(r - c_n2.off[lo], Ns, __popc((uint32_t)lo) + c_n2.ntot);
  return iup | ((uint32_t)lo << Ns);
}

__global__ void __launch_bounds__(256) k_n2_map(int32_t *__restrict__ map, int64_t dim) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < dim) map[i] = (int32_t)n2_unrank(i);
}

struct CountSink {
  int n = 0;
  __device__ void emit(int64_t, double, double) { n++; }
};
struct FillSink {
  int32_t *cols;
  double2 *vals;
  int64_t k;
  __device__ void emit(int64_t col, double re, double im) {
    cols[k] = (int32_t)col;
    vals[k] = make_double2(re, im);
    k++;
  }
};

// direct (ED_SPARSE_H = F) product: the element is applied to the input vector instead of stored
struct ApplySink {
  const double2 *v;
  double are = 0.0, aim = 0.0;
  __device__ void emit(int64_t col, double re, double im) {
    const double2 x = v[col];
    are += re * x.x - im * x.y;
    aim += re * x.y + im * x.x;
  }
};

// c(beta) then cdg(alfa) on the packed state (bit positions 0-based); emits conjg(amp)*sg1*sg2
template <class Sink>
__device__ __forceinline__ void n2_hop(uint32_t m, int alfa, int beta, double are, double aim, Sink &s) {
  if (!((m >> beta) & 1u)) return;
  const uint32_t m1 = m & ~(1u << beta);
  if ((m1 >> alfa) & 1u) return;
  const int par = __popc(m & ((1u << beta) - 1u)) + __popc(m1 & ((1u << alfa) - 1u));
  const double sg = (par & 1) ? -1.0 : 1.0;
  s.emit(n2_rank(m1 | (1u << alfa)), are * sg, -aim * sg);
}

// four-operator chain  cdg(p4) cdg(p3) c(p2) c(p1)  applied right to left as in Hint.f90:73-76
template <class Sink>
__device__ __forceinline__ void n2_chain(uint32_t m, int p1, int p2, int p3, int p4, double amp, Sink &s) {
  int par = __popc(m & ((1u << p1) - 1u));
  m &= ~(1u << p1);
  par += __popc(m & ((1u << p2) - 1u));
  m &= ~(1u << p2);
  par += __popc(m & ((1u << p3) - 1u));
  m |= 1u << p3;
  par += __popc(m & ((1u << p4) - 1u));
  m |= 1u << p4;
  s.emit(n2_rank(m), (par & 1) ? -amp : amp, 0.0);
}

// pair annihilation c(p2) c(p1) (p1 applied first) / pair creation cdg(p2) cdg(p1); the caller
// checked the occupations (Himp.f90:90-121, Hbath.f90:101-131 of ED_SUPERC/stored)
template <class Sink>
__device__ __forceinline__ void n2_pair(uint32_t m, int p1, int p2, bool create, double re, double im, Sink &s) {
  int par = __popc(m & ((1u << p1) - 1u));
  m = create ? (m | (1u << p1)) : (m & ~(1u << p1));
  par += __popc(m & ((1u << p2) - 1u));
  m = create ? (m | (1u << p2)) : (m & ~(1u << p2));
  const double sg = (par & 1) ? -1.0 : 1.0;
  s.emit(n2_rank(m), re * sg, im * sg);
}

// hole-block hop of the Nambu bath (ED_SUPERC/stored/Hbath.f90:73-90): cdg(ibeta) then c(ialfa)
// on the dw bits; the caller checked ib(ibeta)==0, ib(ialfa)==1; emits conjg(amp)*sg1*sg2
template <class Sink>
__device__ __forceinline__ void n2_hole_hop(uint32_t m, int ialfa, int ibeta, double are, double aim, Sink &s) {
  int par = __popc(m & ((1u << ibeta) - 1u));
  m |= 1u << ibeta;
  par += __popc(m & ((1u << ialfa) - 1u));
  m &= ~(1u << ialfa);
  const double sg = (par & 1) ? -1.0 : 1.0;
  s.emit(n2_rank(m), are * sg, -aim * sg);
}

template <class Sink>
__device__ void n2_row(uint32_t m, int64_t i, Sink &s) {
  const edgpu_nonsu2_params &P = c_n2.p;
  const int Ns = P.Ns, No = P.Norb, Nb = P.Nbath;
  const bool replica = c_n2.replica != 0;
  const double *__restrict__ hbp = c_n2.hb;
#define HB(is, js, a, b, k, c) hbp[((((((is) * 2 + (js)) * No + (a)) * No + (b)) * Nb + (k)) << 1) + (c)]
#define HL(is, js, a, b, c) P.hloc[is][js][a][b][c]
  double nup[EDGPU_MAXORB], ndw[EDGPU_MAXORB];
#pragma unroll
  for (int a = 0; a < EDGPU_MAXORB; a++) {
    nup[a] = a < No ? (double)((m >> a) & 1u) : 0.0;
    ndw[a] = a < No ? (double)((m >> (a + Ns)) & 1u) : 0.0;
  }
  // ---- diagonal: Himp.f90:15-21, spin_field z :243-250, Hint.f90:13-52, Hbath.f90:12-27
  double dre = 0.0, dim_ = 0.0;
  for (int a = 0; a < No; a++) {
    dre += HL(0, 0, a, a, 0) * nup[a] + HL(1, 1, a, a, 0) * ndw[a] - P.xmu * (nup[a] + ndw[a]);
    dim_ += HL(0, 0, a, a, 1) * nup[a] + HL(1, 1, a, a, 1) * ndw[a];
  }
  const bool superc = c_n2.mode == MODE_SUPERC;
  bool any_sf = false;
  for (int a = 0; a < No; a++)
    any_sf = any_sf || P.spin_field[a][0] != 0.0 || P.spin_field[a][1] != 0.0 || P.spin_field[a][2] != 0.0;
  if (superc) any_sf = false;
  if (any_sf) {
    double h = 0.0;
    for (int a = 0; a < No; a++) h += P.spin_field[a][2] * (nup[a] - ndw[a]);
    dre += h;
  }
  {
    double h = 0.0;
    for (int a = 0; a < No; a++) h += P.Uloc[a] * nup[a] * ndw[a];
    for (int a = 0; a < No; a++)
      for (int b = a + 1; b < No; b++) {
        h += P.Ust[a][b] * (nup[a] * ndw[b] + nup[b] * ndw[a]);
        h += (P.Ust[a][b] - P.Jh[a][b]) * (nup[a] * nup[b] + ndw[a] * ndw[b]);
      }
    if (P.hfmode) {
      for (int a = 0; a < No; a++) h += -0.5 * P.Uloc[a] * (nup[a] + ndw[a]) + 0.25 * P.Uloc[a];
      for (int a = 0; a < No; a++)
        for (int b = a + 1; b < No; b++) {
          const double nn = nup[a] + ndw[a] + nup[b] + ndw[b];
          h += -0.5 * P.Ust[a][b] * nn + 0.5 * P.Ust[a][b];
          h += -0.5 * (P.Ust[a][b] - P.Jh[a][b]) * nn + 0.5 * (P.Ust[a][b] - P.Jh[a][b]);
        }
    }
    dre += h;
  }
  if (!replica) {
    double h = 0.0;
    for (int a = 0; a < P.Nfoo; a++)
      for (int k = 0; k < Nb; k++) {
        const int st = P.stride[a][k] - 1;
        h += P.bath_e[0][a][k] * (double)((m >> st) & 1u) + P.bath_e[1][a][k] * (double)((m >> (st + Ns)) & 1u);
      }
    dre += h;
  } else {
    // bath_diag(s,a,k) = Hbath_tmp(s,s,a,a,k) (real array in the reference).  superc: the hole block
    // enters with a minus sign (ED_SUPERC/stored/Hbath.f90:33-38)
    const bool sc = c_n2.mode == MODE_SUPERC;
    double h = 0.0;
    for (int k = 0; k < Nb; k++)
      for (int a = 0; a < No; a++) {
        const int st = P.stride[a][k] - 1;
        const double eu = HB(0, 0, a, a, k, 0), ed = HB(1, 1, a, a, k, 0);
        h += eu * (double)((m >> st) & 1u) + (sc ? -ed : ed) * (double)((m >> (st + Ns)) & 1u);
      }
    dre += h;
  }
  s.emit(i, dre, dim_);
  // ---- Himp.f90: same-spin inter-orbital hops :38-80
  for (int a = 0; a < No; a++)
    for (int b = 0; b < No; b++) {
      if (a == b) continue;  // "nup(jorb)==1 .AND. nup(iorb)==0" (Himp.f90:45) never holds for a==b
      if (HL(0, 0, a, b, 0) != 0.0 || HL(0, 0, a, b, 1) != 0.0) n2_hop(m, a, b, HL(0, 0, a, b, 0), HL(0, 0, a, b, 1), s);
      if (HL(1, 1, a, b, 0) != 0.0 || HL(1, 1, a, b, 1) != 0.0)
        n2_hop(m, a + Ns, b + Ns, HL(1, 1, a, b, 0), HL(1, 1, a, b, 1), s);
    }
  if (superc) {
    // anomalous local pairing: impHloc_anomalous + pair_field (ED_SUPERC/stored/Himp.f90:86-125)
    bool any_an = false;
    for (int a = 0; a < No; a++) {
      any_an = any_an || c_n2.pair_field[a] != 0.0;
      for (int b = 0; b < No; b++) any_an = any_an || c_n2.anom[a][b][0] != 0.0 || c_n2.anom[a][b][1] != 0.0;
    }
    if (any_an)
      for (int a = 0; a < No; a++)
        for (int b = 0; b < No; b++) {
          const double pf = a == b ? c_n2.pair_field[a] : 0.0;
          const bool ua = (m >> a) & 1u, db = (m >> (b + Ns)) & 1u;
          if (ua && db) n2_pair(m, a, b + Ns, false, c_n2.anom[a][b][0] + pf, c_n2.anom[a][b][1], s);
          if (!ua && !db) n2_pair(m, b + Ns, a, true, c_n2.anom[a][b][0] + pf, -c_n2.anom[a][b][1], s);
        }
  }
  // spin-flip local terms :85-110 (nonsu2 only)
  for (int is = 0; is < 2 && !superc; is++) {
    const int js = 1 - is;
    for (int a = 0; a < No; a++)
      for (int b = 0; b < No; b++)
        if (HL(is, js, a, b, 0) != 0.0 || HL(is, js, a, b, 1) != 0.0)
          n2_hop(m, a + is * Ns, b + js * Ns, HL(is, js, a, b, 0), HL(is, js, a, b, 1), s);
  }
  // spin_field x / y :252-300: F_x (c+_dw c_up + c+_up c_dw) and -/+ i F_y
  if (any_sf) {
    for (int a = 0; a < No; a++) {
      const double fx = P.spin_field[a][0], fy = P.spin_field[a][1];
      // src up -> dst dw carries (F_x - i F_y); src dw -> dst up carries (F_x + i F_y).  n2_hop
      // emits conjg(amp): pass the conjugates so that the inserted value is the amplitude itself
      n2_hop(m, a + Ns, a, fx, +fy, s);
      n2_hop(m, a, a + Ns, fx, -fy, s);
    }
  }
  // ---- Hint.f90: spin-exchange :63-90 and pair-hopping :96-124
  bool any_jx = false, any_jp = false;
  for (int a = 0; a < No; a++)
    for (int b = 0; b < No; b++) {
      any_jx = any_jx || P.Jx[a][b] != 0.0;
      any_jp = any_jp || P.Jp[a][b] != 0.0;
    }
  if (No > 1 && any_jx)
    for (int a = 0; a < No; a++)
      for (int b = 0; b < No; b++)
        if (a != b && ((m >> b) & 1u) && ((m >> (a + Ns)) & 1u) && !((m >> (b + Ns)) & 1u) && !((m >> a) & 1u))
          n2_chain(m, b, a + Ns, b + Ns, a, P.Jx[a][b], s);
  if (No > 1 && any_jp)
    for (int a = 0; a < No; a++)
      for (int b = 0; b < No; b++)
        if (a != b && ((m >> b) & 1u) && ((m >> (b + Ns)) & 1u) && !((m >> (a + Ns)) & 1u) && !((m >> a) & 1u))
          n2_chain(m, b, b + Ns, a + Ns, a, P.Jp[a][b], s);
  if (replica && !superc) {
    // ED_NONSU2/stored/Hbath.f90:49-100 same-spin bath hops, :103-133 spin-flip bath hops
    for (int k = 0; k < Nb; k++)
      for (int a = 0; a < No; a++)
        for (int b = 0; b < No; b++) {
          if (a == b) continue;  // ib(ibeta)==1 .AND. ib(ialfa)==0 never holds on one site
          const int ia = P.stride[a][k] - 1, ib_ = P.stride[b][k] - 1;
          if (HB(0, 0, a, b, k, 0) != 0.0 || HB(0, 0, a, b, k, 1) != 0.0)
            n2_hop(m, ia, ib_, HB(0, 0, a, b, k, 0), HB(0, 0, a, b, k, 1), s);
          if (HB(1, 1, a, b, k, 0) != 0.0 || HB(1, 1, a, b, k, 1) != 0.0)
            n2_hop(m, ia + Ns, ib_ + Ns, HB(1, 1, a, b, k, 0), HB(1, 1, a, b, k, 1), s);
        }
    for (int k = 0; k < Nb; k++)
      for (int is = 0; is < 2; is++) {
        const int js = 1 - is;
        for (int a = 0; a < No; a++)
          for (int b = 0; b < No; b++)
            if (HB(is, js, a, b, k, 0) != 0.0 || HB(is, js, a, b, k, 1) != 0.0)
              n2_hop(m, P.stride[a][k] - 1 + is * Ns, P.stride[b][k] - 1 + js * Ns, HB(is, js, a, b, k, 0),
                     HB(is, js, a, b, k, 1), s);
      }
  }
  if (replica && superc) {
    // ED_SUPERC/stored/Hbath.f90:48-92: particle block (1,1) on the up bits, hole block (2,2) on
    // the dw bits (creation first); :135-177 anomalous blocks (1,2) pair creation, (2,1) annihilation
    for (int k = 0; k < Nb; k++)
      for (int a = 0; a < No; a++)
        for (int b = 0; b < No; b++) {
          const int ia = P.stride[a][k] - 1, ib_ = P.stride[b][k] - 1;
          if (a != b && (HB(0, 0, a, b, k, 0) != 0.0 || HB(0, 0, a, b, k, 1) != 0.0))
            n2_hop(m, ia, ib_, HB(0, 0, a, b, k, 0), HB(0, 0, a, b, k, 1), s);
          if (a != b && (HB(1, 1, a, b, k, 0) != 0.0 || HB(1, 1, a, b, k, 1) != 0.0) &&
              !((m >> (ib_ + Ns)) & 1u) && ((m >> (ia + Ns)) & 1u))
            n2_hole_hop(m, ia + Ns, ib_ + Ns, HB(1, 1, a, b, k, 0), HB(1, 1, a, b, k, 1), s);
        }
    for (int k = 0; k < Nb; k++)
      for (int a = 0; a < No; a++)
        for (int b = 0; b < No; b++) {
          const int ua = P.stride[a][k] - 1, db = P.stride[b][k] - 1 + Ns;  // (1,2): up a, dw b
          if ((HB(0, 1, a, b, k, 0) != 0.0 || HB(0, 1, a, b, k, 1) != 0.0) && !((m >> db) & 1u) && !((m >> ua) & 1u))
            n2_pair(m, db, ua, true, HB(0, 1, a, b, k, 0), -HB(0, 1, a, b, k, 1), s);
          const int da = P.stride[a][k] - 1 + Ns, ub = P.stride[b][k] - 1;  // (2,1): dw a, up b
          if ((HB(1, 0, a, b, k, 0) != 0.0 || HB(1, 0, a, b, k, 1) != 0.0) && ((m >> ub) & 1u) && ((m >> da) & 1u))
            n2_pair(m, ub, da, false, HB(1, 0, a, b, k, 0), -HB(1, 0, a, b, k, 1), s);
        }
  }
  if (superc && !replica) {
    // bath pairing Delta_l (c_dw c_up + h.c.) on every bath level (ED_SUPERC/stored/Hbath.f90:97-133)
    for (int a = 0; a < P.Nfoo; a++)
      for (int k = 0; k < Nb; k++) {
        const int ms = P.stride[a][k] - 1;
        const double d = c_n2.bath_d[a][k];
        if (d == 0.0) continue;
        const bool u = (m >> ms) & 1u, w = (m >> (ms + Ns)) & 1u;
        if (u && w) n2_pair(m, ms, ms + Ns, false, d, 0.0, s);
        if (!u && !w) n2_pair(m, ms + Ns, ms, true, d, 0.0, s);
      }
  }
  // ---- Himp_bath.f90: spin-conserving hybridisation :10-67
  for (int a = 0; a < No; a++)
    for (int k = 0; k < Nb; k++) {
      const int ms = P.stride[a][k] - 1;
      for (int sp = 0; sp < 2; sp++) {
        const double v = P.bath_v[sp][a][k];
        if (v != 0.0) {
          n2_hop(m, ms + sp * Ns, a + sp * Ns, v, 0.0, s);  // imp -> bath
          n2_hop(m, a + sp * Ns, ms + sp * Ns, v, 0.0, s);  // bath -> imp
        }
      }
    }
  // spin-flip hybridisation u :72-136 (nonsu2 with a normal / hybrid bath only)
  for (int a = 0; a < No && !superc && !replica; a++)
    for (int k = 0; k < Nb; k++) {
      const int ms = P.stride[a][k] - 1;
      const double u1 = P.bath_u[0][a][k], u2 = P.bath_u[1][a][k];
      if (u1 != 0.0) {  // imp up <-> bath dw
        n2_hop(m, ms + Ns, a, u1, 0.0, s);
        n2_hop(m, a, ms + Ns, u1, 0.0, s);
      }
      if (u2 != 0.0) {  // imp dw <-> bath up
        n2_hop(m, ms, a + Ns, u2, 0.0, s);
        n2_hop(m, a + Ns, ms, u2, 0.0, s);
      }
    }
#undef HL
#undef HB
}

__global__ void __launch_bounds__(128)
k_n2_count(const int32_t *__restrict__ map, int64_t row0, int64_t nloc, int32_t *__restrict__ cnt) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  CountSink s;
  n2_row((uint32_t)map[row0 + r], row0 + r, s);
  cnt[r] = s.n;
}

__global__ void __launch_bounds__(128)
k_n2_fill(const int32_t *__restrict__ map, int64_t row0, int64_t nloc, const int64_t *__restrict__ rowptr,
          int32_t *__restrict__ cols, double2 *__restrict__ vals) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  FillSink s{cols, vals, rowptr[r]};
  n2_row((uint32_t)map[row0 + r], row0 + r, s);
}

// Direct on-the-fly H x v of the packed-state modes (directMatVec_nonsu2_main / _superc_main,
// ED_HAMILTONIAN_NONSU2_DIRECT_HxV.f90:22-252, ED_HAMILTONIAN_SUPERC_DIRECT_HxV.f90:22-311, with
// direct/HxVimp.f90, HxVint.f90, HxVbath.f90, HxVimp_bath.f90): the row generator of the stored
// path applied to the (all-gathered) input vector instead of filling a CSR -- the same terms in the
// same order, so the matrix is the STORED path's wherever the reference's two paths disagree
// (DESIGN.md "reference quirks").  Nothing but the sector map is kept in HBM (the stored form
// costs 20 B per element, ~55 elements per row at BASELINE config 5).
__global__ void __launch_bounds__(128)
k_n2_direct(const int32_t *__restrict__ map, int64_t row0, int64_t nloc, const double2 *__restrict__ vin,
            double2 *__restrict__ hv, int accum, double s_acc, double s_old) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  ApplySink s{vin};
  n2_row((uint32_t)map[row0 + r], row0 + r, s);
  double2 o = make_double2(s_acc * s.are, s_acc * s.aim);
  if (accum) {
    const double2 h = hv[r];
    o.x += s_old * h.x;
    o.y += s_old * h.y;
  }
  hv[r] = o;
}

int packed_direct_hxv(Engine &E, const double *d_vin_full, double *d_hv, bool accum, double s_acc, double s_old) {
  const CsrSector &C = E.csr;
  if (C.nloc <= 0) return 0;
  k_n2_direct<<<(unsigned)((C.nloc + 127) / 128), 128, 0, E.stream>>>(
      C.map, C.row0, C.nloc, (const double2 *)d_vin_full, (double2 *)d_hv, (int)accum, s_acc, s_old);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

static Nonsu2Dev g_host_dev;  // host copy of the open sector's constants (~17 KB: not on the stack)

// mode/quantum number are in h; builds map + CSR of this rank's rows and opens the stored-H sector
static int packed_open(Engine &E, Nonsu2Dev &h, const char *who) {
  if (!E.inited) return set_error("edgpu_init was not called");
  if (E.sec.open) return set_error("close the direct-H sector before opening a %s one", who);
  if (E.csr.open) csr_close(E);
  const edgpu_nonsu2_params *p = &h.p;
  const int nbits = 2 * p->Ns, Ns = p->Ns;
  if (p->Ns < 1 || nbits > 31) return set_error("%s: 2*Ns = %d exceeds the 31-bit packed state", who, nbits);
  if (p->Norb < 1 || p->Norb > EDGPU_MAXORB || p->Nbath < 0 || p->Nbath > EDGPU_MAXBATH)
    return set_error("%s: Norb/Nbath out of range", who);
  if (E.Nph > 0)
    return set_error("%s: phonons (H_ph / H_e_ph of the packed-state modes) are not generated on the device: "
                     "hand the host-built spH0 to edgpu_csr_open_z or call edgpu_set_phonons(0,...)", who);
  h.replica = (p->bath_type == EDGPU_BATH_REPLICA || p->bath_type == EDGPU_BATH_GENERAL);
  h.hb = nullptr;
  double *d_hb = nullptr;
  if (h.replica) {
    if (E.hbath_packed.empty() || E.hb_Norb != p->Norb || E.hb_Nbath != p->Nbath)
      return set_error("%s: replica / general bath needs Hbath_tmp for Norb=%d, Nbath=%d "
                       "(edgpu_set_hbath_packed)", who, p->Norb, p->Nbath);
  }
  h.nbits = nbits;
  for (int n = 0; n <= 32; n++)
    for (int k = 0; k <= 32; k++) {
      int64_t c = k > n ? 0 : (k == 0 || k == n ? 1 : (int64_t)h.binom[n - 1][k - 1] + h.binom[n - 1][k]);
      h.binom[n][k] = (int32_t)std::min<int64_t>(c, INT32_MAX);
    }
  int64_t dim = 0;
  int32_t *d_off = nullptr;
  if (h.mode == MODE_NONSU2) {
    if (h.ntot < 0 || h.ntot > nbits) return set_error("nonsu2: Ntot = %d outside [0, %d]", h.ntot, nbits);
    dim = host_binomial(nbits, h.ntot);
  } else {
    if (h.ntot < -Ns || h.ntot > Ns) return set_error("superc: Sz = %d outside [%d, %d]", h.ntot, -Ns, Ns);
    // first sector index of every dw integer (ED_SECTOR.f90:262-281: idw outer, iup inner)
    std::vector<int32_t> off(((size_t)1 << Ns) + 1, 0);
    for (uint32_t idw = 0; idw < (1u << Ns); idw++) {
      const int k = __builtin_popcount(idw) + h.ntot;
      const int64_t c = (k < 0 || k > Ns) ? 0 : host_binomial(Ns, k);
      dim += c;
      if (dim > INT32_MAX) return set_error("superc: sector dimension exceeds 32-bit columns");
      off[(size_t)idw + 1] = (int32_t)dim;
    }
    EDGPU_CUDA(cudaMalloc(&d_off, sizeof(int32_t) * off.size()));
    EDGPU_CUDA(cudaMemcpy(d_off, off.data(), sizeof(int32_t) * off.size(), cudaMemcpyHostToDevice));
  }
  h.off = d_off;
  if (dim > INT32_MAX) {
    cudaFree(d_off);
    return set_error("%s: sector dimension %lld exceeds 32-bit columns", who, (long long)dim);
  }
  if (h.replica) {
    EDGPU_CUDA(cudaMalloc(&d_hb, sizeof(double) * E.hbath_packed.size()));
    EDGPU_CUDA(cudaMemcpy(d_hb, E.hbath_packed.data(), sizeof(double) * E.hbath_packed.size(),
                          cudaMemcpyHostToDevice));
    h.hb = d_hb;
  }
  EDGPU_CUDA(cudaMemcpyToSymbolAsync(c_n2, &h, sizeof(h), 0, cudaMemcpyHostToDevice, E.stream));
  // row split MpiQ = Dim/P, remainder to the last rank (ED_HAMILTONIAN_NONSU2.f90:72-79,
  // ED_HAMILTONIAN_SUPERC.f90:76-88)
  const int P = E.nranks;
  const int64_t q = dim / P;
  const int64_t row0 = q * E.rank, nloc = q + (E.rank == P - 1 ? dim % P : 0);
  int32_t *d_map = nullptr, *d_cnt = nullptr, *d_cols = nullptr;
  int64_t *d_rowptr = nullptr;
  double *d_vals = nullptr;
  auto fail = [&](int rc) {
    cudaFree(d_map);
    cudaFree(d_cnt);
    cudaFree(d_cols);
    cudaFree(d_rowptr);
    cudaFree(d_vals);
    cudaFree(d_off);
    cudaFree(d_hb);
    return rc;
  };
#define N2_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__)); \
  } while (0)
  N2_CUDA(cudaMalloc(&d_map, sizeof(int32_t) * std::max<int64_t>(dim, 1)));
  if (dim > 0) {
    k_n2_map<<<(unsigned)((dim + 255) / 256), 256, 0, E.stream>>>(d_map, dim);
    EDGPU_COUNT_LAUNCH();
  }
  if (!E.sparse_h) {
    // ED_SPARSE_H = F: no stored matrix; the ranking tables (off, Hbath_tmp) live as long as the sector
    N2_CUDA(cudaStreamSynchronize(E.stream));
    int rc = csr_adopt_device(E, true, nloc, dim, row0, nullptr, nullptr, nullptr, 0, d_map);
    if (rc) return fail(rc);
    E.csr.direct = true;
    E.csr.pk_off = d_off;
    E.csr.pk_hb = d_hb;
    E.csr.pk_mode = h.mode;
    E.csr.pk_qn = h.ntot;
    E.csr.pk_Ns = Ns;
    return 0;
  }
  N2_CUDA(cudaMalloc(&d_cnt, sizeof(int32_t) * std::max<int64_t>(nloc, 1)));
  const unsigned grid = (unsigned)std::max<int64_t>(1, (nloc + 127) / 128);
  k_n2_count<<<grid, 128, 0, E.stream>>>(d_map, row0, nloc, d_cnt);
  EDGPU_COUNT_LAUNCH();
  N2_CUDA(cudaGetLastError());
  std::vector<int32_t> cnt((size_t)std::max<int64_t>(nloc, 1));
  N2_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int32_t) * nloc, cudaMemcpyDeviceToHost, E.stream));
  N2_CUDA(cudaStreamSynchronize(E.stream));
  std::vector<int64_t> rowptr((size_t)nloc + 1, 0);
  for (int64_t r = 0; r < nloc; r++) rowptr[(size_t)r + 1] = rowptr[(size_t)r] + cnt[(size_t)r];
  const int64_t nnz = rowptr[(size_t)nloc];
  N2_CUDA(cudaMalloc(&d_rowptr, sizeof(int64_t) * (nloc + 1)));
  N2_CUDA(cudaMalloc(&d_cols, sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
  N2_CUDA(cudaMalloc(&d_vals, sizeof(double) * 2 * std::max<int64_t>(nnz, 1)));
  N2_CUDA(cudaMemcpyAsync(d_rowptr, rowptr.data(), sizeof(int64_t) * (nloc + 1), cudaMemcpyHostToDevice, E.stream));
  k_n2_fill<<<grid, 128, 0, E.stream>>>(d_map, row0, nloc, d_rowptr, d_cols, (double2 *)d_vals);
  EDGPU_COUNT_LAUNCH();
  N2_CUDA(cudaGetLastError());
  N2_CUDA(cudaStreamSynchronize(E.stream));
#undef N2_CUDA
  cudaFree(d_cnt);
  d_cnt = nullptr;
  cudaFree(d_off);  // only the builder kernels rank states
  d_off = nullptr;
  cudaFree(d_hb);   // ... and read Hbath_tmp
  d_hb = nullptr;
  int rc = csr_adopt_device(E, true, nloc, dim, row0, d_rowptr, d_cols, d_vals, nnz, d_map);
  if (rc) return fail(rc);
  E.csr.pk_mode = h.mode;
  E.csr.pk_qn = h.ntot;
  E.csr.pk_Ns = Ns;
  return 0;
}

int64_t packed_sector_dim(int mode, int Ns, int qn) {
  if (mode == MODE_NONSU2) return host_binomial(2 * Ns, qn);
  int64_t dim = 0;
  for (int k = 0; k <= Ns; k++)
    if (k - qn >= 0 && k - qn <= Ns) dim += host_binomial(Ns, k) * host_binomial(Ns, k - qn);
  return dim;
}

// ---------------------------------------------------------------------------------------------
// apply_COps / apply_op_C / apply_op_CDG on packed states (ED_SECTOR.f90:465-1140, nonsu2 and
// superc branches), gather form on the target sector: out(j) = sum_k coef_k sgn_k V(i_k) with
// |j> = O_k |i_k>.  The source index is ranked, not searched.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_pk_apply(const int32_t *__restrict__ map_t, int64_t row0, int64_t nloc, const double2 *__restrict__ vsrc,
           double2 *__restrict__ out, PackedOps ops, int src_mode, const int32_t *__restrict__ src_off,
           int Ns) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nloc) return;
  const uint32_t m = (uint32_t)map_t[row0 + i];
  double2 acc = make_double2(0.0, 0.0);
  for (int k = 0; k < ops.n; k++) {
    const int bit = ops.bit[k];
    const bool occ = (m >> bit) & 1u;
    if (ops.create[k] ? !occ : occ) continue;  // c^+ leaves the bit set, c leaves it empty
    const uint32_t ms = m ^ (1u << bit);
    const double sg = (__popc(ms & ((1u << bit) - 1u)) & 1) ? -1.0 : 1.0;
    const int64_t idx = src_mode == MODE_NONSU2
                            ? comb_rank(ms)
                            : (int64_t)src_off[ms >> Ns] + comb_rank(ms & ((1u << Ns) - 1u));
    const double2 x = vsrc[idx];
    acc.x += sg * (ops.cre[k] * x.x - ops.cim[k] * x.y);
    acc.y += sg * (ops.cre[k] * x.y + ops.cim[k] * x.x);
  }
  out[i] = acc;
}

int packed_apply_ops(Engine &E, const PackedOps &ops, int src_mode, int src_qn, const double *d_vsrc_full,
                     double *d_out) {
  CsrSector &C = E.csr;
  if (!C.open || C.pk_mode < 0 || !C.map) return set_error("apply_ops: no device-built nonsu2/superc sector open");
  const int Ns = C.pk_Ns;
  int32_t *d_off = nullptr;
  if (src_mode == MODE_SUPERC) {
    std::vector<int32_t> off(((size_t)1 << Ns) + 1, 0);
    int64_t dim = 0;
    for (uint32_t idw = 0; idw < (1u << Ns); idw++) {
      const int k = __builtin_popcount(idw) + src_qn;
      dim += (k < 0 || k > Ns) ? 0 : host_binomial(Ns, k);
      off[(size_t)idw + 1] = (int32_t)dim;
    }
    EDGPU_CUDA(cudaMalloc(&d_off, sizeof(int32_t) * off.size()));
    EDGPU_CUDA(cudaMemcpyAsync(d_off, off.data(), sizeof(int32_t) * off.size(), cudaMemcpyHostToDevice, E.stream));
    EDGPU_CUDA(cudaStreamSynchronize(E.stream));  // `off` is a local
  }
  EDGPU_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * C.padded_len(), E.stream));
  if (C.nloc > 0) {
    k_pk_apply<<<(unsigned)((C.nloc + 127) / 128), 128, 0, E.stream>>>(
        C.map, C.row0, C.nloc, (const double2 *)d_vsrc_full, (double2 *)d_out, ops, src_mode, d_off, Ns);
    EDGPU_COUNT_LAUNCH();
  }
  EDGPU_CUDA(cudaGetLastError());
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  cudaFree(d_off);
  return 0;
}

__global__ void __launch_bounds__(128)
k_pk_twin(const int32_t *__restrict__ map_t, int64_t row0, int64_t nloc, const double2 *__restrict__ vsrc,
          double2 *__restrict__ out, int src_mode, const int32_t *__restrict__ src_off, int Ns) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nloc) return;
  const uint32_t m = (uint32_t)map_t[row0 + i];
  const uint32_t lo = (1u << Ns) - 1u;
  // flip_state_other (ED_SECTOR.f90:1797-1817) is an involution: the source state is flip(m)
  const uint32_t ms = src_mode == MODE_NONSU2 ? (~m & ((1u << (2 * Ns)) - 1u)) : ((m >> Ns) | ((m & lo) << Ns));
  const int64_t idx = src_mode == MODE_NONSU2 ? comb_rank(ms) : (int64_t)src_off[ms >> Ns] + comb_rank(ms & lo);
  out[i] = vsrc[idx];
}

int packed_twin(Engine &E, int src_mode, int src_qn, const double *d_vsrc_full, double *d_out) {
  CsrSector &C = E.csr;
  if (!C.open || C.pk_mode < 0 || !C.map) return set_error("state_twin: no device-built nonsu2/superc sector open");
  const int Ns = C.pk_Ns;
  int32_t *d_off = nullptr;
  if (src_mode == MODE_SUPERC) {
    std::vector<int32_t> off(((size_t)1 << Ns) + 1, 0);
    int64_t dim = 0;
    for (uint32_t idw = 0; idw < (1u << Ns); idw++) {
      const int k = __builtin_popcount(idw) + src_qn;
      dim += (k < 0 || k > Ns) ? 0 : host_binomial(Ns, k);
      off[(size_t)idw + 1] = (int32_t)dim;
    }
    EDGPU_CUDA(cudaMalloc(&d_off, sizeof(int32_t) * off.size()));
    EDGPU_CUDA(cudaMemcpy(d_off, off.data(), sizeof(int32_t) * off.size(), cudaMemcpyHostToDevice));
  }
  EDGPU_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * C.padded_len(), E.stream));
  if (C.nloc > 0) {
    k_pk_twin<<<(unsigned)((C.nloc + 127) / 128), 128, 0, E.stream>>>(
        C.map, C.row0, C.nloc, (const double2 *)d_vsrc_full, (double2 *)d_out, src_mode, d_off, Ns);
    EDGPU_COUNT_LAUNCH();
  }
  EDGPU_CUDA(cudaGetLastError());
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  cudaFree(d_off);
  return 0;
}

// dens / docc of a packed-state vector (ED_OBSERVABLES_NONSU2.f90 / _SUPERC.f90:150-165)
__global__ void __launch_bounds__(256)
k_pk_observables(const int32_t *__restrict__ map, int64_t row0, int64_t nloc, const double2 *__restrict__ v,
                 int Ns, int Norb, double *__restrict__ out) {
  double acc[2 * EDGPU_MAXORB];
#pragma unroll
  for (int k = 0; k < 2 * EDGPU_MAXORB; k++) acc[k] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nloc; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t m = (uint32_t)map[row0 + i];
    const double2 x = v[i];
    const double w = x.x * x.x + x.y * x.y;
#pragma unroll
    for (int a = 0; a < EDGPU_MAXORB; a++)
      if (a < Norb) {
        const int nu = (m >> a) & 1u, nd = (m >> (a + Ns)) & 1u;
        acc[a] += w * (nu + nd);
        acc[EDGPU_MAXORB + a] += w * (nu * nd);
      }
  }
#pragma unroll
  for (int k = 0; k < 2 * EDGPU_MAXORB; k++) {
    double x = acc[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(out + k, x);
  }
}

int packed_observables(Engine &E, const double *d_vec, double *h_dens, double *h_docc) {
  CsrSector &C = E.csr;
  if (!C.open || C.pk_mode < 0 || !C.map) return set_error("observables: no device-built nonsu2/superc sector open");
  double *d_out = E.d_scal + 8;
  EDGPU_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * 2 * EDGPU_MAXORB, E.stream));
  const int Norb = g_host_dev.p.Norb;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((C.nloc + 255) / 256, 1024));
  k_pk_observables<<<grid, 256, 0, E.stream>>>(C.map, C.row0, C.nloc, (const double2 *)d_vec, C.pk_Ns, Norb, d_out);
  EDGPU_COUNT_LAUNCH();
  EDGPU_TRY(comm_allreduce_sum(E, d_out, 2 * EDGPU_MAXORB));
  EDGPU_CUDA(cudaMemcpyAsync(E.h_scal + 8, d_out, sizeof(double) * 2 * EDGPU_MAXORB, cudaMemcpyDeviceToHost, E.stream));
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  for (int a = 0; a < Norb; a++) {
    h_dens[a] = E.h_scal[8 + a];
    h_docc[a] = E.h_scal[8 + EDGPU_MAXORB + a];
  }
  return 0;
}

int nonsu2_open(Engine &E, const edgpu_nonsu2_params *p, int ntot) {
  Nonsu2Dev &h = g_host_dev;
  h = Nonsu2Dev();
  h.p = *p;
  h.mode = MODE_NONSU2;
  h.ntot = ntot;
  return packed_open(E, h, "nonsu2");
}

int superc_open(Engine &E, const edgpu_superc_params *p, int sz) {
  Nonsu2Dev &h = g_host_dev;
  h = Nonsu2Dev();
  edgpu_nonsu2_params &q = h.p;
  q.Ns = p->Ns;
  q.Norb = p->Norb;
  q.Nbath = p->Nbath;
  q.bath_type = p->bath_type;
  q.hfmode = p->hfmode;
  q.Nfoo = p->Nfoo;
  q.xmu = p->xmu;
  for (int s = 0; s < 2; s++)
    for (int a = 0; a < EDGPU_MAXORB; a++)
      for (int b = 0; b < EDGPU_MAXORB; b++)
        for (int c = 0; c < 2; c++) q.hloc[s][s][a][b][c] = p->hloc[s][a][b][c];
  memcpy(q.Uloc, p->Uloc, sizeof(q.Uloc));
  memcpy(q.Ust, p->Ust, sizeof(q.Ust));
  memcpy(q.Jh, p->Jh, sizeof(q.Jh));
  memcpy(q.Jx, p->Jx, sizeof(q.Jx));
  memcpy(q.Jp, p->Jp, sizeof(q.Jp));
  memcpy(q.bath_e, p->bath_e, sizeof(q.bath_e));
  memcpy(q.bath_v, p->bath_v, sizeof(q.bath_v));
  memcpy(q.stride, p->stride, sizeof(q.stride));
  memcpy(h.anom, p->hloc_anomalous, sizeof(h.anom));
  memcpy(h.pair_field, p->pair_field, sizeof(h.pair_field));
  memcpy(h.bath_d, p->bath_d, sizeof(h.bath_d));
  h.mode = MODE_SUPERC;
  h.ntot = sz;
  return packed_open(E, h, "superc");
}

}  // namespace edgpu
