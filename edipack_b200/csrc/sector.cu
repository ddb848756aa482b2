// Device-resident sector set-up: Fock maps, ranking tables, diagonal tables, hop tables.
//
// Replaces, on the device, what the reference rebuilds on the host for every sector:
//   build_sector                      ED_SECTOR.f90:165-242   (maps, ascending Fock order)
//   binary_search per matrix element  ED_AUX_FUNX.f90:463-480 (-> two-table ranking)
//   c / cdg sign loops                ED_AUX_FUNX.f90:334-384 (-> popc of a bit window)
//   the per-row term scans of direct/HxV_up.f90, HxV_dw.f90   (-> ELL hop tables, built once)
//   direct/HxV_local.f90              (-> eps_up[iup] + eps_dw[idw] + X[imp_dw][imp_up])
#include "edgpu_internal.cuh"

#include <algorithm>
#include <cstring>

namespace edgpu {

int64_t host_binomial(int n, int k) {
  if (k < 0 || k > n) return 0;
  if (k > n - k) k = n - k;
  int64_t r = 1;
  for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
  return r;
}

// First (n mod P) ranks get one extra element: the dw split of
// ED_HAMILTONIAN_NORMAL.f90:128-142 and the row split of vector_transpose_MPI
// (ED_HAMILTONIAN_NORMAL_COMMON.f90:104-112).
void block_split(int64_t n, int P, int r, int64_t *q, int64_t *start) {
  int64_t base = n / P, rem = n % P;
  *q = base + (r < rem ? 1 : 0);
  *start = r * base + (r < rem ? r : rem);
}

__constant__ int32_t c_binom[33][33];

// r-th (0-based) Ns-bit pattern with nel bits set in ascending order of the PERMUTED integer
// (combinadic unranking), mapped back to the original bit order.  With the identity order this
// is the reference's popcount scan order (ED_SECTOR.f90:217-242).
__global__ void k_build_map(int32_t *__restrict__ map, int32_t *__restrict__ mapp, int64_t dim,
                            int Ns, int nel, SiteOrder ord) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= dim) return;
  int64_t rem = r;
  int k = nel;
  uint32_t mp = 0, m = 0;
  for (int pos = Ns - 1; pos >= 0 && k > 0; --pos) {
    int64_t c = c_binom[pos][k];
    if (rem >= c) {
      mp |= (1u << pos);
      m |= (1u << ord.site[pos]);
      rem -= c;
      --k;
    }
  }
  map[r] = (int32_t)m;
  mapp[r] = (int32_t)mp;
}

// reference index of internal state r = combinadic rank of its Fock integer
__global__ void k_ref_rank(const int32_t *__restrict__ map, int64_t dim, int Ns,
                           int32_t *__restrict__ refidx) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= dim) return;
  const uint32_t m = (uint32_t)map[r];
  int j = 0;
  int64_t rk = 0;
  for (int b = 0; b < Ns; b++)
    if ((m >> b) & 1u) rk += c_binom[b][++j];
  refidx[r] = (int32_t)rk;
}

__global__ void k_lin_ja(const int32_t *__restrict__ mapp, int64_t dim, int lo_bits,
                         int32_t *__restrict__ ja) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= dim) return;
  uint32_t hi = (uint32_t)mapp[r] >> lo_bits;
  if (r == 0 || ((uint32_t)mapp[r - 1] >> lo_bits) != hi) ja[hi] = (int32_t)r;
}
__global__ void k_lin_jb(const int32_t *__restrict__ mapp, int64_t dim, int lo_bits,
                         const int32_t *__restrict__ ja, int32_t *__restrict__ jb) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= dim) return;
  uint32_t m = (uint32_t)mapp[r];
  uint32_t hi = m >> lo_bits, lo = m & ((1u << lo_bits) - 1u);
  jb[lo] = (int32_t)(r - ja[hi]);  // same value from every writer
}

struct EpsCoef {
  double lin[32];                          // per-site linear coefficient
  double pair[EDGPU_MAXORB][EDGPU_MAXORB]; // a<b parallel-spin coefficient
  int Norb;
};

__global__ void k_eps(const int32_t *__restrict__ map, int64_t dim, int Ns, EpsCoef c,
                      double *__restrict__ eps, uint8_t *__restrict__ imp) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= dim) return;
  uint32_t m = (uint32_t)map[r];
  double e = 0.0;
  for (int s = 0; s < Ns; s++)
    if ((m >> s) & 1u) e += c.lin[s];
  for (int a = 0; a < c.Norb; a++)
    for (int b = a + 1; b < c.Norb; b++)
      if (((m >> a) & 1u) && ((m >> b) & 1u)) e += c.pair[a][b];
  eps[r] = e;
  imp[r] = (uint8_t)(m & ((1u << c.Norb) - 1u));
}

// Fermionic sign of c^+_alpha c_beta on m (beta occupied, alpha empty): parity of the
// occupied sites strictly between the two positions = sg1*sg2 of c(), cdg()
// (ED_AUX_FUNX.f90:353-357, 379-383).
__device__ __forceinline__ uint32_t hop_sign(uint32_t m, int alpha, int beta) {
  int lo = min(alpha, beta), hi = max(alpha, beta);
  uint32_t between = ((1u << hi) - 1u) & ~((2u << lo) - 1u);
  return (uint32_t)(__popc(m & between) & 1);
}

// c^+_a c_b between impurity orbitals, all ordered pairs: the building block of the S-E / P-H
// terms (direct/HxV_non_local.f90:23-28, 52-56), tabulated once instead of ranked per product
__global__ void k_imphop_fill(const int32_t *__restrict__ map, int64_t dim, int64_t ld, int Norb,
                              RankView R, int32_t *__restrict__ out) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= ld) return;
  for (int a = 0; a < Norb; a++)
    for (int b = 0; b < Norb; b++) {
      int32_t val = -1;
      if (r < dim && a != b) {
        const uint32_t m = (uint32_t)map[r];
        if (((m >> b) & 1u) && !((m >> a) & 1u)) {
          const uint32_t m2 = (m & ~(1u << b)) | (1u << a);
          val = (int32_t)((uint32_t)rank_of(m2, R) | (hop_sign(m, a, b) << 31));
        }
      }
      out[(int64_t)(a * Norb + b) * ld + r] = val;
    }
}

// per-row counts of local / far allowed terms; far = the term touches a bit >= far_bit[r], or
// (sharded species, shard_q >= 0) its target column lies outside the rank's chunk
// [shard0, shard0 + shard_q) -- only the rows of the chunk are counted then.
// need != nullptr: also marks the columns outside the chunk that the chunk's hops read.
__global__ void k_hop_count(const int32_t *__restrict__ map, int64_t dim,
                            const Term *__restrict__ terms, int nterms,
                            const uint8_t *__restrict__ far_bit, SiteOrder ord, RankView R,
                            int64_t shard0, int64_t shard_q, unsigned char *__restrict__ need,
                            int nl_norb, int *__restrict__ wmax) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= dim) return;
  const bool sharded = shard_q >= 0;
  if (sharded && (r < shard0 || r >= shard0 + shard_q)) return;
  uint32_t m = (uint32_t)map[r];
  // non-local S-E / P-H terms (nl_norb > 0): they move an electron between two impurity orbitals of
  // this species; the (rare) targets outside the chunk join the halo
  if (sharded && need)
    for (int a = 0; a < nl_norb; a++)
      for (int b = 0; b < nl_norb; b++)
        if (a != b && ((m >> b) & 1u) && !((m >> a) & 1u)) {
          const int64_t tgt = rank_of((m & ~(1u << b)) | (1u << a), R);
          if (tgt < shard0 || tgt >= shard0 + shard_q) need[tgt] = 1;
        }
  const int fb = far_bit[r];
  int nl = 0, nf = 0;
  for (int t = 0; t < nterms; t++) {
    int a = terms[t].alpha, b = terms[t].beta;
    if (((m >> b) & 1u) && !((m >> a) & 1u)) {
      bool far = (ord.pos[a] >= fb || ord.pos[b] >= fb);
      if (sharded) {
        const int64_t tgt = rank_of((m & ~(1u << b)) | (1u << a), R);
        if (tgt < shard0 || tgt >= shard0 + shard_q) {
          far = true;
          if (need) need[tgt] = 1;
        }
      }
      if (far) nf++; else nl++;
    }
  }
  atomicMax(wmax, nl);
  atomicMax(wmax + 1, nf);
  atomicMax(wmax + 2, nl + nf);
}

// Row r (source state j of direct/HxV_up.f90): each entry holds the target row i of an
// allowed term (in the reference's term order within its section) and the signed amplitude
// index, so that   Hv(j) += amp2[idx] * v(i)     (gather form of HxV_up.f90:23-27).
// merged == 0 (fast role): local entries in slots [0, 4*Wl4), far entries in [4*Wl4, ...)
// merged == 1 (slow role): one list, far entries (flagged HOP_FAR) first, then the local ones
__global__ void k_hop_fill(const int32_t *__restrict__ map, int64_t dim, int64_t ld,
                           const Term *__restrict__ terms, int nterms,
                           const uint8_t *__restrict__ far_bit, int merged, int Wl4, int Wf4,
                           RankView R, const BlockItem *__restrict__ items,
                           const int32_t *__restrict__ item_of_row, int64_t shard0, int64_t shard_q,
                           const int32_t *__restrict__ colmap, uint32_t *__restrict__ ell) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= ld) return;
  // sharded species (merged only): rows of the chunk get LOCAL targets through colmap (local column
  // or shard_q + halo slot); rows outside the chunk are never used (padding)
  const bool sharded = shard_q >= 0;
  const bool inside = !sharded || (r >= shard0 && r < shard0 + shard_q);
  // slot s of this row lives at ell[((s/4)*ld + r)*4 + s%4]
  auto put = [&](int s, uint32_t val) { ell[((int64_t)(s >> 2) * ld + r) * 4 + (s & 3)] = val; };
  const int nslots = 4 * (Wl4 + Wf4);
  int el = 0, ef = merged ? 0 : 4 * Wl4;
  if (r < dim && inside) {
    uint32_t m = (uint32_t)map[r];
    const int fb = far_bit[r];
    if (merged) {  // first pass: far entries, second pass: local entries behind them
      for (int pass = 0; pass < 2; pass++)
        for (int t = 0; t < nterms; t++) {
          int a = terms[t].alpha, b = terms[t].beta;
          if (((m >> b) & 1u) && !((m >> a) & 1u)) {
            bool far = (R.ord.pos[a] >= fb || R.ord.pos[b] >= fb);
            uint32_t m2 = (m & ~(1u << b)) | (1u << a);
            uint32_t tgt = (uint32_t)rank_of(m2, R);
            if (sharded) {
              far = far || ((int64_t)tgt < shard0 || (int64_t)tgt >= shard0 + shard_q);
              tgt = (uint32_t)colmap[tgt];
            }
            if ((int)far == pass) continue;
            put(el++, tgt | ((uint32_t)(2 * t + hop_sign(m, a, b)) << HOP_AMP_SHIFT) |
                          (far ? HOP_FAR : 0u));
          }
        }
    } else {
      for (int t = 0; t < nterms; t++) {
        int a = terms[t].alpha, b = terms[t].beta;
        if (((m >> b) & 1u) && !((m >> a) & 1u)) {
          uint32_t m2 = (m & ~(1u << b)) | (1u << a);
          uint32_t tgt = (uint32_t)rank_of(m2, R);
          const bool far = (R.ord.pos[a] >= fb || R.ord.pos[b] >= fb);
          // block mode: a local target is addressed by its offset in the work item's tile
          if (items && !far) tgt = (uint32_t)block_tile_offset(items[item_of_row[r]], (int)tgt);
          uint32_t val = tgt | ((uint32_t)(2 * t + hop_sign(m, a, b)) << HOP_AMP_SHIFT);
          if (far) put(ef++, val); else put(el++, val);
        }
      }
    }
  }
  // padding slots gather the row itself (always inside the row's own tile; block mode: tile
  // offset 0) with amplitude 0; far padding points at the row itself
  const uint32_t padamp = (uint32_t)(2 * nterms) << HOP_AMP_SHIFT;
  const uint32_t pad = (uint32_t)(sharded ? (inside ? r - shard0 : 0) : r) | padamp;
  if (merged) {
    for (; el < nslots; el++) put(el, pad);
  } else {
    for (; el < 4 * Wl4; el++) put(el, items ? padamp : pad);
    for (; ef < nslots; ef++) put(ef, pad);
  }
}

// ---------------------------------------------------------------------------------------
// Directed one-body term list of one spin species in the reference's loop order
// (direct/HxV_up.f90:11-122; exc_field follows the stored path, stored/H_up.f90:85-103,
//  i.e. all iorb != jorb, see DESIGN.md "reference quirks").
// ---------------------------------------------------------------------------------------
static void build_terms(const edgpu_normal_params &p, int s, std::vector<Term> &out) {
  out.clear();
  const int No = p.Norb, Nb = p.Nbath;
  for (int io = 0; io < No; io++)
    for (int jo = 0; jo < No; jo++)
      if (io != jo && p.eloc[s][io][jo] != 0.0) out.push_back({io, jo, p.eloc[s][io][jo]});
  if (p.bath_type == EDGPU_BATH_REPLICA || p.bath_type == EDGPU_BATH_GENERAL)
    for (int kp = 0; kp < Nb; kp++)
      for (int io = 0; io < No; io++)
        for (int jo = 0; jo < No; jo++)
          if (io != jo && p.hbath[s][io][jo][kp] != 0.0)
            out.push_back({p.stride[io][kp] - 1, p.stride[jo][kp] - 1, p.hbath[s][io][jo][kp]});
  for (int io = 0; io < No; io++)
    for (int kp = 0; kp < Nb; kp++) {
      double vk = p.diag_hybr[s][io][kp];
      if (vk == 0.0) continue;
      int site = p.stride[io][kp] - 1;
      out.push_back({site, io, vk});  // c^+_bath c_imp   (HxV_up.f90:67-78)
      out.push_back({io, site, vk});  // c^+_imp c_bath   (HxV_up.f90:81-92)
    }
  bool exc = false;
  for (int i = 0; i < 4; i++) exc |= (p.exc_field[i] != 0.0);
  if (exc)
    for (int io = 0; io < No; io++)
      for (int jo = 0; jo < No; jo++) {
        if (io == jo) continue;
        out.push_back({io, jo, p.exc_field[0]});
        out.push_back({io, jo, (s == 0 ? 1.0 : -1.0) * p.exc_field[3]});
      }
}

static void build_eps_coef(const edgpu_normal_params &p, int s, EpsCoef &c) {
  memset(&c, 0, sizeof(c));
  const int No = p.Norb, Nb = p.Nbath;
  c.Norb = No;
  const double sf = (s == 0 ? 1.0 : -1.0);
  for (int a = 0; a < No; a++) {
    c.lin[a] = p.eloc[s][a][a] - p.xmu + sf * p.spin_field_z[a];
    if (p.hfmode) c.lin[a] += -0.5 * p.Uloc[a];
  }
  for (int a = 0; a < No; a++)
    for (int b = a + 1; b < No; b++) {
      c.pair[a][b] = p.Ust[a][b] - p.Jh[a][b];
      if (p.hfmode) {
        double h = -0.5 * p.Ust[a][b] - 0.5 * (p.Ust[a][b] - p.Jh[a][b]);
        c.lin[a] += h;
        c.lin[b] += h;
      }
    }
  for (int a = 0; a < p.Nfoo; a++)
    for (int k = 0; k < Nb; k++) c.lin[p.stride[a][k] - 1] += p.bath_diag[s][a][k];
}

static int free_spin(SpinSpace &S) {
  cudaFree(S.map);
  cudaFree(S.refidx);
  cudaFree(S.lin.ja);
  cudaFree(S.lin.jb);
  cudaFree(S.eps);
  cudaFree(S.imp);
  cudaFree(S.ell4);
  cudaFree(S.amp2);
  cudaFree(S.imphop);
  cudaFree(S.d_range_start);
  cudaFree(S.d_items);
  cudaFree(S.d_item_of_row);
  cudaFree(S.d_colmap);
  S = SpinSpace();
  return 0;
}

// Recursive range construction: the states whose top `t` bits equal a prefix with `p` set
// bits are contiguous in the ascending map (C(Ns-t, nel-p) of them); a prefix whose block
// exceeds `cap` is split on the next bit (0-child first = ascending order).
static void split_ranges(int Ns, int nel, int t, int p, int64_t cap, int64_t &pos,
                         std::vector<int64_t> &start, std::vector<int> &tbits) {
  const int64_t cnt = host_binomial(Ns - t, nel - p);
  if (cnt <= 0) return;
  if (cnt <= cap || t == Ns) {
    start.push_back(pos);
    tbits.push_back(t);
    pos += cnt;
    return;
  }
  split_ranges(Ns, nel, t + 1, p, cap, pos, start, tbits);
  split_ranges(Ns, nel, t + 1, p + 1, cap, pos, start, tbits);
}

// Site order of a species (see SiteOrder): MSB -> LSB = the last T bath sites (range bits),
// the impurity sites, the remaining bath sites; identity when asked (the species whose index is
// split over ranks keeps the reference order so that the rank-local chunk is the reference's).
static SiteOrder make_site_order(int Ns, int Norb, int T, bool identity) {
  SiteOrder o;
  memset(&o, 0, sizeof(o));
  o.Ns = Ns;
  o.identity = identity ? 1 : 0;
  std::vector<int> msb;  // sites from most to least significant
  if (identity) {
    for (int b = Ns - 1; b >= 0; b--) msb.push_back(b);
  } else {
    const int nb = Ns - Norb;
    const int top = std::min(T, nb);
    for (int k = 0; k < top; k++) msb.push_back(Ns - 1 - k);
    for (int a = Norb - 1; a >= 0; a--) msb.push_back(a);
    for (int b = Ns - 1 - top; b >= Norb; b--) msb.push_back(b);
  }
  for (int k = 0; k < Ns; k++) {
    const int p = Ns - 1 - k;
    o.site[p] = (uint8_t)msb[k];
    o.pos[msb[k]] = (uint8_t)p;
  }
  bool id = true;
  for (int b = 0; b < Ns; b++) id = id && (o.pos[b] == b);
  o.identity = id ? 1 : 0;
  return o;
}

RankView rank_view(const LinTable &lin, const SiteOrder &ord) {
  RankView R;
  R.lo_bits = lin.lo_bits;
  R.ja = lin.ja;
  R.jb = lin.jb;
  R.ord = ord;
  return R;
}

static int upload_binom(Engine &E) {
  int32_t hb[33][33];
  for (int n = 0; n < 33; n++)
    for (int k = 0; k < 33; k++) {
      int64_t b = host_binomial(n, k);
      hb[n][k] = (int32_t)std::min<int64_t>(b, INT32_MAX);
    }
  EDGPU_CUDA(cudaMemcpyToSymbolAsync(c_binom, hb, sizeof(hb), 0, cudaMemcpyHostToDevice, E.stream));
  return 0;
}

// Everything that fixes the internal enumeration order and the tiling of one species: role,
// ranges or blocks, number of range bits T, site order.  Depends on (model, nel, communicator)
// only, so a sector opened later with the same data gets the same order (apply_op relies on it).
struct SpeciesPlan {
  int role = ROLE_FAST;
  bool identity = false;
  bool block_mode = false;
  int T = 0;
  SiteOrder ord;
  std::vector<int64_t> range_start;  // nranges+1
  std::vector<int> range_tbits;
  std::vector<BlockItem> items;
  int64_t max_range = 0, max_tile = 0;
};

static void plan_species(Engine &E, const edgpu_normal_params &p, int s, int nel, SpeciesPlan &P) {
  const int Ns = p.Ns, No = p.Norb;
  std::vector<Term> terms;
  build_terms(p, s, terms);
  // opt-in: plan the up species like a slow-role one for the transposed-tile kernel k_fastT
  // (hxv.cu).  Read once per process: stored states depend on the enumeration order.
  static const bool upt = getenv("EDGPU_UPT") && atoi(getenv("EDGPU_UPT")) != 0;
  const bool up_t = (s == 0 && E.nranks == 1 && upt);
  const size_t per_cta = (E.smem_per_sm - 2 * 1024) / 2;  // 2 CTAs per SM
  // amplitudes + the per-column diagonal table (4 columns in k_fastb, 16 in k_fastT)
  const size_t tables = 8 * (2 * terms.size() + 2 + (up_t ? 16 : 4) * ((size_t)1 << No)) + 64 + 512;
  const size_t avail = per_cta > tables ? per_cta - tables : 0;
  P.identity = (s == 1 && E.nranks > 1);
  // the dw species is applied along the slow index: by k_slow on one rank and (halo mode) on the
  // rank's chunk of columns; in transpose mode it acts on v^T, where dw is the fast index
  P.role = (s == 1 && (E.nranks == 1 || E.dw_halo)) ? ROLE_SLOW : ROLE_FAST;
  if (up_t) P.role = ROLE_SLOW;
  P.block_mode = false;
  P.items.clear();
  // ---- block mode (fast role, permuted order): tile = input blocks x 4 columns (2 planes of
  // double2 = 32 B per row)
  if (P.role == ROLE_FAST && !P.identity && !getenv("EDGPU_NO_BLOCKS")) {
    const int64_t cap = (int64_t)(avail / 32);
    for (int T = 0; T <= Ns - No && !P.block_mode; T++) {
      const SiteOrder ord = make_site_order(Ns, No, T, false);
      const int R = Ns - No - T;
      std::vector<int> F;  // distinct impurity flip masks of the local terms
      for (const Term &t : terms) {
        if (ord.pos[t.alpha] >= Ns - T || ord.pos[t.beta] >= Ns - T) continue;
        int f = 0;
        if (t.alpha < No) f ^= 1 << t.alpha;
        if (t.beta < No) f ^= 1 << t.beta;
        if (std::find(F.begin(), F.end(), f) == F.end()) F.push_back(f);
      }
      std::sort(F.begin(), F.end());
      if ((int)F.size() > BLK_MAXIN) break;  // more range bits do not help
      const int nP = 1 << T, nI = 1 << No;
      std::vector<int64_t> start((size_t)nP * nI + 1, 0);
      int64_t pos = 0;
      for (int pr = 0; pr < nP; pr++)
        for (int im = 0; im < nI; im++) {
          start[(size_t)pr * nI + im] = pos;
          pos += host_binomial(R, nel - __builtin_popcount(pr) - __builtin_popcount(im));
        }
      start[(size_t)nP * nI] = pos;
      std::vector<BlockItem> items;
      int64_t max_tile = 0;
      bool ok = true;
      for (int pr = 0; pr < nP && ok; pr++)
        for (int im = 0; im < nI && ok; im++) {
          const size_t b = (size_t)pr * nI + im;
          if (start[b + 1] == start[b]) continue;
          BlockItem it;
          memset(&it, 0, sizeof(it));
          it.out0 = (int32_t)start[b];
          it.out1 = (int32_t)start[b + 1];
          for (int f : F) {
            const size_t bi = (size_t)pr * nI + (im ^ f);
            const int64_t len = start[bi + 1] - start[bi];
            if (len <= 0) continue;
            // Tile offset with the PARITY of the block's first global row, after at least one spare
            // row: the TMA staging (k_fastc) copies 16-byte aligned pieces, i.e. it starts one row
            // early / ends one row late when the block starts / ends on an odd row.
            int32_t off = it.nin == 0 ? 0 : it.tile_rows + 1;
            if ((off ^ (int32_t)start[bi]) & 1) off++;
            it.in0[it.nin] = (int32_t)start[bi];
            it.in_len[it.nin] = (int32_t)len;
            it.in_off[it.nin] = off;
            it.tile_rows = off + (int32_t)len;
            it.nin++;
          }
          it.tile_rows = (it.tile_rows + 2) & ~1;  // room for the row behind an odd end, even size
          if (it.tile_rows > cap) ok = false;
          max_tile = std::max<int64_t>(max_tile, it.tile_rows);
          items.push_back(it);
        }
      if (!ok) continue;
      P.block_mode = true;
      P.T = T;
      P.ord = ord;
      P.items = items;
      P.max_tile = max_tile;
      P.range_start.clear();
      P.range_tbits.clear();
      for (int pr = 0; pr < nP; pr++) {
        if (start[(size_t)(pr + 1) * nI] == start[(size_t)pr * nI]) continue;
        P.range_start.push_back(start[(size_t)pr * nI]);
        P.range_tbits.push_back(T);
      }
      P.range_start.push_back(pos);
    }
  }
  if (!P.block_mode) {
    // ---- range mode: fast role tile[range + 32][2] doubles, slow role tile[range][SLOW_ROWS]
    // slow role: 128 B per range element; the up species of k_fastT: up to 144 B (padded pitch)
    const int64_t cap = P.role == ROLE_SLOW
                            ? std::max<int64_t>((int64_t)(avail / (8 * (up_t ? 18 : SLOW_ROWS))), 1)
                            : std::max<int64_t>((int64_t)(avail / 16) - 32, 1);
    P.range_start.clear();
    P.range_tbits.clear();
    int64_t pos = 0;
    split_ranges(Ns, nel, 0, 0, cap, pos, P.range_start, P.range_tbits);
    P.range_start.push_back(pos);
    P.T = 0;
    for (int t : P.range_tbits) P.T = std::max(P.T, t);
    P.ord = make_site_order(Ns, No, P.T, P.identity);
  }
  P.max_range = 0;
  for (size_t k = 0; k + 1 < P.range_start.size(); k++)
    P.max_range = std::max(P.max_range, P.range_start[k + 1] - P.range_start[k]);
}

// map (original Fock integers in internal order) + ranking tables of one species
static int build_ranking(Engine &E, int Ns, int nel, const SiteOrder &ord, int64_t dim, int64_t ld,
                         int32_t **map, LinTable *lin) {
  cudaStream_t st = E.stream;
  const int T = 256;
  const unsigned gb = (unsigned)((dim + T - 1) / T);
  int32_t *mapp = nullptr;
  EDGPU_CUDA(cudaMalloc(map, sizeof(int32_t) * ld));
  EDGPU_CUDA(cudaMalloc(&mapp, sizeof(int32_t) * ld));
  EDGPU_CUDA(cudaMemsetAsync(*map, 0, sizeof(int32_t) * ld, st));
  k_build_map<<<gb, T, 0, st>>>(*map, mapp, dim, Ns, nel, ord);
  EDGPU_COUNT_LAUNCH();
  lin->lo_bits = Ns / 2;
  const int hi_bits = Ns - lin->lo_bits;
  EDGPU_CUDA(cudaMalloc(&lin->ja, sizeof(int32_t) * ((size_t)1 << hi_bits)));
  EDGPU_CUDA(cudaMalloc(&lin->jb, sizeof(int32_t) * ((size_t)1 << lin->lo_bits)));
  EDGPU_CUDA(cudaMemsetAsync(lin->ja, 0, sizeof(int32_t) * ((size_t)1 << hi_bits), st));
  EDGPU_CUDA(cudaMemsetAsync(lin->jb, 0, sizeof(int32_t) * ((size_t)1 << lin->lo_bits), st));
  k_lin_ja<<<gb, T, 0, st>>>(mapp, dim, lin->lo_bits, lin->ja);
  EDGPU_COUNT_LAUNCH();
  k_lin_jb<<<gb, T, 0, st>>>(mapp, dim, lin->lo_bits, lin->ja, lin->jb);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaStreamSynchronize(st));
  cudaFree(mapp);
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int species_ranking(Engine &E, const edgpu_normal_params &p, int s, int nel, int32_t **map,
                    LinTable *lin, SiteOrder *ord) {
  SpeciesPlan P;
  plan_species(E, p, s, nel, P);
  *ord = P.ord;
  EDGPU_TRY(upload_binom(E));
  const int64_t dim = host_binomial(p.Ns, nel);
  return build_ranking(E, p.Ns, nel, *ord, dim, (dim + 15) / 16 * 16, map, lin);
}

// need_all (sharded species only): receives the all-gathered "columns my hops read from other
// ranks" maps of every rank, [nranks][dim] bytes, for comm_halo_setup
static int build_spin(Engine &E, const edgpu_normal_params &p, int s, int nel, SpinSpace &S,
                      int64_t shard0 = 0, int64_t shard_q = -1,
                      std::vector<unsigned char> *need_all = nullptr) {
  const int Ns = p.Ns;
  cudaStream_t st = E.stream;
  SpeciesPlan P;
  plan_species(E, p, s, nel, P);
  const int role = P.role;
  const bool sharded = shard_q >= 0 && role == ROLE_SLOW;
  if (!sharded) shard_q = -1;
  S.nel = nel;
  S.role = role;
  S.dim = host_binomial(Ns, nel);
  S.ld = (S.dim + 15) / 16 * 16;
  if (S.dim > (int64_t)HOP_TGT_MASK) return set_error("sector species dimension %lld too large", (long long)S.dim);
  const int T = 256;
  const unsigned gb = (unsigned)((S.dim + T - 1) / T), gl = (unsigned)((S.ld + T - 1) / T);
  // ranges / blocks -> per-row far bit (permuted positions >= far_bit are prefix bits)
  std::vector<uint8_t> far_bit((size_t)S.ld, (uint8_t)Ns);
  S.range_start = P.range_start;
  S.range_tbits = P.range_tbits;
  S.nranges = (int)S.range_start.size() - 1;
  if (S.range_start.back() != S.dim) return set_error("internal: range partition mismatch");
  S.max_range = P.max_range;
  for (int k = 0; k < S.nranges; k++)
    for (int64_t r = S.range_start[k]; r < S.range_start[k + 1]; r++)
      far_bit[(size_t)r] = (uint8_t)(Ns - S.range_tbits[k]);
  S.ord = P.ord;
  S.sharded = sharded;
  S.shard0 = sharded ? shard0 : 0;
  S.shard_q = sharded ? shard_q : 0;
  if (sharded) {
    // ranges of the rank's chunk [shard0, shard0 + shard_q) in LOCAL column coordinates: the
    // global prefix ranges cut at the chunk boundaries
    std::vector<int64_t> loc;
    std::vector<int> tb;
    for (int k = 0; k < S.nranges; k++) {
      const int64_t a = std::max(S.range_start[k], shard0), b = std::min(S.range_start[k + 1], shard0 + shard_q);
      if (b <= a) continue;
      loc.push_back(a - shard0);
      tb.push_back(S.range_tbits[k]);
    }
    loc.push_back(shard_q);
    S.range_start = loc;
    S.range_tbits = tb;
    S.nranges = (int)loc.size() - 1;
    S.max_range = 0;
    for (int k = 0; k < S.nranges; k++) S.max_range = std::max(S.max_range, loc[k + 1] - loc[k]);
  }
  EDGPU_CUDA(cudaMalloc(&S.d_range_start, sizeof(int64_t) * S.range_start.size()));
  EDGPU_CUDA(cudaMemcpyAsync(S.d_range_start, S.range_start.data(),
                             sizeof(int64_t) * S.range_start.size(), cudaMemcpyHostToDevice, st));
  S.block_mode = P.block_mode;
  S.items = P.items;
  S.max_tile = P.max_tile;
  if (S.block_mode) {
    std::vector<int32_t> ior((size_t)S.ld, 0);
    for (size_t k = 0; k < S.items.size(); k++)
      for (int32_t r = S.items[k].out0; r < S.items[k].out1; r++) ior[(size_t)r] = (int32_t)k;
    EDGPU_CUDA(cudaMalloc(&S.d_items, sizeof(BlockItem) * S.items.size()));
    EDGPU_CUDA(cudaMalloc(&S.d_item_of_row, sizeof(int32_t) * S.ld));
    EDGPU_CUDA(cudaMemcpyAsync(S.d_items, S.items.data(), sizeof(BlockItem) * S.items.size(),
                               cudaMemcpyHostToDevice, st));
    EDGPU_CUDA(cudaMemcpyAsync(S.d_item_of_row, ior.data(), sizeof(int32_t) * S.ld,
                               cudaMemcpyHostToDevice, st));
    EDGPU_CUDA(cudaStreamSynchronize(st));  // ior goes out of scope
  }
  EDGPU_TRY(build_ranking(E, Ns, nel, S.ord, S.dim, S.ld, &S.map, &S.lin));
  EDGPU_CUDA(cudaMalloc(&S.refidx, sizeof(int32_t) * S.ld));
  EDGPU_CUDA(cudaMemsetAsync(S.refidx, 0, sizeof(int32_t) * S.ld, st));
  k_ref_rank<<<gb, T, 0, st>>>(S.map, S.dim, Ns, S.refidx);
  EDGPU_COUNT_LAUNCH();
  if (p.Norb > 1) {
    EDGPU_CUDA(cudaMalloc(&S.imphop, sizeof(int32_t) * (size_t)p.Norb * p.Norb * S.ld));
    k_imphop_fill<<<gl, T, 0, st>>>(S.map, S.dim, S.ld, p.Norb, rank_view(S.lin, S.ord), S.imphop);
    EDGPU_COUNT_LAUNCH();
  }
  // diagonal single-spin energies
  EpsCoef c;
  build_eps_coef(p, s, c);
  EDGPU_CUDA(cudaMalloc(&S.eps, sizeof(double) * S.ld));
  EDGPU_CUDA(cudaMalloc(&S.imp, S.ld));
  EDGPU_CUDA(cudaMemsetAsync(S.eps, 0, sizeof(double) * S.ld, st));
  EDGPU_CUDA(cudaMemsetAsync(S.imp, 0, S.ld, st));
  k_eps<<<gb, T, 0, st>>>(S.map, S.dim, Ns, c, S.eps, S.imp);
  EDGPU_COUNT_LAUNCH();
  // hop table
  build_terms(p, s, S.terms);
  S.nterms = (int)S.terms.size();
  if (S.nterms > HOP_MAX_TERMS) return set_error("too many one-body terms (%d)", S.nterms);
  Term *d_terms = nullptr;
  int *d_w = nullptr;
  uint8_t *d_far = nullptr;
  EDGPU_CUDA(cudaMalloc(&d_terms, sizeof(Term) * (S.nterms + 1)));
  EDGPU_CUDA(cudaMalloc(&d_w, 3 * sizeof(int)));
  EDGPU_CUDA(cudaMalloc(&d_far, (size_t)S.ld));
  if (S.nterms)
    EDGPU_CUDA(cudaMemcpyAsync(d_terms, S.terms.data(), sizeof(Term) * S.nterms,
                               cudaMemcpyHostToDevice, st));
  EDGPU_CUDA(cudaMemcpyAsync(d_far, far_bit.data(), (size_t)S.ld, cudaMemcpyHostToDevice, st));
  EDGPU_CUDA(cudaMemsetAsync(d_w, 0, 3 * sizeof(int), st));
  unsigned char *d_need = nullptr;
  int32_t *d_colmap = nullptr;
  if (sharded) {
    EDGPU_CUDA(cudaMalloc(&d_need, (size_t)S.ld));
    EDGPU_CUDA(cudaMemsetAsync(d_need, 0, (size_t)S.ld, st));
  }
  bool nl = false;  // nonloc_condition (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:45)
  for (int a = 0; a < p.Norb; a++)
    for (int b = 0; b < p.Norb; b++) nl = nl || p.Jx[a][b] != 0.0 || p.Jp[a][b] != 0.0;
  k_hop_count<<<gb, T, 0, st>>>(S.map, S.dim, d_terms, S.nterms, d_far, S.ord, rank_view(S.lin, S.ord),
                                shard0, shard_q, d_need, (nl && p.Norb > 1) ? p.Norb : 0, d_w);
  EDGPU_COUNT_LAUNCH();
  int W[3] = {0, 0, 0};
  EDGPU_CUDA(cudaMemcpyAsync(W, d_w, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
  EDGPU_CUDA(cudaStreamSynchronize(st));
  if (sharded) {
    // halo slots: the remote columns this chunk reads, in ascending (= owner-major) order
    std::vector<unsigned char> mine((size_t)S.ld, 0);
    EDGPU_CUDA(cudaMemcpy(mine.data(), d_need, (size_t)S.ld, cudaMemcpyDeviceToHost));
    need_all->assign((size_t)E.nranks * (size_t)S.ld, 0);
    EDGPU_TRY(comm_allgather_bytes(E, mine.data(), need_all->data(), (size_t)S.ld));
    std::vector<int32_t> colmap((size_t)S.ld, -1);
    S.halo_cols.clear();
    for (int64_t d = 0; d < S.dim; d++) {
      if (d >= shard0 && d < shard0 + shard_q) {
        colmap[(size_t)d] = (int32_t)(d - shard0);
      } else if (mine[(size_t)d]) {
        colmap[(size_t)d] = (int32_t)(shard_q + (int64_t)S.halo_cols.size());
        S.halo_cols.push_back((int32_t)d);
      }
    }
    S.nhalo = (int64_t)S.halo_cols.size();
    if (shard_q + S.nhalo > (int64_t)HOP_TGT_MASK) return set_error("chunk + halo too large for the hop table");
    EDGPU_CUDA(cudaMalloc(&d_colmap, sizeof(int32_t) * (size_t)S.ld));
    EDGPU_CUDA(cudaMemcpy(d_colmap, colmap.data(), sizeof(int32_t) * (size_t)S.ld, cudaMemcpyHostToDevice));
    cudaFree(d_need);
  }
  const int merged = (role == ROLE_SLOW);
  if (merged) {
    S.Wl = W[2];
    S.Wf = W[1];  // largest number of far entries of a row (they lead the list)
  } else {
    S.Wl = W[0];
    S.Wf = W[1];
  }
  S.Wl4 = (S.Wl + 3) / 4;
  S.Wf4 = merged ? 0 : (S.Wf + 3) / 4;
  const int G = std::max(S.Wl4 + S.Wf4, 1);
  EDGPU_CUDA(cudaMalloc(&S.ell4, sizeof(uint4) * (size_t)G * S.ld));
  k_hop_fill<<<gl, T, 0, st>>>(S.map, S.dim, S.ld, d_terms, S.nterms, d_far, merged, S.Wl4, S.Wf4,
                               rank_view(S.lin, S.ord), S.d_items, S.d_item_of_row, shard0, shard_q,
                               d_colmap, (uint32_t *)S.ell4);
  EDGPU_COUNT_LAUNCH();
  std::vector<double> amp(2 * S.nterms + 2, 0.0);
  for (int t = 0; t < S.nterms; t++) {
    amp[2 * t] = S.terms[t].h;
    amp[2 * t + 1] = -S.terms[t].h;
  }
  EDGPU_CUDA(cudaMalloc(&S.amp2, sizeof(double) * amp.size()));
  EDGPU_CUDA(cudaMemcpyAsync(S.amp2, amp.data(), sizeof(double) * amp.size(),
                             cudaMemcpyHostToDevice, st));
  EDGPU_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_terms);
  cudaFree(d_w);
  cudaFree(d_far);
  S.d_colmap = d_colmap;  // kept: the non-local kernel maps its dw targets through it
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

// also releases what an open that failed half-way left behind (S.open is false then)
int sector_close(Engine &E) {
  Sector &S = E.sec;
  cudaStreamSynchronize(E.stream);
  cudaStreamSynchronize(E.comm_stream);
  cudaStreamSynchronize(E.dw_stream);
  if (E.nranks > 1) comm_p2p_teardown(E);
  free_spin(S.up);
  free_spin(S.dw);
  cudaFree(S.xud);
  cudaFree(S.jx);
  cudaFree(S.jp);
  cudaFree(S.comm_block);  // holds vt and hvr / the halo buffers
  cudaFree(S.hvt);
  cudaFree(S.d_sendlist);
  S.d_sendlist = nullptr;
  S.nsend = 0;
  S.halo[0] = S.halo[1] = nullptr;
  S.halo_mode = false;
  S.comm_block = nullptr;
  S.pipe_err = nullptr;
  S.hvr = nullptr;
  cudaFree(S.sendbuf);
  cudaFree(S.recvbuf);
  cudaFree(S.vfull);
  cudaFree(S.sundry);
  cudaFree(S.gph);
  S.sundry = nullptr;
  S.gph = nullptr;
  S.nsundry = 0;
  S.DimPh = 1;
  S.xud = S.jx = S.jp = nullptr;
  S.vt = S.hvt = S.sendbuf = S.recvbuf = S.vfull = nullptr;
  S.open = false;
  return 0;
}

int sector_open(Engine &E, const edgpu_normal_params *p, int nup, int ndw) {
  if (!E.inited) return set_error("edgpu_init was not called");
  sector_close(E);  // the open sector, or the leftovers of an open that failed half-way
  Sector &S = E.sec;
  if (p->Ns < 1 || p->Ns > 31) return set_error("Ns=%d out of range", p->Ns);
  if (p->Norb < 1 || p->Norb > EDGPU_MAXORB) return set_error("Norb=%d out of range", p->Norb);
  if (p->Nbath < 0 || p->Nbath > EDGPU_MAXBATH) return set_error("Nbath=%d out of range", p->Nbath);
  if (nup < 0 || nup > p->Ns || ndw < 0 || ndw > p->Ns)
    return set_error("sector (nup=%d,ndw=%d) outside [0,%d]", nup, ndw, p->Ns);
  S.prm = *p;
  S.Ns = p->Ns;
  S.Norb = p->Norb;
  EDGPU_TRY(upload_binom(E));
  // dw split (ED_HAMILTONIAN_NORMAL.f90:128-142).  The reference shrinks the communicator when
  // DimDw < MpiSize (:98-126); here the ranks beyond DimDw simply own no column (qdw = 0) and
  // keep taking part in the collectives, which is the same thing seen from the caller.
  block_split(host_binomial(p->Ns, ndw), E.nranks, E.rank, &S.qdw, &S.d0);
  block_split(host_binomial(p->Ns, nup), E.nranks, E.rank, &S.qup, &S.u0);
  // How the Hdw term crosses ranks: halo mode (default; the owners push the few remote columns a
  // chunk's hops read, then pass A runs on the chunk exactly as on one GPU) or the reference's
  // double transpose (EDGPU_DW_MODE=transpose: chunk pipeline over peer memory; EDGPU_NO_P2P=1:
  // NCCL grouped send/recv).  Every rank reads the same environment.
  {
    const char *mode = getenv("EDGPU_DW_MODE"), *nop2p = getenv("EDGPU_NO_P2P");
    E.dw_halo = E.nranks > 1 && E.nranks <= EDGPU_MAXRANKS && !(mode && !strcmp(mode, "transpose")) &&
                !(nop2p && nop2p[0] == '1');
  }
  std::vector<unsigned char> need_all;
  EDGPU_TRY(build_spin(E, *p, 0, nup, S.up));
  if (E.dw_halo)
    EDGPU_TRY(build_spin(E, *p, 1, ndw, S.dw, S.d0, S.qdw, &need_all));
  else
    EDGPU_TRY(build_spin(E, *p, 1, ndw, S.dw));
  // cross-spin interaction table + constants (direct/HxV_local.f90:34-70)
  const int No = p->Norb, nimp = 1 << No;
  std::vector<double> x((size_t)nimp * nimp, 0.0);
  double cst = 0.0;
  if (p->hfmode) {
    for (int a = 0; a < No; a++) cst += 0.25 * p->Uloc[a];
    for (int a = 0; a < No; a++)
      for (int b = a + 1; b < No; b++)
        cst += 0.5 * p->Ust[a][b] + 0.5 * (p->Ust[a][b] - p->Jh[a][b]);
  }
  for (int md = 0; md < nimp; md++)
    for (int mu = 0; mu < nimp; mu++) {
      double e = cst;
      for (int a = 0; a < No; a++) e += p->Uloc[a] * ((mu >> a) & 1) * ((md >> a) & 1);
      for (int a = 0; a < No; a++)
        for (int b = a + 1; b < No; b++)
          e += p->Ust[a][b] * (((mu >> a) & 1) * ((md >> b) & 1) + ((mu >> b) & 1) * ((md >> a) & 1));
      x[(size_t)md * nimp + mu] = e;
    }
  EDGPU_CUDA(cudaMalloc(&S.xud, sizeof(double) * x.size()));
  EDGPU_CUDA(cudaMemcpyAsync(S.xud, x.data(), sizeof(double) * x.size(), cudaMemcpyHostToDevice,
                             E.stream));
  // non-local S-E / P-H (direct/HxV_non_local.f90), nonloc_condition of DIRECT_HxV.f90:45
  S.nonlocal = false;
  if (No > 1)
    for (int a = 0; a < No; a++)
      for (int b = 0; b < No; b++)
        if (p->Jx[a][b] != 0.0 || p->Jp[a][b] != 0.0) S.nonlocal = true;
  {
    std::vector<double> jx(No * No), jp(No * No);
    for (int a = 0; a < No; a++)
      for (int b = 0; b < No; b++) {
        jx[a * No + b] = p->Jx[a][b];
        jp[a * No + b] = p->Jp[a][b];
      }
    EDGPU_CUDA(cudaMalloc(&S.jx, sizeof(double) * No * No));
    EDGPU_CUDA(cudaMalloc(&S.jp, sizeof(double) * No * No));
    EDGPU_CUDA(cudaMemcpyAsync(S.jx, jx.data(), sizeof(double) * No * No, cudaMemcpyHostToDevice, E.stream));
    EDGPU_CUDA(cudaMemcpyAsync(S.jp, jp.data(), sizeof(double) * No * No, cudaMemcpyHostToDevice, E.stream));
  }
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));

  S.halo_mode = false;
  if (E.dw_halo) {
    EDGPU_TRY(comm_halo_setup(E, need_all));  // collective; S.halo_mode = every rank mapped every peer
    if (!S.halo_mode) {
      // peers not mappable: rebuild the dw species for the transposed layout (NCCL transposes)
      E.dw_halo = false;
      free_spin(S.dw);
      EDGPU_TRY(build_spin(E, *p, 1, ndw, S.dw));
    }
  }
  if (E.nranks > 1 && !S.halo_mode) {
    // transposed block [DimDw (fast) x qup] for the dw hops, mapped into every peer when possible
    // (one allocation [flags | vt | hvr]: a single IPC handle per rank, comm.cu)
    const size_t nt = (size_t)S.padded_len_t();
    const size_t off_hvr = pipe_hvr_offset(S.dw.ld, S.qup);
    const size_t total = off_hvr + (sizeof(double) * (size_t)S.slice_len() + 255) / 256 * 256;
    EDGPU_CUDA(cudaMalloc(&S.comm_block, total));
    EDGPU_CUDA(cudaMalloc(&S.hvt, sizeof(double) * std::max<size_t>(nt, 1)));
    S.pipe_err = reinterpret_cast<int32_t *>(E.h_scal + 60);  // pinned host word (UVA: device-visible)
    *S.pipe_err = 0;
    S.vt = reinterpret_cast<double *>(S.comm_block + PIPE_FLAG_BYTES);
    S.hvr = reinterpret_cast<double *>(S.comm_block + off_hvr);
    // flags start at epoch 0, pad rows of vt / hvr stay zero for ever (never written by a peer)
    EDGPU_CUDA(cudaMemsetAsync(S.comm_block, 0, total, E.stream));
    EDGPU_CUDA(cudaMemsetAsync(S.hvt, 0, sizeof(double) * std::max<size_t>(nt, 1), E.stream));
    EDGPU_CUDA(cudaStreamSynchronize(E.stream));
    EDGPU_TRY(comm_p2p_setup(E));
  }
  EDGPU_TRY(hxv_upload_amps(E));  // hop amplitudes of both species -> constant bank (hxv.cu)
  EDGPU_TRY(extra_setup(E));  // coulomb_sundry + phonons (a10)
  S.variant = E.variant_request;
  S.open = true;
  return 0;
}

}  // namespace edgpu
