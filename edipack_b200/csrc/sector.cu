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

// r-th (0-based) Ns-bit pattern with nel bits set in ascending integer order
// = combinadic unranking; identical to the reference's popcount scan order.
__global__ void k_build_map(int32_t *__restrict__ map, int64_t dim, int Ns, int nel) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= dim) return;
  int64_t rem = r;
  int k = nel;
  uint32_t m = 0;
  for (int pos = Ns - 1; pos >= 0 && k > 0; --pos) {
    int64_t c = c_binom[pos][k];
    if (rem >= c) {
      m |= (1u << pos);
      rem -= c;
      --k;
    }
  }
  map[r] = (int32_t)m;
}

__global__ void k_lin_ja(const int32_t *__restrict__ map, int64_t dim, int lo_bits,
                         int32_t *__restrict__ ja) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= dim) return;
  uint32_t hi = (uint32_t)map[r] >> lo_bits;
  if (r == 0 || ((uint32_t)map[r - 1] >> lo_bits) != hi) ja[hi] = (int32_t)r;
}
__global__ void k_lin_jb(const int32_t *__restrict__ map, int64_t dim, int lo_bits,
                         const int32_t *__restrict__ ja, int32_t *__restrict__ jb) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= dim) return;
  uint32_t m = (uint32_t)map[r];
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

__device__ __forceinline__ int lin_rank(uint32_t m, int lo_bits, const int32_t *ja,
                                        const int32_t *jb) {
  return ja[m >> lo_bits] + jb[m & ((1u << lo_bits) - 1u)];
}

// Fermionic sign of c^+_alpha c_beta on m (beta occupied, alpha empty): parity of the
// occupied sites strictly between the two positions = sg1*sg2 of c(), cdg()
// (ED_AUX_FUNX.f90:353-357, 379-383).
__device__ __forceinline__ uint32_t hop_sign(uint32_t m, int alpha, int beta) {
  int lo = min(alpha, beta), hi = max(alpha, beta);
  uint32_t between = ((1u << hi) - 1u) & ~((2u << lo) - 1u);
  return (uint32_t)(__popc(m & between) & 1);
}

// per-row counts of local / far allowed terms; far = the term touches a bit >= far_bit
__global__ void k_hop_count(const int32_t *__restrict__ map, int64_t dim,
                            const Term *__restrict__ terms, int nterms, int far_bit,
                            int *__restrict__ wmax) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= dim) return;
  uint32_t m = (uint32_t)map[r];
  int nl = 0, nf = 0;
  for (int t = 0; t < nterms; t++) {
    int a = terms[t].alpha, b = terms[t].beta;
    if (((m >> b) & 1u) && !((m >> a) & 1u)) {
      if (a >= far_bit || b >= far_bit) nf++; else nl++;
    }
  }
  atomicMax(wmax, nl);
  atomicMax(wmax + 1, nf);
}

// Row r (source state j of direct/HxV_up.f90): each entry holds the target row i of an
// allowed term (in the reference's term order within its section) and the signed amplitude
// index, so that   Hv(j) += amp2[idx] * v(i)     (gather form of HxV_up.f90:23-27)
__global__ void k_hop_fill(const int32_t *__restrict__ map, int64_t dim, int64_t ld,
                           const Term *__restrict__ terms, int nterms, int far_bit, int Wl4,
                           int Wf4, int lo_bits, const int32_t *__restrict__ ja,
                           const int32_t *__restrict__ jb, uint32_t *__restrict__ ell) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= ld) return;
  // slot s of this row lives at ell[((s/4)*ld + r)*4 + s%4]
  auto put = [&](int s, uint32_t val) { ell[((int64_t)(s >> 2) * ld + r) * 4 + (s & 3)] = val; };
  int el = 0, ef = 4 * Wl4;
  if (r < dim) {
    uint32_t m = (uint32_t)map[r];
    for (int t = 0; t < nterms; t++) {
      int a = terms[t].alpha, b = terms[t].beta;
      if (((m >> b) & 1u) && !((m >> a) & 1u)) {
        uint32_t m2 = (m & ~(1u << b)) | (1u << a);
        uint32_t tgt = (uint32_t)lin_rank(m2, lo_bits, ja, jb);
        uint32_t val = tgt | ((uint32_t)(2 * t + hop_sign(m, a, b)) << HOP_AMP_SHIFT);
        if (a >= far_bit || b >= far_bit) put(ef++, val); else put(el++, val);
      }
    }
  }
  // padding slots gather the row itself (always inside the row's own tile) with amplitude 0
  const uint32_t pad = (uint32_t)r | ((uint32_t)(2 * nterms) << HOP_AMP_SHIFT);
  for (; el < 4 * Wl4; el++) put(el, pad);
  for (; ef < 4 * (Wl4 + Wf4); ef++) put(ef, pad);
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
  cudaFree(S.lin.ja);
  cudaFree(S.lin.jb);
  cudaFree(S.eps);
  cudaFree(S.imp);
  cudaFree(S.ell4);
  cudaFree(S.amp2);
  cudaFree(S.d_range_start);
  S = SpinSpace();
  return 0;
}

// cap = largest range (in states of this species) the tiled kernel of this role can stage
static int build_spin(Engine &E, const edgpu_normal_params &p, int s, int nel, SpinSpace &S,
                      int64_t cap) {
  const int Ns = p.Ns;
  cudaStream_t st = E.stream;
  S.nel = nel;
  S.dim = host_binomial(Ns, nel);
  S.ld = (S.dim + 15) / 16 * 16;
  if (S.dim > (int64_t)HOP_TGT_MASK) return set_error("sector species dimension %lld too large", (long long)S.dim);
  const int T = 256;
  const unsigned gb = (unsigned)((S.dim + T - 1) / T), gl = (unsigned)((S.ld + T - 1) / T);
  EDGPU_CUDA(cudaMalloc(&S.map, sizeof(int32_t) * S.ld));
  EDGPU_CUDA(cudaMemsetAsync(S.map, 0, sizeof(int32_t) * S.ld, st));
  k_build_map<<<gb, T, 0, st>>>(S.map, S.dim, Ns, nel);
  EDGPU_COUNT_LAUNCH();
  // ranking tables
  S.lin.lo_bits = Ns / 2;
  const int hi_bits = Ns - S.lin.lo_bits;
  EDGPU_CUDA(cudaMalloc(&S.lin.ja, sizeof(int32_t) * ((size_t)1 << hi_bits)));
  EDGPU_CUDA(cudaMalloc(&S.lin.jb, sizeof(int32_t) * ((size_t)1 << S.lin.lo_bits)));
  EDGPU_CUDA(cudaMemsetAsync(S.lin.ja, 0, sizeof(int32_t) * ((size_t)1 << hi_bits), st));
  EDGPU_CUDA(cudaMemsetAsync(S.lin.jb, 0, sizeof(int32_t) * ((size_t)1 << S.lin.lo_bits), st));
  k_lin_ja<<<gb, T, 0, st>>>(S.map, S.dim, S.lin.lo_bits, S.lin.ja);
  EDGPU_COUNT_LAUNCH();
  k_lin_jb<<<gb, T, 0, st>>>(S.map, S.dim, S.lin.lo_bits, S.lin.ja, S.lin.jb);
  EDGPU_COUNT_LAUNCH();
  // diagonal single-spin energies
  EpsCoef c;
  build_eps_coef(p, s, c);
  EDGPU_CUDA(cudaMalloc(&S.eps, sizeof(double) * S.ld));
  EDGPU_CUDA(cudaMalloc(&S.imp, S.ld));
  EDGPU_CUDA(cudaMemsetAsync(S.eps, 0, sizeof(double) * S.ld, st));
  EDGPU_CUDA(cudaMemsetAsync(S.imp, 0, S.ld, st));
  k_eps<<<gb, T, 0, st>>>(S.map, S.dim, Ns, c, S.eps, S.imp);
  EDGPU_COUNT_LAUNCH();
  // ranges: smallest number of fixed top bits such that every range fits `cap`
  {
    int t = 0;
    auto max_range = [&](int tt) {
      int64_t mx = 0;
      for (int j = 0; j <= tt; j++) mx = std::max(mx, host_binomial(Ns - tt, nel - j));
      return mx;
    };
    while (t < Ns && max_range(t) > cap) t++;
    S.tbits = t;
    S.range_start.clear();
    int64_t pos = 0;
    S.max_range = 0;
    for (uint32_t h = 0; h < (1u << t); h++) {
      int64_t cnt = host_binomial(Ns - t, nel - __builtin_popcount(h));
      if (cnt <= 0) continue;
      S.range_start.push_back(pos);
      pos += cnt;
      S.max_range = std::max(S.max_range, cnt);
    }
    S.range_start.push_back(pos);
    S.nranges = (int)S.range_start.size() - 1;
    if (pos != S.dim) return set_error("internal: range partition mismatch");
    EDGPU_CUDA(cudaMalloc(&S.d_range_start, sizeof(int64_t) * S.range_start.size()));
    EDGPU_CUDA(cudaMemcpyAsync(S.d_range_start, S.range_start.data(),
                               sizeof(int64_t) * S.range_start.size(), cudaMemcpyHostToDevice, st));
  }
  // hop table
  build_terms(p, s, S.terms);
  S.nterms = (int)S.terms.size();
  if (S.nterms > HOP_MAX_TERMS) return set_error("too many one-body terms (%d)", S.nterms);
  const int far_bit = Ns - S.tbits;
  Term *d_terms = nullptr;
  int *d_w = nullptr;
  EDGPU_CUDA(cudaMalloc(&d_terms, sizeof(Term) * (S.nterms + 1)));
  EDGPU_CUDA(cudaMalloc(&d_w, 2 * sizeof(int)));
  if (S.nterms)
    EDGPU_CUDA(cudaMemcpyAsync(d_terms, S.terms.data(), sizeof(Term) * S.nterms,
                               cudaMemcpyHostToDevice, st));
  EDGPU_CUDA(cudaMemsetAsync(d_w, 0, 2 * sizeof(int), st));
  k_hop_count<<<gb, T, 0, st>>>(S.map, S.dim, d_terms, S.nterms, far_bit, d_w);
  EDGPU_COUNT_LAUNCH();
  int W[2] = {0, 0};
  EDGPU_CUDA(cudaMemcpyAsync(W, d_w, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  EDGPU_CUDA(cudaStreamSynchronize(st));
  S.Wl = W[0];
  S.Wf = W[1];
  S.Wl4 = (W[0] + 3) / 4;
  S.Wf4 = (W[1] + 3) / 4;
  const int G = std::max(S.Wl4 + S.Wf4, 1);
  EDGPU_CUDA(cudaMalloc(&S.ell4, sizeof(uint4) * (size_t)G * S.ld));
  k_hop_fill<<<gl, T, 0, st>>>(S.map, S.dim, S.ld, d_terms, S.nterms, far_bit, S.Wl4, S.Wf4,
                               S.lin.lo_bits, S.lin.ja, S.lin.jb, (uint32_t *)S.ell4);
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
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int sector_close(Engine &E) {
  Sector &S = E.sec;
  if (!S.open) return 0;
  cudaStreamSynchronize(E.stream);
  free_spin(S.up);
  free_spin(S.dw);
  cudaFree(S.xud);
  cudaFree(S.jx);
  cudaFree(S.jp);
  cudaFree(S.vt);
  cudaFree(S.hvt);
  cudaFree(S.sendbuf);
  cudaFree(S.recvbuf);
  S.xud = S.jx = S.jp = nullptr;
  S.vt = S.hvt = S.sendbuf = S.recvbuf = nullptr;
  S.open = false;
  return 0;
}

int sector_open(Engine &E, const edgpu_normal_params *p, int nup, int ndw) {
  if (!E.inited) return set_error("edgpu_init was not called");
  if (E.sec.open) sector_close(E);
  Sector &S = E.sec;
  if (p->Ns < 1 || p->Ns > 31) return set_error("Ns=%d out of range", p->Ns);
  if (p->Norb < 1 || p->Norb > EDGPU_MAXORB) return set_error("Norb=%d out of range", p->Norb);
  if (p->Nbath < 0 || p->Nbath > EDGPU_MAXBATH) return set_error("Nbath=%d out of range", p->Nbath);
  if (nup < 0 || nup > p->Ns || ndw < 0 || ndw > p->Ns)
    return set_error("sector (nup=%d,ndw=%d) outside [0,%d]", nup, ndw, p->Ns);
  S.prm = *p;
  S.Ns = p->Ns;
  S.Norb = p->Norb;
  int32_t hb[33][33];
  for (int n = 0; n < 33; n++)
    for (int k = 0; k < 33; k++) {
      int64_t b = host_binomial(n, k);
      hb[n][k] = (int32_t)std::min<int64_t>(b, INT32_MAX);
    }
  EDGPU_CUDA(cudaMemcpyToSymbolAsync(c_binom, hb, sizeof(hb), 0, cudaMemcpyHostToDevice, E.stream));
  // shared-memory capacities of the tiled kernels at 2 CTAs per SM (hxv.cu):
  //   fast role: tile[range + 32][2] doubles ; slow role: tile[range][8] doubles ; + tables
  {
    std::vector<Term> tu, td;
    build_terms(*p, 0, tu);
    build_terms(*p, 1, td);
    const size_t per_cta = (E.smem_per_sm - 2 * 1024) / 2;
    const int nimp_ = 1 << p->Norb;
    auto cap_of = [&](size_t nterms, size_t bytes_per_state, int64_t slack) {
      const size_t tables = 8 * (2 * nterms + 2 + 2 * (size_t)nimp_) + 64;
      const size_t avail = per_cta > tables ? per_cta - tables : 0;
      return std::max<int64_t>((int64_t)(avail / bytes_per_state) - slack, 1);
    };
    EDGPU_TRY(build_spin(E, *p, 0, nup, S.up, cap_of(tu.size(), 16, 32)));
    EDGPU_TRY(build_spin(E, *p, 1, ndw, S.dw,
                         E.nranks == 1 ? cap_of(td.size(), 64, 0) : cap_of(td.size(), 16, 32)));
  }
  // dw split (ED_HAMILTONIAN_NORMAL.f90:128-142).  The reference shrinks the communicator
  // when DimDw < MpiSize (:98-126); here every rank must own at least one column and one row.
  if (E.nranks > 1 && (S.dw.dim < E.nranks || S.up.dim < E.nranks))
    return set_error("sector (DimUp=%lld,DimDw=%lld) smaller than the %d-rank communicator",
                     (long long)S.up.dim, (long long)S.dw.dim, E.nranks);
  block_split(S.dw.dim, E.nranks, E.rank, &S.qdw, &S.d0);
  block_split(S.up.dim, E.nranks, E.rank, &S.qup, &S.u0);
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

  S.variant = E.variant_request;
  S.open = true;
  return 0;
}

}  // namespace edgpu
