// ed_mode=nonsu2 on the device: sector map and stored Hamiltonian built by kernels, replacing
//   build_sector (nonsu2 branch)      ED_SECTOR.f90:335-368   m = iup + idw*2^Ns, popcount = Ntot,
//                                                             ascending m (idw outer, iup inner)
//   build_Hv_sector_nonsu2            ED_HAMILTONIAN_NONSU2.f90:31-130 (row split :72-79)
//   ed_buildH_nonsu2_main             ED_HAMILTONIAN_NONSU2_STORED_HxV.f90:29-190 with the element
//                                     generators ED_NONSU2/stored/Himp.f90, Hint.f90, Hbath.f90,
//                                     Himp_bath.f90 (normal / hybrid bath)
// The reference inserts element by element into a list of rows (`sp_insert_element`, O(row) search
// + realloc per element, ED_SPARSE_MATRIX.f90:346-357) after a recursive binary search of the
// target state.  Here one thread per row enumerates the same terms in the same order twice
// (count, then fill) straight into flat CSR arrays in HBM; the target index is the combinadic
// rank of the packed 2*Ns-bit state (the sector is the set of all Ntot-subsets in ascending
// integer order), fermionic signs are popc of a bit window over ALL lower bits (up and dw),
// as c/cdg do on the packed state (ED_AUX_FUNX.f90:334-384 called with pos+Ns, Himp.f90:62).
// The diagonal contributions (Himp, spin_field z, Hint, Hbath) are summed in the reference's
// order into one entry; off-diagonal entries that hit the same column stay separate entries
// (they add up in the product, like sp_insert_element's accumulation).
// The product itself is the CSR SpMV of csr.cu (complex).
#include <algorithm>

#include "edgpu_internal.cuh"

namespace edgpu {

struct Nonsu2Dev {
  edgpu_nonsu2_params p;
  int32_t binom[33][33];  // C(n,k), n,k <= 32 (entries that overflow int32 are never used)
  int32_t ntot;
  int32_t nbits;  // 2*Ns
};
__constant__ Nonsu2Dev c_n2;

__device__ __forceinline__ int64_t n2_rank(uint32_t m) {
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

__device__ __forceinline__ uint32_t n2_unrank(int64_t r) {
  uint32_t m = 0;
  int p = c_n2.nbits - 1;
  for (int k = c_n2.ntot; k >= 1; k--) {
    while (c_n2.binom[p][k] > r) p--;
    m |= 1u << p;
    r -= c_n2.binom[p][k];
    p--;
  }
  return m;
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

template <class Sink>
__device__ void n2_row(uint32_t m, int64_t i, Sink &s) {
  const edgpu_nonsu2_params &P = c_n2.p;
  const int Ns = P.Ns, No = P.Norb, Nb = P.Nbath;
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
  bool any_sf = false;
  for (int a = 0; a < No; a++)
    any_sf = any_sf || P.spin_field[a][0] != 0.0 || P.spin_field[a][1] != 0.0 || P.spin_field[a][2] != 0.0;
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
  {
    double h = 0.0;
    for (int a = 0; a < P.Nfoo; a++)
      for (int k = 0; k < Nb; k++) {
        const int st = P.stride[a][k] - 1;
        h += P.bath_e[0][a][k] * (double)((m >> st) & 1u) + P.bath_e[1][a][k] * (double)((m >> (st + Ns)) & 1u);
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
  // spin-flip local terms :85-110
  for (int is = 0; is < 2; is++) {
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
  // spin-flip hybridisation u :72-136
  for (int a = 0; a < No; a++)
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

int csr_adopt_device(Engine &E, bool cplx, int64_t nloc, int64_t nglobal, int64_t row0, int64_t *d_rowptr,
                     int32_t *d_cols, double *d_vals, int64_t nnz, int32_t *d_map);

int nonsu2_open(Engine &E, const edgpu_nonsu2_params *p, int ntot) {
  if (!E.inited) return set_error("edgpu_init was not called");
  if (E.sec.open) return set_error("close the direct-H sector before opening a nonsu2 one");
  if (E.csr.open) csr_close(E);
  const int nbits = 2 * p->Ns;
  if (p->Ns < 1 || nbits > 31) return set_error("nonsu2: 2*Ns = %d exceeds the 31-bit packed state", nbits);
  if (p->Norb < 1 || p->Norb > EDGPU_MAXORB || p->Nbath < 0 || p->Nbath > EDGPU_MAXBATH)
    return set_error("nonsu2: Norb/Nbath out of range");
  if (p->bath_type != EDGPU_BATH_NORMAL && p->bath_type != EDGPU_BATH_HYBRID)
    return set_error("nonsu2: only normal / hybrid baths are generated on the device "
                     "(replica / general: hand the host-built spH0 to edgpu_csr_open_z)");
  if (ntot < 0 || ntot > nbits) return set_error("nonsu2: Ntot = %d outside [0, %d]", ntot, nbits);
  static Nonsu2Dev h;  // 10 KB: not on the stack
  h.p = *p;
  h.ntot = ntot;
  h.nbits = nbits;
  for (int n = 0; n <= 32; n++)
    for (int k = 0; k <= 32; k++) {
      int64_t c = k > n ? 0 : (k == 0 || k == n ? 1 : (int64_t)h.binom[n - 1][k - 1] + h.binom[n - 1][k]);
      h.binom[n][k] = (int32_t)std::min<int64_t>(c, INT32_MAX);
    }
  const int64_t dim = host_binomial(nbits, ntot);
  if (dim > INT32_MAX) return set_error("nonsu2: sector dimension %lld exceeds 32-bit columns", (long long)dim);
  EDGPU_CUDA(cudaMemcpyToSymbolAsync(c_n2, &h, sizeof(h), 0, cudaMemcpyHostToDevice, E.stream));
  // row split MpiQ = Dim/P, remainder to the last rank (ED_HAMILTONIAN_NONSU2.f90:72-79)
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
    return rc;
  };
#define N2_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__)); \
  } while (0)
  N2_CUDA(cudaMalloc(&d_map, sizeof(int32_t) * std::max<int64_t>(dim, 1)));
  k_n2_map<<<(unsigned)((dim + 255) / 256), 256, 0, E.stream>>>(d_map, dim);
  EDGPU_COUNT_LAUNCH();
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
  int rc = csr_adopt_device(E, true, nloc, dim, row0, d_rowptr, d_cols, d_vals, nnz, d_map);
  if (rc) return fail(rc);
  return 0;
}

}  // namespace edgpu
