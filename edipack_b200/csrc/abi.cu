// extern "C" entry points declared in include/edgpu.h.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "edgpu_internal.cuh"

namespace edgpu {

Engine g;
int g_status = 0;
int64_t g_launches = 0;
static char g_errbuf[1024] = "";

int set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_errbuf, sizeof(g_errbuf), fmt, ap);
  va_end(ap);
  g_status = 1;
  return 1;
}
void clear_error() {
  g_errbuf[0] = 0;
  g_status = 0;
}

int lanczos_tridiag_dev(Engine &E, double *d_seed, double *d_work, int nlanc, double threshold,
                        double *alanc, double *blanc, int *nused);
int eigh_dev(Engine &E, int neigen, int nblock, int nitermax, double tol, uint64_t seed, double *evals,
             double *resid, std::vector<double *> *vecs, int *nconv, int *nmatvec);

// A state kept on the device (ED_EIGENSPACE state_list entry)
struct StoredState {
  // packed-state sectors (nonsu2 / superc, device-built): kind = 1, vector = this rank's rows
  int kind = 0, pk_mode = -1, pk_qn = 0;
  int64_t nglobal = 0, row0 = 0, nloc = 0, padded = 0;
  int Ns = 0, nup = 0, ndw = 0;
  int64_t dimu = 0, dimd = 0, ldu = 0, qdw = 0, d0 = 0;
  int dimph = 1;  // phonon slices, each [qdw][ldu]
  double *vec = nullptr;
};
static std::map<int, StoredState> g_states;
static double *g_current = nullptr;   // last ground-state vector (device, padded layout)
static int64_t g_current_len = 0;
static double *g_seed = nullptr;      // device-resident GF seed for the open sector
static int64_t g_seed_len = 0;
static std::vector<double *> g_eigvecs;  // eigenvectors of the last edgpu_eigh (open sector)
static void free_eigvecs() {
  for (double *p : g_eigvecs) cudaFree(p);
  g_eigvecs.clear();
}

static int ensure_buf(double **p, int64_t *len, int64_t need) {
  if (*p && *len >= need) return 0;
  cudaFree(*p);
  *p = nullptr;
  EDGPU_CUDA(cudaMalloc(p, sizeof(double) * need));
  *len = need;
  return 0;
}

// reference chunk layout (contiguous columns of DimUp, ascending Fock order of both species,
// i = iup + (idw-1)*DimUp, ED_SECTOR.f90:1681) <-> padded device layout in the internal
// enumeration order (SiteOrder): a plain strided copy when both orders are the reference's,
// else through a device staging buffer and a permutation kernel.
__global__ void __launch_bounds__(256)
k_permute_in(double *__restrict__ dst, int64_t ld, const double *__restrict__ src, int64_t nrow,
             int64_t col_offset, const int32_t *__restrict__ refup,
             const int32_t *__restrict__ refdw) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= ld) return;
  double x = 0.0;
  if (i < nrow) x = src[((int64_t)refdw[c + col_offset] - col_offset) * nrow + refup[i]];
  dst[c * ld + i] = x;
}
__global__ void __launch_bounds__(256)
k_permute_out(double *__restrict__ dst, const double *__restrict__ src, int64_t ld, int64_t nrow,
              int64_t col_offset, const int32_t *__restrict__ refup,
              const int32_t *__restrict__ refdw) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= nrow) return;
  dst[((int64_t)refdw[c + col_offset] - col_offset) * nrow + refup[i]] = src[c * ld + i];
}

static double *g_stage = nullptr;
static int64_t g_stage_len = 0;

static int upload(Engine &E, double *d_dst, const double *h_src) {
  if (E.csr.open) {
    const int64_t n = (E.csr.cplx ? 2 : 1) * E.csr.nloc, npad = E.csr.padded_len();
    EDGPU_CUDA(cudaMemcpyAsync(d_dst, h_src, sizeof(double) * n, cudaMemcpyHostToDevice, E.stream));
    if (npad > n) EDGPU_CUDA(cudaMemsetAsync(d_dst + n, 0, sizeof(double) * (npad - n), E.stream));
    return 0;
  }
  Sector &S = E.sec;
  const int64_t n = S.up.dim * S.qdw, slice = S.slice_len();  // per phonon slice
  if (S.qdw <= 0) return 0;  // a rank without columns (DimDw < nranks)
  if (S.up.ord.identity && S.dw.ord.identity) {
    EDGPU_CUDA(cudaMemsetAsync(d_dst, 0, sizeof(double) * S.padded_len(), E.stream));
    for (int iph = 0; iph < S.DimPh; iph++)
      EDGPU_CUDA(cudaMemcpy2DAsync(d_dst + iph * slice, sizeof(double) * S.up.ld, h_src + iph * n,
                                   sizeof(double) * S.up.dim, sizeof(double) * S.up.dim, (size_t)S.qdw,
                                   cudaMemcpyHostToDevice, E.stream));
    return 0;
  }
  EDGPU_TRY(ensure_buf(&g_stage, &g_stage_len, n * S.DimPh));
  EDGPU_CUDA(cudaMemcpyAsync(g_stage, h_src, sizeof(double) * n * S.DimPh, cudaMemcpyHostToDevice, E.stream));
  dim3 grid((unsigned)((S.up.ld + 255) / 256), (unsigned)S.qdw);
  for (int iph = 0; iph < S.DimPh; iph++) {
    k_permute_in<<<grid, 256, 0, E.stream>>>(d_dst + iph * slice, S.up.ld, g_stage + iph * n, S.up.dim,
                                             S.d0, S.up.refidx, S.dw.refidx);
    EDGPU_COUNT_LAUNCH();
  }
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}
static int download(Engine &E, double *h_dst, const double *d_src) {
  if (E.csr.open) {
    const int64_t n = (E.csr.cplx ? 2 : 1) * E.csr.nloc;
    EDGPU_CUDA(cudaMemcpyAsync(h_dst, d_src, sizeof(double) * n, cudaMemcpyDeviceToHost, E.stream));
    EDGPU_CUDA(cudaStreamSynchronize(E.stream));
    return 0;
  }
  Sector &S = E.sec;
  const int64_t n = S.up.dim * S.qdw, slice = S.slice_len();  // per phonon slice
  if (S.qdw <= 0) {
    EDGPU_CUDA(cudaStreamSynchronize(E.stream));
    return 0;
  }
  if (S.up.ord.identity && S.dw.ord.identity) {
    for (int iph = 0; iph < S.DimPh; iph++)
      EDGPU_CUDA(cudaMemcpy2DAsync(h_dst + iph * n, sizeof(double) * S.up.dim, d_src + iph * slice,
                                   sizeof(double) * S.up.ld, sizeof(double) * S.up.dim, (size_t)S.qdw,
                                   cudaMemcpyDeviceToHost, E.stream));
    EDGPU_CUDA(cudaStreamSynchronize(E.stream));
    return 0;
  }
  EDGPU_TRY(ensure_buf(&g_stage, &g_stage_len, n * S.DimPh));
  dim3 grid((unsigned)((S.up.dim + 255) / 256), (unsigned)S.qdw);
  for (int iph = 0; iph < S.DimPh; iph++) {
    k_permute_out<<<grid, 256, 0, E.stream>>>(g_stage + iph * n, d_src + iph * slice, S.up.ld, S.up.dim,
                                              S.d0, S.up.refidx, S.dw.refidx);
    EDGPU_COUNT_LAUNCH();
  }
  EDGPU_CUDA(cudaGetLastError());
  EDGPU_CUDA(cudaMemcpyAsync(h_dst, g_stage, sizeof(double) * n * S.DimPh, cudaMemcpyDeviceToHost, E.stream));
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  return 0;
}

// ---------------------------------------------------------------------------------------
// apply_op_C / apply_op_CDG (ED_SECTOR.f90:465-531, 654-720), gather form on the target
// sector:  OV(j) = sgn * V(i)  with  |j> = op |i>.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_apply_op(const double *__restrict__ vsrc, int64_t lds, double *__restrict__ out, int64_t ldo,
           int64_t nrow, int64_t ncol, const int32_t *__restrict__ map_t, int op, int bit, int spin,
           RankView Rsrc) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= nrow) return;
  const int64_t t = (spin == 0) ? i : c;  // index of the operated species in the target sector
  const uint32_t m = (uint32_t)map_t[t];
  const bool occ = (m >> bit) & 1u;
  double val = 0.0;
  if ((op < 0 && !occ) || (op > 0 && occ)) {
    const uint32_t ms = m ^ (1u << bit);  // source state of that species
    const double sgn = (__popc(ms & ((1u << bit) - 1u)) & 1) ? -1.0 : 1.0;
    const int64_t r = rank_of(ms, Rsrc);
    const int64_t si = (spin == 0) ? r : i, sc = (spin == 0) ? c : r;
    val = sgn * vsrc[sc * lds + si];
  }
  out[c * ldo + i] = val;
}

// apply_Cops of the NORMAL mode (ED_SECTOR.f90, apply_Cops; used by lanc_build_gf_normal_mix,
// ED_GF_NORMAL.f90:211,227): one term coef * O |state> ADDED to the seed.  Separate from k_apply_op
// (which overwrites) so that the single-operator path is untouched.
__global__ void __launch_bounds__(128)
k_apply_op_acc(const double *__restrict__ vsrc, int64_t lds, double *__restrict__ out, int64_t ldo,
               int64_t nrow, int64_t ncol, const int32_t *__restrict__ map_t, int op, int bit, int spin,
               RankView Rsrc, double coef) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= nrow) return;
  const int64_t t = (spin == 0) ? i : c;
  const uint32_t m = (uint32_t)map_t[t];
  const bool occ = (m >> bit) & 1u;
  if ((op < 0 && !occ) || (op > 0 && occ)) {
    const uint32_t ms = m ^ (1u << bit);
    const double sgn = (__popc(ms & ((1u << bit) - 1u)) & 1) ? -1.0 : 1.0;
    const int64_t r = rank_of(ms, Rsrc);
    const int64_t si = (spin == 0) ? r : i, sc = (spin == 0) ? c : r;
    out[c * ldo + i] += coef * sgn * vsrc[sc * lds + si];
  }
}

// Twin state (es_return_dvector `itwin` branch, ED_EIGENSPACE.f90:640-660 + twin_sector_order,
// ED_SECTOR.f90:1747-1776): the state of sector A = (nup,ndw) re-expressed in the twin sector
// B = (ndw,nup).  The reference sorts the flipped Fock integers mdw + mup*2^Ns of A and reads
// vector_B(i) = vec_A(Order(i)): B's state (mup_B, mdw_B) is A's state (mup_A, mdw_A) = (mdw_B, mup_B),
// i.e. the [DimUp, DimDw] matrix transposed.  Gather form on B, source indices ranked in A's
// internal enumeration orders.
__global__ void __launch_bounds__(128)
k_twin_normal(const double *__restrict__ vsrc, int64_t lds, double *__restrict__ out, int64_t ldo,
              int64_t nrow, const int32_t *__restrict__ map_up, const int32_t *__restrict__ map_dw,
              RankView RupA, RankView RdwA) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= nrow) return;
  const int64_t iuA = rank_of((uint32_t)map_dw[c], RupA);  // A's up pattern = B's dw pattern
  const int64_t idA = rank_of((uint32_t)map_up[i], RdwA);  // A's dw pattern = B's up pattern
  out[c * ldo + i] = vsrc[idA * lds + iuA];
}

// dens / docc partial sums (ED_OBSERVABLES_NORMAL.f90:150-215): out[a] += v^2 (nu+nd),
// out[Norb+a] += v^2 nu nd
__global__ void __launch_bounds__(256)
k_observables(const double *__restrict__ v, int64_t ld, int64_t nrow, int64_t ncol,
              int64_t col_offset, const uint8_t *__restrict__ impu,
              const uint8_t *__restrict__ impd, int Norb, double *__restrict__ out) {
  double acc[2 * EDGPU_MAXORB];
#pragma unroll
  for (int k = 0; k < 2 * EDGPU_MAXORB; k++) acc[k] = 0.0;
  for (int64_t c = blockIdx.y; c < ncol; c += gridDim.y) {
    const int md = impd[c + col_offset];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nrow;
         i += (int64_t)gridDim.x * blockDim.x) {
      const double x = v[c * ld + i];
      const double w = x * x;
      const int mu = impu[i];
#pragma unroll
      for (int a = 0; a < EDGPU_MAXORB; a++) {
        if (a < Norb) {
          const int nu = (mu >> a) & 1, nd = (md >> a) & 1;
          acc[a] += w * (nu + nd);
          acc[EDGPU_MAXORB + a] += w * (nu * nd);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 2 * EDGPU_MAXORB; k++) {
    double x = acc[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(out + k, x);
  }
}

}  // namespace edgpu

using namespace edgpu;

extern "C" {

int edgpu_init(int device) {
  clear_error();
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return set_error("no CUDA device available (%s); this engine has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return set_error("device %d out of range (%d devices)", device, ndev);
  EDGPU_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  EDGPU_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return set_error("device %d is sm_%d%d; this library is built for sm_100a only", device,
                     prop.major, prop.minor);
  if (g.inited && g.device == device) return 0;
  if (g.inited) edgpu_finalize();
  g.device = device;
  g.sm_count = prop.multiProcessorCount;
  g.smem_optin = prop.sharedMemPerBlockOptin;
  g.smem_per_sm = prop.sharedMemPerMultiprocessor;
  EDGPU_CUDA(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  {
    // the communication streams outrank the main one: their small kernels (tile pushes, flag
    // signals / waits) must not queue behind the CTAs of the rank-local pass
    int lo = 0, hi = 0;
    EDGPU_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    EDGPU_CUDA(cudaStreamCreateWithPriority(&g.comm_stream, cudaStreamNonBlocking, hi));
    EDGPU_CUDA(cudaStreamCreateWithPriority(&g.dw_stream, cudaStreamNonBlocking, hi));
  }
  EDGPU_CUDA(cudaEventCreateWithFlags(&g.ev_fork, cudaEventDisableTiming));
  EDGPU_CUDA(cudaEventCreateWithFlags(&g.ev_join, cudaEventDisableTiming));
  EDGPU_CUDA(cudaEventCreateWithFlags(&g.ev_join2, cudaEventDisableTiming));
  for (auto &ev : g.ev) EDGPU_CUDA(cudaEventCreate(&ev));
  g.part_cap = (int64_t)g.sm_count * 8;
  EDGPU_CUDA(cudaMalloc(&g.d_part, sizeof(double) * g.part_cap));
  EDGPU_CUDA(cudaMalloc(&g.d_scal, sizeof(double) * 64));
  EDGPU_CUDA(cudaMallocHost(&g.h_scal, sizeof(double) * 64));
  g.rank = 0;
  g.nranks = 1;
  g.inited = true;
  return 0;
}

int edgpu_release_cache(void) {
  lanczos_release(g);
  return 0;
}

int edgpu_finalize(void) {
  if (!g.inited) return 0;
  lanczos_release(g);
  sector_close(g);
  csr_close(g);
  for (auto &kv : g_states) cudaFree(kv.second.vec);
  g_states.clear();
  free_eigvecs();
  cudaFree(g_current);
  g_current = nullptr;
  g_current_len = 0;
  cudaFree(g_seed);
  g_seed = nullptr;
  g_seed_len = 0;
  cudaFree(g_stage);
  g_stage = nullptr;
  g_stage_len = 0;
  comm_finalize(g);
  cudaFree(g.d_part);
  cudaFree(g.d_scal);
  cudaFreeHost(g.h_scal);
  for (auto &ev : g.ev) cudaEventDestroy(ev);
  for (auto &ev : g.prof_ev) cudaEventDestroy(ev);
  cudaEventDestroy(g.ev_fork);
  cudaEventDestroy(g.ev_join);
  cudaEventDestroy(g.ev_join2);
  cudaStreamDestroy(g.comm_stream);
  cudaStreamDestroy(g.dw_stream);
  cudaStreamDestroy(g.stream);
  g = Engine();
  return 0;
}

int edgpu_comm_unique_id(void *uid) {
  clear_error();
  return comm_unique_id(uid);
}
int edgpu_comm_init(int rank, int nranks, const void *uid) {
  clear_error();
  return comm_init(g, rank, nranks, uid);
}
int edgpu_comm_rank(void) { return g.rank; }
int edgpu_comm_size(void) { return g.nranks; }

int edgpu_sector_open_normal(const edgpu_normal_params *p, int nup, int ndw) {
  clear_error();
  if (!p) return set_error("null params");
  if (g.csr.open) csr_close(g);
  return sector_open(g, p, nup, ndw);
}
int edgpu_sector_open_normal_orbs(const edgpu_normal_params *p, const int32_t *nups, const int32_t *ndws) {
  clear_error();
  if (!p || !nups || !ndws) return set_error("null params");
  free_eigvecs();
  return orbs_open(g, p, nups, ndws);
}

int edgpu_set_coulomb_sundry(int nterms, const edgpu_sundry_term *terms) {
  clear_error();
  if (nterms < 0 || nterms > EDGPU_MAXSUNDRY) return set_error("coulomb_sundry: %d terms (max %d)", nterms, EDGPU_MAXSUNDRY);
  if (nterms > 0 && !terms) return set_error("coulomb_sundry: null terms");
  for (int t = 0; t < nterms; t++) {
    const int32_t *ops[4] = {terms[t].cd_i, terms[t].cd_j, terms[t].c_k, terms[t].c_l};
    for (int k = 0; k < 4; k++) {
      if (ops[k][1] != 1 && ops[k][1] != 2) return set_error("coulomb_sundry term %d: spin must be 1 or 2", t);
      if (ops[k][0] < 1 || ops[k][0] > EDGPU_MAXORB) return set_error("coulomb_sundry term %d: orbital out of range", t);
    }
    // spin balance, direct/HxV_sundry.f90:24-35
    int sc = 0;
    sc += terms[t].c_l[1] == 1 ? 1 : -1;
    sc -= terms[t].cd_j[1] == 1 ? 1 : -1;
    sc += terms[t].c_k[1] == 1 ? 1 : -1;
    sc -= terms[t].cd_i[1] == 1 ? 1 : -1;
    if (sc != 0)
      return set_error("In NORMAL mode, operators that change the total spin are forbidden. Check your umatrix file (term %d)", t);
  }
  g.sundry_terms.assign(terms, terms + nterms);
  return 0;
}

int edgpu_set_phonons(int Nph, double w0_ph, double A_ph, const double *g_ph, int Norb) {
  clear_error();
  if (Nph < 0 || Nph > 4096) return set_error("Nph=%d out of range", Nph);
  if (Norb < 1 || Norb > EDGPU_MAXORB) return set_error("Norb=%d out of range", Norb);
  if (Nph > 0 && !g_ph) return set_error("g_ph is NULL");
  g.Nph = Nph;
  g.w0_ph = w0_ph;
  g.A_ph = A_ph;
  for (int a = 0; a < EDGPU_MAXORB; a++)
    for (int b = 0; b < EDGPU_MAXORB; b++)
      g.g_ph[a][b] = (Nph > 0 && a < Norb && b < Norb) ? g_ph[a * Norb + b] : 0.0;
  return 0;
}

int edgpu_sector_close(void) {
  clear_error();
  free_eigvecs();
  csr_close(g);
  return sector_close(g);
}
int edgpu_csr_open_d(int64_t nloc, int64_t nglobal, int64_t row_offset, const int64_t *rowptr,
                     const int32_t *cols, const double *vals) {
  clear_error();
  if (!rowptr || (!cols && rowptr[nloc] > 0)) return set_error("null CSR arrays");
  if (g.sec.open) sector_close(g);
  return csr_open(g, false, nloc, nglobal, row_offset, rowptr, cols, vals);
}
int edgpu_csr_open_z(int64_t nloc, int64_t nglobal, int64_t row_offset, const int64_t *rowptr,
                     const int32_t *cols, const double *vals_re_im) {
  clear_error();
  if (!rowptr || (!cols && rowptr[nloc] > 0)) return set_error("null CSR arrays");
  if (g.sec.open) sector_close(g);
  return csr_open(g, true, nloc, nglobal, row_offset, rowptr, cols, vals_re_im);
}
static bool any_open() { return g.sec.open || g.csr.open; }
int64_t edgpu_sector_vecdim(void) {
  if (g.csr.open) return g.csr.nloc;
  return g.sec.open ? g.sec.up.dim * g.sec.qdw * g.sec.DimPh : 0;
}
int64_t edgpu_sector_dim(void) {
  if (g.csr.open) return g.csr.nglobal;
  return g.sec.open ? g.sec.up.dim * g.sec.dw.dim * g.sec.DimPh : 0;
}
int edgpu_sector_dims(int64_t *DimUp, int64_t *DimDw, int64_t *qdw, int64_t *dw_start) {
  if (!g.sec.open) return set_error("no sector open");
  if (DimUp) *DimUp = g.sec.up.dim;
  if (DimDw) *DimDw = g.sec.dw.dim;
  if (qdw) *qdw = g.sec.qdw;
  if (dw_start) *dw_start = g.sec.d0;
  return 0;
}

int edgpu_sector_comm_info(int *mode, int64_t *halo_cols, int64_t *send_cols, int *nchunks) {
  if (!g.sec.open) return set_error("no sector open");
  const Sector &S = g.sec;
  if (mode) *mode = g.nranks == 1 ? 0 : (S.halo_mode ? 1 : (S.p2p ? 2 : 3));
  if (halo_cols) *halo_cols = S.halo_mode ? S.dw.nhalo : 0;
  if (send_cols) *send_cols = S.halo_mode ? S.nsend : 0;
  if (nchunks) *nchunks = S.nchunks;
  return 0;
}

int edgpu_halo_plan(int64_t dim_dw, int nranks, int rank, const unsigned char *need, int64_t *halo_cols,
                    int32_t *send_triples, int64_t send_cap, int64_t *nsend) {
  clear_error();
  if (nranks < 1 || rank < 0 || rank >= nranks || dim_dw < 0) return set_error("edgpu_halo_plan: bad arguments");
  std::vector<int64_t> nh;
  std::vector<std::vector<int32_t>> send;
  halo_plan(dim_dw, dim_dw, nranks, rank, need, nh, send);
  for (int r = 0; r < nranks; r++) halo_cols[r] = nh[r];
  int64_t n = 0;
  for (int r = 0; r < nranks; r++)
    for (size_t k = 0; k + 1 < send[r].size(); k += 2) {
      if (n < send_cap) {
        send_triples[3 * n] = send[r][k];
        send_triples[3 * n + 1] = r;
        send_triples[3 * n + 2] = send[r][k + 1];
      }
      n++;
    }
  *nsend = n;
  return 0;
}

int edgpu_sector_open_nonsu2(const edgpu_nonsu2_params *p, int ntot) {
  clear_error();
  if (!p) return set_error("null params");
  if (g.sec.open) sector_close(g);
  free_eigvecs();
  return nonsu2_open(g, p, ntot);
}

int edgpu_set_hbath_packed(const double *hbath_re_im, int Norb, int Nbath) {
  clear_error();
  if (!hbath_re_im) {
    g.hbath_packed.clear();
    g.hb_Norb = g.hb_Nbath = 0;
    return 0;
  }
  if (Norb < 1 || Norb > EDGPU_MAXORB || Nbath < 1 || Nbath > EDGPU_MAXBATH)
    return set_error("edgpu_set_hbath_packed: Norb/Nbath out of range");
  g.hbath_packed.assign(hbath_re_im, hbath_re_im + (size_t)8 * Norb * Norb * Nbath);
  g.hb_Norb = Norb;
  g.hb_Nbath = Nbath;
  return 0;
}

int edgpu_sector_open_superc(const edgpu_superc_params *p, int sz) {
  clear_error();
  if (!p) return set_error("null params");
  if (g.sec.open) sector_close(g);
  free_eigvecs();
  return superc_open(g, p, sz);
}

int64_t edgpu_csr_nnz(void) { return g.csr.open ? g.csr.nnz : -1; }

int edgpu_set_sparse_h(int flag) {
  g.sparse_h = (flag != 0);
  return 0;
}

int edgpu_csr_get(int64_t *rowptr, int32_t *cols, double *vals) {
  clear_error();
  CsrSector &C = g.csr;
  if (!C.open) return set_error("no stored-H sector open");
  if (C.direct) return set_error("the open sector is direct (ED_SPARSE_H=F): no matrix is stored");
  EDGPU_CUDA(cudaMemcpy(rowptr, C.rowptr, sizeof(int64_t) * (C.nloc + 1), cudaMemcpyDeviceToHost));
  if (C.nnz) {
    EDGPU_CUDA(cudaMemcpy(cols, C.cols, sizeof(int32_t) * C.nnz, cudaMemcpyDeviceToHost));
    EDGPU_CUDA(cudaMemcpy(vals, C.vals, sizeof(double) * (C.cplx ? 2 : 1) * C.nnz, cudaMemcpyDeviceToHost));
    for (int64_t k = 0; k < C.nnz; k++) cols[k] += 1;  // the reference's 1-based columns
  }
  return 0;
}

int edgpu_sector_get_map(int spin, int32_t *map) {
  clear_error();
  if (g.csr.open && g.csr.map) {  // device-built nonsu2 sector: H(1)%map(DimEl), packed states
    EDGPU_CUDA(cudaMemcpy(map, g.csr.map, sizeof(int32_t) * g.csr.nglobal, cudaMemcpyDeviceToHost));
    return 0;
  }
  if (!g.sec.open) return set_error("no sector open");
  SpinSpace &S = spin == 0 ? g.sec.up : g.sec.dw;
  // the device map is in the internal enumeration order; report the reference's ascending one
  std::vector<int32_t> m((size_t)S.dim), ref((size_t)S.dim);
  EDGPU_CUDA(cudaMemcpy(m.data(), S.map, sizeof(int32_t) * S.dim, cudaMemcpyDeviceToHost));
  EDGPU_CUDA(cudaMemcpy(ref.data(), S.refidx, sizeof(int32_t) * S.dim, cudaMemcpyDeviceToHost));
  for (int64_t r = 0; r < S.dim; r++) map[ref[(size_t)r]] = m[(size_t)r];
  return 0;
}

// hop table of one species as (row-major) lists: entry s of row r lives in group s/4
static int download_hops(SpinSpace &S, std::vector<uint32_t> &ell, std::vector<double> &amp) {
  if (S.sharded)
    return set_error("hop tables of a species sharded over ranks hold rank-local targets; open the sector on one rank");
  const int G = std::max(S.Wl4 + S.Wf4, 1);
  ell.resize((size_t)4 * G * S.ld);
  amp.resize((size_t)2 * S.nterms + 2);
  EDGPU_CUDA(cudaMemcpy(ell.data(), S.ell4, sizeof(uint32_t) * ell.size(), cudaMemcpyDeviceToHost));
  EDGPU_CUDA(cudaMemcpy(amp.data(), S.amp2, sizeof(double) * amp.size(), cudaMemcpyDeviceToHost));
  return 0;
}

int64_t edgpu_sector_hop_count(int spin) {
  clear_error();
  if (!g.sec.open) {
    set_error("no sector open");
    return -1;
  }
  SpinSpace &S = spin == 0 ? g.sec.up : g.sec.dw;
  std::vector<uint32_t> ell;
  std::vector<double> amp;
  if (download_hops(S, ell, amp)) return -1;
  const int G = S.Wl4 + S.Wf4;
  int64_t n = 0;
  for (int gi = 0; gi < G; gi++)
    for (int64_t r = 0; r < S.dim; r++)
      for (int k = 0; k < 4; k++)
        if (((ell[((size_t)gi * S.ld + r) * 4 + k] >> HOP_AMP_SHIFT) & HOP_AMP_MASK) != (uint32_t)(2 * S.nterms)) n++;
  return n;
}

int edgpu_sector_get_hops(int spin, int64_t *rowptr, int32_t *target, double *value) {
  clear_error();
  if (!g.sec.open) return set_error("no sector open");
  SpinSpace &S = spin == 0 ? g.sec.up : g.sec.dw;
  std::vector<uint32_t> ell;
  std::vector<double> amp;
  EDGPU_TRY(download_hops(S, ell, amp));
  std::vector<int32_t> ref((size_t)S.dim), inv((size_t)S.dim);
  EDGPU_CUDA(cudaMemcpy(ref.data(), S.refidx, sizeof(int32_t) * S.dim, cudaMemcpyDeviceToHost));
  for (int64_t r = 0; r < S.dim; r++) inv[(size_t)ref[(size_t)r]] = (int32_t)r;
  const int G = S.Wl4 + S.Wf4;
  int64_t n = 0;
  for (int64_t rr = 0; rr < S.dim; rr++) {  // rows in the reference's order
    const int64_t r = inv[(size_t)rr];
    rowptr[rr] = n;
    for (int gi = 0; gi < G; gi++)
      for (int k = 0; k < 4; k++) {
        const uint32_t ent = ell[((size_t)gi * S.ld + r) * 4 + k];
        const uint32_t id = (ent >> HOP_AMP_SHIFT) & HOP_AMP_MASK;
        if (id == (uint32_t)(2 * S.nterms)) continue;
        int64_t t = (int64_t)(ent & HOP_TGT_MASK);
        if (S.block_mode && gi < S.Wl4) {  // local entries hold tile offsets of the row's work item
          const BlockItem *it = nullptr;
          for (const BlockItem &b : S.items)
            if (r >= b.out0 && r < b.out1) it = &b;
          int64_t gr = -1;
          for (int k = 0; it && k < it->nin; k++)
            if (t >= it->in_off[k] && t < it->in_off[k] + it->in_len[k]) gr = it->in0[k] + (t - it->in_off[k]);
          if (gr < 0) return set_error("internal: hop entry outside its work item's tile");
          t = gr;
        }
        target[n] = ref[(size_t)t] + 1;
        value[n] = amp[id];
        n++;
      }
  }
  rowptr[S.dim] = n;
  return 0;
}

int64_t edgpu_vec_padded_len(void) { return any_open() ? g.veclen() : 0; }
int edgpu_vec_upload(double *d_dst, const double *h_src) {
  clear_error();
  if (!any_open()) return set_error("no sector open");
  EDGPU_TRY(upload(g, d_dst, h_src));
  EDGPU_CUDA(cudaStreamSynchronize(g.stream));
  return 0;
}
int edgpu_vec_download(double *h_dst, const double *d_src) {
  clear_error();
  if (!any_open()) return set_error("no sector open");
  return download(g, h_dst, d_src);
}

int edgpu_hxv_dev(const double *d_v, double *d_Hv) {
  clear_error();
  return hxv_device(g, d_v, d_Hv, false, true);
}

static double *g_hx_in = nullptr, *g_hx_out = nullptr;
static int64_t g_hx_in_len = 0, g_hx_out_len = 0;

static void hxv_host(const char *who, bool want_cplx, const int32_t *Nloc, const double *v, double *Hv) {
  clear_error();
  if (!any_open()) {
    set_error("%s: no sector open (spHtimesV_p used outside build/delete_Hv_sector)", who);
    return;
  }
  const bool is_cplx = g.csr.open && g.csr.cplx;
  if (is_cplx != want_cplx) {
    set_error("%s: the open sector is %s", who, is_cplx ? "complex (use edgpu_hxv_z)" : "real (use edgpu_hxv_d)");
    return;
  }
  // "if(Nloc/=getdim(isector))stop" (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:52)
  if ((int64_t)*Nloc != edgpu_sector_vecdim()) {
    set_error("%s: Nloc=%d /= vecDim=%lld", who, (int)*Nloc, (long long)edgpu_sector_vecdim());
    return;
  }
  const int64_t n = g.veclen();
  if (ensure_buf(&g_hx_in, &g_hx_in_len, n) || ensure_buf(&g_hx_out, &g_hx_out_len, n)) return;
  if (upload(g, g_hx_in, v)) return;
  if (hxv_device(g, g_hx_in, g_hx_out, false, false)) return;
  if (download(g, Hv, g_hx_out)) return;
  comm_pipe_check(g);
}

void edgpu_hxv_d(const int32_t *Nloc, const double *v, double *Hv) {
  hxv_host("edgpu_hxv_d", false, Nloc, v, Hv);
}
void edgpu_hxv_z(const int32_t *Nloc, const double *v_re_im, double *Hv_re_im) {
  hxv_host("edgpu_hxv_z", true, Nloc, v_re_im, Hv_re_im);
}

int edgpu_status(void) { return g_status; }

int edgpu_host_register(void *ptr, int64_t bytes) {
  clear_error();
  if (!g.inited) return set_error("edgpu_init was not called");
  if (!ptr || bytes <= 0) return 0;
  EDGPU_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
  return 0;
}
int edgpu_host_unregister(void *ptr) {
  clear_error();
  if (!ptr) return 0;
  EDGPU_CUDA(cudaHostUnregister(ptr));
  return 0;
}
const char *edgpu_last_error(void) { return g_errbuf; }

int64_t edgpu_launch_count(int reset) {
  int64_t n = g_launches;
  if (reset) g_launches = 0;
  return n;
}

int edgpu_last_hxv_stage_ms(float *out4) {
  for (int i = 0; i < 4; i++) out4[i] = g.stage_ms[i];
  return 0;
}

void *edgpu_stream(void) { return (void *)g.stream; }

int edgpu_profile_begin(int max_steps) {
  clear_error();
  if (!g.inited) return set_error("edgpu_init was not called");
  if (max_steps < 1) max_steps = 1;
  while ((int)g.prof_ev.size() < 4 * max_steps) {
    cudaEvent_t e;
    EDGPU_CUDA(cudaEventCreate(&e));
    g.prof_ev.push_back(e);
  }
  g.prof_cap = max_steps;
  g.prof_n = 0;
  g.prof_on = true;
  return 0;
}

int edgpu_profile_end(float *ms3, int *nsteps) {
  clear_error();
  g.prof_on = false;
  ms3[0] = ms3[1] = ms3[2] = 0.f;
  *nsteps = g.prof_n;
  if (g.prof_n == 0) return 0;
  EDGPU_CUDA(cudaEventSynchronize(g.prof_ev[(size_t)4 * (g.prof_n - 1) + 3]));
  for (int i = 0; i < g.prof_n; i++)
    for (int k = 0; k < 3; k++) {
      float ms = 0.f;
      EDGPU_CUDA(cudaEventElapsedTime(&ms, g.prof_ev[(size_t)4 * i + k], g.prof_ev[(size_t)4 * i + k + 1]));
      ms3[k] += ms;
    }
  return 0;
}

int edgpu_set_kernel_variant(int variant) {
  g.variant_request = variant;
  if (g.sec.open) g.sec.variant = variant;
  return 0;
}

int edgpu_lanczos_gs(int nitermax, double threshold, int ncheck, int use_start, uint64_t seed,
                     double *egs, double *vec_host, int *niter) {
  clear_error();
  if (!any_open()) return set_error("no sector open");
  const int64_t n = g.veclen();
  EDGPU_TRY(ensure_buf(&g_current, &g_current_len, n));
  double *d_start = nullptr;
  if (use_start) {
    if (!vec_host) return set_error("use_start set but vec_host is NULL");
    EDGPU_TRY(lanczos_work(g, 2, n, &d_start));  // cached: a cudaFree per solve costs up to 0.5 s
    EDGPU_TRY(upload(g, d_start, vec_host));
  }
  int rc = lanczos_gs_dev(g, nitermax, threshold, ncheck, d_start, seed, egs, g_current, niter);
  if (rc) return rc;
  if (vec_host) EDGPU_TRY(download(g, vec_host, g_current));
  return 0;
}

int edgpu_lanczos_last_info(int *nstored, int *nhxv) {
  if (nstored) *nstored = g_lanczos_last_stored;
  if (nhxv) *nhxv = g_lanczos_last_hxv;
  return 0;
}

int edgpu_lanczos_tridiag(const double *seed_host, int nlanc, double threshold, double *alanc,
                          double *blanc, int *nused, double *norm2) {
  clear_error();
  if (!any_open()) return set_error("no sector open");
  const int64_t n = g.veclen();
  if (seed_host) {
    EDGPU_TRY(ensure_buf(&g_seed, &g_seed_len, n));
    EDGPU_TRY(upload(g, g_seed, seed_host));
  } else if (!g_seed || g_seed_len < n) {
    return set_error("no device-resident seed: call edgpu_apply_op first or pass seed_host");
  }
  // norm2 = <vvinit|vvinit>; vvinit /= sqrt(norm2) (ED_HAMILTONIAN_NORMAL.f90:344-347)
  double n2 = 0.0;
  EDGPU_TRY(vec_dot(g, g_seed, g_seed, &n2));
  if (norm2) *norm2 = n2;
  for (int i = 0; i < nlanc; i++) alanc[i] = blanc[i] = 0.0;
  *nused = 0;
  if (n2 == 0.0) return 0;  // "if(norm2/=0d0)" (:355)
  const int64_t dim_global = edgpu_sector_dim();
  if (nlanc > dim_global) nlanc = (int)dim_global;
  double *work = nullptr;
  EDGPU_CUDA(cudaMalloc(&work, sizeof(double) * n));
  int rc = lanczos_tridiag_dev(g, g_seed, work, nlanc, threshold, alanc, blanc, nused);
  cudaFree(work);
  return rc;
}

// es_add_state (ED_EIGENSPACE.f90): keeps a copy of the device vector `src` of the open sector
static int store_state_from(const double *src, int slot) {
  edgpu_state_free(slot);
  StoredState st;
  if (g.csr.open) {
    CsrSector &C = g.csr;
    if (C.pk_mode < 0) return set_error("states of host-supplied stored-H sectors are not kept on the device");
    st.kind = 1;
    st.pk_mode = C.pk_mode;
    st.pk_qn = C.pk_qn;
    st.Ns = C.pk_Ns;
    st.nglobal = C.nglobal;
    st.row0 = C.row0;
    st.nloc = C.nloc;
    st.padded = C.padded_len();
    EDGPU_CUDA(cudaMalloc(&st.vec, sizeof(double) * st.padded));
    EDGPU_CUDA(cudaMemcpyAsync(st.vec, src, sizeof(double) * st.padded, cudaMemcpyDeviceToDevice, g.stream));
    EDGPU_CUDA(cudaStreamSynchronize(g.stream));
    g_states[slot] = st;
    return 0;
  }
  Sector &S = g.sec;
  st.Ns = S.Ns;
  st.nup = S.up.nel;
  st.ndw = S.dw.nel;
  st.dimu = S.up.dim;
  st.dimd = S.dw.dim;
  st.ldu = S.up.ld;
  st.qdw = S.qdw;
  st.d0 = S.d0;
  st.dimph = S.DimPh;
  const int64_t n = S.padded_len();
  EDGPU_CUDA(cudaMalloc(&st.vec, sizeof(double) * n));
  EDGPU_CUDA(cudaMemcpyAsync(st.vec, src, sizeof(double) * n, cudaMemcpyDeviceToDevice, g.stream));
  EDGPU_CUDA(cudaStreamSynchronize(g.stream));
  g_states[slot] = st;
  return 0;
}

int edgpu_state_store(int slot) {
  clear_error();
  if (!any_open()) return set_error("no sector open");
  if (!g_current) return set_error("no current state (run edgpu_lanczos_gs first)");
  return store_state_from(g_current, slot);
}

int edgpu_eigh(int neigen, int nblock, int nitermax, double tol, uint64_t seed, double *evals,
               double *evecs_host, int *nconv, int *nmatvec) {
  clear_error();
  if (!any_open()) return set_error("no sector open");
  if (!evals) return set_error("edgpu_eigh: evals is NULL");
  free_eigvecs();
  std::vector<double> resid((size_t)std::max(neigen, 1));
  int nc = 0, nm = 0;
  EDGPU_TRY(eigh_dev(g, neigen, nblock, nitermax, tol, seed, evals, resid.data(), &g_eigvecs, &nc, &nm));
  if (nconv) *nconv = nc;
  if (nmatvec) *nmatvec = nm;
  if (evecs_host) {
    const int64_t nloc = edgpu_sector_vecdim() * ((g.csr.open && g.csr.cplx) ? 2 : 1);
    for (size_t i = 0; i < g_eigvecs.size(); i++)
      EDGPU_TRY(download(g, evecs_host + (int64_t)i * nloc, g_eigvecs[i]));
  }
  return 0;
}

int edgpu_eigh_state_store(int k, int slot) {
  clear_error();
  if (!any_open()) return set_error("no sector open");
  if (k < 0 || k >= (int)g_eigvecs.size())
    return set_error("eigenvector %d not available (last edgpu_eigh kept %d)", k, (int)g_eigvecs.size());
  return store_state_from(g_eigvecs[k], slot);
}

int edgpu_state_free(int slot) {
  auto it = g_states.find(slot);
  if (it != g_states.end()) {
    cudaFree(it->second.vec);
    g_states.erase(it);
  }
  return 0;
}

int edgpu_apply_ops_packed(int slot, int nops, const double *coef_re_im, const int *op, const int *iorb,
                           const int *spin) {
  clear_error();
  CsrSector &C = g.csr;
  if (!C.open || C.pk_mode < 0) return set_error("no device-built nonsu2/superc sector open");
  auto it = g_states.find(slot);
  if (it == g_states.end()) return set_error("state slot %d is empty", slot);
  StoredState &st = it->second;
  if (st.kind != 1 || st.pk_mode != C.pk_mode || st.Ns != C.pk_Ns)
    return set_error("state %d does not belong to the open sector's mode / model", slot);
  if (nops < 1 || nops > 4) return set_error("apply_ops: 1..4 operators");
  PackedOps ops;
  ops.n = nops;
  const int No = C.pk_Ns;  // bound check below uses Norb through iorb < Ns only; model-level check is the host's
  for (int k = 0; k < nops; k++) {
    if (op[k] != 1 && op[k] != -1) return set_error("op must be +1 (CDG) or -1 (C)");
    if (spin[k] != 0 && spin[k] != 1) return set_error("spin must be 0 or 1");
    if (iorb[k] < 0 || iorb[k] >= No) return set_error("iorb out of range");
    ops.bit[k] = iorb[k] + spin[k] * C.pk_Ns;
    ops.create[k] = op[k] > 0;
    ops.cre[k] = coef_re_im[2 * k];
    ops.cim[k] = coef_re_im[2 * k + 1];
    // quantum number of the target: Ntot +- 1 (nonsu2), Sz +- 1 with the sign of the spin (superc)
    const int dq = C.pk_mode == 0 ? op[k] : (spin[k] == 0 ? op[k] : -op[k]);
    if (st.pk_qn + dq != C.pk_qn)
      return set_error("operator %d maps the state's sector (%d) to %d, not to the open sector (%d)", k,
                       st.pk_qn, st.pk_qn + dq, C.pk_qn);
  }
  const int64_t n = C.padded_len();
  EDGPU_TRY(ensure_buf(&g_seed, &g_seed_len, n));
  // the whole source vector is needed (es_return_cvec gathers it in the reference)
  const double *vsrc = st.vec;
  double *vfull = nullptr;
  if (g.nranks > 1) {
    std::vector<int64_t> counts(g.nranks), offs(g.nranks);
    const int64_t q = st.nglobal / g.nranks;
    for (int p = 0; p < g.nranks; p++) {
      counts[p] = 2 * (q + (p == g.nranks - 1 ? st.nglobal % g.nranks : 0));
      offs[p] = 2 * q * p;
    }
    EDGPU_CUDA(cudaMalloc(&vfull, sizeof(double) * 2 * (size_t)st.nglobal));
    EDGPU_TRY(comm_allgatherv(g, st.vec, vfull, counts, offs));
    vsrc = vfull;
  }
  int rc = packed_apply_ops(g, ops, st.pk_mode, st.pk_qn, vsrc, g_seed);
  cudaFree(vfull);
  return rc;
}

int edgpu_seed_norm2(double *norm2) {
  clear_error();
  if (!any_open()) return set_error("no sector open");
  if (!g_seed || g_seed_len < g.veclen()) return set_error("no device-resident seed");
  return vec_dot(g, g_seed, g_seed, norm2);
}

int edgpu_apply_op(int slot, int op, int iorb, int spin) {
  clear_error();
  if (g.csr.open && g.csr.pk_mode >= 0) {
    const double one[2] = {1.0, 0.0};
    return edgpu_apply_ops_packed(slot, 1, one, &op, &iorb, &spin);
  }
  if (!g.sec.open) return set_error("no sector open");
  auto it = g_states.find(slot);
  if (it == g_states.end()) return set_error("state slot %d is empty", slot);
  StoredState &st = it->second;
  if (st.kind != 0) return set_error("state %d belongs to a nonsu2/superc sector", slot);
  Sector &S = g.sec;
  if (op != 1 && op != -1) return set_error("op must be +1 (CDG) or -1 (C)");
  if (spin != 0 && spin != 1) return set_error("spin must be 0 or 1");
  if (iorb < 0 || iorb >= S.Norb) return set_error("iorb out of range");
  const int tnup = st.nup + (spin == 0 ? op : 0), tndw = st.ndw + (spin == 1 ? op : 0);
  if (S.Ns != st.Ns || S.up.nel != tnup || S.dw.nel != tndw)
    return set_error("open sector (%d,%d) is not the target sector (%d,%d) of the operator",
                     S.up.nel, S.dw.nel, tnup, tndw);
  if (st.dimph != S.DimPh)
    return set_error("state %d has %d phonon slices, the open sector %d", slot, st.dimph, S.DimPh);
  // enumeration order + ranking tables of the operated species in the SOURCE sector
  const int nel_src = spin == 0 ? st.nup : st.ndw;
  int32_t *map = nullptr;
  LinTable lin;
  SiteOrder ord;
  EDGPU_TRY(species_ranking(g, S.prm, spin, nel_src, &map, &lin, &ord));
  const int64_t n = S.padded_len();
  EDGPU_TRY(ensure_buf(&g_seed, &g_seed_len, n));
  EDGPU_CUDA(cudaMemsetAsync(g_seed, 0, sizeof(double) * n, g.stream));
  // An operator on the dw species moves amplitude between dw columns, which live on different
  // ranks in the source and in the target sector (their DimDw differ): gather the stored state
  // first, like the reference does before apply_op (es_return_dvec, ED_EIGENSPACE.f90:723-793).
  // Electronic operators act on every phonon slice alike (ED_SECTOR.f90:489-491: iph loop).
  const double *vsrc = st.vec;
  double *vfull = nullptr;
  int64_t src_slice = st.ldu * st.qdw;
  if (g.nranks > 1 && spin == 1) {
    std::vector<int64_t> counts(g.nranks), offs(g.nranks);
    for (int p = 0; p < g.nranks; p++) {
      int64_t q, d0;
      block_split(st.dimd, g.nranks, p, &q, &d0);
      counts[p] = q * st.ldu;
      offs[p] = d0 * st.ldu;
    }
    const int64_t full_slice = st.dimd * st.ldu;
    EDGPU_CUDA(cudaMalloc(&vfull, sizeof(double) * (size_t)full_slice * (size_t)st.dimph));
    for (int iph = 0; iph < st.dimph; iph++)
      EDGPU_TRY(comm_allgatherv(g, st.vec + iph * src_slice, vfull + iph * full_slice, counts, offs));
    vsrc = vfull;
    src_slice = full_slice;
  }
  dim3 grid((unsigned)((S.up.dim + 127) / 128), (unsigned)S.qdw);
  for (int iph = 0; iph < S.DimPh && S.qdw > 0; iph++) {
    k_apply_op<<<grid, 128, 0, g.stream>>>(vsrc + iph * src_slice, st.ldu, g_seed + iph * S.slice_len(),
                                           S.up.ld, S.up.dim, S.qdw,
                                           spin == 0 ? S.up.map : S.dw.map + S.d0, op, iorb, spin,
                                           rank_view(lin, ord));
    EDGPU_COUNT_LAUNCH();
  }
  EDGPU_CUDA(cudaGetLastError());
  EDGPU_CUDA(cudaStreamSynchronize(g.stream));
  cudaFree(vfull);
  cudaFree(map);
  cudaFree(lin.ja);
  cudaFree(lin.jb);
  return 0;
}

int edgpu_apply_ops_normal(int slot, int nops, const double *coef, int op, const int *iorb, int spin) {
  clear_error();
  if (!g.sec.open) return set_error("no NORMAL sector open");
  auto it = g_states.find(slot);
  if (it == g_states.end()) return set_error("state slot %d is empty", slot);
  StoredState &st = it->second;
  if (st.kind != 0) return set_error("state %d belongs to a nonsu2/superc sector", slot);
  Sector &S = g.sec;
  if (nops < 1 || nops > EDGPU_MAXORB || !coef || !iorb) return set_error("apply_ops_normal: 1..%d operators", EDGPU_MAXORB);
  if (op != 1 && op != -1) return set_error("op must be +1 (CDG) or -1 (C)");
  if (spin != 0 && spin != 1) return set_error("spin must be 0 or 1");
  for (int k = 0; k < nops; k++)
    if (iorb[k] < 0 || iorb[k] >= S.Norb) return set_error("iorb out of range");
  const int tnup = st.nup + (spin == 0 ? op : 0), tndw = st.ndw + (spin == 1 ? op : 0);
  if (S.Ns != st.Ns || S.up.nel != tnup || S.dw.nel != tndw)
    return set_error("open sector (%d,%d) is not the target sector (%d,%d) of the operators", S.up.nel,
                     S.dw.nel, tnup, tndw);
  if (st.dimph != S.DimPh)
    return set_error("state %d has %d phonon slices, the open sector %d", slot, st.dimph, S.DimPh);
  const int nel_src = spin == 0 ? st.nup : st.ndw;
  int32_t *map = nullptr;
  LinTable lin;
  SiteOrder ord;
  EDGPU_TRY(species_ranking(g, S.prm, spin, nel_src, &map, &lin, &ord));
  const int64_t n = S.padded_len();
  EDGPU_TRY(ensure_buf(&g_seed, &g_seed_len, n));
  EDGPU_CUDA(cudaMemsetAsync(g_seed, 0, sizeof(double) * n, g.stream));
  const double *vsrc = st.vec;
  double *vfull = nullptr;
  int64_t src_slice = st.ldu * st.qdw;
  if (g.nranks > 1 && spin == 1) {  // as edgpu_apply_op: the dw operators need every source column
    std::vector<int64_t> counts(g.nranks), offs(g.nranks);
    for (int p = 0; p < g.nranks; p++) {
      int64_t q, d0;
      block_split(st.dimd, g.nranks, p, &q, &d0);
      counts[p] = q * st.ldu;
      offs[p] = d0 * st.ldu;
    }
    const int64_t full_slice = st.dimd * st.ldu;
    EDGPU_CUDA(cudaMalloc(&vfull, sizeof(double) * (size_t)full_slice * (size_t)st.dimph));
    for (int iph = 0; iph < st.dimph; iph++)
      EDGPU_TRY(comm_allgatherv(g, st.vec + iph * src_slice, vfull + iph * full_slice, counts, offs));
    vsrc = vfull;
    src_slice = full_slice;
  }
  dim3 grid((unsigned)((S.up.dim + 127) / 128), (unsigned)S.qdw);
  for (int k = 0; k < nops && S.qdw > 0; k++)
    for (int iph = 0; iph < S.DimPh; iph++) {
      k_apply_op_acc<<<grid, 128, 0, g.stream>>>(vsrc + iph * src_slice, st.ldu, g_seed + iph * S.slice_len(),
                                                 S.up.ld, S.up.dim, S.qdw,
                                                 spin == 0 ? S.up.map : S.dw.map + S.d0, op, iorb[k], spin,
                                                 rank_view(lin, ord), coef[k]);
      EDGPU_COUNT_LAUNCH();
    }
  EDGPU_CUDA(cudaGetLastError());
  EDGPU_CUDA(cudaStreamSynchronize(g.stream));
  cudaFree(vfull);
  cudaFree(map);
  cudaFree(lin.ja);
  cudaFree(lin.jb);
  return 0;
}

int edgpu_state_download(int slot, double *vec_host) {
  clear_error();
  auto it = g_states.find(slot);
  if (it == g_states.end()) return set_error("state slot %d is empty", slot);
  if (!vec_host) return set_error("state_download: null buffer");
  const StoredState &st = it->second;
  if (st.kind == 1) {
    CsrSector &C = g.csr;
    if (!C.open || C.pk_mode != st.pk_mode || C.pk_qn != st.pk_qn || C.pk_Ns != st.Ns)
      return set_error("the state's own sector must be open for state_download");
  } else {
    Sector &S = g.sec;
    if (!S.open || S.up.nel != st.nup || S.dw.nel != st.ndw || S.Ns != st.Ns || S.DimPh != st.dimph)
      return set_error("the state's own sector must be open for state_download");
  }
  return download(g, vec_host, st.vec);
}

int edgpu_state_twin(int src_slot, int dst_slot) {
  clear_error();
  if (src_slot == dst_slot) return set_error("state_twin: source and destination slots coincide");
  auto it = g_states.find(src_slot);
  if (it == g_states.end()) return set_error("state slot %d is empty", src_slot);
  StoredState st = it->second;  // copy: store_state_from below may rehash the map
  if (st.kind == 1) {
    CsrSector &C = g.csr;
    if (!C.open || C.pk_mode != st.pk_mode || C.pk_Ns != st.Ns)
      return set_error("state_twin: the open sector must be the device-built twin sector of state %d", src_slot);
    // get_twin_sector (ED_SECTOR.f90:1826-1843): Sz -> -Sz, Ntot -> Nlevels - Ntot
    const int want = st.pk_mode == 1 ? -st.pk_qn : 2 * st.Ns - st.pk_qn;
    if (C.pk_qn != want)
      return set_error("state_twin: open sector has quantum number %d, the twin of %d is %d", C.pk_qn, st.pk_qn, want);
    const int64_t n = C.padded_len();
    double *tmp = nullptr, *vfull = nullptr;
    EDGPU_CUDA(cudaMalloc(&tmp, sizeof(double) * n));
    const double *vsrc = st.vec;
    if (g.nranks > 1) {  // the flipped states of this rank's rows live anywhere in the source
      std::vector<int64_t> counts(g.nranks), offs(g.nranks);
      const int64_t q = st.nglobal / g.nranks;
      for (int p = 0; p < g.nranks; p++) {
        counts[p] = 2 * (q + (p == g.nranks - 1 ? st.nglobal % g.nranks : 0));
        offs[p] = 2 * q * p;
      }
      EDGPU_CUDA(cudaMalloc(&vfull, sizeof(double) * 2 * (size_t)st.nglobal));
      EDGPU_TRY(comm_allgatherv(g, st.vec, vfull, counts, offs));
      vsrc = vfull;
    }
    int rc = packed_twin(g, st.pk_mode, st.pk_qn, vsrc, tmp);
    if (!rc) rc = store_state_from(tmp, dst_slot);
    cudaFree(vfull);
    cudaFree(tmp);
    return rc;
  }
  Sector &S = g.sec;
  if (!S.open || S.Ns != st.Ns || S.up.nel != st.ndw || S.dw.nel != st.nup)
    return set_error("state_twin: the open sector must be (%d,%d), the twin of state %d's sector (%d,%d)",
                     st.ndw, st.nup, src_slot, st.nup, st.ndw);
  if (st.dimph != S.DimPh)
    return set_error("state %d has %d phonon slices, the open sector %d", src_slot, st.dimph, S.DimPh);
  // enumeration orders + ranking tables of BOTH species of the source sector
  int32_t *mapu = nullptr, *mapd = nullptr;
  LinTable linu, lind;
  SiteOrder ordu, ordd;
  EDGPU_TRY(species_ranking(g, S.prm, 0, st.nup, &mapu, &linu, &ordu));
  EDGPU_TRY(species_ranking(g, S.prm, 1, st.ndw, &mapd, &lind, &ordd));
  const int64_t n = S.padded_len();
  double *tmp = nullptr, *vfull = nullptr;
  EDGPU_CUDA(cudaMalloc(&tmp, sizeof(double) * n));
  EDGPU_CUDA(cudaMemsetAsync(tmp, 0, sizeof(double) * n, g.stream));
  const double *vsrc = st.vec;
  int64_t src_slice = st.ldu * st.qdw;
  if (g.nranks > 1) {  // B's local columns are A's rows: every source column is needed
    std::vector<int64_t> counts(g.nranks), offs(g.nranks);
    for (int p = 0; p < g.nranks; p++) {
      int64_t q, d0;
      block_split(st.dimd, g.nranks, p, &q, &d0);
      counts[p] = q * st.ldu;
      offs[p] = d0 * st.ldu;
    }
    const int64_t full_slice = st.dimd * st.ldu;
    EDGPU_CUDA(cudaMalloc(&vfull, sizeof(double) * (size_t)full_slice * (size_t)st.dimph));
    for (int iph = 0; iph < st.dimph; iph++)
      EDGPU_TRY(comm_allgatherv(g, st.vec + iph * src_slice, vfull + iph * full_slice, counts, offs));
    vsrc = vfull;
    src_slice = full_slice;
  }
  dim3 grid((unsigned)((S.up.dim + 127) / 128), (unsigned)S.qdw);
  for (int iph = 0; iph < S.DimPh && S.qdw > 0; iph++) {
    k_twin_normal<<<grid, 128, 0, g.stream>>>(vsrc + iph * src_slice, st.ldu, tmp + iph * S.slice_len(),
                                              S.up.ld, S.up.dim, S.up.map, S.dw.map + S.d0,
                                              rank_view(linu, ordu), rank_view(lind, ordd));
    EDGPU_COUNT_LAUNCH();
  }
  EDGPU_CUDA(cudaGetLastError());
  EDGPU_CUDA(cudaStreamSynchronize(g.stream));
  int rc = store_state_from(tmp, dst_slot);
  cudaFree(tmp);
  cudaFree(vfull);
  cudaFree(mapu);
  cudaFree(mapd);
  cudaFree(linu.ja);
  cudaFree(linu.jb);
  cudaFree(lind.ja);
  cudaFree(lind.jb);
  return rc;
}

int edgpu_state_observables(int slot, double *dens, double *docc) {
  clear_error();
  auto it = g_states.find(slot);
  if (it == g_states.end()) return set_error("state slot %d is empty", slot);
  StoredState &st = it->second;
  if (st.kind == 1) {
    CsrSector &C = g.csr;
    if (!C.open || C.pk_mode != st.pk_mode || C.pk_qn != st.pk_qn || C.pk_Ns != st.Ns)
      return set_error("the state's own sector must be open for observables");
    return packed_observables(g, st.vec, dens, docc);
  }
  Sector &S = g.sec;
  if (!S.open || S.up.nel != st.nup || S.dw.nel != st.ndw || S.Ns != st.Ns)
    return set_error("the state's own sector must be open for observables");
  double *d_out = g.d_scal + 8;
  EDGPU_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * 2 * EDGPU_MAXORB, g.stream));
  dim3 grid((unsigned)std::min<int64_t>((S.up.dim + 255) / 256, 64),
            (unsigned)std::min<int64_t>(S.qdw, 1024));
  for (int iph = 0; iph < st.dimph && S.qdw > 0; iph++) {  // the occupations do not see the phonon index
    k_observables<<<grid, 256, 0, g.stream>>>(st.vec + iph * st.ldu * st.qdw, st.ldu, S.up.dim, S.qdw, S.d0,
                                              S.up.imp, S.dw.imp, S.Norb, d_out);
    EDGPU_COUNT_LAUNCH();
  }
  EDGPU_TRY(comm_allreduce_sum(g, d_out, 2 * EDGPU_MAXORB));
  EDGPU_CUDA(cudaMemcpyAsync(g.h_scal + 8, d_out, sizeof(double) * 2 * EDGPU_MAXORB,
                             cudaMemcpyDeviceToHost, g.stream));
  EDGPU_CUDA(cudaStreamSynchronize(g.stream));
  for (int a = 0; a < S.Norb; a++) {
    dens[a] = g.h_scal[8 + a];
    docc[a] = g.h_scal[8 + EDGPU_MAXORB + a];
  }
  return 0;
}

}  // extern "C"
