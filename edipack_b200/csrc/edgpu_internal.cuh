// Internal declarations of the edgpu engine (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/edgpu.h"

namespace edgpu {

// ---------------------------------------------------------------------------------------
// error latch (the reference aborts with `stop "msg"`; the ABI returns codes instead)
// ---------------------------------------------------------------------------------------
int set_error(const char *fmt, ...);
void clear_error();
extern int g_status;

#define EDGPU_CUDA(call)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::edgpu::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),   \
                                __FILE__, __LINE__);                                      \
  } while (0)

#define EDGPU_TRY(call)        \
  do {                         \
    int _rc = (call);          \
    if (_rc != 0) return _rc;  \
  } while (0)

extern int64_t g_launches;  // kernels launched by this library (bench "gpu_launches")
#define EDGPU_COUNT_LAUNCH() (++::edgpu::g_launches)

// ---------------------------------------------------------------------------------------
// hop-table entry packing: [19:0] target row, [30:20] signed amplitude index
// (2*term + sign; amp2[2t] = +h_t, amp2[2t+1] = -h_t, amp2[2*nterms] = 0 for padding),
// [31] "far" flag (slow-role tables only: the target lies outside the row's range)
// ---------------------------------------------------------------------------------------
constexpr int EDGPU_MAXRANKS = 16;
constexpr uint32_t HOP_TGT_MASK = 0xFFFFFu;
constexpr int HOP_AMP_SHIFT = 20;
constexpr uint32_t HOP_AMP_MASK = 0x7FFu;
constexpr uint32_t HOP_FAR = 0x80000000u;
constexpr int HOP_MAX_TERMS = 1022;

struct Term {  // one directed one-body term  h * c^+_alpha c_beta  (bit positions, 0-based)
  int32_t alpha, beta;
  double h;
};

// Internal enumeration order of one species: the sector states are listed in ascending order
// of the bit-PERMUTED Fock integer (site `site[p]` sits at permuted position p, p = Ns-1 most
// significant), not of the Fock integer itself.  The permutation puts a few bath sites on top
// (range bits), then the impurity sites, then the other bath sites: consecutive states then
// share their hop structure ("runs"), so that the gathers of a warp hit consecutive rows.
// The reference's ascending order (ED_SECTOR.f90:217-242) is restored at the host boundary
// through refidx[] (internal index -> reference index).
struct SiteOrder {
  uint8_t pos[32];   // original bit b -> permuted position
  uint8_t site[32];  // permuted position p -> original bit
  int32_t Ns;
  int32_t identity;
};

// Lin two-table ranking on the permuted integer mp: rank = ja[mp >> lo_bits] + jb[mp & lo_mask]
struct LinTable {
  int lo_bits = 0;
  int32_t *ja = nullptr;  // [2^(Ns-lo_bits)]
  int32_t *jb = nullptr;  // [2^lo_bits]
};

// device view of the ranking function of a species: ORIGINAL Fock integer -> internal index
struct RankView {
  int lo_bits;
  const int32_t *ja, *jb;
  SiteOrder ord;
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t permute_bits(uint32_t m, const SiteOrder &o) {
  if (o.identity) return m;
  uint32_t r = 0;
  for (int b = 0; b < o.Ns; b++) r |= ((m >> b) & 1u) << o.pos[b];
  return r;
}
__device__ __forceinline__ int rank_of(uint32_t m, const RankView &R) {
  const uint32_t mp = permute_bits(m, R.ord);
  return R.ja[mp >> R.lo_bits] + R.jb[mp & ((1u << R.lo_bits) - 1u)];
}
#endif

enum SpinRole { ROLE_FAST = 0, ROLE_SLOW = 1 };

// Block mode of the fast role: one CTA work item = the states of one range prefix AND one
// impurity configuration ("block", contiguous in the internal order).  Every one-body term maps
// a block onto one other block of the same prefix (its impurity bits flip by a fixed mask), so
// the CTA stages only the blocks its hops read from: for imp<->bath hops that is the other half
// of the range (bipartite structure) and the tile is half as large as the range.
constexpr int BLK_MAXIN = 16;
struct BlockItem {
  int32_t out0, out1;         // output rows [out0, out1)
  int32_t nin;                // number of input blocks
  int32_t tile_rows;          // rows of the tile (input blocks + alignment spares), even
  int32_t in0[BLK_MAXIN];     // first row of input block k
  int32_t in_len[BLK_MAXIN];
  int32_t in_off[BLK_MAXIN];  // its offset inside the tile
};
#ifdef __CUDACC__
// tile offset of global row t for work item `it` (-1: not staged)
__device__ __forceinline__ int block_tile_offset(const BlockItem &it, int t) {
  for (int k = 0; k < it.nin; k++) {
    const int d = t - it.in0[k];
    if ((unsigned)d < (unsigned)it.in_len[k]) return it.in_off[k] + d;
  }
  return -1;
}
__device__ __forceinline__ int block_global_row(const BlockItem &it, int off) {
  for (int k = 0; k < it.nin; k++) {
    const int d = off - it.in_off[k];
    if ((unsigned)d < (unsigned)it.in_len[k]) return it.in0[k] + d;
  }
  return -1;
}
#endif
constexpr int SLOW_ROWS = 16;  // fast-index rows per CTA of the slow-index kernel (128 B segments)

// One spin species of the open sector
struct SpinSpace {
  int nel = 0;
  int64_t dim = 0;         // DimUp or DimDw
  int64_t ld = 0;          // dim padded to a multiple of 16 (128 B)
  SiteOrder ord;
  int32_t *map = nullptr;     // [ld] internal index -> Fock integer (original bit order)
  int32_t *refidx = nullptr;  // [ld] internal index -> reference (ascending Fock) index
  LinTable lin;
  double *eps = nullptr;   // [ld] single-spin diagonal energy
  uint8_t *imp = nullptr;  // [ld] impurity occupation bits (map & (2^Norb-1))
  // Ranges: contiguous runs of states sharing a prefix of top bits, each small enough for the
  // shared-memory tile of the species' role.  A hop is "local" when it stays inside its range
  // and "far" when it moves an electron into / out of the range's prefix bits.
  int role = ROLE_FAST;
  int nranges = 1;
  int64_t max_range = 0;
  std::vector<int64_t> range_start;  // host, nranges+1
  std::vector<int> range_tbits;      // host, prefix length of each range
  int64_t *d_range_start = nullptr;
  // block mode (fast role): local entries of the hop table hold TILE OFFSETS of their work item
  bool block_mode = false;
  std::vector<BlockItem> items;
  BlockItem *d_items = nullptr;
  int32_t *d_item_of_row = nullptr;  // [ld]
  int64_t max_tile = 0;              // largest tile_rows
  // hop table, ELL in groups of 4 entries: ell4[g*ld + row] (uint4).
  //   fast role: groups [0,Wl4) local, [Wl4, Wl4+Wf4) far; unused slots = padding entries
  //   slow role: one list of Wl4 groups (Wf4 = 0), local entries first, far entries flagged
  // padding entries point at the row itself with the zero amplitude.
  int Wl = 0, Wf = 0, Wl4 = 0, Wf4 = 0;
  int nterms = 0;
  uint4 *ell4 = nullptr;
  double *amp2 = nullptr;  // [2*nterms+2]
  // impurity-impurity hop table for the non-local terms (Norb>1): imphop[(a*Norb+b)*ld + row] =
  // target row of c^+_a c_b | sign << 31, or -1 when the hop is not allowed on that state
  int32_t *imphop = nullptr;
  std::vector<Term> terms; // host copy
  // Sharded slow role (the dw species with nranks > 1, halo mode): the table rows of the LOCAL
  // columns [shard0, shard0 + shard_q) hold LOCAL targets: a column index < shard_q of the rank's
  // chunk, or shard_q + slot of the halo (copies of the remote columns the chunk's hops read,
  // pushed by their owners before pass A).  Ranges are in local column coordinates.
  bool sharded = false;
  int64_t shard0 = 0, shard_q = 0;
  int64_t nhalo = 0;
  std::vector<int32_t> halo_cols;   // host: global column of halo slot k (ascending)
  int32_t *d_colmap = nullptr;      // [ld] global column -> local column / shard_q + halo slot / -1
};

// one coulomb_sundry line in application order (c_l, cd_j, c_k, cd_i): bit of the species
// integer, species (0 up / 1 dw), creation flag
struct SundryDev {
  int8_t bit[4], dw[4], create[4];
  int32_t pad;
  double U;
};

struct Sector {
  bool open = false;
  edgpu_normal_params prm;
  int Ns = 0, Norb = 0;
  SpinSpace up, dw;
  // dw split of ED_HAMILTONIAN_NORMAL.f90:128-142
  int64_t qdw = 0, d0 = 0;  // local dw columns [d0, d0+qdw)
  int64_t qup = 0, u0 = 0;  // local up rows of the transposed layout
  double *xud = nullptr;    // [2^Norb (dw imp)][2^Norb (up imp)] cross interaction + constants
  bool nonlocal = false;
  double *jx = nullptr, *jp = nullptr;  // [Norb*Norb] device copies
  int variant = 0;
  // scratch for the distributed transposes: vt = this rank's rows of v^T ([DimDw (fast) x qup]),
  // hvt = Hdw vt, hvr = landing zone of the returned Hdw part (shape of the local vector)
  double *vt = nullptr, *hvt = nullptr, *hvr = nullptr, *sendbuf = nullptr, *recvbuf = nullptr;
  // peer-memory pipeline (comm.cu): ONE allocation [flags | vt | hvr] per rank, mapped into every
  // peer through CUDA IPC; the peers store their tiles and their progress flags straight into it
  bool p2p = false;
  unsigned char *comm_block = nullptr;
  unsigned char *peer_block[EDGPU_MAXRANKS] = {};
  // halo mode (default): no transposes at all.  block = [flags | halo 0 | halo 1]; before pass A
  // every owner stores the columns this rank's dw hops read into halo[epoch & 1] (double-buffered:
  // a peer may run one product ahead), sendlist = (local column, destination rank, slot) triples
  bool halo_mode = false;
  double *halo[2] = {nullptr, nullptr};
  int32_t *d_sendlist = nullptr;
  int64_t nsend = 0;
  uint64_t epoch = 0;              // one per distributed product; flags carry the epoch they belong to
  int nchunks = 1;                 // column chunks of vt the Hdw term is pipelined over
  int32_t *pipe_err = nullptr;     // device word set by a wait kernel that timed out
  // all-gathered vector for the non-local terms with nranks>1
  double *vfull = nullptr;
  std::vector<int64_t> gcounts, goffs;
  // a10: user two-body terms (direct/HxV_sundry.f90) and phonons (HxV_ph.f90, HxV_eph.f90)
  int nsundry = 0;
  SundryDev *sundry = nullptr;
  int DimPh = 1;
  double w0_ph = 0.0, A_ph = 0.0;
  double *gph = nullptr;     // [Norb*Norb] device
  bool eph_offdiag = false;  // some g_ph(a,b) != 0 with a != b
  int64_t slice_len() const { return up.ld * qdw; }  // one phonon slice (electronic chunk)
  int64_t padded_len() const { return up.ld * qdw * DimPh; }
  int64_t padded_len_t() const { return dw.ld * qup; }
};

// Stored-H sector (ED_SPARSE_H=T): the host's sparse_matrix_csr (ED_SPARSE_MATRIX.f90:16-41)
// as device CSR.  Rows = this rank's rows [row0, row0+nloc) of the flat row split
// (ED_HAMILTONIAN_NONSU2.f90:72-79), columns global.  Vectors are plain contiguous arrays of
// nloc reals / complex numbers, padded with zeros to a multiple of 16 doubles.
struct CsrSector {
  bool open = false;
  bool cplx = false;
  int64_t nloc = 0, nglobal = 0, row0 = 0, nnz = 0;
  int64_t *rowptr = nullptr;  // [nloc+1]
  int32_t *cols = nullptr;    // [nnz] 0-based global columns
  double *vals = nullptr;     // [nnz] or [2*nnz] (re,im)
  int lanes = 8;              // threads per row of the SpMV kernel
  double *vfull = nullptr;    // nranks>1: all-gathered input vector
  int32_t *map = nullptr;     // device-built sectors (packed.cu): packed Fock states of the WHOLE sector
  int pk_mode = -1;           // -1: host-supplied CSR; 0: nonsu2 (qn = Ntot); 1: superc (qn = Sz)
  int pk_qn = 0, pk_Ns = 0;
  // ED_SPARSE_H = F for the packed-state modes: nothing is stored, every product re-enumerates the
  // rows' matrix elements (packed.cu: k_n2_direct); the ranking tables stay alive with the sector
  bool direct = false;
  bool direct_orbs = false;  // ... of an orbital-resolved NORMAL sector (orbs.cu: k_ob_direct)
  void *pk_off = nullptr, *pk_hb = nullptr;
  std::vector<int64_t> counts, offs;  // row split of all ranks
  int64_t padded_len() const {
    const int64_t n = cplx ? 2 * nloc : nloc;
    return (n + 15) / 16 * 16;
  }
};

struct Engine {
  bool inited = false;
  int device = -1;
  int sm_count = 148;
  size_t smem_optin = 0;
  size_t smem_per_sm = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t comm_stream = nullptr;  // transposes overlapped with the rank-local pass
  cudaStream_t dw_stream = nullptr;    // Hdw on the received chunks + their return (pipeline)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
  cudaEvent_t ev[8] = {};
  // communicator
  int rank = 0, nranks = 1;
  void *nccl = nullptr;  // ncclComm_t
  bool dw_halo = false;  // sector being opened / open: dw species sharded for the halo mode
  bool sparse_h = true;  // ED_SPARSE_H (ED_INPUT_VARS.f90:664) for the packed-state modes
  // scalar scratch
  double *d_scal = nullptr;   // device scalars (dots etc.)
  double *h_scal = nullptr;   // pinned host mirror
  double *d_part = nullptr;   // block partials
  int64_t part_cap = 0;
  Sector sec;
  CsrSector csr;
  // length (in doubles) of a device vector of whatever sector is open
  int64_t veclen() const { return csr.open ? csr.padded_len() : sec.padded_len(); }
  int variant_request = 0;
  // Pool of device chunks holding the Lanczos vectors of the ground-state driver (lanczos.cu).
  // Allocating HBM costs ~4 ms/GB, more than re-running the recurrence saves on a single solve,
  // so the chunks are kept across solves and sectors (a DMFT run solves dozens of sectors per
  // iteration) and are given back as soon as any other allocation of the library runs short
  // (dev_malloc) or on edgpu_release_cache / edgpu_finalize.
  std::vector<std::pair<double *, size_t>> lz_chunks;  // (pointer, bytes)
  bool lz_in_use = false;  // a ground-state solve holds slots of the pool: do not release it
  // Work vectors of the Lanczos drivers (recurrence pair, uploaded start vector), kept across
  // solves like the pool: cudaFree of a GB-sized buffer was measured to take up to 0.5 s next to a
  // 118 GB pool (tools/diag_lanczos.py), more than the whole 0.29 s solve
  double *lz_work[3] = {nullptr, nullptr, nullptr};
  int64_t lz_work_len[3] = {0, 0, 0};
  float stage_ms[4] = {0, 0, 0, 0};
  // module globals coulomb_sundry / Nph, w0_ph, A_ph, g_ph: copied into the sector at open
  std::vector<edgpu_sundry_term> sundry_terms;
  int Nph = 0;
  double w0_ph = 0.0, A_ph = 0.0;
  double g_ph[EDGPU_MAXORB][EDGPU_MAXORB] = {};
  // Hbath_tmp of replica / general baths for the packed-state modes (edgpu_set_hbath_packed):
  // [2][2][hb_Norb][hb_Norb][hb_Nbath] (re,im)
  std::vector<double> hbath_packed;
  int hb_Norb = 0, hb_Nbath = 0;
  // profiling ring (edgpu_profile_begin/end): 4 events per recorded H x v
  std::vector<cudaEvent_t> prof_ev;
  int prof_cap = 0, prof_n = 0;
  bool prof_on = false;
};

extern Engine g;

// Every device allocation of the library goes through dev_malloc: when device memory is short
// the cached Lanczos vector pool is given back first and the allocation is retried.
cudaError_t dev_malloc(void **p, size_t bytes);
template <class T>
inline cudaError_t dev_malloc_t(T **p, size_t bytes) {
  return dev_malloc((void **)p, bytes);
}
#define cudaMalloc(p, n) ::edgpu::dev_malloc_t((p), (n))

// sector.cu
int sector_open(Engine &E, const edgpu_normal_params *p, int nup, int ndw);
// enumeration order + ranking tables of a species with `nel` electrons of the model `p`, exactly
// as sector_open builds them (apply_op needs those of the SOURCE sector's species)
int species_ranking(Engine &E, const edgpu_normal_params &p, int s, int nel, int32_t **map,
                    LinTable *lin, SiteOrder *ord);
RankView rank_view(const LinTable &lin, const SiteOrder &ord);
int sector_close(Engine &E);
int64_t host_binomial(int n, int k);
void block_split(int64_t n, int P, int r, int64_t *q, int64_t *start);

// hxv.cu
int hxv_upload_amps(Engine &E);  // amp2 tables of the open sector's species -> __constant__ c_amp
int hxv_device(Engine &E, const double *d_v, double *d_hv, bool accum, bool timed);
int hxv_device_ex(Engine &E, const double *d_v, double *d_hv, bool accum, bool timed, double s_acc,
                  double s_old, double *dot_out, const double *d_old = nullptr);

// extra.cu (a10): sundry two-body terms and phonons, applied after the electronic passes.
// vfull = all dw columns of every phonon slice ([DimPh][DimDw][ld]; the local vector itself on
// one rank); hv += s_acc * (H_sundry + H_ph + H_eph) v
int extra_setup(Engine &E);   // sector_open: copies the engine-global settings to the device
int extra_hxv(Engine &E, const double *d_v, const double *d_vfull, double *d_hv, double s_acc);

// csr.cu
int csr_open(Engine &E, bool cplx, int64_t nloc, int64_t nglobal, int64_t row0, const int64_t *rowptr,
             const int32_t *cols, const double *vals);
int csr_close(Engine &E);
// takes ownership of device-built CSR arrays (0-based columns) and of the sector map
int csr_adopt_device(Engine &E, bool cplx, int64_t nloc, int64_t nglobal, int64_t row0, int64_t *d_rowptr,
                     int32_t *d_cols, double *d_vals, int64_t nnz, int32_t *d_map,
                     const std::vector<int64_t> *counts = nullptr, const std::vector<int64_t> *offs = nullptr);

// orbs.cu: ed_total_ud=F sectors (Nups(1:Norb), Ndws(1:Norb)) as a device-built real stored H
int orbs_open(Engine &E, const edgpu_normal_params *p, const int32_t *nups, const int32_t *ndws);

// packed.cu
int nonsu2_open(Engine &E, const edgpu_nonsu2_params *p, int ntot);
int superc_open(Engine &E, const edgpu_superc_params *p, int sz);
// apply_COps (ED_SECTOR.f90) on packed-state sectors: out (this rank's rows of the OPEN sector) =
// sum_k coef_k O_k |src>, O_k = c (create=0) / c^+ (create=1) on bit_k; vsrc = FULL source vector
// (complex) of the sector (src_mode, src_qn)
struct PackedOps {
  int n;
  int bit[4], create[4];
  double cre[4], cim[4];
};
int packed_apply_ops(Engine &E, const PackedOps &ops, int src_mode, int src_qn, const double *d_vsrc_full,
                     double *d_out);
int64_t packed_sector_dim(int mode, int Ns, int qn);
// twin state (twin_sector_order, ED_SECTOR.f90:1747-1817): out (this rank's rows of the OPEN sector,
// which is the twin of (src_mode, src_qn)) = vsrc_full[rank_src(flip(state))], flip = all bits
// complemented (nonsu2) / up and dw halves exchanged (superc)
int packed_twin(Engine &E, int src_mode, int src_qn, const double *d_vsrc_full, double *d_out);
// dens(a), docc(a) partial sums over this rank's rows of a packed-state vector
int packed_observables(Engine &E, const double *d_vec, double *h_dens, double *h_docc);
int csr_hxv_device(Engine &E, const double *d_v, double *d_hv, bool accum, double s_acc, double s_old);
// direct (on-the-fly) product of the open packed-state sector: hv = s_acc * H vin [+ s_old * hv]
int packed_direct_hxv(Engine &E, const double *d_vin_full, double *d_hv, bool accum, double s_acc, double s_old);
int orbs_direct_hxv(Engine &E, const double *d_vin_full, double *d_hv, bool accum, double s_acc, double s_old);

// comm.cu
int comm_unique_id(void *uid);
int comm_allgatherv(Engine &E, const double *d_chunk, double *d_full, const std::vector<int64_t> &counts,
                    const std::vector<int64_t> &offs);
int comm_init(Engine &E, int rank, int nranks, const void *uid);
int comm_allreduce_sum(Engine &E, double *d_buf, int n);
// all-reduce of a few HOST doubles (op: 0 sum, 2 max, 3 min = ncclRedOp_t), synchronous; used for
// collective decisions (all ranks must take the same branch), not on the data path
int comm_allreduce_host(Engine &E, double *h_buf, int n, int op);
int comm_transpose(Engine &E, const double *d_a, int64_t nrow, int64_t lda, int64_t qcol,
                   double *d_b, int64_t ncol, int64_t ldb, int64_t qrow, bool accumulate);
int comm_finalize(Engine &E);
// peer-memory path of the two transposes of one H x v (falls back to comm_transpose when the
// peers' buffers cannot be mapped)
int comm_p2p_setup(Engine &E);     // after S.vt / S.hvt exist: exchange + map IPC handles
int comm_p2p_teardown(Engine &E);  // before they are freed
int comm_barrier(Engine &E);       // in-stream barrier over all ranks (1-element all-reduce)
// Chunked peer-memory pipeline of the Hdw term (hxv.cu drives it; DESIGN.md "Multi-GPU").  The
// local columns of vt (= this rank's up rows) are cut into S.nchunks chunks; for chunk c
//   push   : every rank stores its tile of v^T into the owners' vt           (NVLink stores)
//   signal : st.release.sys of the epoch into the owners' flag words
//   wait   : spin (bounded) on this rank's flag words until every sender signalled
//   return : transposed tiles of hvt chunk c are stored into the owners' hvr (NVLink stores)
//   add    : hv += hvr on the rows of chunk c of every owner
enum { PIPE_PUSH = 0, PIPE_RET = 1 };
constexpr int EDGPU_MAXCHUNKS = 16;
constexpr size_t PIPE_FLAG_BYTES = 4096;  // 2 kinds x EDGPU_MAXCHUNKS x EDGPU_MAXRANKS x 8 B
// layout of a rank's IPC block: [flags | vt (ldD x qup doubles) | hvr], 256-byte aligned parts
inline size_t pipe_hvr_offset(int64_t ldD, int64_t qup) {
  return PIPE_FLAG_BYTES + ((size_t)(ldD * qup) * sizeof(double) + 255) / 256 * 256;
}
int comm_pipe_push(Engine &E, int c, const double *d_v, cudaStream_t st);
int comm_pipe_signal(Engine &E, int kind, int c, cudaStream_t st);
int comm_pipe_wait(Engine &E, int kind, int c, cudaStream_t st);
int comm_pipe_return(Engine &E, int c, cudaStream_t st);
int comm_pipe_add(Engine &E, int c, double *d_hv, cudaStream_t st);
void comm_pipe_chunk_cols(Engine &E, int rank, int c, int64_t *col0, int64_t *ncols);  // of vt, local
int comm_pipe_check(Engine &E);  // after a host sync: did a wait kernel time out?
// halo mode: need[p][d] = 1 when rank p's dw hops read column d of another rank (host, all ranks'
// maps all-gathered by the caller): allocates + maps the halo block, builds the send list
int comm_allgather_bytes(Engine &E, const unsigned char *h_mine, unsigned char *h_all, size_t nbytes);
int comm_halo_setup(Engine &E, const std::vector<unsigned char> &need_all);
void halo_plan(int64_t dim, int64_t ld, int P, int me, const unsigned char *need_all, std::vector<int64_t> &nh,
               std::vector<std::vector<int32_t>> &send);
// the halo travels in S.nchunks ROW chunks (pass A consumes 16-row tiles)
void comm_halo_rows(Engine &E, int c, int64_t *row0, int64_t *row1);
int comm_halo_push(Engine &E, int c, const double *d_v, cudaStream_t st);
int comm_halo_signal(Engine &E, int c, cudaStream_t st);
int comm_halo_wait(Engine &E, int c, cudaStream_t st);

// vecops.cu
int vec_fill_random(Engine &E, double *d_v, uint64_t seed);
int ensure_partials(Engine &E, int64_t n);                 // E.d_part holds >= n doubles
int final_sum(Engine &E, int nblocks, double *d_out);      // *d_out = sum(E.d_part[0..n)), on device
int vec_dot_dev(Engine &E, const double *a, const double *b, double *d_out);  // local, device scalar
int scalar_to_host(Engine &E, double *d_scalar, double *h_out);  // all-reduce + copy + sync
int vec_zero(Engine &E, double *d_v, int64_t n);
int vec_dot(Engine &E, const double *a, const double *b, double *h_out);  // all-reduced
int vec_scale(Engine &E, double *a, double s);
// w -= alpha v ; beta2 = <w,w> ; store (optional) receives a copy of the new w
int vec_axpy_norm(Engine &E, double *w, const double *v, double alpha, double *h_beta2,
                  double *store = nullptr);
int vec_lincomb(Engine &E, double *out, const std::vector<double *> &vecs, const std::vector<double> &coef);
int vec_axpy(Engine &E, double *y, const double *x, double a);

// lanczos.cu
void lanczos_release(Engine &E);  // frees the Lanczos vector pool and the cached work vectors
int lanczos_work(Engine &E, int k, int64_t n, double **p);  // cached work vector k (>= n doubles)
int tridiag_eig(int n, const double *diag, const double *sub, double *evals, double *evecs,
                bool want_vecs, int track_row = 0);
int lanczos_gs_dev(Engine &E, int nitermax, double threshold, int ncheck, const double *d_start,
                   uint64_t seed, double *egs, double *d_vect, int *niter, double resid_tol = 0.0);
extern int g_lanczos_last_stored, g_lanczos_last_hxv;
extern double g_lanczos_last_resid;

}  // namespace edgpu
