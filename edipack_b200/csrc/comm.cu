// Rank plumbing for the dw-split layout (ED_HAMILTONIAN_NORMAL.f90:128-142): NCCL (loaded at run
// time) for the collectives, and three ways to get the Hdw term across ranks, i.e. replacements of
// vector_transpose_MPI (ED_HAMILTONIAN_NORMAL_COMMON.f90:66-178; one MPI_Alltoallv per local
// column, :115-162):
//   * halo mode (default, second half of this file): no transpose; the owners store the few remote
//     dw columns a chunk's hops read straight into the reader's memory over NVLink (TMA copy kernel,
//     st.release.sys / ld.acquire.sys flags), pass A then runs on the chunk as on one GPU;
//   * chunk-pipelined peer-memory transposes with the same flag protocol (EDGPU_DW_MODE=transpose);
//   * NCCL fallback (EDGPU_NO_P2P=1 / unmappable peers): each rank packs ONE transposed tile per peer
//     with a shared-memory tile-transpose kernel, exchanges all tiles in a single
//     ncclGroupStart/End of ncclSend/ncclRecv, and unpacks (or accumulates) with a strided copy
//     kernel; the rank's own tile never leaves the device.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "edgpu_internal.cuh"

namespace edgpu {

// ---- minimal NCCL surface, resolved with dlopen so that the library has no link-time
// ---- dependency (torch ships its own libnccl.so.2; the system one is used otherwise)
typedef struct {
  char internal[128];
} nccl_uid_t;
typedef void *nccl_comm_t;
enum { NCCL_SUM = 0, NCCL_CHAR = 0, NCCL_FLOAT64 = 8 };
static struct {
  void *h = nullptr;
  int (*GetUniqueId)(nccl_uid_t *) = nullptr;
  int (*CommInitRank)(nccl_comm_t *, int, nccl_uid_t, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
} N;

static int nccl_load() {
  if (N.h) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    N.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (N.h) break;
  }
  if (!N.h) return set_error("cannot load libnccl.so.2: %s", dlerror());
#define EDGPU_SYM(field, name)                                            \
  *(void **)(&N.field) = dlsym(N.h, name);                                \
  if (!N.field) return set_error("NCCL symbol %s not found", name);
  EDGPU_SYM(GetUniqueId, "ncclGetUniqueId");
  EDGPU_SYM(CommInitRank, "ncclCommInitRank");
  EDGPU_SYM(CommDestroy, "ncclCommDestroy");
  EDGPU_SYM(Send, "ncclSend");
  EDGPU_SYM(Recv, "ncclRecv");
  EDGPU_SYM(GroupStart, "ncclGroupStart");
  EDGPU_SYM(GroupEnd, "ncclGroupEnd");
  EDGPU_SYM(AllReduce, "ncclAllReduce");
  EDGPU_SYM(AllGather, "ncclAllGather");
  EDGPU_SYM(Broadcast, "ncclBroadcast");
  EDGPU_SYM(GetErrorString, "ncclGetErrorString");
#undef EDGPU_SYM
  return 0;
}

#define EDGPU_NCCL(call)                                                                   \
  do {                                                                                     \
    int _r = (call);                                                                       \
    if (_r != 0) return set_error("%s failed: %s", #call, N.GetErrorString(_r));           \
  } while (0)

int comm_unique_id(void *uid) {
  EDGPU_TRY(nccl_load());
  static_assert(sizeof(nccl_uid_t) == EDGPU_UID_BYTES, "uid size");
  nccl_uid_t id;
  EDGPU_NCCL(N.GetUniqueId(&id));
  memcpy(uid, &id, sizeof(id));
  return 0;
}

int comm_init(Engine &E, int rank, int nranks, const void *uid) {
  if (!E.inited) return set_error("edgpu_init was not called");
  if (nranks < 1 || rank < 0 || rank >= nranks) return set_error("bad rank %d / %d", rank, nranks);
  if (E.sec.open) return set_error("close the sector before (re)initialising the communicator");
  comm_finalize(E);
  E.rank = rank;
  E.nranks = nranks;
  if (nranks == 1) return 0;
  EDGPU_TRY(nccl_load());
  nccl_uid_t id;
  memcpy(&id, uid, sizeof(id));
  nccl_comm_t c = nullptr;
  EDGPU_NCCL(N.CommInitRank(&c, nranks, id, rank));
  E.nccl = c;
  return 0;
}

int comm_finalize(Engine &E) {
  if (E.nccl) {
    N.CommDestroy((nccl_comm_t)E.nccl);
    E.nccl = nullptr;
  }
  E.rank = 0;
  E.nranks = 1;
  return 0;
}

int comm_allreduce_sum(Engine &E, double *d_buf, int n) {
  if (E.nranks == 1) return 0;
  EDGPU_NCCL(N.AllReduce(d_buf, d_buf, (size_t)n, NCCL_FLOAT64, NCCL_SUM, (nccl_comm_t)E.nccl,
                         E.stream));
  return 0;
}

int comm_allreduce_host(Engine &E, double *h_buf, int n, int op) {
  if (E.nranks == 1) return 0;
  if (n < 1 || n > 16) return set_error("comm_allreduce_host: n=%d out of range", n);
  double *d = E.d_scal + 40;  // scratch slots 40..55 of the scalar buffer
  EDGPU_CUDA(cudaMemcpyAsync(d, h_buf, sizeof(double) * n, cudaMemcpyHostToDevice, E.stream));
  EDGPU_NCCL(N.AllReduce(d, d, (size_t)n, NCCL_FLOAT64, op, (nccl_comm_t)E.nccl, E.stream));
  EDGPU_CUDA(cudaMemcpyAsync(h_buf, d, sizeof(double) * n, cudaMemcpyDeviceToHost, E.stream));
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  return 0;
}

// MPI_Allgatherv of the input vector (ED_HAMILTONIAN_NONSU2_STORED_HxV.f90:256-259): every rank's
// chunk lands at its offset of the full vector; unequal counts -> one grouped broadcast per rank.
int comm_allgatherv(Engine &E, const double *d_chunk, double *d_full, const std::vector<int64_t> &counts,
                    const std::vector<int64_t> &offs) {
  const int P = E.nranks, me = E.rank;
  if (P == 1) {
    EDGPU_CUDA(cudaMemcpyAsync(d_full + offs[0], d_chunk, sizeof(double) * counts[0],
                               cudaMemcpyDeviceToDevice, E.stream));
    return 0;
  }
  EDGPU_NCCL(N.GroupStart());
  for (int p = 0; p < P; p++)
    EDGPU_NCCL(N.Broadcast(p == me ? (const void *)d_chunk : (const void *)(d_full + offs[p]),
                           d_full + offs[p], (size_t)counts[p], NCCL_FLOAT64, p, (nccl_comm_t)E.nccl,
                           E.stream));
  EDGPU_NCCL(N.GroupEnd());
  return 0;
}

// out[(i - i0) * ldo + j] = in[j * ldi + i]   for i in [i0, i0+ni), j in [0, nj)
// (tile transpose of a ni x nj block through shared memory; in has i fast, out has j fast)
template <bool ACCUM>
__global__ void __launch_bounds__(256)
k_transpose_tile(const double *__restrict__ in, int64_t ldi, double *__restrict__ out,
                 int64_t ldo, int64_t i0, int64_t ni, int64_t nj) {
  __shared__ double t[32][33];
  const int64_t ib = (int64_t)blockIdx.x * 32, jb = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    int64_t i = ib + tx, j = jb + ty + k;
    if (i < ni && j < nj) t[ty + k][tx] = in[j * ldi + i0 + i];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    int64_t j = jb + tx, i = ib + ty + k;
    if (i < ni && j < nj) {
      double *o = out + i * ldo + j;
      *o = ACCUM ? (*o + t[tx][ty + k]) : t[tx][ty + k];
    }
  }
}

// out[i * ldo + j] (+)= in[i * nj + j] : unpack a received (already transposed) tile
template <bool ACCUM>
__global__ void __launch_bounds__(256)
k_unpack(const double *__restrict__ in, double *__restrict__ out, int64_t ldo, int64_t ni,
         int64_t nj) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j < nj && i < ni) {
    double *o = out + i * ldo + j;
    *o = ACCUM ? (*o + in[i * nj + j]) : in[i * nj + j];
  }
}

// B(j_global, i_loc) = A_global(i0_me + i_loc, j): A is this rank's [nrow x qcol] column
// block (lda), B its [ncol x qrow] block of the transposed matrix (ldb).
// vector_transpose_MPI(nrow,qcol,a,ncol,qrow,b), ..._COMMON.f90:66.
int comm_transpose(Engine &E, const double *d_a, int64_t nrow, int64_t lda, int64_t qcol,
                   double *d_b, int64_t ncol, int64_t ldb, int64_t qrow, bool accumulate) {
  const int P = E.nranks, me = E.rank;
  Sector &S = E.sec;
  cudaStream_t st = E.stream;
  int64_t c0_me, qc_me, r0_me, qr_me;
  block_split(ncol, P, me, &qc_me, &c0_me);
  block_split(nrow, P, me, &qr_me, &r0_me);
  if (qc_me != qcol || qr_me != qrow) return set_error("comm_transpose: split mismatch");
  const size_t need = (size_t)nrow * qcol > (size_t)ncol * qrow ? (size_t)nrow * qcol
                                                                 : (size_t)ncol * qrow;
  if (P > 1 && !S.sendbuf) {
    // both directions of one H x v use the same element count
    EDGPU_CUDA(cudaMalloc(&S.sendbuf, sizeof(double) * need));
    EDGPU_CUDA(cudaMalloc(&S.recvbuf, sizeof(double) * need));
  }
  // pack: for peer p the tile rows [r0_p, r0_p+qr_p) x my columns, stored [i_loc][jc]
  std::vector<int64_t> soff(P + 1, 0), roff(P + 1, 0);
  for (int p = 0; p < P; p++) {
    int64_t qr, r0, qc, c0;
    block_split(nrow, P, p, &qr, &r0);
    block_split(ncol, P, p, &qc, &c0);
    soff[p + 1] = soff[p] + (p == me ? 0 : qr * qcol);
    roff[p + 1] = roff[p] + (p == me ? 0 : qrow * qc);
  }
  for (int p = 0; p < P; p++) {
    int64_t qr, r0;
    block_split(nrow, P, p, &qr, &r0);
    dim3 grid((unsigned)((qr + 31) / 32), (unsigned)((qcol + 31) / 32));
    if (p == me) {
      // own tile: straight into B at column offset c0_me
      if (accumulate)
        k_transpose_tile<true><<<grid, 256, 0, st>>>(d_a, lda, d_b + c0_me, ldb, r0, qr, qcol);
      else
        k_transpose_tile<false><<<grid, 256, 0, st>>>(d_a, lda, d_b + c0_me, ldb, r0, qr, qcol);
    } else {
      k_transpose_tile<false><<<grid, 256, 0, st>>>(d_a, lda, S.sendbuf + soff[p], qcol, r0, qr,
                                                    qcol);
    }
    EDGPU_COUNT_LAUNCH();
  }
  EDGPU_CUDA(cudaGetLastError());
  if (P == 1) return 0;
  EDGPU_NCCL(N.GroupStart());
  for (int p = 0; p < P; p++) {
    if (p == me) continue;
    size_t ns = (size_t)(soff[p + 1] - soff[p]), nr = (size_t)(roff[p + 1] - roff[p]);
    EDGPU_NCCL(N.Send(S.sendbuf + soff[p], ns, NCCL_FLOAT64, p, (nccl_comm_t)E.nccl, st));
    EDGPU_NCCL(N.Recv(S.recvbuf + roff[p], nr, NCCL_FLOAT64, p, (nccl_comm_t)E.nccl, st));
  }
  EDGPU_NCCL(N.GroupEnd());
  for (int p = 0; p < P; p++) {
    if (p == me) continue;
    int64_t qc, c0;
    block_split(ncol, P, p, &qc, &c0);
    dim3 grid((unsigned)((qc + 255) / 256), (unsigned)qrow);
    if (accumulate)
      k_unpack<true><<<grid, 256, 0, st>>>(S.recvbuf + roff[p], d_b + c0, ldb, qrow, qc);
    else
      k_unpack<false><<<grid, 256, 0, st>>>(S.recvbuf + roff[p], d_b + c0, ldb, qrow, qc);
    EDGPU_COUNT_LAUNCH();
  }
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}


// ---------------------------------------------------------------------------------------
// Peer-memory pipeline of the Hdw term (replaces both vector_transpose_MPI of one H x v,
// ED_HAMILTONIAN_NORMAL_COMMON.f90:66-178, and the MPI_Alltoallv per column inside them).
//
// Every rank owns ONE IPC-mapped block [flags | vt | hvr]:
//   vt   [ldD x qup]  this rank's rows of v^T: written by all ranks (push)
//   hvr  [ldU x qdw]  Hdw part of this rank's columns of Hv: written by all ranks (return)
//   flags             epoch words, written by all ranks (st.release.sys), spun on locally
// The qup local columns of vt (= up rows u0 .. u0+qup) are cut into nchunks chunks and chunk c
// flows  push(c) -> [flag] -> Hdw on chunk c -> return(c) -> [flag] -> hv += hvr on chunk c's rows
// so that NVLink traffic in both directions, the Hdw pass and the rank-local pass overlap.  The
// tiles are transposed through shared memory: both the local and the remote side move 256-byte
// segments, issued by the kernels themselves (st.global on mapped peer pointers) -- no pack /
// send / recv / unpack passes, no collective call on the data path.
//
// Buffer reuse across products needs no extra synchronisation: a rank starts product e+1 only
// after it has seen the return flags of product e from every peer, i.e. after every peer has
// finished reading its vt of product e.
// ---------------------------------------------------------------------------------------
static_assert(2 * EDGPU_MAXCHUNKS * EDGPU_MAXRANKS * sizeof(uint64_t) <= PIPE_FLAG_BYTES, "flag area");

struct PipeTable {
  unsigned char *block[EDGPU_MAXRANKS];
  int32_t u0[EDGPU_MAXRANKS + 1];   // up rows [u0[p], u0[p+1]) belong to rank p (columns of its vt)
  int32_t d0[EDGPU_MAXRANKS + 1];   // dw columns [d0[p], d0[p+1]) belong to rank p
  int32_t cb[EDGPU_MAXRANKS][EDGPU_MAXCHUNKS + 1];  // chunk c of rank p = its local rows [cb[c], cb[c+1])
  int64_t off_vt, off_hvr[EDGPU_MAXRANKS];  // byte offsets of vt / hvr inside rank p's block
  int64_t ldU, ldD;
  int nranks, me, nchunks;
};

__device__ __forceinline__ int flag_slot(int kind, int c, int sender) {
  return (kind * EDGPU_MAXCHUNKS + c) * EDGPU_MAXRANKS + sender;
}

static PipeTable g_pipe;  // of the open sector (host copy, passed to the kernels by value)

// chunk boundaries: multiples of 32 rows (tile height), the last chunk takes the remainder
static void chunk_bounds(int64_t q, int K, int32_t *cb) {
  const int64_t per = ((q + K - 1) / K + 31) / 32 * 32;
  for (int c = 0; c <= K; c++) cb[c] = (int32_t)std::min<int64_t>(q, per * c);
  cb[K] = (int32_t)q;
}

void comm_pipe_chunk_cols(Engine &E, int rank, int c, int64_t *col0, int64_t *ncols) {
  *col0 = g_pipe.cb[rank][c];
  *ncols = g_pipe.cb[rank][c + 1] - g_pipe.cb[rank][c];
}

// ---- push: v[c * ldU + i] -> vt_p[(i - u0_p) * ldD + d0_me + c] for the rows i of chunk `ch` of
// every owner p.  grid = (ceil(qdw/32), nranks * row tiles of the longest chunk); blockIdx.y
// rotates over the owners, starting at a different one on every rank (all-to-all schedule), and
// column tiles run fastest: consecutive CTAs extend the SAME 32 remote rows by consecutive
// 256-byte pieces (row tiles fastest would put every remote segment on another 2 MB page of a
// multi-GB peer buffer: measured 200 ms instead of ~5 ms per transpose at Ns=18 on 8 GPUs).
__global__ void __launch_bounds__(256)
k_pipe_push(const double *__restrict__ a, int qcol, int ch, PipeTable T) {
  __shared__ double t[32][33];
  const int p = ((int)(blockIdx.y % T.nranks) + T.me) % T.nranks;
  const int ib = T.u0[p] + T.cb[p][ch] + (int)(blockIdx.y / T.nranks) * 32;
  const int iend = T.u0[p] + T.cb[p][ch + 1];
  if (ib >= iend) return;
  const int jb = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int i = ib + tx, j = jb + ty + k;
    if (i < iend && j < qcol) t[ty + k][tx] = a[(int64_t)j * T.ldU + i];
  }
  __syncthreads();
  double *dst = reinterpret_cast<double *>(T.block[p] + T.off_vt) + T.d0[T.me];
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int j = jb + tx, i = ib + ty + k;  // the 32 lanes write 32 consecutive dw columns of row i
    if (i < iend && j < qcol) dst[(int64_t)(i - T.u0[p]) * T.ldD + j] = t[tx][ty + k];
  }
}

// ---- return: hvt[j * ldD + d] (j = local column of chunk `ch`, up row u0_me + j) ->
// hvr_q[(d - d0_q) * ldU + u0_me + j] for every owner q of dw column d.
// grid = (ceil(chunk cols / 32), nranks * dw tiles of the widest rank).
__global__ void __launch_bounds__(256)
k_pipe_return(const double *__restrict__ hvt, int ch, PipeTable T) {
  __shared__ double t[32][33];
  const int q = ((int)(blockIdx.y % T.nranks) + T.me) % T.nranks;
  const int db = T.d0[q] + (int)(blockIdx.y / T.nranks) * 32;
  const int dend = T.d0[q + 1];
  if (db >= dend) return;
  const int j0 = T.cb[T.me][ch], jend = T.cb[T.me][ch + 1];
  const int jb = j0 + blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int d = db + tx, j = jb + ty + k;
    if (d < dend && j < jend) t[ty + k][tx] = hvt[(int64_t)j * T.ldD + d];
  }
  __syncthreads();
  double *dst = reinterpret_cast<double *>(T.block[q] + T.off_hvr[q]) + T.u0[T.me];
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int j = jb + tx, d = db + ty + k;  // 32 lanes: 32 consecutive up rows of dw column d
    if (d < dend && j < jend) dst[(int64_t)(d - T.d0[q]) * T.ldU + j] = t[tx][ty + k];
  }
}

// ---- signal: one thread per peer.  The kernel is stream-ordered behind the kernel whose stores
// it announces; the fence orders those (already performed) stores before the flag at system scope.
__global__ void k_pipe_signal(int kind, int c, unsigned long long epoch, PipeTable T) {
  const int p = threadIdx.x;
  if (p >= T.nranks) return;
  __threadfence_system();
  unsigned long long *f = reinterpret_cast<unsigned long long *>(T.block[p]) + flag_slot(kind, c, T.me);
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
}

// ---- wait: one thread per sender spins on this rank's own flag word.  Bounded: after
// `timeout_ns` the error word is set and the kernel returns (a dead peer must not hang the box).
__global__ void k_pipe_wait(int kind, int c, unsigned long long epoch, PipeTable T,
                            unsigned long long timeout_ns, int32_t *err) {
  const int p = threadIdx.x;
  if (p >= T.nranks) return;
  const unsigned long long *f =
      reinterpret_cast<const unsigned long long *>(T.block[T.me]) + flag_slot(kind, c, p);
  unsigned long long t0, t1, seen;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f) : "memory");
    if (seen >= epoch) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) {
      *(volatile int32_t *)err = 1 + kind;  // pinned host word, read after the next stream sync
      __threadfence_system();
      break;
    }
    __nanosleep(200);
  }
}

// ---- add: hv[col * ldU + i] += hvr[col * ldU + i] on the rows i of chunk `ch` of every owner.
// grid = (qdw, row blocks of the longest chunk, nranks)
__global__ void __launch_bounds__(256)
k_pipe_add(double *__restrict__ hv, const double *__restrict__ hvr, int ch, PipeTable T) {
  const int p = blockIdx.z;
  const int i = T.u0[p] + T.cb[p][ch] + (int)(blockIdx.y * blockDim.x + threadIdx.x);
  if (i >= T.u0[p] + T.cb[p][ch + 1]) return;
  const int64_t o = (int64_t)blockIdx.x * T.ldU + i;
  hv[o] += hvr[o];
}

int comm_barrier(Engine &E) {
  if (E.nranks == 1) return 0;
  EDGPU_NCCL(N.AllReduce(E.d_scal + 32, E.d_scal + 32, 1, NCCL_FLOAT64, NCCL_SUM, (nccl_comm_t)E.nccl,
                         E.stream));
  return 0;
}

static int pipe_default_chunks(Engine &E) {
  if (const char *e = getenv("EDGPU_CHUNKS")) {
    const int k = atoi(e);
    if (k >= 1) return std::min(k, EDGPU_MAXCHUNKS);
  }
  // ~48 MB of vt per chunk, between 2 and 8 chunks
  const double mb = (double)E.sec.padded_len_t() * 8.0 / (1 << 20);
  return (int)std::min(8.0, std::max(2.0, mb / 48.0 + 0.5));
}

int comm_p2p_setup(Engine &E) {
  Sector &S = E.sec;
  S.p2p = false;
  const int P = E.nranks, me = E.rank;
  if (P == 1 || P > EDGPU_MAXRANKS) return 0;
  const char *off = getenv("EDGPU_NO_P2P");
  // every rank must take the same decision: all-reduce the local "ok" flags
  int ok = (off && off[0] == '1') ? 0 : 1;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaIpcGetMemHandle(&mine, S.comm_block) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  unsigned char *d_h = nullptr;
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  EDGPU_CUDA(cudaMalloc(&d_h, hb * P));
  EDGPU_CUDA(cudaMemcpyAsync(d_h + hb * me, &mine, hb, cudaMemcpyHostToDevice, E.stream));
  EDGPU_NCCL(N.AllGather(d_h + hb * me, d_h, hb, NCCL_CHAR, (nccl_comm_t)E.nccl, E.stream));
  std::vector<cudaIpcMemHandle_t> all((size_t)P);
  EDGPU_CUDA(cudaMemcpyAsync(all.data(), d_h, hb * P, cudaMemcpyDeviceToHost, E.stream));
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  cudaFree(d_h);
  for (int p = 0; p < P && ok; p++) {
    if (p == me) {
      S.peer_block[p] = S.comm_block;
      continue;
    }
    void *a = nullptr;
    if (cudaIpcOpenMemHandle(&a, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
      break;
    }
    S.peer_block[p] = (unsigned char *)a;
  }
  double flag = ok ? 0.0 : 1.0;
  EDGPU_TRY(comm_allreduce_host(E, &flag, 1, NCCL_SUM));
  S.p2p = (flag == 0.0);
  if (!S.p2p) {  // someone could not map: everybody unmaps and uses the NCCL path
    for (int p = 0; p < P; p++) {
      if (p != me && S.peer_block[p]) cudaIpcCloseMemHandle(S.peer_block[p]);
      S.peer_block[p] = nullptr;
    }
    cudaGetLastError();
    return 0;
  }
  S.nchunks = pipe_default_chunks(E);
  S.epoch = 0;
  PipeTable &T = g_pipe;
  memset(&T, 0, sizeof(T));
  T.nranks = P;
  T.me = me;
  T.nchunks = S.nchunks;
  T.off_vt = (int64_t)PIPE_FLAG_BYTES;
  T.ldU = S.up.ld;
  T.ldD = S.dw.ld;
  for (int p = 0; p < P; p++) {
    int64_t q, r0;
    block_split(S.up.dim, P, p, &q, &r0);
    T.u0[p] = (int32_t)r0;
    T.u0[p + 1] = (int32_t)(r0 + q);
    chunk_bounds(q, S.nchunks, T.cb[p]);
    T.off_hvr[p] = (int64_t)pipe_hvr_offset(S.dw.ld, q);
    block_split(S.dw.dim, P, p, &q, &r0);
    T.d0[p] = (int32_t)r0;
    T.d0[p + 1] = (int32_t)(r0 + q);
    T.block[p] = S.peer_block[p];
  }
  return 0;
}

int comm_p2p_teardown(Engine &E) {
  Sector &S = E.sec;
  if (!S.p2p) return 0;
  // nobody may still be reading / writing a peer's buffer
  comm_barrier(E);
  cudaStreamSynchronize(E.stream);
  for (int p = 0; p < E.nranks; p++) {
    if (p != E.rank && S.peer_block[p]) cudaIpcCloseMemHandle(S.peer_block[p]);
    S.peer_block[p] = nullptr;
  }
  comm_barrier(E);  // every rank unmapped before anybody frees
  cudaStreamSynchronize(E.stream);
  S.p2p = false;
  return 0;
}

static int longest_chunk(const PipeTable &T, int c) {
  int m = 0;
  for (int p = 0; p < T.nranks; p++) m = std::max(m, T.cb[p][c + 1] - T.cb[p][c]);
  return m;
}

static unsigned long long pipe_timeout_ns();

int comm_pipe_push(Engine &E, int c, const double *d_v, cudaStream_t st) {
  Sector &S = E.sec;
  const int rt = (longest_chunk(g_pipe, c) + 31) / 32;
  if (S.qdw <= 0 || rt == 0) return 0;
  dim3 grid((unsigned)((S.qdw + 31) / 32), (unsigned)(E.nranks * rt));
  k_pipe_push<<<grid, 256, 0, st>>>(d_v, (int)S.qdw, c, g_pipe);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int comm_pipe_return(Engine &E, int c, cudaStream_t st) {
  Sector &S = E.sec;
  const PipeTable &T = g_pipe;
  const int nc = T.cb[T.me][c + 1] - T.cb[T.me][c];
  int wd = 0;
  for (int p = 0; p < T.nranks; p++) wd = std::max(wd, T.d0[p + 1] - T.d0[p]);
  if (nc <= 0 || wd == 0) return 0;
  dim3 grid((unsigned)((nc + 31) / 32), (unsigned)(E.nranks * ((wd + 31) / 32)));
  k_pipe_return<<<grid, 256, 0, st>>>(S.hvt, c, T);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int comm_pipe_signal(Engine &E, int kind, int c, cudaStream_t st) {
  k_pipe_signal<<<1, 32, 0, st>>>(kind, c, (unsigned long long)E.sec.epoch, g_pipe);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int comm_pipe_wait(Engine &E, int kind, int c, cudaStream_t st) {
  k_pipe_wait<<<1, 32, 0, st>>>(kind, c, (unsigned long long)E.sec.epoch, g_pipe, pipe_timeout_ns(),
                                E.sec.pipe_err);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int comm_pipe_add(Engine &E, int c, double *d_hv, cudaStream_t st) {
  Sector &S = E.sec;
  const int len = longest_chunk(g_pipe, c);
  if (S.qdw <= 0 || len == 0) return 0;
  dim3 grid((unsigned)S.qdw, (unsigned)((len + 255) / 256), (unsigned)E.nranks);
  k_pipe_add<<<grid, 256, 0, st>>>(d_hv, S.hvr, c, g_pipe);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------
// Halo mode (default for nranks > 1): the Hdw term without any transpose.  The hops of a rank's
// chunk of dw columns read a few columns of other ranks (those reached by moving an electron
// into / out of the top bath sites: ~ 8 T / Nbath columns per local column for T ~ log2(nranks)
// "rank bits").  Before pass A every owner stores those columns -- contiguous runs of DimUp
// doubles -- straight into the reader's halo buffer over NVLink and raises a flag; pass A (k_slow)
// then runs on the chunk exactly as on one GPU, far gathers landing in the halo.  Per product and
// state this moves <= the bytes of ONE transpose and adds no HBM pass (the double transpose of
// vector_transpose_MPI costs push + Hdw on v^T + return + accumulate = 48 B/state on top).
// block = [flags | halo 0 | halo 1]: double-buffered by the parity of the product counter, because
// a peer may run one product ahead (it cannot run two ahead: its push of product e+2 needs this
// rank's push flag of e+1, which is stream-ordered behind this rank's pass A of product e).
// ---------------------------------------------------------------------------------------
struct HaloTable {
  unsigned char *block[EDGPU_MAXRANKS];
  int64_t hbytes[EDGPU_MAXRANKS];  // bytes of ONE halo buffer of rank p
  int64_t ldU;
  int nranks, me;
};
static HaloTable g_halo;

int comm_allgather_bytes(Engine &E, const unsigned char *h_mine, unsigned char *h_all, size_t nbytes) {
  const int P = E.nranks;
  if (P == 1) {
    memcpy(h_all, h_mine, nbytes);
    return 0;
  }
  unsigned char *d = nullptr;
  EDGPU_CUDA(cudaMalloc(&d, nbytes * P));
  EDGPU_CUDA(cudaMemcpyAsync(d + nbytes * E.rank, h_mine, nbytes, cudaMemcpyHostToDevice, E.stream));
  EDGPU_NCCL(N.AllGather(d + nbytes * E.rank, d, nbytes, NCCL_CHAR, (nccl_comm_t)E.nccl, E.stream));
  EDGPU_CUDA(cudaMemcpyAsync(h_all, d, nbytes * P, cudaMemcpyDeviceToHost, E.stream));
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  cudaFree(d);
  return 0;
}

// maps `mine` (this rank's block) into every peer; *ok_all = every rank mapped every peer
static int ipc_exchange(Engine &E, unsigned char *mine_block, unsigned char **peer, bool *ok_all) {
  const int P = E.nranks, me = E.rank;
  int ok = 1;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (cudaIpcGetMemHandle(&mine, mine_block) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  std::vector<cudaIpcMemHandle_t> all((size_t)P);
  EDGPU_TRY(comm_allgather_bytes(E, (const unsigned char *)&mine, (unsigned char *)all.data(), sizeof(mine)));
  for (int p = 0; p < P; p++) peer[p] = nullptr;
  for (int p = 0; p < P && ok; p++) {
    if (p == me) {
      peer[p] = mine_block;
      continue;
    }
    void *a = nullptr;
    if (cudaIpcOpenMemHandle(&a, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
      break;
    }
    peer[p] = (unsigned char *)a;
  }
  double flag = ok ? 0.0 : 1.0;
  EDGPU_TRY(comm_allreduce_host(E, &flag, 1, NCCL_SUM));
  *ok_all = (flag == 0.0);
  if (!*ok_all) {
    for (int p = 0; p < P; p++) {
      if (p != me && peer[p]) cudaIpcCloseMemHandle(peer[p]);
      peer[p] = nullptr;
    }
    cudaGetLastError();
  }
  return 0;
}

// The exchange plan, pure host logic (no CUDA; exported as edgpu_halo_plan for the CPU tests):
// need[r * ld + d] != 0 when rank r's chunk reads column d of another rank.  Rank r's halo holds its
// needed columns in ascending order; nh[r] = its size; send[r] = (local column of `me`, slot in
// r's halo) pairs of the columns `me` owns.  Chunks are the reference's dw split (block_split).
void halo_plan(int64_t dim, int64_t ld, int P, int me, const unsigned char *need_all, std::vector<int64_t> &nh,
               std::vector<std::vector<int32_t>> &send) {
  int64_t q_me, d0_me;
  block_split(dim, P, me, &q_me, &d0_me);
  nh.assign(P, 0);
  send.assign(P, {});
  for (int r = 0; r < P; r++) {
    int64_t qr, d0r;
    block_split(dim, P, r, &qr, &d0r);
    const unsigned char *need = need_all + (size_t)r * (size_t)ld;
    int64_t slot = 0;
    for (int64_t d = 0; d < dim; d++) {
      if (d >= d0r && d < d0r + qr) continue;
      if (!need[d]) continue;
      if (d >= d0_me && d < d0_me + q_me) {
        send[r].push_back((int32_t)(d - d0_me));
        send[r].push_back((int32_t)slot);
      }
      slot++;
    }
    nh[r] = slot;
  }
}

int comm_halo_setup(Engine &E, const std::vector<unsigned char> &need_all) {
  Sector &S = E.sec;
  const int P = E.nranks, me = E.rank;
  const int64_t ld = S.dw.ld, dim = S.dw.dim, ldU = S.up.ld;
  S.halo_mode = false;
  // halo size of every rank and the slot of each of MY columns in each reader's halo
  std::vector<int64_t> nh;
  std::vector<std::vector<int32_t>> send;  // per destination: (local col, slot) pairs
  halo_plan(dim, ld, P, me, need_all.data(), nh, send);
  if (nh[me] != S.dw.nhalo) return set_error("internal: halo size mismatch");
  HaloTable &T = g_halo;
  memset(&T, 0, sizeof(T));
  T.nranks = P;
  T.me = me;
  T.ldU = ldU;
  for (int p = 0; p < P; p++) T.hbytes[p] = (int64_t)(((size_t)nh[p] * (size_t)ldU * sizeof(double) + 255) / 256 * 256);
  const size_t total = PIPE_FLAG_BYTES + 2 * (size_t)T.hbytes[me];
  EDGPU_CUDA(cudaMalloc(&S.comm_block, total));
  EDGPU_CUDA(cudaMemsetAsync(S.comm_block, 0, total, E.stream));
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  bool ok = false;
  EDGPU_TRY(ipc_exchange(E, S.comm_block, S.peer_block, &ok));
  if (!ok) {
    cudaFree(S.comm_block);
    S.comm_block = nullptr;
    return 0;
  }
  for (int p = 0; p < P; p++) T.block[p] = S.peer_block[p];
  S.halo[0] = reinterpret_cast<double *>(S.comm_block + PIPE_FLAG_BYTES);
  S.halo[1] = reinterpret_cast<double *>(S.comm_block + PIPE_FLAG_BYTES + T.hbytes[me]);
  // send list, destinations interleaved (starting at a different one on every rank) so that all
  // NVLink endpoints are busy at any instant
  std::vector<int32_t> list;
  size_t longest = 0;
  for (int r = 0; r < P; r++) longest = std::max(longest, send[r].size() / 2);
  for (size_t k = 0; k < longest; k++)
    for (int dr = 1; dr < P; dr++) {
      const int r = (me + dr) % P;
      if (k < send[r].size() / 2) {
        list.push_back(send[r][2 * k]);
        list.push_back(r);
        list.push_back(send[r][2 * k + 1]);
      }
    }
  S.nsend = (int64_t)list.size() / 3;
  EDGPU_CUDA(cudaMalloc(&S.d_sendlist, sizeof(int32_t) * std::max<size_t>(list.size(), 3)));
  if (!list.empty())
    EDGPU_CUDA(cudaMemcpy(S.d_sendlist, list.data(), sizeof(int32_t) * list.size(), cudaMemcpyHostToDevice));
  S.pipe_err = reinterpret_cast<int32_t *>(E.h_scal + 60);  // pinned host word (UVA: device-visible)
  *S.pipe_err = 0;
  S.epoch = 0;
  S.p2p = true;
  S.halo_mode = true;
  {
    // row chunks of the halo exchange (pipelined against pass A); at least 16 tiles of 16 rows each
    const char *e = getenv("EDGPU_HALO_CHUNKS");
    // (by the size of the local vector: 8 chunks above 1 GB -- the last chunk's pass A is the tail --,
    // 4 above 256 MB, else 2: every chunk costs four more kernel launches)
    const int64_t lb = S.slice_len() * 8;
    int k = e && atoi(e) > 0 ? atoi(e) : (lb > ((int64_t)1 << 30) ? 8 : (lb > ((int64_t)256 << 20) ? 4 : 2));
    k = std::min<int>(k, EDGPU_MAXCHUNKS);
    k = (int)std::max<int64_t>(1, std::min<int64_t>(k, S.up.ld / SLOW_ROWS / 16));
    S.nchunks = k;
  }
  return 0;
}

constexpr int HALO_THREADS = 512, HALO_UNROLL = 8;
constexpr int HALO_PIECE = HALO_THREADS * HALO_UNROLL;  // double2 per work item: <= 64 KB of one column

// Persistent copy kernel with a SMALL footprint (EDGPU_PUSH_CTAS CTAs, default 40): it runs
// concurrently with the rank-local pass B, whose CTAs need whole SMs (2 x 512 threads x 64
// registers); a grid of one CTA per 16 KB piece was measured to halve pass B's speed.  NVLink
// needs ~1 MB in flight (770 GB/s x ~1 us): each CTA keeps up to 64 KB in flight -- all loads of a
// work item (send-list entry, piece of the column) are issued before its 16-byte stores, which go
// straight into the reader's halo slot over NVLink.
// The halo travels in ROW chunks (rows [r0, r1) of every column, in double2 units): pass A works on
// 16-row tiles, so its tiles of chunk c can start as soon as chunk c of every owner has arrived.
__global__ void __launch_bounds__(HALO_THREADS)
k_halo_push(const double *__restrict__ v, const int32_t *__restrict__ list, int64_t nsend, int par,
            int r0, int r1, HaloTable T) {
  const int n2 = (int)(T.ldU / 2);
  const int len = r1 - r0;
  const int npiece = (len + HALO_PIECE - 1) / HALO_PIECE;
  const int plen = (len + npiece - 1) / npiece;  // even split of the column chunk
  const int64_t nwork = nsend * npiece;
  for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int64_t e = w / npiece;
    const int piece = (int)(w - e * npiece);
    const int32_t src = list[3 * e], dst = list[3 * e + 1], slot = list[3 * e + 2];
    const double2 *s = reinterpret_cast<const double2 *>(v + (int64_t)src * T.ldU);
    double2 *d = reinterpret_cast<double2 *>(T.block[dst] + PIPE_FLAG_BYTES + (int64_t)par * T.hbytes[dst]) +
                 (int64_t)slot * n2;
    const int i0 = r0 + piece * plen + threadIdx.x, iend = min(r1, r0 + (piece + 1) * plen);
    double2 r[HALO_UNROLL];
#pragma unroll
    for (int k = 0; k < HALO_UNROLL; k++)
      if (i0 + k * HALO_THREADS < iend) r[k] = s[i0 + k * HALO_THREADS];
#pragma unroll
    for (int k = 0; k < HALO_UNROLL; k++)
      if (i0 + k * HALO_THREADS < iend) d[i0 + k * HALO_THREADS] = r[k];
  }
}

// The same copy driven by the TMA unit (cp.async.bulk): ONE thread per CTA moves 16 KB pieces
// global -> shared (mbarrier complete_tx) -> the reader's halo in peer memory (bulk_group) through
// a ring of HALO_TMA_STAGES buffers.  No registers, no load/store-unit instructions: the SMs keep
// their issue slots and L1 bandwidth for pass B / pass A running beside it, and 32 such CTAs reach
// the same NVLink rate as 80 CTAs of 512 threads storing with st.global (tools/micro/p2p_bw.cu:
// 694 vs 696 GB/s of the 702 GB/s a kernel can push).
constexpr int HALO_TMA_CHUNK = 16384, HALO_TMA_STAGES = 4;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32)
k_halo_push_tma(const double *__restrict__ v, const int32_t *__restrict__ list, int64_t nsend, int par,
                int r0, int r1, HaloTable T) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) uint64_t full[HALO_TMA_STAGES];
  if (threadIdx.x != 0) return;
  for (int i = 0; i < HALO_TMA_STAGES; i++)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&full[i])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // rows [r0, r1) in double2 units -> bytes [16 r0, 16 r1) of every listed column
  const int64_t cbytes = (int64_t)(r1 - r0) * 16;
  const int64_t npiece = (cbytes + HALO_TMA_CHUNK - 1) / HALO_TMA_CHUNK;
  const int64_t plen = ((cbytes + npiece - 1) / npiece + 15) / 16 * 16;  // even split, 16-byte pieces
  const int64_t nwork = nsend * npiece;
  const int64_t mine = nwork > blockIdx.x ? (nwork - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t colbytes = T.ldU * 8;
  auto item = [&](int64_t i, const unsigned char **sp, unsigned char **dp, uint32_t *bytes) {
    const int64_t w = blockIdx.x + i * gridDim.x;
    const int64_t e = w / npiece, p = w - e * npiece;
    const int32_t src = list[3 * e], dst = list[3 * e + 1], slot = list[3 * e + 2];
    const int64_t off = (int64_t)r0 * 16 + p * plen;
    *sp = reinterpret_cast<const unsigned char *>(v) + (int64_t)src * colbytes + off;
    *dp = T.block[dst] + PIPE_FLAG_BYTES + (int64_t)par * T.hbytes[dst] + (int64_t)slot * colbytes + off;
    const int64_t rem = cbytes - p * plen;
    *bytes = (uint32_t)(rem < plen ? rem : plen);
  };
  int64_t issued = 0, done = 0;
  uint32_t phase = 0;  // bit s = parity of stage s
  while (done < mine) {
    while (issued < mine && issued < done + HALO_TMA_STAGES) {  // keep the ring full of loads
      const int st = (int)(issued % HALO_TMA_STAGES);
      if (issued >= HALO_TMA_STAGES)  // the stage's previous store has finished READING shared memory
        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(HALO_TMA_STAGES - 1) : "memory");
      const unsigned char *sp;
      unsigned char *dp;
      uint32_t bytes;
      item(issued, &sp, &dp, &bytes);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&full[st])), "r"(bytes)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_addr(ring + st * HALO_TMA_CHUNK)),
                   "l"(sp), "r"(bytes), "r"(smem_addr(&full[st]))
                   : "memory");
      issued++;
    }
    const int st = (int)(done % HALO_TMA_STAGES);
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok)
                   : "r"(smem_addr(&full[st])), "r"((phase >> st) & 1u)
                   : "memory");
    phase ^= 1u << st;
    const unsigned char *sp;
    unsigned char *dp;
    uint32_t bytes;
    item(done, &sp, &dp, &bytes);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dp),
                 "r"(smem_addr(ring + st * HALO_TMA_CHUNK)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    done++;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the stores have been performed
}

__device__ __forceinline__ int halo_flag_slot(int par, int c, int sender) {
  return (par * EDGPU_MAXCHUNKS + c) * EDGPU_MAXRANKS + sender;
}

__global__ void k_halo_signal(int par, int c, unsigned long long epoch, HaloTable T) {
  const int p = threadIdx.x;
  if (p >= T.nranks) return;
  __threadfence_system();
  unsigned long long *f = reinterpret_cast<unsigned long long *>(T.block[p]) + halo_flag_slot(par, c, T.me);
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
}

__global__ void k_halo_wait(int par, int c, unsigned long long epoch, HaloTable T,
                            unsigned long long timeout_ns, int32_t *err) {
  const int p = threadIdx.x;
  if (p >= T.nranks) return;
  const unsigned long long *f =
      reinterpret_cast<const unsigned long long *>(T.block[T.me]) + halo_flag_slot(par, c, p);
  unsigned long long t0, t1, seen;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f) : "memory");
    if (seen >= epoch) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) {
      *(volatile int32_t *)err = 1;
      __threadfence_system();
      break;
    }
    __nanosleep(200);
  }
}

static unsigned long long pipe_timeout_ns() {
  static const unsigned long long t = [] {
    const char *e = getenv("EDGPU_PIPE_TIMEOUT_S");
    return (unsigned long long)((e ? atof(e) : 20.0) * 1e9);
  }();
  return t;
}

// rows [*row0, *row1) of row chunk c of the halo (multiples of SLOW_ROWS = 16)
void comm_halo_rows(Engine &E, int c, int64_t *row0, int64_t *row1) {
  const int64_t tiles = E.sec.up.ld / SLOW_ROWS, K = E.sec.nchunks;
  *row0 = tiles * c / K * SLOW_ROWS;
  *row1 = tiles * (c + 1) / K * SLOW_ROWS;
}

int comm_halo_push(Engine &E, int c, const double *d_v, cudaStream_t st) {
  Sector &S = E.sec;
  if (S.nsend <= 0) return 0;
  static const int ctas = [] {
    const char *e = getenv("EDGPU_PUSH_CTAS");
    return e && atoi(e) > 0 ? atoi(e) : 40;
  }();
  int64_t r0, r1;
  comm_halo_rows(E, c, &r0, &r1);
  if (r1 <= r0) return 0;
  // EDGPU_PUSH_MODE=st: threads copy with ld/st.global; default: the TMA unit copies (cp.async.bulk)
  static const bool use_tma = !(getenv("EDGPU_PUSH_MODE") && !strcmp(getenv("EDGPU_PUSH_MODE"), "st"));
  static const int tma_ctas = [] {
    const char *e = getenv("EDGPU_PUSH_CTAS");
    return e && atoi(e) > 0 ? atoi(e) : 48;
  }();
  if (use_tma) {
    const int64_t nwork = S.nsend * (((r1 - r0) * 8 + HALO_TMA_CHUNK - 1) / HALO_TMA_CHUNK);
    static bool attr = false;
    if (!attr) {
      EDGPU_CUDA(cudaFuncSetAttribute(k_halo_push_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      HALO_TMA_CHUNK * HALO_TMA_STAGES));
      attr = true;
    }
    k_halo_push_tma<<<(unsigned)std::min<int64_t>(nwork, tma_ctas), 32, HALO_TMA_CHUNK * HALO_TMA_STAGES, st>>>(
        d_v, S.d_sendlist, S.nsend, (int)(S.epoch & 1), (int)(r0 / 2), (int)(r1 / 2), g_halo);
  } else {
    const int64_t nwork = S.nsend * (((r1 - r0) / 2 + HALO_PIECE - 1) / HALO_PIECE);
    k_halo_push<<<(unsigned)std::min<int64_t>(nwork, ctas), HALO_THREADS, 0, st>>>(
        d_v, S.d_sendlist, S.nsend, (int)(S.epoch & 1), (int)(r0 / 2), (int)(r1 / 2), g_halo);
  }
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int comm_halo_signal(Engine &E, int c, cudaStream_t st) {
  k_halo_signal<<<1, 32, 0, st>>>((int)(E.sec.epoch & 1), c, (unsigned long long)E.sec.epoch, g_halo);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int comm_halo_wait(Engine &E, int c, cudaStream_t st) {
  k_halo_wait<<<1, 32, 0, st>>>((int)(E.sec.epoch & 1), c, (unsigned long long)E.sec.epoch, g_halo,
                                pipe_timeout_ns(), E.sec.pipe_err);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int comm_pipe_check(Engine &E) {
  Sector &S = E.sec;
  if (!S.p2p || !S.pipe_err) return 0;
  const int32_t e = *(volatile int32_t *)S.pipe_err;
  if (e != 0) {
    *(volatile int32_t *)S.pipe_err = 0;
    return set_error("distributed H x v: timed out waiting for the %s flags of a peer (rank %d)",
                     e == 1 ? "push" : "return", E.rank);
  }
  return 0;
}

}  // namespace edgpu
