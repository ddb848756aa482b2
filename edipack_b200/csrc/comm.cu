// Rank plumbing for the dw-split layout: NCCL (loaded at run time) + the distributed
// matrix transpose that replaces vector_transpose_MPI
// (ED_HAMILTONIAN_NORMAL_COMMON.f90:66-178).
//
// The reference issues one MPI_Alltoallv per local column plus one MPI_Alltoall of counts
// per column (:115-162).  Here each rank packs ONE transposed tile per peer with a
// shared-memory tile-transpose kernel, exchanges all tiles in a single
// ncclGroupStart/End of ncclSend/ncclRecv over NVLink, and unpacks (or accumulates) with a
// strided copy kernel; the rank's own tile never leaves the device.
#include <dlfcn.h>

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
// Peer-memory transposes.  Every rank maps the vt / hvt buffers of all ranks (CUDA IPC, handles
// all-gathered through NCCL) and the two vector_transpose_MPI of one H x v become
//   push : vt_p[(i - u0_p) * ldD + d0_me + c] = v[c * ldU + i]      for every row i, p = owner(i)
//   pull : hv[c * ldU + i] += hvt_p[(i - u0_p) * ldD + d0_me + c]
// one kernel each, 32x32 tiles transposed through shared memory so that both the local and
// the remote side move 256-byte segments; the NVLink traffic is issued by the kernel itself
// (st.global / ld.global on mapped peer pointers), no pack / send / recv / unpack passes.
// Ordering between ranks: in-stream 1-element NCCL all-reduces (comm_barrier).
// ---------------------------------------------------------------------------------------
struct PeerTable {
  double *ptr[EDGPU_MAXRANKS];
  int32_t row0[EDGPU_MAXRANKS + 1];  // rows [row0[p], row0[p+1]) of the fast index belong to rank p
  int nranks, me;
};

// blockIdx.y -> (peer, first row of a 32-row tile of that peer's rows).  Consecutive y rotate over
// the peers, starting at a different peer on every rank: at any instant each GPU exchanges with a
// different partner (the all-to-all schedule) instead of all of them hammering the same NVLink
// endpoint.  Returns false for the padding tiles of peers with fewer rows.
__device__ __forceinline__ bool peer_tile(const PeerTable &T, int y, int *peer, int *ib) {
  const int p = (y % T.nranks + T.me) % T.nranks;
  const int r = T.row0[p] + (y / T.nranks) * 32;
  *peer = p;
  *ib = r;
  return r < T.row0[p + 1];
}

__device__ __forceinline__ int owner_of(const PeerTable &T, int i) {
  int p = 0;
#pragma unroll 1
  while (p + 1 < T.nranks && i >= T.row0[p + 1]) p++;
  return p;
}

// grid = (ceil(qcol/32), ceil(nrow/32)); a = [nrow (fast, lda) x qcol] local block of v.
// Column tiles run fastest: consecutive CTAs then extend the SAME 32 remote rows by consecutive
// 256-byte pieces.  (Row tiles fastest makes every remote segment land on another 2 MB page of a
// multi-GB peer buffer: measured 200 ms instead of ~5 ms per transpose at Ns=18 on 8 GPUs.)
__global__ void __launch_bounds__(256)
k_push_transpose(const double *__restrict__ a, int64_t lda, int qcol, int64_t c_off, int64_t ldb,
                 PeerTable T) {
  __shared__ double t[32][33];
  int p, ib;
  if (!peer_tile(T, blockIdx.y, &p, &ib)) return;
  const int iend = T.row0[p + 1];
  const int jb = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int i = ib + tx, j = jb + ty + k;
    if (i < iend && j < qcol) t[ty + k][tx] = a[(int64_t)j * lda + i];
  }
  __syncthreads();
  double *dst = T.ptr[p] + c_off;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int j = jb + tx, i = ib + ty + k;  // the 32 lanes write 32 consecutive columns of row i
    if (i < iend && j < qcol) dst[(int64_t)(i - T.row0[p]) * ldb + j] = t[tx][ty + k];
  }
  // one system-scope fence per CTA (cumulative over the CTA's stores through the barrier)
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
}

__global__ void __launch_bounds__(256)
k_pull_transpose_acc(double *__restrict__ hv, int64_t lda, int qcol, int64_t c_off, int64_t ldb,
                     PeerTable T) {
  __shared__ double t[32][33];
  int p, ib;
  if (!peer_tile(T, blockIdx.y, &p, &ib)) return;
  const int iend = T.row0[p + 1];
  const int jb = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const double *src = T.ptr[p] + c_off;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int j = jb + tx, i = ib + ty + k;
    if (i < iend && j < qcol) t[ty + k][tx] = src[(int64_t)(i - T.row0[p]) * ldb + j];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int i = ib + tx, j = jb + ty + k;
    if (i < iend && j < qcol) hv[(int64_t)j * lda + i] += t[tx][ty + k];
  }
}

int comm_barrier(Engine &E) {
  if (E.nranks == 1) return 0;
  EDGPU_NCCL(N.AllReduce(E.d_scal + 32, E.d_scal + 32, 1, NCCL_FLOAT64, NCCL_SUM, (nccl_comm_t)E.nccl,
                         E.stream));
  return 0;
}

static PeerTable peer_table(Engine &E, double *const *ptrs) {
  PeerTable T;
  memset(&T, 0, sizeof(T));
  T.nranks = E.nranks;
  T.me = E.rank;
  for (int p = 0; p < E.nranks; p++) {
    int64_t q, r0;
    block_split(E.sec.up.dim, E.nranks, p, &q, &r0);
    T.ptr[p] = ptrs[p];
    T.row0[p] = (int32_t)r0;
    T.row0[p + 1] = (int32_t)(r0 + q);
  }
  return T;
}

// 32-row tiles of the rank with the most rows of the fast index
static int peer_row_tiles(Engine &E) {
  int64_t q, r0;
  block_split(E.sec.up.dim, E.nranks, 0, &q, &r0);  // rank 0 has the longest share
  return (int)((q + 31) / 32);
}

int comm_p2p_setup(Engine &E) {
  Sector &S = E.sec;
  S.p2p = false;
  const int P = E.nranks, me = E.rank;
  if (P == 1 || P > EDGPU_MAXRANKS) return 0;
  const char *off = getenv("EDGPU_NO_P2P");
  // every rank must take the same decision: all-reduce the local "ok" flags
  int ok = (off && off[0] == '1') ? 0 : 1;
  cudaIpcMemHandle_t mine[2];
  if (ok && (cudaIpcGetMemHandle(&mine[0], S.vt) != cudaSuccess ||
             cudaIpcGetMemHandle(&mine[1], S.hvt) != cudaSuccess)) {
    cudaGetLastError();
    ok = 0;
  }
  unsigned char *d_h = nullptr;
  const size_t hb = 2 * sizeof(cudaIpcMemHandle_t);
  EDGPU_CUDA(cudaMalloc(&d_h, hb * P));
  EDGPU_CUDA(cudaMemcpyAsync(d_h + hb * me, mine, hb, cudaMemcpyHostToDevice, E.stream));
  EDGPU_NCCL(N.AllGather(d_h + hb * me, d_h, hb, NCCL_CHAR, (nccl_comm_t)E.nccl, E.stream));
  std::vector<cudaIpcMemHandle_t> all(2 * (size_t)P);
  EDGPU_CUDA(cudaMemcpyAsync(all.data(), d_h, hb * P, cudaMemcpyDeviceToHost, E.stream));
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  cudaFree(d_h);
  for (int p = 0; p < P && ok; p++) {
    if (p == me) {
      S.peer_vt[p] = S.vt;
      S.peer_hvt[p] = S.hvt;
      continue;
    }
    void *a = nullptr, *b = nullptr;
    if (cudaIpcOpenMemHandle(&a, all[2 * p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&b, all[2 * p + 1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
      break;
    }
    S.peer_vt[p] = (double *)a;
    S.peer_hvt[p] = (double *)b;
  }
  double flag = ok ? 0.0 : 1.0;
  EDGPU_CUDA(cudaMemcpyAsync(E.d_scal + 33, &flag, sizeof(double), cudaMemcpyHostToDevice, E.stream));
  EDGPU_NCCL(N.AllReduce(E.d_scal + 33, E.d_scal + 33, 1, NCCL_FLOAT64, NCCL_SUM, (nccl_comm_t)E.nccl,
                         E.stream));
  EDGPU_CUDA(cudaMemcpyAsync(&flag, E.d_scal + 33, sizeof(double), cudaMemcpyDeviceToHost, E.stream));
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  S.p2p = (flag == 0.0);
  if (!S.p2p) {  // someone could not map: everybody unmaps and uses the NCCL path
    for (int p = 0; p < P; p++) {
      if (p != me && S.peer_vt[p]) cudaIpcCloseMemHandle(S.peer_vt[p]);
      if (p != me && S.peer_hvt[p]) cudaIpcCloseMemHandle(S.peer_hvt[p]);
      S.peer_vt[p] = S.peer_hvt[p] = nullptr;
    }
    cudaGetLastError();
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
    if (p != E.rank) {
      cudaIpcCloseMemHandle(S.peer_vt[p]);
      cudaIpcCloseMemHandle(S.peer_hvt[p]);
    }
    S.peer_vt[p] = S.peer_hvt[p] = nullptr;
  }
  comm_barrier(E);  // every rank unmapped before anybody frees
  cudaStreamSynchronize(E.stream);
  S.p2p = false;
  return 0;
}

int comm_push_transpose(Engine &E, const double *d_a) {
  Sector &S = E.sec;
  const PeerTable T = peer_table(E, S.peer_vt);
  dim3 grid((unsigned)((S.qdw + 31) / 32), (unsigned)(E.nranks * peer_row_tiles(E)));
  k_push_transpose<<<grid, 256, 0, E.stream>>>(d_a, S.up.ld, (int)S.qdw, S.d0, S.dw.ld, T);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int comm_pull_transpose_acc(Engine &E, double *d_hv) {
  Sector &S = E.sec;
  const PeerTable T = peer_table(E, S.peer_hvt);
  dim3 grid((unsigned)((S.qdw + 31) / 32), (unsigned)(E.nranks * peer_row_tiles(E)));
  k_pull_transpose_acc<<<grid, 256, 0, E.stream>>>(d_hv, S.up.ld, (int)S.qdw, S.d0, S.dw.ld, T);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace edgpu
