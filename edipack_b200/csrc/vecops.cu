// Vector kernels of the device-resident Lanczos recurrence (what SciFortran's
// lanczos_iteration does with Fortran array syntax + MPI_Allreduce): every kernel streams
// its operands once with 16-byte accesses and reduces through warp shuffles -> per-block
// partials -> one fixed-order final block (deterministic), then NCCL all-reduce of the scalar.
#include <algorithm>

#include "edgpu_internal.cuh"

namespace edgpu {

constexpr int VT = 256;

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

__device__ __forceinline__ void block_partial(double x, double *__restrict__ part) {
  __shared__ double sh[VT / 32];
  x = warp_sum(x);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    double y = threadIdx.x < VT / 32 ? sh[threadIdx.x] : 0.0;
    y = warp_sum(y);
    if (threadIdx.x == 0) part[blockIdx.x] = y;
  }
}

__global__ void __launch_bounds__(1024) k_final_sum(const double *__restrict__ part, int n,
                                                    double *__restrict__ out) {
  __shared__ double sh[32];
  double x = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) x += part[i];
  x = warp_sum(x);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    double y = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    y = warp_sum(y);
    if (threadIdx.x == 0) *out = y;
  }
}

// n2 = number of double2 elements (vectors are padded to multiples of 16 doubles)
__global__ void __launch_bounds__(VT) k_dot(const double2 *__restrict__ a,
                                            const double2 *__restrict__ b, int64_t n2,
                                            double *__restrict__ part) {
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * VT) {
    double2 x = a[i], y = b[i];
    s += x.x * y.x + x.y * y.y;
  }
  block_partial(s, part);
}

__global__ void __launch_bounds__(VT) k_scale(double2 *__restrict__ a, int64_t n2, double s) {
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * VT) {
    double2 x = a[i];
    x.x *= s;
    x.y *= s;
    a[i] = x;
  }
}

// w -= alpha*v ; partial of <w,w>; optionally a second copy of the new w (the Lanczos vector
// store of the ground-state driver: 8 B/state more, written while the value is in registers)
template <bool STORE>
__global__ void __launch_bounds__(VT) k_axpy_norm(double2 *__restrict__ w,
                                                  const double2 *__restrict__ v, int64_t n2,
                                                  double alpha, double *__restrict__ part,
                                                  double2 *__restrict__ store) {
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * VT) {
    double2 x = w[i], y = v[i];
    x.x -= alpha * y.x;
    x.y -= alpha * y.y;
    w[i] = x;
    if (STORE) store[i] = x;
    s += x.x * x.x + x.y * x.y;
  }
  block_partial(s, part);
}

// out = [out +] sum_b c_b V_b for a batch of up to 8 vectors (second pass of the ground-state
// driver when the Lanczos vectors were kept in HBM)
struct LcBatch {
  const double2 *p[8];
  double c[8];
};
template <bool ACC>
__global__ void __launch_bounds__(VT) k_lincomb(double2 *__restrict__ out, LcBatch B, int nb, int64_t n2) {
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * VT) {
    double2 a = ACC ? out[i] : make_double2(0.0, 0.0);
#pragma unroll
    for (int b = 0; b < 8; b++) {
      if (b < nb) {
        const double2 x = B.p[b][i];
        a.x += B.c[b] * x.x;
        a.y += B.c[b] * x.y;
      }
    }
    out[i] = a;
  }
}

__global__ void __launch_bounds__(VT) k_axpy(double2 *__restrict__ y, const double2 *__restrict__ x,
                                             int64_t n2, double a) {
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * VT) {
    double2 p = y[i], q = x[i];
    p.x += a * q.x;
    p.y += a * q.y;
    y[i] = p;
  }
}

// splitmix64 -> uniform(0,1), indexed by the GLOBAL state index so that the start vector
// does not depend on the number of ranks (the test-side checker regenerates it).
__global__ void __launch_bounds__(VT) k_random(double *__restrict__ v, int64_t nrow, int64_t ld,
                                               int64_t ncol, int64_t col_offset, uint64_t seed,
                                               const int32_t *__restrict__ refup,
                                               const int32_t *__restrict__ refdw, int64_t dimel) {
  const int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= ld) return;
  double x = 0.0;
  v += blockIdx.z * (ld * ncol);  // phonon slice
  if (i < nrow) {
    // reference (ascending Fock order) state index: independent of ranks and internal order
    uint64_t z = ((uint64_t)(refup[i] + (int64_t)refdw[c + col_offset] * nrow + blockIdx.z * dimel) + seed) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    x = ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  v[c * ld + i] = x;
}

static int grid_for(Engine &E, int64_t n2) {
  int64_t want = (n2 + VT - 1) / VT;
  int64_t cap = (int64_t)E.sm_count * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

int ensure_partials(Engine &E, int64_t n) {
  if (E.part_cap >= n) return 0;
  cudaFree(E.d_part);
  E.d_part = nullptr;
  EDGPU_CUDA(cudaMalloc(&E.d_part, sizeof(double) * n));
  E.part_cap = n;
  return 0;
}

int final_sum(Engine &E, int nblocks, double *d_out) {
  k_final_sum<<<1, 1024, 0, E.stream>>>(E.d_part, nblocks, d_out);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int vec_dot_dev(Engine &E, const double *a, const double *b, double *d_out) {
  const int64_t n2 = E.veclen() / 2;
  int gb = grid_for(E, n2);
  k_dot<<<gb, VT, 0, E.stream>>>((const double2 *)a, (const double2 *)b, n2, E.d_part);
  EDGPU_COUNT_LAUNCH();
  return final_sum(E, gb, d_out);
}

int scalar_to_host(Engine &E, double *d_scalar, double *h_out) {
  EDGPU_TRY(comm_allreduce_sum(E, d_scalar, 1));
  EDGPU_CUDA(cudaMemcpyAsync(E.h_scal, d_scalar, sizeof(double), cudaMemcpyDeviceToHost, E.stream));
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  *h_out = E.h_scal[0];
  return comm_pipe_check(E);
}

static int finish_scalar(Engine &E, int nblocks, double *h_out) {
  k_final_sum<<<1, 1024, 0, E.stream>>>(E.d_part, nblocks, E.d_scal);
  EDGPU_COUNT_LAUNCH();
  EDGPU_TRY(comm_allreduce_sum(E, E.d_scal, 1));
  EDGPU_CUDA(cudaMemcpyAsync(E.h_scal, E.d_scal, sizeof(double), cudaMemcpyDeviceToHost, E.stream));
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  *h_out = E.h_scal[0];
  return comm_pipe_check(E);
}

int vec_zero(Engine &E, double *d_v, int64_t n) {
  EDGPU_CUDA(cudaMemsetAsync(d_v, 0, sizeof(double) * n, E.stream));
  return 0;
}

// CSR sectors: plain vector, entry k of the padded array <- hash(global element index)
__global__ void __launch_bounds__(VT) k_random_flat(double *__restrict__ v, int64_t n, int64_t npad,
                                                    int64_t offset, uint64_t seed) {
  const int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x;
  if (i >= npad) return;
  double x = 0.0;
  if (i < n) {
    uint64_t z = ((uint64_t)(i + offset) + seed) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    x = ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  v[i] = x;
}

int vec_fill_random(Engine &E, double *d_v, uint64_t seed) {
  if (E.csr.open) {
    const CsrSector &C = E.csr;
    const int64_t w = C.cplx ? 2 : 1, npad = C.padded_len();
    k_random_flat<<<(unsigned)((npad + VT - 1) / VT), VT, 0, E.stream>>>(d_v, w * C.nloc, npad,
                                                                        w * C.row0, seed);
    EDGPU_COUNT_LAUNCH();
    EDGPU_CUDA(cudaGetLastError());
    return 0;
  }
  Sector &S = E.sec;
  if (S.qdw <= 0) return 0;  // a rank without columns (DimDw < nranks)
  dim3 grid((unsigned)((S.up.ld + VT - 1) / VT), (unsigned)S.qdw, (unsigned)S.DimPh);
  k_random<<<grid, VT, 0, E.stream>>>(d_v, S.up.dim, S.up.ld, S.qdw, S.d0, seed, S.up.refidx,
                                      S.dw.refidx, S.up.dim * S.dw.dim);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int vec_dot(Engine &E, const double *a, const double *b, double *h_out) {
  const int64_t n2 = E.veclen() / 2;
  int gb = grid_for(E, n2);
  k_dot<<<gb, VT, 0, E.stream>>>((const double2 *)a, (const double2 *)b, n2, E.d_part);
  EDGPU_COUNT_LAUNCH();
  return finish_scalar(E, gb, h_out);
}

int vec_scale(Engine &E, double *a, double s) {
  const int64_t n2 = E.veclen() / 2;
  k_scale<<<grid_for(E, n2), VT, 0, E.stream>>>((double2 *)a, n2, s);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int vec_axpy_norm(Engine &E, double *w, const double *v, double alpha, double *h_beta2, double *store) {
  const int64_t n2 = E.veclen() / 2;
  int gb = grid_for(E, n2);
  if (store)
    k_axpy_norm<true><<<gb, VT, 0, E.stream>>>((double2 *)w, (const double2 *)v, n2, alpha, E.d_part,
                                               (double2 *)store);
  else
    k_axpy_norm<false><<<gb, VT, 0, E.stream>>>((double2 *)w, (const double2 *)v, n2, alpha, E.d_part,
                                                nullptr);
  EDGPU_COUNT_LAUNCH();
  return finish_scalar(E, gb, h_beta2);
}

// out = sum_j coef[j] * vecs[j]  (overwrites out)
int vec_lincomb(Engine &E, double *out, const std::vector<double *> &vecs, const std::vector<double> &coef) {
  const int64_t n2 = E.veclen() / 2;
  const int gb = grid_for(E, n2);
  const int m = (int)vecs.size();
  if (m == 0) return vec_zero(E, out, E.veclen());
  for (int f = 0; f < m; f += 8) {
    LcBatch B;
    const int nb = std::min(8, m - f);
    for (int b = 0; b < 8; b++) {
      B.p[b] = (const double2 *)vecs[f + (b < nb ? b : 0)];
      B.c[b] = b < nb ? coef[f + b] : 0.0;
    }
    if (f == 0) k_lincomb<false><<<gb, VT, 0, E.stream>>>((double2 *)out, B, nb, n2);
    else k_lincomb<true><<<gb, VT, 0, E.stream>>>((double2 *)out, B, nb, n2);
    EDGPU_COUNT_LAUNCH();
  }
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int vec_axpy(Engine &E, double *y, const double *x, double a) {
  const int64_t n2 = E.veclen() / 2;
  k_axpy<<<grid_for(E, n2), VT, 0, E.stream>>>((double2 *)y, (const double2 *)x, n2, a);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace edgpu
