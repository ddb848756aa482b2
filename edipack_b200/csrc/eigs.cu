// Device backend of the thick-restart Lanczos solver (trlan.hpp): the replacement of SciFortran's
// sp_eigh = (P-)ARPACK as called from ed_diag_d / ed_diag_c (ED_DIAG_NORMAL.f90:179-192,
// ED_DIAG_NONSU2.f90:179-192).  The whole ncv-vector basis lives in HBM; per Lanczos step the
// host sees the projection coefficients (a few doubles) only.
//
// Kernels (all streaming, 16-byte accesses, deterministic two-level reductions):
//   k_mdot   h_b = <V_b, w> for a batch of up to 8 basis vectors, w read once per batch
//            (+ |w|^2 for free); complex sectors: conj(V_b) . w, real and imaginary parts
//   k_maxpy  w -= sum_b h_b V_b with the coefficients read from device memory (no host round trip
//            between the dots and the update), |w|^2 of the result accumulated in the last batch
//   k_rotate V[:, 0..k) = V[:, 0..m) Y  in place (restart / final Ritz vectors), Y in shared memory
#include <cmath>
#include <cstdlib>

#include "edgpu_internal.cuh"
#include "trlan.hpp"

namespace edgpu {

constexpr int EIG_MAXCV = 128;            // largest basis (Nblock)
constexpr int EIG_LD = EIG_MAXCV + 8;     // coefficient array: re[0..LD) im[LD..2LD) nb na
constexpr int EIG_NB = 8;                 // vectors per k_mdot / k_maxpy batch
constexpr int EIG_T = 256;

struct VecBatch {
  const double2 *p[EIG_NB];
};
struct VecTable {
  double2 *p[EIG_MAXCV + 1];
};

__device__ __forceinline__ double eig_warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// coefficient slots: vector b of this batch -> re at (first + b), im at (EIG_LD + first + b);
// |w|^2 -> slot 2*EIG_LD when with_norm
template <bool CPLX>
__global__ void __launch_bounds__(EIG_T)
k_mdot(VecBatch B, int nb, int first, const double2 *__restrict__ w, int64_t n2, int with_norm,
       double *__restrict__ part) {
  constexpr int NA = (CPLX ? 2 : 1) * EIG_NB + 1;
  double acc[NA];
#pragma unroll
  for (int a = 0; a < NA; a++) acc[a] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)EIG_T + threadIdx.x; i < n2; i += (int64_t)gridDim.x * EIG_T) {
    const double2 x = w[i];
    acc[NA - 1] += x.x * x.x + x.y * x.y;
#pragma unroll
    for (int b = 0; b < EIG_NB; b++) {
      if (b < nb) {
        const double2 v = B.p[b][i];
        acc[b] += v.x * x.x + v.y * x.y;
        if (CPLX) acc[EIG_NB + b] += v.x * x.y - v.y * x.x;  // Im(conj(v) w)
      }
    }
  }
  int slot[NA];
#pragma unroll
  for (int b = 0; b < EIG_NB; b++) {
    slot[b] = first + (b < nb ? b : 0);
    if (CPLX) slot[EIG_NB + b] = EIG_LD + first + (b < nb ? b : 0);
  }
  slot[NA - 1] = 2 * EIG_LD;
  // inactive accumulators (b >= nb) are written by nobody: mark them through nact ordering
  // (they are zero and alias an active slot, so skip them explicitly)
  __shared__ double sh[NA][EIG_T / 32];
#pragma unroll
  for (int a = 0; a < NA; a++) {
    double x = eig_warp_sum(acc[a]);
    if ((threadIdx.x & 31) == 0) sh[a][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
#pragma unroll
    for (int a = 0; a < NA; a++) {
      double y = threadIdx.x < EIG_T / 32 ? sh[a][threadIdx.x] : 0.0;
      y = eig_warp_sum(y);
      const int b = a == NA - 1 ? -1 : (a % EIG_NB);
      const bool active = (b < 0) ? (with_norm != 0) : (b < nb);
      if (threadIdx.x == 0 && active) part[(size_t)slot[a] * gridDim.x + blockIdx.x] = y;
    }
  }
}

// out[first + blockIdx.x] = sum of the block partials of that slot (fixed order)
__global__ void __launch_bounds__(EIG_T)
k_eig_final(const double *__restrict__ part, int nblk, int first, double *__restrict__ out) {
  __shared__ double sh[EIG_T / 32];
  const int s = first + blockIdx.x;
  double x = 0.0;
  for (int i = threadIdx.x; i < nblk; i += EIG_T) x += part[(size_t)s * nblk + i];
  x = eig_warp_sum(x);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    double y = threadIdx.x < EIG_T / 32 ? sh[threadIdx.x] : 0.0;
    y = eig_warp_sum(y);
    if (threadIdx.x == 0) out[s] = y;
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(EIG_T)
k_maxpy(VecBatch B, int nb, int first, double2 *__restrict__ w, int64_t n2,
        const double *__restrict__ coef, int with_norm, double *__restrict__ part) {
  double cr[EIG_NB], ci[EIG_NB];
#pragma unroll
  for (int b = 0; b < EIG_NB; b++) {
    cr[b] = b < nb ? coef[first + b] : 0.0;
    ci[b] = (CPLX && b < nb) ? coef[EIG_LD + first + b] : 0.0;
  }
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)EIG_T + threadIdx.x; i < n2; i += (int64_t)gridDim.x * EIG_T) {
    double2 x = w[i];
#pragma unroll
    for (int b = 0; b < EIG_NB; b++) {
      if (b < nb) {
        const double2 v = B.p[b][i];
        x.x -= cr[b] * v.x;
        x.y -= cr[b] * v.y;
        if (CPLX) {  // (cr + i ci)(vx + i vy)
          x.x += ci[b] * v.y;
          x.y -= ci[b] * v.x;
        }
      }
    }
    w[i] = x;
    s += x.x * x.x + x.y * x.y;
  }
  if (with_norm) {
    __shared__ double sh[EIG_T / 32];
    s = eig_warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      double y = threadIdx.x < EIG_T / 32 ? sh[threadIdx.x] : 0.0;
      y = eig_warp_sum(y);
      if (threadIdx.x == 0) part[(size_t)(2 * EIG_LD + 1) * gridDim.x + blockIdx.x] = y;
    }
  }
}

// In-place basis rotation: every thread owns one double2 element of all m vectors.
__global__ void __launch_bounds__(EIG_T)
k_rotate(VecTable V, int m, int k, const double *__restrict__ Y, int64_t n2) {
  extern __shared__ double sY[];  // column-major m x k
  for (int t = threadIdx.x; t < m * k; t += EIG_T) sY[t] = Y[t];
  __syncthreads();
  double2 x[EIG_MAXCV];
  for (int64_t i = blockIdx.x * (int64_t)EIG_T + threadIdx.x; i < n2; i += (int64_t)gridDim.x * EIG_T) {
    for (int a = 0; a < m; a++) x[a] = V.p[a][i];
    for (int j = 0; j < k; j++) {
      const double *y = sY + (size_t)j * m;
      double2 acc = make_double2(0.0, 0.0);
      for (int a = 0; a < m; a++) {
        acc.x += x[a].x * y[a];
        acc.y += x[a].y * y[a];
      }
      V.p[j][i] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
struct DeviceOps {
  Engine &E;
  bool cplx;
  int64_t n;  // doubles per vector (padded)
  int nblk;
  std::vector<double *> V;
  double *d_coef = nullptr;  // [2*EIG_LD + 2]
  double *h_coef = nullptr;  // pinned mirror
  double *d_Y = nullptr;     // [EIG_MAXCV * EIG_MAXCV]
  uint64_t seed0 = 0;

  explicit DeviceOps(Engine &e) : E(e), cplx(e.csr.open && e.csr.cplx), n(e.veclen()), nblk(1) {}

  int init(int nslots) {
    const int64_t n2 = n / 2;
    int64_t want = (n2 + EIG_T - 1) / EIG_T;
    const int64_t cap = (int64_t)E.sm_count * 8;
    nblk = (int)std::max<int64_t>(1, std::min(want, cap));
    EDGPU_TRY(ensure_partials(E, (int64_t)(2 * EIG_LD + 2) * nblk));
    V.assign(nslots, nullptr);
    for (auto &p : V) {
      EDGPU_CUDA(cudaMalloc(&p, sizeof(double) * n));
      // the pad entries of a vector must be exactly zero and stay so: the stored-H product does not
      // write them, every other kernel maps zero pads to zero pads
      EDGPU_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * n, E.stream));
    }
    EDGPU_CUDA(cudaMalloc(&d_coef, sizeof(double) * (2 * EIG_LD + 2)));
    EDGPU_CUDA(cudaMemsetAsync(d_coef, 0, sizeof(double) * (2 * EIG_LD + 2), E.stream));
    EDGPU_CUDA(cudaMallocHost(&h_coef, sizeof(double) * (2 * EIG_LD + 2)));
    EDGPU_CUDA(cudaMalloc(&d_Y, sizeof(double) * EIG_MAXCV * EIG_MAXCV));
    EDGPU_CUDA(cudaFuncSetAttribute(k_rotate, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(double) * EIG_MAXCV * EIG_MAXCV));
    return 0;
  }
  // frees everything except slots [0, keep) which are handed to `out`
  void release(int keep, std::vector<double *> *out) {
    for (size_t i = 0; i < V.size(); i++) {
      if ((int)i < keep && out) out->push_back(V[i]);
      else cudaFree(V[i]);
    }
    V.clear();
    cudaFree(d_coef);
    cudaFreeHost(h_coef);
    cudaFree(d_Y);
    d_coef = h_coef = d_Y = nullptr;
  }

  int matvec(int s, int d) { return hxv_device(E, V[s], V[d], false, false); }

  int project_out(int m, int w, double *h, double *nb2, double *na2) {
    const int64_t n2 = n / 2;
    double2 *wv = (double2 *)V[w];
    for (int f = 0; f < m; f += EIG_NB) {
      VecBatch B;
      const int nb = std::min(EIG_NB, m - f);
      for (int b = 0; b < EIG_NB; b++) B.p[b] = (const double2 *)V[f + (b < nb ? b : 0)];
      if (cplx) k_mdot<true><<<nblk, EIG_T, 0, E.stream>>>(B, nb, f, wv, n2, f == 0, E.d_part);
      else k_mdot<false><<<nblk, EIG_T, 0, E.stream>>>(B, nb, f, wv, n2, f == 0, E.d_part);
      EDGPU_COUNT_LAUNCH();
    }
    k_eig_final<<<m, EIG_T, 0, E.stream>>>(E.d_part, nblk, 0, d_coef);
    if (cplx) k_eig_final<<<m, EIG_T, 0, E.stream>>>(E.d_part, nblk, EIG_LD, d_coef);
    k_eig_final<<<1, EIG_T, 0, E.stream>>>(E.d_part, nblk, 2 * EIG_LD, d_coef);
    EDGPU_COUNT_LAUNCH();
    EDGPU_CUDA(cudaGetLastError());
    // one all-reduce of the whole coefficient block (re, im, |w|^2 before)
    EDGPU_TRY(comm_allreduce_sum(E, d_coef, 2 * EIG_LD + 1));
    for (int f = 0; f < m; f += EIG_NB) {
      VecBatch B;
      const int nb = std::min(EIG_NB, m - f);
      for (int b = 0; b < EIG_NB; b++) B.p[b] = (const double2 *)V[f + (b < nb ? b : 0)];
      const int last = f + EIG_NB >= m;
      if (cplx) k_maxpy<true><<<nblk, EIG_T, 0, E.stream>>>(B, nb, f, wv, n2, d_coef, last, E.d_part);
      else k_maxpy<false><<<nblk, EIG_T, 0, E.stream>>>(B, nb, f, wv, n2, d_coef, last, E.d_part);
      EDGPU_COUNT_LAUNCH();
    }
    k_eig_final<<<1, EIG_T, 0, E.stream>>>(E.d_part, nblk, 2 * EIG_LD + 1, d_coef);
    EDGPU_COUNT_LAUNCH();
    EDGPU_CUDA(cudaGetLastError());
    EDGPU_TRY(comm_allreduce_sum(E, d_coef + 2 * EIG_LD + 1, 1));
    EDGPU_CUDA(cudaMemcpyAsync(h_coef, d_coef, sizeof(double) * (2 * EIG_LD + 2), cudaMemcpyDeviceToHost,
                               E.stream));
    EDGPU_CUDA(cudaStreamSynchronize(E.stream));
    for (int i = 0; i < m; i++) h[i] = h_coef[i];
    *nb2 = h_coef[2 * EIG_LD];
    *na2 = h_coef[2 * EIG_LD + 1];
    return 0;
  }

  int scale(int w, double s) { return vec_scale(E, V[w], s); }

  int rotate(int m, int k, const double *Y) {
    if (m > EIG_MAXCV) return set_error("edgpu_eigh: basis larger than %d", EIG_MAXCV);
    if (k <= 0) return 0;
    EDGPU_CUDA(cudaMemcpyAsync(d_Y, Y, sizeof(double) * (size_t)m * k, cudaMemcpyHostToDevice, E.stream));
    VecTable T;
    for (int a = 0; a <= EIG_MAXCV; a++) T.p[a] = (double2 *)V[std::min<size_t>(a, V.size() - 1)];
    const int64_t n2 = n / 2;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n2 + EIG_T - 1) / EIG_T, (int64_t)E.sm_count * 4));
    k_rotate<<<grid, EIG_T, sizeof(double) * (size_t)m * k, E.stream>>>(T, m, k, d_Y, n2);
    EDGPU_COUNT_LAUNCH();
    EDGPU_CUDA(cudaGetLastError());
    EDGPU_CUDA(cudaStreamSynchronize(E.stream));  // Y (host) may be reused by the caller
    return 0;
  }

  int swap(int a, int b) {
    std::swap(V[a], V[b]);
    return 0;
  }
  int randomize(int w, uint64_t seed) { return vec_fill_random(E, V[w], seed); }
  int norm2(int w, double *out) { return vec_dot(E, V[w], V[w], out); }
};

// Lowest `neigen` eigenpairs of the open sector.  On success `vecs` receives neigen device
// vectors (padded layout, ownership passes to the caller).
int eigh_dev(Engine &E, int neigen, int nblock, int nitermax, double tol, uint64_t seed, double *evals,
             double *resid, std::vector<double *> *vecs, int *nconv, int *nmatvec) {
  const int64_t dim = E.csr.open ? E.csr.nglobal : E.sec.up.dim * E.sec.dw.dim;
  if (neigen < 1) return set_error("edgpu_eigh: Neigen must be >= 1");
  if ((int64_t)neigen > dim) neigen = (int)dim;
  if (nblock > EIG_MAXCV) nblock = EIG_MAXCV;
  if ((int64_t)nblock > dim) nblock = (int)dim;
  if (nblock < neigen) return set_error("edgpu_eigh: Nblock=%d < Neigen=%d", nblock, neigen);
  // A single eigenpair does not need a re-orthogonalised basis: the plain Lanczos recurrence
  // converges to the lowest pair regardless of the loss of orthogonality, at 1/5 of the memory
  // traffic per step.  Same stopping rule as below (Ritz estimate against tol), vectors of
  // pass 1 kept in HBM.  EDGPU_EIGH_FULL=1 forces the thick-restart path.
  const char *env_full = getenv("EDGPU_EIGH_FULL");
  if (neigen == 1 && dim > 1 && !(env_full && env_full[0] == '1')) {
    const int64_t n = E.veclen();
    double *vec = nullptr;
    EDGPU_CUDA(cudaMalloc(&vec, sizeof(double) * n));
    int nit = 0;
    const double eps = 2.220446049250313e-16;
    const int nmax = (int)std::min<int64_t>(dim, std::max(nitermax, 64));
    int rc = lanczos_gs_dev(E, nmax, 0.0, 2, nullptr, seed, evals, vec, &nit, std::max(tol, eps));
    if (rc) {
      cudaFree(vec);
      return rc;
    }
    const bool ok = g_lanczos_last_resid <= std::max(tol, eps) * std::max(3.666852862501036e-11, std::fabs(evals[0]))
                    || nit >= dim;
    if (resid) resid[0] = g_lanczos_last_resid;
    if (nconv) *nconv = ok ? 1 : 0;
    if (nmatvec) *nmatvec = g_lanczos_last_hxv;
    vecs->push_back(vec);
    return 0;
  }
  DeviceOps ops(E);
  int rc = ops.init(nblock + 1);
  if (rc) {
    ops.release(0, nullptr);
    return rc;
  }
  TrlanResult R;
  rc = trlan_solve(ops, dim, neigen, nblock, nitermax, tol, seed, evals, resid, &R);
  if (rc > 0 && g_status == 0) set_error("edgpu_eigh: thick-restart Lanczos failed (code %d)", rc);
  if (rc) {
    ops.release(0, nullptr);
    return rc;
  }
  if (nconv) *nconv = R.nconv;
  if (nmatvec) *nmatvec = R.nmatvec;
  ops.release(neigen, vecs);
  return 0;
}

}  // namespace edgpu
