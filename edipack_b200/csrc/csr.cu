// Stored-H path: CSR sparse matrix - vector product, real and complex, replacing
//   spMatVec_normal_main / spMatVec_mpi_*      ED_HAMILTONIAN_NORMAL_STORED_HxV.f90:517-929
//   spMatVec_nonsu2_main / spMatVec_mpi_nonsu2 ED_HAMILTONIAN_NONSU2_STORED_HxV.f90:194-265
//   spMatVec_superc_main / _mpi_               ED_HAMILTONIAN_SUPERC_STORED_HxV.f90:312-432
//   sp_matvec_matrix_csr_d/_c                  ED_SPARSE_MATRIX.f90:778
// The reference's "CSR" is a list of separately allocated rows (cols(:), vals(:) in insertion
// order, duplicates accumulated at insertion); here the host dumps it into three flat arrays
// once per sector and the device keeps it resident.  Kernel: L threads per row (L = 4..32 by
// average row length), the L threads read consecutive (col,val) pairs (coalesced), gather v
// through L1/L2 and reduce with shuffles: Hv(i) = sum_k vals(k) * v(cols(k)).
// nranks>1: flat row split, the input vector is all-gathered first (the reference does
// MPI_Allgatherv of the whole vector on every call, ..._NONSU2_STORED_HxV.f90:256-259).
#include "edgpu_internal.cuh"

namespace edgpu {

template <int L, bool CPLX, bool ACCUM>
__global__ void __launch_bounds__(256)
k_csr_spmv(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ cols,
           const double *__restrict__ vals, const double *__restrict__ vin,
           double *__restrict__ hv, int64_t nloc, double s_acc, double s_old) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = t / L;
  const int lane = (int)(t % L);
  const bool live = row < nloc;  // whole groups fall out together; keep them for the shuffles
  int64_t k0 = 0, k1 = 0;
  if (live) {
    k0 = rowptr[row];
    k1 = rowptr[row + 1];
  }
  double ar = 0.0, ai = 0.0;
  // four entries of this lane per trip: the (col,val) loads and the four gathers of v are
  // independent and issue back to back (the row's k-range is short: latency, not bandwidth)
  constexpr int U = 4;
  int64_t k = k0 + lane;
  for (; k + (U - 1) * L < k1; k += U * L) {
    int32_t c[U];
#pragma unroll
    for (int u = 0; u < U; u++) c[u] = cols[k + u * L];
    if (CPLX) {
      double2 a[U], x[U];
#pragma unroll
      for (int u = 0; u < U; u++) a[u] = reinterpret_cast<const double2 *>(vals)[k + u * L];
#pragma unroll
      for (int u = 0; u < U; u++) x[u] = reinterpret_cast<const double2 *>(vin)[c[u]];
#pragma unroll
      for (int u = 0; u < U; u++) {
        ar += a[u].x * x[u].x - a[u].y * x[u].y;
        ai += a[u].x * x[u].y + a[u].y * x[u].x;
      }
    } else {
      double a[U], x[U];
#pragma unroll
      for (int u = 0; u < U; u++) a[u] = vals[k + u * L];
#pragma unroll
      for (int u = 0; u < U; u++) x[u] = vin[c[u]];
#pragma unroll
      for (int u = 0; u < U; u++) ar += a[u] * x[u];
    }
  }
  for (; k < k1; k += L) {
    const int32_t c = cols[k];
    if (CPLX) {
      const double2 a = reinterpret_cast<const double2 *>(vals)[k];
      const double2 x = reinterpret_cast<const double2 *>(vin)[c];
      ar += a.x * x.x - a.y * x.y;
      ai += a.x * x.y + a.y * x.x;
    } else {
      ar += vals[k] * vin[c];
    }
  }
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) {
    ar += __shfl_xor_sync(0xffffffffu, ar, o);
    if (CPLX) ai += __shfl_xor_sync(0xffffffffu, ai, o);
  }
  if (live && lane == 0) {
    if (CPLX) {
      double2 *o = reinterpret_cast<double2 *>(hv) + row;
      double2 r = make_double2(s_acc * ar, s_acc * ai);
      if (ACCUM) {
        const double2 h = *o;
        r.x += s_old * h.x;
        r.y += s_old * h.y;
      }
      *o = r;
    } else {
      double r = s_acc * ar;
      if (ACCUM) r += s_old * hv[row];
      hv[row] = r;
    }
  }
}

int csr_close(Engine &E) {
  CsrSector &C = E.csr;
  if (!C.open) return 0;
  cudaStreamSynchronize(E.stream);
  cudaFree(C.rowptr);
  cudaFree(C.cols);
  cudaFree(C.vals);
  cudaFree(C.vfull);
  cudaFree(C.map);
  cudaFree(C.pk_off);
  cudaFree(C.pk_hb);
  C = CsrSector();
  return 0;
}

// lanes per row + the row split of all ranks; marks the sector open
static int csr_finish(Engine &E, const std::vector<int64_t> *counts = nullptr,
                      const std::vector<int64_t> *offs = nullptr) {
  CsrSector &C = E.csr;
  const int64_t nloc = C.nloc, nglobal = C.nglobal, nnz = C.nnz, row0 = C.row0;
  const size_t w = C.cplx ? 2 : 1;
  const double avg = nloc ? (double)nnz / (double)nloc : 0.0;
  // ~6-8 entries per lane: enough independent loads per lane, few idle lanes in the last trip
  C.lanes = avg > 160 ? 32 : (avg > 80 ? 16 : (avg > 20 ? 8 : 4));
  if (E.nranks > 1 && counts && offs) {
    // the builder's own row split (orbital-resolved NORMAL sectors: along the last factor)
    C.counts = *counts;
    C.offs = *offs;
    EDGPU_CUDA(cudaMalloc(&C.vfull, sizeof(double) * w * nglobal));
  } else if (E.nranks > 1) {
    // row split of every rank: MpiQ = Dim/P, remainder to the LAST rank
    // (ED_HAMILTONIAN_NONSU2.f90:72-79, ED_HAMILTONIAN_SUPERC.f90:76-88)
    const int P = E.nranks;
    const int64_t q = nglobal / P;
    C.counts.assign(P, 0);
    C.offs.assign(P, 0);
    for (int p = 0; p < P; p++) {
      const int64_t qp = q + (p == P - 1 ? nglobal % P : 0);
      C.counts[p] = (int64_t)w * qp;
      C.offs[p] = (int64_t)w * q * p;
    }
    if (C.offs[E.rank] != (int64_t)w * row0 || C.counts[E.rank] != (int64_t)w * nloc)
      return set_error("csr_open: (row0=%lld,nloc=%lld) is not rank %d's chunk of the reference row split",
                       (long long)row0, (long long)nloc, E.rank);
    EDGPU_CUDA(cudaMalloc(&C.vfull, sizeof(double) * w * nglobal));
  }
  C.open = true;
  return 0;
}

int csr_adopt_device(Engine &E, bool cplx, int64_t nloc, int64_t nglobal, int64_t row0, int64_t *d_rowptr,
                     int32_t *d_cols, double *d_vals, int64_t nnz, int32_t *d_map,
                     const std::vector<int64_t> *counts, const std::vector<int64_t> *offs) {
  if (E.csr.open) csr_close(E);
  CsrSector &C = E.csr;
  C.cplx = cplx;
  C.nloc = nloc;
  C.nglobal = nglobal;
  C.row0 = row0;
  C.nnz = nnz;
  C.rowptr = d_rowptr;
  C.cols = d_cols;
  C.vals = d_vals;
  C.map = d_map;
  int rc = csr_finish(E, counts, offs);
  if (rc) {  // the caller keeps ownership on failure
    C.rowptr = nullptr;
    C.cols = nullptr;
    C.vals = nullptr;
    C.map = nullptr;
    cudaFree(C.vfull);
    C = CsrSector();
  }
  return rc;
}

int csr_open(Engine &E, bool cplx, int64_t nloc, int64_t nglobal, int64_t row0, const int64_t *rowptr,
             const int32_t *cols, const double *vals) {
  if (!E.inited) return set_error("edgpu_init was not called");
  if (E.sec.open) return set_error("close the direct-H sector before opening a stored-H one");
  if (E.csr.open) csr_close(E);
  if (nloc < 0 || nglobal < nloc || row0 < 0 || row0 + nloc > nglobal)
    return set_error("csr_open: inconsistent sizes nloc=%lld nglobal=%lld row0=%lld", (long long)nloc,
                     (long long)nglobal, (long long)row0);
  if (nglobal > INT32_MAX) return set_error("csr_open: 32-bit column indices limit Dim to 2^31-1");
  if (E.nranks == 1 && (row0 != 0 || nloc != nglobal))
    return set_error("csr_open: a single rank must own all rows");
  CsrSector &C = E.csr;
  C.cplx = cplx;
  C.nloc = nloc;
  C.nglobal = nglobal;
  C.row0 = row0;
  const int64_t nnz = rowptr[nloc] - rowptr[0];
  if (rowptr[0] != 0) return set_error("csr_open: rowptr[0] must be 0");
  C.nnz = nnz;
  // 1-based -> 0-based columns, range check (the reference stops on out-of-range inserts)
  std::vector<int32_t> c0((size_t)std::max<int64_t>(nnz, 1));
  for (int64_t k = 0; k < nnz; k++) {
    if (cols[k] < 1 || cols[k] > nglobal)
      return set_error("csr_open: column %d out of range at entry %lld", (int)cols[k], (long long)k);
    c0[(size_t)k] = cols[k] - 1;
  }
  const size_t w = cplx ? 2 : 1;
  EDGPU_CUDA(cudaMalloc(&C.rowptr, sizeof(int64_t) * (nloc + 1)));
  EDGPU_CUDA(cudaMalloc(&C.cols, sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
  EDGPU_CUDA(cudaMalloc(&C.vals, sizeof(double) * w * std::max<int64_t>(nnz, 1)));
  EDGPU_CUDA(cudaMemcpyAsync(C.rowptr, rowptr, sizeof(int64_t) * (nloc + 1), cudaMemcpyHostToDevice, E.stream));
  if (nnz) {
    EDGPU_CUDA(cudaMemcpyAsync(C.cols, c0.data(), sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, E.stream));
    EDGPU_CUDA(cudaMemcpyAsync(C.vals, vals, sizeof(double) * w * nnz, cudaMemcpyHostToDevice, E.stream));
  }
  EDGPU_CUDA(cudaStreamSynchronize(E.stream));
  return csr_finish(E);
}

template <int L>
static int launch_csr(Engine &E, const double *vin, double *hv, bool accum, double s_acc, double s_old) {
  const CsrSector &C = E.csr;
  const int64_t threads = C.nloc * L;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  if (grid == 0) return 0;
#define EDGPU_CSR(CC, AA)                                                                         \
  k_csr_spmv<L, CC, AA><<<grid, 256, 0, E.stream>>>(C.rowptr, C.cols, C.vals, vin, hv, C.nloc, s_acc, \
                                                    s_old)
  if (C.cplx) {
    if (accum) EDGPU_CSR(true, true); else EDGPU_CSR(true, false);
  } else {
    if (accum) EDGPU_CSR(false, true); else EDGPU_CSR(false, false);
  }
#undef EDGPU_CSR
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

int csr_hxv_device(Engine &E, const double *d_v, double *d_hv, bool accum, double s_acc, double s_old) {
  CsrSector &C = E.csr;
  if (!C.open) return set_error("no stored-H sector open");
  const double *vin = d_v;
  if (E.nranks > 1) {
    EDGPU_TRY(comm_allgatherv(E, d_v, C.vfull, C.counts, C.offs));
    vin = C.vfull;
  }
  if (C.direct)
    return C.direct_orbs ? orbs_direct_hxv(E, vin, d_hv, accum, s_acc, s_old)
                         : packed_direct_hxv(E, vin, d_hv, accum, s_acc, s_old);
  switch (C.lanes) {
    case 32: return launch_csr<32>(E, vin, d_hv, accum, s_acc, s_old);
    case 16: return launch_csr<16>(E, vin, d_hv, accum, s_acc, s_old);
    case 8: return launch_csr<8>(E, vin, d_hv, accum, s_acc, s_old);
    default: return launch_csr<4>(E, vin, d_hv, accum, s_acc, s_old);
  }
}

}  // namespace edgpu
