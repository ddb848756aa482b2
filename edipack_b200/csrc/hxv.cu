// H x v kernels, NORMAL mode, ed_total_ud=T, DimPh=1:   Hv = (Hd + 1 (x) Hup + Hdw (x) 1 + Hnd) v
// on the vector viewed as a [DimUp (fast), qdw] matrix with padded leading dimension.
//
// Reference loops being replaced (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:23-130, 236-375):
//   direct/HxV_local.f90      -> diagonal, fused into the "up" kernel
//   direct/HxV_up.f90         -> k_up_*   : sparse Hup applied along the fast index
//   direct/HxV_dw.f90         -> k_dw_*   : sparse Hdw applied along the slow index
//   direct/HxV_non_local.f90  -> k_nonlocal
// The reference recomputes every matrix element (bdecomp, c/cdg sign loops, recursive
// binary_search) for every state on every call; here the two small operators are ELL hop
// tables built once per sector (sector.cu) and the kernels are pure streaming + gathers.
//
// Kernel variants
//   generic : thread per state, gathers through L1/L2 (any size; also the fallback for
//             ranges that do not fit shared memory)
//   tiled   : a CTA stages a window of v in shared memory with TMA bulk copies
//             (cp.async.bulk + mbarrier) and serves all in-window hops from it; hops that
//             leave the window fall back to global loads.
#include "edgpu_internal.cuh"

namespace edgpu {

__device__ __forceinline__ double signed_amp(const double *__restrict__ amp, uint32_t ent) {
  double a = amp[(ent >> HOP_AMP_SHIFT) & HOP_AMP_MASK];
  return __hiloint2double(__double2hiint(a) ^ (int)(ent & HOP_SIGN), __double2loint(a));
}

struct SpinView {
  int64_t dim, ld;
  int W;
  int nterms;
  const uint32_t *ell;
  const double *amp;
  const double *eps;
  const uint8_t *imp;
};

static SpinView view_of(const SpinSpace &S) {
  SpinView v;
  v.dim = S.dim;
  v.ld = S.ld;
  v.W = S.W;
  v.nterms = S.nterms;
  v.ell = S.ell;
  v.amp = S.amp;
  v.eps = S.eps;
  v.imp = S.imp;
  return v;
}

// ---------------------------------------------------------------------------------------
// generic kernels
// ---------------------------------------------------------------------------------------
// Hv[:,c] = diag * v[:,c] + Hfast v[:,c]  (+ optional Hslow contribution when all slow
// columns are local).  grid = (ceil(rows/128), ncols)
template <bool WITH_DIAG, bool WITH_SLOW>
__global__ void __launch_bounds__(128)
k_generic(const double *__restrict__ v, double *__restrict__ hv, int64_t nrow, int64_t ldv,
          int64_t ncol, int64_t col_offset, SpinView F, SpinView S,
          const double *__restrict__ xud, int nimp, int accum) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= nrow) return;
  const int64_t cg = c + col_offset;  // global slow index (for eps / imp / slow hops)
  const double *vc = v + c * ldv;
  double acc = accum ? hv[c * ldv + i] : 0.0;
  if (WITH_DIAG) {
    double d = F.eps[i] + S.eps[cg] + xud[(int)S.imp[cg] * nimp + (int)F.imp[i]];
    acc += d * vc[i];
  }
  for (int e = 0; e < F.W; e++) {
    uint32_t ent = F.ell[(int64_t)e * F.ld + i];
    acc += signed_amp(F.amp, ent) * vc[ent & HOP_TGT_MASK];
  }
  if (WITH_SLOW) {
    for (int e = 0; e < S.W; e++) {
      uint32_t ent = S.ell[(int64_t)e * S.ld + cg];
      acc += signed_amp(S.amp, ent) * v[(int64_t)(ent & HOP_TGT_MASK) * ldv + i];
    }
  }
  hv[c * ldv + i] = acc;
}

// ---------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (sm_90+; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  }
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------------------
// tiled "fast index" kernel: CTA = C columns x rows [r0, r0+tl) staged in shared memory.
// shared layout: [C][tile] doubles | amp[nterms+1] | mbarrier
// ---------------------------------------------------------------------------------------
constexpr int UP_THREADS = 512;
constexpr uint32_t BULK_CHUNK = 32768;  // bytes per bulk copy

template <int C, bool WITH_DIAG, bool ACCUM>
__global__ void __launch_bounds__(UP_THREADS, 2)
k_fast_tiled(const double *__restrict__ v, double *__restrict__ hv, int64_t nrow, int64_t ldv,
             int64_t ncol, int64_t col_offset, int64_t tile, SpinView F, SpinView S,
             const double *__restrict__ xud, int nimp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tl = reinterpret_cast<double *>(smem_raw);
  double *amp_s = tl + (size_t)C * tile;
  uint64_t *bar = reinterpret_cast<uint64_t *>(amp_s + ((F.nterms + 1 + 1) & ~1));

  const int64_t r0 = (int64_t)blockIdx.x * tile;
  const int64_t rows = min(tile, F.ld - r0);  // padded rows are zero and 16-aligned
  const int64_t c0 = (int64_t)blockIdx.y * C;
  const int tid = threadIdx.x;

  if (tid == 0) mbar_init(bar, 1);
  for (int t = tid; t <= F.nterms; t += UP_THREADS) amp_s[t] = F.amp[t];
  __syncthreads();
  if (tid == 0) {
    uint32_t total = 0;
#pragma unroll
    for (int c = 0; c < C; c++)
      if (c0 + c < ncol) total += (uint32_t)(rows * 8);
    mbar_expect_tx(bar, total);
#pragma unroll
    for (int c = 0; c < C; c++) {
      if (c0 + c >= ncol) continue;
      const char *src = reinterpret_cast<const char *>(v + (c0 + c) * ldv + r0);
      char *dst = reinterpret_cast<char *>(tl + (size_t)c * tile);
      uint32_t left = (uint32_t)(rows * 8);
      while (left) {
        uint32_t n = left < BULK_CHUNK ? left : BULK_CHUNK;
        tma_bulk_g2s(dst, src, n, bar);
        dst += n;
        src += n;
        left -= n;
      }
    }
  }
  mbar_wait(bar, 0);

  double xrow_d[C];  // slow-index diagonal part per column
  const double *xrow[C];
#pragma unroll
  for (int c = 0; c < C; c++) {
    int64_t cg = min(c0 + c, ncol - 1) + col_offset;
    xrow_d[c] = WITH_DIAG ? S.eps[cg] : 0.0;
    xrow[c] = WITH_DIAG ? xud + (int)S.imp[cg] * nimp : xud;
  }

  for (int64_t il = tid; il < rows; il += UP_THREADS) {
    const int64_t i = r0 + il;
    if (i >= nrow) break;
    double acc[C];
    if (WITH_DIAG) {
      const double eu = F.eps[i];
      const int iu = (int)F.imp[i];
#pragma unroll
      for (int c = 0; c < C; c++) acc[c] = (eu + xrow_d[c] + xrow[c][iu]) * tl[(size_t)c * tile + il];
    } else {
#pragma unroll
      for (int c = 0; c < C; c++) acc[c] = 0.0;
    }
    for (int e = 0; e < F.W; e++) {
      const uint32_t ent = F.ell[(int64_t)e * F.ld + i];
      const double a = signed_amp(amp_s, ent);
      const int64_t t = (int64_t)(ent & HOP_TGT_MASK) - r0;
      if ((uint64_t)t < (uint64_t)rows) {
#pragma unroll
        for (int c = 0; c < C; c++) acc[c] += a * tl[(size_t)c * tile + t];
      } else {
#pragma unroll
        for (int c = 0; c < C; c++)
          if (c0 + c < ncol) acc[c] += a * v[(c0 + c) * ldv + r0 + t];
      }
    }
#pragma unroll
    for (int c = 0; c < C; c++)
      if (c0 + c < ncol) {
        double *o = hv + (c0 + c) * ldv + i;
        *o = ACCUM ? (*o + acc[c]) : acc[c];
      }
  }
}

// ---------------------------------------------------------------------------------------
// tiled "slow index" kernel (single rank): CTA = R consecutive fast rows x the slow range
// [s0, s1) of one segment (states sharing their top bits), staged as tile[j][R].
//   hv[r, s0+j] += sum_e amp * v[r, tgt_e]     in-range targets from shared memory
// ---------------------------------------------------------------------------------------
constexpr int DW_THREADS = 512;

template <int R>
__global__ void __launch_bounds__(DW_THREADS, 2)
k_slow_tiled(const double *__restrict__ v, double *__restrict__ hv, int64_t nrow, int64_t ldv,
             const int64_t *__restrict__ seg_start, SpinView S) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tl = reinterpret_cast<double *>(smem_raw);
  const int seg = blockIdx.y;
  const int64_t s0 = seg_start[seg], s1 = seg_start[seg + 1];
  const int64_t len = s1 - s0;
  double *amp_s = tl + (size_t)len * R;
  const int64_t r0 = (int64_t)blockIdx.x * R;
  const int tid = threadIdx.x;
  for (int t = tid; t <= S.nterms; t += DW_THREADS) amp_s[t] = S.amp[t];
  // stage: each R-row piece (R*8 bytes, contiguous) with 16-byte loads
  constexpr int V2 = R / 2;  // double2 per piece
  for (int64_t q = tid; q < len * V2; q += DW_THREADS) {
    const int64_t j = q / V2;
    const int h = (int)(q % V2);
    const double2 x = *reinterpret_cast<const double2 *>(v + (s0 + j) * ldv + r0 + 2 * h);
    *reinterpret_cast<double2 *>(tl + j * R + 2 * h) = x;
  }
  __syncthreads();
  for (int64_t q = tid; q < len * R; q += DW_THREADS) {
    const int64_t j = q / R;
    const int r = (int)(q % R);
    if (r0 + r >= nrow) continue;
    const int64_t cg = s0 + j;
    double acc = 0.0;
    for (int e = 0; e < S.W; e++) {
      const uint32_t ent = S.ell[(int64_t)e * S.ld + cg];
      const double a = signed_amp(amp_s, ent);
      const int64_t t = (int64_t)(ent & HOP_TGT_MASK);
      const int64_t tloc = t - s0;
      double x;
      if ((uint64_t)tloc < (uint64_t)len)
        x = tl[tloc * R + r];
      else
        x = v[t * ldv + r0 + r];
      acc += a * x;
    }
    hv[cg * ldv + r0 + r] += acc;
  }
}

// ---------------------------------------------------------------------------------------
// non-local S-E / P-H terms (direct/HxV_non_local.f90): gather from anywhere in the vector.
//   S-E: nup(j)=1, ndw(i)=1, ndw(j)=0, nup(i)=0 :  dw: c^+_j c_i ; up: c^+_i c_j ; Jx(i,j)
//   P-H: nup(j)=1, ndw(j)=1, ndw(i)=0, nup(i)=0 :  dw: c^+_i c_j ; up: c^+_i c_j ; Jp(i,j)
// All four operators act on impurity bits, so signs depend on impurity bits only.
// vfull is the full (all-gathered) vector with leading dimension ldv.
// ---------------------------------------------------------------------------------------
struct LinView {
  int lo_bits;
  const int32_t *ja, *jb;
};
__device__ __forceinline__ int lin_rank_v(uint32_t m, LinView L) {
  return L.ja[m >> L.lo_bits] + L.jb[m & ((1u << L.lo_bits) - 1u)];
}
__device__ __forceinline__ double pair_sign(uint32_t m, int a, int b) {
  int lo = min(a, b), hi = max(a, b);
  uint32_t between = ((1u << hi) - 1u) & ~((2u << lo) - 1u);
  return (__popc(m & between) & 1) ? -1.0 : 1.0;
}

__global__ void __launch_bounds__(128)
k_nonlocal(const double *__restrict__ vfull, double *__restrict__ hv, int64_t nrow, int64_t ldv,
           int64_t ncol, int64_t col_offset, const int32_t *__restrict__ mapu,
           const int32_t *__restrict__ mapd, LinView Lu, LinView Ld, int Norb,
           const double *__restrict__ jx, const double *__restrict__ jp) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= nrow) return;
  const uint32_t mu = (uint32_t)mapu[i], md = (uint32_t)mapd[c + col_offset];
  double acc = 0.0;
  for (int io = 0; io < Norb; io++)
    for (int jo = 0; jo < Norb; jo++) {
      if (io == jo) continue;
      const uint32_t bi = 1u << io, bj = 1u << jo;
      const double x = jx[io * Norb + jo];
      if (x != 0.0 && (mu & bj) && (md & bi) && !(md & bj) && !(mu & bi)) {
        // dw: c(iorb), cdg(jorb) ; up: c(jorb), cdg(iorb)
        const uint32_t md2 = (md & ~bi) | bj, mu2 = (mu & ~bj) | bi;
        const double sg = pair_sign(md, io, jo) * pair_sign(mu, io, jo);
        const int64_t iu = lin_rank_v(mu2, Lu), id = lin_rank_v(md2, Ld);
        acc += x * sg * vfull[id * ldv + iu];
      }
      const double y = jp[io * Norb + jo];
      if (y != 0.0 && (mu & bj) && (md & bj) && !(md & bi) && !(mu & bi)) {
        const uint32_t md2 = (md & ~bj) | bi, mu2 = (mu & ~bj) | bi;
        const double sg = pair_sign(md, io, jo) * pair_sign(mu, io, jo);
        const int64_t iu = lin_rank_v(mu2, Lu), id = lin_rank_v(md2, Ld);
        acc += y * sg * vfull[id * ldv + iu];
      }
    }
  if (acc != 0.0) hv[c * ldv + i] += acc;
}

// ---------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------
template <int C, bool WITH_DIAG, bool ACCUM>
static int launch_fast_tiled(Engine &E, const double *v, double *hv, int64_t nrow, int64_t ldv,
                             int64_t ncol, int64_t col_offset, int64_t tile, const SpinView &F,
                             const SpinView &S, const double *xud, int nimp) {
  size_t smem = sizeof(double) * ((size_t)C * tile + ((F.nterms + 2) & ~1)) + 16;
  auto kern = k_fast_tiled<C, WITH_DIAG, ACCUM>;
  EDGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((F.ld + tile - 1) / tile), (unsigned)((ncol + C - 1) / C));
  kern<<<grid, UP_THREADS, smem, E.stream>>>(v, hv, nrow, ldv, ncol, col_offset, tile, F, S, xud,
                                             nimp);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

// Applies (diag +) the fast-index operator F to an [F.ld x ncol] block.
static int apply_fast(Engine &E, bool tiled, bool with_diag, bool accum, const double *v, double *hv,
                      int64_t ncol, int64_t col_offset, int64_t tile, int cols,
                      const SpinView &F, const SpinView &S, const double *xud, int nimp) {
  if (ncol <= 0) return 0;
  if (!tiled) {
    dim3 grid((unsigned)((F.dim + 127) / 128), (unsigned)ncol);
    if (with_diag)
      k_generic<true, false><<<grid, 128, 0, E.stream>>>(v, hv, F.dim, F.ld, ncol, col_offset, F,
                                                          S, xud, nimp, (int)accum);
    else
      k_generic<false, false><<<grid, 128, 0, E.stream>>>(v, hv, F.dim, F.ld, ncol, col_offset,
                                                           F, S, xud, nimp, (int)accum);
    EDGPU_COUNT_LAUNCH();
    EDGPU_CUDA(cudaGetLastError());
    return 0;
  }
#define EDGPU_FAST2(CC, DD, AA) \
  launch_fast_tiled<CC, DD, AA>(E, v, hv, F.dim, F.ld, ncol, col_offset, tile, F, S, xud, nimp)
#define EDGPU_FAST(CC)                                                      \
  (with_diag ? (accum ? EDGPU_FAST2(CC, true, true) : EDGPU_FAST2(CC, true, false)) \
             : (accum ? EDGPU_FAST2(CC, false, true) : EDGPU_FAST2(CC, false, false)))
  switch (cols) {
    case 4: return EDGPU_FAST(4);
    case 2: return EDGPU_FAST(2);
    default: return EDGPU_FAST(1);
  }
#undef EDGPU_FAST2
#undef EDGPU_FAST
}

int hxv_device(Engine &E, const double *d_v, double *d_hv, bool accum, bool timed) {
  Sector &S = E.sec;
  if (!S.open) return set_error("no sector open (build_Hv_sector_normal not called)");
  const SpinView U = view_of(S.up), D = view_of(S.dw);
  const int nimp = 1 << S.Norb;
  const int variant = S.variant == 0 ? 2 : S.variant;
  const bool tiled = (variant == 2);
  cudaStream_t st = E.stream;
  // event sink: the profiling ring when armed, else the 4 scratch events (timed calls only)
  cudaEvent_t *evs = E.ev;
  if (E.prof_on && E.prof_n < E.prof_cap) {
    evs = &E.prof_ev[(size_t)4 * E.prof_n];
    E.prof_n++;
    timed = false;  // no host sync; edgpu_profile_end reads the ring
  } else if (!timed) {
    evs = nullptr;
  }
#define EDGPU_MARK(k) \
  if (evs) cudaEventRecord(evs[k], st)
  EDGPU_MARK(0);

  if (E.nranks == 1) {
    if (!tiled) {
      // one fused gather kernel: diagonal + up hops + dw hops
      dim3 grid((unsigned)((U.dim + 127) / 128), (unsigned)S.qdw);
      k_generic<true, true><<<grid, 128, 0, st>>>(d_v, d_hv, U.dim, U.ld, S.qdw, 0, U, D, S.xud,
                                                  nimp, (int)accum);
      EDGPU_COUNT_LAUNCH();
      EDGPU_CUDA(cudaGetLastError());
      EDGPU_MARK(1);
      EDGPU_MARK(2);
    } else {
      EDGPU_TRY(apply_fast(E, true, true, accum, d_v, d_hv, S.qdw, 0, S.up_tile, S.up_cols, U, D, S.xud,
                           nimp));
      EDGPU_MARK(1);
      if (D.W > 0) {
        constexpr int R = 8;
        size_t smem = sizeof(double) * ((size_t)S.max_seg * R + ((D.nterms + 2) & ~1));
        EDGPU_CUDA(cudaFuncSetAttribute(k_slow_tiled<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
        dim3 grid((unsigned)((U.dim + R - 1) / R), (unsigned)S.nseg);
        k_slow_tiled<R><<<grid, DW_THREADS, smem, st>>>(d_v, d_hv, U.dim, U.ld, S.d_seg_start, D);
        EDGPU_COUNT_LAUNCH();
        EDGPU_CUDA(cudaGetLastError());
      }
      EDGPU_MARK(2);
    }
    if (S.nonlocal) {
      dim3 grid((unsigned)((U.dim + 127) / 128), (unsigned)S.qdw);
      LinView Lu{S.up.lin.lo_bits, S.up.lin.ja, S.up.lin.jb};
      LinView Ld{S.dw.lin.lo_bits, S.dw.lin.ja, S.dw.lin.jb};
      k_nonlocal<<<grid, 128, 0, st>>>(d_v, d_hv, U.dim, U.ld, S.qdw, 0, S.up.map, S.dw.map, Lu, Ld,
                                       S.Norb, S.jx, S.jp);
      EDGPU_COUNT_LAUNCH();
      EDGPU_CUDA(cudaGetLastError());
    }
    EDGPU_MARK(3);
  } else {
    // dw-split over ranks (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:236-375):
    //   Hv  = (Hd + 1 (x) Hup) v                      local columns
    //   vt  = transpose(v)                            NCCL all-to-all of tiles
    //   Hvt = Hdw vt                                  dw is now the fast index
    //   Hv += transpose(Hvt)
    EDGPU_TRY(apply_fast(E, tiled, true, accum, d_v, d_hv, S.qdw, S.d0, S.up_tile, S.up_cols, U, D, S.xud,
                         nimp));
    EDGPU_MARK(1);
    const size_t nt = (size_t)S.padded_len_t();
    if (!S.vt) {
      EDGPU_CUDA(cudaMalloc(&S.vt, sizeof(double) * nt));
      EDGPU_CUDA(cudaMalloc(&S.hvt, sizeof(double) * nt));
      EDGPU_CUDA(cudaMemsetAsync(S.vt, 0, sizeof(double) * nt, st));
      EDGPU_CUDA(cudaMemsetAsync(S.hvt, 0, sizeof(double) * nt, st));
    }
    EDGPU_TRY(comm_transpose(E, d_v, U.dim, U.ld, S.qdw, S.vt, D.dim, D.ld, S.qup, false));
    {
      // tiling plan for the transposed block (fast index = dw)
      const size_t budget = std::min<size_t>(E.smem_optin, 227 * 1024) / 2 - 2048;
      size_t amp_bytes = sizeof(double) * (D.nterms + 2);
      size_t avail = budget - amp_bytes;
      int64_t tile = D.ld;
      int cols = 1;
      if ((size_t)tile * 8 <= avail) {
        while (cols < 4 && (size_t)tile * 8 * (cols * 2) <= avail && cols * 2 <= S.qup) cols *= 2;
      } else {
        int64_t parts = ((size_t)tile * 8 + avail - 1) / avail;
        tile = ((D.ld + parts - 1) / parts + 15) / 16 * 16;
      }
      EDGPU_TRY(apply_fast(E, tiled, false, false, S.vt, S.hvt, S.qup, S.u0, tile, cols, D, U, S.xud, nimp));
    }
    EDGPU_MARK(2);
    EDGPU_TRY(comm_transpose(E, S.hvt, D.dim, D.ld, S.qup, d_hv, U.dim, U.ld, S.qdw, true));
    if (S.nonlocal) return set_error("non-local (Jx/Jp) terms with nranks>1 are not implemented yet");
    EDGPU_MARK(3);
  }
#undef EDGPU_MARK
  if (timed) {
    cudaEventSynchronize(E.ev[3]);
    cudaEventElapsedTime(&E.stage_ms[0], E.ev[0], E.ev[1]);
    cudaEventElapsedTime(&E.stage_ms[1], E.ev[1], E.ev[2]);
    cudaEventElapsedTime(&E.stage_ms[2], E.ev[2], E.ev[3]);
    E.stage_ms[3] = 0.f;
  }
  return 0;
}

}  // namespace edgpu
