// H x v kernels, NORMAL mode, ed_total_ud=T, DimPh=1:   Hv = (Hd + 1 (x) Hup + Hdw (x) 1 + Hnd) v
// on the vector viewed as a [DimUp (fast), qdw] matrix with padded leading dimension.
//
// Reference loops being replaced (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:23-130, 236-375):
//   direct/HxV_local.f90      -> diagonal, fused into the "fast" kernel
//   direct/HxV_up.f90         -> k_fast    : sparse Hup applied along the fast index
//   direct/HxV_dw.f90         -> k_slow    : sparse Hdw applied along the slow index
//   direct/HxV_non_local.f90  -> k_nonlocal
// The reference recomputes every matrix element (bdecomp, c/cdg sign loops, recursive
// binary_search) for every state on every call; here the two small operators are ELL hop
// tables built once per sector (sector.cu) and the kernels are streaming passes whose gathers
// are served from shared memory.
//
// Two-pass structure (DESIGN.md "Why two passes"): a tile that is closed under the up hops is a
// set of full columns, a tile closed under the dw hops is a set of full rows; no 227 KB tile is
// closed under both, so
//   pass B  k_fastc (k_fastb / k_fast) : diagonal + all up hops (+ the non-local terms on one rank);
//                              CTA = one work item of the up species x 4 columns, tile staged by the TMA
//                              unit (k_fastc, default); k_fastb = the thread-staged round-1 form, k_fast =
//                              range mode (2 columns) for species without a block plan
//   pass A  k_slow           : all dw hops; CTA = 16 rows x one dw range, accumulates onto pass B.
//                              With nranks > 1 it runs on the rank's chunk of dw columns; far targets on
//                              other ranks are read from the halo their owners pushed (comm.cu)
// A "range" is a maximal run of sector states sharing a prefix of top (permuted) bits; a "block"
// is the part of a range with one impurity configuration (sector.cu).  Hops that leave the
// range ("far", they move an electron into / out of the prefix bits) are read from global
// memory (L2: sibling ranges of the same rows / columns are scheduled next to each other).
#include <cstring>

#include "edgpu_internal.cuh"

namespace edgpu {

struct SpinView {
  int64_t dim, ld;
  int Wl4, Wf4;  // local / far groups of 4 entries
  int Wf;        // largest number of far entries of a row
  int nterms;
  const uint4 *ell4;
  const double *amp2;
  const double *eps;
  const uint8_t *imp;
  const int64_t *range_start;
  int nranges;
  // block mode (fast role): local entries are tile offsets of the row's work item
  int block_mode, nitems;
  const BlockItem *items;
  const int32_t *item_of_row;
  // sharded slow role (halo mode): table row of local column 0, number of local columns (targets
  // >= qloc are halo slots); unsharded: 0 and UINT32_MAX
  int64_t row_off;
  uint32_t qloc;
};

// non-local S-E / P-H terms fused into pass B (single rank): impurity hop tables of both species
// (sector.cu, k_imphop_fill), couplings, and the whole vector the gathers read
struct NlDev {
  const int32_t *upT, *dwT;
  int64_t ldu, ldd;
  const double *jx, *jp;
  const double *vfull;
  int Norb;
};

static SpinView view_of(const SpinSpace &S) {
  SpinView v;
  v.dim = S.dim;
  v.ld = S.ld;
  v.Wl4 = S.Wl4;
  v.Wf4 = S.Wf4;
  v.Wf = S.Wf;
  v.nterms = S.nterms;
  v.ell4 = S.ell4;
  v.amp2 = S.amp2;
  v.eps = S.eps;
  v.imp = S.imp;
  v.range_start = S.d_range_start;
  v.nranges = S.nranges;
  v.block_mode = S.block_mode ? 1 : 0;
  v.nitems = (int)S.items.size();
  v.items = S.d_items;
  v.item_of_row = S.d_item_of_row;
  v.row_off = S.sharded ? S.shard0 : 0;
  v.qloc = S.sharded ? (uint32_t)S.shard_q : 0xFFFFFFFFu;
  return v;
}

// explicit shared-window loads (32-bit addresses: no generic->shared conversion per access)
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ double2 lds128(uint32_t a) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ double lds64(uint32_t a) {
  double r;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a));
  return r;
}
// byte offset of amp2[idx] from the packed entry
__device__ __forceinline__ uint32_t amp_off(uint32_t ent) {
  return (ent >> (HOP_AMP_SHIFT - 3)) & (HOP_AMP_MASK << 3);
}

__device__ __forceinline__ uint32_t ent_of(const uint4 &q, int k) {
  return k == 0 ? q.x : (k == 1 ? q.y : (k == 2 ? q.z : q.w));
}

// ---------------------------------------------------------------------------------------
// generic kernel: thread per state, gathers through L1/L2 (any size; parity cross-check of
// the tiled kernels = "variant 1")
// ---------------------------------------------------------------------------------------
template <bool WITH_DIAG, bool WITH_SLOW>
__global__ void __launch_bounds__(128)
k_generic(const double *__restrict__ v, double *__restrict__ hv, int64_t nrow, int64_t ldv,
          int64_t ncol, int64_t col_offset, SpinView F, SpinView S,
          const double *__restrict__ xud, int nimp, int accum, double s_acc, double s_old,
          const double *__restrict__ halo) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= nrow) return;
  const int64_t cg = c + col_offset;  // global slow index (for eps / imp / slow hops)
  const double *vc = v + c * ldv;
  double acc = 0.0;
  if (WITH_DIAG) {
    double d = F.eps[i] + S.eps[cg] + xud[(int)S.imp[cg] * nimp + (int)F.imp[i]];
    acc += d * vc[i];
  }
  for (int g = 0; g < F.Wl4 + F.Wf4; g++) {
    const uint4 q = F.ell4[(int64_t)g * F.ld + i];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t ent = ent_of(q, k);
      const uint32_t idx = (ent >> HOP_AMP_SHIFT) & HOP_AMP_MASK;
      uint32_t t = ent & HOP_TGT_MASK;
      if (F.block_mode && g < F.Wl4) {  // tile offset -> global row (padding: amplitude 0)
        if (idx == 2u * (uint32_t)F.nterms) continue;
        t = (uint32_t)block_global_row(F.items[F.item_of_row[i]], (int)t);
      }
      acc += F.amp2[idx] * vc[t];
    }
  }
  if (WITH_SLOW) {
    for (int g = 0; g < S.Wl4 + S.Wf4; g++) {
      const uint4 q = S.ell4[(int64_t)g * S.ld + cg];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t ent = ent_of(q, k);
        const uint32_t t = ent & HOP_TGT_MASK;  // sharded species: local column or qloc + halo slot
        const double x = t >= S.qloc ? halo[(int64_t)(t - S.qloc) * ldv + i] : v[(int64_t)t * ldv + i];
        acc += S.amp2[(ent >> HOP_AMP_SHIFT) & HOP_AMP_MASK] * x;
      }
    }
  }
  hv[c * ldv + i] = accum ? s_acc * acc + s_old * hv[c * ldv + i] : s_acc * acc;
}

// ---------------------------------------------------------------------------------------
// pass B, "fast index" kernel.  CTA = rows [r0, r1) of one range x 2 columns, staged in shared
// memory interleaved as tile[row][2] so that one 16-byte shared load serves both columns.
// grid = (nranges, ceil(ncol / 2)); range index fastest so that the ranges of one column pair
// run next to each other and the far gathers hit L2.
// shared layout: tile[trows][2] | xc[2][nimp] | amp[2*nterms+2]
// ---------------------------------------------------------------------------------------
constexpr int FAST_THREADS = 512;

// WL4 = local entry groups per row (0 = dynamic widths, plain loop), NFAR = far slots per row
// (all in the first far group).  The per-row metadata (entries, eps, impurity bits) of the
// thread's next row is prefetched one iteration ahead; far gathers are issued before the local
// hops and consumed after them.
template <int WL4, int NFAR, bool WITH_DIAG, bool ACCUM>
__global__ void __launch_bounds__(FAST_THREADS, 2)
k_fast(const double *__restrict__ v, double *__restrict__ hv, int64_t ldv, int64_t ncol,
       int64_t col_offset, SpinView F, SpinView S, const double *__restrict__ xud, int nimp,
       double s_acc, double s_old) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int r0 = (int)F.range_start[blockIdx.x], r1 = (int)F.range_start[blockIdx.x + 1];
  const int tr0 = r0 & ~15;
  const int tr1 = (r1 + 15) & ~15;  // <= F.ld
  const int trows = tr1 - tr0;
  double2 *tile = reinterpret_cast<double2 *>(smem_raw);
  double *xc = reinterpret_cast<double *>(tile + trows);
  double *amp_s = xc + 2 * nimp;

  const int64_t c0 = (int64_t)blockIdx.y * 2;
  const bool has2 = (c0 + 1 < ncol);
  const double *v0 = v + c0 * ldv;
  const double *v1 = v + (has2 ? c0 + 1 : c0) * ldv;

  // stage: one row of both columns per thread: coalesced 8-byte global loads, contiguous
  // (conflict-free) 16-byte shared stores
#pragma unroll 4
  for (int p = tid; p < trows; p += FAST_THREADS)
    tile[p] = make_double2(v0[tr0 + p], v1[tr0 + p]);
  for (int t = tid; t < 2 * F.nterms + 2; t += FAST_THREADS) amp_s[t] = F.amp2[t];
  if (WITH_DIAG) {
    // xc[cc][m] = eps_S(c) + X[imp_S(c)][m]
    for (int t = tid; t < 2 * nimp; t += FAST_THREADS) {
      const int cc = t / nimp, m = t - cc * nimp;
      const int64_t cg = (cc == 0 || has2 ? c0 + cc : c0) + col_offset;
      xc[t] = S.eps[cg] + xud[(int)S.imp[cg] * nimp + m];
    }
  }
  __syncthreads();

  const uint32_t PAD = 2u * (uint32_t)F.nterms;
  double *h0 = hv + c0 * ldv;
  double *h1 = hv + (c0 + 1) * ldv;
  const uint32_t tile_sa = smem_u32(tile) - (uint32_t)tr0 * 16u;  // + row * 16
  const uint32_t amp_sa = smem_u32(amp_s);
  const uint32_t xc_sa = smem_u32(xc);
  // the last range also owns the pad rows [dim, ld): their entries are all padding -> zeros
  const int rend = (blockIdx.x + 1 == gridDim.x) ? (int)F.ld : r1;

  if (WL4 > 0) {
    constexpr int NG = WL4 > 0 ? WL4 : 1;
    const uint4 *ellf = F.ell4 + (int64_t)F.Wl4 * F.ld;  // first far group
    uint4 nq[NG], nqf = make_uint4(0, 0, 0, 0);
    double neu = 0.0;
    uint32_t nm = 0;
    int i = r0 + tid;
    if (i < rend) {
#pragma unroll
      for (int g = 0; g < WL4; g++) nq[g] = F.ell4[(int64_t)g * F.ld + i];
      if (NFAR > 0) nqf = ellf[i];
      if (WITH_DIAG) {
        neu = F.eps[i];
        nm = (uint32_t)F.imp[i];
      }
    }
    for (; i < rend; i += FAST_THREADS) {
      uint4 q[NG];
#pragma unroll
      for (int g = 0; g < WL4; g++) q[g] = nq[g];
      const uint4 qf = nqf;
      const double eu = neu;
      const uint32_t m = nm;
      double2 hacc = make_double2(0.0, 0.0);
      if (ACCUM) {
        hacc.x = h0[i];
        if (has2) hacc.y = h1[i];
      }
      // far gathers first (L2), consumed after the local hops
      double2 xf[NFAR > 0 ? NFAR : 1];
#pragma unroll
      for (int k = 0; k < NFAR; k++) {
        const uint32_t ent = ent_of(qf, k);
        const uint32_t t = ent & HOP_TGT_MASK;  // padding entries point at the row itself
        xf[k] = make_double2(v0[t], v1[t]);
      }
      const int inext = i + FAST_THREADS;
      if (inext < rend) {
#pragma unroll
        for (int g = 0; g < WL4; g++) nq[g] = F.ell4[(int64_t)g * F.ld + inext];
        if (NFAR > 0) nqf = ellf[inext];
        if (WITH_DIAG) {
          neu = F.eps[inext];
          nm = (uint32_t)F.imp[inext];
        }
      }
      double2 acc = make_double2(0.0, 0.0);
      if (WITH_DIAG) {
        const double2 own = lds128(tile_sa + (uint32_t)i * 16u);
        acc.x = (eu + lds64(xc_sa + m * 8u)) * own.x;
        acc.y = (eu + lds64(xc_sa + ((uint32_t)nimp + m) * 8u)) * own.y;
      }
#pragma unroll
      for (int g = 0; g < WL4; g++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t ent = ent_of(q[g], k);
          const double a = lds64(amp_sa + amp_off(ent));
          const double2 x = lds128(tile_sa + (ent & HOP_TGT_MASK) * 16u);
          acc.x += a * x.x;
          acc.y += a * x.y;
        }
      }
#pragma unroll
      for (int k = 0; k < NFAR; k++) {
        const double a = lds64(amp_sa + amp_off(ent_of(qf, k)));  // padding -> amplitude 0
        acc.x += a * xf[k].x;
        acc.y += a * xf[k].y;
      }
      // Hv = s_acc * (H v) [+ s_old * Hv_old] : the Lanczos drivers fold 1/|v| and -beta here
      acc.x *= s_acc;
      acc.y *= s_acc;
      if (ACCUM) {
        acc.x += s_old * hacc.x;
        acc.y += s_old * hacc.y;
      }
      h0[i] = acc.x;
      if (has2) h1[i] = acc.y;
    }
  } else {
    for (int i = r0 + tid; i < rend; i += FAST_THREADS) {
      double2 acc = make_double2(0.0, 0.0);
      if (WITH_DIAG) {
        const double eu = F.eps[i];
        const uint32_t m = (uint32_t)F.imp[i];
        const double2 own = lds128(tile_sa + (uint32_t)i * 16u);
        acc.x = (eu + lds64(xc_sa + m * 8u)) * own.x;
        acc.y = (eu + lds64(xc_sa + ((uint32_t)nimp + m) * 8u)) * own.y;
      }
      double2 hacc = make_double2(0.0, 0.0);
      if (ACCUM) {
        hacc.x = h0[i];
        if (has2) hacc.y = h1[i];
      }
      for (int g = 0; g < F.Wl4; g++) {
        const uint4 qq = F.ell4[(int64_t)g * F.ld + i];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t ent = ent_of(qq, k);
          const double a = lds64(amp_sa + amp_off(ent));
          const double2 x = lds128(tile_sa + (ent & HOP_TGT_MASK) * 16u);
          acc.x += a * x.x;
          acc.y += a * x.y;
        }
      }
      for (int g = 0; g < F.Wf4; g++) {
        const uint4 qq = F.ell4[(int64_t)(F.Wl4 + g) * F.ld + i];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (4 * g + k >= F.Wf) break;  // uniform: slots beyond the widest far row
          const uint32_t ent = ent_of(qq, k);
          const uint32_t idx = (ent >> HOP_AMP_SHIFT) & HOP_AMP_MASK;
          if (idx != PAD) {
            const double a = lds64(amp_sa + idx * 8u);
            const uint32_t t = ent & HOP_TGT_MASK;
            acc.x += a * v0[t];
            acc.y += a * v1[t];
          }
        }
      }
      acc.x *= s_acc;
      acc.y *= s_acc;
      if (ACCUM) {
        acc.x += s_old * hacc.x;
        acc.y += s_old * hacc.y;
      }
      h0[i] = acc.x;
      if (has2) h1[i] = acc.y;
    }
  }
}

// ---------------------------------------------------------------------------------------
// pass B in block mode.  CTA = one work item (the rows of one range prefix and one impurity
// configuration) x 4 columns.  Only the blocks the item's hops read from are staged (for
// imp<->bath hops: the other half of the range), as two planes tile[row][2] of column pairs, so
// that one 16-byte shared load serves two columns and per-row metadata (entries, amplitudes,
// eps) are amortised over 4 columns.  The rows' own values (diagonal term) come from global
// memory.  grid = (nitems, ceil(ncol / 4)), item index fastest (far gathers hit L2).
// shared layout: planeA[tile_cap] | planeB[tile_cap] (double2) | xc[4][nimp] | amp[2*nterms+2]
// ---------------------------------------------------------------------------------------
constexpr int FASTB_THREADS = 512;

template <int WL4, int NFAR, bool WITH_DIAG, bool ACCUM>
__global__ void __launch_bounds__(FASTB_THREADS, 2)
k_fastb(const double *__restrict__ v, double *__restrict__ hv, int64_t ldv, int64_t ncol,
        int64_t col_offset, SpinView F, SpinView S, const double *__restrict__ xud, int nimp,
        double s_acc, double s_old, int tile_cap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  double2 *pa = reinterpret_cast<double2 *>(smem_raw);
  double2 *pb = pa + tile_cap;
  double *xc = reinterpret_cast<double *>(pb + tile_cap);
  double *amp_s = xc + 4 * nimp;
  const BlockItem &it = F.items[blockIdx.x];
  const int out0 = it.out0, out1 = it.out1, nin = it.nin;

  const int64_t c0 = (int64_t)blockIdx.y * 4;
  const int nc = (int)min((int64_t)4, ncol - c0);  // live columns of this CTA
  // column k of the CTA lives at element offset ko[k] from the first one (CTA-uniform; missing
  // columns of the last CTA alias column 0 and are never stored)
  const double *vb = v + c0 * ldv;
  double *hb = hv + c0 * ldv;
  const uint32_t ld32 = (uint32_t)ldv;
  const uint32_t ko[4] = {0u, nc > 1 ? ld32 : 0u, nc > 2 ? 2u * ld32 : 0u, nc > 3 ? 3u * ld32 : 0u};
  // stage the input blocks: coalesced 8-byte global loads, contiguous 16-byte shared stores
  for (int b = 0; b < nin; b++) {
    const int g0 = it.in0[b], len = it.in_len[b], off = it.in_off[b];
#pragma unroll 2
    for (int p = tid; p < len; p += FASTB_THREADS) {
      const uint32_t g = (uint32_t)(g0 + p);
      pa[off + p] = make_double2(vb[ko[0] + g], vb[ko[1] + g]);
      pb[off + p] = make_double2(vb[ko[2] + g], vb[ko[3] + g]);
    }
  }
  if ((nin == 0 || it.in_off[0] != 0) && tid == 0) {  // padding entries read slot 0 (not staged then)
    pa[0] = make_double2(0.0, 0.0);
    pb[0] = make_double2(0.0, 0.0);
  }
  // The rows' own values (diagonal term) and, when accumulating, the old Hv are read from
  // global memory late in each row iteration; pull their lines into L2 now so that those loads
  // see L2 latency only (they are not kept in registers across the hop loop).
  {
    const int l0 = out0 & ~15, nl = ((out1 + 15) & ~15) - l0;  // whole 128-byte lines
    for (int p = tid * 16; p < nl; p += FASTB_THREADS * 16) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (WITH_DIAG) asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + ko[k] + (uint32_t)(l0 + p)));
        if (ACCUM) asm volatile("prefetch.global.L2 [%0];" ::"l"(hb + ko[k] + (uint32_t)(l0 + p)));
      }
    }
  }
  for (int t = tid; t < 2 * F.nterms + 2; t += FASTB_THREADS) amp_s[t] = F.amp2[t];
  if (WITH_DIAG) {
    // xc[k][m] = eps_S(c) + X[imp_S(c)][m]
    for (int t = tid; t < 4 * nimp; t += FASTB_THREADS) {
      const int k = t / nimp, m = t - k * nimp;
      const int64_t cg = c0 + (k < nc ? k : 0) + col_offset;
      xc[t] = S.eps[cg] + xud[(int)S.imp[cg] * nimp + m];
    }
  }
  __syncthreads();

  const uint32_t pa_sa = smem_u32(pa), pb_sa = smem_u32(pb);
  const uint32_t amp_sa = smem_u32(amp_s), xc_sa = smem_u32(xc);
  // first far group, entry e of row i at far32[4 * i + e]
  const uint32_t *far32 = reinterpret_cast<const uint32_t *>(F.ell4 + (int64_t)F.Wl4 * F.ld);
  constexpr int NG = WL4 > 0 ? WL4 : 1;
  constexpr int NFE = NFAR > 0 ? NFAR : 1;
  uint4 nq[NG];
  uint32_t nqf[NFE];
  double neu = 0.0;
  uint32_t nm = 0;
  int i = out0 + tid;
  if (i < out1) {
#pragma unroll
    for (int g = 0; g < WL4; g++) nq[g] = F.ell4[(int64_t)g * F.ld + i];
#pragma unroll
    for (int e = 0; e < NFAR; e++) nqf[e] = far32[4 * (int64_t)i + e];
    if (WITH_DIAG) {
      neu = F.eps[i];
      nm = (uint32_t)F.imp[i];
    }
  }
  for (; i < out1; i += FASTB_THREADS) {
    uint4 q[NG];
#pragma unroll
    for (int g = 0; g < WL4; g++) q[g] = nq[g];
    uint32_t qf[NFE];
#pragma unroll
    for (int e = 0; e < NFAR; e++) qf[e] = nqf[e];
    const double eu = neu;
    const uint32_t m = nm;
    const int inext = i + FASTB_THREADS;
    if (inext < out1) {
#pragma unroll
      for (int g = 0; g < WL4; g++) nq[g] = F.ell4[(int64_t)g * F.ld + inext];
#pragma unroll
      for (int e = 0; e < NFAR; e++) nqf[e] = far32[4 * (int64_t)inext + e];
      if (WITH_DIAG) {
        neu = F.eps[inext];
        nm = (uint32_t)F.imp[inext];
      }
    }
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int g = 0; g < WL4; g++) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t ent = ent_of(q[g], k);
        const double a = lds64(amp_sa + amp_off(ent));
        const uint32_t o = (ent & HOP_TGT_MASK) * 16u;
        const double2 xa = lds128(pa_sa + o), xb = lds128(pb_sa + o);
        acc[0] += a * xa.x;
        acc[1] += a * xa.y;
        acc[2] += a * xb.x;
        acc[3] += a * xb.y;
      }
    }
    if (WL4 == 0) {  // dynamic widths
      for (int g = 0; g < F.Wl4; g++) {
        const uint4 qq = F.ell4[(int64_t)g * F.ld + i];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t ent = ent_of(qq, k);
          const double a = lds64(amp_sa + amp_off(ent));
          const uint32_t o = (ent & HOP_TGT_MASK) * 16u;
          const double2 xa = lds128(pa_sa + o), xb = lds128(pb_sa + o);
          acc[0] += a * xa.x;
          acc[1] += a * xa.y;
          acc[2] += a * xb.x;
          acc[3] += a * xb.y;
        }
      }
      for (int g = 0; g < F.Wf4; g++) {
        const uint4 qq = F.ell4[(int64_t)(F.Wl4 + g) * F.ld + i];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (4 * g + k >= F.Wf) break;  // uniform
          const uint32_t ent = ent_of(qq, k);
          const double a = lds64(amp_sa + amp_off(ent));  // padding -> amplitude 0, target = row
          const uint32_t t = ent & HOP_TGT_MASK;
#pragma unroll
          for (int c = 0; c < 4; c++) acc[c] += a * vb[ko[c] + t];
        }
      }
    }
    // far gathers, own values, old Hv: L2 hits (prefetched above / staged by the sibling items)
    double own[4], hold[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      own[k] = WITH_DIAG ? vb[ko[k] + (uint32_t)i] : 0.0;
      hold[k] = ACCUM ? hb[ko[k] + (uint32_t)i] : 0.0;
    }
#pragma unroll
    for (int e = 0; e < NFAR; e++) {
      const uint32_t t = qf[e] & HOP_TGT_MASK;  // padding entries point at the row itself
      const double a = lds64(amp_sa + amp_off(qf[e]));
#pragma unroll
      for (int k = 0; k < 4; k++) acc[k] += a * vb[ko[k] + t];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double r = acc[k];
      if (WITH_DIAG) r += (eu + lds64(xc_sa + ((uint32_t)(k * nimp) + m) * 8u)) * own[k];
      r *= s_acc;
      if (ACCUM) r += s_old * hold[k];
      if (k < nc) hb[ko[k] + (uint32_t)i] = r;
    }
  }
  // the last work item also owns the pad rows [dim, ld): zeros (scaled old value when accumulating)
  if (blockIdx.x + 1 == gridDim.x) {
    for (int r = (int)F.dim + tid; r < (int)F.ld; r += FASTB_THREADS) {
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (k < nc) hb[ko[k] + (uint32_t)r] = ACCUM ? s_old * hb[ko[k] + (uint32_t)r] : 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Signed hop amplitudes of the open sector in the constant bank: c_amp[0] = the species applied
// along the fast index (up), c_amp[1] = along the slow index (dw).  Indexed per lane (LDC with a
// register index): the lanes of a warp sit in a run of consecutive rows and read at most a few
// distinct entries, and the look-up leaves the shared-memory / LSU pipe, which is what bounds
// the tiled passes (ncu l1tex__throughput 72-76 %).
// ---------------------------------------------------------------------------------------
__constant__ double c_amp[2][2048];

int hxv_upload_amps(Engine &E) {
  Sector &S = E.sec;
  const SpinSpace *sp[2] = {&S.up, &S.dw};
  for (int k = 0; k < 2; k++) {
    const size_t n = 2 * (size_t)sp[k]->nterms + 2;
    if (n > 2048) return set_error("too many one-body terms for the constant amplitude table");
    EDGPU_CUDA(cudaMemcpyToSymbolAsync(c_amp, sp[k]->amp2, sizeof(double) * n, sizeof(double) * 2048 * k,
                                       cudaMemcpyDeviceToDevice, E.stream));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// pass B in block mode, TMA-staged ("k_fastc", the default).  Same work decomposition as k_fastb:
// CTA = one work item (rows of one range prefix and one impurity configuration) x 4 columns,
// only the blocks its hops read from are staged.  Differences:
//   * the tile is staged by the TMA unit: ONE thread issues a cp.async.bulk per (input block,
//     column) -- up to 27 KB each -- completing on an mbarrier; no thread executes a load / store
//     for the staging, and the tile arrives while all threads prefetch their rows' hop entries;
//   * the tile is four column planes plane[k][row] (what a bulk copy of a column piece produces);
//     a hop gathers one 8-byte value per column;
//   * hop amplitudes come from the constant bank (c_amp) instead of shared memory.
// Bulk copies move 16-byte aligned pieces: the block plan (sector.cu) gives every input block a
// tile offset with the parity of its first row, so the copy may start one row early / end one row
// late into spare slots.
// shared layout: plane[4][tile_cap] doubles | xc[4][nimp] | mbarrier
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(mbar), "r"(parity)
                 : "memory");
}

template <int WL4, int NFAR, bool WITH_DIAG, bool ACCUM, bool NL>
__global__ void __launch_bounds__(FASTB_THREADS, 2)
k_fastc(const double *__restrict__ v, double *__restrict__ hv, int64_t ldv, int64_t ncol,
        int64_t col_offset, SpinView F, SpinView S, const double *__restrict__ xud, int nimp,
        double s_acc, double s_old, int tile_cap, NlDev nl, const double *__restrict__ old) {
  // `old` (ACCUM): where the old values s_old multiplies are read, laid out like hv; it may be hv
  // itself (in-place accumulate) or another vector (the Lanczos drivers write T = s_acc H x +
  // s_old X_{j-1} straight into the slot of the vector store, X_{j-1} staying intact in its own)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  double *plane = reinterpret_cast<double *>(smem_raw);
  double *xc = plane + 4 * (size_t)tile_cap;
  uint64_t *mbar = reinterpret_cast<uint64_t *>(xc + 4 * nimp);
  int32_t *nld = reinterpret_cast<int32_t *>(mbar + 1);  // NL: dw impurity hops of the 4 columns
  const BlockItem &it = F.items[blockIdx.x];
  const int out0 = it.out0, out1 = it.out1, nin = it.nin;

  const int64_t c0 = (int64_t)blockIdx.y * 4;
  const int nc = (int)min((int64_t)4, ncol - c0);  // live columns of this CTA
  const double *vb = v + c0 * ldv;
  double *hb = hv + c0 * ldv;
  const double *ob = old + c0 * ldv;
  const uint32_t ld32 = (uint32_t)ldv;
  const uint32_t ko[4] = {0u, nc > 1 ? ld32 : 0u, nc > 2 ? 2u * ld32 : 0u, nc > 3 ? 3u * ld32 : 0u};
  const uint32_t plane_sa = smem_u32(plane), mbar_sa = smem_u32(mbar);
  const uint32_t pstride = (uint32_t)tile_cap * 8u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_sa));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint32_t total = 0;
    for (int b = 0; b < nin; b++) {
      const int g0 = it.in0[b] & ~1, g1 = (it.in0[b] + it.in_len[b] + 1) & ~1;
      total += 4u * (uint32_t)(g1 - g0) * 8u;
    }
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_sa), "r"(total) : "memory");
    for (int b = 0; b < nin; b++) {
      const int g0 = it.in0[b] & ~1, g1 = (it.in0[b] + it.in_len[b] + 1) & ~1;
      const uint32_t bytes = (uint32_t)(g1 - g0) * 8u;
      const uint32_t dst = plane_sa + (uint32_t)(it.in_off[b] - (it.in0[b] & 1)) * 8u;
#pragma unroll
      for (int k = 0; k < 4; k++)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         dst + (uint32_t)k * pstride),
                     "l"(vb + ko[k] + (uint32_t)g0), "r"(bytes), "r"(mbar_sa)
                     : "memory");
    }
    if (nin == 0) {  // padding entries read slot 0
#pragma unroll
      for (int k = 0; k < 4; k++) plane[(size_t)k * tile_cap] = 0.0;
    }
  }
  // own values / old Hv are read from global memory late in each row iteration: pull their lines
  // into L2 now (see k_fastb)
  {
    const int l0 = out0 & ~15, nl = ((out1 + 15) & ~15) - l0;
    for (int p = tid * 16; p < nl; p += FASTB_THREADS * 16) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (WITH_DIAG) asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + ko[k] + (uint32_t)(l0 + p)));
        if (ACCUM) asm volatile("prefetch.global.L2 [%0];" ::"l"(ob + ko[k] + (uint32_t)(l0 + p)));
      }
    }
  }
  if (WITH_DIAG) {
    for (int t = tid; t < 4 * nimp; t += FASTB_THREADS) {
      const int k = t / nimp, m = t - k * nimp;
      const int64_t cg = c0 + (k < nc ? k : 0) + col_offset;
      xc[t] = S.eps[cg] + xud[(int)S.imp[cg] * nimp + m];
    }
  }
  const int no2 = NL ? nl.Norb * nl.Norb : 1;
  if (NL) {
    // nld[k][a * Norb + b] = target dw column of c^+_a c_b on column k | sign << 31, or -1
    for (int t = tid; t < 4 * no2; t += FASTB_THREADS) {
      const int k = t / no2, ab = t - k * no2;
      const int64_t cg = c0 + (k < nc ? k : 0) + col_offset;
      nld[t] = nl.dwT[(int64_t)ab * nl.ldd + cg];
    }
  }
  const uint32_t xc_sa = smem_u32(xc);
  const uint32_t *far32 = reinterpret_cast<const uint32_t *>(F.ell4 + (int64_t)F.Wl4 * F.ld);
  constexpr int NG = WL4 > 0 ? WL4 : 1;
  constexpr int NFE = NFAR > 0 ? NFAR : 1;
  uint4 nq[NG];
  uint32_t nqf[NFE];
  double neu = 0.0;
  uint32_t nm = 0;
  int i = out0 + tid;
  if (i < out1) {
#pragma unroll
    for (int g = 0; g < WL4; g++) nq[g] = F.ell4[(int64_t)g * F.ld + i];
#pragma unroll
    for (int e = 0; e < NFAR; e++) nqf[e] = far32[4 * (int64_t)i + e];
    if (WITH_DIAG) {
      neu = F.eps[i];
      nm = (uint32_t)F.imp[i];
    }
  }
  __syncthreads();          // xc, nld, the mbarrier initialisation
  mbar_wait(mbar_sa, 0);    // the tile has landed

  for (; i < out1; i += FASTB_THREADS) {
    uint4 q[NG];
#pragma unroll
    for (int g = 0; g < WL4; g++) q[g] = nq[g];
    uint32_t qf[NFE];
#pragma unroll
    for (int e = 0; e < NFAR; e++) qf[e] = nqf[e];
    const double eu = neu;
    const uint32_t m = nm;
    const int inext = i + FASTB_THREADS;
    if (inext < out1) {
#pragma unroll
      for (int g = 0; g < WL4; g++) nq[g] = F.ell4[(int64_t)g * F.ld + inext];
#pragma unroll
      for (int e = 0; e < NFAR; e++) nqf[e] = far32[4 * (int64_t)inext + e];
      if (WITH_DIAG) {
        neu = F.eps[inext];
        nm = (uint32_t)F.imp[inext];
      }
    }
    // far gathers, own values, old Hv: issued first (L2), consumed after the local hops
    double own[4], hold[4], xf[NFE][4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      own[k] = WITH_DIAG ? vb[ko[k] + (uint32_t)i] : 0.0;
      hold[k] = ACCUM ? ob[ko[k] + (uint32_t)i] : 0.0;
    }
#pragma unroll
    for (int e = 0; e < NFAR; e++) {
      const uint32_t t = qf[e] & HOP_TGT_MASK;  // padding entries point at the row itself
#pragma unroll
      for (int k = 0; k < 4; k++) xf[e][k] = vb[ko[k] + t];
    }
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int g = 0; g < WL4; g++) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t ent = ent_of(q[g], k);
        const double a = c_amp[0][(ent >> HOP_AMP_SHIFT) & HOP_AMP_MASK];
        const uint32_t o = plane_sa + (ent & HOP_TGT_MASK) * 8u;
#pragma unroll
        for (int c = 0; c < 4; c++) acc[c] += a * lds64(o + (uint32_t)c * pstride);
      }
    }
    if (WL4 == 0) {  // dynamic widths
      for (int g = 0; g < F.Wl4; g++) {
        const uint4 qq = F.ell4[(int64_t)g * F.ld + i];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t ent = ent_of(qq, k);
          const double a = c_amp[0][(ent >> HOP_AMP_SHIFT) & HOP_AMP_MASK];
          const uint32_t o = plane_sa + (ent & HOP_TGT_MASK) * 8u;
#pragma unroll
          for (int c = 0; c < 4; c++) acc[c] += a * lds64(o + (uint32_t)c * pstride);
        }
      }
      for (int g = 0; g < F.Wf4; g++) {
        const uint4 qq = F.ell4[(int64_t)(F.Wl4 + g) * F.ld + i];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (4 * g + k >= F.Wf) break;  // uniform
          const uint32_t ent = ent_of(qq, k);
          const double a = c_amp[0][(ent >> HOP_AMP_SHIFT) & HOP_AMP_MASK];  // padding -> 0, target = row
          const uint32_t t = ent & HOP_TGT_MASK;
#pragma unroll
          for (int c = 0; c < 4; c++) acc[c] += a * vb[ko[c] + t];
        }
      }
    }
#pragma unroll
    for (int e = 0; e < NFAR; e++) {
      const double a = c_amp[0][(qf[e] >> HOP_AMP_SHIFT) & HOP_AMP_MASK];
#pragma unroll
      for (int k = 0; k < 4; k++) acc[k] += a * xf[e][k];
    }
    if (NL) {
      // non-local S-E / P-H terms (direct/HxV_non_local.f90): up c^+_io c_jo together with dw
      // c^+_jo c_io (Jx) or dw c^+_io c_jo (Jp); all four operators act on impurity bits
      for (int io = 0; io < nl.Norb; io++)
        for (int jo = 0; jo < nl.Norb; jo++) {
          if (io == jo) continue;
          const int32_t uu = nl.upT[(int64_t)(io * nl.Norb + jo) * nl.ldu + i];
          if (uu == -1) continue;
          const double xj = nl.jx[io * nl.Norb + jo], yj = nl.jp[io * nl.Norb + jo];
          const double *vu = nl.vfull + (uu & 0x7FFFFFFF);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int32_t d1 = nld[k * no2 + jo * nl.Norb + io], d2 = nld[k * no2 + io * nl.Norb + jo];
            if (xj != 0.0 && d1 != -1)
              acc[k] += (((uu ^ d1) < 0) ? -xj : xj) * vu[(int64_t)(d1 & 0x7FFFFFFF) * ldv];
            if (yj != 0.0 && d2 != -1)
              acc[k] += (((uu ^ d2) < 0) ? -yj : yj) * vu[(int64_t)(d2 & 0x7FFFFFFF) * ldv];
          }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double r = acc[k];
      if (WITH_DIAG) r += (eu + lds64(xc_sa + ((uint32_t)(k * nimp) + m) * 8u)) * own[k];
      r *= s_acc;
      if (ACCUM) r += s_old * hold[k];
      if (k < nc) hb[ko[k] + (uint32_t)i] = r;
    }
  }
  // the last work item also owns the pad rows [dim, ld): zeros (scaled old value when accumulating)
  if (blockIdx.x + 1 == gridDim.x) {
    for (int r = (int)F.dim + tid; r < (int)F.ld; r += FASTB_THREADS) {
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (k < nc) hb[ko[k] + (uint32_t)r] = ACCUM ? s_old * ob[ko[k] + (uint32_t)r] : 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------
// pass A, "slow index" kernel.  CTA = SLOW_ROWS (16) consecutive fast rows x the slow range
// [s0, s1), staged as tile[j][16] with cp.async (LDGSTS) 16-byte copies.  A thread owns two rows
// of one column; the 8 threads of a column read one contiguous 128-byte segment per hop
// (= one conflict-free shared-memory wavefront) and share the column's hop entries.
//   hv[r, s0+j] (+)= sum_e amp_e * v[r, tgt_e]
// Entries: one list per column, local targets first, far ones flagged (read from global / L2).
// grid = (nranges, ceil(nrow / 16)), range index fastest (far gathers hit L2).
// ---------------------------------------------------------------------------------------
constexpr int SLOW_THREADS = 512;
constexpr int SLOW_R = SLOW_ROWS;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// deterministic CTA reduction of one double per thread -> dot_part[linear CTA index]
__device__ __forceinline__ void slow_block_sum(double x, double *__restrict__ dot_part, int tile0) {
  __shared__ double red[SLOW_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < SLOW_THREADS / 32; w++) t += red[w];
    dot_part[(size_t)(blockIdx.y + tile0) * gridDim.x + blockIdx.x] = t;
  }
}

// W4 = entry groups per column (0 = dynamic, slow generic loop), NF = number of leading slots
// that may hold far entries (the table is sorted far-first, sector.cu): their loads are issued
// first and consumed last, so that the L2 latency of the far gathers hides behind the local
// hops.  Entries of the next column are prefetched into registers.
// DOT: also accumulates <v, Hv_final> over the CTA's states into dot_part[CTA] (fused alpha of
// the Lanczos recurrence: v is in the tile, Hv_final in registers -> no extra memory traffic).
template <int W4, int NF, bool ACCUM, bool DOT>
__global__ void __launch_bounds__(SLOW_THREADS, 2)
k_slow(const double *__restrict__ v, double *__restrict__ hv, int64_t ldv, SpinView S,
       double s_acc, double *__restrict__ dot_part, const double *__restrict__ halo, int tile0) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tile = reinterpret_cast<double *>(smem_raw);
  const int tid = threadIdx.x;
  const int s0 = (int)S.range_start[blockIdx.x], s1 = (int)S.range_start[blockIdx.x + 1];
  const int len = s1 - s0;
  double *amp_s = tile + (size_t)len * SLOW_R;
  const int64_t i0 = (int64_t)(blockIdx.y + tile0) * SLOW_R;  // tile0: first row tile of a sub-grid
  constexpr int PARTS = SLOW_R / 2;              // 16-byte pieces (= threads) per column
  constexpr int CSTEP = SLOW_THREADS / PARTS;    // columns per sweep of the CTA
  const int rp2 = (tid % PARTS) * 2;             // first of the thread's two rows
  const int jl = tid / PARTS;
  const uint32_t ld32 = (uint32_t)ldv;

  {
    const double *src = v + (int64_t)(s0 + jl) * ldv + i0 + rp2;
    double *dst = tile + jl * SLOW_R + rp2;
    for (int j = jl; j < len; j += CSTEP) {
      cp_async16(dst, src);
      src += (int64_t)CSTEP * ldv;
      dst += CSTEP * SLOW_R;
    }
  }
  if (ACCUM) {
    // The Hv values this CTA will read-modify-write come from DRAM; their latency does not fit
    // the one-iteration software pipeline below.  Pull the CTA's whole Hv footprint (one 128-byte
    // line per column) into L2 now, while the tile loads: costs no registers.
    for (int j = tid; j < len; j += SLOW_THREADS)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(hv + (int64_t)(s0 + j) * ldv + i0));
  }
  for (int t = tid; t < 2 * S.nterms + 2; t += SLOW_THREADS) amp_s[t] = S.amp2[t];
  cp_async_wait_all();
  __syncthreads();

  const double *vrow = v + i0 + rp2;  // + t * ldv : far target column
  // sharded species: far targets >= qloc are halo slots (copies of remote columns, comm.cu)
  const uint32_t qloc = S.qloc;
  const double *hrow = qloc == 0xFFFFFFFFu ? vrow : halo + i0 + rp2 - (int64_t)qloc * ldv;
  // + t * 128 : local target column (wraps mod 2^32 before the add, exact after it)
  const uint32_t trow_sa = smem_u32(tile) + (uint32_t)rp2 * 8u - (uint32_t)s0 * (SLOW_R * 8u);
  const uint32_t amp_sa = smem_u32(amp_s);
  const uint4 *ell = S.ell4 + S.row_off + s0;
  constexpr int NG = W4 > 0 ? W4 : 1;

  if (W4 > 0) {
    // Software pipeline over the thread's columns j0, j0+CSTEP, ...:
    //   iteration j : (A) finish column j-CSTEP = consume its deferred loads (far gathers, Hv)
    //                 (B) issue the deferred loads of column j, prefetch the entries of j+CSTEP
    //                 (C) local hops of column j from shared memory
    // so that the L2 / DRAM latency of (B) is covered by a whole iteration.
    uint4 nq[NG];
    if (jl < len) {
#pragma unroll
      for (int g = 0; g < W4; g++) nq[g] = ell[(int64_t)g * S.ld + jl];
    }
    double2 xf[NF > 0 ? NF : 1];
    double dsum = 0.0;
    double2 hold = make_double2(0.0, 0.0), accp = make_double2(0.0, 0.0);
    uint4 q0p = make_uint4(0, 0, 0, 0);
    double *op = nullptr;
    for (int j = jl; j < len + CSTEP; j += CSTEP) {
      // (A)
      if (j > jl) {
#pragma unroll
        for (int e = 0; e < NF; e++) {
          const double a = c_amp[1][(ent_of(q0p, e & 3) >> HOP_AMP_SHIFT) & HOP_AMP_MASK];
          accp.x += a * xf[e].x;
          accp.y += a * xf[e].y;
        }
        accp.x *= s_acc;
        accp.y *= s_acc;
        if (ACCUM) {
          accp.x += hold.x;
          accp.y += hold.y;
        }
        *reinterpret_cast<double2 *>(op) = accp;
        if (DOT) {
          const double2 own = lds128(trow_sa + (uint32_t)(s0 + j - CSTEP) * (SLOW_R * 8u));
          dsum += own.x * accp.x + own.y * accp.y;
        }
      }
      if (j >= len) break;
      // (B)
      uint4 q[NG];
#pragma unroll
      for (int g = 0; g < W4; g++) q[g] = nq[g];
      double *o = hv + (int64_t)(s0 + j) * ldv + i0 + rp2;
      if (ACCUM) hold = *reinterpret_cast<const double2 *>(o);
#pragma unroll
      for (int e = 0; e < NF; e++) {
        const uint32_t ent = ent_of(q[e >> 2], e & 3);
        const uint32_t t = ent & HOP_TGT_MASK;
        if (ent & HOP_FAR)
          xf[e] = *reinterpret_cast<const double2 *>((t >= qloc ? hrow : vrow) + (size_t)(t * ld32));
        else
          xf[e] = lds128(trow_sa + t * (SLOW_R * 8u));
      }
      if (j + CSTEP < len) {
#pragma unroll
        for (int g = 0; g < W4; g++) nq[g] = ell[(int64_t)g * S.ld + j + CSTEP];
      }
      // (C)
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int e = NF; e < 4 * W4; e++) {
        const uint32_t ent = ent_of(q[e >> 2], e & 3);
        const double a = c_amp[1][(ent >> HOP_AMP_SHIFT) & HOP_AMP_MASK];
        const double2 x = lds128(trow_sa + (ent & HOP_TGT_MASK) * (SLOW_R * 8u));
        acc.x += a * x.x;
        acc.y += a * x.y;
      }
      accp = acc;
      q0p = q[0];
      op = o;
    }
    if (DOT) slow_block_sum(dsum, dot_part, tile0);
  } else {
    double dsum = 0.0;
    for (int j = jl; j < len; j += CSTEP) {
      const int c = s0 + j;
      double *o = hv + (int64_t)c * ldv + i0 + rp2;
      const double2 hold = ACCUM ? *reinterpret_cast<const double2 *>(o) : make_double2(0.0, 0.0);
      double2 acc = make_double2(0.0, 0.0);
      for (int g = 0; g < S.Wl4; g++) {
        const uint4 qq = ell[(int64_t)g * S.ld + j];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t ent = ent_of(qq, k);
          const double a = amp_s[(ent >> HOP_AMP_SHIFT) & HOP_AMP_MASK];
          const uint32_t t = ent & HOP_TGT_MASK;
          double2 x;
          if (ent & HOP_FAR)
            x = *reinterpret_cast<const double2 *>((t >= qloc ? hrow : vrow) + (size_t)t * ldv);
          else
            x = lds128(trow_sa + t * (SLOW_R * 8u));
          acc.x += a * x.x;
          acc.y += a * x.y;
        }
      }
      acc.x = s_acc * acc.x + hold.x;
      acc.y = s_acc * acc.y + hold.y;
      *reinterpret_cast<double2 *>(o) = acc;
      if (DOT) {
        const double2 own = lds128(trow_sa + (uint32_t)c * (SLOW_R * 8u));
        dsum += own.x * acc.x + own.y * acc.y;
      }
    }
    if (DOT) slow_block_sum(dsum, dot_part, tile0);
  }
}

// ---------------------------------------------------------------------------------------
// pass B', "transposed-tile" kernel for the FAST index (opt-in, EDGPU_UPT=1): the structure of
// k_slow applied to the up hops.  The up species is planned like a slow-role species (ranges of
// <= ~880 rows, one merged entry list per row, far entries first).  CTA = the up range [s0, s1) x
// FT_COLS (16) consecutive columns, staged TRANSPOSED as tile[row][16 columns] (128 B per row) by
// 8-byte cp.async: a warp copies 32 consecutive rows of one column (coalesced) into
// tile[r][j ^ (r & 15)] -- the XOR swizzle spreads the 32 stores over all banks.  8 threads own
// one row, each a pair of columns: a hop reads one conflict-free 128-byte segment (the pair of a
// thread sits at pair index p ^ ((r' & 15) >> 1), swapped when r' is odd) and the row's entries,
// amplitudes, eps and impurity bits are shared by 16 columns.
//   hv[s0+j, c] = s_acc * ( (eps_up + eps_dw + X[imp_dw][imp_up]) v[s0+j, c] + sum_e amp_e v[tgt_e, c] )
//                 (+ s_old * hv_old when ACCUM)
// grid = (nranges, ceil(ncol / 16)), range index fastest (far gathers hit L2).
// ---------------------------------------------------------------------------------------
constexpr int FT_COLS = 16;

__device__ __forceinline__ void cp_async8(uint32_t smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

// PAD = true: no swizzle, row pitch FT_PITCH = 18 doubles (144 B) instead: the 16-byte pairs of 32
// consecutive rows fall on 8 distinct 16-byte slots x 4 lanes (conflict-free), the tile address is
// one multiply-add and there is no swap.
constexpr int FT_PITCH = 18;

template <int W4, int NF, bool ACCUM, bool PAD>
__global__ void __launch_bounds__(SLOW_THREADS, 2)
k_fastT(const double *__restrict__ v, double *__restrict__ hv, int64_t ldv, int64_t ncol,
        int64_t col_offset, SpinView F, SpinView S, const double *__restrict__ xud, int nimp,
        double s_acc, double s_old, int lane_rows) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tile = reinterpret_cast<double *>(smem_raw);
  const int tid = threadIdx.x;
  const int s0 = (int)F.range_start[blockIdx.x], s1 = (int)F.range_start[blockIdx.x + 1];
  const int len = s1 - s0;
  constexpr int PITCH = PAD ? FT_PITCH : FT_COLS;
  double *amp_s = tile + (size_t)len * PITCH;
  double *xc = amp_s + 2 * F.nterms + 2;
  const int64_t c0 = (int64_t)blockIdx.y * FT_COLS;
  const int nc = (int)min((int64_t)FT_COLS, ncol - c0);
  constexpr int PARTS = FT_COLS / 2;            // threads per row
  constexpr int CSTEP = SLOW_THREADS / PARTS;   // rows per sweep of the CTA
  // thread <-> (row of the sweep, column pair).  lane_rows = 0: the 8 threads of a row are
  // neighbours (entries / amplitudes broadcast, but every global access of a warp -- far gathers,
  // Hv -- touches 8 columns = 8 lines: measured 1.97 ms at cfg 2, 1.8 L1 wavefronts per state).
  // lane_rows = 1: a warp = 32 consecutive rows of ONE column pair: far gathers and stores are
  // 256-byte runs, the tile reads stay conflict-free (4 lanes per 16-byte slot of the swizzle).
  const int p = lane_rows ? tid / CSTEP : tid % PARTS;   // the thread's column pair (2p, 2p+1)
  const int jl = lane_rows ? tid % CSTEP : tid / PARTS;
  const uint32_t tile_sa = smem_u32(tile);

  // stage: column j (missing columns of the last CTA alias the last live one, never stored)
  for (int j = 0; j < FT_COLS; j++) {
    const double *src = v + (c0 + (j < nc ? j : nc - 1)) * ldv + s0;
    for (int r = tid; r < len; r += SLOW_THREADS)
      cp_async8(tile_sa + (uint32_t)(r * PITCH + (PAD ? j : (j ^ (r & 15)))) * 8u, src + r);
  }
  if (ACCUM) {  // old Hv of the CTA's footprint -> L2 (one 128-byte line per 16 rows and column)
    const int nl = (len + 15) / 16;
    for (int t = tid; t < nl * nc; t += SLOW_THREADS) {
      const int j = t / nl, k = t - j * nl;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(hv + (c0 + j) * ldv + s0 + 16 * k));
    }
  }
  for (int t = tid; t < 2 * F.nterms + 2; t += SLOW_THREADS) amp_s[t] = F.amp2[t];
  // xc[j][m] = eps_dw(c) + X[imp_dw(c)][m]
  for (int t = tid; t < FT_COLS * nimp; t += SLOW_THREADS) {
    const int j = t / nimp, m = t - j * nimp;
    const int64_t cg = c0 + (j < nc ? j : nc - 1) + col_offset;
    xc[t] = S.eps[cg] + xud[(int)S.imp[cg] * nimp + m];
  }
  cp_async_wait_all();
  __syncthreads();

  const int ca = 2 * p < nc ? 2 * p : nc - 1, cb = 2 * p + 1 < nc ? 2 * p + 1 : nc - 1;
  const double *va = v + (c0 + ca) * ldv, *vb = v + (c0 + cb) * ldv;   // far gathers / own columns
  double *ha = hv + (c0 + ca) * ldv, *hb = hv + (c0 + cb) * ldv;
  const bool live_a = 2 * p < nc, live_b = 2 * p + 1 < nc;
  // + t * row pitch (wraps, exact after the add); PAD: the thread's pair offset folded in
  const uint32_t base_sa = tile_sa - (uint32_t)s0 * (PITCH * 8u) + (PAD ? (uint32_t)p * 16u : 0u);
  const uint32_t s0m = (uint32_t)s0 & 15u;
  const uint32_t amp_sa = smem_u32(amp_s);
  const uint32_t xc_sa = smem_u32(xc) + (uint32_t)(2 * p) * (uint32_t)nimp * 8u;
  const uint4 *ell = F.ell4 + s0;
  // pair of columns (2p, 2p+1) of global row t from the swizzled tile
  auto tile_pair = [&](uint32_t t) -> double2 {
    if (PAD) return lds128(base_sa + t * (PITCH * 8u));
    const uint32_t sw = (t - s0m) & 15u;
    double2 x = lds128(base_sa + t * (FT_COLS * 8u) + (((uint32_t)p ^ (sw >> 1)) << 4));
    if (sw & 1u) {
      const double y = x.x;
      x.x = x.y;
      x.y = y;
    }
    return x;
  };

  // Software pipeline over the thread's rows j0, j0+CSTEP, ... as in k_slow: iteration j consumes
  // the deferred loads (far gathers, old Hv) of row j-CSTEP, issues those of row j and does the
  // local hops of row j.
  constexpr int NG = W4;
  uint4 nq[NG];
  double neu = 0.0;
  uint32_t nm = 0;
  if (jl < len) {
#pragma unroll
    for (int g = 0; g < W4; g++) nq[g] = ell[(int64_t)g * F.ld + jl];
    neu = F.eps[s0 + jl];
    nm = (uint32_t)F.imp[s0 + jl];
  }
  double2 xf[NF > 0 ? NF : 1];
  double2 hold = make_double2(0.0, 0.0), accp = make_double2(0.0, 0.0);
  uint4 q0p = make_uint4(0, 0, 0, 0);
  int rowp = 0;
  for (int j = jl; j < len + CSTEP; j += CSTEP) {
    // (A)
    if (j > jl) {
#pragma unroll
      for (int e = 0; e < NF; e++) {
        const double a = lds64(amp_sa + amp_off(ent_of(q0p, e & 3)));
        accp.x += a * xf[e].x;
        accp.y += a * xf[e].y;
      }
      accp.x *= s_acc;
      accp.y *= s_acc;
      if (ACCUM) {
        accp.x += s_old * hold.x;
        accp.y += s_old * hold.y;
      }
      if (live_a) ha[rowp] = accp.x;
      if (live_b) hb[rowp] = accp.y;
    }
    if (j >= len) break;
    // (B)
    uint4 q[NG];
#pragma unroll
    for (int g = 0; g < W4; g++) q[g] = nq[g];
    const double eu = neu;
    const uint32_t mimp = nm;
    const int row = s0 + j;
    if (ACCUM) hold = make_double2(ha[row], hb[row]);
#pragma unroll
    for (int e = 0; e < NF; e++) {
      const uint32_t ent = ent_of(q[e >> 2], e & 3);
      const uint32_t t = ent & HOP_TGT_MASK;
      if (ent & HOP_FAR)
        xf[e] = make_double2(va[t], vb[t]);
      else
        xf[e] = tile_pair(t);
    }
    if (j + CSTEP < len) {
#pragma unroll
      for (int g = 0; g < W4; g++) nq[g] = ell[(int64_t)g * F.ld + j + CSTEP];
      neu = F.eps[row + CSTEP];
      nm = (uint32_t)F.imp[row + CSTEP];
    }
    // (C) diagonal + local hops
    const double2 own = tile_pair((uint32_t)row);
    double2 acc;
    acc.x = (eu + lds64(xc_sa + mimp * 8u)) * own.x;
    acc.y = (eu + lds64(xc_sa + ((uint32_t)nimp + mimp) * 8u)) * own.y;
#pragma unroll
    for (int e = NF; e < 4 * W4; e++) {
      const uint32_t ent = ent_of(q[e >> 2], e & 3);
      const double a = lds64(amp_sa + amp_off(ent));
      const double2 x = tile_pair(ent & HOP_TGT_MASK);
      acc.x += a * x.x;
      acc.y += a * x.y;
    }
    accp = acc;
    q0p = q[0];
    rowp = row;
  }
  // the last range also owns the pad rows [dim, ld): zeros (scaled old value when accumulating)
  if (blockIdx.x + 1 == gridDim.x) {
    const int npad = (int)(F.ld - F.dim);
    for (int t = tid; t < npad * nc; t += SLOW_THREADS) {
      const int j = t / npad, r = (int)F.dim + (t - j * npad);
      double *o = hv + (c0 + j) * ldv + r;
      *o = ACCUM ? s_old * *o : 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------
// non-local S-E / P-H terms (direct/HxV_non_local.f90): gather from anywhere in the vector.
//   S-E: nup(j)=1, ndw(i)=1, ndw(j)=0, nup(i)=0 :  dw: c^+_j c_i ; up: c^+_i c_j ; Jx(i,j)
//   P-H: nup(j)=1, ndw(j)=1, ndw(i)=0, nup(i)=0 :  dw: c^+_i c_j ; up: c^+_i c_j ; Jp(i,j)
// All four operators act on impurity bits, so signs depend on impurity bits only.
// vfull is the full (all-gathered) vector with leading dimension ldv.
// ---------------------------------------------------------------------------------------
// upT / dwT: impurity hop tables of the two species (sector.cu, k_imphop_fill).  The dw entries
// are uniform over a column (= block): columns on which no term can act exit at once.
// Sharded dw species (halo mode): colmap maps the global dw target to a local column (read from
// vfull = the rank's chunk) or to qloc + halo slot (read from halo); unsharded: colmap = nullptr.
__global__ void __launch_bounds__(128)
k_nonlocal(const double *__restrict__ vfull, double *__restrict__ hv, int64_t nrow, int64_t ldv,
           int64_t col_offset, const int32_t *__restrict__ upT, int64_t ldu,
           const int32_t *__restrict__ dwT, int64_t ldd, int Norb, const double *__restrict__ jx,
           const double *__restrict__ jp, double s_acc, const int32_t *__restrict__ colmap,
           const double *__restrict__ halo, int64_t qloc) {
  const int64_t c = blockIdx.y;
  const int64_t cg = c + col_offset;
  __shared__ int any_dw;
  if (threadIdx.x == 0) {
    int f = 0;
    for (int io = 0; io < Norb; io++)
      for (int jo = 0; jo < Norb; jo++) {
        if (io == jo) continue;
        if (jx[io * Norb + jo] != 0.0 && dwT[(int64_t)(jo * Norb + io) * ldd + cg] != -1) f = 1;
        if (jp[io * Norb + jo] != 0.0 && dwT[(int64_t)(io * Norb + jo) * ldd + cg] != -1) f = 1;
      }
    any_dw = f;
  }
  __syncthreads();
  if (!any_dw) return;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nrow) return;
  double acc = 0.0;
  for (int io = 0; io < Norb; io++)
    for (int jo = 0; jo < Norb; jo++) {
      if (io == jo) continue;
      const int32_t uu = upT[(int64_t)(io * Norb + jo) * ldu + i];  // up: c^+_io c_jo (both terms)
      if (uu == -1) continue;
      // S-E: dw c^+_jo c_io, Jx(io,jo) ; P-H: dw c^+_io c_jo, Jp(io,jo)
      const double x = jx[io * Norb + jo];
      const int32_t d1 = dwT[(int64_t)(jo * Norb + io) * ldd + cg];
      if (x != 0.0 && d1 != -1) {
        const double sg = ((uu ^ d1) < 0) ? -1.0 : 1.0;
        int64_t t = d1 & 0x7FFFFFFF;
        const double *src = vfull;
        if (colmap) {
          t = colmap[t];
          if (t >= qloc) {
            t -= qloc;
            src = halo;
          }
        }
        acc += x * sg * src[t * ldv + (uu & 0x7FFFFFFF)];
      }
      const double y = jp[io * Norb + jo];
      const int32_t d2 = dwT[(int64_t)(io * Norb + jo) * ldd + cg];
      if (y != 0.0 && d2 != -1) {
        const double sg = ((uu ^ d2) < 0) ? -1.0 : 1.0;
        int64_t t = d2 & 0x7FFFFFFF;
        const double *src = vfull;
        if (colmap) {
          t = colmap[t];
          if (t >= qloc) {
            t -= qloc;
            src = halo;
          }
        }
        acc += y * sg * src[t * ldv + (uu & 0x7FFFFFFF)];
      }
    }
  if (acc != 0.0) hv[c * ldv + i] += s_acc * acc;
}

// ---------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------
size_t fast_smem_bytes(int64_t max_range, int nterms, int nimp) {
  return sizeof(double) * (2 * (size_t)(max_range + 32) + 2 * (size_t)nimp + 2 * (size_t)nterms + 2);
}
size_t fastb_smem_bytes(int64_t max_tile, int nterms, int nimp) {
  return sizeof(double) * (4 * (size_t)std::max<int64_t>(max_tile, 1) + 4 * (size_t)nimp + 2 * (size_t)nterms + 2);
}
size_t slow_smem_bytes(int64_t max_range, int nterms) {
  return sizeof(double) * ((size_t)max_range * SLOW_R + 2 * (size_t)nterms + 2);
}

template <int WL4, int NFAR, bool WITH_DIAG, bool ACCUM>
static int launch_fast(Engine &E, const double *v, double *hv, int64_t ldv, int64_t ncol,
                       int64_t col_offset, const SpinView &F, const SpinView &S, int64_t max_range,
                       const double *xud, int nimp, double s_acc, double s_old) {
  const size_t smem = fast_smem_bytes(max_range, F.nterms, nimp);
  auto kern = k_fast<WL4, NFAR, WITH_DIAG, ACCUM>;
  EDGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)F.nranges, (unsigned)((ncol + 1) / 2));
  kern<<<grid, FAST_THREADS, smem, E.stream>>>(v, hv, ldv, ncol, col_offset, F, S, xud, nimp,
                                               s_acc, s_old);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

size_t fastc_smem_bytes(int64_t max_tile, int nimp) {
  return sizeof(double) * (4 * (size_t)std::max<int64_t>(max_tile, 2) + 4 * (size_t)nimp) + 16 +
         4 * EDGPU_MAXORB * EDGPU_MAXORB * sizeof(int32_t);
}

// EDGPU_FASTB=legacy: thread-staged k_fastb (pair planes, amplitudes in shared memory)
static bool fastb_legacy() {
  static const bool l = getenv("EDGPU_FASTB") && !strcmp(getenv("EDGPU_FASTB"), "legacy");
  return l;
}

// set by hxv_device_ex around the pass-B launch when the non-local terms are to be fused into it
static const NlDev *g_fuse_nl = nullptr;
static bool g_nl_fused = false;  // the last pass-B launch applied them
// source of the old values of an accumulating pass B when it is not the output vector itself
// (set by hxv_device_ex around the launch; only k_fastc reads old values from a second vector)
static const double *g_old_src = nullptr;

static bool fastc_serves(Engine &E, const double *amp2, const double *v, int64_t ldv) {
  return !fastb_legacy() && amp2 == E.sec.up.amp2 && (ldv & 1) == 0 && ((uintptr_t)v & 15) == 0;
}

template <int WL4, int NFAR, bool WITH_DIAG, bool ACCUM>
static int launch_fastc(Engine &E, const double *v, double *hv, int64_t ldv, int64_t ncol,
                        int64_t col_offset, const SpinView &F, const SpinView &S, int64_t max_tile,
                        const double *xud, int nimp, double s_acc, double s_old) {
  const size_t smem = fastc_smem_bytes(max_tile, nimp);
  dim3 grid((unsigned)F.nitems, (unsigned)((ncol + 3) / 4));
  const int cap = (int)std::max<int64_t>(max_tile, 2);
  if (WITH_DIAG && g_fuse_nl) {
    auto kern = k_fastc<WL4, NFAR, WITH_DIAG, ACCUM, WITH_DIAG>;  // NL only exists with the diagonal
    EDGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, FASTB_THREADS, smem, E.stream>>>(v, hv, ldv, ncol, col_offset, F, S, xud, nimp, s_acc, s_old,
                                                  cap, *g_fuse_nl, g_old_src ? g_old_src : hv);
    g_nl_fused = true;
  } else {
    auto kern = k_fastc<WL4, NFAR, WITH_DIAG, ACCUM, false>;
    EDGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, FASTB_THREADS, smem, E.stream>>>(v, hv, ldv, ncol, col_offset, F, S, xud, nimp, s_acc, s_old,
                                                  cap, NlDev(), g_old_src ? g_old_src : hv);
  }
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

template <int WL4, int NFAR, bool WITH_DIAG, bool ACCUM>
static int launch_fastb(Engine &E, const double *v, double *hv, int64_t ldv, int64_t ncol,
                        int64_t col_offset, const SpinView &F, const SpinView &S, int64_t max_tile,
                        const double *xud, int nimp, double s_acc, double s_old) {
  // the TMA-staged kernel serves the species whose amplitudes sit in c_amp[0] (the up species of
  // the open sector) with 16-byte aligned columns
  if (fastc_serves(E, F.amp2, v, ldv))
    return launch_fastc<WL4, NFAR, WITH_DIAG, ACCUM>(E, v, hv, ldv, ncol, col_offset, F, S, max_tile, xud, nimp,
                                                      s_acc, s_old);
  const size_t smem = fastb_smem_bytes(max_tile, F.nterms, nimp);
  auto kern = k_fastb<WL4, NFAR, WITH_DIAG, ACCUM>;
  EDGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)F.nitems, (unsigned)((ncol + 3) / 4));
  kern<<<grid, FASTB_THREADS, smem, E.stream>>>(v, hv, ldv, ncol, col_offset, F, S, xud, nimp, s_acc,
                                                s_old, (int)std::max<int64_t>(max_tile, 1));
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

size_t fastT_smem_bytes(int64_t max_range, int nterms, int nimp) {
  return sizeof(double) * ((size_t)max_range * FT_PITCH + 2 * (size_t)nterms + 2 + (size_t)FT_COLS * nimp);
}

// EDGPU_UPT: 1 = swizzled tile, 8 threads per row; 2 = swizzled tile, lanes along rows;
// 3 = padded tile (no swizzle), lanes along rows
static int fastT_mode() {
  static const int mode = getenv("EDGPU_UPT") ? atoi(getenv("EDGPU_UPT")) : 0;
  return mode;
}

template <int W4, int NF, bool ACCUM>
static int launch_fastT(Engine &E, const double *v, double *hv, int64_t ncol, int64_t col_offset,
                        const SpinSpace &Fs, const SpinView &F, const SpinView &S, const double *xud,
                        int nimp, double s_acc, double s_old) {
  const size_t smem = fastT_smem_bytes(Fs.max_range, F.nterms, nimp);
  const bool pad = fastT_mode() == 3;
  auto kern = pad ? k_fastT<W4, NF, ACCUM, true> : k_fastT<W4, NF, ACCUM, false>;
  EDGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)F.nranges, (unsigned)((ncol + FT_COLS - 1) / FT_COLS));
  const int lane_rows = fastT_mode() == 1 ? 0 : 1;
  kern<<<grid, SLOW_THREADS, smem, E.stream>>>(v, hv, F.ld, ncol, col_offset, F, S, xud, nimp, s_acc, s_old,
                                               lane_rows);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

// diag + up hops with the transposed-tile kernel; returns -1 when the table shape has no
// instantiation (the caller falls back to the generic gather kernel)
static int apply_fastT(Engine &E, bool accum, const double *v, double *hv, int64_t ncol, int64_t col_offset,
                       const SpinSpace &Fs, const SpinView &F, const SpinView &S, const double *xud, int nimp,
                       double s_acc, double s_old) {
  const int W4 = Fs.Wl4, NF = Fs.Wf;
  if (fastT_smem_bytes(Fs.max_range, F.nterms, nimp) > E.smem_optin) return -1;
#define EDGPU_FT(WW, FF)                                                                                  \
  (accum ? launch_fastT<WW, FF, true>(E, v, hv, ncol, col_offset, Fs, F, S, xud, nimp, s_acc, s_old)      \
         : launch_fastT<WW, FF, false>(E, v, hv, ncol, col_offset, Fs, F, S, xud, nimp, s_acc, s_old))
  if (W4 >= 1 && W4 <= 3 && NF <= 4 && NF <= 4 * W4) {
    switch (W4 * 8 + NF) {
      case 8 + 0: return EDGPU_FT(1, 0);
      case 8 + 1: return EDGPU_FT(1, 1);
      case 8 + 2: return EDGPU_FT(1, 2);
      case 8 + 3: return EDGPU_FT(1, 3);
      case 8 + 4: return EDGPU_FT(1, 4);
      case 16 + 0: return EDGPU_FT(2, 0);
      case 16 + 1: return EDGPU_FT(2, 1);
      case 16 + 2: return EDGPU_FT(2, 2);
      case 16 + 3: return EDGPU_FT(2, 3);
      case 16 + 4: return EDGPU_FT(2, 4);
      case 24 + 0: return EDGPU_FT(3, 0);
      case 24 + 1: return EDGPU_FT(3, 1);
      case 24 + 2: return EDGPU_FT(3, 2);
      case 24 + 3: return EDGPU_FT(3, 3);
      case 24 + 4: return EDGPU_FT(3, 4);
    }
  }
#undef EDGPU_FT
  return -1;
}

// Applies (diag +) the fast-index operator F to an [F.ld x ncol] block.
static int apply_fast(Engine &E, bool tiled, bool with_diag, bool accum, const double *v, double *hv,
                      int64_t ncol, int64_t col_offset, const SpinSpace &Fs, const SpinView &F,
                      const SpinView &S, const double *xud, int nimp, double s_acc = 1.0,
                      double s_old = 1.0) {
  if (ncol <= 0) return 0;
  if (tiled && Fs.role == ROLE_SLOW) {
    // the species was planned for the transposed-tile kernel (EDGPU_UPT=1, single rank)
    const int rc = with_diag ? apply_fastT(E, accum, v, hv, ncol, col_offset, Fs, F, S, xud, nimp, s_acc, s_old) : -1;
    if (rc >= 0) return rc;
    tiled = false;  // no instantiation for this table shape: generic gather kernel
  }
  if (!tiled) {
    dim3 grid((unsigned)((F.dim + 127) / 128), (unsigned)ncol);
    if (with_diag)
      k_generic<true, false><<<grid, 128, 0, E.stream>>>(v, hv, F.dim, F.ld, ncol, col_offset, F,
                                                          S, xud, nimp, (int)accum, s_acc, s_old, nullptr);
    else
      k_generic<false, false><<<grid, 128, 0, E.stream>>>(v, hv, F.dim, F.ld, ncol, col_offset,
                                                           F, S, xud, nimp, (int)accum, s_acc,
                                                           s_old, nullptr);
    EDGPU_COUNT_LAUNCH();
    EDGPU_CUDA(cudaGetLastError());
    return 0;
  }
#define EDGPU_FAST2(WW, FF, DD, AA)                                                               \
  (Fs.block_mode ? launch_fastb<WW, FF, DD, AA>(E, v, hv, F.ld, ncol, col_offset, F, S, Fs.max_tile, \
                                                xud, nimp, s_acc, s_old)                           \
                 : launch_fast<WW, FF, DD, AA>(E, v, hv, F.ld, ncol, col_offset, F, S, Fs.max_range, \
                                               xud, nimp, s_acc, s_old))
#define EDGPU_FAST(WW, FF)                                                                  \
  (with_diag ? (accum ? EDGPU_FAST2(WW, FF, true, true) : EDGPU_FAST2(WW, FF, true, false)) \
             : (accum ? EDGPU_FAST2(WW, FF, false, true) : EDGPU_FAST2(WW, FF, false, false)))
  if (F.Wl4 >= 1 && F.Wl4 <= 3 && F.Wf <= 2) {
    switch (F.Wl4 * 4 + F.Wf) {
      case 4 + 0: return EDGPU_FAST(1, 0);
      case 4 + 1: return EDGPU_FAST(1, 1);
      case 4 + 2: return EDGPU_FAST(1, 2);
      case 8 + 0: return EDGPU_FAST(2, 0);
      case 8 + 1: return EDGPU_FAST(2, 1);
      case 8 + 2: return EDGPU_FAST(2, 2);
      case 12 + 0: return EDGPU_FAST(3, 0);
      case 12 + 1: return EDGPU_FAST(3, 1);
      case 12 + 2: return EDGPU_FAST(3, 2);
    }
  }
  return EDGPU_FAST(0, 0);
#undef EDGPU_FAST2
#undef EDGPU_FAST
}

template <int W4, int NF, bool ACCUM, bool DOT>
static int launch_slow(Engine &E, const double *v, double *hv, const SpinSpace &Ss,
                       const SpinView &Fv, const SpinView &S, double s_acc, double *dot_part,
                       const double *halo, int tile0, int ntiles) {
  const size_t smem = slow_smem_bytes(Ss.max_range, S.nterms);
  dim3 grid((unsigned)S.nranges, (unsigned)ntiles);
  auto kern = k_slow<W4, NF, ACCUM, DOT>;
  EDGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, SLOW_THREADS, smem, E.stream>>>(v, hv, Fv.ld, S, s_acc, dot_part, halo, tile0);
  EDGPU_COUNT_LAUNCH();
  EDGPU_CUDA(cudaGetLastError());
  return 0;
}

// number of CTAs (= dot partials) of the slow pass
static int64_t slow_grid_size(const SpinView &Fv, const SpinView &S) {
  return (int64_t)S.nranges * ((Fv.ld + SLOW_R - 1) / SLOW_R);
}

static int apply_slow(Engine &E, bool accum, const double *v, double *hv, const SpinSpace &Ss,
                      const SpinView &Fv, const SpinView &S, double s_acc, double *dot_part,
                      const double *halo = nullptr, int tile0 = 0, int ntiles = -1) {
  if (ntiles < 0) ntiles = (int)((Fv.ld + SLOW_R - 1) / SLOW_R) - tile0;  // all row tiles from tile0 on
  if (S.nranges <= 0 || ntiles <= 0) return 0;  // a rank without columns (DimDw < nranks)
  // the element offset t * ld of a far gather is formed in 32 bits
  const uint64_t ncols = Ss.sharded ? (uint64_t)(Ss.shard_q + Ss.nhalo) : (uint64_t)Ss.dim;
  const bool small = ncols * (uint64_t)Fv.ld < (1ull << 32);
  const int W4 = Ss.Wl4, NF = Ss.Wf;  // Wf = largest number of far entries of a column
  const bool dot = dot_part != nullptr;
#define EDGPU_SLOW(WW, FF)                                                                       \
  (accum ? (dot ? launch_slow<WW, FF, true, true>(E, v, hv, Ss, Fv, S, s_acc, dot_part, halo, tile0, ntiles)    \
                : launch_slow<WW, FF, true, false>(E, v, hv, Ss, Fv, S, s_acc, dot_part, halo, tile0, ntiles))  \
         : (dot ? launch_slow<WW, FF, false, true>(E, v, hv, Ss, Fv, S, s_acc, dot_part, halo, tile0, ntiles)   \
                : launch_slow<WW, FF, false, false>(E, v, hv, Ss, Fv, S, s_acc, dot_part, halo, tile0, ntiles)))
  if (small && W4 >= 1 && W4 <= 3 && NF <= 4 && NF <= 4 * W4) {
    switch (W4 * 8 + NF) {
      case 8 + 0: return EDGPU_SLOW(1, 0);
      case 8 + 1: return EDGPU_SLOW(1, 1);
      case 8 + 2: return EDGPU_SLOW(1, 2);
      case 8 + 3: return EDGPU_SLOW(1, 3);
      case 8 + 4: return EDGPU_SLOW(1, 4);
      case 16 + 0: return EDGPU_SLOW(2, 0);
      case 16 + 1: return EDGPU_SLOW(2, 1);
      case 16 + 2: return EDGPU_SLOW(2, 2);
      case 16 + 3: return EDGPU_SLOW(2, 3);
      case 16 + 4: return EDGPU_SLOW(2, 4);
      case 24 + 0: return EDGPU_SLOW(3, 0);
      case 24 + 1: return EDGPU_SLOW(3, 1);
      case 24 + 2: return EDGPU_SLOW(3, 2);
      case 24 + 3: return EDGPU_SLOW(3, 3);
      case 24 + 4: return EDGPU_SLOW(3, 4);
    }
  }
  return EDGPU_SLOW(0, 0);
#undef EDGPU_SLOW
}

// Hv = s_acc * (H v) [+ s_old * Hv_old when accum].  When `dot_out` is given the device scalar
// *dot_out receives <v, Hv> of the LOCAL chunk, fused into the last pass when possible
// (single rank, tiled kernels), else through a separate dot kernel; the caller all-reduces.
int hxv_device_ex(Engine &E, const double *d_v, double *d_hv, bool accum, bool timed, double s_acc,
                  double s_old, double *dot_out, const double *d_old) {
  // d_old (accum only): the vector s_old multiplies when it is not d_hv itself.  Pass B of the tiled
  // NORMAL path (k_fastc) reads it in place of the old output; every other path copies it first.
  if (!accum || d_old == d_hv) d_old = nullptr;
  if (d_old) {
    const Sector &S0 = E.sec;
    const int var = S0.variant == 0 ? 2 : S0.variant;
    const bool native = !E.csr.open && S0.open && var == 2 && S0.up.block_mode && S0.up.role == ROLE_FAST &&
                        fastc_serves(E, S0.up.amp2, d_v, S0.up.ld) && (S0.slice_len() & 1) == 0;
    if (!native) {
      EDGPU_CUDA(cudaMemcpyAsync(d_hv, d_old, sizeof(double) * (size_t)E.veclen(), cudaMemcpyDeviceToDevice,
                                 E.stream));
      d_old = nullptr;
    }
  }
  if (E.csr.open) {  // stored-H sector: one SpMV kernel (csr.cu)
    if (timed) cudaEventRecord(E.ev[0], E.stream);
    EDGPU_TRY(csr_hxv_device(E, d_v, d_hv, accum, s_acc, s_old));
    if (dot_out) EDGPU_TRY(vec_dot_dev(E, d_v, d_hv, dot_out));
    if (timed) {
      cudaEventRecord(E.ev[1], E.stream);
      cudaEventSynchronize(E.ev[1]);
      cudaEventElapsedTime(&E.stage_ms[0], E.ev[0], E.ev[1]);
      E.stage_ms[1] = E.stage_ms[2] = E.stage_ms[3] = 0.f;
    }
    return 0;
  }
  Sector &S = E.sec;
  if (!S.open) return set_error("no sector open (build_Hv_sector_normal not called)");
  const SpinView U = view_of(S.up), D = view_of(S.dw);
  const int nimp = 1 << S.Norb;
  const int variant = S.variant == 0 ? 2 : S.variant;
  const bool tiled = (variant == 2);
  cudaStream_t st = E.stream;
  bool dot_done = false;
  // event sink: the profiling ring when armed, else the 4 scratch events (timed calls only)
  cudaEvent_t *evs = E.ev;
  if (E.prof_on && E.prof_n < E.prof_cap) {
    evs = &E.prof_ev[(size_t)4 * E.prof_n];
    E.prof_n++;
    timed = false;  // no host sync; edgpu_profile_end reads the ring
  } else if (!timed) {
    evs = nullptr;
  }
  // a10: DimPh > 1 phonon slices (each an electronic chunk: HxV_local / HxV_up / HxV_dw act on
  // every slice, direct/HxV_local.f90:2-4) and the terms applied after the electronic passes
  const int DimPh = S.DimPh;
  const int64_t slice = S.slice_len(), slice_full = U.ld * D.dim;
  const bool extras = (DimPh > 1) || (S.nsundry > 0);
  // nranks>1: terms that change the dw index gather from the columns of every rank, like
  // allgather_vector_MPI (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:355-360)
  // (halo mode: the non-local terms read their few remote dw columns from the halo instead)
  const bool need_full = E.nranks > 1 && ((S.nonlocal && !S.halo_mode) || S.nsundry > 0 ||
                                          (DimPh > 1 && S.eph_offdiag));
  if (need_full && !S.vfull) {
    EDGPU_CUDA(cudaMalloc(&S.vfull, sizeof(double) * (size_t)slice_full * (size_t)DimPh));
    S.gcounts.assign(E.nranks, 0);
    S.goffs.assign(E.nranks, 0);
    for (int p = 0; p < E.nranks; p++) {
      int64_t q, d0;
      block_split(D.dim, E.nranks, p, &q, &d0);
      S.gcounts[p] = q * U.ld;
      S.goffs[p] = d0 * U.ld;
    }
  }
#define EDGPU_MARK(k) \
  if (evs && (k == 0 ? iph == 0 : iph == DimPh - 1)) cudaEventRecord(evs[k], st)
  for (int iph = 0; iph < DimPh; iph++) {
    const double *v_s = d_v + iph * slice;
    double *hv_s = d_hv + iph * slice;
    g_old_src = d_old ? d_old + iph * slice : nullptr;  // read by the accumulating pass-B launch
    EDGPU_MARK(0);

    if (E.nranks == 1) {
      if (!tiled) {
        // one fused gather kernel: diagonal + up hops + dw hops
        dim3 grid((unsigned)((U.dim + 127) / 128), (unsigned)S.qdw);
        k_generic<true, true><<<grid, 128, 0, st>>>(v_s, hv_s, U.dim, U.ld, S.qdw, 0, U, D, S.xud,
                                                    nimp, (int)accum, s_acc, s_old, nullptr);
        EDGPU_COUNT_LAUNCH();
        EDGPU_CUDA(cudaGetLastError());
        EDGPU_MARK(1);
        EDGPU_MARK(2);
      } else {
        // Order of the two passes.  "BA" (default): pass B (diag + up hops) writes / accumulates first,
        // pass A (dw hops) accumulates and fuses the Lanczos dot product <v, Hv>.  EDGPU_ORDER=AB
        // (plain products only): pass A writes Hv without reading it and pass B accumulates; measured
        // at cfg 2 (profiles/r2_experiments.md) it only moves the read-modify-write: 0.67 + 0.92 ms
        // against 0.74 + 0.83 ms.
        static const int order_env = getenv("EDGPU_ORDER") ? (!strcmp(getenv("EDGPU_ORDER"), "AB") ? 2 : 1) : 0;
        const bool has_dw = (D.Wl4 + D.Wf4) > 0;
        // the non-local S-E / P-H terms ride in pass B (k_fastc<NL>): no third pass over Hv
        // (EDGPU_NL_FUSE=0: separate k_nonlocal kernel)
        static const bool nl_fuse_on = !(getenv("EDGPU_NL_FUSE") && getenv("EDGPU_NL_FUSE")[0] == '0');
        NlDev nl;
        nl.upT = S.up.imphop;
        nl.dwT = S.dw.imphop;
        nl.ldu = S.up.ld;
        nl.ldd = S.dw.ld;
        nl.jx = S.jx;
        nl.jp = S.jp;
        nl.vfull = v_s;
        nl.Norb = S.Norb;
        g_nl_fused = false;
        g_fuse_nl = (S.nonlocal && nl_fuse_on) ? &nl : nullptr;
        const bool a_first = has_dw && !accum && !dot_out && order_env == 2;
        if (a_first) {
          EDGPU_TRY(apply_slow(E, false, v_s, hv_s, S.dw, U, D, s_acc, nullptr));
          EDGPU_MARK(1);
          EDGPU_TRY(apply_fast(E, true, true, true, v_s, hv_s, S.qdw, 0, S.up, U, D, S.xud, nimp, s_acc, 1.0));
        } else {
          EDGPU_TRY(apply_fast(E, true, true, accum, v_s, hv_s, S.qdw, 0, S.up, U, D, S.xud, nimp, s_acc,
                               s_old));
          EDGPU_MARK(1);
          if (has_dw) {
            double *part = nullptr;
            const int64_t nblk = slow_grid_size(U, D);
            if (dot_out && (!S.nonlocal || g_nl_fused) && !extras) {
              EDGPU_TRY(ensure_partials(E, nblk));
              part = E.d_part;
            }
            EDGPU_TRY(apply_slow(E, true, v_s, hv_s, S.dw, U, D, s_acc, part));
            if (part) {
              EDGPU_TRY(final_sum(E, (int)nblk, dot_out));
              dot_done = true;
            }
          }
        }
        g_fuse_nl = nullptr;
        EDGPU_MARK(2);
      }
      if (S.nonlocal && !(tiled && g_nl_fused)) {
        dim3 grid((unsigned)((U.dim + 127) / 128), (unsigned)S.qdw);
        k_nonlocal<<<grid, 128, 0, st>>>(v_s, hv_s, U.dim, U.ld, 0, S.up.imphop, S.up.ld, S.dw.imphop,
                                         S.dw.ld, S.Norb, S.jx, S.jp, s_acc, nullptr, nullptr, 0);
        EDGPU_COUNT_LAUNCH();
        EDGPU_CUDA(cudaGetLastError());
      }
      if (!extras) EDGPU_MARK(3);
    } else {
      // dw-split over ranks (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:236-375):
      //   Hv  = (Hd + 1 (x) Hup) v                      local columns
      //   vt  = transpose(v)                            tiles exchanged over NVLink
      //   Hvt = Hdw vt                                  dw is now the fast index
      //   Hv += transpose(Hvt)
      // Phonon slices are processed one after the other like the reference's `do iph=1,DimPh`
      // (:322-337).
      EDGPU_CUDA(cudaEventRecord(E.ev_fork, st));
      if (S.halo_mode) {
        // Halo mode (comm.cu): the owners push the remote columns this chunk's dw hops read while
        // the rank-local pass B runs; pass A then runs on the chunk exactly as on one GPU.
        S.epoch++;
        const int K = S.nchunks;
        EDGPU_CUDA(cudaStreamWaitEvent(E.comm_stream, E.ev_fork, 0));
        for (int c = 0; c < K; c++) {  // the halo travels in K row chunks, each with its own flag
          EDGPU_TRY(comm_halo_push(E, c, v_s, E.comm_stream));
          EDGPU_TRY(comm_halo_signal(E, c, E.comm_stream));
        }
        EDGPU_CUDA(cudaEventRecord(E.ev_join, E.comm_stream));
        const double *halo = S.halo[S.epoch & 1];
        if (!tiled) {
          for (int c = 0; c < K; c++) EDGPU_TRY(comm_halo_wait(E, c, st));
          if (S.qdw > 0) {
            dim3 grid((unsigned)((U.dim + 127) / 128), (unsigned)S.qdw);
            k_generic<true, true><<<grid, 128, 0, st>>>(v_s, hv_s, U.dim, U.ld, S.qdw, S.d0, U, D, S.xud, nimp,
                                                        (int)accum, s_acc, s_old, halo);
            EDGPU_COUNT_LAUNCH();
            EDGPU_CUDA(cudaGetLastError());
          }
          EDGPU_MARK(1);
          EDGPU_MARK(2);
        } else {
          // diagnostic knob (profiling only): EDGPU_HALO_DIAG=1 skips pass B, =2 skips pass A, =3 both
          static const int diag = getenv("EDGPU_HALO_DIAG") ? atoi(getenv("EDGPU_HALO_DIAG")) : 0;
          if (!(diag & 1))
            EDGPU_TRY(apply_fast(E, true, true, accum, v_s, hv_s, S.qdw, S.d0, S.up, U, D, S.xud, nimp, s_acc,
                                 s_old));
          EDGPU_MARK(1);
          double *part = nullptr;
          const int64_t nblk = slow_grid_size(U, D);
          const bool has_dw = (D.Wl4 + D.Wf4) > 0 && !(diag & 2);
          if (has_dw && dot_out && !S.nonlocal && !extras && nblk > 0) {
            EDGPU_TRY(ensure_partials(E, nblk));
            part = E.d_part;
          }
          // pass A on the row tiles of chunk c as soon as chunk c of every owner has arrived
          for (int c = 0; c < K; c++) {
            int64_t r0, r1;
            comm_halo_rows(E, c, &r0, &r1);
            EDGPU_TRY(comm_halo_wait(E, c, st));
            if (has_dw)
              EDGPU_TRY(apply_slow(E, true, v_s, hv_s, S.dw, U, D, s_acc, part, halo, (int)(r0 / SLOW_R),
                                   (int)((r1 - r0) / SLOW_R)));
          }
          if (part) {
            EDGPU_TRY(final_sum(E, (int)nblk, dot_out));
            dot_done = true;
          }
          EDGPU_MARK(2);
        }
        EDGPU_CUDA(cudaStreamWaitEvent(st, E.ev_join, 0));
      } else if (S.p2p) {
        // Chunk pipeline over peer memory (comm.cu): three streams per rank,
        //   comm_stream : push(0) sig push(1) sig ...
        //   dw_stream   : wait(push 0) Hdw(0) return(0) sig  wait(push 1) Hdw(1) ...
        //   main        : local pass (diag + up hops), then wait(ret c) + add(c) for every chunk
        // so that the NVLink stores of both directions, the Hdw pass and the local pass overlap.
        S.epoch++;
        const int K = S.nchunks;
        EDGPU_CUDA(cudaStreamWaitEvent(E.comm_stream, E.ev_fork, 0));
        EDGPU_CUDA(cudaStreamWaitEvent(E.dw_stream, E.ev_fork, 0));
        for (int c = 0; c < K; c++) {
          EDGPU_TRY(comm_pipe_push(E, c, v_s, E.comm_stream));
          EDGPU_TRY(comm_pipe_signal(E, PIPE_PUSH, c, E.comm_stream));
        }
        EDGPU_CUDA(cudaEventRecord(E.ev_join, E.comm_stream));
        EDGPU_TRY(apply_fast(E, tiled, true, accum, v_s, hv_s, S.qdw, S.d0, S.up, U, D, S.xud, nimp, s_acc,
                             s_old));
        EDGPU_MARK(1);
        {
          cudaStream_t keep = E.stream;
          E.stream = E.dw_stream;  // apply_fast launches on E.stream
          int rc = 0;
          for (int c = 0; c < K && !rc; c++) {
            int64_t col0, ncols;
            comm_pipe_chunk_cols(E, E.rank, c, &col0, &ncols);
            rc = comm_pipe_wait(E, PIPE_PUSH, c, E.dw_stream);
            if (!rc && ncols > 0)
              rc = apply_fast(E, tiled, false, false, S.vt + col0 * D.ld, S.hvt + col0 * D.ld, ncols,
                              S.u0 + col0, S.dw, D, U, S.xud, nimp, s_acc, 1.0);
            if (!rc) rc = comm_pipe_return(E, c, E.dw_stream);
            if (!rc) rc = comm_pipe_signal(E, PIPE_RET, c, E.dw_stream);
          }
          E.stream = keep;
          if (rc) return rc;
        }
        EDGPU_CUDA(cudaEventRecord(E.ev_join2, E.dw_stream));
        for (int c = 0; c < K; c++) {
          EDGPU_TRY(comm_pipe_wait(E, PIPE_RET, c, st));
          if (c == K - 1) EDGPU_MARK(2);
          EDGPU_TRY(comm_pipe_add(E, c, hv_s, st));
        }
        EDGPU_CUDA(cudaStreamWaitEvent(st, E.ev_join, 0));
        EDGPU_CUDA(cudaStreamWaitEvent(st, E.ev_join2, 0));
      } else {
        // NCCL fallback (EDGPU_NO_P2P=1 or unmappable peers): grouped send/recv tile transposes.
        // The first one only reads v: it runs on the communication stream concurrently with the
        // rank-local pass on the main stream.
        EDGPU_CUDA(cudaStreamWaitEvent(E.comm_stream, E.ev_fork, 0));
        {
          cudaStream_t keep = E.stream;
          E.stream = E.comm_stream;  // the comm_* helpers and apply_fast launch on E.stream
          int rc = comm_transpose(E, v_s, U.dim, U.ld, S.qdw, S.vt, D.dim, D.ld, S.qup, false);
          if (!rc)
            rc = apply_fast(E, tiled, false, false, S.vt, S.hvt, S.qup, S.u0, S.dw, D, U, S.xud, nimp, s_acc,
                            1.0);
          E.stream = keep;
          if (rc) return rc;
        }
        EDGPU_CUDA(cudaEventRecord(E.ev_join, E.comm_stream));
        EDGPU_TRY(apply_fast(E, tiled, true, accum, v_s, hv_s, S.qdw, S.d0, S.up, U, D, S.xud, nimp, s_acc,
                             s_old));
        EDGPU_MARK(1);
        EDGPU_CUDA(cudaStreamWaitEvent(st, E.ev_join, 0));
        EDGPU_MARK(2);
        EDGPU_TRY(comm_transpose(E, S.hvt, D.dim, D.ld, S.qup, hv_s, U.dim, U.ld, S.qdw, true));
      }
      if (need_full) EDGPU_TRY(comm_allgatherv(E, v_s, S.vfull + iph * slice_full, S.gcounts, S.goffs));
      if (S.nonlocal && S.qdw > 0) {
        dim3 grid((unsigned)((U.dim + 127) / 128), (unsigned)S.qdw);
        if (S.halo_mode)
          k_nonlocal<<<grid, 128, 0, st>>>(v_s, hv_s, U.dim, U.ld, S.d0, S.up.imphop, S.up.ld, S.dw.imphop,
                                           S.dw.ld, S.Norb, S.jx, S.jp, s_acc, S.dw.d_colmap,
                                           S.halo[S.epoch & 1], S.qdw);
        else
          k_nonlocal<<<grid, 128, 0, st>>>(S.vfull + iph * slice_full, hv_s, U.dim, U.ld, S.d0, S.up.imphop,
                                           S.up.ld, S.dw.imphop, S.dw.ld, S.Norb, S.jx, S.jp, s_acc, nullptr,
                                           nullptr, 0);
        EDGPU_COUNT_LAUNCH();
        EDGPU_CUDA(cudaGetLastError());
      }
      if (!extras) EDGPU_MARK(3);
    }
  }  // phonon slices
  g_old_src = nullptr;
  if (extras) {
    EDGPU_TRY(extra_hxv(E, d_v, need_full ? S.vfull : d_v, d_hv, s_acc));
    const int iph = DimPh - 1;
    EDGPU_MARK(3);
  }
#undef EDGPU_MARK
  if (dot_out && !dot_done) EDGPU_TRY(vec_dot_dev(E, d_v, d_hv, dot_out));
  if (timed) {
    cudaEventSynchronize(E.ev[3]);
    cudaEventElapsedTime(&E.stage_ms[0], E.ev[0], E.ev[1]);
    cudaEventElapsedTime(&E.stage_ms[1], E.ev[1], E.ev[2]);
    cudaEventElapsedTime(&E.stage_ms[2], E.ev[2], E.ev[3]);
    E.stage_ms[3] = 0.f;
  }
  return 0;
}

int hxv_device(Engine &E, const double *d_v, double *d_hv, bool accum, bool timed) {
  return hxv_device_ex(E, d_v, d_hv, accum, timed, 1.0, 1.0, nullptr, nullptr);
}

}  // namespace edgpu
