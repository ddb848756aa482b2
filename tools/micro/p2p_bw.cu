// Peer-store bandwidth microbenchmark (diagnostic, not part of the product): GPU0 copies columns of
// a local buffer into GPU1's memory the way the halo push does, with different store flavours.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_bw p2p_bw.cu && ./p2p_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int THREADS = 512, UNROLL = 8;

// flavour 0: ld/st .v2.f64 (current halo push)
__global__ void __launch_bounds__(THREADS) k_v2(const double2 *__restrict__ s, double2 *__restrict__ d, int64_t ncol, int n2, int64_t sstride, int64_t dstride) {
  const int piece = THREADS * UNROLL;
  const int npiece = (n2 + piece - 1) / piece;
  for (int64_t w = blockIdx.x; w < ncol * npiece; w += gridDim.x) {
    const int64_t c = w / npiece;
    const int p = (int)(w - c * npiece);
    const double2 *sp = s + c * sstride;
    double2 *dp = d + c * dstride;
    const int i0 = p * piece + threadIdx.x;
    double2 r[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; k++) if (i0 + k * THREADS < n2) r[k] = sp[i0 + k * THREADS];
#pragma unroll
    for (int k = 0; k < UNROLL; k++) if (i0 + k * THREADS < n2) dp[i0 + k * THREADS] = r[k];
  }
}

// flavour 1: 256-bit ld/st (.v4.f64, sm_100+)
__global__ void __launch_bounds__(THREADS) k_v4(const double *__restrict__ s, double *__restrict__ d, int64_t ncol, int n4, int64_t sstride, int64_t dstride) {
  constexpr int U = UNROLL / 2;
  const int piece = THREADS * U;
  const int npiece = (n4 + piece - 1) / piece;
  for (int64_t w = blockIdx.x; w < ncol * npiece; w += gridDim.x) {
    const int64_t c = w / npiece;
    const int p = (int)(w - c * npiece);
    const double *sp = s + c * sstride * 2;
    double *dp = d + c * dstride * 2;
    const int i0 = p * piece + threadIdx.x;
    double r[U][4];
#pragma unroll
    for (int k = 0; k < U; k++) if (i0 + k * THREADS < n4)
      asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r[k][0]), "=d"(r[k][1]), "=d"(r[k][2]), "=d"(r[k][3]) : "l"(sp + 4 * (int64_t)(i0 + k * THREADS)));
#pragma unroll
    for (int k = 0; k < U; k++) if (i0 + k * THREADS < n4)
      asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(dp + 4 * (int64_t)(i0 + k * THREADS)), "d"(r[k][0]), "d"(r[k][1]), "d"(r[k][2]), "d"(r[k][3]) : "memory");
  }
}

// flavour 2: TMA bulk copies global -> shared -> peer global, STAGES x CHUNK ring, one issuing thread
constexpr int CHUNK = 16384, STAGES = 4;
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32) k_tma(const char *__restrict__ s, char *__restrict__ d, int64_t ncol, int64_t colbytes, int64_t sstride, int64_t dstride) {
  extern __shared__ __align__(128) char ring[];
  __shared__ __align__(8) uint64_t full[STAGES];
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x != 0) return;
  const int64_t npiece = (colbytes + CHUNK - 1) / CHUNK;
  const int64_t nwork = ncol * npiece;
  // work items of this CTA: w = blockIdx.x + i * gridDim.x
  int64_t issued = 0, done = 0;
  const int64_t mine = nwork > blockIdx.x ? (nwork - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  uint32_t phase[STAGES] = {0, 0, 0, 0};
  auto item = [&](int64_t i, const char **sp, char **dp, uint32_t *bytes) {
    const int64_t w = blockIdx.x + i * gridDim.x;
    const int64_t c = w / npiece, p = w - c * npiece;
    *sp = s + c * sstride + p * CHUNK;
    *dp = d + c * dstride + p * CHUNK;
    const int64_t rem = colbytes - p * CHUNK;
    *bytes = (uint32_t)(rem < CHUNK ? rem : CHUNK);
  };
  while (done < mine) {
    // keep STAGES loads in flight
    while (issued < mine && issued < done + STAGES) {
      const int st = (int)(issued % STAGES);
      if (issued >= STAGES)  // the stage's previous store must have finished READING shared memory
        asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(STAGES - 1) : "memory");
      const char *sp; char *dp; uint32_t bytes;
      item(issued, &sp, &dp, &bytes);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&full[st])), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(s32(ring + st * CHUNK)), "l"(sp), "r"(bytes), "r"(s32(&full[st])) : "memory");
      issued++;
    }
    const int st = (int)(done % STAGES);
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(s32(&full[st])), "r"(phase[st]) : "memory");
    phase[st] ^= 1;
    const char *sp; char *dp; uint32_t bytes;
    item(done, &sp, &dp, &bytes);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dp), "r"(s32(ring + st * CHUNK)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    done++;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  const int64_t ld = 12880, ncol = 3300;  // ~ the cfg2 halo of one rank at N=2 (340 MB)
  const int64_t colbytes = ld * 8, total = ncol * colbytes;
  double *src, *dst, *dstl;
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&dst, total));
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&src, 2 * total));  // every second column is sent (like the imp-bit pattern)
  CK(cudaMalloc(&dstl, total));
  CK(cudaMemset(src, 1, 2 * total));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, CHUNK * STAGES));
  auto run = [&](const char *name, int flavour, int ctas, double *d) {
    float best = 1e9;
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      if (flavour == 0) k_v2<<<ctas, THREADS>>>((const double2 *)src, (double2 *)d, ncol, (int)(ld / 2), ld, ld / 2);
      else if (flavour == 1) k_v4<<<ctas, THREADS>>>(src, d, ncol, (int)(ld / 4), ld / 2 * 2, ld / 4 * 2);
      else if (flavour == 2) k_tma<<<ctas, 32, CHUNK * STAGES>>>((const char *)src, (char *)d, ncol, colbytes, 2 * colbytes, colbytes);
      else cudaMemcpyPeerAsync(d, d == dst ? 1 : 0, src, 0, total, 0);
      cudaEventRecord(e1);
      if (cudaEventSynchronize(e1) != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("%-28s ctas=%4d  %7.3f ms  %7.1f GB/s  (%s)\n", name, ctas, best, total / best / 1e6, d == dst ? "peer" : "local");
  };
  run("cudaMemcpyPeerAsync", 3, 0, dst);
  for (int ctas : {20, 40, 80, 148, 296}) run("st.v2.f64 x8", 0, ctas, dst);
  for (int ctas : {20, 40, 80, 148, 296}) run("st.v4.f64 x4 (256-bit)", 1, ctas, dst);
  for (int ctas : {16, 32, 64, 148, 296, 592}) run("TMA bulk 4x16KB ring", 2, ctas, dst);
  for (int ctas : {40, 148}) run("st.v2.f64 x8", 0, ctas, dstl);
  for (int ctas : {32, 148}) run("TMA bulk 4x16KB ring", 2, ctas, dstl);
  return 0;
}
