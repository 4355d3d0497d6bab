// Microbenchmark: tcgen05.mma.cta_group::2 (M = 256 over a CTA pair) issue / completion rate vs N.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t lo, uint32_t hi) { uint64_t d; asm("mov.b64 %0, {%1,%2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; }
constexpr uint32_t HI128 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t HI64 = (512u >> 4) | (1u << 14) | (4u << 29);

template <int N, bool TS, int NMMA, int ALT>
__global__ void __launch_bounds__(128, 1) k(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("fence.mbarrier_init.release.cluster;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;"); asm volatile("barrier.cluster.wait.acquire.aligned;");
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm0 = slot; const uint32_t tm = tm0; (void)tm;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (16u << 24);
    const uint32_t a_lo = ((base >> 4) & 0x3FFF) | 0x10000u;
    const uint32_t b_lo = (((base + 65536) >> 4) & 0x3FFF) | 0x10000u;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < NMMA; ++i) {
      const uint64_t db = mkdesc(b_lo + (uint32_t)((i & 15) * 2), HI64);
      const uint32_t tm = tm0 + (ALT ? (uint32_t)(((i / ALT) & 1) * 256) : 0u);   // ALT: switch accumulator (and TMEM A) every ALT MMAs
      if (TS) asm volatile("{.reg .pred p; setp.eq.u32 p,1,1; tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(tm),
                           "r"(tm + 128u + (uint32_t)((i & 7) * 8)), "l"(db), "r"(idesc));
      else { const uint64_t da = mkdesc(a_lo + (uint32_t)((i & 3) * 2), HI128);
             asm volatile("{.reg .pred p; setp.eq.u32 p,1,1; tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(tm), "l"(da), "l"(db), "r"(idesc)); }
    }
    long long t1 = clock64();
    asm volatile("{.reg .b16 m; mov.b16 m, 1; tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;}" ::"r"(smem_u32(&bar)));
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar)));
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;"); asm volatile("barrier.cluster.wait.acquire.aligned;");
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

template <int N, bool TS, int ALT = 0>
void run(const char* name) {
  constexpr int NMMA = 64;
  long long* d; cudaMalloc(&d, 16);
  auto kern = k<N, TS, NMMA, ALT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, kern, d);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-30s issue %6lld cyc (%5.1f/mma)  complete %6lld cyc (%5.1f/mma)  formula %d  %s\n", name, h[0], h[0] / 64.0, h[1],
         h[1] / 64.0, 256 * N / 512, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<256, false>("2SM N=256 SS");
  run<128, false>("2SM N=128 SS");
  run<64, false>("2SM N=64  SS");
  run<256, true>("2SM N=256 TS");
  run<128, true>("2SM N=128 TS");
  run<64, true>("2SM N=64  TS");
  run<32, true>("2SM N=32  TS");
  run<128, true, 4>("2SM N=128 TS alt-D every 4");
  run<128, true, 1>("2SM N=128 TS alt-D every 1");
  run<128, false, 4>("2SM N=128 SS alt-D every 4");
  run<128, true, 16>("2SM N=128 TS alt-D every 16");
  return 0;
}
