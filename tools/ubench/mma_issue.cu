// Microbenchmark: how fast can ONE thread issue tcgen05.mma?  N = 32 MMAs (16 cycles of tensor work each) so that the
// issue path, not the tensor pipe, is what is measured.  Variants:
//   A  if (threadIdx.x == 0)  one asm statement per MMA               (divergent single thread)
//   B  whole warp + elect.sync, one asm statement per MMA              (warp-uniform control flow)
//   C  whole warp + elect.sync, ONE asm block with 8 MMAs, all descriptor words precomputed before the block
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t lo, uint32_t hi) { uint64_t d; asm("mov.b64 %0, {%1,%2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{.reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p;}" : "=r"(pred));
  return pred != 0;
}
constexpr uint32_t HI64 = (512u >> 4) | (1u << 14) | (4u << 29);

template <int N, int MODE, int NMMA>
__global__ void __launch_bounds__(128, 1) k(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  const uint32_t b_lo = (((base + 65536) >> 4) & 0x3FFF) | 0x10000u;
  if (threadIdx.x < 32) {
    long long t0 = clock64(), t1 = 0;
    if (MODE == 3) {
      // sustained: reps x (64 MMAs, commit, wait) on every CTA of the grid
      uint32_t par = 0;
      for (int r = 0; r < reps; ++r) {
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < NMMA; ++i)
            asm volatile("{.reg .pred p; setp.eq.u32 p,1,1; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(tm),
                         "r"(tm + 256u + (uint32_t)((i & 15) * 8)), "l"(mkdesc(b_lo + (uint32_t)((i & 15) * 2), HI64)), "r"(idesc));
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        }
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(par));
        par ^= 1;
      }
      long long t2 = clock64();
      if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t2 - t0; out[1] = t2 - t0; }
    } else if (MODE == 0) {
      if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i)
          asm volatile("{.reg .pred p; setp.eq.u32 p,1,1; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(tm),
                       "r"(tm + 256u + (uint32_t)((i & 15) * 8)), "l"(mkdesc(b_lo + (uint32_t)((i & 15) * 2), HI64)), "r"(idesc));
      }
    } else if (MODE == 1) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i)
          asm volatile("{.reg .pred p; setp.eq.u32 p,1,1; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(tm),
                       "r"(tm + 256u + (uint32_t)((i & 15) * 8)), "l"(mkdesc(b_lo + (uint32_t)((i & 15) * 2), HI64)), "r"(idesc));
      }
    } else {
      if (elect_one()) {
#pragma unroll
        for (int g = 0; g < NMMA / 8; ++g) {
          const uint32_t a0 = tm + 256u + (uint32_t)((g & 1) * 64), b0 = b_lo + (uint32_t)((g & 1) * 16);
          asm volatile(
              "{.reg .pred p; setp.eq.u32 p,1,1;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %10, %1, p;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%3], %11, %1, p;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%4], %12, %1, p;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%5], %13, %1, p;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%6], %14, %1, p;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%7], %15, %1, p;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%8], %16, %1, p;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%9], %17, %1, p;}\n" ::"r"(tm), "r"(idesc), "r"(a0), "r"(a0 + 8), "r"(a0 + 16),
              "r"(a0 + 24), "r"(a0 + 32), "r"(a0 + 40), "r"(a0 + 48), "r"(a0 + 56), "l"(mkdesc(b0, HI64)), "l"(mkdesc(b0 + 2, HI64)),
              "l"(mkdesc(b0 + 4, HI64)), "l"(mkdesc(b0 + 6, HI64)), "l"(mkdesc(b0 + 8, HI64)), "l"(mkdesc(b0 + 10, HI64)),
              "l"(mkdesc(b0 + 12, HI64)), "l"(mkdesc(b0 + 14, HI64)));
        }
      }
    }
    __syncwarp();
    t1 = clock64();
    if (MODE != 3) {
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar)));
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

// Commit cost: ITER x (8 MMAs N=128, [dummy work], tcgen05.commit, [dummy work]).  WHERE = 0 none, 1 dummy before the commit, 2 after.
template <int WHERE, int NCOMMIT>
__global__ void __launch_bounds__(128, 1) kc(long long* out, int iters, int dummy) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar[2];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(iters));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(iters));
  }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | (8u << 24);
  const uint32_t b_lo = (((base + 65536) >> 4) & 0x3FFF) | 0x10000u;
  if (threadIdx.x < 32) {
    uint32_t acc = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("{.reg .pred p; setp.eq.u32 p,1,1; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(tm),
                       "r"(tm + 256u + (uint32_t)(i * 8)), "l"(mkdesc(b_lo + (uint32_t)(i * 2), HI64)), "r"(idesc));
      }
      __syncwarp();
      if (WHERE == 1) for (int d = 0; d < dummy; ++d) acc = acc * 1664525u + 1013904223u;
      if (elect_one()) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])));
        if (NCOMMIT == 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])));
      }
      __syncwarp();
      if (WHERE == 2) for (int d = 0; d < dummy; ++d) acc = acc * 1664525u + 1013904223u;
    }
    long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar[0])));
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = acc; }
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
template <int WHERE, int NCOMMIT>
void runc(const char* name, int dummy) {
  const int iters = 200;
  long long* d; cudaMalloc(&d, 32);
  auto kern = kc<WHERE, NCOMMIT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) kern<<<1, 128, 200 * 1024>>>(d, iters, dummy);
  long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-56s thread %7.1f cyc / iter, complete %7.1f cyc / iter (8 MMAs = 534 at the floor) %s\n", name, h[0] / (double)iters,
         h[1] / (double)iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}


// Contention: warp 17 issues ITER x (8 MMAs N=128 TS + commit) while 16 other warps (4 per SM sub-partition, as in the
// field kernel) do nothing (LOAD=0), ALU work (1), or tcgen05.ld x32 + tcgen05.st x16 epilogue-like TMEM traffic (2).
template <int LOAD>
__global__ void __launch_bounds__(576, 1) kx(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  __shared__ volatile int done;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(iters)); done = 0; }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 576) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  if (warp == 17) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | (8u << 24);
    const uint32_t b_lo = (((base + 65536) >> 4) & 0x3FFF) | 0x10000u;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("{.reg .pred p; setp.eq.u32 p,1,1; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(tm),
                       "r"(tm + 256u + (uint32_t)(i * 8)), "l"(mkdesc(b_lo + (uint32_t)(i * 2), HI64)), "r"(idesc));
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
      }
      __syncwarp();
    }
    long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar)));
    long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; done = 1; }
  } else if (warp < 16) {
    const uint32_t taddr = tm + ((uint32_t)((warp & 3) * 32) << 16) + 128u + (uint32_t)((warp >> 2) * 32);   // accumulator half 1 area: not used by the MMAs
    uint32_t v[32];
    uint32_t acc = threadIdx.x;
    long long n = 0;
    while (!done) {
      if (LOAD == 1) {
#pragma unroll
        for (int d = 0; d < 64; ++d) acc = acc * 1664525u + 1013904223u;
      } else if (LOAD == 2) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += v[j] * 3u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr + 256u),
                     "r"(v[0] + acc), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else {
        __nanosleep(200);
      }
      ++n;
    }
    if (threadIdx.x == 0) { out[2] = n; out[3] = acc; }
  }
  __syncthreads();
  if (warp == 17) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
template <int LOAD>
void runx(const char* name) {
  const int iters = 400;
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  auto kern = kx<LOAD>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) kern<<<1, 576, 200 * 1024>>>(d, iters);
  long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-62s %7.1f cyc per 8 MMAs (thread), %7.1f (complete); other-warp iterations %lld  %s\n", name, h[0] / (double)iters,
         h[1] / (double)iters, h[2], e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

template <int N, int MODE>
void run(const char* name) {
  constexpr int NMMA = 64;
  long long* d; cudaMalloc(&d, 16);
  auto kern = k<N, MODE, NMMA>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) kern<<<1, 128, 200 * 1024>>>(d, 1);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-44s issue %6lld cyc (%5.1f/mma)  complete %6lld cyc (%5.1f/mma)  %s\n", name, h[0], h[0] / 64.0, h[1], h[1] / 64.0,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}
template <int N>
void sustained(int nblk, int reps) {
  long long* d; cudaMalloc(&d, 16);
  auto kern = k<N, 3, 64>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<nblk, 128, 200 * 1024>>>(d, reps);
  cudaEventRecord(e0);
  kern<<<nblk, 128, 200 * 1024>>>(d, reps);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double flop = 2.0 * 128 * N * 16 * 64.0 * reps * nblk;
  printf("sustained N=%d on %3d CTAs x %d x 64 MMAs: %7.1f cycles per MMA (block 0), %.3f ms, %.0f TFLOP/s  %s\n", N, nblk, reps,
         (double)h[0] / (64.0 * reps), ms, flop / (ms * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  runx<0>("8 MMAs + commit, 16 other warps idle (nanosleep)");
  runx<1>("8 MMAs + commit, 16 other warps doing ALU work");
  runx<2>("8 MMAs + commit, 16 other warps doing tcgen05.ld/st (TMEM traffic)");
  runc<0, 1>("8 MMAs + 1 commit", 0);
  runc<0, 2>("8 MMAs + 2 commits", 0);
  runc<1, 1>("8 MMAs + ~300 cyc of ALU work + 1 commit", 60);
  runc<2, 1>("8 MMAs + 1 commit + ~300 cyc of ALU work", 60);
  runc<1, 1>("8 MMAs + ~600 cyc of ALU work + 1 commit", 120);
  runc<2, 1>("8 MMAs + 1 commit + ~600 cyc of ALU work", 120);
  sustained<128>(1, 2000);
  sustained<128>(148, 2000);
  sustained<128>(148, 20000);
  sustained<256>(148, 10000);
  run<32, 0>("N=32  A: thread 0, asm per MMA");
  run<32, 1>("N=32  B: elect, asm per MMA");
  run<32, 2>("N=32  C: elect, 8 MMAs per asm block");
  run<128, 0>("N=128 A: thread 0, asm per MMA");
  run<128, 1>("N=128 B: elect, asm per MMA");
  run<128, 2>("N=128 C: elect, 8 MMAs per asm block");
  return 0;
}
