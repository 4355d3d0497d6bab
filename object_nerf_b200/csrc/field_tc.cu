// tcgen05 (5th-gen tensor core) implementation of the fused encode + two-branch MLP for sm_100a.
//
// One persistent CTA per SM; a CTA owns one M = 128 tile of consecutive samples at a time.
//   warps 0-15 (512 thr)  encode the tile into shared memory (X, bf16, UMMA K-major SWIZZLE_128B atoms) and run
//                         every layer's epilogue: TMEM accumulator -> registers (tcgen05.ld) -> bias / per-ray
//                         constant -> LeakyReLU -> bf16x2 -> back into TMEM (tcgen05.st) as the NEXT layer's A
//                         operand.  Hidden activations never touch shared or global memory.  Heads (sigma,
//                         rgb) are CUDA-core dot products on the fp32 values.
//   warp 16    (1 lane)   streams the weights global -> shared with cp.async.bulk (TMA): one (N/2 x 32) bf16
//                         half K-slab per stage (pre-swizzled SWIZZLE_64B stage images written by pack.cu),
//                         NSTAGE-deep mbarrier ring running ahead across layers and tiles.
//   warp 17    (1 lane)   issues tcgen05.mma (M=128, K=16, bf16 x bf16 -> fp32 in TMEM): A from shared memory
//                         (X slabs) or from TMEM (hidden slabs), B from the weight ring; owns the TMEM allocation.
// Overlap: the N = 256 layers are computed as two halves of 128 outputs.  While the epilogue warps drain half 0, the
// tensor pipe computes half 1; the next layer's half 0 starts on the K range produced by the first epilogue half as
// soon as that is written, and waits for the second only for the remaining K slabs.  Layers with N <= 128 (dir and
// object layers) use one accumulator and full-width MMAs: narrow MMAs are bound by the issue rate of the single
// MMA thread (~45-65 cycles per instruction whatever N), so two N = 64 halves would cost twice the issue time and
// twice the barrier hand-offs for the same tensor work (measured: 700 -> 763 TFLOP/s, profiles/r01_experiments.md).
// TMEM map (512 columns): [0,256) accumulator halves at 0 / 128, [256,384) and [384,512) activation ping-pong
// (bf16 pairs packed in 32-bit columns: K = 2c, 2c+1 in column c).
// Skip / dir / code concatenations never materialise: a skip layer takes K-slabs from both X and H, and the
// per-ray-constant terms arrive through ray_const (see layout.h).
//
// Reference semantics: models/rendering.py:85-137, models/nerf_model.py:97-152,
// models/embedding_helper.py:325-411, render_tools/multi_rendering.py:16-93.
#include <cuda_bf16.h>
#include <cstdlib>

#include "encode.cuh"
#include "field_common.cuh"

// In-kernel clock64() timeline (tools/timeline.py): compiled in only with -DONERF_TIMELINE (make TIMELINE=1).
#ifdef ONERF_TIMELINE
#define ONERF_TL_ON true
#else
#define ONERF_TL_ON false
#endif

namespace {

constexpr int TM = 128;             // samples per tile (UMMA M)
constexpr int NSTAGE = 3;           // weight ring depth
constexpr int SLAB_BYTES = 8192;    // 128 rows x 64 B: one half K-slab (32 of K) of an N = 256 layer
constexpr int STAGE_SLABS = 4;      // a ring stage carries up to 4 consecutive K-slabs (128 of K) of one layer half
constexpr int STAGE_BYTES = STAGE_SLABS * SLAB_BYTES;
constexpr int MAX_GROUPS = 6;
constexpr int TM_ACC1 = 128;        // TMEM column of accumulator half 1
constexpr int TM_HA = 256, TM_HB = 384;
constexpr int ATOM_BYTES = 16384;   // 128 rows x 128 B (64 bf16 of K)
constexpr int NUM_COMPUTE = 512;    // 16 encode/epilogue warps: 4 per TMEM lane quarter
constexpr int PRODUCER_WARP = 16, MMA_WARP = 17;
constexpr int NUM_THREADS = 576;
constexpr int MAX_LAYERS = 16;
constexpr float kLeaky = 0.01f;

enum Epi { EPI_HIDDEN = 0, EPI_HIDDEN_RC = 1, EPI_HIDDEN_SIGMA = 2, EPI_FINAL = 3, EPI_DIR = 4 };

struct TcLayer {
  int N;           // outputs (UMMA N)
  int nslab_x;     // leading K slabs (32 wide) taken from X
  int nslab_h;     // following K slabs taken from H
  int epi;         // Epi
  int branch;      // 0 scene, 1 object
  int rc_base;     // ray_const offset for *_RC / DIR epilogues
  int h_in_col;    // TMEM column of the input activations (K pairs), if nslab_h > 0
  int h_out_col;   // TMEM column the epilogue writes the output activations to
  int64_t img_off;   // byte offset of this layer's stage images in the packed blob
  int64_t bias_off;  // float offset of the bias vector
  // K-slab groups (one ring stage each), identical for both halves of the layer:
  //   bits [0,5) first slab (index inside X or H), [5,8) slab count (1..4), bit 8: from H, bit 9: needs the
  //   second epilogue half of the previous layer (high-K half of the input activations)
  int ngroups;
  int groups[MAX_GROUPS];
  int nhalf;       // 2: the N outputs are computed as two halves (accumulators 0 / 1); 1: one N <= 128 accumulator
  int prev_two;    // the previous layer (cyclically) has two halves, i.e. posts a second epilogue-done arrival
};

struct TcParams {
  FieldParams f;
  long long* timeline;   // debug: per-event clock64() of block 0, second tile (null = off); see tools/timeline.py
  TcLayer layers[MAX_LAYERS];
  int n_layers;
  int x_atoms;     // 6 (voxel) or 1 (plain)
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must become an error, not a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("onerf field_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
             bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] . B[smem desc]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with A taken from TMEM (bf16 pairs per column)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4 (8-row group pitch),
//   [46,48) version = 1, [61,64) layout type (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
}
// same, from a precomputed high word and a low word ((addr >> 4) & 0x3FFF) | (1 << 16)
constexpr uint32_t DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t DESC_HI_SW64 = (512u >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ uint64_t make_desc_hl(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int N) {  // N = columns of ONE mma (a layer half)
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// 16-byte chunk `chunk` (8 bf16 of K) of row `row` in an A buffer made of SWIZZLE_128B atoms (64 K per atom)
__device__ __forceinline__ uint32_t a_chunk_addr(uint32_t base, int row, int chunk) {
  return base + (uint32_t)(chunk >> 3) * ATOM_BYTES + (uint32_t)row * 128u + (uint32_t)(((chunk & 7) ^ (row & 7)) << 4);
}
__device__ __forceinline__ void st_chunk(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// PE of 8 channels -> 13 chunks [f | sin 2^k f | cos 2^k f]_k at chunk0 + stride * block
__device__ __forceinline__ void pe8_to_chunks(uint32_t xbase, int row, int chunk0, int stride, const float* f) {
  float s[8], c[8];
  st_chunk(a_chunk_addr(xbase, row, chunk0), pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
           pack_bf16(f[6], f[7]));
#pragma unroll
  for (int j = 0; j < 8; ++j) __sincosf(f[j], &s[j], &c[j]);  // |f| = O(1), 6 octaves: error stays << bf16 ulp
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    st_chunk(a_chunk_addr(xbase, row, chunk0 + stride * (1 + 2 * k)), pack_bf16(s[0], s[1]), pack_bf16(s[2], s[3]),
             pack_bf16(s[4], s[5]), pack_bf16(s[6], s[7]));
    st_chunk(a_chunk_addr(xbase, row, chunk0 + stride * (2 + 2 * k)), pack_bf16(c[0], c[1]), pack_bf16(c[2], c[3]),
             pack_bf16(c[4], c[5]), pack_bf16(c[6], c[7]));
    if (k < 5) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // double-angle step to the next octave
        const float s2 = 2.0f * s[j] * c[j];
        c[j] = fmaf(-2.0f * s[j], s[j], 1.0f);
        s[j] = s2;
      }
    }
  }
}

// PE10(xyz): 63 values in reference order + one zero -> 8 chunks starting at chunk0.  Values are packed to
// bf16 pairs as they are produced (the stream position is a compile-time constant after unrolling).
__device__ __forceinline__ void pe_xyz_to_chunks(uint32_t xbase, int row, int chunk0, float x, float y, float z) {
  uint32_t pk[32];
  float pend = 0.0f;
  int pos = 0;
  auto emit = [&](float val) {
    if ((pos & 1) == 0) pend = val;
    else pk[pos >> 1] = pack_bf16(pend, val);
    ++pos;
  };
  emit(x); emit(y); emit(z);
  float s[3], c[3];
  sincosf(x, &s[0], &c[0]);
  sincosf(y, &s[1], &c[1]);
  sincosf(z, &s[2], &c[2]);
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    emit(s[0]); emit(s[1]); emit(s[2]);
    emit(c[0]); emit(c[1]); emit(c[2]);
    if (k < 9) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {  // double-angle step to the next octave
        const float s2 = 2.0f * s[j] * c[j];
        c[j] = fmaf(-2.0f * s[j], s[j], 1.0f);
        s[j] = s2;
      }
    }
  }
  emit(0.0f);
#pragma unroll
  for (int q = 0; q < 8; ++q)
    st_chunk(a_chunk_addr(xbase, row, chunk0 + q), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

// ------------------------------------------------------------------------------------------------
// epilogue of one layer half for one thread: NC accumulator columns of its row
// ------------------------------------------------------------------------------------------------
template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v) {
  if constexpr (NC == 32) { tmem_ld16(taddr, v); tmem_ld16(taddr + 16, v + 16); }
  else if constexpr (NC == 16) { tmem_ld16(taddr, v); }
  else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
  }
}
template <int NP>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t* v) {
  if constexpr (NP == 16) tmem_st16(taddr, v);
  else if constexpr (NP == 8) tmem_st8(taddr, v);
  else asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                    "r"(v[2]), "r"(v[3]) : "memory");
}

__device__ __forceinline__ uint32_t leaky_bf16x2(uint32_t x) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&x);
  const __nv_bfloat162 slope = __floats2bfloat162_rn(kLeaky, kLeaky);
  v = __hmax2(v, __hmul2(v, slope));
  return *reinterpret_cast<uint32_t*>(&v);
}

// Hidden / final layer: t = acc + bias in fp32, one rounding to bf16, LeakyReLU on packed bf16 pairs,
// result written to TMEM as the next layer's A operand.  BIAS_GLOBAL: per-ray constant from global memory.
template <int NC, bool ACT, bool BIAS_GLOBAL>
__device__ __forceinline__ void epi_hidden(uint32_t acc_addr, const float* bias, uint32_t out_addr) {
  uint32_t v[NC];
  tmem_ld_cols<NC>(acc_addr, v);
  tmem_ld_wait();
  uint32_t pk[NC / 2];
#pragma unroll
  for (int j4 = 0; j4 < NC / 4; ++j4) {
    float4 b;
    if (BIAS_GLOBAL) b = __ldg(reinterpret_cast<const float4*>(bias) + j4);
    else b = *(reinterpret_cast<const float4*>(bias) + j4);
    uint32_t p0 = pack_bf16(__uint_as_float(v[4 * j4 + 0]) + b.x, __uint_as_float(v[4 * j4 + 1]) + b.y);
    uint32_t p1 = pack_bf16(__uint_as_float(v[4 * j4 + 2]) + b.z, __uint_as_float(v[4 * j4 + 3]) + b.w);
    if (ACT) { p0 = leaky_bf16x2(p0); p1 = leaky_bf16x2(p1); }
    pk[2 * j4] = p0;
    pk[2 * j4 + 1] = p1;
  }
  tmem_st_cols<NC / 2>(out_addr, pk);
  tmem_st_wait();
}

// Last hidden layer of a branch: like epi_hidden, plus the sigma head as an fp32 dot product on the
// un-rounded activations (reference: sigma = Linear(h), models/nerf_model.py:108,140).
template <int NC>
__device__ __forceinline__ float epi_hidden_sigma(uint32_t acc_addr, const float* bias, const float* headw, uint32_t out_addr) {
  uint32_t v[NC];
  tmem_ld_cols<NC>(acc_addr, v);
  tmem_ld_wait();
  uint32_t pk[NC / 2];
  float part = 0.0f;
#pragma unroll
  for (int j4 = 0; j4 < NC / 4; ++j4) {
    const float4 b = *(reinterpret_cast<const float4*>(bias) + j4);
    const float4 w = __ldg(reinterpret_cast<const float4*>(headw) + j4);
    float t0 = __uint_as_float(v[4 * j4 + 0]) + b.x, t1 = __uint_as_float(v[4 * j4 + 1]) + b.y;
    float t2 = __uint_as_float(v[4 * j4 + 2]) + b.z, t3 = __uint_as_float(v[4 * j4 + 3]) + b.w;
    t0 = fmaxf(t0, t0 * kLeaky); t1 = fmaxf(t1, t1 * kLeaky); t2 = fmaxf(t2, t2 * kLeaky); t3 = fmaxf(t3, t3 * kLeaky);
    part = fmaf(t0, w.x, part); part = fmaf(t1, w.y, part); part = fmaf(t2, w.z, part); part = fmaf(t3, w.w, part);
    pk[2 * j4] = pack_bf16(t0, t1);
    pk[2 * j4 + 1] = pack_bf16(t2, t3);
  }
  tmem_st_cols<NC / 2>(out_addr, pk);
  tmem_st_wait();
  return part;
}

// Direction layer: LeakyReLU(acc + per-ray constant) feeds the 3-wide rgb head directly (fp32 dots).
template <int NC>
__device__ __forceinline__ void epi_dir(uint32_t acc_addr, const float* rcbias, const float* headw, int head_ld,
                                        float& p0, float& p1, float& p2) {
  uint32_t v[NC];
  tmem_ld_cols<NC>(acc_addr, v);
  tmem_ld_wait();
#pragma unroll
  for (int j4 = 0; j4 < NC / 4; ++j4) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(rcbias) + j4);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(headw) + j4);
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(headw + head_ld) + j4);
    const float4 w2 = __ldg(reinterpret_cast<const float4*>(headw + 2 * head_ld) + j4);
    float t0 = __uint_as_float(v[4 * j4 + 0]) + b.x, t1 = __uint_as_float(v[4 * j4 + 1]) + b.y;
    float t2 = __uint_as_float(v[4 * j4 + 2]) + b.z, t3 = __uint_as_float(v[4 * j4 + 3]) + b.w;
    t0 = fmaxf(t0, t0 * kLeaky); t1 = fmaxf(t1, t1 * kLeaky); t2 = fmaxf(t2, t2 * kLeaky); t3 = fmaxf(t3, t3 * kLeaky);
    p0 = fmaf(t0, w0.x, p0); p0 = fmaf(t1, w0.y, p0); p0 = fmaf(t2, w0.z, p0); p0 = fmaf(t3, w0.w, p0);
    p1 = fmaf(t0, w1.x, p1); p1 = fmaf(t1, w1.y, p1); p1 = fmaf(t2, w1.z, p1); p1 = fmaf(t3, w1.w, p1);
    p2 = fmaf(t0, w2.x, p2); p2 = fmaf(t1, w2.y, p2); p2 = fmaf(t2, w2.z, p2); p2 = fmaf(t3, w2.w, p2);
  }
}

// dispatch on the (warp-uniform) layer width; NC = N / 8 columns per thread per layer half
template <int NC>
__device__ __forceinline__ void epilogue_half(const TcLayer& Ly, uint32_t acc_addr, uint32_t out_addr, const float* bias_smem,
                                              const float* rc, const float* headw, int n, float& part0, float& part1,
                                              float& part2) {
  if constexpr (NC >= 16) {
    switch (Ly.epi) {
      case EPI_HIDDEN: epi_hidden<NC, true, false>(acc_addr, bias_smem + n, out_addr); break;
      case EPI_HIDDEN_RC: epi_hidden<NC, true, true>(acc_addr, rc + Ly.rc_base + n, out_addr); break;
      case EPI_FINAL: epi_hidden<NC, false, false>(acc_addr, bias_smem + n, out_addr); break;
      case EPI_HIDDEN_SIGMA: part0 += epi_hidden_sigma<NC>(acc_addr, bias_smem + n, headw + n, out_addr); break;
      default: epi_dir<NC>(acc_addr, rc + Ly.rc_base + n, headw + n, Ly.N, part0, part1, part2); break;
    }
  } else {
    epi_dir<NC>(acc_addr, rc + Ly.rc_base + n, headw + n, Ly.N, part0, part1, part2);  // only the N = 64 dir layer
  }
}

template <bool VOXEL>
__global__ void __launch_bounds__(NUM_THREADS, 1) field_tc_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const FieldParams& p = P.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int X_ATOMS = VOXEL ? 6 : 1;

  // ---- shared memory carve-up (base is 1024-byte aligned: required by the 128B swizzle) ----
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = sbase;
  const uint32_t sB = sX + X_ATOMS * ATOM_BYTES;
  const uint32_t sBias = sB + NSTAGE * STAGE_BYTES;                 // [MAX_LAYERS][256] floats
  const uint32_t sScratch = sBias + MAX_LAYERS * 256 * 4;           // [128][4][4] floats
  const uint32_t sBar = sScratch + TM * 4 * 4 * 4;
  const uint32_t bar_full = sBar;                                   // NSTAGE x 8 B
  const uint32_t bar_empty = sBar + 8 * NSTAGE;
  const uint32_t bar_x_ready = sBar + 16 * NSTAGE;                  // compute -> MMA, once per tile
  const uint32_t bar_acc_ready = bar_x_ready + 8;                   // [2] MMA -> compute, per layer half
  const uint32_t bar_epi_done = bar_acc_ready + 16;                 // [2] compute -> MMA, per layer half
  const uint32_t tmem_slot = bar_epi_done + 16;
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));
  float* bias_tab = reinterpret_cast<float*>(gen_base + (sBias - sbase));
  float* scratch = reinterpret_cast<float*>(gen_base + (sScratch - sbase));
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - sbase));
  const float* Pf = reinterpret_cast<const float*>(p.packed);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    // compute -> MMA barriers take ONE arrive per warp (after __syncwarp): 512 serialized shared-memory
    // atomics per phase would cost more than the epilogue math
    mbar_init(bar_x_ready, NUM_COMPUTE / 32);
    for (int h = 0; h < 2; ++h) {
      mbar_init(bar_acc_ready + 8 * h, 1);
      mbar_init(bar_epi_done + 8 * h, NUM_COMPUTE / 32);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, 512);
  // per-column biases of every layer -> shared memory (layers with a per-ray constant read ray_const instead)
  for (int i = threadIdx.x; i < P.n_layers * 256; i += NUM_THREADS) {
    const int l = i >> 8, c = i & 255;
    bias_tab[i] = (c < P.layers[l].N) ? __ldg(Pf + P.layers[l].bias_off + c) : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int64_t total = (int64_t)p.n_rays * p.S;
  const int64_t n_tiles = (total + TM - 1) / TM;
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(p.packed);

  if (warp == PRODUCER_WARP) {
    // =============================== weight producer (TMA bulk copies) ===============================
    // The whole warp runs the (uniform) loop; one elected lane talks to the barriers / TMA.
    uint32_t stage = 0, phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& Ly = P.layers[l];
        const uint32_t slab_bytes = (uint32_t)Ly.N * 64u, half_bytes = slab_bytes >> (Ly.nhalf - 1);
        const uint8_t* src = blob + Ly.img_off;
        for (int h = 0; h < Ly.nhalf; ++h) {
          for (int gi = 0; gi < Ly.ngroups; ++gi) {
            const int grp = Ly.groups[gi];
            const int first = grp & 31, cnt = (grp >> 5) & 7;
            const int gslab = ((grp >> 8) & 1) ? Ly.nslab_x + first : first;   // slab index inside the layer
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(bar_full + 8 * stage, (uint32_t)cnt * half_bytes);
              for (int i2 = 0; i2 < cnt; ++i2)
                tma_bulk_g2s(sB + stage * STAGE_BYTES + (uint32_t)i2 * half_bytes,
                             src + (size_t)(gslab + i2) * slab_bytes + (size_t)h * half_bytes, half_bytes,
                             bar_full + 8 * stage);
            }
            __syncwarp();
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =============================== MMA issuer ===============================
    // Warp-uniform control flow (barrier waits by all lanes), tcgen05.mma / commit by one elected lane.
    uint32_t stage = 0, phase = 0, x_phase = 0, ed_phase0 = 0, ed_phase1 = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& Ly = P.layers[l];
        const uint32_t idesc = make_idesc(Ly.N >> (Ly.nhalf - 1));
        const uint32_t half_bytes = ((uint32_t)Ly.N * 64u) >> (Ly.nhalf - 1);
        if (l == 0) {  // this tile's X is encoded
          mbar_wait(bar_x_ready, x_phase);
          x_phase ^= 1;
        }
        // accumulator half 0 drained and the low-K half of the input activations written (previous layer,
        // or the previous tile's last layer)
        mbar_wait(bar_epi_done, ed_phase0);
        ed_phase0 ^= 1;
        tc_fence_after();
        bool waited1 = !Ly.prev_two;   // a second epilogue-done arrival exists only after a two-half layer
        for (int h = 0; h < Ly.nhalf; ++h) {
          if (h == 1 && !waited1) {
            mbar_wait(bar_epi_done + 8, ed_phase1);
            ed_phase1 ^= 1;
            tc_fence_after();
            waited1 = true;
          }
          const uint32_t d_tmem = tmem_base + (uint32_t)(h * TM_ACC1);
          if (ONERF_TL_ON && P.timeline && blockIdx.x == 0 && tile == (int64_t)gridDim.x && lane == 0) P.timeline[(l * 2 + h) * 4 + 0] = clock64();
          for (int gi = 0; gi < Ly.ngroups; ++gi) {
            const int grp = Ly.groups[gi];
            const int first = grp & 31, cnt = (grp >> 5) & 7;
            const bool from_h = (grp >> 8) & 1;
            // Descriptor words are formed BEFORE the barrier waits (the empty asm pins them there): whatever sits between
            // a satisfied wait and the first tcgen05.mma is pure latency on the layer-to-layer dependency chain.
            // High words are constants; the low words advance by (bytes >> 4) per K step.
            const uint32_t b_lo0 = (((sB + stage * STAGE_BYTES) >> 4) & 0x3FFFu) | 0x10000u;
            const uint32_t hb16 = half_bytes >> 4;
            // X slabs: `first` is a multiple of 4, i.e. atom aligned; slab i2 sits at atom (i2 >> 1), half (i2 & 1)
            const uint32_t a0 = from_h ? tmem_base + (uint32_t)(Ly.h_in_col + first * 16)
                                       : ((((sX + (uint32_t)(first >> 1) * ATOM_BYTES) >> 4) & 0x3FFFu) | 0x10000u);
            const uint32_t accum0 = (gi > 0) ? 1u : 0u;
            asm volatile("" ::"r"(b_lo0), "r"(hb16), "r"(a0), "r"(accum0), "r"(d_tmem), "r"(idesc));
            if (((grp >> 9) & 1) && !waited1) {   // high-K half of the input activations
              mbar_wait(bar_epi_done + 8, ed_phase1);
              ed_phase1 ^= 1;
              tc_fence_after();
              waited1 = true;
            }
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            if (elect_one()) {
              uint32_t accum = accum0;
              if (!from_h) {
#pragma unroll
                for (int i2 = 0; i2 < STAGE_SLABS; ++i2) {
                  if (i2 < cnt) {
                    const uint32_t a_lo = a0 + (uint32_t)(i2 >> 1) * (ATOM_BYTES >> 4) + (uint32_t)(i2 & 1) * 4u;
                    const uint32_t b_lo = b_lo0 + (uint32_t)i2 * hb16;
                    umma_bf16(d_tmem, make_desc_hl(a_lo, DESC_HI_SW128), make_desc_hl(b_lo, DESC_HI_SW64), idesc, accum);
                    umma_bf16(d_tmem, make_desc_hl(a_lo + 2u, DESC_HI_SW128), make_desc_hl(b_lo + 2u, DESC_HI_SW64), idesc, 1u);
                    accum = 1u;
                  }
                }
              } else {
#pragma unroll
                for (int i2 = 0; i2 < STAGE_SLABS; ++i2) {
                  if (i2 < cnt) {
                    const uint32_t b_lo = b_lo0 + (uint32_t)i2 * hb16;
                    umma_bf16_ts(d_tmem, a0 + (uint32_t)i2 * 16u, make_desc_hl(b_lo, DESC_HI_SW64), idesc, accum);
                    umma_bf16_ts(d_tmem, a0 + (uint32_t)i2 * 16u + 8u, make_desc_hl(b_lo + 2u, DESC_HI_SW64), idesc, 1u);
                    accum = 1u;
                  }
                }
              }
              umma_commit(bar_empty + 8 * stage);
              if (gi == Ly.ngroups - 1) umma_commit(bar_acc_ready + 8 * h);
            }
            if (gi == Ly.ngroups - 1 && ONERF_TL_ON && P.timeline && blockIdx.x == 0 && tile == (int64_t)gridDim.x && lane == 0)
              P.timeline[(l * 2 + h) * 4 + 1] = clock64();
            __syncwarp();
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
        if (!waited1) {   // one-half layer without high-K input after a two-half layer: keep the barrier phase in step
          mbar_wait(bar_epi_done + 8, ed_phase1);
          ed_phase1 ^= 1;
        }
      }
    }
  } else {
    // =============================== encode + epilogue warps ===============================
    const int q = warp & 3, cq = warp >> 2;          // TMEM lane quarter (rows), column quarter
    const int row = q * 32 + lane;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_phase0 = 0, acc_phase1 = 0;
    // nothing to drain before the very first layer
    if (lane == 0) {
      mbar_arrive(bar_epi_done);
      if (P.layers[0].prev_two) mbar_arrive(bar_epi_done + 8);
    }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (ONERF_TL_ON && P.timeline && blockIdx.x == 0 && tile == (int64_t)gridDim.x && threadIdx.x == 0) P.timeline[201] = clock64();
      const int64_t e = tile * TM + row;
      const bool live = e < total;
      const int ray = live ? (int)(e / p.S) : 0;
      const int si = live ? (int)(e - (int64_t)ray * p.S) : 0;
      const float* rr = p.rays + (int64_t)ray * 8;
      const float zz = live ? __ldg(p.z + (int64_t)ray * p.z_stride + si) : 0.0f;
      float x = fmaf(__ldg(rr + 3), zz, __ldg(rr + 0));
      float y = fmaf(__ldg(rr + 4), zz, __ldg(rr + 1));
      float z = fmaf(__ldg(rr + 5), zz, __ldg(rr + 2));
      if (p.xyz && live) {
        const float* qq = p.xyz + ((int64_t)ray * p.S + si) * 3;
        x = __ldg(qq); y = __ldg(qq + 1); z = __ldg(qq + 2);
      }
      if (!live) { x = 0.f; y = 0.f; z = 0.f; }
      int mute = 0;  // bit 0: scene sigma muted, bit 1: object sigma muted
      if (cq == 0) {
        if (live && p.mute_zero_rays && __ldg(p.z + (int64_t)ray * p.z_stride + (p.S - 1)) == 0.0f) mute = 3;
        if (live && mute == 0 && p.n_boxes > 0 && point_in_boxes(p.boxes, p.n_boxes, x, y, z)) mute = 1;
      }
      const float* rc = p.ray_const + (int64_t)ray * ONERF_RAY_CONST_FLOATS;

      // ---- encode this thread's quarter of the row of X ----
      if (VOXEL) {
        const GridView g = load_grid_view(p.grid);
        float f[8];
        if (cq == 0) {
          voxel_trilinear<0, 8, false>(g, x, y, z, f);
          pe8_to_chunks(sX, row, 0, 2, f);        // scene channels 0-7 : chunks 0, 2, 4, ...
        } else if (cq == 1) {
          voxel_trilinear<8, 8, false>(g, x, y, z, f);
          pe8_to_chunks(sX, row, 1, 2, f);        // scene channels 8-15: chunks 1, 3, 5, ...
        } else if (cq == 2) {
          voxel_trilinear<16, 8, false>(g, x, y, z, f);
          pe8_to_chunks(sX, row, 34, 1, f);       // object voxel block starts at column 272 = chunk 34
        } else {
          pe_xyz_to_chunks(sX, row, 26, x, y, z); // columns 208..271
          st_chunk(a_chunk_addr(sX, row, 47), 0u, 0u, 0u, 0u);  // columns 376..383
        }
      } else {
        if (cq == 0) pe_xyz_to_chunks(sX, row, 0, x, y, z);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_x_ready);
      if (ONERF_TL_ON && P.timeline && blockIdx.x == 0 && tile == (int64_t)gridDim.x && threadIdx.x == 0) P.timeline[200] = clock64();

      float sigma_part = 0.0f;
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& Ly = P.layers[l];
        const int HW = Ly.N >> (Ly.nhalf - 1);               // accumulator width of one layer half
        const int NC = HW >> 2;                        // columns this thread handles per layer half (32/16/8)
        const float* headw = nullptr;
        if (Ly.epi == EPI_HIDDEN_SIGMA) headw = Pf + (Ly.branch ? p.L.osigma_w : p.L.sigma_w);
        if (Ly.epi == EPI_DIR) headw = Pf + (Ly.branch ? p.L.orgb_w : p.L.rgb_w);
        float part0 = 0.0f, part1 = 0.0f, part2 = 0.0f;
#pragma unroll 1
        for (int h = 0; h < Ly.nhalf; ++h) {
          const int n = h * HW + cq * NC;              // first output column of this thread in this half
          const uint32_t acc_addr = lane_taddr + (uint32_t)(h * TM_ACC1 + cq * NC);
          const uint32_t out_addr = lane_taddr + (uint32_t)(Ly.h_out_col + (n >> 1));
          if (h == 0) { mbar_wait(bar_acc_ready, acc_phase0); acc_phase0 ^= 1; }
          else { mbar_wait(bar_acc_ready + 8, acc_phase1); acc_phase1 ^= 1; }
          tc_fence_after();
          if (ONERF_TL_ON && P.timeline && blockIdx.x == 0 && tile == (int64_t)gridDim.x && threadIdx.x == 0) P.timeline[(l * 2 + h) * 4 + 2] = clock64();
          if (NC == 32) epilogue_half<32>(Ly, acc_addr, out_addr, bias_tab + l * 256, rc, headw, n, part0, part1, part2);
          else if (NC == 16) epilogue_half<16>(Ly, acc_addr, out_addr, bias_tab + l * 256, rc, headw, n, part0, part1, part2);
          else epilogue_half<8>(Ly, acc_addr, out_addr, bias_tab + l * 256, rc, headw, n, part0, part1, part2);
          // accumulator half h drained, output activations of this half written
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_epi_done + 8 * h);
          if (ONERF_TL_ON && P.timeline && blockIdx.x == 0 && tile == (int64_t)gridDim.x && threadIdx.x == 0) P.timeline[(l * 2 + h) * 4 + 3] = clock64();
        }
        if (Ly.epi == EPI_HIDDEN_SIGMA) sigma_part = part0;
        if (Ly.epi == EPI_DIR) {
          // combine the four column quarters of this row through shared memory, finish the heads, write out
          float* sc = scratch + (row * 4 + cq) * 4;
          sc[0] = sigma_part; sc[1] = part0; sc[2] = part1; sc[3] = part2;
          asm volatile("bar.sync 1, %0;" ::"n"(NUM_COMPUTE) : "memory");
          if (cq == 0 && live) {
            const float4 a1 = *reinterpret_cast<const float4*>(scratch + (row * 4 + 1) * 4);
            const float4 a2 = *reinterpret_cast<const float4*>(scratch + (row * 4 + 2) * 4);
            const float4 a3 = *reinterpret_cast<const float4*>(scratch + (row * 4 + 3) * 4);
            const float* hb = Pf + (Ly.branch ? p.L.orgb_b : p.L.rgb_b);
            float sg = sigma_part + a1.x + a2.x + a3.x + __ldg(Pf + (Ly.branch ? p.L.osigma_b : p.L.sigma_b));
            const float r = 1.0f / (1.0f + __expf(-(part0 + a1.y + a2.y + a3.y + __ldg(hb + 0))));
            const float gch = 1.0f / (1.0f + __expf(-(part1 + a1.z + a2.z + a3.z + __ldg(hb + 1))));
            const float b = 1.0f / (1.0f + __expf(-(part2 + a1.w + a2.w + a3.w + __ldg(hb + 2))));
            if (mute & (Ly.branch ? 2 : 1)) sg = -1e5f;
            float* outp = Ly.branch ? p.obj_out : p.scene_out;
            reinterpret_cast<float4*>(outp)[(int64_t)ray * p.out_stride + si] = make_float4(r, gch, b, sg);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(NUM_COMPUTE) : "memory");  // scratch reusable
        }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, 512);
}

}  // namespace

static long long* g_timeline = nullptr;
extern "C" void onerf_debug_timeline(void* dev_buf) { g_timeline = reinterpret_cast<long long*>(dev_buf); }

int onerf_launch_field_bf16(onerf_ctx* ctx, const FieldParams& fp, cudaStream_t stream) {
  const PackLayout& L = fp.L;
  TcParams P;
  memset(&P, 0, sizeof(P));
  P.f = fp;
  P.timeline = g_timeline;
  const int xs = L.KX / 32, xo = L.KO / 32;
  // Layers with N <= 128 use ONE accumulator and full-width MMAs: half as many tcgen05.mma issues and barrier
  // hand-offs as two N = 64 halves (the MMA issue rate, not the tensor pipe, bounds narrow layers).
  // ONERF_TC_SPLIT_ALL=1 restores two halves everywhere (A/B measurements).
  static const int single_max_n = [] { const char* v = getenv("ONERF_TC_SPLIT_ALL"); return (v && v[0] == '1') ? 0 : 128; }();
  int n = 0;
  auto add = [&](int gemm, int nx, int nh, int epi, int branch, int rc_base) {
    TcLayer& t = P.layers[n];
    t.N = L.g[gemm].N; t.nslab_x = nx; t.nslab_h = nh; t.epi = epi; t.branch = branch; t.rc_base = rc_base;
    t.h_in_col = (n & 1) ? TM_HB : TM_HA;     // layer n reads what layer n-1 wrote
    t.h_out_col = (n & 1) ? TM_HA : TM_HB;
    t.img_off = L.g[gemm].img_off; t.bias_off = L.g[gemm].bias_off;
    t.nhalf = (t.N <= single_max_n) ? 1 : 2;
    const bool in_two = n > 0 && P.layers[n - 1].nhalf == 2;   // the input activations were written in two halves
    // K-slab groups: X slabs in runs of 4, then the low-K and high-K halves of H in runs of 4
    int ng = 0;
    auto emit = [&](int first, int count, int from_h, int needs_hi) {
      for (int o = 0; o < count; o += STAGE_SLABS) {
        const int c = (count - o < STAGE_SLABS) ? count - o : STAGE_SLABS;
        t.groups[ng++] = (first + o) | (c << 5) | (from_h << 8) | (needs_hi << 9);
      }
    };
    emit(0, nx, 0, 0);
    if (in_two) {
      emit(0, nh / 2, 1, 0);
      emit(nh / 2, nh - nh / 2, 1, 1);
    } else {
      emit(0, nh, 1, 0);
    }
    t.ngroups = ng;
    ++n;
  };
  if (fp.want_scene) {
    add(G_S0, xs, 0, EPI_HIDDEN, 0, 0);
    add(G_S1, 0, 8, EPI_HIDDEN, 0, 0);
    add(G_S2, 0, 8, EPI_HIDDEN, 0, 0);
    add(G_S3, 0, 8, EPI_HIDDEN, 0, 0);
    add(G_S4, xs, 8, EPI_HIDDEN, 0, 0);
    add(G_S5, 0, 8, EPI_HIDDEN, 0, 0);
    add(G_S6, 0, 8, EPI_HIDDEN, 0, 0);
    add(G_S7, 0, 8, EPI_HIDDEN_SIGMA, 0, 0);
    add(G_SFIN, 0, 8, EPI_FINAL, 0, 0);
    add(G_SDIR, 0, 8, EPI_DIR, 0, RC_SDIR);
  }
  if (fp.want_object) {
    add(G_O0, xo, 0, EPI_HIDDEN_RC, 1, RC_OL0);
    add(G_O1, 0, 4, EPI_HIDDEN, 1, 0);
    add(G_O2, xo, 4, EPI_HIDDEN_RC, 1, RC_OL2);
    add(G_O3, 0, 4, EPI_HIDDEN_SIGMA, 1, 0);
    add(G_OFIN, 0, 4, EPI_FINAL, 1, 0);
    add(G_ODIR, 0, 4, EPI_DIR, 1, RC_ODIR);
  }
  P.n_layers = n;
  for (int i = 0; i < n; ++i) P.layers[i].prev_two = P.layers[(i + n - 1) % n].nhalf == 2;
  P.x_atoms = L.use_voxel ? 6 : 1;
  const int64_t total = (int64_t)fp.n_rays * fp.S;
  const int64_t tiles = (total + TM - 1) / TM;
  const int blocks = (int)(tiles < ctx->num_sms ? tiles : ctx->num_sms);
  const size_t smem = 1024 + (size_t)P.x_atoms * ATOM_BYTES + NSTAGE * STAGE_BYTES + MAX_LAYERS * 256 * 4 +
                      TM * 4 * 4 * 4 + 512;
  if (L.use_voxel) {
    ONERF_CUDA(cudaFuncSetAttribute(field_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    field_tc_kernel<true><<<blocks, NUM_THREADS, smem, stream>>>(P);
  } else {
    ONERF_CUDA(cudaFuncSetAttribute(field_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    field_tc_kernel<false><<<blocks, NUM_THREADS, smem, stream>>>(P);
  }
  ONERF_LAUNCH_CHECK(ctx);
  return ONERF_OK;
}
