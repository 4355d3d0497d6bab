// tcgen05 (5th-gen tensor core) implementation of the fused encode + two-branch MLP for sm_100a.
//
// Persistent CTA PAIRS (cluster of 2, one CTA per SM, tcgen05 cta_group::2): each CTA owns one M = 128 tile of
// consecutive samples at a time (own X, own TMEM), the pair's leader issues M = 256 MMAs for both, and each CTA
// stages only HALF of every weight slab in its shared memory (the tensor core reads B from both CTAs), which
// halves the shared-memory and L2 traffic per FLOP - the limiter of the 1-CTA version (tools/ubench/mma_rate.cu:
// N=128 MMAs run at 67 cycles alone but ~100 when the B slabs are written and read at 64 B/clk each).
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
// Overlap: every layer's N outputs are computed as two halves.  While the epilogue warps drain half 0, the tensor
// pipe computes half 1; the next layer's half 0 starts on the K range produced by the first epilogue half as soon
// as that is written, and waits for the second only for the remaining K slabs.
// TMEM map (512 columns): [0,256) accumulator halves at 0 / 128, [256,384) and [384,512) activation ping-pong
// (bf16 pairs packed in 32-bit columns: K = 2c, 2c+1 in column c).
// Skip / dir / code concatenations never materialise: a skip layer takes K-slabs from both X and H, and the
// per-ray-constant terms arrive through ray_const (see layout.h).
//
// Reference semantics: models/rendering.py:85-137, models/nerf_model.py:97-152,
// models/embedding_helper.py:325-411, render_tools/multi_rendering.py:16-93.
#include <cuda_bf16.h>

#include "encode.cuh"
#include "field_common.cuh"

namespace {

constexpr int TM = 128;             // samples per tile (UMMA M)
constexpr int NSTAGE = 6;           // weight ring depth
constexpr int SLAB_BYTES = 4096;    // 64 rows x 64 B: this CTA's half of one half K-slab (32 of K) of an N = 256 layer
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
  int pad;
};

// One weight-ring stage of the flattened per-tile program (built on the host, copied to shared memory):
// up to STAGE_SLABS consecutive K-slabs of one layer half.
struct StageDesc {
  uint32_t a_off;       // A operand: from X: offset from sX in 16-byte units; from H: TMEM column of the first slab
  uint32_t src_off;     // byte offset in the packed blob of the first half-slab image
  uint32_t slab_bytes;  // distance between consecutive slabs in the blob (N * 64)
  uint16_t hb16;        // bytes >> 4 of one half-slab (N/2 * 64 >> 4): 512 / 256 / 128
  uint8_t cnt;          // slabs in this stage (1..4)
  uint8_t flags;        // ST_* bits
};
enum : uint8_t {
  ST_FROM_H = 1, ST_FIRST = 2, ST_WAIT_X = 4, ST_WAIT_E0 = 8, ST_WAIT_E1 = 16, ST_COMMIT_ACC = 32, ST_HALF1 = 64
};
constexpr int MAX_STAGES = 112;

struct TcParams {
  FieldParams f;
  StageDesc stages[MAX_STAGES];
  int n_stages;
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
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {   // non-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
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
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive on an mbarrier anywhere in the cluster (address from mapa).  Default .release.cta semantics, as CUTLASS's
// ClusterBarrier::arrive(cta_id): nothing is handed over through generic memory (operands live in each CTA's own
// TMEM / shared memory and are read by the tensor core; tcgen05 fences order them), and cluster-scope
// release / acquire would flush L1 on every epilogue step.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
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
// accumulate flag as a compile-time constant (folds to UPT / !UPT: no register -> predicate round trip)
template <bool ACC>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  if constexpr (ACC)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "l"(desc_a), "l"(desc_b), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "l"(desc_a), "l"(desc_b), "r"(idesc) : "memory");
}
template <bool ACC>
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc) {
  if constexpr (ACC)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "r"(tmem_a), "l"(desc_b), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "r"(tmem_a), "l"(desc_b), "r"(idesc) : "memory");
}
// completion of all prior tcgen05.mma of this thread -> arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar)
      : "memory");
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
__device__ __forceinline__ uint32_t make_idesc(int N) {  // N = columns of ONE mma (a layer half); M = 256 (cta_group::2)
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);
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

// ------------------------------------------------------------------------------------------------
// MMA issue of one ring stage: CNT K-slabs (2 K-steps each) as straight-line code with immediate offsets.
//   FROM_H: A from TMEM (a0 = TMEM address of the first slab's packed K columns), else A from shared memory
//           (a0 = descriptor low word of the first slab; X groups start atom-aligned)
//   HB16:   bytes >> 4 of this CTA's part of one half K-slab in the ring stage (N/4 rows x 64 B)
// ------------------------------------------------------------------------------------------------
template <bool FROM_H, int CNT, int HB16, bool FIRST>
__device__ __forceinline__ void issue_stage(uint32_t d_tmem, uint32_t a0, uint32_t b_lo0, uint32_t idesc) {
#pragma unroll
  for (int i2 = 0; i2 < CNT; ++i2) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t db = make_desc_hl(b_lo0 + (uint32_t)(i2 * HB16 + ks * 2), DESC_HI_SW64);
      if constexpr (FROM_H) {
        const uint32_t ta = a0 + (uint32_t)(i2 * 16 + ks * 8);
        if (FIRST && i2 == 0 && ks == 0) umma_ts<false>(d_tmem, ta, db, idesc);
        else umma_ts<true>(d_tmem, ta, db, idesc);
      } else {
        const uint64_t da = make_desc_hl(a0 + (uint32_t)((i2 >> 1) * (ATOM_BYTES >> 4) + (i2 & 1) * 4 + ks * 2), DESC_HI_SW128);
        if (FIRST && i2 == 0 && ks == 0) umma_ss<false>(d_tmem, da, db, idesc);
        else umma_ss<true>(d_tmem, da, db, idesc);
      }
    }
  }
}
template <bool FROM_H, int HB16>
__device__ __forceinline__ void issue_stage_cnt(int cnt, bool first, uint32_t d_tmem, uint32_t a0, uint32_t b_lo0, uint32_t idesc) {
  if (first) {
    switch (cnt) {
      case 4: issue_stage<FROM_H, 4, HB16, true>(d_tmem, a0, b_lo0, idesc); break;
      case 3: issue_stage<FROM_H, 3, HB16, true>(d_tmem, a0, b_lo0, idesc); break;
      case 2: issue_stage<FROM_H, 2, HB16, true>(d_tmem, a0, b_lo0, idesc); break;
      default: issue_stage<FROM_H, 1, HB16, true>(d_tmem, a0, b_lo0, idesc); break;
    }
  } else {
    switch (cnt) {
      case 4: issue_stage<FROM_H, 4, HB16, false>(d_tmem, a0, b_lo0, idesc); break;
      case 3: issue_stage<FROM_H, 3, HB16, false>(d_tmem, a0, b_lo0, idesc); break;
      case 2: issue_stage<FROM_H, 2, HB16, false>(d_tmem, a0, b_lo0, idesc); break;
      default: issue_stage<FROM_H, 1, HB16, false>(d_tmem, a0, b_lo0, idesc); break;
    }
  }
}
template <bool FROM_H>
__device__ __forceinline__ void issue_stage_n(int N, int cnt, bool first, uint32_t d_tmem, uint32_t a0, uint32_t b_lo0, uint32_t idesc) {
  if (N == 256) issue_stage_cnt<FROM_H, 256>(cnt, first, d_tmem, a0, b_lo0, idesc);
  else if (N == 128) issue_stage_cnt<FROM_H, 128>(cnt, first, d_tmem, a0, b_lo0, idesc);
  else issue_stage_cnt<FROM_H, 64>(cnt, first, d_tmem, a0, b_lo0, idesc);
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
  const uint32_t sStages = sScratch + TM * 4 * 4 * 4;               // [MAX_STAGES] StageDesc
  const uint32_t sBar = sStages + MAX_STAGES * 16;
  const uint32_t bar_full = sBar;                                   // NSTAGE x 8 B
  const uint32_t bar_empty = sBar + 8 * NSTAGE;
  const uint32_t bar_x_ready = sBar + 16 * NSTAGE;                  // compute -> MMA, once per tile
  const uint32_t bar_acc_ready = bar_x_ready + 8;                   // [2] MMA -> compute, per layer half
  const uint32_t bar_epi_done = bar_acc_ready + 16;                 // [2] compute -> MMA, per layer half
  const uint32_t tmem_slot = bar_epi_done + 16;
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));
  float* bias_tab = reinterpret_cast<float*>(gen_base + (sBias - sbase));
  float* scratch = reinterpret_cast<float*>(gen_base + (sScratch - sbase));
  uint4* stage_tab = reinterpret_cast<uint4*>(gen_base + (sStages - sbase));
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - sbase));
  const float* Pf = reinterpret_cast<const float*>(p.packed);

  const uint32_t cta_rank = cluster_ctarank();        // 0 = leader (issues the pair's MMAs)
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      // leader: its own TMA (arrive.expect_tx) + the peer's "my half has landed" remote arrive
      mbar_init(bar_full + 8 * s, cta_rank == 0 ? 2 : 1);
      mbar_init(bar_empty + 8 * s, 1);                // multicast tcgen05.commit
    }
    // compute -> MMA barriers live in the leader and take ONE arrive per compute warp of BOTH CTAs
    mbar_init(bar_x_ready, 2 * NUM_COMPUTE / 32);
    for (int h = 0; h < 2; ++h) {
      mbar_init(bar_acc_ready + 8 * h, 1);            // multicast tcgen05.commit
      mbar_init(bar_epi_done + 8 * h, 2 * NUM_COMPUTE / 32);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, 512);
  // per-column biases of every layer -> shared memory (layers with a per-ray constant read ray_const instead)
  for (int i = threadIdx.x; i < P.n_layers * 256; i += NUM_THREADS) {
    const int l = i >> 8, c = i & 255;
    bias_tab[i] = (c < P.layers[l].N) ? __ldg(Pf + P.layers[l].bias_off + c) : 0.0f;
  }
  for (int i = threadIdx.x; i < P.n_stages; i += NUM_THREADS)
    stage_tab[i] = *reinterpret_cast<const uint4*>(&P.stages[i]);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int64_t total = (int64_t)p.n_rays * p.S;
  const int64_t n_tiles = (total + TM - 1) / TM;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  const int64_t pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;
  // barriers of the leader, as seen from this CTA
  const uint32_t ld_x_ready = mapa(bar_x_ready, 0), ld_epi_done = mapa(bar_epi_done, 0), ld_full = mapa(bar_full, 0);
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(p.packed);

  if (warp == PRODUCER_WARP) {
    // =============================== weight producer (TMA bulk copies) ===============================
    // The whole warp runs the (uniform) loop over the flattened stage program; one elected lane talks to the
    // barriers / TMA.  It runs ahead of the MMA warp by the ring depth, across layers and tiles.
    uint32_t stage = 0, phase = 0;
    for (int64_t pair = pair0; pair < n_pairs; pair += pair_step) {
      for (int si = 0; si < P.n_stages; ++si) {
        const uint4 raw = stage_tab[si];
        const uint32_t my_bytes = (raw.w & 0xFFFFu) << 4, cnt = (raw.w >> 16) & 0xFFu;   // this CTA's rows of a half-slab
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(bar_full + 8 * stage, cnt * my_bytes);
          const uint8_t* src = blob + raw.y + cta_rank * my_bytes;
          for (uint32_t i2 = 0; i2 < cnt; ++i2)
            tma_bulk_g2s(sB + stage * STAGE_BYTES + i2 * my_bytes, src + (size_t)i2 * raw.z, my_bytes,
                         bar_full + 8 * stage);
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == MMA_WARP) {
    // =============================== MMA issuer ===============================
    // Warp-uniform loop over the flattened stage program.  The next stage's descriptor is fetched and its
    // `full` barrier probed BEFORE the current stage's MMAs are issued, so neither latency sits between two
    // bursts of tcgen05.mma.  Barrier waits are executed by all lanes, mma / commit by one elected lane.
    uint32_t stage = 0, phase = 0, x_phase = 0, ed_phase0 = 0, ed_phase1 = 0;
    if (cta_rank != 0) {
      // peer CTA: no MMA issue; forward "my half of stage s has landed" to the leader's full barrier
      for (int64_t pair = pair0; pair < n_pairs; pair += pair_step) {
        for (int si = 0; si < P.n_stages; ++si) {
          mbar_wait(bar_full + 8 * stage, phase);
          if (elect_one()) mbar_arrive_cluster(ld_full + 8 * stage);
          __syncwarp();
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    } else {
    int tl_idx = 0;
    uint4 cur = stage_tab[0];
    bool cur_ready = false;
    for (int64_t pair = pair0; pair < n_pairs; pair += pair_step) {
      for (int si = 0; si < P.n_stages; ++si) {
        const uint32_t flags = cur.w >> 24, cnt = (cur.w >> 16) & 0xFFu, hb16 = cur.w & 0xFFFFu;
        const bool tl = P.timeline && blockIdx.x == 0 && pair == pair0 + pair_step && lane == 0;
        if (tl) P.timeline[256 + (tl_idx * 3 + 0)] = clock64();
        if (flags & ST_WAIT_X) { mbar_wait(bar_x_ready, x_phase); x_phase ^= 1; }
        if (flags & ST_WAIT_E0) { mbar_wait(bar_epi_done, ed_phase0); ed_phase0 ^= 1; }
        if (flags & ST_WAIT_E1) { mbar_wait(bar_epi_done + 8, ed_phase1); ed_phase1 ^= 1; }
        if (tl) P.timeline[640 + tl_idx] = clock64() * 256 + (long long)(flags & 0xff);
        if (!cur_ready) mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (tl) P.timeline[256 + (tl_idx * 3 + 1)] = clock64();
        // prefetch the next stage's descriptor and probe its barrier (the stage table wraps around per tile pair)
        const int sn = (si + 1 == P.n_stages) ? 0 : si + 1;
        const uint4 nxt = stage_tab[sn];
        const uint32_t nstage = (stage + 1 == NSTAGE) ? 0u : stage + 1, nphase = (stage + 1 == NSTAGE) ? phase ^ 1u : phase;
        const bool nxt_ready = mbar_test_wait(bar_full + 8 * nstage, nphase);
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + ((flags & ST_HALF1) ? (uint32_t)TM_ACC1 : 0u);
          const uint32_t b_lo0 = (((sB + stage * STAGE_BYTES) >> 4) & 0x3FFFu) | 0x10000u;
          const int N = (int)hb16;                   // per-CTA slab bytes >> 4 == N (N/4 rows x 64 B)
          const uint32_t idesc = make_idesc(N >> 1);
          const bool first = (flags & ST_FIRST) != 0;
          if (flags & ST_FROM_H) issue_stage_n<true>(N, (int)cnt, first, d_tmem, tmem_base + cur.x, b_lo0, idesc);
          else issue_stage_n<false>(N, (int)cnt, first, d_tmem, ((((sX >> 4) + cur.x)) & 0x3FFFu) | 0x10000u, b_lo0, idesc);
          umma_commit(bar_empty + 8 * stage);
          if (flags & ST_COMMIT_ACC) umma_commit(bar_acc_ready + ((flags & ST_HALF1) ? 8u : 0u));
        }
        __syncwarp();
        if (tl) { P.timeline[256 + (tl_idx * 3 + 2)] = clock64(); ++tl_idx; }
        stage = nstage; phase = nphase;
        cur = nxt; cur_ready = nxt_ready;
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
      mbar_arrive_cluster(ld_epi_done);
      mbar_arrive_cluster(ld_epi_done + 8);
    }
    for (int64_t pair = pair0; pair < n_pairs; pair += pair_step) {
      const int64_t tile = 2 * pair + cta_rank;        // may be == n_tiles (odd tail): a dead tile, nothing written
      if (P.timeline && blockIdx.x == 0 && pair == pair0 + pair_step && threadIdx.x == 0) P.timeline[201] = clock64();
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
      if (lane == 0) mbar_arrive_cluster(ld_x_ready);
      if (P.timeline && blockIdx.x == 0 && pair == pair0 + pair_step && threadIdx.x == 0) P.timeline[200] = clock64();

      float sigma_part = 0.0f;
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& Ly = P.layers[l];
        const int NC = Ly.N >> 3;                      // columns this thread handles per layer half (32/16/8)
        const float* headw = nullptr;
        if (Ly.epi == EPI_HIDDEN_SIGMA) headw = Pf + (Ly.branch ? p.L.osigma_w : p.L.sigma_w);
        if (Ly.epi == EPI_DIR) headw = Pf + (Ly.branch ? p.L.orgb_w : p.L.rgb_w);
        float part0 = 0.0f, part1 = 0.0f, part2 = 0.0f;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int n = h * (Ly.N >> 1) + cq * NC;     // first output column of this thread in this half
          const uint32_t acc_addr = lane_taddr + (uint32_t)(h * TM_ACC1 + cq * NC);
          const uint32_t out_addr = lane_taddr + (uint32_t)(Ly.h_out_col + (n >> 1));
          if (h == 0) { mbar_wait(bar_acc_ready, acc_phase0); acc_phase0 ^= 1; }
          else { mbar_wait(bar_acc_ready + 8, acc_phase1); acc_phase1 ^= 1; }
          tc_fence_after();
          if (P.timeline && blockIdx.x < 2 && pair == pair0 + pair_step && threadIdx.x == 0) P.timeline[768 * blockIdx.x + (l * 2 + h) * 4 + 2] = clock64();
          if (NC == 32) epilogue_half<32>(Ly, acc_addr, out_addr, bias_tab + l * 256, rc, headw, n, part0, part1, part2);
          else if (NC == 16) epilogue_half<16>(Ly, acc_addr, out_addr, bias_tab + l * 256, rc, headw, n, part0, part1, part2);
          else epilogue_half<8>(Ly, acc_addr, out_addr, bias_tab + l * 256, rc, headw, n, part0, part1, part2);
          // accumulator half h drained, output activations of this half written
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(ld_epi_done + 8 * h);
          if (P.timeline && blockIdx.x < 2 && pair == pair0 + pair_step && threadIdx.x == 0) P.timeline[768 * blockIdx.x + (l * 2 + h) * 4 + 3] = clock64();
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
  cluster_sync_all();          // no remote arrive / multicast commit may target a CTA that has exited
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
  int n = 0;
  auto add = [&](int gemm, int nx, int nh, int epi, int branch, int rc_base) {
    TcLayer& t = P.layers[n];
    t.N = L.g[gemm].N; t.nslab_x = nx; t.nslab_h = nh; t.epi = epi; t.branch = branch; t.rc_base = rc_base;
    t.h_in_col = (n & 1) ? TM_HB : TM_HA;     // layer n reads what layer n-1 wrote
    t.h_out_col = (n & 1) ? TM_HA : TM_HB;
    t.img_off = L.g[gemm].img_off; t.bias_off = L.g[gemm].bias_off;
    // K-slab groups: X slabs in runs of 4, then the low-K and high-K halves of H in runs of 4
    int ng = 0;
    auto emit = [&](int first, int count, int from_h, int needs_hi) {
      for (int o = 0; o < count; o += STAGE_SLABS) {
        const int c = (count - o < STAGE_SLABS) ? count - o : STAGE_SLABS;
        t.groups[ng++] = (first + o) | (c << 5) | (from_h << 8) | (needs_hi << 9);
      }
    };
    emit(0, nx, 0, 0);
    emit(0, nh / 2, 1, 0);
    emit(nh / 2, nh - nh / 2, 1, 1);
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
  P.x_atoms = L.use_voxel ? 6 : 1;
  // flatten: for each layer, for each half, one stage per K-slab group
  int ns = 0;
  for (int l = 0; l < n; ++l) {
    const TcLayer& t = P.layers[l];
    const uint32_t slab_bytes = (uint32_t)t.N * 64u, half_bytes = slab_bytes / 2;
    bool waited1 = false;   // per layer: the second epilogue half of the previous layer is waited for once
    for (int h = 0; h < 2; ++h) {
      for (int gi = 0; gi < t.ngroups; ++gi) {
        const int grp = t.groups[gi];
        const int first = grp & 31, cnt = (grp >> 5) & 7, from_h = (grp >> 8) & 1, needs_hi = (grp >> 9) & 1;
        if (ns >= MAX_STAGES) { onerf_set_error("field_tc: stage program too long"); return ONERF_ERR_UNSUPPORTED; }
        StageDesc& sd = P.stages[ns++];
        uint8_t fl = 0;
        if (from_h) fl |= ST_FROM_H;
        if (gi == 0) fl |= ST_FIRST;
        if (h == 1) fl |= ST_HALF1;
        if (l == 0 && h == 0 && gi == 0) fl |= ST_WAIT_X;
        if (h == 0 && gi == 0) fl |= ST_WAIT_E0;               // accumulator half 0 drained, low-K activations written
        if (!waited1 && (h == 1 || needs_hi)) { fl |= ST_WAIT_E1; waited1 = true; }   // half 1 drained / high-K written
        if (gi == t.ngroups - 1) fl |= ST_COMMIT_ACC;
        sd.flags = fl;
        sd.cnt = (uint8_t)cnt;
        sd.hb16 = (uint16_t)(half_bytes >> 5);   // this CTA's rows of the half-slab: N/4 rows x 64 B, >> 4  (== N)
        sd.slab_bytes = slab_bytes;
        const int gslab = from_h ? t.nslab_x + first : first;
        sd.src_off = (uint32_t)(t.img_off + (int64_t)gslab * slab_bytes + (int64_t)h * half_bytes);
        sd.a_off = from_h ? (uint32_t)(t.h_in_col + first * 16) : (uint32_t)((first >> 1) * (ATOM_BYTES >> 4));
      }
    }
  }
  P.n_stages = ns;
  const int64_t total = (int64_t)fp.n_rays * fp.S;
  const int64_t tiles = (total + TM - 1) / TM;
  const int64_t pairs = (tiles + 1) / 2;
  int blocks = (int)(2 * pairs < (int64_t)(ctx->num_sms & ~1) ? 2 * pairs : (int64_t)(ctx->num_sms & ~1));
  const size_t smem = 1024 + (size_t)P.x_atoms * ATOM_BYTES + NSTAGE * STAGE_BYTES + MAX_LAYERS * 256 * 4 +
                      TM * 4 * 4 * 4 + MAX_STAGES * 16 + 512;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;   // CTA pair: tcgen05 cta_group::2
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (L.use_voxel) {
    ONERF_CUDA(cudaFuncSetAttribute(field_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ONERF_CUDA(cudaLaunchKernelEx(&cfg, field_tc_kernel<true>, P));
  } else {
    ONERF_CUDA(cudaFuncSetAttribute(field_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ONERF_CUDA(cudaLaunchKernelEx(&cfg, field_tc_kernel<false>, P));
  }
  ONERF_LAUNCH_CHECK(ctx);
  return ONERF_OK;
}
