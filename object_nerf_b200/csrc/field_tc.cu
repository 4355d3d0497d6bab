// tcgen05 (5th-gen tensor core) implementation of the fused encode + two-branch MLP for sm_100a.
//
// One persistent CTA per SM; a CTA owns one M = 128 tile of consecutive samples at a time.
//   warps 0-7  (256 thr)  encode the tile into shared memory (X, bf16, UMMA K-major SWIZZLE_128B atoms) and run
//                         every layer's epilogue: TMEM -> registers (tcgen05.ld) -> bias / per-ray constant ->
//                         LeakyReLU -> bf16 -> shared memory (H, next layer's A operand); heads (sigma, rgb) on
//                         CUDA cores from the fp32 accumulators.
//   warp 8     (1 lane)   streams the weight K-slabs (N x 32 bf16, pre-swizzled SWIZZLE_64B stage images written by
//                         pack.cu) global -> shared with cp.async.bulk (TMA) through a NSTAGE mbarrier ring.
//   warp 9     (1 lane)   issues tcgen05.mma (M=128, N=256/128/64, K=16, bf16 x bf16 -> fp32 in TMEM) and
//                         tcgen05.commit; also owns the TMEM allocation.
// Skip / dir / code concatenations never materialise: a skip layer simply takes K-slabs from both X and H, and
// the per-ray-constant terms arrive through ray_const (see layout.h).
//
// Reference semantics: models/rendering.py:85-137, models/nerf_model.py:97-152,
// models/embedding_helper.py:325-411, render_tools/multi_rendering.py:16-93.
#include <cuda_bf16.h>

#include "encode.cuh"
#include "field_common.cuh"

namespace {

constexpr int TM = 128;             // samples per tile (UMMA M)
constexpr int NSTAGE = 3;           // weight ring depth
constexpr int STAGE_BYTES = 16384;  // 256 rows x 64 B
constexpr int ATOM_BYTES = 16384;   // 128 rows x 128 B (64 bf16 of K)
constexpr int NUM_COMPUTE = 256;
constexpr int NUM_THREADS = 320;
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
  int acc_col;     // TMEM column of the accumulator
  int pad;
  int64_t img_off;   // byte offset of this layer's stage images in the packed blob
  int64_t bias_off;  // float offset of the bias vector
};

struct TcParams {
  FieldParams f;
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4 (8-row group pitch),
//   [46,48) version = 1, [61,64) layout type (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int N) {
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
  for (int j = 0; j < 8; ++j) sincosf(f[j], &s[j], &c[j]);
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

// PE10(xyz): 63 values in reference order + one zero -> 8 chunks starting at chunk0
__device__ __forceinline__ void pe_xyz_to_chunks(uint32_t xbase, int row, int chunk0, float x, float y, float z) {
  float v[64];
  v[0] = x; v[1] = y; v[2] = z;
  float s[3], c[3];
  sincosf(x, &s[0], &c[0]);
  sincosf(y, &s[1], &c[1]);
  sincosf(z, &s[2], &c[2]);
#pragma unroll
  for (int k = 0; k < 10; ++k) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      v[3 * (1 + 2 * k) + j] = s[j];
      v[3 * (2 + 2 * k) + j] = c[j];
    }
    if (k < 9) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float s2 = 2.0f * s[j] * c[j];
        c[j] = fmaf(-2.0f * s[j], s[j], 1.0f);
        s[j] = s2;
      }
    }
  }
  v[63] = 0.0f;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    st_chunk(a_chunk_addr(xbase, row, chunk0 + q), pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
             pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
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
  const uint32_t sH = sX + X_ATOMS * ATOM_BYTES;
  const uint32_t sB = sH + 4 * ATOM_BYTES;
  const uint32_t sScratch = sB + NSTAGE * STAGE_BYTES;            // [128][2][4] floats
  const uint32_t sBar = sScratch + TM * 2 * 4 * 4;
  const uint32_t bar_full = sBar;                                   // NSTAGE x 8 B
  const uint32_t bar_empty = sBar + 8 * NSTAGE;
  const uint32_t bar_a_ready = sBar + 16 * NSTAGE;                  // compute -> MMA
  const uint32_t bar_acc_ready = bar_a_ready + 8;                   // MMA -> compute
  const uint32_t tmem_slot = bar_acc_ready + 8;
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));
  float* scratch = reinterpret_cast<float*>(gen_base + (sScratch - sbase));
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - sbase));

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_a_ready, NUM_COMPUTE);
    mbar_init(bar_acc_ready, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int64_t total = (int64_t)p.n_rays * p.S;
  const int64_t n_tiles = (total + TM - 1) / TM;
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(p.packed);
  const float* Pf = reinterpret_cast<const float*>(p.packed);

  if (warp == 8) {
    // =============================== weight producer (TMA bulk copies) ===============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < P.n_layers; ++l) {
          const TcLayer& Ly = P.layers[l];
          const uint32_t bytes = (uint32_t)Ly.N * 64u;
          const int nslab = Ly.nslab_x + Ly.nslab_h;
          const uint8_t* src = blob + Ly.img_off;
          for (int j = 0; j < nslab; ++j) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            mbar_expect_tx(bar_full + 8 * stage, bytes);
            tma_bulk_g2s(sB + stage * STAGE_BYTES, src + (size_t)j * bytes, bytes, bar_full + 8 * stage);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 9) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < P.n_layers; ++l) {
          const TcLayer& Ly = P.layers[l];
          const uint32_t idesc = make_idesc(Ly.N);
          const uint32_t d_tmem = tmem_base + (uint32_t)Ly.acc_col;
          const int nslab = Ly.nslab_x + Ly.nslab_h;
          // A operand (X at tile start, H after the previous layer's epilogue) is in shared memory, and the
          // previous accumulator has been drained
          mbar_wait(bar_a_ready, a_phase);
          a_phase ^= 1;
          tc_fence_after();
          for (int j = 0; j < nslab; ++j) {
            const bool from_x = j < Ly.nslab_x;
            const int sj = from_x ? j : j - Ly.nslab_x;                  // 32-wide slab inside X or H
            const uint32_t a_addr = (from_x ? sX : sH) + (uint32_t)(sj >> 1) * ATOM_BYTES + (uint32_t)(sj & 1) * 64u;
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t b_addr = sB + stage * STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t da = make_desc(a_addr + ks * 32u, 1024u, 2u);
              const uint64_t db = make_desc(b_addr + ks * 32u, 512u, 4u);
              umma_bf16(d_tmem, da, db, idesc, (j > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit(bar_empty + 8 * stage);   // slab consumed -> producer may refill
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
          umma_commit(bar_acc_ready);             // accumulator complete -> epilogue
        }
      }
    }
  } else {
    // =============================== encode + epilogue warps ===============================
    const int q = warp & 3, hf = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
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
      if (live && p.mute_zero_rays && __ldg(p.z + (int64_t)ray * p.z_stride + (p.S - 1)) == 0.0f) mute = 3;
      if (live && mute == 0 && p.n_boxes > 0 && point_in_boxes(p.boxes, p.n_boxes, x, y, z)) mute = 1;
      const float* rc = p.ray_const + (int64_t)ray * ONERF_RAY_CONST_FLOATS;

      // ---- encode this row's half of X ----
      if (VOXEL) {
        const GridView g = load_grid_view(p.grid);
        if (hf == 0) {
          float f[16];
          voxel_trilinear<0, 16, false>(g, x, y, z, f);
          pe8_to_chunks(sX, row, 0, 2, f);        // channels 0-7 : chunks 0, 2, 4, ...
          pe8_to_chunks(sX, row, 1, 2, f + 8);    // channels 8-15: chunks 1, 3, 5, ...
        } else {
          float f[8];
          voxel_trilinear<16, 8, false>(g, x, y, z, f);
          pe8_to_chunks(sX, row, 34, 1, f);       // object voxel block starts at column 272 = chunk 34
          pe_xyz_to_chunks(sX, row, 26, x, y, z); // columns 208..271
          st_chunk(a_chunk_addr(sX, row, 47), 0u, 0u, 0u, 0u);  // columns 376..383
        }
      } else {
        if (hf == 0) pe_xyz_to_chunks(sX, row, 0, x, y, z);
      }
      fence_async_smem();
      mbar_arrive(bar_a_ready);

      float sigma_part = 0.0f;
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& Ly = P.layers[l];
        const int ncol = Ly.N >> 1;                 // columns handled by this thread
        const int col0 = hf * ncol;
        mbar_wait(bar_acc_ready, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        const bool to_h = Ly.epi != EPI_DIR;
        const bool act = Ly.epi != EPI_FINAL;
        const bool per_ray = (Ly.epi == EPI_HIDDEN_RC) || (Ly.epi == EPI_DIR);
        const float* bias = per_ray ? (rc + Ly.rc_base) : (Pf + Ly.bias_off);
        const float* headw = nullptr;
        if (Ly.epi == EPI_HIDDEN_SIGMA) headw = Pf + (Ly.branch ? p.L.osigma_w : p.L.sigma_w);
        if (Ly.epi == EPI_DIR) headw = Pf + (Ly.branch ? p.L.orgb_w : p.L.rgb_w);
        float part0 = 0.0f, part1 = 0.0f, part2 = 0.0f;
        for (int c = 0; c < ncol; c += 32) {
          uint32_t v[32];
          tmem_ld32(lane_taddr + (uint32_t)(Ly.acc_col + col0 + c), v);
          tmem_ld_wait();
          float t[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + c) + j4);
            t[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) + b.x;
            t[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + b.y;
            t[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + b.z;
            t[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + b.w;
          }
          if (act) {
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = fmaxf(t[j], t[j] * kLeaky);
          }
          if (Ly.epi == EPI_HIDDEN_SIGMA) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(headw + col0 + c) + j4);
              part0 = fmaf(t[4 * j4 + 0], w.x, part0);
              part0 = fmaf(t[4 * j4 + 1], w.y, part0);
              part0 = fmaf(t[4 * j4 + 2], w.z, part0);
              part0 = fmaf(t[4 * j4 + 3], w.w, part0);
            }
          } else if (Ly.epi == EPI_DIR) {
            const int hw = Ly.N;  // rgb head weights are [3][N]
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(headw + col0 + c) + j4);
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(headw + hw + col0 + c) + j4);
              const float4 w2 = __ldg(reinterpret_cast<const float4*>(headw + 2 * hw + col0 + c) + j4);
              part0 = fmaf(t[4 * j4 + 0], w0.x, part0); part0 = fmaf(t[4 * j4 + 1], w0.y, part0);
              part0 = fmaf(t[4 * j4 + 2], w0.z, part0); part0 = fmaf(t[4 * j4 + 3], w0.w, part0);
              part1 = fmaf(t[4 * j4 + 0], w1.x, part1); part1 = fmaf(t[4 * j4 + 1], w1.y, part1);
              part1 = fmaf(t[4 * j4 + 2], w1.z, part1); part1 = fmaf(t[4 * j4 + 3], w1.w, part1);
              part2 = fmaf(t[4 * j4 + 0], w2.x, part2); part2 = fmaf(t[4 * j4 + 1], w2.y, part2);
              part2 = fmaf(t[4 * j4 + 2], w2.z, part2); part2 = fmaf(t[4 * j4 + 3], w2.w, part2);
            }
          }
          if (to_h) {
            const int chunk0 = (col0 + c) >> 3;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd)
              st_chunk(a_chunk_addr(sH, row, chunk0 + qd), pack_bf16(t[8 * qd + 0], t[8 * qd + 1]),
                       pack_bf16(t[8 * qd + 2], t[8 * qd + 3]), pack_bf16(t[8 * qd + 4], t[8 * qd + 5]),
                       pack_bf16(t[8 * qd + 6], t[8 * qd + 7]));
          }
        }
        if (Ly.epi == EPI_HIDDEN_SIGMA) sigma_part = part0;
        if (Ly.epi == EPI_DIR) {
          // combine the two column halves of this row through shared memory, finish the heads, write out
          float* sc = scratch + (row * 2 + hf) * 4;
          sc[0] = sigma_part; sc[1] = part0; sc[2] = part1; sc[3] = part2;
          asm volatile("bar.sync 1, %0;" ::"n"(NUM_COMPUTE) : "memory");
          if (hf == 0 && live) {
            const float* o = scratch + (row * 2 + 1) * 4;
            const float* hb = Pf + (Ly.branch ? p.L.orgb_b : p.L.rgb_b);
            float sg = sigma_part + o[0] + __ldg(Pf + (Ly.branch ? p.L.osigma_b : p.L.sigma_b));
            const float r = 1.0f / (1.0f + __expf(-(part0 + o[1] + __ldg(hb + 0))));
            const float gch = 1.0f / (1.0f + __expf(-(part1 + o[2] + __ldg(hb + 1))));
            const float b = 1.0f / (1.0f + __expf(-(part2 + o[3] + __ldg(hb + 2))));
            if (mute & (Ly.branch ? 2 : 1)) sg = -1e5f;
            float* outp = Ly.branch ? p.obj_out : p.scene_out;
            reinterpret_cast<float4*>(outp)[(int64_t)ray * p.out_stride + si] = make_float4(r, gch, b, sg);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(NUM_COMPUTE) : "memory");  // scratch reusable
        }
        // this layer's accumulator is drained and (if any) H is written: release the MMA warp.
        // The last layer of a tile releases the first layer of the next tile together with the X arrive.
        if (l + 1 < P.n_layers) {
          tc_fence_before();
          fence_async_smem();
          mbar_arrive(bar_a_ready);
        } else {
          tc_fence_before();
        }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int onerf_launch_field_bf16(onerf_ctx* ctx, const FieldParams& fp, cudaStream_t stream) {
  const PackLayout& L = fp.L;
  TcParams P;
  memset(&P, 0, sizeof(P));
  P.f = fp;
  const int xs = L.KX / 32, xo = L.KO / 32;
  int n = 0;
  auto add = [&](int gemm, int nx, int nh, int epi, int branch, int rc_base, int acc_col) {
    TcLayer& t = P.layers[n++];
    t.N = L.g[gemm].N; t.nslab_x = nx; t.nslab_h = nh; t.epi = epi; t.branch = branch; t.rc_base = rc_base;
    t.acc_col = acc_col; t.img_off = L.g[gemm].img_off; t.bias_off = L.g[gemm].bias_off;
  };
  if (fp.want_scene) {
    add(G_S0, xs, 0, EPI_HIDDEN, 0, 0, 0);
    add(G_S1, 0, 8, EPI_HIDDEN, 0, 0, 0);
    add(G_S2, 0, 8, EPI_HIDDEN, 0, 0, 0);
    add(G_S3, 0, 8, EPI_HIDDEN, 0, 0, 0);
    add(G_S4, xs, 8, EPI_HIDDEN, 0, 0, 0);
    add(G_S5, 0, 8, EPI_HIDDEN, 0, 0, 0);
    add(G_S6, 0, 8, EPI_HIDDEN, 0, 0, 0);
    add(G_S7, 0, 8, EPI_HIDDEN_SIGMA, 0, 0, 0);
    add(G_SFIN, 0, 8, EPI_FINAL, 0, 0, 0);
    add(G_SDIR, 0, 8, EPI_DIR, 0, RC_SDIR, 0);
  }
  if (fp.want_object) {
    add(G_O0, xo, 0, EPI_HIDDEN_RC, 1, RC_OL0, 256);
    add(G_O1, 0, 4, EPI_HIDDEN, 1, 0, 256);
    add(G_O2, xo, 4, EPI_HIDDEN_RC, 1, RC_OL2, 256);
    add(G_O3, 0, 4, EPI_HIDDEN_SIGMA, 1, 0, 256);
    add(G_OFIN, 0, 4, EPI_FINAL, 1, 0, 256);
    add(G_ODIR, 0, 4, EPI_DIR, 1, RC_ODIR, 256);
  }
  P.n_layers = n;
  P.x_atoms = L.use_voxel ? 6 : 1;
  const int64_t total = (int64_t)fp.n_rays * fp.S;
  const int64_t tiles = (total + TM - 1) / TM;
  const int blocks = (int)(tiles < ctx->num_sms ? tiles : ctx->num_sms);
  const size_t smem = 1024 + (size_t)(P.x_atoms + 4) * ATOM_BYTES + NSTAGE * STAGE_BYTES + TM * 2 * 4 * 4 + 256;
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
