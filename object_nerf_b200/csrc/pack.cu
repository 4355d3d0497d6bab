// Weight re-layout: 20 nn.Linear (W[out,in], b[out]) fp32 tensors -> the packed blob of layout.h, in ONE kernel
// launch driven by a per-layout job table (built once per context, kept in device memory).
// Reference layer shapes / concat orders: models/nerf_model.py:41-58 (scene), :77-95 (object),
// :105 and :138 (skip concat puts the INPUT first), :116 and :147 (dir concat puts it LAST),
// :130 (object input = [emb_xyz | obj_voxel | obj_code]).
// onerf_unpack_grads is the inverse map for the tensor-core backward: kernel-layout weight gradients -> the
// reference's [out,in] gradient tensors.
#include "common.cuh"
#include "layout.h"

#include <cuda_bf16.h>

#include <vector>

namespace {

struct Seg {
  int dst, len, src;
};
struct ColMap {
  Seg s[3];
};

__host__ __device__ __forceinline__ int map_col(const ColMap& m, int k) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (k >= m.s[i].dst && k < m.s[i].dst + m.s[i].len) return m.s[i].src + (k - m.s[i].dst);
  return -1;
}

enum JobType {
  JOB_GEMM = 0,    // forward layouts of one GEMM layer: fp32 W^T [Kd][N] and bf16 images [Kd/32][N x 32]
  JOB_VEC = 1,     // plain copy of n floats
  JOB_TIMG = 2,    // transposed bf16 images: rows r < N (mapped to source columns), K = Kd source rows (outputs)
};

struct PackJob {
  int type;
  int src;          // index into the 40 source pointers: 0..19 weights, 20..39 biases
  int src_ld;
  int N, Kd;
  ColMap map;       // JOB_GEMM: kernel-K column -> source column; JOB_TIMG: image row -> source column
  int64_t dst_f;    // float offset of the fp32 destination (-1: none)
  int64_t dst_b;    // byte offset of the bf16 image destination (-1: none)
};

struct SrcPtrs {
  const float* p[2 * ONERF_N_LINEAR];
};

// byte offset of element (row n, k) inside the stack of [rows x 32] SWIZZLE_64B images (rows * 64 B per image)
__device__ __forceinline__ int64_t img_byte(int rows, int n, int k) {
  const int s = k >> 5, kk = k & 31;
  return (int64_t)s * rows * 64 + (int64_t)n * 64 + ((((kk >> 3) ^ ((n >> 1) & 3))) << 4) + (kk & 7) * 2;
}

__global__ void __launch_bounds__(256) pack_all_kernel(SrcPtrs src, const PackJob* __restrict__ jobs, char* __restrict__ blob) {
  const PackJob j = jobs[blockIdx.y];
  const float* __restrict__ W = src.p[j.src];
  float* f = reinterpret_cast<float*>(blob);
  if (j.type == JOB_VEC) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.N; i += gridDim.x * blockDim.x) f[j.dst_f + i] = W[i];
    return;
  }
  const int total = j.Kd * j.N;
  if (j.type == JOB_GEMM) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
      const int n = e / j.Kd, k = e - n * j.Kd;  // consecutive threads walk k: coalesced reads of W rows
      const int c = map_col(j.map, k);
      const float v = (c >= 0) ? W[(int64_t)n * j.src_ld + c] : 0.0f;
      if (j.dst_f >= 0) f[j.dst_f + (int64_t)k * j.N + n] = v;
      if (j.dst_b >= 0) *reinterpret_cast<__nv_bfloat16*>(blob + j.dst_b + img_byte(j.N, n, k)) = __float2bfloat16_rn(v);
    }
  } else {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
      const int k = e / j.N, n = e - k * j.N;    // consecutive threads walk the image row = source column
      const int c = map_col(j.map, n);
      const float v = (c >= 0) ? W[(int64_t)k * j.src_ld + c] : 0.0f;
      *reinterpret_cast<__nv_bfloat16*>(blob + j.dst_b + img_byte(j.N, n, k)) = __float2bfloat16_rn(v);
    }
  }
}

// ---- per-GEMM description shared by pack and unpack ----
struct GemmSrc {
  int src_idx, src_ld;
  ColMap map;       // kernel-K column -> reference column
};

void gemm_sources(int use_voxel, const PackLayout& L, GemmSrc* g) {
  const int xin = use_voxel ? 271 : 63;      // width of the reference scene input
  const int ovx = L.n_obj_vox;               // 104 / 0
  const int oin = xin + ovx + ONERF_NCODE;   // 439 / 127
  const int xo = use_voxel ? 272 : 0;        // where the object voxel block starts in X (unused for plain)
  const ColMap ident256 = {{{0, 256, 0}, {0, 0, 0}, {0, 0, 0}}};
  const ColMap ident128 = {{{0, 128, 0}, {0, 0, 0}, {0, 0, 0}}};
  g[G_S0] = {0, xin, {{{0, xin, 0}, {0, 0, 0}, {0, 0, 0}}}};
  g[G_S1] = {1, 256, ident256};
  g[G_S2] = {2, 256, ident256};
  g[G_S3] = {3, 256, ident256};
  g[G_S4] = {4, xin + 256, {{{0, xin, 0}, {L.KX, 256, xin}, {0, 0, 0}}}};
  g[G_S5] = {5, 256, ident256};
  g[G_S6] = {6, 256, ident256};
  g[G_S7] = {7, 256, ident256};
  g[G_SFIN] = {9, 256, ident256};
  g[G_SDIR] = {10, 256 + 27, ident256};
  // object branch: X = [scene-in | pad | obj voxel | pad]
  g[G_O0] = {12, oin, {{{0, xin, 0}, {xo, ovx, xin}, {0, 0, 0}}}};
  g[G_O1] = {13, 128, ident128};
  g[G_O2] = {14, oin + 128, {{{0, xin, 0}, {xo, ovx, xin}, {L.KO, 128, oin}}}};
  g[G_O3] = {15, 128, ident128};
  g[G_OFIN] = {17, 128, ident128};
  g[G_ODIR] = {18, 128 + 27, ident128};
}

std::vector<PackJob> build_jobs(int use_voxel) {
  const PackLayout L = onerf_make_layout(use_voxel);
  GemmSrc gs[G_COUNT];
  gemm_sources(use_voxel, L, gs);
  const int xin = use_voxel ? 271 : 63, ovx = L.n_obj_vox, oin = xin + ovx + ONERF_NCODE;
  std::vector<PackJob> jobs;
  auto vec = [&](int src, int n, int64_t dst) { jobs.push_back(PackJob{JOB_VEC, src, 0, n, 1, ColMap{}, dst, -1}); };
  for (int i = 0; i < G_COUNT; ++i) {
    const GemmDesc& g = L.g[i];
    jobs.push_back(PackJob{JOB_GEMM, gs[i].src_idx, gs[i].src_ld, g.N, g.K, gs[i].map, g.wt_off, g.img_off});
    vec(ONERF_N_LINEAR + gs[i].src_idx, g.N, g.bias_off);
    if (g.hid_n > 0) {
      // transposed hidden block: image row n <-> kernel column hid_col0 + n <-> reference column
      const int ref0 = map_col(gs[i].map, g.hid_col0);
      jobs.push_back(PackJob{JOB_TIMG, gs[i].src_idx, gs[i].src_ld, g.hid_n, g.N, ColMap{{{0, g.hid_n, ref0}, {0, 0, 0}, {0, 0, 0}}},
                             -1, g.bimg_off});
    }
  }
  // X blocks of the four X-fed layers, transposed (input gradient w.r.t. the encoding)
  {
    const int xg[4] = {G_S0, G_S4, G_O0, G_O2};
    int64_t off = L.ximg_off;
    for (int q = 0; q < 4; ++q) {
      const int i = xg[q];
      ColMap m = gs[i].map;               // kernel column -> reference column, restricted to the X block
      const int kx = (q < 2) ? L.KX : L.KO;
      for (int t = 0; t < 3; ++t)
        if (m.s[t].dst >= kx) m.s[t] = Seg{0, 0, 0};
      jobs.push_back(PackJob{JOB_TIMG, gs[i].src_idx, gs[i].src_ld, ONERF_DX_N, L.g[i].N, m, -1, off});
      off += (int64_t)(L.g[i].N / 32) * ONERF_DX_N * 64;
    }
  }
  // heads
  vec(8, 256, L.sigma_w); vec(ONERF_N_LINEAR + 8, 1, L.sigma_b);
  vec(11, 3 * 128, L.rgb_w); vec(ONERF_N_LINEAR + 11, 3, L.rgb_b);
  vec(16, 128, L.osigma_w); vec(ONERF_N_LINEAR + 16, 1, L.osigma_b);
  vec(19, 3 * 64, L.orgb_w); vec(ONERF_N_LINEAR + 19, 3, L.orgb_b);
  // per-ray-constant blocks: fp32 W^T [len][N] of the hoisted columns + the layer's bias
  auto hoist = [&](int layer, int ld, int src0, int len, int N, int64_t dst, int64_t bias_dst) {
    jobs.push_back(PackJob{JOB_GEMM, layer, ld, N, len, ColMap{{{0, len, src0}, {0, 0, 0}, {0, 0, 0}}}, dst, -1});
    vec(ONERF_N_LINEAR + layer, N, bias_dst);
  };
  hoist(10, 256 + 27, 256, 27, 128, L.h_sdir, L.b_sdir);
  hoist(18, 128 + 27, 128, 27, 64, L.h_odir, L.b_odir);
  hoist(12, oin, xin + ovx, 64, 128, L.h_ol0, L.b_ol0);
  hoist(14, oin + 128, xin + ovx, 64, 128, L.h_ol2, L.b_ol2);
  return jobs;
}

// ------------------------------------------------------------------------------------------------
// unpack: kernel-layout gradients -> reference-layout gradients (+=)
//   dWk[g]: [N][K] fp32 (row = output, kernel-K columns), dbk[g]: [N]
// ------------------------------------------------------------------------------------------------
struct UnpackJob {
  int N, K, ref_ld;
  ColMap map;
  int64_t src_w, src_b;   // float offsets into the kernel-layout gradient buffer
  int dst;                // reference layer index (0..19)
};
struct DstPtrs {
  float* w[ONERF_N_LINEAR];
  float* b[ONERF_N_LINEAR];
};

__global__ void __launch_bounds__(256) unpack_kernel(const float* __restrict__ gk, const UnpackJob* __restrict__ jobs, DstPtrs dst) {
  const UnpackJob j = jobs[blockIdx.y];
  float* __restrict__ W = dst.w[j.dst];
  const int total = j.N * j.K;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int n = e / j.K, k = e - n * j.K;
    const int c = map_col(j.map, k);
    if (c >= 0) W[(int64_t)n * j.ref_ld + c] += gk[j.src_w + e];
  }
  if (blockIdx.x == 0)
    for (int n = threadIdx.x; n < j.N; n += blockDim.x) dst.b[j.dst][n] += gk[j.src_b + n];
}

struct JobTable {
  void* pack[2] = {nullptr, nullptr};
  int n_pack[2] = {0, 0};
  void* unpack[2] = {nullptr, nullptr};
};

}  // namespace

// kernel-layout gradient buffer of one model (fp32): per GEMM [N][K] then [N]; then the four heads
//   sigma_w[256] sigma_b[1] rgb_w[3*128] rgb_b[3] osigma_w[128] osigma_b[1] orgb_w[3*64] orgb_b[3]  (padded to 4)
GradLayout onerf_make_grad_layout(int use_voxel) {
  const PackLayout L = onerf_make_layout(use_voxel);
  GradLayout G;
  int64_t f = 0;
  for (int i = 0; i < G_COUNT; ++i) {
    G.w_off[i] = f;
    f += (int64_t)L.g[i].N * L.g[i].K;
    G.b_off[i] = f;
    f += L.g[i].N;
    f = (f + 3) & ~3ll;
  }
  auto take = [&](int64_t n) { int64_t o = f; f += (n + 3) & ~3ll; return o; };
  G.sigma_w = take(256); G.sigma_b = take(1); G.rgb_w = take(3 * 128); G.rgb_b = take(3);
  G.osigma_w = take(128); G.osigma_b = take(1); G.orgb_w = take(3 * 64); G.orgb_b = take(3);
  G.total_floats = f;
  return G;
}

static JobTable* tables(onerf_ctx* ctx) {
  if (!ctx->pack_tables) ctx->pack_tables = new JobTable();
  return reinterpret_cast<JobTable*>(ctx->pack_tables);
}

void onerf_free_pack_tables(onerf_ctx* ctx) {
  if (!ctx->pack_tables) return;
  JobTable* t = reinterpret_cast<JobTable*>(ctx->pack_tables);
  for (int v = 0; v < 2; ++v) {
    if (t->pack[v]) cudaFree(t->pack[v]);
    if (t->unpack[v]) cudaFree(t->unpack[v]);
  }
  delete t;
  ctx->pack_tables = nullptr;
}

extern "C" size_t onerf_packed_weights_bytes(int use_voxel) {
  return (size_t)onerf_make_layout(use_voxel ? 1 : 0).total_bytes;
}

extern "C" int onerf_pack_weights(onerf_ctx* ctx, int use_voxel, const float* const* W, const float* const* b,
                                  void* packed, size_t packed_bytes, void* stream_) {
  ONERF_CHECK_ARG(ctx && W && b && packed, "null argument");
  use_voxel = use_voxel ? 1 : 0;
  const PackLayout L = onerf_make_layout(use_voxel);
  if (packed_bytes < (size_t)L.total_bytes) {
    onerf_set_error("onerf_pack_weights: packed buffer too small (%zu < %lld)", packed_bytes, (long long)L.total_bytes);
    return ONERF_ERR_WORKSPACE;
  }
  ONERF_CHECK_ARG(onerf_aligned16(packed) && (reinterpret_cast<uintptr_t>(packed) & 1023u) == 0, "packed must be 1024-byte aligned");
  SrcPtrs src;
  for (int i = 0; i < ONERF_N_LINEAR; ++i) {
    ONERF_CHECK_ARG(W[i] && b[i], "null layer tensor");
    src.p[i] = W[i];
    src.p[ONERF_N_LINEAR + i] = b[i];
  }
  JobTable* t = tables(ctx);
  if (!t->pack[use_voxel]) {   // one-time: the job table of this layout, kept in device memory
    std::vector<PackJob> jobs = build_jobs(use_voxel);
    ONERF_CUDA(cudaMalloc(&t->pack[use_voxel], jobs.size() * sizeof(PackJob)));
    ONERF_CUDA(cudaMemcpy(t->pack[use_voxel], jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
    t->n_pack[use_voxel] = (int)jobs.size();
  }
  pack_all_kernel<<<dim3(48, t->n_pack[use_voxel]), 256, 0, (cudaStream_t)stream_>>>(
      src, reinterpret_cast<const PackJob*>(t->pack[use_voxel]), reinterpret_cast<char*>(packed));
  ONERF_LAUNCH_CHECK(ctx);
  return ONERF_OK;
}

extern "C" size_t onerf_grad_buffer_floats(int use_voxel) { return (size_t)onerf_make_grad_layout(use_voxel ? 1 : 0).total_floats; }

extern "C" int onerf_unpack_grads(onerf_ctx* ctx, int use_voxel, const float* grad_kernel_layout, float* const* dW,
                                  float* const* db, void* stream_) {
  ONERF_CHECK_ARG(ctx && grad_kernel_layout && dW && db, "null argument");
  use_voxel = use_voxel ? 1 : 0;
  const PackLayout L = onerf_make_layout(use_voxel);
  const GradLayout G = onerf_make_grad_layout(use_voxel);
  DstPtrs dst;
  for (int i = 0; i < ONERF_N_LINEAR; ++i) {
    ONERF_CHECK_ARG(dW[i] && db[i], "null gradient tensor");
    dst.w[i] = dW[i];
    dst.b[i] = db[i];
  }
  JobTable* t = tables(ctx);
  if (!t->unpack[use_voxel]) {
    GemmSrc gs[G_COUNT];
    gemm_sources(use_voxel, L, gs);
    std::vector<UnpackJob> jobs;
    for (int i = 0; i < G_COUNT; ++i)
      jobs.push_back(UnpackJob{L.g[i].N, L.g[i].K, gs[i].src_ld, gs[i].map, G.w_off[i], G.b_off[i], gs[i].src_idx});
    const ColMap id256 = {{{0, 256, 0}, {0, 0, 0}, {0, 0, 0}}}, id128 = {{{0, 128, 0}, {0, 0, 0}, {0, 0, 0}}},
                 id64 = {{{0, 64, 0}, {0, 0, 0}, {0, 0, 0}}};
    jobs.push_back(UnpackJob{1, 256, 256, id256, G.sigma_w, G.sigma_b, 8});
    jobs.push_back(UnpackJob{3, 128, 128, id128, G.rgb_w, G.rgb_b, 11});
    jobs.push_back(UnpackJob{1, 128, 128, id128, G.osigma_w, G.osigma_b, 16});
    jobs.push_back(UnpackJob{3, 64, 64, id64, G.orgb_w, G.orgb_b, 19});
    ONERF_CUDA(cudaMalloc(&t->unpack[use_voxel], jobs.size() * sizeof(UnpackJob)));
    ONERF_CUDA(cudaMemcpy(t->unpack[use_voxel], jobs.data(), jobs.size() * sizeof(UnpackJob), cudaMemcpyHostToDevice));
  }
  unpack_kernel<<<dim3(32, G_COUNT + 4), 256, 0, (cudaStream_t)stream_>>>(
      grad_kernel_layout, reinterpret_cast<const UnpackJob*>(t->unpack[use_voxel]), dst);
  ONERF_LAUNCH_CHECK(ctx);
  return ONERF_OK;
}
