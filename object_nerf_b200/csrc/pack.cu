// Weight re-layout: 20 nn.Linear (W[out,in], b[out]) fp32 tensors -> the packed blob of layout.h.
// Reference layer shapes / concat orders: models/nerf_model.py:41-58 (scene), :77-95 (object),
// :105 and :138 (skip concat puts the INPUT first), :116 and :147 (dir concat puts it LAST),
// :130 (object input = [emb_xyz | obj_voxel | obj_code]).
#include "common.cuh"
#include "layout.h"

#include <cuda_bf16.h>

namespace {

struct Seg {
  int dst, len, src;
};
struct ColMap {
  Seg s[3];
};

__device__ __forceinline__ int map_col(const ColMap& m, int k) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (k >= m.s[i].dst && k < m.s[i].dst + m.s[i].len) return m.s[i].src + (k - m.s[i].dst);
  return -1;
}

// dst_wt [Kd][N] fp32 (may be null), dst_img bf16 SW64 stage images (may be null)
__global__ void __launch_bounds__(256)
pack_gemm_kernel(const float* __restrict__ W, int src_ld, int N, int Kd, ColMap map, float* __restrict__ dst_wt,
                 __nv_bfloat16* __restrict__ dst_img) {
  const int total = Kd * N;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int n = e / Kd, k = e - n * Kd;  // consecutive threads walk k: coalesced reads of W rows
    const int c = map_col(map, k);
    const float v = (c >= 0) ? W[(int64_t)n * src_ld + c] : 0.0f;
    if (dst_wt) dst_wt[(int64_t)k * N + n] = v;
    if (dst_img) {
      const int s = k >> 5, kk = k & 31;
      const int64_t byte_off = (int64_t)s * N * 64 + (int64_t)n * 64 + ((((kk >> 3) ^ ((n >> 1) & 3))) << 4) + (kk & 7) * 2;
      dst_img[byte_off >> 1] = __float2bfloat16_rn(v);
    }
  }
}

__global__ void copy_vec_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace

extern "C" size_t onerf_packed_weights_bytes(int use_voxel) {
  return (size_t)onerf_make_layout(use_voxel ? 1 : 0).total_bytes;
}

extern "C" int onerf_pack_weights(onerf_ctx* ctx, int use_voxel, const float* const* W, const float* const* b,
                                  void* packed, size_t packed_bytes, void* stream_) {
  ONERF_CHECK_ARG(ctx && W && b && packed, "null argument");
  const PackLayout L = onerf_make_layout(use_voxel ? 1 : 0);
  if (packed_bytes < (size_t)L.total_bytes) {
    onerf_set_error("onerf_pack_weights: packed buffer too small (%zu < %lld)", packed_bytes, (long long)L.total_bytes);
    return ONERF_ERR_WORKSPACE;
  }
  ONERF_CHECK_ARG(onerf_aligned16(packed) && (reinterpret_cast<uintptr_t>(packed) & 1023u) == 0, "packed must be 1024-byte aligned");
  for (int i = 0; i < ONERF_N_LINEAR; ++i) ONERF_CHECK_ARG(W[i] && b[i], "null layer tensor");
  cudaStream_t stream = (cudaStream_t)stream_;
  float* f = reinterpret_cast<float*>(packed);
  char* bytes = reinterpret_cast<char*>(packed);

  const int xin = use_voxel ? 271 : 63;      // width of the reference scene input
  const int ovx = L.n_obj_vox;               // 104 / 0
  const int oin = xin + ovx + ONERF_NCODE;   // 439 / 127
  // source layer index (header order) and source leading dimension per GEMM
  const int src_idx[G_COUNT] = {0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 12, 13, 14, 15, 17, 18};
  const int src_ld[G_COUNT] = {xin, 256, 256, 256, xin + 256, 256, 256, 256, 256, 256 + 27,
                               oin, 128, oin + 128, 128, 128, 128 + 27};
  auto gemm = [&](int gi, ColMap m) -> int {
    const GemmDesc& g = L.g[gi];
    const int total = g.K * g.N;
    pack_gemm_kernel<<<(total + 255) / 256, 256, 0, stream>>>(
        W[src_idx[gi]], src_ld[gi], g.N, g.K, m, f + g.wt_off,
        reinterpret_cast<__nv_bfloat16*>(bytes + g.img_off));
    ONERF_LAUNCH_CHECK(ctx);
    copy_vec_kernel<<<1, 256, 0, stream>>>(b[src_idx[gi]], f + g.bias_off, g.N);
    ONERF_LAUNCH_CHECK(ctx);
    return ONERF_OK;
  };
  auto hoist = [&](int layer, int ld, int src0, int len, int N, int64_t dst, int64_t bias_dst) -> int {
    ColMap m = {{{0, len, src0}, {0, 0, 0}, {0, 0, 0}}};
    pack_gemm_kernel<<<(len * N + 255) / 256, 256, 0, stream>>>(W[layer], ld, N, len, m, f + dst, nullptr);
    ONERF_LAUNCH_CHECK(ctx);
    copy_vec_kernel<<<1, 256, 0, stream>>>(b[layer], f + bias_dst, N);
    ONERF_LAUNCH_CHECK(ctx);
    return ONERF_OK;
  };
  auto vec = [&](const float* src, int64_t dst, int n) -> int {
    copy_vec_kernel<<<1, 256, 0, stream>>>(src, f + dst, n);
    ONERF_LAUNCH_CHECK(ctx);
    return ONERF_OK;
  };
  const ColMap ident256 = {{{0, 256, 0}, {0, 0, 0}, {0, 0, 0}}};
  const ColMap ident128 = {{{0, 128, 0}, {0, 0, 0}, {0, 0, 0}}};
  int rc;
#define TRY(x) do { rc = (x); if (rc != ONERF_OK) return rc; } while (0)
  // scene branch
  TRY(gemm(G_S0, ColMap{{{0, xin, 0}, {0, 0, 0}, {0, 0, 0}}}));
  TRY(gemm(G_S1, ident256)); TRY(gemm(G_S2, ident256)); TRY(gemm(G_S3, ident256));
  TRY(gemm(G_S4, ColMap{{{0, xin, 0}, {L.KX, 256, xin}, {0, 0, 0}}}));
  TRY(gemm(G_S5, ident256)); TRY(gemm(G_S6, ident256)); TRY(gemm(G_S7, ident256));
  TRY(gemm(G_SFIN, ident256));
  TRY(gemm(G_SDIR, ident256));
  // object branch: X = [scene-in | pad | obj voxel | pad]
  const int xo = use_voxel ? 272 : 0;  // where the object voxel block starts in X (unused for plain)
  TRY(gemm(G_O0, ColMap{{{0, xin, 0}, {xo, ovx, xin}, {0, 0, 0}}}));
  TRY(gemm(G_O1, ident128));
  TRY(gemm(G_O2, ColMap{{{0, xin, 0}, {xo, ovx, xin}, {L.KO, 128, oin}}}));
  TRY(gemm(G_O3, ident128));
  TRY(gemm(G_OFIN, ident128));
  TRY(gemm(G_ODIR, ident128));
  // heads
  TRY(vec(W[8], L.sigma_w, 256)); TRY(vec(b[8], L.sigma_b, 1));
  TRY(vec(W[11], L.rgb_w, 3 * 128)); TRY(vec(b[11], L.rgb_b, 3));
  TRY(vec(W[16], L.osigma_w, 128)); TRY(vec(b[16], L.osigma_b, 1));
  TRY(vec(W[19], L.orgb_w, 3 * 64)); TRY(vec(b[19], L.orgb_b, 3));
  // per-ray-constant blocks
  TRY(hoist(10, 256 + 27, 256, 27, 128, L.h_sdir, L.b_sdir));
  TRY(hoist(18, 128 + 27, 128, 27, 64, L.h_odir, L.b_odir));
  TRY(hoist(12, oin, xin + ovx, 64, 128, L.h_ol0, L.b_ol0));
  TRY(hoist(14, oin + 128, xin + ovx, 64, 128, L.h_ol2, L.b_ol2));
#undef TRY
  return ONERF_OK;
}
