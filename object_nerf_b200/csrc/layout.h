// Packed-weight layout shared by the pack kernels and both field kernels (host + device).
//
// One ObjectNeRF (reference models/nerf_model.py:18-95) is re-laid into one blob:
//   [fp32 section]  every GEMM layer as W^T [K][N] (K-major, kernel-K order), head vectors, biases and
//                   the per-ray-constant ("hoisted") column blocks
//   [bf16 section]  the tcgen05 stage images: per GEMM layer, per 32-wide K slab, an N x 32 bf16 tile
//                   in the UMMA K-major SWIZZLE_64B shared-memory layout, in program order
//
// Kernel-K order ("X layout").  The encoded input of a sample is kept as one vector X:
//   voxel model:  X[0..271)   = scene input  [PE6(scene voxel ftr 16) | PE10(xyz)]  (reference order)
//                 X[271]      = 0 (pad)   (scene GEMMs read X[0..288): their weights for 271..287 are 0)
//                 X[272..376) = PE6(object voxel ftr 8)
//                 X[376..384) = 0 (pad to a multiple of 32/64)
//   plain model:  X[0..63) = PE10(xyz), X[63] = 0
// The scene branch consumes X[0..KX), the object branch X[0..KO).  Terms that are constant along a ray
// (direction encoding into the two dir layers, object code into object layers 0 and 2) are hoisted
// into a per-ray vector (ONERF_RAY_CONST_FLOATS) that also carries those layers' biases.
#pragma once
#include <stdint.h>

#define ONERF_W 256
#define ONERF_IW 128
#define ONERF_NCODE 64
#define ONERF_NDIR 27

// ray_const layout (floats)
#define RC_SDIR 0     // 128: b_dir  + W_dir [:,256:283] . PE4(d)
#define RC_ODIR 128   // 64 : b_odir + W_odir[:,128:155] . PE4(d)
#define RC_OL0 192    // 128: b_o0   + W_o0[:, code cols] . code
#define RC_OL2 320    // 128: b_o2   + W_o2[:, code cols] . code

enum GemmId {
  G_S0 = 0, G_S1, G_S2, G_S3, G_S4, G_S5, G_S6, G_S7, G_SFIN, G_SDIR,
  G_O0, G_O1, G_O2, G_O3, G_OFIN, G_ODIR, G_COUNT
};

struct GemmDesc {
  int K;            // kernel K (rows of W^T), multiple of 32
  int N;            // outputs
  int64_t wt_off;   // fp32 W^T [K][N], float offset into the blob
  int64_t bias_off; // fp32 [N] (for hoisted layers this bias is folded into ray_const and unused)
  int64_t img_off;  // bf16 stage images, BYTE offset into the blob; K/32 images of N*64 bytes
};

struct PackLayout {
  int use_voxel;
  int KX, KO;       // scene / object widths of X (288/384 voxel, 64/64 plain)
  int n_obj_vox;    // 104 or 0
  GemmDesc g[G_COUNT];
  int64_t sigma_w, sigma_b;     // [256], [1]
  int64_t rgb_w, rgb_b;         // [3][128], [3]
  int64_t osigma_w, osigma_b;   // [128], [1]
  int64_t orgb_w, orgb_b;       // [3][64], [3]
  int64_t h_sdir, h_odir;       // [27][128], [27][64]   (K-major)
  int64_t h_ol0, h_ol2;         // [64][128] each
  int64_t b_sdir, b_odir, b_ol0, b_ol2;  // biases folded into ray_const
  int64_t fp32_floats;          // size of the fp32 section
  int64_t total_bytes;
};

static inline PackLayout onerf_make_layout(int use_voxel) {
  PackLayout L;
  L.use_voxel = use_voxel;
  L.KX = use_voxel ? 288 : 64;  // 272 rounded up to a 32-wide K slab (columns 272..287 carry zero scene weights)
  L.KO = use_voxel ? 384 : 64;
  L.n_obj_vox = use_voxel ? 104 : 0;
  const int K[G_COUNT] = {L.KX, 256, 256, 256, L.KX + 256, 256, 256, 256, 256, 256,
                          L.KO, 128, L.KO + 128, 128, 128, 128};
  const int N[G_COUNT] = {256, 256, 256, 256, 256, 256, 256, 256, 256, 128, 128, 128, 128, 128, 128, 64};
  int64_t f = 0;
  for (int i = 0; i < G_COUNT; ++i) {
    L.g[i].K = K[i];
    L.g[i].N = N[i];
    L.g[i].wt_off = f;
    f += (int64_t)K[i] * N[i];
    L.g[i].bias_off = f;
    f += N[i];
  }
  auto take = [&](int64_t n) { int64_t o = f; f += (n + 3) & ~3ll; return o; };
  L.sigma_w = take(256); L.sigma_b = take(1);
  L.rgb_w = take(3 * 128); L.rgb_b = take(3);
  L.osigma_w = take(128); L.osigma_b = take(1);
  L.orgb_w = take(3 * 64); L.orgb_b = take(3);
  L.h_sdir = take(27 * 128); L.h_odir = take(27 * 64);
  L.h_ol0 = take(64 * 128); L.h_ol2 = take(64 * 128);
  L.b_sdir = take(128); L.b_odir = take(64); L.b_ol0 = take(128); L.b_ol2 = take(128);
  L.fp32_floats = f;
  int64_t bytes = (f * 4 + 1023) & ~1023ll;
  for (int i = 0; i < G_COUNT; ++i) {
    L.g[i].img_off = bytes;
    bytes += (int64_t)(K[i] / 32) * N[i] * 64;
  }
  L.total_bytes = bytes;
  return L;
}
