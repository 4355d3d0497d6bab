// Device-side encoding: sparse-voxel trilinear gather and positional encoding.
// Reference: models/embedding_helper.py:57-74 (Embedding.forward), :331-352 (sparse lookup),
// :354-411 (8-corner trilinear blend, channel split, PE), :325-329 (concat with PE10(xyz)).
#pragma once
#include "common.cuh"

struct GridView {
  const float* table;
  const int64_t* idx_map;
  float off[3];
  float vsize;
  int sx, sy, sz;
};

__device__ __forceinline__ GridView load_grid_view(const onerf_grid& g) {
  GridView v;
  v.table = g.table;
  v.idx_map = g.idx_map;
  v.off[0] = __ldg(g.voxel_offset + 0);
  v.off[1] = __ldg(g.voxel_offset + 1);
  v.off[2] = __ldg(g.voxel_offset + 2);
  v.vsize = __ldg(g.voxel_size);
  v.sx = (int)__ldg(g.voxel_shape + 0);
  v.sy = (int)__ldg(g.voxel_shape + 1);
  v.sz = (int)__ldg(g.voxel_shape + 2);
  return v;
}

// Trilinear blend of channels [C0, C0+NC) of the 8 surrounding voxel rows (NC multiple of 4).
// Empty (-1) or out-of-range corners contribute zero (embedding_helper.py:336-351); corner order and
// weight association follow :364-385.  EXACT = true keeps every product / sum individually rounded in
// the reference's order; false lets the compiler contract to FMA.
template <int C0, int NC, bool EXACT>
__device__ __forceinline__ void voxel_trilinear(const GridView& g, float x, float y, float z, float* out) {
  const float px = __fdiv_rn(__fadd_rn(x, g.off[0]), g.vsize);
  const float py = __fdiv_rn(__fadd_rn(y, g.off[1]), g.vsize);
  const float pz = __fdiv_rn(__fadd_rn(z, g.off[2]), g.vsize);
  const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
  const float u = __fsub_rn(px, fx), v = __fsub_rn(py, fy), w = __fsub_rn(pz, fz);
  const float lu = __fsub_rn(1.0f, u), lv = __fsub_rn(1.0f, v), lw = __fsub_rn(1.0f, w);
  // anything further than one cell outside the grid has no valid corner
  const bool any = (fx >= -1.0f) && (fy >= -1.0f) && (fz >= -1.0f) && (fx < (float)g.sx) &&
                   (fy < (float)g.sy) && (fz < (float)g.sz);
  const int qx = any ? (int)fx : -2, qy = any ? (int)fy : -2, qz = any ? (int)fz : -2;
#pragma unroll
  for (int c = 0; c < NC; ++c) out[c] = 0.0f;
#pragma unroll
  for (int corner = 0; corner < 8; ++corner) {
    const int cx = (corner >> 2) & 1, cy = (corner >> 1) & 1, cz = corner & 1;  // x-major product order
    const int ix = qx + cx, iy = qy + cy, iz = qz + cz;
    const bool ok = any && ix >= 0 && iy >= 0 && iz >= 0 && ix < g.sx && iy < g.sy && iz < g.sz;
    long long row = -1;
    if (ok) row = __ldg(g.idx_map + ((int64_t)ix * g.sy + iy) * g.sz + iz);
    const float wt = EXACT ? __fmul_rn(__fmul_rn(cx ? u : lu, cy ? v : lv), cz ? w : lw)
                           : (cx ? u : lu) * (cy ? v : lv) * (cz ? w : lw);
    if (row >= 0) {
      const float4* src = reinterpret_cast<const float4*>(g.table + row * 24 + C0);
#pragma unroll
      for (int q = 0; q < NC / 4; ++q) {
        const float4 f = __ldg(src + q);
        if (EXACT) {
          out[4 * q + 0] = __fadd_rn(out[4 * q + 0], __fmul_rn(f.x, wt));
          out[4 * q + 1] = __fadd_rn(out[4 * q + 1], __fmul_rn(f.y, wt));
          out[4 * q + 2] = __fadd_rn(out[4 * q + 2], __fmul_rn(f.z, wt));
          out[4 * q + 3] = __fadd_rn(out[4 * q + 3], __fmul_rn(f.w, wt));
        } else {
          out[4 * q + 0] += f.x * wt;
          out[4 * q + 1] += f.y * wt;
          out[4 * q + 2] += f.z * wt;
          out[4 * q + 3] += f.w * wt;
        }
      }
    }
  }
}
