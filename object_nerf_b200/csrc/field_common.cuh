// Parameters shared by the field kernels (FFMA and tcgen05) and their launchers.
#pragma once
#include "common.cuh"
#include "layout.h"

struct FieldParams {
  const float* rays;      // (N,8)
  const float* xyz;       // optional (N,S,3)
  const float* z;         // depth of sample i of ray r at z[r * z_stride + i]
  int64_t z_stride;
  const float* codes;     // (N,64) or null
  const float* code_row;  // (64,) or null
  int n_rays, S;
  onerf_grid grid;        // device pointers (unused for the plain-PE model)
  const void* packed;
  PackLayout L;
  int want_scene, want_object;
  int mute_zero_rays;
  const float* boxes;     // (n_boxes, 18)
  int n_boxes;
  float* scene_out;       // float4 (rgb, sigma) of sample i of ray r at [r * out_stride + i]
  float* obj_out;
  int64_t out_stride;
  float* ray_const;       // (N, ONERF_RAY_CONST_FLOATS)
  // backward support (FFMA kernel only): dump every layer's activations as [samples x width] row-major matrices
  //   dump_x: X (KO wide);  dump_s[0..7] scene hidden (256), [8] final (256), [9] dir (128);
  //   dump_o[0..3] object hidden (128), [4] final (128), [5] dir (64).  All null = no dump.
  float* dump_x;
  float* dump_s[10];
  float* dump_o[6];
  // training forward of the tensor-core kernel: bf16 activation atoms, X atoms and LeakyReLU sign masks of every tile
  // go to this workspace (layout.h: onerf_make_train_layout(use_voxel, n_rays * S)); null = inference
  void* train_ws;
};

// removed-object mask: inside any box <=> lo <= A p + t <= hi (inclusive), utils/bbox_utils.py:158-207
__device__ __forceinline__ bool point_in_boxes(const float* __restrict__ boxes, int n_boxes, float x, float y, float z) {
  bool inside = false;
  for (int b = 0; b < n_boxes; ++b) {
    const float* B = boxes + b * 18;
    bool in = true;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      // row-by-row dot product in the order of the reference's xyz @ A.T + t
      float v = __fmul_rn(x, __ldg(B + r * 3 + 0));
      v = __fadd_rn(v, __fmul_rn(y, __ldg(B + r * 3 + 1)));
      v = __fadd_rn(v, __fmul_rn(z, __ldg(B + r * 3 + 2)));
      v = __fadd_rn(v, __ldg(B + 9 + r));
      in = in && (v >= __ldg(B + 12 + r)) && (v <= __ldg(B + 15 + r));
    }
    inside = inside || in;
  }
  return inside;
}

int onerf_launch_ray_const(onerf_ctx* ctx, const FieldParams& p, cudaStream_t stream);
int onerf_launch_field_fp32(onerf_ctx* ctx, const FieldParams& p, cudaStream_t stream);
int onerf_launch_field_bf16(onerf_ctx* ctx, const FieldParams& p, cudaStream_t stream);
