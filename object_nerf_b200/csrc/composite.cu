// sigma -> alpha -> transmittance-weighted compositing, single-scene (scene + object branch) and the
// multi-object joint-sort variant.  One warp per ray, samples strided over lanes (coalesced float /
// float4 access), multiplicative warp scan for the exclusive transmittance product.
//
// Reference behaviour: models/rendering.py:139-229; render_tools/multi_rendering.py:96-157.
#include "common.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inclusive multiplicative warp scan
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}

__device__ __forceinline__ float alpha_from(float sigma, float delta) {
  // 1 - exp(-delta * relu(sigma))   (models/rendering.py:157)
  return __fsub_rn(1.0f, expf(__fmul_rn(-delta, fmaxf(sigma, 0.0f))));
}

struct Acc {
  float opacity, r, g, b, depth;
};

// Composite one branch of one ray.  field = (S,4) rgb,sigma.  Returns warp-reduced sums on all lanes.
// If w_out != nullptr the per-sample weights are stored.
__device__ __forceinline__ Acc composite_branch(const float* __restrict__ z, const float4* __restrict__ field,
                                                int S, float last_delta, float noise_std,
                                                const float* __restrict__ noise, uint64_t seed,
                                                uint32_t stream_id, int64_t ray, bool use_mask, float z_limit,
                                                float* __restrict__ w_out, int lane) {
  Acc acc = {0.f, 0.f, 0.f, 0.f, 0.f};
  float carry = 1.0f;  // prod_{j < chunk start} (1 - alpha_j + 1e-10)
  for (int base = 0; base < S; base += 32) {
    const int i = base + lane;
    float alpha = 0.0f, zi = 0.0f;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < S) {
      zi = __ldg(z + i);
      const float delta = (i + 1 < S) ? __fsub_rn(__ldg(z + i + 1), zi) : last_delta;
      f = __ldg(field + i);
      float s = f.w;
      if (noise_std > 0.0f) {
        const float nz = noise ? __ldg(noise + i) : philox_normal(seed, stream_id, (uint64_t)ray * S + i);
        s = __fadd_rn(s, __fmul_rn(nz, noise_std));
      }
      alpha = alpha_from(s, delta);
      if (use_mask && z_limit < zi) alpha = 0.0f;  // occlusion mask, models/rendering.py:192-202
    }
    const float t = (i < S) ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
    const float incl = warp_scan_mul(t, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float w = alpha * (carry * excl);
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    if (i < S) {
      if (w_out) w_out[i] = w;
      acc.opacity += w;
      acc.r += w * f.x;
      acc.g += w * f.y;
      acc.b += w * f.z;
      acc.depth += w * zi;
    }
  }
  acc.opacity = warp_sum(acc.opacity);
  acc.r = warp_sum(acc.r);
  acc.g = warp_sum(acc.g);
  acc.b = warp_sum(acc.b);
  acc.depth = warp_sum(acc.depth);
  return acc;
}

__global__ void __launch_bounds__(256) composite_kernel(onerf_composite_args a) {
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.n_samples;
  for (int r = blockIdx.x * warps_per_block + warp; r < a.n_rays; r += gridDim.x * warps_per_block) {
    const float* z = a.z + (int64_t)r * S;
    const bool obj_weights_out = (a.obj != nullptr) && a.rays_in_bbox;
    Acc sc = composite_branch(z, reinterpret_cast<const float4*>(a.scene) + (int64_t)r * S, S,
                              a.zero_last_delta ? 0.0f : 1e10f, a.noise_std,
                              a.noise_scene ? a.noise_scene + (int64_t)r * S : nullptr, a.seed, 2u, r, false,
                              0.0f, obj_weights_out ? nullptr : a.weights + (int64_t)r * S, lane);
    if (lane == 0) {
      a.opacity[r] = sc.opacity;
      a.depth[r] = sc.depth;
      // white background: rgb + 1 - opacity, models/rendering.py:178-179
      a.rgb[r * 3 + 0] = a.white_back ? __fadd_rn(__fadd_rn(sc.r, 1.0f), -sc.opacity) : sc.r;
      a.rgb[r * 3 + 1] = a.white_back ? __fadd_rn(__fadd_rn(sc.g, 1.0f), -sc.opacity) : sc.g;
      a.rgb[r * 3 + 2] = a.white_back ? __fadd_rn(__fadd_rn(sc.b, 1.0f), -sc.opacity) : sc.b;
    }
    if (a.obj != nullptr) {
      bool use_mask = (!a.is_eval) && (a.frustum_bound_th > 0.0f);
      if (use_mask && a.pass_through_mask && a.pass_through_mask[r]) use_mask = false;
      const float z_limit = __fadd_rn(sc.depth, a.frustum_bound_th);
      Acc ob = composite_branch(z, reinterpret_cast<const float4*>(a.obj) + (int64_t)r * S, S, 0.0f,
                                a.noise_std, a.noise_obj ? a.noise_obj + (int64_t)r * S : nullptr, a.seed, 3u,
                                r, use_mask, z_limit, obj_weights_out ? a.weights + (int64_t)r * S : nullptr,
                                lane);
      if (lane == 0) {
        a.opacity_instance[r] = ob.opacity;
        a.depth_instance[r] = ob.depth;
        // always composited on white, models/rendering.py:223
        a.rgb_instance[r * 3 + 0] = __fadd_rn(__fadd_rn(ob.r, 1.0f), -ob.opacity);
        a.rgb_instance[r * 3 + 1] = __fadd_rn(__fadd_rn(ob.g, 1.0f), -ob.opacity);
        a.rgb_instance[r * 3 + 2] = __fadd_rn(__fadd_rn(ob.b, 1.0f), -ob.opacity);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// multi-object: joint stable sort by depth, then composite (last delta = 0)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_order_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// One warp per ray.  Shared memory per warp: keys[P] (uint64: orderable z << 32 | concat index).
__global__ void __launch_bounds__(128)
composite_multi_kernel(const float* __restrict__ z_all, const float4* __restrict__ field_all, int n_rays,
                       int n_obj, int S, int P, int white_back, float* __restrict__ z_sorted,
                       float* __restrict__ weights, float* __restrict__ obj_ids,
                       float* __restrict__ weights_unsorted, float* __restrict__ opacity,
                       float* __restrict__ rgb, float* __restrict__ depth) {
  extern __shared__ unsigned long long keys_all[];
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* keys = keys_all + (size_t)warp * P;
  const int T = n_obj * S;
  for (int r = blockIdx.x * warps_per_block + warp; r < n_rays; r += gridDim.x * warps_per_block) {
    // concatenated index c = obj * S + s  <->  object-major storage [obj][ray][s]
    const int64_t obj_stride = (int64_t)n_rays * S;
    const float* z = z_all + (int64_t)r * S;
    const float4* fld = field_all + (int64_t)r * S;
#define SRC_OFF(c) ((int64_t)((c) / S) * obj_stride + ((c) % S))
    for (int i = lane; i < P; i += 32)
      keys[i] = (i < T) ? (((unsigned long long)float_order_key(__ldg(z + SRC_OFF(i))) << 32) | (unsigned)i)
                        : 0xffffffffffffffffull;
    __syncwarp();
    for (int k2 = 2; k2 <= P; k2 <<= 1) {
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < (P >> 1); t += 32) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int l = i | j;
          const bool up = ((i & k2) == 0);
          const unsigned long long a = keys[i], b = keys[l];
          if ((a > b) == up) { keys[i] = b; keys[l] = a; }
        }
        __syncwarp();
      }
    }
    // composite in sorted order
    Acc acc = {0.f, 0.f, 0.f, 0.f, 0.f};
    float carry = 1.0f;
    for (int base = 0; base < T; base += 32) {
      const int i = base + lane;
      float alpha = 0.0f, zi = 0.0f;
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      int src = 0;
      if (i < T) {
        src = (int)(keys[i] & 0xffffffffu);
        zi = __ldg(z + SRC_OFF(src));
        const float zn = (i + 1 < T) ? __ldg(z + SRC_OFF((int)(keys[i + 1] & 0xffffffffu))) : zi;
        const float delta = (i + 1 < T) ? __fsub_rn(zn, zi) : 0.0f;  // multi_rendering.py:125-128
        f = __ldg(fld + SRC_OFF(src));
        alpha = alpha_from(f.w, delta);
      }
      const float t = (i < T) ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
      const float incl = warp_scan_mul(t, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float w = alpha * (carry * excl);
      carry *= __shfl_sync(0xffffffffu, incl, 31);
      if (i < T) {
        const int64_t o = (int64_t)r * T + i;
        z_sorted[o] = zi;
        weights[o] = w;
        if (obj_ids) obj_ids[o] = (float)(src / S);
        if (weights_unsorted) weights_unsorted[(int64_t)r * S + SRC_OFF(src)] = w;
        acc.opacity += w;
        acc.r += w * f.x;
        acc.g += w * f.y;
        acc.b += w * f.z;
        acc.depth += w * zi;
      }
    }
    acc.opacity = warp_sum(acc.opacity);
    acc.r = warp_sum(acc.r);
    acc.g = warp_sum(acc.g);
    acc.b = warp_sum(acc.b);
    acc.depth = warp_sum(acc.depth);
    if (lane == 0) {
      opacity[r] = acc.opacity;
      depth[r] = acc.depth;
      rgb[r * 3 + 0] = white_back ? __fadd_rn(__fadd_rn(acc.r, 1.0f), -acc.opacity) : acc.r;
      rgb[r * 3 + 1] = white_back ? __fadd_rn(__fadd_rn(acc.g, 1.0f), -acc.opacity) : acc.g;
      rgb[r * 3 + 2] = white_back ? __fadd_rn(__fadd_rn(acc.b, 1.0f), -acc.opacity) : acc.b;
    }
    __syncwarp();
  }
#undef SRC_OFF
}

}  // namespace

extern "C" int onerf_composite(onerf_ctx* ctx, const onerf_composite_args* a, void* stream) {
  ONERF_CHECK_ARG(ctx && a, "null argument");
  ONERF_CHECK_ARG(a->z && a->scene && a->weights && a->opacity && a->rgb && a->depth, "null buffer");
  ONERF_CHECK_ARG(a->n_rays >= 0 && a->n_samples >= 1, "bad shape");
  ONERF_CHECK_ARG(onerf_aligned16(a->scene) && (!a->obj || onerf_aligned16(a->obj)), "field buffers must be 16-byte aligned");
  if (a->obj) ONERF_CHECK_ARG(a->rgb_instance && a->depth_instance && a->opacity_instance, "null instance output");
  if (a->n_rays == 0) return ONERF_OK;
  const int warps = 8;
  int blocks = (a->n_rays + warps - 1) / warps;
  const int cap = ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  composite_kernel<<<blocks, warps * 32, 0, (cudaStream_t)stream>>>(*a);
  ONERF_LAUNCH_CHECK(ctx);
  return ONERF_OK;
}

extern "C" int onerf_composite_multi(onerf_ctx* ctx, const float* z_all, const float* field_all, int n_rays,
                                     int n_obj, int n_samples, int white_back, float* z_sorted,
                                     float* weights, float* obj_ids, float* weights_unsorted, float* opacity,
                                     float* rgb, float* depth, void* stream) {
  ONERF_CHECK_ARG(ctx && z_all && field_all && z_sorted && weights && opacity && rgb && depth, "null argument");
  ONERF_CHECK_ARG(n_rays >= 0 && n_obj >= 1 && n_samples >= 1, "bad shape");
  ONERF_CHECK_ARG(onerf_aligned16(field_all), "field buffer must be 16-byte aligned");
  const int T = n_obj * n_samples;
  ONERF_UNSUPPORTED(T > 4096, "n_obj * n_samples > 4096");
  if (n_rays == 0) return ONERF_OK;
  int P = 2;
  while (P < T) P <<= 1;
  const int warps = 4;
  const size_t smem = (size_t)warps * P * sizeof(unsigned long long);
  ONERF_CUDA(cudaFuncSetAttribute(composite_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = (n_rays + warps - 1) / warps;
  const int cap = ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  composite_multi_kernel<<<blocks, warps * 32, smem, (cudaStream_t)stream>>>(
      z_all, reinterpret_cast<const float4*>(field_all), n_rays, n_obj, n_samples, P, white_back, z_sorted,
      weights, obj_ids, weights_unsorted, opacity, rgb, depth);
  ONERF_LAUNCH_CHECK(ctx);
  return ONERF_OK;
}
