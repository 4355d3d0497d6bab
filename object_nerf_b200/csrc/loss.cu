// TotalLoss of the training step and its gradient w.r.t. the rendered maps in two small kernels (SURVEY.md section 8f
// row 3).  Reference: models/losses.py:5-135 - five masked-MSE terms (color, depth, opacity, instance color, instance
// depth), each over the coarse and the fine maps; under autograd the reference runs ~120 elementwise / index / reduce
// kernels and several host syncs (`mask.sum() == 0`) for 2 048 rays.  Here: one reduction pass (counts and weighted
// squared-error sums, fp64 accumulators) and one pass that writes d(loss_sum)/d(map) for all ten maps and the loss values.
#include "common.cuh"

namespace {

enum { T_COLOR = 0, T_DEPTH, T_OPACITY, T_ICOLOR, T_IDEPTH, N_TERMS };
// workspace (doubles): [0..5) mask counts per term (elements of the masked mean), [5] number of targets > 0,
// [6..16) squared-error sums: term * 2 + (0 coarse / 1 fine)
enum { WS_COUNT = 0, WS_TPOS = 5, WS_SUM = 6, WS_DOUBLES = 16 };

struct LossParams {
  onerf_loss_args a;
};

__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) loss_reduce_kernel(LossParams P, double* __restrict__ ws) {
  const onerf_loss_args& a = P.a;
  double acc[WS_DOUBLES];
#pragma unroll
  for (int i = 0; i < WS_DOUBLES; ++i) acc[i] = 0.0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < a.n_rays; r += (int64_t)gridDim.x * blockDim.x) {
    const bool valid = a.valid_mask[r] != 0, inst = a.instance_mask[r] != 0;
    const float t = a.depths[r], w = a.instance_mask_weight[r];
    const bool tpos = t > 0.0f;
    const float tr = a.rgbs[3 * r], tg = a.rgbs[3 * r + 1], tb = a.rgbs[3 * r + 2];
    if (tpos) acc[WS_TPOS] += 1.0;
    if (valid) {
      acc[WS_COUNT + T_COLOR] += 3.0;
      acc[WS_COUNT + T_OPACITY] += 1.0;
      if (tpos) acc[WS_COUNT + T_DEPTH] += 1.0;
      if (inst) acc[WS_COUNT + T_ICOLOR] += 3.0;
      if (inst && tpos) acc[WS_COUNT + T_IDEPTH] += 1.0;
    }
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const onerf_loss_maps& m = f ? a.fine : a.coarse;
      if (f && !a.has_fine) break;
      if (!valid) continue;
      {
        const float e0 = m.rgb[3 * r] - tr, e1 = m.rgb[3 * r + 1] - tg, e2 = m.rgb[3 * r + 2] - tb;
        acc[WS_SUM + 2 * T_COLOR + f] += (double)(e0 * e0) + (double)(e1 * e1) + (double)(e2 * e2);
      }
      if (tpos) { const float e = m.depth[r] - t; acc[WS_SUM + 2 * T_DEPTH + f] += (double)(e * e); }
      { const float e = clamp01(m.opacity_instance[r]) - (inst ? 1.0f : 0.0f); acc[WS_SUM + 2 * T_OPACITY + f] += (double)(e * e * w); }
      if (inst) {
        const float e0 = m.rgb_instance[3 * r] - tr, e1 = m.rgb_instance[3 * r + 1] - tg, e2 = m.rgb_instance[3 * r + 2] - tb;
        acc[WS_SUM + 2 * T_ICOLOR + f] += (double)(e0 * e0 * w) + (double)(e1 * e1 * w) + (double)(e2 * e2 * w);
        if (tpos) { const float e = m.depth_instance[r] - t; acc[WS_SUM + 2 * T_IDEPTH + f] += (double)(e * e * w); }
      }
    }
  }
  __shared__ double sh[8][WS_DOUBLES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < WS_DOUBLES; ++i) {
    const double v = warp_sum(acc[i]);
    if (lane == 0) sh[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < WS_DOUBLES) {
    double v = 0.0;
    for (int w8 = 0; w8 < 8; ++w8) v += sh[w8][threadIdx.x];
    if (v != 0.0) atomicAdd(ws + threadIdx.x, v);
  }
}

// term present (the reference returns None otherwise): models/losses.py:13-14, :46-47, :51-52, :80-81
__device__ __forceinline__ bool term_present(const double* ws, int t) {
  switch (t) {
    case T_COLOR: return true;                                                   // never skipped (mean of an empty set = NaN)
    case T_DEPTH: return ws[WS_TPOS] > 0;                                        // skipped only if no target depth at all
    case T_OPACITY: return ws[WS_COUNT + T_OPACITY] > 0;
    case T_ICOLOR: return ws[WS_COUNT + T_ICOLOR] > 0;
    default: return ws[WS_TPOS] > 0 && ws[WS_COUNT + T_IDEPTH] > 0;
  }
}

__global__ void __launch_bounds__(256) loss_grad_kernel(LossParams P, const double* __restrict__ ws) {
  const onerf_loss_args& a = P.a;
  const float wt[N_TERMS] = {a.color_weight, a.depth_weight, a.opacity_weight, a.instance_color_weight, a.instance_depth_weight};
  float scale[N_TERMS];   // d(weighted mean)/d(squared error) = weight / count
  bool present[N_TERMS];
#pragma unroll
  for (int t = 0; t < N_TERMS; ++t) {
    present[t] = term_present(ws, t);
    scale[t] = present[t] ? (float)((double)wt[t] / ws[WS_COUNT + t]) : 0.0f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double total = 0.0;
    for (int t = 0; t < N_TERMS; ++t) {
      // mean over the mask in fp32 like torch (sum / count), coarse + fine, times the weight
      float v = 0.0f;
      if (present[t]) {
        v = (float)(ws[WS_SUM + 2 * t] / ws[WS_COUNT + t]);
        if (a.has_fine) v += (float)(ws[WS_SUM + 2 * t + 1] / ws[WS_COUNT + t]);
      }
      a.terms_out[t] = v;                     // unweighted, as the reference's loss_dict (:129-131)
      a.present_out[t] = present[t] ? 1 : 0;
      if (present[t]) total += (double)(wt[t] * v);
    }
    *a.loss_sum_out = (float)total;
  }
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < a.n_rays; r += (int64_t)gridDim.x * blockDim.x) {
    const bool valid = a.valid_mask[r] != 0, inst = a.instance_mask[r] != 0;
    const float t = a.depths[r], w = a.instance_mask_weight[r];
    const bool tpos = t > 0.0f;
    const float tr = a.rgbs[3 * r], tg = a.rgbs[3 * r + 1], tb = a.rgbs[3 * r + 2];
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      if (f && !a.has_fine) break;
      const onerf_loss_maps& m = f ? a.fine : a.coarse;
      const onerf_loss_maps& g = f ? a.grad_fine : a.grad_coarse;
      float gc[3] = {0.f, 0.f, 0.f}, gi[3] = {0.f, 0.f, 0.f}, gd = 0.f, go = 0.f, gid = 0.f;
      if (valid) {
        gc[0] = 2.0f * (m.rgb[3 * r] - tr) * scale[T_COLOR];
        gc[1] = 2.0f * (m.rgb[3 * r + 1] - tg) * scale[T_COLOR];
        gc[2] = 2.0f * (m.rgb[3 * r + 2] - tb) * scale[T_COLOR];
        if (tpos) gd = 2.0f * (m.depth[r] - t) * scale[T_DEPTH];
        const float o = m.opacity_instance[r];
        if (o >= 0.0f && o <= 1.0f) go = 2.0f * (o - (inst ? 1.0f : 0.0f)) * w * scale[T_OPACITY];   // clamp backward
        if (inst) {
          gi[0] = 2.0f * (m.rgb_instance[3 * r] - tr) * w * scale[T_ICOLOR];
          gi[1] = 2.0f * (m.rgb_instance[3 * r + 1] - tg) * w * scale[T_ICOLOR];
          gi[2] = 2.0f * (m.rgb_instance[3 * r + 2] - tb) * w * scale[T_ICOLOR];
          if (tpos) gid = 2.0f * (m.depth_instance[r] - t) * w * scale[T_IDEPTH];
        }
      }
      float* grgb = const_cast<float*>(g.rgb);
      float* girgb = const_cast<float*>(g.rgb_instance);
      grgb[3 * r] = gc[0]; grgb[3 * r + 1] = gc[1]; grgb[3 * r + 2] = gc[2];
      girgb[3 * r] = gi[0]; girgb[3 * r + 1] = gi[1]; girgb[3 * r + 2] = gi[2];
      const_cast<float*>(g.depth)[r] = gd;
      const_cast<float*>(g.opacity_instance)[r] = go;
      const_cast<float*>(g.depth_instance)[r] = gid;
    }
  }
}

bool maps_ok(const onerf_loss_maps& m) { return m.rgb && m.depth && m.opacity_instance && m.rgb_instance && m.depth_instance; }

}  // namespace

extern "C" size_t onerf_total_loss_workspace_bytes(void) { return WS_DOUBLES * sizeof(double); }

extern "C" int onerf_total_loss(onerf_ctx* ctx, const onerf_loss_args* a, void* stream_) {
  ONERF_CHECK_ARG(ctx && a, "null argument");
  ONERF_CHECK_ARG(a->n_rays > 0, "n_rays must be positive");
  ONERF_CHECK_ARG(a->rgbs && a->depths && a->valid_mask && a->instance_mask && a->instance_mask_weight, "null batch buffer");
  ONERF_CHECK_ARG(maps_ok(a->coarse) && maps_ok(a->grad_coarse), "null coarse map / gradient");
  ONERF_CHECK_ARG(!a->has_fine || (maps_ok(a->fine) && maps_ok(a->grad_fine)), "null fine map / gradient");
  ONERF_CHECK_ARG(a->loss_sum_out && a->terms_out && a->present_out && a->workspace, "null output / workspace");
  ONERF_CHECK_ARG((reinterpret_cast<uintptr_t>(a->workspace) & 7u) == 0, "workspace must be 8-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  double* ws = reinterpret_cast<double*>(a->workspace);
  ONERF_CUDA(cudaMemsetAsync(ws, 0, WS_DOUBLES * sizeof(double), stream));
  LossParams P;
  P.a = *a;
  const int64_t want = (a->n_rays + 255) / 256;
  const int grid = (int)(want < (int64_t)ctx->num_sms * 4 ? want : (int64_t)ctx->num_sms * 4);
  loss_reduce_kernel<<<grid, 256, 0, stream>>>(P, ws);
  ONERF_LAUNCH_CHECK(ctx);
  loss_grad_kernel<<<grid, 256, 0, stream>>>(P, ws);
  ONERF_LAUNCH_CHECK(ctx);
  return ONERF_OK;
}
