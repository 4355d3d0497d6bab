// C-ABI plumbing: context, errors, and the field entry point's argument validation / dispatch.
#include <stdarg.h>
#include <string.h>

#include "field_common.cuh"

static thread_local char g_err[512] = "";

void onerf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int onerf_abi_version(void) { return ONERF_ABI_VERSION; }
extern "C" const char* onerf_last_error(void) { return g_err; }

extern "C" int onerf_ctx_create(int device, onerf_ctx** out) {
  ONERF_CHECK_ARG(out, "null out pointer");
  int count = 0;
  ONERF_CUDA(cudaGetDeviceCount(&count));
  ONERF_CHECK_ARG(device >= 0 && device < count, "no such CUDA device");
  cudaDeviceProp prop;
  ONERF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    onerf_set_error("onerf_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only (no fallback)",
                    device, prop.major, prop.minor);
    return ONERF_ERR_UNSUPPORTED;
  }
  ONERF_CUDA(cudaSetDevice(device));
  onerf_ctx* c = new onerf_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->launches = 0;
  *out = c;
  return ONERF_OK;
}

extern "C" int onerf_ctx_destroy(onerf_ctx* ctx) {
  delete ctx;
  return ONERF_OK;
}

extern "C" int64_t onerf_ctx_launch_count(const onerf_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int onerf_field_fwd(onerf_ctx* ctx, const onerf_field_args* a, void* stream_) {
  ONERF_CHECK_ARG(ctx && a, "null argument");
  ONERF_CHECK_ARG(a->rays && a->z && a->packed && a->ray_const, "null buffer");
  ONERF_CHECK_ARG(a->n_rays >= 0 && a->n_samples >= 1, "bad shape");
  ONERF_CHECK_ARG(a->want_scene || a->want_object, "nothing to compute");
  if (a->want_scene) ONERF_CHECK_ARG(a->scene_out && onerf_aligned16(a->scene_out), "scene_out null or misaligned");
  if (a->want_object) {
    ONERF_CHECK_ARG(a->obj_out && onerf_aligned16(a->obj_out), "obj_out null or misaligned");
    ONERF_CHECK_ARG(a->codes || a->code_row, "object branch needs codes or code_row");
  }
  ONERF_CHECK_ARG(a->n_boxes == 0 || a->boxes, "n_boxes > 0 with null boxes");
  ONERF_CHECK_ARG(a->z_stride >= a->n_samples && a->out_stride >= a->n_samples, "bad strides");
  if (a->grid)
    ONERF_CHECK_ARG(a->grid->table && a->grid->idx_map && a->grid->voxel_offset && a->grid->voxel_size &&
                        a->grid->voxel_shape && onerf_aligned16(a->grid->table),
                    "null / misaligned grid buffer");
  if (a->n_rays == 0) return ONERF_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  FieldParams p;
  memset(&p, 0, sizeof(p));
  p.rays = a->rays; p.xyz = a->xyz; p.z = a->z; p.z_stride = a->z_stride;
  p.codes = a->codes; p.code_row = a->code_row;
  p.n_rays = a->n_rays; p.S = a->n_samples;
  if (a->grid) p.grid = *a->grid;
  p.packed = a->packed;
  p.L = onerf_make_layout(a->grid ? 1 : 0);
  p.want_scene = a->want_scene; p.want_object = a->want_object;
  p.mute_zero_rays = a->mute_zero_rays;
  p.boxes = a->boxes; p.n_boxes = a->n_boxes;
  p.scene_out = a->scene_out; p.obj_out = a->obj_out; p.out_stride = a->out_stride;
  p.ray_const = a->ray_const;
  if (a->activations) {
    ONERF_UNSUPPORTED(a->precision != ONERF_PREC_FP32, "activation dump is built for ONERF_PREC_FP32 only");
    ONERF_UNSUPPORTED(a->z_stride != a->n_samples || a->out_stride != a->n_samples, "activation dump needs dense z / outputs");
    for (int i = 0; i < 17; ++i) ONERF_CHECK_ARG(a->activations[i], "null activation matrix");
    p.dump_x = a->activations[0];
    for (int i = 0; i < 10; ++i) p.dump_s[i] = a->activations[1 + i];
    for (int i = 0; i < 6; ++i) p.dump_o[i] = a->activations[11 + i];
  }
  int rc = onerf_launch_ray_const(ctx, p, stream);
  if (rc != ONERF_OK) return rc;
  if (a->precision == ONERF_PREC_FP32) return onerf_launch_field_fp32(ctx, p, stream);
  if (a->precision == ONERF_PREC_BF16) return onerf_launch_field_bf16(ctx, p, stream);
  onerf_set_error("onerf_field_fwd: unknown precision %d", a->precision);
  return ONERF_ERR_BAD_ARG;
}
