/* Development hooks exported by libonerf_sm100.so next to the product ABI (include/onerf.h).  Not part of the drop-in
 * boundary: tools/ and tests/ use them to compare kernel variants, to read timing stamps and to check the host-built
 * tables of the two-tile field kernel.  None of them changes results. */
#ifndef ONERF_DEBUG_H
#define ONERF_DEBUG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* 1: every bf16 field launch uses the one-tile kernel (csrc/field_tc.cu); 0: default dispatch (two-tile kernel for the voxel
 * model).  Overrides the environment variable ONERF_TC_ONE_TILE.  tests/test_gpu_train_tc.py compares the two training dumps. */
void onerf_debug_force_one_tile(int on);

/* Device buffer of at least 1024 int64 for clock64() stamps / per-role wait statistics of block 0; null switches them off.
 * Only written by libraries built with `make TIMELINE=1` (one-tile kernel) or `make EXPERIMENT=ONERF_WAITSTATS` (two-tile). */
void onerf_debug_timeline(void* dev_buf);
void onerf_debug_timeline2(void* dev_buf);

/* The static program of the two-tile kernel as the launcher builds it (no GPU needed): slot records of the MMA warp, event
 * table of the epilogue warps, the weight producer's group lists.  Layout of `out`: csrc/field_tc2.cu; returns the number of
 * 32-bit words written, or a negative value (error code, or minus the capacity needed).  tests/test_two_tile_program_cpu.py. */
int onerf_debug_two_tile_program(int want_scene, int want_object, int train, uint32_t* out, int cap);

#ifdef __cplusplus
}
#endif
#endif
