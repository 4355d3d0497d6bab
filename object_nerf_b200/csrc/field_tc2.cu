// Two-tile tcgen05 implementation of the fused encode + two-branch MLP for sm_100a (voxel model, inference).
//
// The one-tile kernel (field_tc.cu) leaves the tensor pipe idle two thirds of the time: one 128-row tile has a single
// dependency chain  MMA(layer l) -> epilogue(l) -> MMA(l + 1)  and the epilogues of its warps cannot hide behind its own
// MMAs (profiles/r01_experiments.md).  Here a CTA owns TWO 128-row tiles (A, B) that walk the layer program together:
// while the 16 epilogue warps drain a layer half of one tile, the tensor pipe runs a layer half of the other.
//   slot order per two-half layer l :  (A,l,h0) (A,l,h1) (B,l,h0) (B,l,h1);  one-half layers (N <= 128): (A,l) (B,l)
//                                      object branch first, then the scene branch (see the launcher)
//   TMEM (512 columns)              :  two 128-column accumulators SHARED by the tiles (half h of a two-half layer uses
//                                      accumulator h, a one-half layer of tile t uses accumulator t), so consecutive slots
//                                      never wait for each other's epilogue; activations of tile t at [256 + 128 t, +128)
//                                      IN PLACE (bf16 pairs): the outputs of half 0 wait in 16 registers per thread until
//                                      the MMAs of half 1 have read the old activations
//   shared memory                   :  ONE 96 KB buffer XS for the encoded input X (the A operand of the four X-fed layers),
//                                      regenerated from 27 raw features per sample (24 trilinear voxel channels + xyz,
//                                      kept in shared memory for both tiles) each time a tile reaches an X-fed layer;
//                                      3 x 32 KB weight ring; per-row metadata; 2 KB head scratch
//   warps 0-15  epilogue of both tiles, alternating, AND the X productions (events of one program, ev_tab)
//   warp 16     weight producer (cp.async.bulk)     warp 17  tcgen05.mma issuer (owns TMEM)
//   warps 18-19 trilinear gather of the NEXT tile pair's raw features
// Registers are re-divided between the warpgroups with setmaxnreg (epilogue 88, the other four warps 128).
// mbarriers: full / empty (ring), acc_ready[a] (MMA -> epilogue), acc_free[a] (the epilogue has loaded accumulator a: the
// next MMAs may overwrite it), h_ready[t] (a layer's activations of tile t are written), xs_ready (XS holds the next X),
// f_ready[t] / f_free[t] (raw features of tile t: gather warps <-> epilogue warps).
// Arithmetic is that of the one-tile kernel (same K order, same epilogue math): results are bit-identical to it.
//
// Reference semantics: models/rendering.py:85-137, models/nerf_model.py:97-152,
// models/embedding_helper.py:325-411, render_tools/multi_rendering.py:16-93.
#include <cuda_bf16.h>
#include <cstdlib>

#include "encode.cuh"
#include "field_common.cuh"
#include "field_pe.cuh"

namespace {

using namespace tc;

constexpr int T2_NSTAGE = 3;
constexpr int T2_SLAB_BYTES = 8192;          // 128 rows x 64 B: one K-slab (32 of K) of one N = 128 half
constexpr int T2_STAGE_SLABS = 4;            // every ring stage costs the issuing warp ~500 cycles of waits / commit /
constexpr int T2_STAGE_BYTES = T2_STAGE_SLABS * T2_SLAB_BYTES;   // reconvergence whatever it holds: 8 MMAs per 32 KB stage
constexpr int T2_MAX_GROUPS = 8;
constexpr int T2_MAX_LAYERS = 16;
constexpr int T2_MAX_SLOTS = 56;
constexpr int T2_MAX_XUSE = 4;
constexpr int T2_MAX_EVENTS = T2_MAX_SLOTS + 2 * T2_MAX_XUSE;
constexpr int T2_EPI_THREADS = 512;
// warpgroups (setmaxnreg works per group of 4 warps): 0-3 epilogue, 4 = {producer, MMA, encode x 2}
constexpr int T2_PRODUCER_WARP = 16, T2_MMA_WARP = 17, T2_ENC_WARP0 = 18, T2_ENC_WARPS = 2;
constexpr int T2_THREADS = 32 * (T2_ENC_WARP0 + T2_ENC_WARPS);   // 640: 96 registers per thread at launch
// setmaxnreg moves registers inside the pool the CTA was LAUNCHED with (640 threads x 96), not the whole register file
#ifndef T2_REGS_EPI_V
#define T2_REGS_EPI_V 88
#define T2_REGS_CTRL_V 128
#endif
constexpr int T2_REGS_EPI = T2_REGS_EPI_V, T2_REGS_CTRL = T2_REGS_CTRL_V;   // 512 x EPI + 128 x CTRL = 61 440 = 640 x 96
constexpr int T2_ENC_ROWS = TM / (32 * T2_ENC_WARPS);             // rows of a tile each encode thread handles (2)
constexpr int T2_NF = 27;                     // raw features per sample: 24 trilinear channels, x, y, z
constexpr float kLeaky = 0.01f;

enum Epi { EPI_HIDDEN = 0, EPI_HIDDEN_RC = 1, EPI_HIDDEN_SIGMA = 2, EPI_FINAL = 3, EPI_DIR = 4, EV_XGEN = 5 };
enum SlotFlags { SLOT_WAIT_H = 1, SLOT_WAIT_XS = 2 };

struct T2Layer {
  int N;            // outputs of the layer
  int nhalf;        // 2: computed as two N = 128 halves; 1: N <= 128
  int nslab_x, nslab_h;
  int epi, branch, rc_base;
  int writes_h;     // the epilogue leaves activations for the next layer (everything except the dir layers)
  int act_slot;     // training dump: activation slot of the layer's outputs (layout.h: TrainLayout)
  int64_t img_off, bias_off;
  int ngroups, n_xgroups;
  int groups[T2_MAX_GROUPS];   // bits [0,5) first slab (inside X or H), [5,8) slab count (1..4), bit 8: from H
  // what the MMA warp adds to the operand base for slab i of group g (formed on the host: the issuing warp shares its
  // scheduler with four epilogue warps, every instruction it does not execute shortens the slot):
  //   X slab s: (s / 2) * (ATOM_BYTES / 16) + (s % 2) * 4   (16-byte units, added to the XS descriptor word)
  //   H slab s: 16 s                                         (TMEM columns, added to the tile's activation base)
  uint32_t a_rel[T2_MAX_GROUPS][T2_STAGE_SLABS];
};
struct T2Slot {
  uint8_t tile, layer, half, flags;   // flags: SlotFlags | accumulator index << 4
};
// Everything the MMA warp needs for one slot in 64 bytes (four independent 16-byte constant-bank loads, fetched one slot
// ahead): the per-slot walk through slots[] -> layers[] -> groups[] / a_rel[] was a chain of dependent loads in front of
// every slot.
//   q[0].x  instruction descriptor          q[0].y  half bytes / 16 | acc << 16 | tile << 17 | wait_h << 18 | wait_xs << 19 | ngroups << 20
//   q[0].z  4 bits per group: slab count | from_h << 3
//   q[1..3] 16-bit operand offsets, half-word 4 g + i = slab i of group g (same meaning as T2Layer::a_rel)
constexpr int T2_REC_GROUPS = 5;
struct T2Rec {
  uint4 q[4];
};
struct T2Params {
  FieldParams f;
  T2Rec rec[T2_MAX_SLOTS];
  T2Layer layers[T2_MAX_LAYERS];
  int n_layers;
  T2Slot slots[T2_MAX_SLOTS];
  int n_slots;
  int n_xuse;                       // X-fed layers per tile, in program order
  int xuse_full[T2_MAX_XUSE];       // 1: the layer reads all 384 columns (object branch), 0: the first 288 (scene)
  long long* timeline;              // -DONERF_TIMELINE: clock64() stamps of block 0, second tile pair (tools/timeline2.py)
  // what the epilogue of slot i needs, precomputed (one 16-byte constant-bank load per event instead of address arithmetic):
  //   x: accumulator column | activation column << 16      y: float offset of the bias row in the packed blob
  //   z: ray_const / head-weight column offset (floats)     w: SlotEpi flags | accumulator index << 16 | layer N << 20
  // An EV_XGEN event (the 16 epilogue warps write a tile's encoded input X into XS) has w = EV_XGEN | XgenFlags << 8 and
  //   x: accumulator whose completion says that XS is free (the last slot of the other tile's X-fed layer), 0xff: none
  uint4 ev_tab[T2_MAX_EVENTS];
  int n_events;
  // training forward (kernel template DUMP): where the event's activations go in the training workspace
  //   ev_dump[i] = (activation slot + 1) | (first mask word + 1) << 8   (0 in a field: nothing to store)
  // T2Rec header bit 28: the slot's X is also the tile's dump of X (bulk store by the MMA warp), bit 29: drain that copy
  // before the slot's accumulator is announced complete (the next production of X waits for exactly that)
  uint32_t ev_dump[T2_MAX_EVENTS];
  uint8_t* dump;
  TrainLayout TL;
};
enum SlotEpi { SE_KIND = 7, SE_TWO = 8, SE_H1 = 16, SE_TILE = 32, SE_BRANCH = 64, SE_N64 = 128 };
enum XgenFlags { XG_TILE = 1, XG_FULL = 2, XG_WAIT_F = 4, XG_RELEASE_F = 8 };

// -DONERF_WAITSTATS: block 0 accumulates the cycles each role spends in each kind of barrier wait (timeline[900 + k])
#ifdef ONERF_WAITSTATS
#define T2_WAIT(k, call) do { const long long w0_ = clock64(); call; wstat[k] += clock64() - w0_; } while (0)
#else
#define T2_WAIT(k, call) do { call; } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// epilogue math (same operations as field_tc.cu)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t leaky_bf16x2(uint32_t x) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&x);
  const __nv_bfloat162 slope = __floats2bfloat162_rn(kLeaky, kLeaky);
  v = __hmax2(v, __hmul2(v, slope));
  return *reinterpret_cast<uint32_t*>(&v);
}

// acc + bias -> bf16 (-> LeakyReLU on packed pairs): NC columns -> NC / 2 packed words
template <int NC, bool ACT, bool BIAS_GLOBAL>
__device__ __forceinline__ void math_hidden(const uint32_t* v, const float* bias, uint32_t* pk) {
#pragma unroll
  for (int j4 = 0; j4 < NC / 4; ++j4) {
    float4 b;
#ifdef T2_EXP_NO_BIAS
    b = make_float4(0.f, 0.f, 0.f, 0.f);
#else
    b = __ldg(reinterpret_cast<const float4*>(bias) + j4);
#endif
    uint32_t p0 = pack_bf16(__uint_as_float(v[4 * j4 + 0]) + b.x, __uint_as_float(v[4 * j4 + 1]) + b.y);
    uint32_t p1 = pack_bf16(__uint_as_float(v[4 * j4 + 2]) + b.z, __uint_as_float(v[4 * j4 + 3]) + b.w);
    if (ACT) { p0 = leaky_bf16x2(p0); p1 = leaky_bf16x2(p1); }
    pk[2 * j4] = p0;
    pk[2 * j4 + 1] = p1;
  }
}
// last hidden layer of a branch: plus the sigma head as an fp32 dot product on the un-rounded activations
template <int NC>
__device__ __forceinline__ float math_hidden_sigma(const uint32_t* v, const float* bias, const float* headw, uint32_t* pk) {
  float part = 0.0f;
#pragma unroll
  for (int j4 = 0; j4 < NC / 4; ++j4) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + j4);
    const float4 w = __ldg(reinterpret_cast<const float4*>(headw) + j4);
    float t0 = __uint_as_float(v[4 * j4 + 0]) + b.x, t1 = __uint_as_float(v[4 * j4 + 1]) + b.y;
    float t2 = __uint_as_float(v[4 * j4 + 2]) + b.z, t3 = __uint_as_float(v[4 * j4 + 3]) + b.w;
    t0 = fmaxf(t0, t0 * kLeaky); t1 = fmaxf(t1, t1 * kLeaky); t2 = fmaxf(t2, t2 * kLeaky); t3 = fmaxf(t3, t3 * kLeaky);
    part = fmaf(t0, w.x, part); part = fmaf(t1, w.y, part); part = fmaf(t2, w.z, part); part = fmaf(t3, w.w, part);
    pk[2 * j4] = pack_bf16(t0, t1);
    pk[2 * j4 + 1] = pack_bf16(t2, t3);
  }
  return part;
}
// direction layer: LeakyReLU(acc + per-ray constant) feeds the 3-wide rgb head directly (fp32 dots)
template <int NC>
__device__ __forceinline__ void math_dir(const uint32_t* v, const float* rcbias, const float* headw, int head_ld, float& p0,
                                         float& p1, float& p2) {
#pragma unroll
  for (int j4 = 0; j4 < NC / 4; ++j4) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(rcbias) + j4);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(headw) + j4);
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(headw + head_ld) + j4);
    const float4 w2 = __ldg(reinterpret_cast<const float4*>(headw + 2 * head_ld) + j4);
    float t0 = __uint_as_float(v[4 * j4 + 0]) + b.x, t1 = __uint_as_float(v[4 * j4 + 1]) + b.y;
    float t2 = __uint_as_float(v[4 * j4 + 2]) + b.z, t3 = __uint_as_float(v[4 * j4 + 3]) + b.w;
    t0 = fmaxf(t0, t0 * kLeaky); t1 = fmaxf(t1, t1 * kLeaky); t2 = fmaxf(t2, t2 * kLeaky); t3 = fmaxf(t3, t3 * kLeaky);
    p0 = fmaf(t0, w0.x, p0); p0 = fmaf(t1, w0.y, p0); p0 = fmaf(t2, w0.z, p0); p0 = fmaf(t3, w0.w, p0);
    p1 = fmaf(t0, w1.x, p1); p1 = fmaf(t1, w1.y, p1); p1 = fmaf(t2, w1.z, p1); p1 = fmaf(t3, w1.w, p1);
    p2 = fmaf(t0, w2.x, p2); p2 = fmaf(t1, w2.y, p2); p2 = fmaf(t2, w2.z, p2); p2 = fmaf(t3, w2.w, p2);
  }
}

// training forward: the same, and the activations packed for the dump (they feed the rgb head's weight gradient)
template <int NC>
__device__ __forceinline__ void math_dir_pk(const uint32_t* v, const float* rcbias, const float* headw, int head_ld, float& p0,
                                            float& p1, float& p2, uint32_t* pk) {
#pragma unroll
  for (int j4 = 0; j4 < NC / 4; ++j4) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(rcbias) + j4);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(headw) + j4);
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(headw + head_ld) + j4);
    const float4 w2 = __ldg(reinterpret_cast<const float4*>(headw + 2 * head_ld) + j4);
    float t0 = __uint_as_float(v[4 * j4 + 0]) + b.x, t1 = __uint_as_float(v[4 * j4 + 1]) + b.y;
    float t2 = __uint_as_float(v[4 * j4 + 2]) + b.z, t3 = __uint_as_float(v[4 * j4 + 3]) + b.w;
    t0 = fmaxf(t0, t0 * kLeaky); t1 = fmaxf(t1, t1 * kLeaky); t2 = fmaxf(t2, t2 * kLeaky); t3 = fmaxf(t3, t3 * kLeaky);
    p0 = fmaf(t0, w0.x, p0); p0 = fmaf(t1, w0.y, p0); p0 = fmaf(t2, w0.z, p0); p0 = fmaf(t3, w0.w, p0);
    p1 = fmaf(t0, w1.x, p1); p1 = fmaf(t1, w1.y, p1); p1 = fmaf(t2, w1.z, p1); p1 = fmaf(t3, w1.w, p1);
    p2 = fmaf(t0, w2.x, p2); p2 = fmaf(t1, w2.y, p2); p2 = fmaf(t2, w2.z, p2); p2 = fmaf(t3, w2.w, p2);
    pk[2 * j4] = pack_bf16(t0, t1);
    pk[2 * j4 + 1] = pack_bf16(t2, t3);
  }
}

// A non-dir layer half is 128 wide: 32 accumulator columns per thread, consumed in two batches of 16 (16 live registers
// instead of 32); after the last load the accumulator is handed back.  Returns the partial sigma dot product.
__device__ __forceinline__ float epi_batches(int epi, uint32_t acc_addr, const float* bias, const float* rcbias,
                                             const float* sigw, uint32_t bar_free, int lane, uint32_t (&out)[16]) {
  float part = 0.0f;
#pragma unroll
  for (int bt = 0; bt < 2; ++bt) {
    uint32_t v[16];
    tmem_ld16(acc_addr + 16 * bt, v);
    tmem_ld_wait();
    if (bt == 1) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free);
    }
    switch (epi) {
      case EPI_HIDDEN: math_hidden<16, true, false>(v, bias + 16 * bt, out + 8 * bt); break;
      case EPI_HIDDEN_RC: math_hidden<16, true, true>(v, rcbias + 16 * bt, out + 8 * bt); break;
      case EPI_FINAL: math_hidden<16, false, false>(v, bias + 16 * bt, out + 8 * bt); break;
      default: part += math_hidden_sigma<16>(v, bias + 16 * bt, sigw + 16 * bt, out + 8 * bt); break;   // EPI_HIDDEN_SIGMA
    }
  }
  return part;
}

// shared-memory carve-up (byte offsets from the 1024-byte aligned base)
constexpr uint32_t OFF_X = 0;
constexpr uint32_t OFF_B = OFF_X + 6 * ATOM_BYTES;                    // weight ring
constexpr uint32_t OFF_F = OFF_B + T2_NSTAGE * T2_STAGE_BYTES;        // [2][27][128] floats: raw features of both tiles
constexpr uint32_t OFF_META = OFF_F + 2 * T2_NF * 128 * 4;            // [2 parities][2 tiles][128] int2: ray, sample | flags
constexpr uint32_t OFF_SCRATCH = OFF_META + 2 * 2 * 128 * 8;          // [4][128] floats (head partial sums, one quantity at a time)
constexpr uint32_t OFF_BAR = OFF_SCRATCH + 4 * 128 * 4;
constexpr uint32_t T2_SMEM_BYTES = OFF_BAR + 256 + 1024;

// One epilogue event (slot `si` of the pair's program).  Live state across events: the 16-register stash, two partial
// sigma sums and the barrier phase bits; everything else comes from the slot table (constant bank) and the per-row
// metadata the encode warps left in shared memory.
template <bool DUMP>
__device__ __forceinline__ void epi_event(const T2Params& P, const uint4 e, uint32_t ed, int64_t tile_idx, int64_t n_tiles,
                                          uint8_t* smem, uint32_t sbase, uint32_t lane_taddr,
                                          int parity, uint32_t (&stash)[16], uint32_t& acc_bits, float& sigma_a, float& sigma_b
#ifdef ONERF_WAITSTATS
                                          , long long* wstat
#endif
                                          ) {
  const FieldParams& p = P.f;
  const int flags = (int)(e.w & 0xffffu), A = (int)((e.w >> 16) & 15u), N = (int)(e.w >> 20);
  const int kind = flags & SE_KIND, T = (flags & SE_TILE) ? 1 : 0, branch = (flags & SE_BRANCH) ? 1 : 0;
  const int lane = threadIdx.x & 31, cq = threadIdx.x >> 7, row = threadIdx.x & 127;
  const int ncol = (flags & SE_N64) ? 16 : 32;            // accumulator columns of this thread
  const uint32_t acc_addr = lane_taddr + (e.x & 0xffffu) + (uint32_t)(cq * ncol);
  const uint32_t h_addr = lane_taddr + (e.x >> 16);
  const uint32_t bar_acc_ready = sbase + OFF_BAR + 16 * T2_NSTAGE, bar_acc_free = bar_acc_ready + 16, bar_h_ready = bar_acc_free + 16;
  const float* Pf = reinterpret_cast<const float*>(p.packed);
  T2_WAIT(6, mbar_wait(bar_acc_ready + 8 * A, (acc_bits >> A) & 1u));
  acc_bits ^= 1u << A;
  tc_fence_after();
  // per-row metadata of this pair, written by the encode warps before the pair's first X was produced (read it only
  // after an accumulator of the pair is ready: that orders it after the gather)
  // (only the two event kinds with per-ray constants touch it: the common hidden-layer event does no address arithmetic)
  int2 meta = make_int2(0, 0);
  const float* rc = nullptr;
  if (kind == EPI_DIR || kind == EPI_HIDDEN_RC) {
    meta = *reinterpret_cast<const int2*>(smem + OFF_META + ((parity * 2 + T) * 128 + row) * 8);
    rc = p.ray_const + (int64_t)meta.x * ONERF_RAY_CONST_FLOATS + e.z + cq * ncol;   // per-ray constants of this thread's columns
  }
  uint32_t v[32];
#ifdef ONERF_WAITSTATS
  const long long ph0 = clock64();
#endif
  tmem_ld16(acc_addr, v);
  if (ncol == 32) tmem_ld16(acc_addr + 16, v + 16);
  tmem_ld_wait();
  // the accumulator is in registers: the next MMAs may overwrite it
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar_acc_free + 8 * A);
#ifdef ONERF_WAITSTATS
  const long long ph1 = clock64();
  wstat[0] += ph1 - ph0;
  if (kind == EPI_DIR) wstat[4] -= ph1; else if ((flags & SE_TWO) && !(flags & SE_H1)) wstat[1] -= ph1; else wstat[2] -= ph1;
#endif
  // training forward: this thread's row of the layer's activation slot and of its mask words
  DumpDst dd;
  dd.row = nullptr; dd.mask = nullptr; dd.swz = row & 7;
  const int half = (flags & SE_H1) ? 1 : 0, dword = half * 4 + cq, dcol = half * 128 + cq * ncol;
  if (DUMP && (ed & 0xffu) != 0u && tile_idx < n_tiles) {
    const int slot = (int)(ed & 0xffu) - 1, mw = (int)((ed >> 8) & 0xffu) - 1;
    dd.row = P.dump + P.TL.act_off[slot] + ((size_t)tile_idx * P.TL.act_atoms[slot]) * ATOM_BYTES + (size_t)row * 128;
    if (mw >= 0) dd.mask = reinterpret_cast<uint32_t*>(P.dump + P.TL.mask_off) + ((size_t)tile_idx * ONERF_MASK_WORDS + mw) * 128 + row;
  }
  if (kind == EPI_DIR) {
    const float* headw = Pf + (branch ? p.L.orgb_w : p.L.rgb_w) + cq * ncol;
    float part0 = 0.0f, part1 = 0.0f, part2 = 0.0f;
    if (DUMP) {
      uint32_t dpk[16];
      if (ncol == 32) { math_dir_pk<32>(v, rc, headw, N, part0, part1, part2, dpk); dump_packed<16>(dd, dcol, dword, dpk); }
      else { math_dir_pk<16>(v, rc, headw, N, part0, part1, part2, dpk); dump_packed<8>(dd, dcol, dword, dpk); }
    } else {
      if (ncol == 32) math_dir<32>(v, rc, headw, N, part0, part1, part2);
      else math_dir<16>(v, rc, headw, N, part0, part1, part2);
    }
    // combine the four column quarters of this row through shared memory: one float4 per thread, one exchange.  The
    // scratch is the last atom of XS (columns 320..383: only the object layers' X lives there, their MMAs completed long
    // ago, and the next object X is written after every warp has passed this event's barrier); the two tiles use
    // different halves, so the other tile's head event, which follows immediately, cannot overwrite what is being read.
    float4* scratch = reinterpret_cast<float4*>(smem + OFF_X + 5 * ATOM_BYTES) + T * 512;
    scratch[cq * 128 + row] = make_float4(T ? sigma_b : sigma_a, part0, part1, part2);
    asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    if (cq == 0) {
      const float4 q0 = scratch[row], q1 = scratch[128 + row], q2 = scratch[256 + row], q3 = scratch[384 + row];
      tot[0] = q0.x + q1.x + q2.x + q3.x;   // (same order of additions as before: bit-identical results)
      tot[1] = q0.y + q1.y + q2.y + q3.y;
      tot[2] = q0.z + q1.z + q2.z + q3.z;
      tot[3] = q0.w + q1.w + q2.w + q3.w;
    }
    // (the reads above are done before any warp can reach the next production of an object X, which rewrites this atom)
    asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");
    if (cq == 0 && (meta.y & (1 << 30))) {
      const float* hb = Pf + (branch ? p.L.orgb_b : p.L.rgb_b);
      float sg = tot[0] + __ldg(Pf + (branch ? p.L.osigma_b : p.L.sigma_b));
      const float r = 1.0f / (1.0f + __expf(-(tot[1] + __ldg(hb + 0))));
      const float gch = 1.0f / (1.0f + __expf(-(tot[2] + __ldg(hb + 1))));
      const float b = 1.0f / (1.0f + __expf(-(tot[3] + __ldg(hb + 2))));
      if ((meta.y >> 28) & (branch ? 2 : 1)) sg = -1e5f;
      float* outp = branch ? p.obj_out : p.scene_out;
      reinterpret_cast<float4*>(outp)[(int64_t)meta.x * p.out_stride + (meta.y & 0x0fffffff)] = make_float4(r, gch, b, sg);
    }
#ifdef ONERF_WAITSTATS
    wstat[4] += clock64();
#endif
    return;
  }
#ifdef T2_EXP_NO_EPI
  if (kind != EPI_DIR) {   // experiment: no epilogue math (garbage results), only the barrier protocol
    if ((flags & SE_TWO) && !(flags & SE_H1)) return;
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_h_ready + 8 * T);
    return;
  }
#endif
  // hidden / final layers: 32 columns per thread; biases come straight from the packed blob (14 KB per model: L1 hits)
  const float* bias = Pf + (int64_t)e.y + cq * 32;
  if ((flags & SE_TWO) && !(flags & SE_H1)) {
    // half 0 of a two-half layer: the outputs wait in registers until the MMAs of half 1 have read the old activations
    float part = 0.0f;
    switch (kind) {
      case EPI_HIDDEN: math_hidden<32, true, false>(v, bias, stash); break;
      case EPI_HIDDEN_RC: math_hidden<32, true, true>(v, rc, stash); break;
      case EPI_FINAL: math_hidden<32, false, false>(v, bias, stash); break;
      default: part = math_hidden_sigma<32>(v, bias, Pf + (branch ? p.L.osigma_w : p.L.sigma_w) + cq * 32, stash); break;
    }
    if (kind == EPI_HIDDEN_SIGMA) { if (T) sigma_b = part; else sigma_a = part; }
    if (DUMP) dump_packed<16>(dd, dcol, dword, stash);
#ifdef ONERF_WAITSTATS
    asm volatile("" ::"r"(stash[0]), "r"(stash[15]));
    wstat[1] += clock64();
#endif
    return;
  }
  // this half's accumulator was complete, so every MMA of the layer has finished reading the old activations: overwrite
  // them in place, half 0 first (its registers are free before this half's outputs are formed)
  uint32_t out_col = (uint32_t)(cq * 16);
  if (flags & SE_TWO) {
    tmem_st16(h_addr + out_col, stash);
    out_col += 64;
  }
  uint32_t pk[16];
  float part = 0.0f;
  switch (kind) {
    case EPI_HIDDEN: math_hidden<32, true, false>(v, bias, pk); break;
    case EPI_HIDDEN_RC: math_hidden<32, true, true>(v, rc, pk); break;
    case EPI_FINAL: math_hidden<32, false, false>(v, bias, pk); break;
    default:
      part = math_hidden_sigma<32>(v, bias, Pf + (branch ? p.L.osigma_w : p.L.sigma_w) + ((flags & SE_TWO) ? 128 : 0) + cq * 32, pk);
      break;
  }
  if (kind == EPI_HIDDEN_SIGMA) {
    if (flags & SE_TWO) part += T ? sigma_b : sigma_a;
    if (T) sigma_b = part; else sigma_a = part;
  }
#ifdef ONERF_WAITSTATS
  asm volatile("" ::"r"(pk[0]), "r"(pk[15]));
  const long long ph2 = clock64();
  wstat[2] += ph2;
#endif
  tmem_st16(h_addr + out_col, pk);
  if (DUMP) dump_packed<16>(dd, dcol, dword, pk);
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar_h_ready + 8 * T);
#ifdef ONERF_WAITSTATS
  wstat[3] += clock64() - ph2;
#endif
}

// Trilinear gather of one row's 27 raw features (24 voxel channels, x, y, z) and its mute flags.  Off the critical path: it
// prefetches the NEXT tile pair.  Inlined at ONE call site: a separately compiled function would not know the register
// budget setmaxnreg left to the gather warps.
// meta_out: {ray, sample index | live << 30 | mute bits << 28} for the epilogue's output stage.
__device__ __forceinline__ void gather_tile(const FieldParams& p, const GridView& g, float* F, int2* meta_out, int64_t e,
                                         int64_t total) {
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
  *meta_out = make_int2(ray, si | (live ? (1 << 30) : 0) | (mute << 28));
  float f[8];
  voxel_trilinear<0, 8, false>(g, x, y, z, f);
#pragma unroll
  for (int c = 0; c < 8; ++c) F[c * 128] = f[c];
  voxel_trilinear<8, 8, false>(g, x, y, z, f);
#pragma unroll
  for (int c = 0; c < 8; ++c) F[(8 + c) * 128] = f[c];
  voxel_trilinear<16, 8, false>(g, x, y, z, f);
#pragma unroll
  for (int c = 0; c < 8; ++c) F[(16 + c) * 128] = f[c];
  F[24 * 128] = x; F[25 * 128] = y; F[26 * 128] = z;
}

template <bool DUMP>
__global__ void __launch_bounds__(T2_THREADS, 1) field_tc2_kernel(const __grid_constant__ T2Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const FieldParams& p = P.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- shared memory carve-up (base is 1024-byte aligned: required by the 128B swizzle) ----
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = sbase + OFF_X, sB = sbase + OFF_B, sBar = sbase + OFF_BAR;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * T2_NSTAGE;
  const uint32_t bar_acc_ready = sBar + 16 * T2_NSTAGE;              // [2]
  const uint32_t bar_acc_free = bar_acc_ready + 16;                  // [2]
  const uint32_t bar_h_ready = bar_acc_free + 16;                    // [2]
  const uint32_t bar_xs_ready = bar_h_ready + 16;
  const uint32_t bar_f_ready = bar_xs_ready + 8, bar_f_free = bar_f_ready + 16;   // [2] each: raw features of tile t
  const uint32_t tmem_slot = bar_f_free + 16;
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));
  float* feat = reinterpret_cast<float*>(gen_base + OFF_F);
  int2* meta_tab = reinterpret_cast<int2*>(gen_base + OFF_META);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - sbase));
  const float* Pf = reinterpret_cast<const float*>(p.packed);

  if (threadIdx.x == 0) {
    for (int s = 0; s < T2_NSTAGE; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int t = 0; t < 2; ++t) {    // two accumulators, two tiles
      mbar_init(bar_acc_ready + 8 * t, 1);
      mbar_init(bar_acc_free + 8 * t, T2_EPI_THREADS / 32);
      mbar_init(bar_h_ready + 8 * t, T2_EPI_THREADS / 32);
    }
    mbar_init(bar_xs_ready, T2_EPI_THREADS / 32);
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_f_ready + 8 * t, T2_ENC_WARPS);
      mbar_init(bar_f_free + 8 * t, T2_EPI_THREADS / 32);
    }
    fence_barrier_init();
  }
  if (warp == T2_MMA_WARP) tmem_alloc(tmem_slot, 512);
  // XS starts as zeros: columns a layer's weights do not reach (scene layers: 272..287, pads) must stay finite
  for (uint32_t i = threadIdx.x; i < 6u * ATOM_BYTES / 16u; i += T2_THREADS)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sX + i * 16u), "r"(0u) : "memory");
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int64_t total = (int64_t)p.n_rays * p.S;
  const int64_t n_tiles = (total + TM - 1) / TM;
  const int64_t n_pairs = (n_tiles + 1) / 2;
#ifdef ONERF_WAITSTATS
  long long wstat[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_role0 = clock64();
#endif
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(p.packed);

  // The register budget ptxas compiles a region with is the one of the setmaxnreg that DOMINATES it: the two
  // instructions sit at the top of the two role regions (a common "if dec else inc" ahead of the role switch made the whole
  // kernel compile for the smaller budget: every epilogue loop variable lived in local memory).
  if (warp >= 16) {
    if (T2_REGS_CTRL < 96) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(T2_REGS_CTRL));
    if (T2_REGS_CTRL > 96) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(T2_REGS_CTRL));
    if (warp == T2_PRODUCER_WARP) {
      // =============================== weight producer ===============================
      {
        uint32_t stage = 0, phase = 0;
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
          for (int si = 0; si < P.n_slots; ++si) {
            const T2Slot sl = P.slots[si];
            const T2Layer& Ly = P.layers[sl.layer];
            const uint32_t slab_bytes = (uint32_t)Ly.N * 64u, half_bytes = slab_bytes >> (Ly.nhalf - 1);
            const uint8_t* src = blob + Ly.img_off + (size_t)sl.half * half_bytes;
            for (int gi = 0; gi < Ly.ngroups; ++gi) {
              const int grp = Ly.groups[gi];
              const int first = grp & 31, cnt = (grp >> 5) & 7;
              const int gslab = ((grp >> 8) & 1) ? Ly.nslab_x + first : first;
              T2_WAIT(0, mbar_wait(bar_empty + 8 * stage, phase ^ 1));
              if (elect_one()) {   // (the uniform-datapath instructions want a region the compiler KNOWS is one lane wide)
                mbar_expect_tx(bar_full + 8 * stage, (uint32_t)cnt * half_bytes);
                for (int i2 = 0; i2 < cnt; ++i2)
                  tma_bulk_g2s(sB + stage * T2_STAGE_BYTES + (uint32_t)i2 * half_bytes, src + (size_t)(gslab + i2) * slab_bytes,
                               half_bytes, bar_full + 8 * stage);
              }
              __syncwarp();
              if (++stage == T2_NSTAGE) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    } else if (warp == T2_MMA_WARP) {
      // =============================== MMA issuer ===============================
      // The warp stays converged for the loop control and the waits; ONE elect.sync region per ring stage holds the
      // tcgen05.mma and tcgen05.commit instructions (inside a plain "if (lane == 0)" the compiler wraps every uniform-
      // datapath instruction in its own elect loop).
      {
        uint32_t stage = 0, phase = 0, xs_phase = 0;
        uint32_t free_bits = 0, h_bits = 0;   // per-tile barrier phases, bit t
        const uint32_t ring16 = ((sB >> 4) & 0x3FFFu) | 0x10000u, xs16 = ((sX >> 4) & 0x3FFFu) | 0x10000u;   // descriptor low words
        T2Rec cur = P.rec[0];
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
          for (int si = 0; si < P.n_slots; ++si) {
            const T2Rec nxt = P.rec[(si + 1 == P.n_slots) ? 0 : si + 1];   // in flight while this slot waits and issues
            const uint32_t idesc = cur.q[0].x, hdr = cur.q[0].y, gmeta = cur.q[0].z;
            const uint32_t hb16 = hdr & 0xffffu;
            const int acc = (hdr >> 16) & 1, t = (hdr >> 17) & 1, ngroups = (int)((hdr >> 20) & 0xffu);
            const uint32_t rel[2 * T2_REC_GROUPS] = {cur.q[1].x, cur.q[1].y, cur.q[1].z, cur.q[1].w, cur.q[2].x,
                                                     cur.q[2].y, cur.q[2].z, cur.q[2].w, cur.q[3].x, cur.q[3].y};
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
            const uint32_t h_tmem = tmem_base + (uint32_t)(256 + t * 128);
            // the epilogue of the accumulator's previous user has loaded it into registers
            T2_WAIT(1, mbar_wait(bar_acc_free + 8 * acc, (free_bits >> acc) & 1u));
            free_bits ^= 1u << acc;
            if (hdr & (1u << 18)) {   // the previous layer's activations are written
              T2_WAIT(2, mbar_wait(bar_h_ready + 8 * t, (h_bits >> t) & 1u));
              h_bits ^= 1u << t;
            }
            tc_fence_after();
#pragma unroll
            for (int gi = 0; gi < T2_REC_GROUPS; ++gi) {
              if (gi < ngroups) {
                const uint32_t gm = (gmeta >> (4 * gi)) & 15u;
                const int cnt = (int)(gm & 7u);
                const bool from_h = (gm >> 3) != 0;
                // Operand words are formed BEFORE the barrier waits (the empty asm pins them there): whatever sits between
                // a satisfied wait and the tcgen05.mma instructions is pure latency on the slot-to-slot dependency chain.
                const uint32_t b_lo0 = ring16 + stage * (T2_STAGE_BYTES >> 4);
                const uint32_t a_base = from_h ? h_tmem : xs16;
                uint32_t a_w[T2_STAGE_SLABS], b_w[T2_STAGE_SLABS];
#pragma unroll
                for (int i2 = 0; i2 < T2_STAGE_SLABS; ++i2) {
                  a_w[i2] = a_base + ((rel[2 * gi + (i2 >> 1)] >> (16 * (i2 & 1))) & 0xffffu);
                  b_w[i2] = b_lo0 + (uint32_t)i2 * hb16;
                }
                uint32_t accum = (gi > 0) ? 1u : 0u;
                asm volatile("" ::"r"(a_w[0]), "r"(a_w[1]), "r"(a_w[2]), "r"(a_w[3]), "r"(b_w[0]), "r"(b_w[1]), "r"(b_w[2]), "r"(b_w[3]), "r"(accum), "r"(d_tmem), "r"(idesc));
                if (gi == 0 && (hdr & (1u << 19))) {   // XS holds this tile's X
                  T2_WAIT(3, mbar_wait(bar_xs_ready, xs_phase));
                  xs_phase ^= 1;
                  if (DUMP && (hdr & (1u << 28)) && 2 * pair + t < n_tiles) {   // training: the tile's X atoms -> workspace
                    if (elect_one()) {
                      bulk_s2g(P.dump + P.TL.act_off[0] + (size_t)(2 * pair + t) * 6 * ATOM_BYTES, sX, 6u * ATOM_BYTES);
                      bulk_commit_group();
                    }
                    __syncwarp();
                  }
                }
                T2_WAIT(4, mbar_wait(bar_full + 8 * stage, phase));
                tc_fence_after();
#ifdef ONERF_WAITSTATS
                const long long mi0 = clock64();
#endif
                if (elect_one()) {
#ifndef T2_EXP_NO_MMA
                  if (!from_h) {
#pragma unroll
                    for (int i2 = 0; i2 < T2_STAGE_SLABS; ++i2) {
                      if (i2 < cnt) {
                        umma_bf16(d_tmem, make_desc_hl(a_w[i2], DESC_HI_SW128), make_desc_hl(b_w[i2], DESC_HI_SW64), idesc, accum);
                        umma_bf16(d_tmem, make_desc_hl(a_w[i2] + 2u, DESC_HI_SW128), make_desc_hl(b_w[i2] + 2u, DESC_HI_SW64), idesc, 1u);
                        accum = 1u;
                      }
                    }
                  } else {
#pragma unroll
                    for (int i2 = 0; i2 < T2_STAGE_SLABS; ++i2) {
                      if (i2 < cnt) {
                        umma_bf16_ts(d_tmem, a_w[i2], make_desc_hl(b_w[i2], DESC_HI_SW64), idesc, accum);
                        umma_bf16_ts(d_tmem, a_w[i2] + 8u, make_desc_hl(b_w[i2] + 2u, DESC_HI_SW64), idesc, 1u);
                        accum = 1u;
                      }
                    }
                  }
#endif
#ifdef ONERF_WAITSTATS
                  const long long mi1 = clock64();
                  wstat[5] += mi1 - mi0;
#endif
                  umma_commit(bar_empty + 8 * stage);
                  if (gi == ngroups - 1) {
                    if (DUMP && (hdr & (1u << 29))) bulk_wait_group_read0();   // XS may be rewritten once this accumulator is seen
                    umma_commit(bar_acc_ready + 8 * acc);
                  }
#ifdef ONERF_WAITSTATS
                  wstat[7] += clock64() - mi1;
#endif
                }
                __syncwarp();
                if (++stage == T2_NSTAGE) { stage = 0; phase ^= 1; }
              }
            }
            cur = nxt;
          }
        }
      }
    } else {
      // =============================== gather warps: raw features of the NEXT tile pair ===============================
      const int row0 = (warp - T2_ENC_WARP0) * 32 + lane;     // this thread's rows: row0 + 64 k
      const GridView g = load_grid_view(p.grid);
      uint32_t it = 0;
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          // the epilogue warps have encoded the tile's last X of the previous pair: its features may be overwritten
          if (it > 0) T2_WAIT(5, mbar_wait(bar_f_free + 8 * t, (it - 1) & 1u));
#pragma unroll 1
          for (int k = 0; k < T2_ENC_ROWS; ++k) {
            const int row = row0 + 32 * T2_ENC_WARPS * k;
            gather_tile(p, g, feat + (size_t)t * T2_NF * 128 + row, meta_tab + ((it & 1) * 2 + t) * 128 + row,
                        (2 * pair + t) * TM + row, total);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_f_ready + 8 * t);
        }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    if (T2_REGS_EPI > 96) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(T2_REGS_EPI));
    if (T2_REGS_EPI < 96) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(T2_REGS_EPI));
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc_bits = 0;
    uint32_t stash[16];
    float sigma_a = 0.0f, sigma_b = 0.0f;
    // nothing holds the two accumulators before their first use
    if (lane == 0) {
      mbar_arrive(bar_acc_free);
      mbar_arrive(bar_acc_free + 8);
    }
    int parity = 0;
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, parity ^= 1) {
#pragma unroll 1
      uint4 e_next = P.ev_tab[0];
      for (int ei = 0; ei < P.n_events; ++ei) {
        const uint4 e = e_next;
        e_next = P.ev_tab[(ei + 1 == P.n_events) ? 0 : ei + 1];   // one event ahead: the table read hides behind this event
        if ((e.w & 0xffu) == EV_XGEN) {
          // ---- X of one tile into XS: row = thread & 127, the four column quarters of the CTA take one block each ----
          const int xf = (int)(e.w >> 8), t = xf & XG_TILE;
#ifdef ONERF_WAITSTATS
          const long long x0 = clock64();
#endif
          if (e.x != 0xffu) mbar_wait(bar_acc_ready + 8 * e.x, (acc_bits >> e.x) & 1u);   // (peek) the MMAs that read XS are done
          if (xf & XG_WAIT_F) mbar_wait(bar_f_ready + 8 * t, (uint32_t)parity);
          tc_fence_after();
#ifndef T2_EXP_NO_XGEN
          const int row = threadIdx.x & 127, q = threadIdx.x >> 7;
          const float* F = feat + (size_t)t * T2_NF * 128 + row;
          if (q == 2) {
            pe_xyz_to_chunks(sX, row, 26, F[24 * 128], F[25 * 128], F[26 * 128]);   // columns 208..271
          } else if (q < 2 || (xf & XG_FULL)) {
            float f[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) f[c] = F[((q == 3 ? 16 : 8 * q) + c) * 128];
            // scene channels 0-7: chunks 0, 2, 4, ...; 8-15: chunks 1, 3, 5, ...; object voxel block: column 272 = chunk 34
            pe8_to_chunks(sX, row, q == 3 ? 34 : q, q == 3 ? 1 : 2, f);
            if (q == 3) st_chunk(a_chunk_addr(sX, row, 47), 0u, 0u, 0u, 0u);  // columns 376..383
          }
#endif
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar_xs_ready);
            if (xf & XG_RELEASE_F) mbar_arrive(bar_f_free + 8 * t);
          }
#ifdef ONERF_WAITSTATS
          wstat[5] += clock64() - x0;
#endif
          continue;
        }
        epi_event<DUMP>(P, e, DUMP ? P.ev_dump[ei] : 0u, 2 * pair + ((e.w & SE_TILE) ? 1 : 0), n_tiles, gen_base, sbase, lane_taddr,
                        parity, stash, acc_bits, sigma_a, sigma_b
#ifdef ONERF_WAITSTATS
                  , wstat
#endif
        );
      }
    }
  }

#ifdef ONERF_WAITSTATS
  if (P.timeline && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp >= 16)) {
    long long* o = P.timeline + 900 + (warp == 0 ? 0 : warp - 15) * 10;
    for (int k = 0; k < 8; ++k) o[k] = wstat[k];
    o[8] = clock64() - t_role0;
  }
#endif
  // ---- teardown ----
  if (DUMP && warp == T2_MMA_WARP) {   // (the elected lane of an elect region is the same lane every time)
    if (elect_one()) bulk_wait_group0();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == T2_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

}  // namespace

static long long* g_timeline2 = nullptr;
extern "C" void onerf_debug_timeline2(void* dev_buf) { g_timeline2 = reinterpret_cast<long long*>(dev_buf); }

// The static program of one tile pair (host): layers, MMA slots and their 64-byte records, the epilogue's event table.
static int t2_build_program(const FieldParams& fp, T2Params& P) {
  const PackLayout& L = fp.L;
  memset(&P, 0, sizeof(P));
  P.f = fp;
  P.timeline = g_timeline2;
  const int xs = L.KX / 32, xo = L.KO / 32;
  int n = 0;
  auto add = [&](int gemm, int nx, int nh, int epi, int branch, int rc_base, int act_slot) {
    T2Layer& t = P.layers[n++];
    t.act_slot = act_slot;
    t.N = L.g[gemm].N; t.nhalf = t.N > 128 ? 2 : 1;
    t.nslab_x = nx; t.nslab_h = nh; t.epi = epi; t.branch = branch; t.rc_base = rc_base;
    t.writes_h = epi != EPI_DIR;
    t.img_off = L.g[gemm].img_off; t.bias_off = L.g[gemm].bias_off;
    int ng = 0;
    auto emit = [&](int count, int from_h) {
      for (int o = 0; o < count; o += T2_STAGE_SLABS) {
        const int c = (count - o < T2_STAGE_SLABS) ? count - o : T2_STAGE_SLABS;
        t.groups[ng++] = o | (c << 5) | (from_h << 8);
      }
    };
    emit(nx, 0);
    t.n_xgroups = ng;
    emit(nh, 1);
    t.ngroups = ng;
    for (int g = 0; g < ng; ++g)
      for (int i2 = 0; i2 < T2_STAGE_SLABS; ++i2) {
        const int s = (t.groups[g] & 31) + i2;
        t.a_rel[g][i2] = ((t.groups[g] >> 8) & 1) ? (uint32_t)s * 16u : (uint32_t)(s >> 1) * (ATOM_BYTES >> 4) + (uint32_t)(s & 1) * 4u;
      }
  };
  // Object branch first: a pair's raw features can only be replaced (by the gather of the NEXT pair, ~27K cycles on two
  // warps) after the tile's last X-fed layer; with the scene branch last that layer is S4 and ten long slots still follow.
  if (fp.want_object) {
    add(G_O0, xo, 0, EPI_HIDDEN_RC, 1, RC_OL0, 11);
    add(G_O1, 0, 4, EPI_HIDDEN, 1, 0, 12);
    add(G_O2, xo, 4, EPI_HIDDEN_RC, 1, RC_OL2, 13);
    add(G_O3, 0, 4, EPI_HIDDEN_SIGMA, 1, 0, 14);
    add(G_OFIN, 0, 4, EPI_FINAL, 1, 0, 15);
    add(G_ODIR, 0, 4, EPI_DIR, 1, RC_ODIR, 16);
  }
  if (fp.want_scene) {
    add(G_S0, xs, 0, EPI_HIDDEN, 0, 0, 1);
    add(G_S1, 0, 8, EPI_HIDDEN, 0, 0, 2);
    add(G_S2, 0, 8, EPI_HIDDEN, 0, 0, 3);
    add(G_S3, 0, 8, EPI_HIDDEN, 0, 0, 4);
    add(G_S4, xs, 8, EPI_HIDDEN, 0, 0, 5);
    add(G_S5, 0, 8, EPI_HIDDEN, 0, 0, 6);
    add(G_S6, 0, 8, EPI_HIDDEN, 0, 0, 7);
    add(G_S7, 0, 8, EPI_HIDDEN_SIGMA, 0, 0, 8);
    add(G_SFIN, 0, 8, EPI_FINAL, 0, 0, 9);
    add(G_SDIR, 0, 8, EPI_DIR, 0, RC_SDIR, 10);
  }
  P.n_layers = n;
  // slots of one tile pair
  int ns = 0, nx = 0;
  auto slot = [&](int tile, int layer, int half, int flags) {
    P.slots[ns].tile = (uint8_t)tile; P.slots[ns].layer = (uint8_t)layer; P.slots[ns].half = (uint8_t)half;
    P.slots[ns].flags = (uint8_t)flags;
    ++ns;
  };
  auto layer_slots = [&](int tile, int l) {
    const T2Layer& t = P.layers[l];
    const int wait_h = t.nslab_h > 0 ? SLOT_WAIT_H : 0;
    const int wait_x = t.nslab_x > 0 ? SLOT_WAIT_XS : 0;
    // a tile's slots of a layer are consecutive; an X-fed layer keeps XS for all of them
    if (t.nhalf == 2) {
      slot(tile, l, 0, wait_h | wait_x | (0 << 4));
      slot(tile, l, 1, 1 << 4);
    } else {
      slot(tile, l, 0, wait_h | wait_x | (tile << 4));
    }
  };
  for (int l = 0; l < n; ++l) {
    if (P.layers[l].nslab_x > 0) P.xuse_full[nx++] = P.layers[l].branch ? 1 : 0;
    for (int tile = 0; tile < 2; ++tile) layer_slots(tile, l);
  }
  P.n_slots = ns;
  P.n_xuse = nx;
  for (int i = 0; i < ns; ++i) {
    const T2Slot& sl = P.slots[i];
    const T2Layer& t = P.layers[sl.layer];
    if (t.ngroups > T2_REC_GROUPS) {
      onerf_set_error("two-tile field kernel: a layer needs %d ring stages per slot (max %d)", t.ngroups, T2_REC_GROUPS);
      return ONERF_ERR_UNSUPPORTED;
    }
    const uint32_t half_bytes = ((uint32_t)t.N * 64u) >> (t.nhalf - 1);
    uint32_t w[16] = {0};
    w[0] = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((t.N >> (t.nhalf - 1)) >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);   // = make_idesc
    w[1] = (half_bytes >> 4) | ((uint32_t)(sl.flags >> 4) << 16) | ((uint32_t)sl.tile << 17) |
           ((sl.flags & SLOT_WAIT_H) ? 1u << 18 : 0u) | ((sl.flags & SLOT_WAIT_XS) ? 1u << 19 : 0u) | ((uint32_t)t.ngroups << 20);
    // training dump of X: at the tile's FIRST X-fed layer (the object branch's O0: all 384 columns are produced there)
    if (fp.train_ws && sl.layer == 0 && t.nslab_x > 0) {
      if (sl.flags & SLOT_WAIT_XS) w[1] |= 1u << 28;
      if (sl.half == t.nhalf - 1) w[1] |= 1u << 29;
    }
    for (int g = 0; g < t.ngroups; ++g) {
      w[2] |= ((uint32_t)((t.groups[g] >> 5) & 7) | ((uint32_t)((t.groups[g] >> 8) & 1) << 3)) << (4 * g);
      for (int i2 = 0; i2 < T2_STAGE_SLABS; ++i2) w[4 + 2 * g + (i2 >> 1)] |= (t.a_rel[g][i2] & 0xffffu) << (16 * (i2 & 1));
    }
    for (int q = 0; q < 4; ++q) P.rec[i].q[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  }
  // The epilogue's program: one event per slot, in slot order, plus the X productions.  XS is one buffer:
  //   X of tile A for X-fed layer l  : right after A's last event of layer l - 1 (first event of the pair for layer 0); the
  //                                    previous reader of XS (tile B, an earlier layer) has long completed
  //   X of tile B for X-fed layer l  : right before the event of A's LAST slot of layer l, once that slot's accumulator is
  //                                    complete (every MMA of A's layer l has then read XS); the MMA warp is waiting for it
  int ne = 0, xi = 0;
  auto xgen = [&](int tile, int u, uint32_t peek_acc) {
    const uint32_t xf = (tile ? XG_TILE : 0) | (P.xuse_full[u] ? XG_FULL : 0) | (u == 0 ? XG_WAIT_F : 0) | (u == nx - 1 ? XG_RELEASE_F : 0);
    P.ev_tab[ne].x = peek_acc; P.ev_tab[ne].y = 0; P.ev_tab[ne].z = 0;
    P.ev_tab[ne].w = (uint32_t)EV_XGEN | (xf << 8);
    ++ne;
  };
  if (n > 0 && P.layers[0].nslab_x > 0) xgen(0, xi, 0xffu);
  for (int i = 0; i < ns; ++i) {
    const T2Slot& sl = P.slots[i];
    const T2Layer& t = P.layers[sl.layer];
    const int acc = sl.flags >> 4, HW = t.N >> (t.nhalf - 1);
    const bool last_slot_of_layer = (sl.half == t.nhalf - 1);
    if (sl.tile == 0 && last_slot_of_layer && t.nslab_x > 0) xgen(1, xi++, (uint32_t)acc);
    uint32_t flags = (uint32_t)t.epi | (t.nhalf == 2 ? SE_TWO : 0) | (sl.half ? SE_H1 : 0) | (sl.tile ? SE_TILE : 0) |
                     (t.branch ? SE_BRANCH : 0) | (HW == 64 ? SE_N64 : 0);
    P.ev_tab[ne].x = (uint32_t)(acc * 128) | ((uint32_t)(256 + sl.tile * 128) << 16);
    P.ev_tab[ne].y = (uint32_t)(t.bias_off + sl.half * 128);
    P.ev_tab[ne].z = (uint32_t)(t.rc_base + sl.half * 128);
    P.ev_tab[ne].w = flags | ((uint32_t)acc << 16) | ((uint32_t)t.N << 20);
    P.ev_dump[ne] = (uint32_t)(t.act_slot + 1) | ((uint32_t)(onerf_mask_word0(t.act_slot) + 1) << 8);
    ++ne;
    if (sl.tile == 0 && last_slot_of_layer && sl.layer + 1 < n && P.layers[sl.layer + 1].nslab_x > 0) xgen(0, xi, 0xffu);
  }
  P.n_events = ne;
  return ONERF_OK;
}

int onerf_launch_field_bf16_two_tile(onerf_ctx* ctx, const FieldParams& fp, cudaStream_t stream) {
  const PackLayout& L = fp.L;
  T2Params P;
  const int rc = t2_build_program(fp, P);
  if (rc != ONERF_OK) return rc;
  const int64_t total = (int64_t)fp.n_rays * fp.S;
  const int64_t tiles = (total + TM - 1) / TM, pairs = (tiles + 1) / 2;
  const int blocks = (int)(pairs < ctx->num_sms ? pairs : ctx->num_sms);
  const size_t smem = T2_SMEM_BYTES;
  if (fp.train_ws) {   // training forward: both branches, the object branch's first layer produces all of X
    if (!(fp.want_scene && fp.want_object)) {
      onerf_set_error("two-tile field kernel: the training dump needs both branches");
      return ONERF_ERR_UNSUPPORTED;
    }
    P.dump = reinterpret_cast<uint8_t*>(fp.train_ws);
    P.TL = onerf_make_train_layout(L.use_voxel, total);
    ONERF_CUDA(cudaFuncSetAttribute(field_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    field_tc2_kernel<true><<<blocks, T2_THREADS, smem, stream>>>(P);
  } else {
    ONERF_CUDA(cudaFuncSetAttribute(field_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    field_tc2_kernel<false><<<blocks, T2_THREADS, smem, stream>>>(P);
  }
  ONERF_LAUNCH_CHECK(ctx);
  return ONERF_OK;
}

// Test hook (tests/test_two_tile_program_cpu.py, no GPU needed): the program the launcher would hand to the kernel.
//   out[0..3] = n_slots, n_events, n_layers, words per slot (4 + 16)
//   per slot  : tile, layer, half, flags (SlotFlags | accumulator << 4), then the 16 record words
//   per event : the 4 ev_tab words, then ev_dump
//   per layer : N, halves, X slabs, H slabs, ring stages per slot, groups[8] (what the weight producer walks)
// Returns the number of words written, or a negative error code.
extern "C" int onerf_debug_two_tile_program(int want_scene, int want_object, int train, uint32_t* out, int cap) {
  FieldParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.L = onerf_make_layout(1);
  fp.want_scene = want_scene; fp.want_object = want_object;
  fp.n_rays = 1024; fp.S = 128;
  fp.train_ws = train ? reinterpret_cast<void*>(uintptr_t(1024)) : nullptr;   // only tested against null by the builder
  T2Params* P = new T2Params;
  const int rc = t2_build_program(fp, *P);
  if (rc != ONERF_OK) { delete P; return rc < 0 ? rc : -rc; }
  const int need = 4 + P->n_slots * 20 + P->n_events * 5 + P->n_layers * (5 + T2_MAX_GROUPS);
  if (need > cap) { delete P; return -need; }
  int o = 0;
  out[o++] = (uint32_t)P->n_slots; out[o++] = (uint32_t)P->n_events; out[o++] = (uint32_t)P->n_layers; out[o++] = 20u;
  for (int i = 0; i < P->n_slots; ++i) {
    out[o++] = P->slots[i].tile; out[o++] = P->slots[i].layer; out[o++] = P->slots[i].half; out[o++] = P->slots[i].flags;
    for (int q = 0; q < 4; ++q) {
      out[o++] = P->rec[i].q[q].x; out[o++] = P->rec[i].q[q].y; out[o++] = P->rec[i].q[q].z; out[o++] = P->rec[i].q[q].w;
    }
  }
  for (int i = 0; i < P->n_events; ++i) {
    out[o++] = P->ev_tab[i].x; out[o++] = P->ev_tab[i].y; out[o++] = P->ev_tab[i].z; out[o++] = P->ev_tab[i].w;
    out[o++] = P->ev_dump[i];
  }
  // per layer, the producer's view of the same walk: N, halves, X slabs, H slabs, ring stages and their packed descriptions
  for (int l = 0; l < P->n_layers; ++l) {
    const T2Layer& t = P->layers[l];
    out[o++] = (uint32_t)t.N; out[o++] = (uint32_t)t.nhalf; out[o++] = (uint32_t)t.nslab_x; out[o++] = (uint32_t)t.nslab_h;
    out[o++] = (uint32_t)t.ngroups;
    for (int g = 0; g < T2_MAX_GROUPS; ++g) out[o++] = (uint32_t)t.groups[g];
  }
  delete P;
  return o;
}
