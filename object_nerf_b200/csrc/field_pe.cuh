// Positional encoding of one sample row into UMMA K-major SWIZZLE_128B shared-memory atoms (the A operand "X" of the
// X-fed layers), shared by the fused forward kernels (field_tc.cu, field_tc2.cu).
// Reference: models/embedding_helper.py:57-74 (Embedding.forward), :403-409 (channel split + PE of the voxel features).
#pragma once
#include "tc_common.cuh"

namespace tc {

// 16-byte chunk `chunk` (8 bf16 of K) of row `row` in an A buffer made of SWIZZLE_128B atoms (64 K per atom)
__device__ __forceinline__ uint32_t a_chunk_addr(uint32_t base, int row, int chunk) {
  return base + (uint32_t)(chunk >> 3) * ATOM_BYTES + (uint32_t)row * 128u + (uint32_t)(((chunk & 7) ^ (row & 7)) << 4);
}
__device__ __forceinline__ void st_chunk(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// PE of 8 channels -> 13 chunks [f | sin 2^k f | cos 2^k f]_k at chunk0 + stride * block
__device__ __forceinline__ void pe8_to_chunks(uint32_t xbase, int row, int chunk0, int stride, const float* f) {
  float s[8], c[8];
  st_chunk(a_chunk_addr(xbase, row, chunk0), pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
           pack_bf16(f[6], f[7]));
#pragma unroll
  for (int j = 0; j < 8; ++j) __sincosf(f[j], &s[j], &c[j]);  // |f| = O(1), 6 octaves: error stays << bf16 ulp
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    st_chunk(a_chunk_addr(xbase, row, chunk0 + stride * (1 + 2 * k)), pack_bf16(s[0], s[1]), pack_bf16(s[2], s[3]),
             pack_bf16(s[4], s[5]), pack_bf16(s[6], s[7]));
    st_chunk(a_chunk_addr(xbase, row, chunk0 + stride * (2 + 2 * k)), pack_bf16(c[0], c[1]), pack_bf16(c[2], c[3]),
             pack_bf16(c[4], c[5]), pack_bf16(c[6], c[7]));
    if (k < 5) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // double-angle step to the next octave
        const float s2 = 2.0f * s[j] * c[j];
        c[j] = fmaf(-2.0f * s[j], s[j], 1.0f);
        s[j] = s2;
      }
    }
  }
}

// PE10(xyz): 63 values in reference order + one zero -> 8 chunks starting at chunk0.  Values are packed to
// bf16 pairs as they are produced (the stream position is a compile-time constant after unrolling).
__device__ __forceinline__ void pe_xyz_to_chunks(uint32_t xbase, int row, int chunk0, float x, float y, float z) {
  uint32_t pk[32];
  float pend = 0.0f;
  int pos = 0;
  auto emit = [&](float val) {
    if ((pos & 1) == 0) pend = val;
    else pk[pos >> 1] = pack_bf16(pend, val);
    ++pos;
  };
  emit(x); emit(y); emit(z);
  float s[3], c[3];
  sincosf(x, &s[0], &c[0]);
  sincosf(y, &s[1], &c[1]);
  sincosf(z, &s[2], &c[2]);
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    emit(s[0]); emit(s[1]); emit(s[2]);
    emit(c[0]); emit(c[1]); emit(c[2]);
    if (k < 9) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {  // double-angle step to the next octave
        const float s2 = 2.0f * s[j] * c[j];
        c[j] = fmaf(-2.0f * s[j], s[j], 1.0f);
        s[j] = s2;
      }
    }
  }
  emit(0.0f);
#pragma unroll
  for (int q = 0; q < 8; ++q)
    st_chunk(a_chunk_addr(xbase, row, chunk0 + q), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

// ---- training dump (field_tc.cu, field_tc2.cu): bf16 activation atoms and LeakyReLU sign masks, layout.h: TrainLayout ----
// where one thread's output of one layer half goes in the training dump
struct DumpDst {
  uint8_t* row;      // byte address of (tile, atom 0, this row) of the layer's activation slot; null = no dump
  uint32_t* mask;    // &masks[tile][word 0 of the layer][this row], stride 128 words per mask word; null = no mask
  int swz;           // row & 7
};
// store NP packed bf16 pairs (columns n .. n + 2 NP - 1 of the layer) and their sign bits
template <int NP>
__device__ __forceinline__ void dump_packed(const DumpDst& d, int n, int word, const uint32_t* pk) {
  if (d.row == nullptr) return;
  uint8_t* base = d.row + (size_t)(n >> 6) * ATOM_BYTES;
  const int chunk0 = (n & 63) >> 3;
#pragma unroll
  for (int j = 0; j < NP / 4; ++j)
    *reinterpret_cast<uint4*>(base + (((chunk0 + j) ^ d.swz) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  if (d.mask) {
    // bit 2 j / 2 j + 1 = 1 where the low / high bf16 of pk[j] is non-negative.  Four values at a time: the two sign-carrying
    // bytes of two words gathered by one PRMT, their top bits squeezed into a nibble by one multiply.
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < NP / 2; ++k) {
      const uint32_t y = (__byte_perm(pk[2 * k], pk[2 * k + 1], 0x7531) >> 7) & 0x01010101u;
      m |= ((((y * 0x01020408u) >> 24) & 15u) ^ 15u) << (4 * k);
    }
    d.mask[word * 128] = m;
  }
}


}  // namespace tc
