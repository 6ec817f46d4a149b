// Shared by the two convolution kernels (conv_igemm.cu: one TMA box per filter tap; conv_patch.cu: one halo patch
// per channel chunk): kernel parameters and the epilogue that turns fp32 accumulators into the stored result.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace stx {

struct ConvKernelParams {
  int H, W, nimg;
  int cout;            // total output channels (row pitch of out/addend/mask)
  int cchunks;         // C_in / 64
  int taps;            // 9 or 1
  int bw_log2, bh_log2;
  int tiles_w, tiles_h;
  int n_step;          // images advanced per tile along the batch axis
  int per_image_w;     // weight "tap" index = image index (Gram dF mode); only rows of the first image are stored
  int relu;
  const float* bias;   // [cout] or null
  const bf16* addend;  // [nimg,H,W,cout] or null (may alias out)
  const bf16* mask;    // [nimg,H,W,cout] or null: result is zeroed where mask <= 0
  bf16* out;           // [nimg,H,W,cout]
  bf16* out_lo;        // SPLIT: low halves of the output (out holds the high halves)
};

union Pack8 {
  uint4 u;
  __nv_bfloat162 h2[4];
};

// Epilogue for 32 consecutive output channels of one pixel: v[] are the fp32 accumulators.
//   v += bias; v += addend; relu; zero where mask <= 0; store bf16 (hi) and, in split mode, lo = bf16(v - hi).
template <bool SPLIT>
__device__ __forceinline__ void conv_epilogue_store(const ConvKernelParams& p, float (&v)[32], size_t off, int ch0) {
  if (p.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + ch0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      v[4 * j + 0] += b.x;
      v[4 * j + 1] += b.y;
      v[4 * j + 2] += b.z;
      v[4 * j + 3] += b.w;
    }
  }
  if (p.addend) {
    const uint4* a4 = reinterpret_cast<const uint4*>(p.addend + off);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      Pack8 pk;
      pk.u = a4[j];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(pk.h2[e]);
        v[8 * j + 2 * e] += f.x;
        v[8 * j + 2 * e + 1] += f.y;
      }
    }
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (p.mask) {
    const uint4* m4 = reinterpret_cast<const uint4*>(p.mask + off);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      Pack8 pk;
      pk.u = __ldg(m4 + j);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(pk.h2[e]);
        if (!(f.x > 0.0f)) v[8 * j + 2 * e] = 0.0f;
        if (!(f.y > 0.0f)) v[8 * j + 2 * e + 1] = 0.0f;
      }
    }
  }
  uint4* o4 = reinterpret_cast<uint4*>(p.out + off);
  uint4* l4 = SPLIT ? reinterpret_cast<uint4*>(p.out_lo + off) : nullptr;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    Pack8 pk, pl;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = v[8 * j + 2 * e], b = v[8 * j + 2 * e + 1];
      pk.h2[e] = __floats2bfloat162_rn(a, b);
      if (SPLIT) {
        const float2 h = __bfloat1622float2(pk.h2[e]);
        pl.h2[e] = __floats2bfloat162_rn(a - h.x, b - h.y);
      }
    }
    o4[j] = pk.u;
    if (SPLIT) l4[j] = pl.u;
  }
}

}  // namespace stx
