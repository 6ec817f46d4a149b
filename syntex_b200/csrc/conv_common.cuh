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
  int pad;             // taps reach from -pad to 2 - pad around the output pixel (1 = SAME)
  const float* bias;   // [cout] or null
  const bf16* addend;  // [nimg,H,W,cout] or null (may alias out)
  const bf16* mask;    // [nimg,H,W,cout] or null: result is zeroed where mask <= 0
  bf16* out;           // [nimg,H,W,cout]
  bf16* out_lo;        // SPLIT: low halves of the output (out holds the high halves)
  // Optional copy of v right after the addend (before ReLU / mask / scatter): the total gradient dL/dF_k that the
  // PO models consume (tf.gradients(loss, feat[k]), po/improved_model.py:227-231)
  bf16* export_out;
  // Fused 2x2 average pooling of the result (patch kernel only; vgg19.py:90-91): pooled hi/lo planes [n,H/2,W/2,cout]
  bf16* pool_out;
  bf16* pool_out_lo;
  // Fused AvgPoolGrad + ReluGrad: instead of `out`, write 0.25 * v to the four fine pixels of [n,2H,2W,cout],
  // zeroed where up_mask <= 0 (the activations of the convolution that fed the pool)
  bf16* up_out;
  const bf16* up_mask;
  // conv1_1 input-gradient mode (N tile 16, channels 0..2 real): fp32 [nimg,H,W,3] = acc * 255 (* sigmoid')
  float* rgb_out;
  const float* rgb_image;
  int rgb_preimage;
};

union Pack8 {
  uint4 u;
  __nv_bfloat162 h2[4];
};

// Epilogue for 32 consecutive output channels of one pixel: v[] are the fp32 accumulators.
//   v += bias; v += addend; relu; zero where mask <= 0; then one of
//     store bf16 (hi) and, in split mode, lo = bf16(v - hi)          [+ fused 2x2 average pool of v]
//     or scatter 0.25 * v to the four fine pixels it was pooled from  [fused AvgPoolGrad + ReluGrad]
// Called by all 32 lanes of the warp (the pooling uses shuffles); `valid` guards the memory operations.
// Pooling relies on the patch kernel's row mapping: lane = 8 * (h & 3) + (w & 7), so a 2x2 window is
// {lane, lane ^ 1, lane ^ 8, lane ^ 9}.
template <bool SPLIT>
__device__ __forceinline__ void conv_epilogue_store(const ConvKernelParams& p, float (&v)[32], bool valid, int n, int h,
                                                    int w, int ch0) {
  const size_t off = ((static_cast<size_t>(n) * p.H + h) * p.W + w) * p.cout + ch0;
  if (valid) {
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
    if (p.export_out) {
      uint4* e4 = reinterpret_cast<uint4*>(p.export_out + off);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Pack8 pk;
#pragma unroll
        for (int e = 0; e < 4; ++e) pk.h2[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
        e4[j] = pk.u;
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
    if (p.up_out) {
      const int H2 = 2 * p.H, W2 = 2 * p.W;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const size_t uo = ((static_cast<size_t>(n) * H2 + 2 * h + dy) * W2 + 2 * w + dx) * p.cout + ch0;
          const uint4* m4 = reinterpret_cast<const uint4*>(p.up_mask + uo);
          uint4* o4 = reinterpret_cast<uint4*>(p.up_out + uo);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            Pack8 pm, po;
            pm.u = __ldg(m4 + j);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(pm.h2[e]);
              po.h2[e] = __floats2bfloat162_rn(f.x > 0.0f ? 0.25f * v[8 * j + 2 * e] : 0.0f,
                                               f.y > 0.0f ? 0.25f * v[8 * j + 2 * e + 1] : 0.0f);
            }
            o4[j] = po.u;
          }
        }
      }
    } else {
      uint4* o4 = reinterpret_cast<uint4*>(p.out + off);
      // the low halves are optional even in split mode: a layer that only feeds a fused pool (whose hi/lo outputs
      // carry the precision forward) and the backward mask needs its high plane only
      uint4* l4 = (SPLIT && p.out_lo) ? reinterpret_cast<uint4*>(p.out_lo + off) : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Pack8 pk, pl;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = v[8 * j + 2 * e], b = v[8 * j + 2 * e + 1];
          pk.h2[e] = __floats2bfloat162_rn(a, b);
          if (SPLIT) {
            const float2 hh = __bfloat1622float2(pk.h2[e]);
            pl.h2[e] = __floats2bfloat162_rn(a - hh.x, b - hh.y);
          }
        }
        o4[j] = pk.u;
        if (SPLIT && l4) l4[j] = pl.u;
      }
    }
  }
  if (p.pool_out) {
    // 2x2 average of the fp32 results (H, W even: a window is valid as a whole)
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float t = v[j];
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 8);
      v[j] = 0.25f * t;
    }
    const int lane = threadIdx.x & 31;
    if (valid && (lane & 9) == 0) {
      const size_t po = ((static_cast<size_t>(n) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1)) * p.cout + ch0;
      uint4* o4 = reinterpret_cast<uint4*>(p.pool_out + po);
      uint4* l4 = p.pool_out_lo ? reinterpret_cast<uint4*>(p.pool_out_lo + po) : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Pack8 pk, pl;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = v[8 * j + 2 * e], b = v[8 * j + 2 * e + 1];
          pk.h2[e] = __floats2bfloat162_rn(a, b);
          const float2 hh = __bfloat1622float2(pk.h2[e]);
          pl.h2[e] = __floats2bfloat162_rn(a - hh.x, b - hh.y);
        }
        o4[j] = pk.u;
        if (l4) l4[j] = pl.u;
      }
    }
  }
}

}  // namespace stx
