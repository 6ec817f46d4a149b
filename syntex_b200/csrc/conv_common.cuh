// Shared by the two convolution kernels (conv_igemm.cu: one TMA box per filter tap; conv_patch.cu: one halo patch
// per channel chunk): kernel parameters and the epilogue that turns fp32 accumulators into the stored result.
#pragma once
#include <cuda_fp16.h>
#include <cuda_fp8.h>

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
  // One-bit ReLU masks for a convolution whose full-resolution result is read by nobody but the AvgPoolGrad + ReluGrad
  // scatter of the backward pass (the layers in front of a fused pool: conv1_2, conv2_2, conv3_4, conv4_4 in the
  // Gatys plans): bit j of word [ch0 / 32][pixel] = result of channel ch0 + j > 0. The forward writes bits_out INSTEAD
  // of `out` (which may then be null), the scatter reads up_bits instead of up_mask: 1/16 of the bytes both ways.
  uint32_t* bits_out;          // [cout / 32][nimg * H * W]
  const uint32_t* up_bits;     // [cout / 32][nimg * 2H * 2W]
  // conv1_1 input-gradient mode (N tile 16, channels 0..2 real): fp32 [nimg,H,W,3] = acc * 255 (* sigmoid')
  float* rgb_out;
  const float* rgb_image;
  int rgb_preimage;
  // f16+f8 forward (plan precision 2, see kF8*): the correction accumulator is multiplied by corr_scale before it
  // is added to the main one. aux_out / pool_aux (nullable): a bf16 copy of the stored / pooled result, kept for the
  // loss layers - their features are the A operand of the Gram gradient against bf16 matrices, and kind::f16 does
  // NOT take an fp16 A with a bf16 B (measured on the B200: illegal instruction).
  float corr_scale;
  bf16* aux_out;
  bf16* pool_aux;
  // Per-image switch (nullable): images whose flag is 0 are skipped by the persistent work loops - their outputs keep
  // whatever they held. The device optimiser clears the flag of a job that has finished while the other jobs of the
  // batch still run (csrc/lbfgs.cu), so the tail of a batch costs what its remaining jobs cost.
  const int* active;
};

union Pack8 {
  uint4 u;
  __nv_bfloat162 h2[4];
};

// The epilogue reads and writes 32 bytes per access (st256 / ld256 below).
inline bool epilogue_pointers_aligned(const ConvKernelParams& p) {
  const void* ptrs[] = {p.addend, p.mask, p.out, p.out_lo, p.export_out, p.pool_out, p.pool_out_lo, p.up_out, p.up_mask,
                        p.aux_out, p.pool_aux};
  for (const void* q : ptrs)
    if (reinterpret_cast<uintptr_t>(q) & 31) return false;
  return true;
}

// ---- f16 + f8 activation format (plan precision 2) ----------------------------------------------------------
// A recorded layer is TWO planes of 2 bytes per element:
//   hi   [n,H,W,C] fp16: A16 = fp16(A)                        (main MMA operand, Gram operand, ReLU mask)
//   corr [n,H,W,C/64][128 B]: per pixel and 64-channel chunk, 64 bytes e4m3(A * kF8ScaleA) followed by 64 bytes
//        e4m3((A - A16) * kF8ScaleLo)
// so that one K = 128-byte row of the correction MMA (kind::f8f6f4) is [A8 | Alo8] against a weight row [Wlo8 | W8]:
//   A8 * Wlo8 + Alo8 * W8  =  2^S (A * W_lo + A_lo * W)   up to e4m3 rounding of quantities that are already 2^-11
//   of the main term (relative error ~2^-15: bf16x3 class, measured tests/tools/precision_emulation.py).
// The scales are powers of two: activations are data, so theirs are fixed (A / 8 covers 2^-12 .. 3584, A_lo * 512
// the same range relative to A's half-ulp); the weights' are chosen per layer from max |W| (vgg_create), with
// scale(W_lo) = 4096 * scale(W) so that both products carry the same factor 2^S = 512 * scale(W).
constexpr float kF8ScaleA = 0.125f;
constexpr float kF8ScaleLo = 512.0f;

__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}

union Pack8h {
  uint4 u;
  __half2 h2[4];
};

// 256-bit global accesses (sm_100: LDG/STG.E.256): one FULL 32-byte sector per lane. A thread of the epilogue owns 64
// contiguous bytes per plane; with 128-bit accesses every instruction touched 32 half-filled sectors of 32 different
// lines (the LSU works per sector per instruction: conv1_1 forward spent 28 M sector operations on 0.9 M requests,
// profiles/r02_ncu_full_conv1_1_fwd_cuda_cores.md). Addresses must be 32-byte aligned (checked by the launchers).
__device__ __forceinline__ void st256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void ld256_nc(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void ld256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p)
               : "memory");
}

// 32 consecutive channels (ch0 a multiple of 32) of one pixel -> fp16 plane and correction plane.
// pix = pixel index, C = channels of the layer; lo_plane may be null (layer whose low half nothing reads).
__device__ __forceinline__ void store_f16f8_32(const float (&v)[32], bf16* hi_plane, bf16* corr_plane, size_t pix, int C,
                                               int ch0) {
  Pack8h pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      pk[j].h2[e] = __floats2half2_rn(fminf(v[8 * j + 2 * e], 65504.0f), fminf(v[8 * j + 2 * e + 1], 65504.0f));
  }
  uint8_t* hb = reinterpret_cast<uint8_t*>(hi_plane + pix * C + ch0);
  st256(hb, pk[0].u, pk[1].u);
  st256(hb + 32, pk[2].u, pk[3].u);
  if (corr_plane) {
    uint32_t a8[8], l8[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 hh = __half22float2(pk[j].h2[e]);
        lo[2 * e] = (fminf(v[8 * j + 2 * e], 65504.0f) - hh.x) * kF8ScaleLo;
        lo[2 * e + 1] = (fminf(v[8 * j + 2 * e + 1], 65504.0f) - hh.y) * kF8ScaleLo;
      }
      a8[2 * j] = pack_e4m3x4(v[8 * j] * kF8ScaleA, v[8 * j + 1] * kF8ScaleA, v[8 * j + 2] * kF8ScaleA,
                              v[8 * j + 3] * kF8ScaleA);
      a8[2 * j + 1] = pack_e4m3x4(v[8 * j + 4] * kF8ScaleA, v[8 * j + 5] * kF8ScaleA, v[8 * j + 6] * kF8ScaleA,
                                  v[8 * j + 7] * kF8ScaleA);
      l8[2 * j] = pack_e4m3x4(lo[0], lo[1], lo[2], lo[3]);
      l8[2 * j + 1] = pack_e4m3x4(lo[4], lo[5], lo[6], lo[7]);
    }
    uint8_t* cb = reinterpret_cast<uint8_t*>(corr_plane) + pix * C * 2 + static_cast<size_t>(ch0 >> 6) * 128 + (ch0 & 63);
    st256(cb, make_uint4(a8[0], a8[1], a8[2], a8[3]), make_uint4(a8[4], a8[5], a8[6], a8[7]));
    st256(cb + 64, make_uint4(l8[0], l8[1], l8[2], l8[3]), make_uint4(l8[4], l8[5], l8[6], l8[7]));
  }
}

// 32 consecutive channels of one pixel -> bf16
__device__ __forceinline__ void store_bf16_32(const float (&v)[32], bf16* dst) {
  Pack8 pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int e = 0; e < 4; ++e) pk[j].h2[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
  }
  st256(dst, pk[0].u, pk[1].u);
  st256(reinterpret_cast<uint8_t*>(dst) + 32, pk[2].u, pk[3].u);
}

// Sign test of a stored activation, valid for bf16 AND fp16 bit patterns: positive <=> the 16 bits, read as a signed
// integer, are > 0 (the ReLU masks of the backward pass only need the sign of the recorded activation).
__device__ __forceinline__ bool pos16_lo(uint32_t w) { return static_cast<short>(w & 0xffffu) > 0; }
__device__ __forceinline__ bool pos16_hi(uint32_t w) { return static_cast<int>(w) >> 16 > 0; }


// Epilogue for 32 consecutive output channels of one pixel: v[] are the fp32 accumulators.
//   v += bias; v += addend; relu; zero where mask <= 0; then one of
//     store bf16 (hi) and, in split mode, lo = bf16(v - hi)          [+ fused 2x2 average pool of v]
//     or scatter 0.25 * v to the four fine pixels it was pooled from  [fused AvgPoolGrad + ReluGrad]
// Called by all 32 lanes of the warp (the pooling uses shuffles); `valid` guards the memory operations.
// Pooling relies on the patch kernel's row mapping: lane = 8 * (h & 3) + (w & 7), so a 2x2 window is
// {lane, lane ^ 1, lane ^ 8, lane ^ 9}.
// FMT: 0 = bf16 planes (hi, and lo in split mode), 2 = fp16 + e4m3 correction planes (see kF8*)
template <bool SPLIT, int FMT = 0>
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
      uint4 ad[4];
      ld256(p.addend + off, ad[0], ad[1]);          // may alias `out`: not through the read-only path
      ld256(p.addend + off + 16, ad[2], ad[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Pack8 pk;
        pk.u = ad[j];
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
      uint4 mk[4];
      ld256_nc(p.mask + off, mk[0], mk[1]);
      ld256_nc(p.mask + off + 16, mk[2], mk[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t mw[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (!pos16_lo(mw[e])) v[8 * j + 2 * e] = 0.0f;
          if (!pos16_hi(mw[e])) v[8 * j + 2 * e + 1] = 0.0f;
        }
      }
    }
    if (p.bits_out) {
      uint32_t word = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) word |= (v[j] > 0.0f ? 1u : 0u) << j;
      p.bits_out[static_cast<size_t>(ch0 >> 5) * (static_cast<size_t>(p.nimg) * p.H * p.W) +
                 (static_cast<size_t>(n) * p.H + h) * p.W + w] = word;
    }
    if (p.up_out && p.up_bits) {
      const int H2 = 2 * p.H, W2 = 2 * p.W;
      const size_t plane = static_cast<size_t>(p.nimg) * H2 * W2;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const size_t pix = (static_cast<size_t>(n) * H2 + 2 * h + dy) * W2 + 2 * w;      // even: the pair is 8-byte aligned
        const uint2 wd = __ldg(reinterpret_cast<const uint2*>(p.up_bits + static_cast<size_t>(ch0 >> 5) * plane + pix));
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const uint32_t word = dx ? wd.y : wd.x;
          Pack8 po[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              po[j].h2[e] = __floats2bfloat162_rn(((word >> (8 * j + 2 * e)) & 1u) ? 0.25f * v[8 * j + 2 * e] : 0.0f,
                                                  ((word >> (8 * j + 2 * e + 1)) & 1u) ? 0.25f * v[8 * j + 2 * e + 1] : 0.0f);
            }
          }
          const size_t uo = (pix + dx) * p.cout + ch0;
          st256(p.up_out + uo, po[0].u, po[1].u);
          st256(p.up_out + uo + 16, po[2].u, po[3].u);
        }
      }
    } else if (p.up_out) {
      const int H2 = 2 * p.H, W2 = 2 * p.W;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const size_t uo = ((static_cast<size_t>(n) * H2 + 2 * h + dy) * W2 + 2 * w + dx) * p.cout + ch0;
          uint4 mk[4];
          ld256_nc(p.up_mask + uo, mk[0], mk[1]);
          ld256_nc(p.up_mask + uo + 16, mk[2], mk[3]);
          Pack8 po[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t mw[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              po[j].h2[e] = __floats2bfloat162_rn(pos16_lo(mw[e]) ? 0.25f * v[8 * j + 2 * e] : 0.0f,
                                                  pos16_hi(mw[e]) ? 0.25f * v[8 * j + 2 * e + 1] : 0.0f);
            }
          }
          st256(p.up_out + uo, po[0].u, po[1].u);
          st256(p.up_out + uo + 16, po[2].u, po[3].u);
        }
      }
    } else if (FMT == 2) {
      if (p.out) store_f16f8_32(v, p.out, p.out_lo, (static_cast<size_t>(n) * p.H + h) * p.W + w, p.cout, ch0);
      if (p.aux_out) store_bf16_32(v, p.aux_out + off);
    } else if (p.out) {
      // the low halves are optional even in split mode: a layer that only feeds a fused pool (whose hi/lo outputs
      // carry the precision forward) and the backward mask needs its high plane only
      bf16* lo_dst = (SPLIT && p.out_lo) ? p.out_lo + off : nullptr;
      Pack8 pk[4], pl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = v[8 * j + 2 * e], b = v[8 * j + 2 * e + 1];
          pk[j].h2[e] = __floats2bfloat162_rn(a, b);
          if (SPLIT) {
            const float2 hh = __bfloat1622float2(pk[j].h2[e]);
            pl[j].h2[e] = __floats2bfloat162_rn(a - hh.x, b - hh.y);
          }
        }
      }
      st256(p.out + off, pk[0].u, pk[1].u);
      st256(p.out + off + 16, pk[2].u, pk[3].u);
      if (SPLIT && lo_dst) {
        st256(lo_dst, pl[0].u, pl[1].u);
        st256(lo_dst + 16, pl[2].u, pl[3].u);
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
    if (FMT == 2) {
      if (valid && (lane & 9) == 0) {
        const size_t ppix = (static_cast<size_t>(n) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
        store_f16f8_32(v, p.pool_out, p.pool_out_lo, ppix, p.cout, ch0);
        if (p.pool_aux) store_bf16_32(v, p.pool_aux + ppix * p.cout + ch0);
      }
    } else if (valid && (lane & 9) == 0) {
      const size_t po = ((static_cast<size_t>(n) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1)) * p.cout + ch0;
      Pack8 pk[4], pl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = v[8 * j + 2 * e], b = v[8 * j + 2 * e + 1];
          pk[j].h2[e] = __floats2bfloat162_rn(a, b);
          const float2 hh = __bfloat1622float2(pk[j].h2[e]);
          pl[j].h2[e] = __floats2bfloat162_rn(a - hh.x, b - hh.y);
        }
      }
      st256(p.pool_out + po, pk[0].u, pk[1].u);
      st256(p.pool_out + po + 16, pk[2].u, pk[3].u);
      if (p.pool_out_lo) {
        st256(p.pool_out_lo + po, pl[0].u, pl[1].u);
        st256(p.pool_out_lo + po + 16, pl[2].u, pl[3].u);
      }
    }
  }
}

}  // namespace stx
