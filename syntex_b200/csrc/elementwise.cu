// Bandwidth-bound pieces of the VGG19 path: the 3-channel first layer (fused with preprocessing), 2x2 average
// pooling forward / backward (+ReLU mask), weight packing and dtype conversion. All are simple coalesced
// 16-byte-vector kernels; none has data reuse worth staging.
#include "conv_common.cuh"
#include "ops.h"

namespace stx {

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------------------
// conv1_1 forward. Reference: vgg19.py:142-146 (x*255 - mean) and :171-179 (conv + bias + relu), optional
// sigmoid pre-image (gatys/synthesizer.py:59). Each thread computes 32 output channels of TWO horizontally adjacent
// pixels, so every shared-memory weight vector (broadcast LDS.128) feeds 8 FMAs and the 3x4 input patch is shared.
// Zero padding applies to the *preprocessed* image, as in the TF graph.
// Packed fp32 FMA (FFMA2 on sm_100): d.{x,y} += a.{x,y} * b.{x,y} in one issue slot.
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

// The kernel is issue-bound on the 1728 FMAs per pixel (27 taps x 64 channels): the packed FFMA2 form halves that, and
// a grid-stride loop over pixel pairs lets a block load the 7 KB of weights once instead of once per 128 pixels
// (4608 blocks at batch 9 spent ~40 % of their life in that prologue; 120 us per launch, profiles/r01_ncu_full_final_b9.md).
__global__ void __launch_bounds__(128, 4)
conv1_1_fwd_kernel(const float* __restrict__ image, int nimg, int H, int W, int is_preimage, float m0, float m1,
                   float m2, const float* __restrict__ w, const float* __restrict__ b, bf16* __restrict__ out,
                   bf16* __restrict__ out_lo, int fmt, bf16* __restrict__ aux) {
  __shared__ __align__(16) float sw[27 * 64];
  __shared__ __align__(16) float sb[64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) sw[i] = w[i];      // constants: safe before pdl_wait
  if (threadIdx.x < 64) sb[threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  pdl_launch_dependents();   // programmatic dependent launch: see launch_pdl (common.h)
  pdl_wait();
  const int Wp = (W + 1) >> 1;                                   // pixel pairs per row
  const long long npairs = static_cast<long long>(nimg) * H * Wp;
  const float mean[3] = {m0, m1, m2};
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; (gid >> 1) < npairs;
       gid += stride) {
    const long long pair = gid >> 1;
    const int half = static_cast<int>(gid & 1);
    const int x = static_cast<int>(pair % Wp) * 2;
    const int y = static_cast<int>((pair / Wp) % H);
    const long long img_base = (pair / (static_cast<long long>(Wp) * H)) * H * W;

    float patch[3][4][3];                                          // rows y-1..y+1, columns x-1..x+2
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int yy = y + r - 1, xx = x + s - 1;
        const bool in = (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
        const float* src = image + (img_base + static_cast<long long>(yy) * W + xx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = 0.0f;
          if (in) {
            v = __ldg(src + c);
            if (is_preimage) v = sigmoidf_(v);
            v = v * 255.0f - mean[c];
          }
          patch[r][s][c] = v;
        }
      }
    }
    unsigned long long acc0[16], acc1[16];                         // channel pairs (2k, 2k+1) of the two pixels
    const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(sb + half * 32);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const ulonglong2 bb = b2[k];
      acc0[2 * k] = acc1[2 * k] = bb.x;
      acc0[2 * k + 1] = acc1[2 * k + 1] = bb.y;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const ulonglong2* w2 = reinterpret_cast<const ulonglong2*>(sw + ((r * 3 + s) * 3 + c) * 64 + half * 32);
          const unsigned long long p0 = pack2(patch[r][s][c], patch[r][s][c]);
          const unsigned long long p1 = pack2(patch[r][s + 1][c], patch[r][s + 1][c]);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const ulonglong2 ww = w2[k];                           // one LDS.128 = two channel pairs
            ffma2(acc0[2 * k], p0, ww.x);
            ffma2(acc0[2 * k + 1], p0, ww.y);
            ffma2(acc1[2 * k], p1, ww.x);
            ffma2(acc1[2 * k + 1], p1, ww.y);
          }
        }
      }
    }
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      if (x + px >= W) break;
      const unsigned long long* acc = px ? acc1 : acc0;
      const long long pix = img_base + static_cast<long long>(y) * W + x + px;
      if (fmt == 2) {      // fp16 plane + e4m3 correction plane (conv_common.cuh, kF8*)
        float v32[32];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 v = unpack2(acc[k]);
          v32[2 * k] = fmaxf(v.x, 0.0f);
          v32[2 * k + 1] = fmaxf(v.y, 0.0f);
        }
        store_f16f8_32(v32, out, out_lo, static_cast<size_t>(pix), 64, half * 32);
        if (aux) store_bf16_32(v32, aux + pix * 64 + half * 32);
        continue;
      }
      uint4* o4 = reinterpret_cast<uint4*>(out + pix * 64 + half * 32);
      uint4* l4 = out_lo ? reinterpret_cast<uint4*>(out_lo + pix * 64 + half * 32) : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Pack8 pk, pl;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 v = unpack2(acc[4 * j + e]);
          const float a = fmaxf(v.x, 0.0f), bb = fmaxf(v.y, 0.0f);
          pk.h2[e] = __floats2bfloat162_rn(a, bb);
          const float2 h = __bfloat1622float2(pk.h2[e]);
          pl.h2[e] = __floats2bfloat162_rn(a - h.x, bb - h.y);
        }
        o4[j] = pk.u;
        if (l4) l4[j] = pl.u;
      }
    }
  }
}

// conv1_1 input gradient. g is dL/d(pre-activation of conv1_1) [nimg,H,W,64] bf16. Four lanes per pixel, 16
// channels each, reduced with shuffles. Chained with d(x*255 - mean)/dx = 255 and the optional sigmoid'.
__global__ void __launch_bounds__(256)
conv1_1_bwd_kernel(const bf16* __restrict__ g, const float* __restrict__ image, int nimg, int H, int W,
                   int is_preimage, const float* __restrict__ w, float* __restrict__ grad_out) {
  __shared__ float sw[27 * 64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long npix = static_cast<long long>(nimg) * H * W;
  long long pix = gid >> 2;
  const int part = static_cast<int>(gid & 3);
  const bool live = pix < npix;
  if (!live) pix = npix - 1;   // keep the whole warp in the shuffles
  const int x = static_cast<int>(pix % W);
  const int y = static_cast<int>((pix / W) % H);
  const long long img_base = (pix / (static_cast<long long>(W) * H)) * H * W;

  float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      // dx[y,x,c] = sum_{r,s,k} g[y - r + 1, x - s + 1, k] * w[r,s,c,k]
      const int yy = y - r + 1, xx = x - s + 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const uint4* g4 = reinterpret_cast<const uint4*>(g + (img_base + static_cast<long long>(yy) * W + xx) * 64 + part * 16);
      const float* w0 = sw + ((r * 3 + s) * 3 + 0) * 64 + part * 16;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        Pack8 pk;
        pk.u = __ldg(g4 + j);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(pk.h2[e]);
          const int k = 8 * j + 2 * e;
          d0 = fmaf(f.x, w0[k], d0);
          d0 = fmaf(f.y, w0[k + 1], d0);
          d1 = fmaf(f.x, w0[64 + k], d1);
          d1 = fmaf(f.y, w0[64 + k + 1], d1);
          d2 = fmaf(f.x, w0[128 + k], d2);
          d2 = fmaf(f.y, w0[128 + k + 1], d2);
        }
      }
    }
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    d0 += __shfl_xor_sync(0xffffffffu, d0, o);
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
  }
  if (live && part < 3) {
    float d = (part == 0) ? d0 : (part == 1 ? d1 : d2);
    d *= 255.0f;
    if (is_preimage) {
      const float sg = sigmoidf_(image[pix * 3 + part]);
      d *= sg * (1.0f - sg);
    }
    grad_out[pix * 3 + part] = d;
  }
}

// ---------------------------------------------------------------------------------------------------------
// 2x2 stride-2 average pooling, even H and W (vgg19.py:90-91; SAME == VALID for even sizes).
__global__ void __launch_bounds__(256)
avgpool2_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ x_lo, int nimg, int H, int W, int C,
                    bf16* __restrict__ out, bf16* __restrict__ out_lo) {
  const int Ho = H >> 1, Wo = W >> 1, C8 = C >> 3;
  const long long total = static_cast<long long>(nimg) * Ho * Wo * C8;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int c8 = static_cast<int>(gid % C8);
  const long long opix = gid / C8;
  const int xo = static_cast<int>(opix % Wo);
  const int yo = static_cast<int>((opix / Wo) % Ho);
  const long long n = opix / (static_cast<long long>(Wo) * Ho);
  const long long base = ((n * H + 2 * yo) * W + 2 * xo) * C + c8 * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const long long o = base + (static_cast<long long>(dy) * W + dx) * C;
      Pack8 pk;
      pk.u = __ldg(reinterpret_cast<const uint4*>(x + o));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(pk.h2[e]);
        acc[2 * e] += f.x;
        acc[2 * e + 1] += f.y;
      }
      if (x_lo) {
        pk.u = __ldg(reinterpret_cast<const uint4*>(x_lo + o));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(pk.h2[e]);
          acc[2 * e] += f.x;
          acc[2 * e + 1] += f.y;
        }
      }
    }
  }
  Pack8 o, ol;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a = acc[2 * e] * 0.25f, b = acc[2 * e + 1] * 0.25f;
    o.h2[e] = __floats2bfloat162_rn(a, b);
    const float2 h = __bfloat1622float2(o.h2[e]);
    ol.h2[e] = __floats2bfloat162_rn(a - h.x, b - h.y);
  }
  *reinterpret_cast<uint4*>(out + opix * C + c8 * 8) = o.u;
  if (out_lo) *reinterpret_cast<uint4*>(out_lo + opix * C + c8 * 8) = ol.u;
}

// AvgPoolGrad fused with the ReluGrad of the conv that feeds the pool.
__global__ void __launch_bounds__(256)
avgpool2_bwd_mask_kernel(const bf16* __restrict__ g_in, const bf16* __restrict__ act, int nimg, int H, int W, int C,
                         bf16* __restrict__ out) {
  const int Ho = H >> 1, Wo = W >> 1, C8 = C >> 3;
  const long long total = static_cast<long long>(nimg) * H * W * C8;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int c8 = static_cast<int>(gid % C8);
  const long long pix = gid / C8;
  const int x = static_cast<int>(pix % W);
  const int y = static_cast<int>((pix / W) % H);
  const long long n = pix / (static_cast<long long>(W) * H);
  Pack8 gi, a, o;
  gi.u = __ldg(reinterpret_cast<const uint4*>(g_in + ((n * Ho + (y >> 1)) * Wo + (x >> 1)) * C + c8 * 8));
  if (act) a.u = __ldg(reinterpret_cast<const uint4*>(act + pix * C + c8 * 8));
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 f = __bfloat1622float2(gi.h2[e]);
    f.x *= 0.25f;
    f.y *= 0.25f;
    if (act) {
      const uint32_t mw = reinterpret_cast<const uint32_t*>(&a.u)[e];     // sign test valid for bf16 and fp16 planes
      if (!pos16_lo(mw)) f.x = 0.0f;
      if (!pos16_hi(mw)) f.y = 0.0f;
    }
    o.h2[e] = __floats2bfloat162_rn(f.x, f.y);
  }
  *reinterpret_cast<uint4*>(out + pix * C + c8 * 8) = o.u;
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, size_t n, bf16* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(x[i]);
}
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ x, size_t n, float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(x[i]);
}

__global__ void bf16pair_to_f32_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, size_t n,
                                       float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}

__global__ void pack_weights_kernel(const float* __restrict__ w, int cin, int cout, bf16* __restrict__ wf,
                                    bf16* __restrict__ wf_lo, bf16* __restrict__ wd) {
  const long long total = 9LL * cin * cout;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // w is HWIO: i = ((tap * cin) + ci) * cout + co
  const int co = static_cast<int>(i % cout);
  const int ci = static_cast<int>((i / cout) % cin);
  const int tap = static_cast<int>(i / (static_cast<long long>(cout) * cin));
  const bf16 v = __float2bfloat16_rn(w[i]);
  if (wf) wf[(static_cast<long long>(tap) * cout + co) * cin + ci] = v;
  if (wf_lo) wf_lo[(static_cast<long long>(tap) * cout + co) * cin + ci] = __float2bfloat16_rn(w[i] - __bfloat162float(v));
  if (wd) wd[(static_cast<long long>(8 - tap) * cin + ci) * cout + co] = v;
}

// f16 + f8 forward packs (conv_common.cuh, kF8*): w16 [tap][cout][cin] fp16 and the correction rows
// wcorr [tap][cout][cin / 64][128 B] = 64 bytes e4m3(W_lo * 4096 sw) followed by 64 bytes e4m3(W * sw), W_lo = W - fp16(W):
// the K order matches the activations' [A8 | Alo8] rows, so that A8 * Wlo8 + Alo8 * W8 = 512 sw (A W_lo + A_lo W).
__global__ void pack_weights_f16f8_kernel(const float* __restrict__ w, int cin, int cout, float sw,
                                          __half* __restrict__ w16, uint8_t* __restrict__ wcorr) {
  const long long total = 9LL * cin * cout;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = static_cast<int>(i % cout);
  const int ci = static_cast<int>((i / cout) % cin);
  const int tap = static_cast<int>(i / (static_cast<long long>(cout) * cin));
  const float x = w[i];
  const __half h = __float2half_rn(x);
  const float lo = x - __half2float(h);
  const long long row = static_cast<long long>(tap) * cout + co;
  w16[row * cin + ci] = h;
  uint8_t* dst = wcorr + row * cin * 2 + static_cast<long long>(ci >> 6) * 128 + (ci & 63);
  dst[0] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(lo * 4096.0f * sw, __NV_SATFINITE, __NV_E4M3));
  dst[64] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(x * sw, __NV_SATFINITE, __NV_E4M3));
}

// fp16 plane (+ optional correction plane: the e4m3 low parts, scaled by kF8ScaleLo) -> fp32
__global__ void f16f8_to_f32_kernel(const __half* __restrict__ hi, const uint8_t* __restrict__ corr, size_t n, int C,
                                    float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = __half2float(hi[i]);
  if (corr) {
    const size_t pix = i / C;
    const int ch = static_cast<int>(i % C);
    const uint8_t b = corr[pix * C * 2 + static_cast<size_t>(ch >> 6) * 128 + 64 + (ch & 63)];
    const __half_raw hr = __nv_cvt_fp8_to_halfraw(b, __NV_E4M3);
    v += __half2float(__half(hr)) * (1.0f / kF8ScaleLo);
  }
  out[i] = v;
}

// The same from torch's OIHW layout [cout][cin][3][3] (the generator's nn.Parameters), with the filter bank zero-padded
// to cout_p outputs (the 64 -> 3 colour layer runs as one 64-channel tile). One thread per element of the forward
// pack (coalesced bf16 writes; the fp32 reads are 36-byte-strided but L2-resident: at most 9.4 MB).
__global__ void pack_weights_oihw_kernel(const float* __restrict__ w, int cin, int cout, int cout_p,
                                         bf16* __restrict__ wf, bf16* __restrict__ wd) {
  const long long total = 9LL * cin * cout_p;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // forward pack index: i = ((tap * cout_p) + co) * cin + ci
  const int ci = static_cast<int>(i % cin);
  const int co = static_cast<int>((i / cin) % cout_p);
  const int tap = static_cast<int>(i / (static_cast<long long>(cin) * cout_p));
  const float x = co < cout ? w[(static_cast<long long>(co) * cin + ci) * 9 + tap] : 0.0f;
  const bf16 v = __float2bfloat16_rn(x);
  if (wf) wf[i] = v;
  if (wd) wd[(static_cast<long long>(8 - tap) * cin + ci) * cout_p + co] = v;
}

// Tiled, pre-swizzled packs for the halo-patch kernel: [tap][K/64][N/64] blocks of 64 rows x 128 bytes laid out
// exactly as a 128B-swizzled shared-memory tile (16-byte chunk j of row r stored at chunk j ^ (r & 7)), so that a
// [BN x 64] weight tile is ONE contiguous BN*128-byte run and lands with a single 1-D bulk copy instead of BN
// strided 128-byte TMA rows. forward: N = cout, K = cin; dgrad: N = cin, K = cout, taps rotated by 180 degrees.
__global__ void pack_weights_tiled_kernel(const float* __restrict__ w, int cin, int cout, bf16* __restrict__ f_hi,
                                          bf16* __restrict__ f_lo, bf16* __restrict__ d_hi) {
  const long long total = 9LL * cin * cout;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = static_cast<int>(i % cout);
  const int ci = static_cast<int>((i / cout) % cin);
  const int tap = static_cast<int>(i / (static_cast<long long>(cout) * cin));
  const float x = w[i];
  const bf16 hi = __float2bfloat16_rn(x);
  auto offset = [](int tap_, int n, int k, int N, int K) -> long long {
    const int r = n & 63, kk = k & 63;
    const long long block = (static_cast<long long>(tap_) * (K / 64) + (k >> 6)) * (N / 64) + (n >> 6);
    return block * 4096 + r * 64 + (((kk >> 3) ^ (r & 7)) << 3) + (kk & 7);     // in bf16 elements
  };
  if (f_hi) f_hi[offset(tap, co, ci, cout, cin)] = hi;
  if (f_lo) f_lo[offset(tap, co, ci, cout, cin)] = __float2bfloat16_rn(x - __bfloat162float(hi));
  if (d_hi) d_hi[offset(8 - tap, ci, co, cin, cout)] = hi;
}

__global__ void pack_conv1_1_dgrad_kernel(const float* __restrict__ w, bf16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over [9][16][64]
  if (i >= 9 * 16 * 64) return;
  const int k = i & 63, c = (i >> 6) & 15, tap = i >> 10;
  float v = 0.0f;
  if (c < 3) v = w[((8 - tap) * 3 + c) * 64 + k];
  out[i] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256)
sum_partials_kernel(const float* __restrict__ part, int count, int stride, float* __restrict__ out) {
  const int img = blockIdx.x;
  const float* p = part + static_cast<size_t>(img) * stride;
  float s = 0.0f;
  for (int i = threadIdx.x; i < count; i += 256) s += p[i];
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    out[img] = t;
  }
}

inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

}  // namespace

int conv1_1_fwd_launch(const float* image, int nimg, int H, int W, int is_preimage, const float* mean3,
                       const float* w, const float* b, bf16* out, bf16* out_lo, cudaStream_t stream, int fmt, bf16* aux) {
  STX_REQUIRE(image && w && b && out && mean3, "conv1_1 fwd: null operand");
  const long long threads = 2LL * nimg * H * ((W + 1) / 2);
  unsigned blocks = blocks_for(threads, 128);
  if (blocks > 148u * 4u * 4u) blocks = 148u * 4u * 4u;      // four resident blocks per SM, ~four rounds each
  STX_TRY(launch_pdl(conv1_1_fwd_kernel, dim3(blocks), dim3(128), 0, stream, image, nimg, H, W,
                     is_preimage, mean3[0], mean3[1], mean3[2], w, b, out, out_lo, fmt, aux));
  count_launch();
  return STX_OK;
}

int conv1_1_bwd_launch(const bf16* g, const float* image, int nimg, int H, int W, int is_preimage, const float* w,
                       float* grad_out, cudaStream_t stream) {
  STX_REQUIRE(g && w && grad_out, "conv1_1 bwd: null operand");
  STX_REQUIRE(!is_preimage || image, "conv1_1 bwd: pre-image mode needs the input");
  const long long threads = 4LL * nimg * H * W;
  STX_TRY(launch_pdl(conv1_1_bwd_kernel, dim3(blocks_for(threads, 256)), dim3(256), 0, stream, g, image, nimg, H, W,
                     is_preimage, w, grad_out));
  count_launch();
  return STX_OK;
}

int avgpool2_fwd_launch(const bf16* x, const bf16* x_lo, int nimg, int H, int W, int C, bf16* out, bf16* out_lo,
                        cudaStream_t stream) {
  STX_REQUIRE(x && out, "avgpool fwd: null operand");
  STX_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "avgpool: H, W must be even and C a multiple of 8");
  const long long total = static_cast<long long>(nimg) * (H / 2) * (W / 2) * (C / 8);
  avgpool2_fwd_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(x, x_lo, nimg, H, W, C, out, out_lo);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int avgpool2_bwd_mask_launch(const bf16* g_in, const bf16* act, int nimg, int H, int W, int C, bf16* out,
                             cudaStream_t stream) {
  STX_REQUIRE(g_in && out, "avgpool bwd: null operand");
  STX_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "avgpool: H, W must be even and C a multiple of 8");
  const long long total = static_cast<long long>(nimg) * H * W * (C / 8);
  avgpool2_bwd_mask_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(g_in, act, nimg, H, W, C, out);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int f32_to_bf16_launch(const float* x, size_t n, bf16* out, cudaStream_t stream) {
  if (n == 0) return STX_OK;
  f32_to_bf16_kernel<<<blocks_for(static_cast<long long>(n), 256), 256, 0, stream>>>(x, n, out);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}
int bf16_to_f32_launch(const bf16* x, size_t n, float* out, cudaStream_t stream) {
  if (n == 0) return STX_OK;
  bf16_to_f32_kernel<<<blocks_for(static_cast<long long>(n), 256), 256, 0, stream>>>(x, n, out);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int bf16pair_to_f32_launch(const bf16* hi, const bf16* lo, size_t n, float* out, cudaStream_t stream) {
  if (n == 0) return STX_OK;
  bf16pair_to_f32_kernel<<<blocks_for(static_cast<long long>(n), 256), 256, 0, stream>>>(hi, lo, n, out);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int pack_weights_launch(const float* w_hwio, int cin, int cout, bf16* w_fwd, bf16* w_fwd_lo, bf16* w_dgrad,
                        cudaStream_t stream) {
  STX_REQUIRE(w_hwio, "pack_weights: null weights");
  const long long total = 9LL * cin * cout;
  pack_weights_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(w_hwio, cin, cout, w_fwd, w_fwd_lo, w_dgrad);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int pack_weights_f16f8_launch(const float* w_hwio, int cin, int cout, float sw, bf16* w16, bf16* wcorr,
                              cudaStream_t stream) {
  STX_REQUIRE(w_hwio && w16 && wcorr, "pack_weights_f16f8: null argument");
  STX_REQUIRE(cin % 64 == 0 && sw > 0.0f, "pack_weights_f16f8: cin must be a multiple of 64 and the scale positive");
  const long long total = 9LL * cin * cout;
  pack_weights_f16f8_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(w_hwio, cin, cout, sw,
                                                                       reinterpret_cast<__half*>(w16),
                                                                       reinterpret_cast<uint8_t*>(wcorr));
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int f16f8_to_f32_launch(const bf16* hi, const bf16* corr, size_t n, int C, float* out, cudaStream_t stream) {
  if (n == 0) return STX_OK;
  f16f8_to_f32_kernel<<<blocks_for(static_cast<long long>(n), 256), 256, 0, stream>>>(
      reinterpret_cast<const __half*>(hi), reinterpret_cast<const uint8_t*>(corr), n, C, out);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int pack_weights_oihw_launch(const float* w_oihw, int cin, int cout, int cout_p, bf16* w_fwd, bf16* w_dgrad,
                             cudaStream_t stream) {
  STX_REQUIRE(w_oihw, "pack_weights_oihw: null weights");
  STX_REQUIRE(cin > 0 && cout > 0 && cout_p >= cout, "pack_weights_oihw: bad channel counts");
  const long long total = 9LL * cin * cout_p;
  pack_weights_oihw_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(w_oihw, cin, cout, cout_p, w_fwd, w_dgrad);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int pack_weights_tiled_launch(const float* w_hwio, int cin, int cout, bf16* f_hi, bf16* f_lo, bf16* d_hi,
                              cudaStream_t stream) {
  STX_REQUIRE(w_hwio, "pack_weights_tiled: null weights");
  STX_REQUIRE(cin % 64 == 0 && cout % 64 == 0, "pack_weights_tiled: channels must be multiples of 64");
  const long long total = 9LL * cin * cout;
  pack_weights_tiled_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(w_hwio, cin, cout, f_hi, f_lo, d_hi);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int pack_conv1_1_dgrad_launch(const float* w27x64, bf16* out, cudaStream_t stream) {
  STX_REQUIRE(w27x64 && out, "pack conv1_1 dgrad: null argument");
  pack_conv1_1_dgrad_kernel<<<(9 * 16 * 64 + 255) / 256, 256, 0, stream>>>(w27x64, out);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int sum_partials_launch(const float* part, int nimg, int count, int stride, float* out, cudaStream_t stream) {
  sum_partials_kernel<<<nimg, 256, 0, stream>>>(part, count, stride, out);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

}  // namespace stx
