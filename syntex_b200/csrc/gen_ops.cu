// HBM-bound building blocks of the PO generator, NHWC bf16 with fp32 arithmetic:
//   * tf.pad(mode="SYMMETRIC" / "REFLECT") by one pixel and its gradient, optionally fused with ReLU / LeakyReLU
//     (po/progressive_model.py:171-175: pad -> Conv2D(VALID); :77-99 relu before the first conv of a block)
//   * tensorpack InstanceNorm (per image and channel over H x W, biased variance, affine), forward and backward, with
//     the ReLU that follows it and the residual shortcut that is added to it fused into the same pass
//     (po/progressive_model.py:77-121).
// Every kernel moves 16-byte vectors (8 channels) with consecutive threads on consecutive addresses; reductions are
// two-level (per-block partials in a fixed order, then one small finishing kernel): deterministic, no atomics.
#include "ops.h"
#include "ptx.cuh"

namespace stx {

namespace {

union Vec8 {
  uint4 u;
  __nv_bfloat162 h2[4];
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  Vec8 v;
  v.u = u;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __bfloat1622float2(v.h2[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  Vec8 v;
#pragma unroll
  for (int e = 0; e < 4; ++e) v.h2[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return v.u;
}

// Source index of padded position p (0 .. n + 1) along an axis of n pixels.
//   mode 0: SYMMETRIC by one pixel = edge replication (p = 0 -> 0, p = n + 1 -> n - 1)
//   mode 1: REFLECT (p = 0 -> 1, p = n + 1 -> n - 2; needs n >= 2)
__device__ __forceinline__ int pad_src(int p, int n, int mode) {
  if (p == 0) return mode;
  if (p == n + 1) return n - 1 - mode;
  return p - 1;
}

// out [n, H+2, W+2, C] <- act(x [n, H, W, C]) with a one-pixel SYMMETRIC / REFLECT ring.
// act: 0 none, 1 = x > 0 ? x : slope * x  (slope 0 = ReLU, 0.2 = the reference's LeakyReLU).
__global__ void __launch_bounds__(256)
pad_rep_kernel(const uint4* __restrict__ x, int nimg, int H, int W, int c8, int mode, int act, float slope,
               uint4* __restrict__ out) {
  const int Hp = H + 2, Wp = W + 2;
  const size_t total = static_cast<size_t>(nimg) * Hp * Wp * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c8);
    size_t t = i / c8;
    const int wp = static_cast<int>(t % Wp);
    t /= Wp;
    const int hp = static_cast<int>(t % Hp);
    const int n = static_cast<int>(t / Hp);
    const int h = pad_src(hp, H, mode), w = pad_src(wp, W, mode);
    uint4 v = __ldg(x + ((static_cast<size_t>(n) * H + h) * W + w) * c8 + c);
    if (act) {
      float f[8];
      unpack8(v, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = f[e] > 0.0f ? f[e] : slope * f[e];
      v = pack8(f);
    }
    out[i] = v;
  }
}

// Padded positions that read source pixel h along an axis of n pixels: always h + 1, plus one ring position for the
// pixels the ring copies (mode 0: pixel 0 <- ring 0, pixel n-1 <- ring n+1; mode 1: pixel 1 <- ring 0,
// pixel n-2 <- ring n+1). Returns the count (1..3) and fills idx.
__device__ __forceinline__ int pad_readers(int h, int n, int mode, int (&idx)[3]) {
  int k = 0;
  idx[k++] = h + 1;
  if (h == mode) idx[k++] = 0;
  if (h == n - 1 - mode) idx[k++] = n + 1;
  return k;
}

// dx [n, H, W, C] <- gradient w.r.t. the padded tensor gp [n, H+2, W+2, C]: the ring folds back onto the pixels it
// copied; with `x` (the forward input) given, the activation's derivative is applied as well (1 where x > 0, else
// slope).
__global__ void __launch_bounds__(256)
pad_rep_bwd_kernel(const uint4* __restrict__ gp, const uint4* __restrict__ x, int nimg, int H, int W, int c8, int mode,
                   float slope, uint4* __restrict__ dx) {
  const int Wp = W + 2, Hp = H + 2;
  const size_t total = static_cast<size_t>(nimg) * H * W * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c8);
    size_t t = i / c8;
    const int w = static_cast<int>(t % W);
    t /= W;
    const int h = static_cast<int>(t % H);
    const int n = static_cast<int>(t / H);
    int hs[3], ws[3];
    const int nh = pad_readers(h, H, mode, hs), nw = pad_readers(w, W, mode, ws);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
    for (int a = 0; a < nh; ++a) {
      for (int b = 0; b < nw; ++b) {
        float f[8];
        unpack8(__ldg(gp + ((static_cast<size_t>(n) * Hp + hs[a]) * Wp + ws[b]) * c8 + c), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    }
    if (x) {
      float f[8];
      unpack8(__ldg(x + i), f);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (!(f[e] > 0.0f)) acc[e] *= slope;
    }
    dx[i] = pack8(acc);
  }
}

// ---------------------------------------------------------------- instance norm
// Per-(image, channel) sums over the pixels. One block = one slab of pixels of one image; thread t owns the channel
// group t % c8 and walks pixels t / c8, t / c8 + 256 / c8, ... of the slab.
//   MODE 0: a = sum x,                b = sum x^2
//   MODE 1: a = sum g,                b = sum g * xhat       with g = dy * act'(xhat * gamma + beta)
constexpr int kInThreads = 256;

struct InParams {
  int HW, C, c8, slabs, pix_per_slab;
  int relu;             // activation after the affine: 0 none, 1 = v > 0 ? v : slope * v
  float slope;          // 0 = ReLU, 0.2 = LeakyReLU of po/improved_model.py:30-47
  float eps;
  const float* mean;    // [n][C]
  const float* rstd;    // [n][C]
  const float* gamma;   // [C]
  const float* beta;    // [C]
};

template <int MODE>
__global__ void __launch_bounds__(kInThreads)
in_partial_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, const InParams p, float* __restrict__ part) {
  __shared__ float red[2][kInThreads][8 + 1];
  const int img = blockIdx.y, slab = blockIdx.x;
  const int cg = threadIdx.x % p.c8, pl = threadIdx.x / p.c8, lanes = kInThreads / p.c8;
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = b[e] = 0.0f;
  float mu[8], rs[8], ga[8], be[8];
  if (MODE == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ch = cg * 8 + e;
      mu[e] = p.mean[img * p.C + ch];
      rs[e] = p.rstd[img * p.C + ch];
      ga[e] = p.gamma[ch];
      be[e] = p.beta[ch];
    }
  }
  const int p_begin = slab * p.pix_per_slab;
  const int p_end = min(p_begin + p.pix_per_slab, p.HW);
  const uint4* xi = x + static_cast<size_t>(img) * p.HW * p.c8;
  const uint4* di = (MODE == 1) ? dy + static_cast<size_t>(img) * p.HW * p.c8 : nullptr;
#pragma unroll 8
  for (int px = p_begin + pl; px < p_end; px += lanes) {
    float f[8];
    unpack8(__ldg(xi + static_cast<size_t>(px) * p.c8 + cg), f);
    if (MODE == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        a[e] += f[e];
        b[e] += f[e] * f[e];
      }
    } else {
      float g[8];
      unpack8(__ldg(di + static_cast<size_t>(px) * p.c8 + cg), g);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float t = f[e] - mu[e];
        float gg = g[e];
        if (p.relu && !(t * (rs[e] * ga[e]) + be[e] > 0.0f)) gg *= p.slope;   // the forward's expression, bit for bit
        a[e] += gg;
        b[e] += gg * (t * rs[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[0][threadIdx.x][e] = a[e];
    red[1][threadIdx.x][e] = b[e];
  }
  __syncthreads();
  // thread (cg, e) of the first C threads sums the pixel lanes in a fixed order
  for (int ch = threadIdx.x; ch < p.C; ch += kInThreads) {
    const int g8 = ch >> 3, e = ch & 7;
    float sa = 0.0f, sb = 0.0f;
    for (int l = 0; l < lanes; ++l) {
      sa += red[0][l * p.c8 + g8][e];
      sb += red[1][l * p.c8 + g8][e];
    }
    float* dst = part + ((static_cast<size_t>(img) * p.slabs + slab) * p.C + ch) * 2;
    dst[0] = sa;
    dst[1] = sb;
  }
}

// MODE 0: (mean, rstd) from the slab sums; MODE 1: the two sums themselves. One warp per (image, channel): lane l sums
// slabs l, l + 32, ... and the lanes meet in a fixed-order shuffle tree (deterministic).
template <int MODE>
__global__ void __launch_bounds__(256)
in_finish_kernel(const float* __restrict__ part, int nimg, int C, int slabs, int HW, float eps, float* __restrict__ o0,
                 float* __restrict__ o1) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= nimg * C) return;
  const int img = i / C, ch = i - img * C;
  float sa = 0.0f, sb = 0.0f;
#pragma unroll 4
  for (int s = lane; s < slabs; s += 32) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(part + ((static_cast<size_t>(img) * slabs + s) * C + ch) * 2));
    sa += v.x;
    sb += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  if (lane != 0) return;
  if (MODE == 0) {
    const float m = sa / static_cast<float>(HW);
    const float var = fmaxf(sb / static_cast<float>(HW) - m * m, 0.0f);
    o0[i] = m;
    o1[i] = rsqrtf(var + eps);
  } else {
    o0[i] = sa;
    o1[i] = sb;
  }
}

// out = [relu]((x - mean) * (rstd * gamma) + beta) [+ addend]. Grid (slabs, images): a thread keeps its channel group
// for the whole image (the grid stride is a multiple of C / 8), so the eight per-channel constants live in registers
// and the loop body is one 16-byte load, eight FMAs and one 16-byte store.
__global__ void __launch_bounds__(256)
in_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ addend, const InParams p, uint4* __restrict__ out) {
  const int img = blockIdx.y;
  const size_t per_img = static_cast<size_t>(p.HW) * p.c8;
  const size_t first = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const int cg = static_cast<int>(first % p.c8);
  float mu[8], a[8], be[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ch = cg * 8 + e;
    mu[e] = __ldg(p.mean + img * p.C + ch);
    a[e] = __ldg(p.rstd + img * p.C + ch) * __ldg(p.gamma + ch);
    be[e] = __ldg(p.beta + ch);
  }
  const uint4* xi = x + img * per_img;
  const uint4* ai = addend ? addend + img * per_img : nullptr;
  uint4* oi = out + img * per_img;
#pragma unroll 2
  for (size_t i = first; i < per_img; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8(__ldg(xi + i), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      f[e] = (f[e] - mu[e]) * a[e] + be[e];
      if (p.relu) f[e] = f[e] > 0.0f ? f[e] : p.slope * f[e];
    }
    if (ai) {
      float s[8];
      unpack8(__ldg(ai + i), s);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += s[e];
    }
    oi[i] = pack8(f);
  }
}

// dx = rstd * gamma * (g - s1 / HW - xhat * s2 / HW),  g = dy * [relu mask], s1 = sum g, s2 = sum g * xhat
__global__ void __launch_bounds__(256)
in_bwd_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, const InParams p,
                    const float* __restrict__ s1, const float* __restrict__ s2, uint4* __restrict__ dx) {
  const int img = blockIdx.y;
  const size_t per_img = static_cast<size_t>(p.HW) * p.c8;
  const size_t first = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const int cg = static_cast<int>(first % p.c8);
  const float inv = 1.0f / static_cast<float>(p.HW);
  float mu[8], rs[8], a[8], be[8], c1[8], c2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ch = cg * 8 + e;
    mu[e] = __ldg(p.mean + img * p.C + ch);
    rs[e] = __ldg(p.rstd + img * p.C + ch);
    a[e] = rs[e] * __ldg(p.gamma + ch);
    be[e] = __ldg(p.beta + ch);
    c1[e] = inv * __ldg(s1 + img * p.C + ch);
    c2[e] = inv * __ldg(s2 + img * p.C + ch);
  }
  const uint4* xi = x + img * per_img;
  const uint4* di = dy + img * per_img;
  uint4* oi = dx + img * per_img;
#pragma unroll 2
  for (size_t i = first; i < per_img; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float f[8], g[8];
    unpack8(__ldg(xi + i), f);
    unpack8(__ldg(di + i), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float t = f[e] - mu[e];
      float gg = g[e];
      if (p.relu && !(t * a[e] + be[e] > 0.0f)) gg *= p.slope;
      f[e] = a[e] * (gg - c1[e] - (t * rs[e]) * c2[e]);
    }
    oi[i] = pack8(f);
  }
}

int ew_blocks(size_t total) {
  size_t b = (total + 255) / 256;
  const size_t cap = 148 * 16;
  return static_cast<int>(b < cap ? (b ? b : 1) : cap);
}

int in_apply_blocks(int nimg, size_t per_img) {
  size_t b = (per_img + 255) / 256;                      // one vector per thread ...
  const size_t cap = (148 * 16 + nimg - 1) / nimg;       // ... up to ~16 blocks per SM over the batch, then grid-stride
  return static_cast<int>(b < cap ? (b ? b : 1) : cap);
}

int in_slabs(int nimg, int HW) {
  // ~4 blocks per SM over the whole batch (more slabs only lengthen the finishing reduction: with 1184 of them the
  // 8-block finish kernel took 15 us at 512x512, longer than the 13 us streaming pass it follows), >= 64 pixels each
  int slabs = (148 * 4 + nimg - 1) / nimg;
  const int max_slabs = (HW + 63) / 64;
  if (slabs > max_slabs) slabs = max_slabs;
  return slabs < 1 ? 1 : slabs;
}

}  // namespace

int pad_rep_launch(const bf16* x, int nimg, int H, int W, int C, int mode, int act, float slope, bf16* out,
                   cudaStream_t stream) {
  STX_REQUIRE(x && out, "pad: null operand");
  STX_REQUIRE(nimg > 0 && H > 0 && W > 0 && C % 8 == 0, "pad: C must be a multiple of 8");
  STX_REQUIRE(mode == 0 || (mode == 1 && H >= 2 && W >= 2), "pad: mode must be 0 (symmetric) or 1 (reflect, H, W >= 2)");
  const size_t total = static_cast<size_t>(nimg) * (H + 2) * (W + 2) * (C / 8);
  pad_rep_kernel<<<ew_blocks(total), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), nimg, H, W, C / 8, mode, act,
                                                      slope, reinterpret_cast<uint4*>(out));
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int pad_rep_bwd_launch(const bf16* gp, const bf16* x_or_null, int nimg, int H, int W, int C, int mode, float slope,
                       bf16* dx, cudaStream_t stream) {
  STX_REQUIRE(gp && dx, "pad backward: null operand");
  STX_REQUIRE(nimg > 0 && H > 0 && W > 0 && C % 8 == 0, "pad backward: C must be a multiple of 8");
  STX_REQUIRE(mode == 0 || (mode == 1 && H >= 2 && W >= 2), "pad backward: mode must be 0 (symmetric) or 1 (reflect)");
  const size_t total = static_cast<size_t>(nimg) * H * W * (C / 8);
  pad_rep_bwd_kernel<<<ew_blocks(total), 256, 0, stream>>>(reinterpret_cast<const uint4*>(gp),
                                                          reinterpret_cast<const uint4*>(x_or_null), nimg, H, W, C / 8,
                                                          mode, slope, reinterpret_cast<uint4*>(dx));
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

size_t inorm_workspace_bytes(int nimg, int HW, int C) {
  if (nimg <= 0 || HW <= 0 || C <= 0) return 0;
  return static_cast<size_t>(nimg) * in_slabs(nimg, HW) * C * 2 * sizeof(float);
}

static int in_check(int nimg, int HW, int C) {
  STX_REQUIRE(nimg > 0 && HW > 0, "instance norm: empty tensor");
  STX_REQUIRE(C % 8 == 0 && C >= 8 && (kInThreads % (C / 8) == 0) && C <= 8 * kInThreads,
              "instance norm: C must be 8 * a divisor of 256");
  return STX_OK;
}

static InParams in_params(int nimg, int HW, int C, int relu, float slope, float eps, const float* mean,
                          const float* rstd, const float* gamma, const float* beta) {
  InParams p;
  p.slope = slope;
  p.HW = HW;
  p.C = C;
  p.c8 = C / 8;
  p.slabs = in_slabs(nimg, HW);
  p.pix_per_slab = (HW + p.slabs - 1) / p.slabs;
  p.relu = relu;
  p.eps = eps;
  p.mean = mean;
  p.rstd = rstd;
  p.gamma = gamma;
  p.beta = beta;
  return p;
}

int inorm_stats_launch(const bf16* x, int nimg, int HW, int C, float eps, float* mean, float* rstd, float* workspace,
                       cudaStream_t stream) {
  STX_REQUIRE(x && mean && rstd && workspace, "instance norm: null operand");
  STX_TRY(in_check(nimg, HW, C));
  const InParams p = in_params(nimg, HW, C, 0, 0.0f, eps, nullptr, nullptr, nullptr, nullptr);
  in_partial_kernel<0><<<dim3(p.slabs, nimg), kInThreads, 0, stream>>>(reinterpret_cast<const uint4*>(x), nullptr, p,
                                                                      workspace);
  STX_CUDA(cudaGetLastError());
  in_finish_kernel<0><<<(nimg * C + 7) / 8, 256, 0, stream>>>(workspace, nimg, C, p.slabs, HW, eps, mean, rstd);
  STX_CUDA(cudaGetLastError());
  count_launch(2);
  return STX_OK;
}

int inorm_apply_launch(const bf16* x, int nimg, int HW, int C, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, int relu, float slope, const bf16* addend, bf16* out, cudaStream_t stream) {
  STX_REQUIRE(x && mean && rstd && gamma && beta && out, "instance norm: null operand");
  STX_TRY(in_check(nimg, HW, C));
  const InParams p = in_params(nimg, HW, C, relu, slope, 0.0f, mean, rstd, gamma, beta);
  const size_t per_img = static_cast<size_t>(HW) * p.c8;
  in_apply_kernel<<<dim3(in_apply_blocks(nimg, per_img), nimg), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(addend), p, reinterpret_cast<uint4*>(out));
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int inorm_bwd_launch(const bf16* x, const bf16* dy, int nimg, int HW, int C, const float* mean, const float* rstd,
                     const float* gamma, const float* beta, int relu, float slope, float* s1, float* s2, bf16* dx,
                     float* workspace, cudaStream_t stream) {
  STX_REQUIRE(x && dy && mean && rstd && gamma && beta && s1 && s2 && dx && workspace, "instance norm: null operand");
  STX_TRY(in_check(nimg, HW, C));
  const InParams p = in_params(nimg, HW, C, relu, slope, 0.0f, mean, rstd, gamma, beta);
  in_partial_kernel<1><<<dim3(p.slabs, nimg), kInThreads, 0, stream>>>(reinterpret_cast<const uint4*>(x),
                                                                      reinterpret_cast<const uint4*>(dy), p, workspace);
  STX_CUDA(cudaGetLastError());
  in_finish_kernel<1><<<(nimg * C + 7) / 8, 256, 0, stream>>>(workspace, nimg, C, p.slabs, HW, 0.0f, s1, s2);
  STX_CUDA(cudaGetLastError());
  const size_t per_img = static_cast<size_t>(HW) * p.c8;
  in_bwd_apply_kernel<<<dim3(in_apply_blocks(nimg, per_img), nimg), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(dy), p, s1, s2, reinterpret_cast<uint4*>(dx));
  STX_CUDA(cudaGetLastError());
  count_launch(3);
  return STX_OK;
}

// ------------------------------------------------------------------------------------------------ Adam (TF rule)
// tf.train.AdamOptimizer as the reference's trainers use it (po/progressive_model.py:268-271, epsilon 1e-3):
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);  p -= lr_t m / (sqrt(v) + eps)
// ("epsilon hat": eps is added to sqrt(v), NOT to sqrt(v / (1 - b2^t)) as torch.optim.Adam does - with eps = 1e-3 the
// two differ by up to 30x in the effective epsilon over the first thousand steps.)
// lr and the step count t (already incremented, as a float) live on the device so that a captured step replays
// with the current schedule value.
namespace {
__global__ void __launch_bounds__(256) adam_tf_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ m, float* __restrict__ v, size_t n,
                                                      const float* __restrict__ lr, const float* __restrict__ step,
                                                      float b1, float b2, float eps) {
  const float t = *step;
  const float lr_t = *lr * sqrtf(1.0f - powf(b2, t)) / (1.0f - powf(b1, t));
  const size_t i4 = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    float4 pp = *reinterpret_cast<const float4*>(p + i4);
    const float4 gg = *reinterpret_cast<const float4*>(g + i4);
    float4 mm = *reinterpret_cast<const float4*>(m + i4);
    float4 vv = *reinterpret_cast<const float4*>(v + i4);
#define STX_ADAM1(c)                                 \
    mm.c = b1 * mm.c + (1.0f - b1) * gg.c;             \
    vv.c = b2 * vv.c + (1.0f - b2) * gg.c * gg.c;      \
    pp.c -= lr_t * mm.c / (sqrtf(vv.c) + eps);
    STX_ADAM1(x) STX_ADAM1(y) STX_ADAM1(z) STX_ADAM1(w)
#undef STX_ADAM1
    *reinterpret_cast<float4*>(p + i4) = pp;
    *reinterpret_cast<float4*>(m + i4) = mm;
    *reinterpret_cast<float4*>(v + i4) = vv;
  } else {
    for (size_t i = i4; i < n; ++i) {
      const float gi = g[i];
      const float mi = b1 * m[i] + (1.0f - b1) * gi;
      const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      p[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
  }
}
}  // namespace

int adam_tf_launch(float* p, const float* g, float* m, float* v, size_t n, const float* lr_dev, const float* step_dev,
                   float b1, float b2, float eps, cudaStream_t stream) {
  STX_REQUIRE(p && g && m && v && lr_dev && step_dev, "adam: null argument");
  STX_REQUIRE((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
               reinterpret_cast<uintptr_t>(v)) % 16 == 0, "adam: buffers must be 16-byte aligned");
  if (n == 0) return STX_OK;
  const size_t threads = (n + 3) / 4;
  adam_tf_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(p, g, m, v, n, lr_dev, step_dev, b1,
                                                                                  b2, eps);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

}  // namespace stx
