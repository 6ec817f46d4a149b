// conv1_1 (3 -> 64 channels) of the f16 + f8 plans on the tensor cores.
//
// Reference: vgg19.py:142-146 (x * 255 - mean) and :171-179 (conv + bias + ReLU), optional sigmoid pre-image
// (gatys/synthesizer.py:59); zero padding applies to the PREPROCESSED image, as in the TF graph.
//
// The CUDA-core kernel (elementwise.cu) spends 243 us on 18 jobs of 256^2 for 0.4 GFLOP per job - 3.5x the time its
// 453 MB of output take at the HBM roof. Here the layer is an implicit GEMM with K = 27 (9 taps x 3 colours, padded
// to 32): the 128 worker threads of a CTA build the im2col rows of a 16 x 8 pixel tile directly in shared memory, in
// the 128-byte-swizzled K-major layout tcgen05 reads (chunk c of row r lives at chunk position c ^ (r & 7)), in the
// same two precisions as every other layer of the plan (conv_common.cuh, kF8*):
//     hi row   : 32 fp16 = fp16(xp)                          -> 2 kind::f16 MMAs (K = 16 each) against fp16(W)
//     corr row : 32 e4m3(xp / 8) | .. | 32 e4m3(lo * 512)    -> 2 kind::f8f6f4 MMAs (K = 32 each) against [W_lo8 | .. | W8]
// Four MMAs per tile (N = 64) - the tensor work is nothing; the kernel is its epilogue (scale + bias + ReLU, fp16 /
// e4m3 / bf16 conversions, stores), which is why it runs four small CTAs per SM instead of one big one.
#include "conv_common.cuh"
#include "ops.h"

namespace stx {

namespace {

constexpr int kRows = 128;                    // pixels of a tile = TMEM lanes
constexpr int kWorkers = 256;                 // two per pixel: thread (r, half) owns output channels [32 half, 32 half + 32)
constexpr int kThreadsTc = kWorkers + 32;     // + the MMA / TMEM-allocation warp
constexpr int kRowBytes = 128;
constexpr int kABytes = kRows * kRowBytes;    // 16 KB: hi rows (first 64 bytes used) / correction rows / one staged plane
constexpr int kWBytes = 64 * kRowBytes;       // 8 KB per weight tile
constexpr int kPatchFloats = 18 * 10 * 3;     // preprocessed halo patch of a tile
constexpr int kSmemTc = 5 * kABytes + 2 * kWBytes + kPatchFloats * 4 + 64 + 1024;   // A hi, A corr, staging hi / corr / bf16

// Byte offset of 16-byte chunk `c` of row `r` in a 128-byte-swizzled tile whose base is 1024-byte aligned.
__host__ __device__ __forceinline__ int sw128(int r, int c) { return r * kRowBytes + ((c ^ (r & 7)) << 4); }

// w27x64 fp32 ((tap * 3 + colour) major) -> [W16 tile | Wcorr tile], both pre-swizzled (16 KB)
__global__ void pack_conv1_1_tc_kernel(const float* __restrict__ w, float sw, uint8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // (co, k), k < 32
  if (i >= 64 * 32) return;
  const int co = i >> 5, k = i & 31;
  const float x = k < 27 ? w[k * 64 + co] : 0.0f;
  const __half h = __float2half_rn(x);
  const float lo = x - __half2float(h);
  *reinterpret_cast<__half*>(out + sw128(co, k >> 3) + (k & 7) * 2) = h;
  uint8_t* corr = out + kWBytes;
  corr[sw128(co, k >> 4) + (k & 15)] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(lo * 4096.0f * sw, __NV_SATFINITE, __NV_E4M3));
  corr[sw128(co, 4 + (k >> 4)) + (k & 15)] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(x * sw, __NV_SATFINITE, __NV_E4M3));
}

__global__ void __launch_bounds__(kThreadsTc, 2)
conv1_1_tc_kernel(const float* __restrict__ image, int is_preimage, float m0, float m1, float m2,
                  const uint8_t* __restrict__ wpack, const __grid_constant__ CUtensorMap tm_hi,
                  const __grid_constant__ CUtensorMap tm_corr, const __grid_constant__ CUtensorMap tm_aux,
                  const ConvKernelParams p, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_hi = smem;
  uint8_t* a_corr = smem + kABytes;
  uint8_t* s_hi = smem + 2 * kABytes;        // staged output planes of the tile, [pixel][128 B], 128-byte swizzle
  uint8_t* s_corr = smem + 3 * kABytes;
  uint8_t* s_aux = smem + 4 * kABytes;
  uint8_t* w16 = smem + 5 * kABytes;
  uint8_t* wcorr = w16 + kWBytes;
  float* patch = reinterpret_cast<float*>(wcorr + kWBytes);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(patch + kPatchFloats + 2);      // 8-byte aligned: 542 floats
  uint64_t* acc_full = a_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5;

  // weights (constants: safe before pdl_wait) and the parts of the A rows no tile ever rewrites
  for (int i = threadIdx.x; i < 2 * kWBytes / 16; i += kThreadsTc)
    reinterpret_cast<uint4*>(w16)[i] = __ldg(reinterpret_cast<const uint4*>(wpack) + i);
  for (int i = threadIdx.x; i < 2 * kABytes / 16; i += kThreadsTc) reinterpret_cast<uint4*>(a_hi)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_corr);
    if (p.aux_out) tma_prefetch_desc(&tm_aux);
    mbar_init(a_full, kWorkers);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 128);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const float mean[3] = {m0, m1, m2};
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, phase ^= 1) {
    const int n = tile / tiles_per_img;
    if (p.active && !__ldg(p.active + n)) {     // a finished job's image (uniform over the block)
      phase ^= 1;                               // undone by the loop increment: the barriers' phase follows processed tiles
      continue;
    }
    const int t2 = tile % tiles_per_img;
    const int h0 = (t2 / p.tiles_w) * 16, w0 = (t2 % p.tiles_w) * 8;
    if (warp < 8) {
      // ---- the preprocessed (18 x 10 pixel x 3 colour) halo patch of the tile, once per tile in shared memory (zero
      // outside the image: the padding applies to the preprocessed image), instead of 27 bounds-checked global loads
      // and preprocessings per thread
      for (int i = threadIdx.x; i < kPatchFloats; i += kWorkers) {
        const int pr = i / 30, rem = i - pr * 30, pc = rem / 3, c = rem - pc * 3;
        const int yy = h0 - 1 + pr, xx = w0 - 1 + pc;
        float v = 0.0f;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
          v = __ldg(image + ((static_cast<size_t>(n) * p.H + yy) * p.W + xx) * 3 + c);
          if (is_preimage) v = 1.0f / (1.0f + __expf(-v));
          v = v * 255.0f - mean[c];
        }
        patch[i] = v;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
      // ---- im2col row of this thread's pixel: k = (dy * 3 + dx) * 3 + colour; half 0 writes the fp16 row, half 1 the
      // correction row
      const int r = threadIdx.x & (kRows - 1), half = threadIdx.x >> 7;
      float xp[32];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const float* prow = patch + (((r >> 3) + dy) * 10 + (r & 7)) * 3;      // 9 consecutive floats: 3 pixels x 3 colours
#pragma unroll
        for (int q = 0; q < 9; ++q) xp[dy * 9 + q] = prow[q];
      }
#pragma unroll
      for (int k = 27; k < 32; ++k) xp[k] = 0.0f;
      if (half == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {        // fp16 row: 4 chunks of 8 values
          Pack8h pk;
#pragma unroll
          for (int e = 0; e < 4; ++e) pk.h2[e] = __floats2half2_rn(xp[8 * c + 2 * e], xp[8 * c + 2 * e + 1]);
          *reinterpret_cast<uint4*>(a_hi + sw128(r, c)) = pk.u;
        }
      } else {
        // correction row: A8 in bytes [0, 32), Alo8 in bytes [64, 96)
        uint32_t a8[8], l8[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {        // 4 values each
          float lo[4];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const __half2 hh = __floats2half2_rn(xp[4 * c + 2 * e], xp[4 * c + 2 * e + 1]);
            const float2 hf = __half22float2(hh);
            lo[2 * e] = (xp[4 * c + 2 * e] - hf.x) * kF8ScaleLo;
            lo[2 * e + 1] = (xp[4 * c + 2 * e + 1] - hf.y) * kF8ScaleLo;
          }
          a8[c] = pack_e4m3x4(xp[4 * c] * kF8ScaleA, xp[4 * c + 1] * kF8ScaleA, xp[4 * c + 2] * kF8ScaleA,
                              xp[4 * c + 3] * kF8ScaleA);
          l8[c] = pack_e4m3x4(lo[0], lo[1], lo[2], lo[3]);
        }
        *reinterpret_cast<uint4*>(a_corr + sw128(r, 0)) = make_uint4(a8[0], a8[1], a8[2], a8[3]);
        *reinterpret_cast<uint4*>(a_corr + sw128(r, 1)) = make_uint4(a8[4], a8[5], a8[6], a8[7]);
        *reinterpret_cast<uint4*>(a_corr + sw128(r, 4)) = make_uint4(l8[0], l8[1], l8[2], l8[3]);
        *reinterpret_cast<uint4*>(a_corr + sw128(r, 5)) = make_uint4(l8[4], l8[5], l8[6], l8[7]);
      }
      fence_proxy_async_smem();              // generic-proxy writes -> visible to the tensor core's reads
      mbar_arrive(a_full);
      // ---- epilogue: thread (r, half) reads TMEM lane r, columns [32 half, 32 half + 32), and stages its 64 output
      // bytes per plane in shared memory; one thread then issues the tile's TMA stores (one box per plane: whole
      // 128-byte lines instead of 16-byte stores to 32 different lines per instruction)
      mbar_wait(acc_full, phase, 500);      // a short wait (four MMAs): spin, do not sleep
      tc_fence_after();
      const uint32_t acc = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
      const int c0 = half * 32;
      float v[32];
      {
        float u[32];
        tmem_ld_32x32(acc + c0, v);
        tmem_ld_32x32(acc + 64 + c0, u);
        tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(b4 + j);
          v[4 * j + 0] = fmaxf(fmaf(u[4 * j + 0], p.corr_scale, v[4 * j + 0]) + bb.x, 0.0f);
          v[4 * j + 1] = fmaxf(fmaf(u[4 * j + 1], p.corr_scale, v[4 * j + 1]) + bb.y, 0.0f);
          v[4 * j + 2] = fmaxf(fmaf(u[4 * j + 2], p.corr_scale, v[4 * j + 2]) + bb.z, 0.0f);
          v[4 * j + 3] = fmaxf(fmaf(u[4 * j + 3], p.corr_scale, v[4 * j + 3]) + bb.w, 0.0f);
        }
      }
      tc_fence_before();     // ordered before this thread's next arrival on a_full: the next tile's MMAs wait for all
      // the previous tile's stores must have finished READING the staging buffers
      if (threadIdx.x == 0) tma_store_wait_read();
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
      uint32_t a8[8], l8[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Pack8h pk;
        Pack8 pb;
        float lo[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = fminf(v[8 * j + 2 * e], 65504.0f), b = fminf(v[8 * j + 2 * e + 1], 65504.0f);
          pk.h2[e] = __floats2half2_rn(a, b);
          pb.h2[e] = __floats2bfloat162_rn(a, b);
          const float2 hh = __half22float2(pk.h2[e]);
          lo[2 * e] = (a - hh.x) * kF8ScaleLo;
          lo[2 * e + 1] = (b - hh.y) * kF8ScaleLo;
        }
        *reinterpret_cast<uint4*>(s_hi + sw128(r, half * 4 + j)) = pk.u;
        *reinterpret_cast<uint4*>(s_aux + sw128(r, half * 4 + j)) = pb.u;
        a8[2 * j] = pack_e4m3x4(v[8 * j] * kF8ScaleA, v[8 * j + 1] * kF8ScaleA, v[8 * j + 2] * kF8ScaleA,
                                v[8 * j + 3] * kF8ScaleA);
        a8[2 * j + 1] = pack_e4m3x4(v[8 * j + 4] * kF8ScaleA, v[8 * j + 5] * kF8ScaleA, v[8 * j + 6] * kF8ScaleA,
                                    v[8 * j + 7] * kF8ScaleA);
        l8[2 * j] = pack_e4m3x4(lo[0], lo[1], lo[2], lo[3]);
        l8[2 * j + 1] = pack_e4m3x4(lo[4], lo[5], lo[6], lo[7]);
      }
      // correction plane row: [A8 x 64 | Alo8 x 64]; this thread's 32 channels = 2 chunks of each half
      *reinterpret_cast<uint4*>(s_corr + sw128(r, half * 2)) = make_uint4(a8[0], a8[1], a8[2], a8[3]);
      *reinterpret_cast<uint4*>(s_corr + sw128(r, half * 2 + 1)) = make_uint4(a8[4], a8[5], a8[6], a8[7]);
      *reinterpret_cast<uint4*>(s_corr + sw128(r, 4 + half * 2)) = make_uint4(l8[0], l8[1], l8[2], l8[3]);
      *reinterpret_cast<uint4*>(s_corr + sw128(r, 4 + half * 2 + 1)) = make_uint4(l8[4], l8[5], l8[6], l8[7]);
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
      if (threadIdx.x == 0) {
        tma_store_4d(&tm_hi, s_hi, 0, w0, h0, n);
        if (p.out_lo) tma_store_4d(&tm_corr, s_corr, 0, w0, h0, n);
        if (p.aux_out) tma_store_4d(&tm_aux, s_aux, 0, w0, h0, n);
        tma_store_commit();
      }
    } else {
      // ---- MMA issuer (whole warp walks the loop, one elected lane issues: see conv_patch.cu)
      mbar_wait(a_full, phase, 510);
      tc_fence_after();
      if (elect_one()) {
        constexpr uint32_t idesc16 = make_idesc(128, 64, kFmtF16, kFmtF16, 0, 0);
        constexpr uint32_t idesc8 = make_idesc(128, 64, kFmtE4M3, kFmtE4M3, 0, 0);
        constexpr uint32_t hi = desc_hi_sw128(1024);
        const uint32_t a16 = desc_lo_base(smem_u32(a_hi)), b16 = desc_lo_base(smem_u32(w16));
        const uint32_t a8 = desc_lo_base(smem_u32(a_corr)), b8 = desc_lo_base(smem_u32(wcorr));
        umma_bf16(tmem_base, desc_pack(a16, hi), desc_pack(b16, hi), idesc16, 0);           // k = 0..15
        umma_bf16(tmem_base, desc_pack(a16 + 2, hi), desc_pack(b16 + 2, hi), idesc16, 1);   // k = 16..31
        umma_f8(tmem_base + 64, desc_pack(a8, hi), desc_pack(b8, hi), idesc8, 0);           // A8 * Wlo8
        umma_f8(tmem_base + 64, desc_pack(a8 + 4, hi), desc_pack(b8 + 4, hi), idesc8, 1);   // Alo8 * W8
        umma_commit(acc_full);
      }
      __syncwarp();
    }
  }
  if (threadIdx.x == 0) tma_store_wait_all();      // the stores of the last tile must be complete before the CTA exits
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 128);
}

}  // namespace

int pack_conv1_1_tc_launch(const float* w27x64, float sw, bf16* out16k, cudaStream_t stream) {
  STX_REQUIRE(w27x64 && out16k && sw > 0.0f, "pack_conv1_1_tc: bad argument");
  STX_CUDA(cudaMemsetAsync(out16k, 0, 2 * kWBytes, stream));
  pack_conv1_1_tc_kernel<<<(64 * 32 + 255) / 256, 256, 0, stream>>>(w27x64, sw, reinterpret_cast<uint8_t*>(out16k));
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int conv1_1_tc_launch(const float* image, int nimg, int H, int W, int is_preimage, const float* mean3, const bf16* wpack,
                      float corr_scale, const float* bias, bf16* out, bf16* out_corr, bf16* aux, cudaStream_t stream,
                      const int* active) {
  STX_REQUIRE(image && wpack && bias && out && mean3, "conv1_1 (tensor cores): null operand");
  static bool attr_set = false;
  if (!attr_set) {
    STX_CUDA(cudaFuncSetAttribute(conv1_1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTc));
    attr_set = true;
  }
  ConvKernelParams p{};
  p.H = H;
  p.W = W;
  p.nimg = nimg;
  p.cout = 64;
  p.tiles_w = (W + 7) / 8;
  p.tiles_h = (H + 15) / 16;
  p.relu = 1;
  p.pad = 1;
  p.bias = bias;
  p.out = out;
  p.out_lo = out_corr;
  p.aux_out = aux;
  p.corr_scale = corr_scale;
  p.active = active;
  // store maps: one 8 x 16 pixel x 64 channel box per tile and plane (2-byte elements; the correction plane's 128
  // bytes per pixel are 64 such elements), 128-byte swizzle as staged
  const uint64_t dims[4] = {64u, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(nimg)};
  const uint32_t box[4] = {64u, 8u, 16u, 1u};
  CUtensorMap tm_hi, tm_corr, tm_aux;
  STX_TRY(encode_tmap_bf16(&tm_hi, out, 4, dims, box));
  STX_TRY(encode_tmap_bf16(&tm_corr, out_corr ? out_corr : out, 4, dims, box));
  STX_TRY(encode_tmap_bf16(&tm_aux, aux ? aux : out, 4, dims, box));
  const int total = p.tiles_w * p.tiles_h * nimg;
  const int grid = total < 148 * 2 ? total : 148 * 2;
  STX_TRY(launch_pdl(conv1_1_tc_kernel, dim3(grid), dim3(kThreadsTc), static_cast<size_t>(kSmemTc), stream, image,
                     is_preimage, mean3[0], mean3[1], mean3[2], reinterpret_cast<const uint8_t*>(wpack), tm_hi, tm_corr,
                     tm_aux, p, total));
  count_launch();
  return STX_OK;
}

}  // namespace stx
