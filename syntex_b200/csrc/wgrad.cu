// Filter gradient of a 3x3 convolution (the PO generator's trainable layers; the
// reference gets it from tf.gradients over tensorpack Conv2D, po/progressive_model.py:171-175):
//
//     dW[tap][ci][co] = sum over images and output pixels (h, w) of  X[h + r - pad, w + s - pad][ci] * dY[h, w][co]
//
// as a split-K tcgen05 contraction over the pixels. Both operands are NHWC, i.e. MN-major with the pixel as K - the
// same shape of problem as the Gram matrix (gram.cu) - so a TMA box of bw x bh pixels x 64 channels lands as 64 rows
// of 128 B (128-byte swizzle) and feeds the MMA directly; the tap shift is nothing but the box coordinate of X
// (out-of-range pixels read zeros when pad = 1; pad = 0 addresses an explicitly padded input).
//
// The M axis of the accumulators is a flat list of "A chunks" q = tap * (Cin / 64) + ci_chunk (64 rows each): a CTA
// owns six consecutive chunks = three 128-row accumulators, so that one staged dY tile (the N operand) is reused by
// twelve MMAs, and a 64-channel layer still fills M = 128 (two taps stacked). Each CTA reduces one slab of pixels
// into an fp32 partial tile; wgrad_reduce_kernel sums the slabs in a fixed order (bit-reproducible, no atomics).
// The partial layout [split][tap][ci][co] makes the result HWIO, the layout of texture_utils/caffe2npz.py:16-18.
#include "ops.h"
#include "ptx.cuh"

namespace stx {

namespace {

constexpr int kWgStages = 3;
constexpr int kWgAChunks = 6;                       // 64-row A chunks per CTA (three M = 128 accumulators)
constexpr int kWgChunkBytes = 64 * 128;             // 64 pixels x 64 channels bf16
constexpr int kWgStageBytes = 64 * 1024;            // 6 A chunks + up to 2 B chunks
constexpr int kWgThreads = 192;                     // warp 0: TMA, warp 1: MMA, warps 2..5: epilogue
constexpr int kWgSmemTotal = kWgStages * kWgStageBytes + 128 + 1024;
constexpr int kWgMaxGroups = 16;                    // 9 * (Cin / 64) / 6 groups: Cin <= 640

struct WgradKernelParams {
  int cchunks;          // Cin / 64
  int nq;               // 9 * cchunks: number of A chunks
  int cout;
  int pad;              // 0: X is explicitly padded (H + 2 rows); 1: taps reach outside X and read zeros
  int bw_log2, bh_log2; // pixel box of one K stage (bw * bh = 64)
  int tiles_w, tiles_h;
  int total_iters;      // nimg * tiles_h * tiles_w
  int groups;           // CTA groups along M (six A chunks each)
  // Pixel slabs per group: a group with fewer valid chunks (the tail of the chunk list: 3 of 6 for a 64-channel
  // layer) stages fewer bytes per iteration, so it gets proportionally fewer, longer slabs and every CTA of the launch
  // finishes at the same time. split_begin[g] = first blockIdx.x of group g (prefix sums, groups + 1 entries).
  int split_begin[kWgMaxGroups + 1];
  int iters_per_split[kWgMaxGroups];
  float* partial;       // [max splits][nq * 64][cout]
};

template <int BN>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
             const WgradKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kBOffset = kWgAChunks * kWgChunkBytes;
  constexpr int kBChunks = BN / 64;
  constexpr int kTmemCols = BN == 128 ? 512 : 256;  // three accumulators of BN columns
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
  uint64_t* empty_bar = full_bar + kWgStages;
  uint64_t* accum_bar = empty_bar + kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int grp = 0;
  while (grp + 1 < p.groups && static_cast<int>(blockIdx.x) >= p.split_begin[grp + 1]) ++grp;
  const int q0 = grp * kWgAChunks;
  const int n0 = blockIdx.y * BN;
  const int split = blockIdx.x - p.split_begin[grp];
  int valid = p.nq - q0;
  if (valid > kWgAChunks) valid = kWgAChunks;
  const int n_acc = (valid + 1) >> 1;

  const int it_begin = split * p.iters_per_split[grp];
  const int it_end = min(it_begin + p.iters_per_split[grp], p.total_iters);
  const int iters = max(it_end - it_begin, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp walks, one lane issues)
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    for (int it = 0; it < iters; ++it) {
      const int s = it % kWgStages;
      const uint32_t ph = (it / kWgStages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1, 500 + s);
      const int g = it_begin + it;
      const int img = g / tiles_per_img;
      const int rem = g - img * tiles_per_img;
      const int th = rem / p.tiles_w;
      const int tw = rem - th * p.tiles_w;
      const int w0 = tw << p.bw_log2, h0 = th << p.bh_log2;
      uint8_t* a_dst = smem + s * kWgStageBytes;
      uint8_t* b_dst = a_dst + kBOffset;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[s], (valid + kBChunks) * kWgChunkBytes);
        for (int j = 0; j < valid; ++j) {
          const int q = q0 + j;
          const int tap = q / p.cchunks;
          const int cc = q - tap * p.cchunks;
          const int r = tap / 3, sx = tap - 3 * r;
          tma_load_4d(a_dst + j * kWgChunkBytes, &tmX, &full_bar[s], cc * 64, w0 + sx - p.pad, h0 + r - p.pad, img);
        }
        for (int c = 0; c < kBChunks; ++c)
          tma_load_4d(b_dst + c * kWgChunkBytes, &tmY, &full_bar[s], n0 + 64 * c, w0, h0, img);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
    for (int it = 0; it < iters; ++it) {
      const int s = it % kWgStages;
      const uint32_t ph = (it / kWgStages) & 1;
      mbar_wait(&full_bar[s], ph, 520 + s);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * kWgStageBytes);
      const uint32_t b_addr = a_addr + kBOffset;
      if (elect_one()) {
        for (int j = 0; j < n_acc; ++j) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // 16 pixel rows per MMA: advance 16 * 128 B along K; the two 64-row halves of M (and of N) are
            // kWgChunkBytes apart (LBO)
            const uint64_t adesc = make_smem_desc_sw128(a_addr + j * 2 * kWgChunkBytes + k * 2048, kWgChunkBytes, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * 2048, kWgChunkBytes, 1024);
            umma_bf16(tmem_base + j * BN, adesc, bdesc, idesc, (it | k) != 0);
          }
        }
        umma_commit(&empty_bar[s]);
        if (it == iters - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue: TMEM -> fp32 partial slab
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    if (iters > 0) {
      mbar_wait(accum_bar, 0, 540);
      tc_fence_after();
    }
    for (int j = 0; j < n_acc; ++j) {
      const int chunk = q0 + 2 * j + (r >> 6);
      const bool live = chunk < p.nq;
      float* dst = p.partial + ((static_cast<size_t>(split) * p.nq + chunk) * 64 + (r & 63)) * p.cout + n0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        if (iters > 0) {
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + j * BN + c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0.0f;
        }
        if (live) {
          float4* d4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
          for (int e = 0; e < 8; ++e) d4[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

struct WgradSplits {
  int n[kWgMaxGroups];   // slabs written by group g
};

// out[e] = sum_s partial[s][e], the slabs of the element's group in a fixed order.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float4* __restrict__ partial, size_t count4, int cout4, const WgradSplits sp,
                    float4* __restrict__ out) {
  for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < count4;
       e += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int chunk = static_cast<int>(e / (static_cast<size_t>(64) * cout4));
    const int splits = sp.n[chunk / kWgAChunks];
    float4 a = partial[e];
    for (int s = 1; s < splits; ++s) {
      const float4 b = partial[static_cast<size_t>(s) * count4 + e];
      a.x += b.x;
      a.y += b.y;
      a.z += b.z;
      a.w += b.w;
    }
    out[e] = a;
  }
}

// The same reduction written out in torch's OIHW layout: e runs over the HWIO partial layout (coalesced float4 reads
// along co), each of the four sums goes to out[(co * cin + ci) * 9 + tap] for co < cout_real.
__global__ void __launch_bounds__(256)
wgrad_reduce_oihw_kernel(const float4* __restrict__ partial, size_t count4, int cout4, int cin, int cout_real,
                         const WgradSplits sp, float* __restrict__ out) {
  for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < count4;
       e += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = e / cout4;                        // tap * cin + ci
    const int co0 = static_cast<int>(e - row * cout4) * 4;
    if (co0 >= cout_real) continue;
    const int chunk = static_cast<int>(row >> 6);
    const int splits = sp.n[chunk / kWgAChunks];
    float4 a = partial[e];
    for (int s = 1; s < splits; ++s) {
      const float4 b = partial[static_cast<size_t>(s) * count4 + e];
      a.x += b.x;
      a.y += b.y;
      a.z += b.z;
      a.w += b.w;
    }
    const int tap = static_cast<int>(row / cin), ci = static_cast<int>(row - static_cast<size_t>(tap) * cin);
    const float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (co0 + j < cout_real) out[(static_cast<size_t>(co0 + j) * cin + ci) * 9 + tap] = v[j];
  }
}

struct WgradGeom {
  int bw, bh, tiles_w, tiles_h, total_iters, groups, nblocks, bn;
  int splits[kWgMaxGroups], iters_per_split[kWgMaxGroups], max_splits, total_splits;
};

int wgrad_geometry(int nimg, int H, int W, int cin, int cout, WgradGeom* g) {
  STX_REQUIRE(nimg > 0 && H > 0 && W > 0, "wgrad: empty tensor");
  STX_REQUIRE(cin % 64 == 0 && cout % 64 == 0, "wgrad: channels must be multiples of 64");
  int bw = 64;
  while (bw > 1 && (bw >> 1) >= W) bw >>= 1;     // smallest power of two >= W, capped at 64
  g->bw = bw;
  g->bh = 64 / bw;
  g->tiles_w = (W + g->bw - 1) / g->bw;
  g->tiles_h = (H + g->bh - 1) / g->bh;
  g->total_iters = nimg * g->tiles_w * g->tiles_h;
  const int nq = 9 * (cin / 64);
  g->groups = (nq + kWgAChunks - 1) / kWgAChunks;
  STX_REQUIRE(g->groups <= kWgMaxGroups, "wgrad: too many input channels");
  g->bn = (cout % 128 == 0) ? 128 : 64;
  g->nblocks = cout / g->bn;
  // About one CTA per SM over the launch, shared among the groups in proportion to the bytes a group stages per
  // iteration (its valid A chunks + the dY chunks): the kernel is bound by L2 -> SMEM delivery.
  int weight[kWgMaxGroups], wsum = 0;
  for (int i = 0; i < g->groups; ++i) {
    int valid = nq - i * kWgAChunks;
    if (valid > kWgAChunks) valid = kWgAChunks;
    weight[i] = valid + g->bn / 64;
    wsum += weight[i];
  }
  const int budget = (148 + g->nblocks - 1) / g->nblocks;      // CTAs along x
  g->max_splits = 0;
  g->total_splits = 0;
  for (int i = 0; i < g->groups; ++i) {
    int want = (budget * weight[i] + wsum / 2) / wsum;
    if (want > g->total_iters) want = g->total_iters;
    if (want < 1) want = 1;
    g->iters_per_split[i] = (g->total_iters + want - 1) / want;
    g->splits[i] = (g->total_iters + g->iters_per_split[i] - 1) / g->iters_per_split[i];
    g->total_splits += g->splits[i];
    if (g->splits[i] > g->max_splits) g->max_splits = g->splits[i];
  }
  return STX_OK;
}

}  // namespace

size_t wgrad_workspace_bytes(int nimg, int H, int W, int cin, int cout) {
  WgradGeom g;
  if (wgrad_geometry(nimg, H, W, cin, cout, &g) != STX_OK) return 0;
  return static_cast<size_t>(g.max_splits) * 9 * cin * cout * sizeof(float);
}

static int wgrad_run(const bf16* x, const bf16* dy, int nimg, int H, int W, int cin, int cout, int pad, float* dw,
                     int cout_real_or_0, float* workspace, cudaStream_t stream);

int wgrad_launch(const bf16* x, const bf16* dy, int nimg, int H, int W, int cin, int cout, int pad, float* dw_hwio,
                 float* workspace, cudaStream_t stream) {
  return wgrad_run(x, dy, nimg, H, W, cin, cout, pad, dw_hwio, 0, workspace, stream);
}

int wgrad_oihw_launch(const bf16* x, const bf16* dy, int nimg, int H, int W, int cin, int cout, int pad, int cout_real,
                      float* dw_oihw, float* workspace, cudaStream_t stream) {
  STX_REQUIRE(cout_real > 0 && cout_real <= cout, "wgrad: cout_real out of range");
  return wgrad_run(x, dy, nimg, H, W, cin, cout, pad, dw_oihw, cout_real, workspace, stream);
}

static int wgrad_run(const bf16* x, const bf16* dy, int nimg, int H, int W, int cin, int cout, int pad, float* dw_hwio,
                     int cout_real_or_0, float* workspace, cudaStream_t stream) {
  STX_REQUIRE(x && dy && dw_hwio && workspace, "wgrad: null operand");
  STX_REQUIRE(pad == 0 || pad == 1, "wgrad: pad must be 0 (explicitly padded input) or 1 (zero padding)");
  WgradGeom g;
  STX_TRY(wgrad_geometry(nimg, H, W, cin, cout, &g));
  static bool attr_set = false;
  if (!attr_set) {
    STX_CUDA(cudaFuncSetAttribute(wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemTotal));
    STX_CUDA(cudaFuncSetAttribute(wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemTotal));
    attr_set = true;
  }
  const int Hx = H + 2 - 2 * pad, Wx = W + 2 - 2 * pad;
  CUtensorMap tmX, tmY;
  const uint64_t xdims[4] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(Wx), static_cast<uint64_t>(Hx),
                             static_cast<uint64_t>(nimg)};
  const uint64_t ydims[4] = {static_cast<uint64_t>(cout), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                             static_cast<uint64_t>(nimg)};
  const uint32_t box[4] = {64u, static_cast<uint32_t>(g.bw), static_cast<uint32_t>(g.bh), 1u};
  STX_TRY(encode_tmap_bf16(&tmX, x, 4, xdims, box));
  STX_TRY(encode_tmap_bf16(&tmY, dy, 4, ydims, box));
  WgradKernelParams p;
  p.cchunks = cin / 64;
  p.nq = 9 * p.cchunks;
  p.cout = cout;
  p.pad = pad;
  p.bw_log2 = ilog2(g.bw);
  p.bh_log2 = ilog2(g.bh);
  p.tiles_w = g.tiles_w;
  p.tiles_h = g.tiles_h;
  p.total_iters = g.total_iters;
  p.groups = g.groups;
  WgradSplits sp;
  p.split_begin[0] = 0;
  for (int i = 0; i < kWgMaxGroups; ++i) {
    const bool live = i < g.groups;
    sp.n[i] = live ? g.splits[i] : 0;
    p.iters_per_split[i] = live ? g.iters_per_split[i] : 1;
    p.split_begin[i + 1] = p.split_begin[i] + sp.n[i];
  }
  p.partial = workspace;
  dim3 grid(g.total_splits, g.nblocks, 1);
  if (g.bn == 64)
    wgrad_kernel<64><<<grid, kWgThreads, kWgSmemTotal, stream>>>(tmX, tmY, p);
  else
    wgrad_kernel<128><<<grid, kWgThreads, kWgSmemTotal, stream>>>(tmX, tmY, p);
  STX_CUDA(cudaGetLastError());
  count_launch();
  const size_t count4 = static_cast<size_t>(9) * cin * cout / 4;
  int blocks = static_cast<int>((count4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (cout_real_or_0 > 0)
    wgrad_reduce_oihw_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(workspace), count4, cout / 4,
                                                         cin, cout_real_or_0, sp, dw_hwio);
  else
    wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(workspace), count4, cout / 4, sp,
                                                    reinterpret_cast<float4*>(dw_hwio));
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

}  // namespace stx
