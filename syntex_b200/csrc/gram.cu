// Gram matrix G = F^T F / (HW + eps) of an NHWC bf16 feature map (reference: texture_utils/utils.py:8-21) as a
// symmetric split-K tensor-core contraction, plus the fused style-loss / difference kernel (utils.py:44-49).
//
// F is [HW, C] with C contiguous, so both MMA operands are MN-major: a TMA box of 64 pixels x 64 channels lands
// as 64 rows of 128 B (128-byte swizzle) and is consumed with MN-major shared-memory descriptors
// (LBO = distance between 64-channel chunks, SBO = 8 pixel rows). Only the upper-triangle 128x128 tiles are
// computed; diagonal tiles reuse the A stage as the B operand. Each CTA reduces one slab of pixels and writes
// an fp32 partial tile; gram_finish_kernel sums the slabs in a fixed order (bit-reproducible, no atomics),
// mirrors the triangle and produces the loss terms.
#include "ops.h"
#include "ptx.cuh"

namespace stx {

int g_debug_gram_lbo = 0;   // 0 = default; set through stx_debug_set for descriptor bring-up
int g_debug_gram_sbo = 0;

namespace {

// Pipeline: 6 stages of 32 KiB. C >= 128: a stage is 64 pixels x (2 A chunks + 2 B chunks) of 64 channels; C == 64:
// 128 pixels x (1 A chunk [+ 1 B chunk in cross mode]). The conv1_1 / pool1 Grams are pure streaming (8 MB in, 16 KB
// out per image): with 4 stages of 8 KiB in flight per SM they ran at 1.6 TB/s, a quarter of HBM speed.
constexpr int kStages = 6;
constexpr int kSlotBytes = 32 * 1024;
__host__ __device__ constexpr int rows_per_stage(int C) { return C == 64 ? 128 : 64; }
constexpr int kThreads = 192;
constexpr int kSmemTotal = kStages * kSlotBytes + 128 + 1024;

struct GramKernelParams {
  int C, splits, iters_per_split, total_iters, tiles_per_dim;
  int cross;   // 1 = A^T B of two different maps (all T x T tiles, B operand from tmU); 0 = symmetric F^T F
  int f16;     // the feature planes are fp16 (f16+f8 plans)
  uint32_t lbo, sbo;
  float* partial;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gram_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmU, const GramKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kStageBytes = kSlotBytes;
  constexpr int kRowsPerStage = rows_per_stage(BN == 64 ? 64 : 128);   // pixels (K) per pipeline stage
  constexpr int kChunkBytes = kRowsPerStage * 128;                     // one 64-channel chunk
  constexpr int kBOffset = (BN == 64 ? 1 : 2) * kChunkBytes;           // B chunks follow the A chunks
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* accum_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Upper-triangle tile (ti <= tj).
  int ti = 0, tj = 0;
  if (p.cross) {
    ti = blockIdx.x / p.tiles_per_dim;
    tj = blockIdx.x % p.tiles_per_dim;
  } else {
    int t = blockIdx.x;
    int row_len = p.tiles_per_dim;
    while (t >= row_len) {
      t -= row_len;
      ++ti;
      --row_len;
    }
    tj = ti + t;
  }
  const int split = blockIdx.y;
  const int img = blockIdx.z;
  const int m0 = ti * 128, n0 = tj * 128;
  const bool diag = (ti == tj) && !p.cross;
  const int a_chunks = (p.C >= 128) ? 2 : 1;
  const int b_chunks = diag ? 0 : BN / 64;

  int it_begin = split * p.iters_per_split;
  int it_end = min(it_begin + p.iters_per_split, p.total_iters);
  const int iters = max(it_end - it_begin, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmF);
    if (p.cross) tma_prefetch_desc(&tmU);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // whole warp in the loop, one elected lane issuing (keeps the operands provably warp-uniform)
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1, 100 + s);
      uint8_t* a_dst = smem + s * kStageBytes;
      uint8_t* b_dst = a_dst + kBOffset;
      const int row = (it_begin + it) * kRowsPerStage;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[s], (a_chunks + b_chunks) * kChunkBytes);
        for (int c = 0; c < a_chunks; ++c) tma_load_3d(a_dst + c * kChunkBytes, &tmF, &full_bar[s], m0 + 64 * c, row, img);
        for (int c = 0; c < b_chunks; ++c)
          tma_load_3d(b_dst + c * kChunkBytes, p.cross ? &tmU : &tmF, &full_bar[s], n0 + 64 * c, row, img);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // both operands are feature planes: fp16 in the f16+f8 plans (p.f16), bf16 otherwise
    const uint32_t idesc = p.f16 ? make_idesc(128, BN, kFmtF16, kFmtF16, 1, 1) : make_idesc_bf16(128, BN, 1, 1);
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(&full_bar[s], ph, 200 + s);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
      const uint32_t b_addr = diag ? a_addr : a_addr + kBOffset;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kRowsPerStage / 16; ++k) {
          // 16 pixel rows per MMA: advance 16 * 128 B along K
          const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 2048, p.lbo, p.sbo);
          const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * 2048, p.lbo, p.sbo);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (it | k) != 0);
        }
        umma_commit(&empty_bar[s]);
        if (it == iters - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int i = m0 + r;
    float* dst = p.partial + ((static_cast<size_t>(img) * p.splits + split) * p.C + i) * p.C + n0;
    if (iters > 0) {
      mbar_wait(accum_bar, 0, 300);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      if (iters > 0) {
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.0f;
      }
      if (i < p.C) {
        float4* d4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        if (!p.cross && ti != tj) {
          // Mirror of an off-diagonal tile, written here (lane = row: every store of the warp is one coalesced
          // 128-byte line) so that the finish kernel never has to read a slab transposed (stride-C gathers made it
          // move 8x the bytes for 3/8 of the elements at C = 512).
          float* mir = p.partial + ((static_cast<size_t>(img) * p.splits + split) * p.C + n0 + c0) * p.C + i;
#pragma unroll
          for (int j = 0; j < 32; ++j) mir[static_cast<size_t>(j) * p.C] = v[j];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, BN);
}

constexpr int kFinishThreads = 256;
constexpr int kFinishElems = 256;                                 // Gram elements per block, scalar path (= one loss partial)
constexpr int kFinishVecElems = 1024;                             // vector path: four per thread = four loss-partial slots
constexpr int kFinishWarps = kFinishThreads / 32;

// 256 consecutive Gram elements per block, 32 per pass: `kw` warps (1, 2, 4 or 8, chosen from the slab count) share
// the slabs of one 32-element group (warp w takes slabs w, w+kw, ...: every read is one coalesced 128-byte line) and
// the 8/kw groups of a pass run side by side; partial sums meet in shared memory in a fixed order. The slab count
// ranges from 4 (pool4, batch 8) to ~300 (conv1_1, batch 1). Deterministic.
__device__ __forceinline__ void gram_finish_block(int block, const float* __restrict__ partial, int C, int splits,
                                                  int kw, int cross, float norm, float* __restrict__ G_out,
                                                  const float* __restrict__ target, long long target_stride,
                                                  float dscale, bf16* __restrict__ dscaled, float loss_scale,
                                                  float* __restrict__ loss_part, int part_stride) {
  __shared__ float red[kFinishWarps][32];
  __shared__ float sq_warp[kFinishWarps];
  const int img = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = kFinishWarps / kw;        // 32-element groups handled per pass
  const int grp = warp / kw, sub = warp % kw;  // this warp's group and slab subset
  const int CC = C * C;
  float sq = 0.0f;
  if (kw == 1) {
    // Few slabs (every layer at batch >= 2): no sharing of slabs between warps, so no shared-memory pass either -
    // four consecutive elements per thread, 128-bit accesses (C * C is a multiple of 4). The scalar path below launched
    // one thread per Gram element: 18 k blocks of 256 threads for pool4 at 18 jobs, 53 us for 63 MB of traffic.
    const int e = block * kFinishVecElems + threadIdx.x * 4;
    if (e < CC) {
      const float* src = partial + static_cast<size_t>(img) * splits * CC + e;
      float4 t = *reinterpret_cast<const float4*>(src);
      for (int k = 1; k < splits; ++k) {
        const float4 u = *reinterpret_cast<const float4*>(src + static_cast<size_t>(k) * CC);
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      const float4 g = make_float4(t.x / norm, t.y / norm, t.z / norm, t.w / norm);
      if (G_out) *reinterpret_cast<float4*>(G_out + static_cast<size_t>(img) * CC + e) = g;
      if (target) {
        const float4 tg = *reinterpret_cast<const float4*>(target + static_cast<size_t>(img) * target_stride + e);
        const float4 d = make_float4(g.x - tg.x, g.y - tg.y, g.z - tg.z, g.w - tg.w);
        if (dscaled) {
          union { uint2 u; __nv_bfloat162 h2[2]; } pk;
          pk.h2[0] = __floats2bfloat162_rn(d.x * dscale, d.y * dscale);
          pk.h2[1] = __floats2bfloat162_rn(d.z * dscale, d.w * dscale);
          *reinterpret_cast<uint2*>(dscaled + static_cast<size_t>(img) * CC + e) = pk.u;
        }
        sq = ((d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w)) * loss_scale;
      }
    }
  } else
  for (int pass = 0; pass < kFinishElems / 32; pass += groups) {
    const int e = block * kFinishElems + (pass + grp) * 32 + lane;
    float s = 0.0f;
    if (e < CC) {
      const int i = e / C, j = e % C;
      // (the Gram kernel writes the upper-triangle tiles and their mirrors: every slab is a full matrix)
      const float* src = partial + static_cast<size_t>(img) * splits * CC + e;
      for (int k = sub; k < splits; k += kw) s += src[static_cast<size_t>(k) * CC];
    }
    red[warp][lane] = s;
    __syncthreads();
    if (sub == 0 && e < CC) {
      float t = 0.0f;
      for (int w = 0; w < kw; ++w) t += red[warp + w][lane];
      const float g = t / norm;
      if (G_out) G_out[static_cast<size_t>(img) * CC + e] = g;
      if (target) {
        const float d = g - target[static_cast<size_t>(img) * target_stride + e];
        if (dscaled) dscaled[static_cast<size_t>(img) * CC + e] = __float2bfloat16_rn(d * dscale);
        sq += d * d * loss_scale;
      }
    }
    __syncthreads();
  }
  if (loss_part) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) sq_warp[warp] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.0f;
#pragma unroll
      for (int w = 0; w < kFinishWarps; ++w) t += sq_warp[w];
      if (kw == 1) {
        // a vector-path block covers four slots of the per-layer partial table (gram_finish_blocks counts 256-element
        // blocks): its sum goes to the first, zeros to the rest
        float* lp = loss_part + static_cast<size_t>(img) * part_stride + block * (kFinishVecElems / kFinishElems);
        const int slots = min(kFinishVecElems / kFinishElems, (CC + kFinishElems - 1) / kFinishElems - block * (kFinishVecElems / kFinishElems));
        for (int j = 0; j < slots; ++j) lp[j] = j == 0 ? t : 0.0f;
      } else {
        loss_part[static_cast<size_t>(img) * part_stride + block] = t;
      }
    }
  }
}


__global__ void __launch_bounds__(kFinishThreads)
gram_finish_kernel(const float* __restrict__ partial, int C, int splits, int kw, int cross, float norm,
                   float* __restrict__ G_out, const float* __restrict__ target, long long target_stride, float dscale,
                   bf16* __restrict__ dscaled, float loss_scale, float* __restrict__ loss_part, int part_stride) {
  gram_finish_block(blockIdx.x, partial, C, splits, kw, cross, norm, G_out, target, target_stride, dscale, dscaled,
                    loss_scale, loss_part, part_stride);
}

// All loss layers of a plan in ONE launch: blockIdx.x runs over the concatenated block ranges of the layers.
__global__ void __launch_bounds__(kFinishThreads) gram_finish_multi_kernel(const GramFinishBatch fb) {
  int k = 0;
  while (k + 1 < fb.n && static_cast<int>(blockIdx.x) >= fb.l[k + 1].block_begin) ++k;
  const GramFinishLayer& f = fb.l[k];
  gram_finish_block(blockIdx.x - f.block_begin, f.partial, f.C, f.splits, f.kw, 0, f.norm, f.G_out, f.target,
                    f.target_stride, f.dscale, f.dscaled, f.loss_scale, f.loss_part, fb.part_stride);
}

}  // namespace

static void gram_geometry(int nimg, int HW, int C, int* ntiles, int* splits, int* iters_per_split, int* total_iters,
                          int cross = 0) {
  const int T = (C + 127) / 128;
  *ntiles = cross ? T * T : T * (T + 1) / 2;
  *total_iters = (HW + rows_per_stage(C) - 1) / rows_per_stage(C);
  // Aim for about one CTA per SM overall (more slabs only make the fixed-order reduction longer), at least 2
  // pipeline iterations per CTA.
  int want = (148 + (*ntiles) * nimg - 1) / ((*ntiles) * nimg);
  int max_splits = (*total_iters + 1) / 2;
  if (max_splits < 1) max_splits = 1;
  int s = want < max_splits ? want : max_splits;
  if (s < 1) s = 1;
  *iters_per_split = (*total_iters + s - 1) / s;
  *splits = (*total_iters + *iters_per_split - 1) / (*iters_per_split);
}

size_t gram_partial_bytes(int nimg, int HW, int C) {
  int ntiles, splits, ips, total;
  gram_geometry(nimg, HW, C, &ntiles, &splits, &ips, &total);
  return static_cast<size_t>(nimg) * splits * C * C * sizeof(float);
}

int gram_prepare(GramOp* op, const bf16* feat, int nimg, int HW, int C, float* partial) {
  STX_REQUIRE(feat && partial, "gram: null operand");
  STX_REQUIRE(C == 64 || C % 128 == 0, "gram: C must be 64 or a multiple of 128");
  STX_REQUIRE(nimg > 0 && HW > 0, "gram: empty feature map");
  op->nimg = nimg;
  op->HW = HW;
  op->C = C;
  op->bn = (C == 64) ? 64 : 128;
  op->cross = 0;
  int total;
  gram_geometry(nimg, HW, C, &op->ntiles, &op->splits, &op->iters_per_split, &total);
  op->partial = partial;
  const uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(HW), static_cast<uint64_t>(nimg)};
  const uint32_t box[3] = {64u, static_cast<uint32_t>(rows_per_stage(C)), 1u};
  return encode_tmap_bf16(&op->tmF, feat, 3, dims, box);
}

// A^T B / (HW + eps) of two feature maps of the same shape (the F^T U contraction of the second-order term of the
// per-layer style gradient). Every 128x128 tile is computed (no symmetry).
size_t gram_cross_partial_bytes(int nimg, int HW, int C) {
  int ntiles, splits, ips, total;
  gram_geometry(nimg, HW, C, &ntiles, &splits, &ips, &total, 1);
  return static_cast<size_t>(nimg) * splits * C * C * sizeof(float);
}

int gram_cross_prepare(GramOp* op, const bf16* a, const bf16* b, int nimg, int HW, int C, float* partial) {
  STX_REQUIRE(a && b && partial, "gram_cross: null operand");
  STX_REQUIRE(C == 64 || C % 128 == 0, "gram_cross: C must be 64 or a multiple of 128");
  STX_REQUIRE(nimg > 0 && HW > 0, "gram_cross: empty feature map");
  op->nimg = nimg;
  op->HW = HW;
  op->C = C;
  op->bn = (C == 64) ? 64 : 128;
  op->cross = 1;
  int total;
  gram_geometry(nimg, HW, C, &op->ntiles, &op->splits, &op->iters_per_split, &total, 1);
  op->partial = partial;
  const uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(HW), static_cast<uint64_t>(nimg)};
  const uint32_t box[3] = {64u, static_cast<uint32_t>(rows_per_stage(C)), 1u};
  STX_TRY(encode_tmap_bf16(&op->tmF, a, 3, dims, box));
  return encode_tmap_bf16(&op->tmU, b, 3, dims, box);
}

int gram_launch(const GramOp& op, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    STX_CUDA(cudaFuncSetAttribute(gram_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    STX_CUDA(cudaFuncSetAttribute(gram_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    attr_set = true;
  }
  GramKernelParams p;
  p.C = op.C;
  p.splits = op.splits;
  p.iters_per_split = op.iters_per_split;
  p.total_iters = (op.HW + rows_per_stage(op.C) - 1) / rows_per_stage(op.C);
  p.tiles_per_dim = (op.C + 127) / 128;
  p.lbo = g_debug_gram_lbo ? static_cast<uint32_t>(g_debug_gram_lbo) : static_cast<uint32_t>(rows_per_stage(op.C) * 128);
  p.sbo = g_debug_gram_sbo ? static_cast<uint32_t>(g_debug_gram_sbo) : 1024u;
  p.partial = op.partial;
  p.cross = op.cross;
  p.f16 = op.f16;
  dim3 grid(op.ntiles, op.splits, op.nimg);
  if (op.bn == 64)
    gram_kernel<64><<<grid, kThreads, kSmemTotal, stream>>>(op.tmF, op.cross ? op.tmU : op.tmF, p);
  else
    gram_kernel<128><<<grid, kThreads, kSmemTotal, stream>>>(op.tmF, op.cross ? op.tmU : op.tmF, p);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

// blocks a layer's finish occupies in a launch: the vector path (kw == 1) covers 1024 elements per block
static int finish_grid_blocks(int C, int kw) {
  const int per = kw == 1 ? kFinishVecElems : kFinishElems;
  return (C * C + per - 1) / per;
}

static int finish_kw(int splits) {
  int kw = 1;
  while (kw < kFinishWarps && splits > 8 * kw) kw *= 2;      // keep the serial chain per thread <= ~8..37 slabs
  return kw;
}

int gram_finish_batch_add(GramFinishBatch* fb, const GramOp& op, float eps, float* G_out, const float* target,
                          long long target_stride, float dscale, bf16* dscaled, float loss_scale, float* loss_part) {
  STX_REQUIRE(fb->n < kMaxFinishLayers, "gram_finish_batch: too many layers");
  GramFinishLayer& f = fb->l[fb->n];
  f.partial = op.partial;
  f.C = op.C;
  f.splits = op.splits;
  f.kw = finish_kw(op.splits);
  f.norm = static_cast<float>(op.HW) + eps;
  f.G_out = G_out;
  f.target = target;
  f.target_stride = target_stride;
  f.dscale = dscale;
  f.dscaled = dscaled;
  f.loss_scale = loss_scale;
  f.loss_part = loss_part;
  f.block_begin = fb->total_blocks;
  fb->total_blocks += finish_grid_blocks(op.C, f.kw);
  fb->nimg = op.nimg;
  fb->n += 1;
  return STX_OK;
}

int gram_finish_batch_launch(const GramFinishBatch& fb, cudaStream_t stream) {
  if (fb.n == 0) return STX_OK;
  dim3 grid(fb.total_blocks, fb.nimg);
  gram_finish_multi_kernel<<<grid, kFinishThreads, 0, stream>>>(fb);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

int gram_finish_blocks(int C) { return (C * C + kFinishElems - 1) / kFinishElems; }

int gram_finish_launch(const GramOp& op, float eps, float* G_out, const float* target, long long target_stride,
                       float dscale, bf16* dscaled, float loss_scale, float* loss_part, int part_stride,
                       cudaStream_t stream) {
  // fp32 arithmetic as in the reference: normalizer = float32(HW) + float32(eps)  (utils.py:15, :20)
  const float norm = static_cast<float>(op.HW) + eps;
  const int kw = finish_kw(op.splits);
  dim3 grid(finish_grid_blocks(op.C, kw), op.nimg);
  gram_finish_kernel<<<grid, kFinishThreads, 0, stream>>>(op.partial, op.C, op.splits, kw, op.cross, norm, G_out, target,
                                                          target_stride, dscale, dscaled, loss_scale, loss_part, part_stride);
  STX_CUDA(cudaGetLastError());
  count_launch();
  return STX_OK;
}

}  // namespace stx
