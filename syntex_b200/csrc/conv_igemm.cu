// 3x3 (and 1x1) convolution over NHWC bf16 activations as an implicit GEMM on tcgen05 tensor cores.
//
//   D[M = 128 output pixels, N = BN output channels] = sum over (tap, 64-channel chunk) A_tap * W_tap^T
//
// * A tile  : a BNI x BH x BW patch of pixels (BNI*BH*BW = 128) x 64 input channels, fetched by ONE 4-D TMA box
//             whose (w, h) coordinates are shifted by the tap offset. Halo pixels are out of range for the
//             tensor map and arrive as zeros, which is exactly TF's SAME padding (reference:
//             texture_utils/vgg19.py:176, tf.nn.conv2d(..., padding="SAME")).
// * W tile  : BN x 64 slice of the pre-packed weights [tap][C_out][C_in] (K-major), one 3-D TMA box.
// * Both operands land in shared memory with the 128-byte swizzle and are consumed directly by
//   tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) through shared-memory descriptors; accumulators live in TMEM.
// * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM -> registers ->
//   bias / addend / ReLU / ReLU-mask -> bf16 -> global).
//
// SPLIT mode (forward convolutions of the accurate path): activations and weights are carried as bf16 pairs
// hi + lo (x = hi + lo to ~16 mantissa bits) and each K step issues three MMAs, A_hi*W_hi + A_lo*W_hi + A_hi*W_lo,
// into the same fp32 TMEM accumulator ("bf16x3"). A single bf16 MMA flips the sign of ~1e-3 of the ReLU
// pre-activations, which alone puts a 2-3 % error on the input gradient (measured, DESIGN.md "precision");
// the split keeps the forward at ~1e-5 and the gradient inside the 1e-2 tolerance of BASELINE.json.
//
// The same kernel serves
//   forward      relu(conv(x, W) + b)                          (vgg19.py:171-179)
//   dgrad        (conv(dy, rot180(W)^T) [+ addend]) * [a > 0]   (Conv2DBackpropInput + ReluGrad of tf.gradients,
//                                                               gatys/synthesizer.py:79)
//   Gram dF      F * Dscaled as a 1x1 convolution with per-image weights (utils.py:54), accumulated in place.
#include "conv_common.cuh"
#include "ops.h"

namespace stx {

namespace {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;                       // bf16 elements per 128-byte swizzle row
constexpr int kABytes = kTileM * kChunkK * 2;     // 16 KiB per stage
constexpr int kThreads = 192;

template <int BN, bool SPLIT>
struct SmemLayout {
  static constexpr int kStages = SPLIT ? (BN == 64 ? 4 : 3) : 4;
  static constexpr int kBBytes = BN * kChunkK * 2;
  static constexpr int kOperandBytes = kABytes + kBBytes;              // one (A, W) pair
  static constexpr int kStageBytes = (SPLIT ? 2 : 1) * kOperandBytes;  // SPLIT: [A_hi | W_hi | A_lo | W_lo]
  static constexpr int kBarrierOffset = kStages * kStageBytes;
  static constexpr int kTotal = kBarrierOffset + 128 /*barriers + tmem slot*/ + 1024 /*alignment slack*/;
};

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(kThreads, SPLIT ? 1 : 2)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                  const ConvKernelParams p) {
  using L = SmemLayout<BN, SPLIT>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* accum_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Tile coordinates.
  const int tile = blockIdx.x;
  const int tw = tile % p.tiles_w;
  const int th = (tile / p.tiles_w) % p.tiles_h;
  const int tn = tile / (p.tiles_w * p.tiles_h);
  const int w0 = tw << p.bw_log2;
  const int h0 = th << p.bh_log2;
  const int n0 = tn * p.n_step;
  const int nblk = blockIdx.y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (SPLIT) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
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
  pdl_launch_dependents();   // programmatic dependent launch: see launch_pdl (common.h)
  pdl_wait();

  const int iters = p.taps * p.cchunks;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // Whole warp in the loop, one elected lane issuing: keeps every operand provably warp-uniform (no
    // ELECT / BRA.U.ANY serialisation loops around the uniform-datapath instructions).
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1, 100 + s);
      const int tap = it / p.cchunks;
      const int cc = it - tap * p.cchunks;
      int dy = 0, dx = 0;
      if (p.taps == 9) {
        dy = tap / 3 - p.pad;
        dx = tap % 3 - p.pad;
      }
      uint8_t* a_dst = smem + s * L::kStageBytes;
      uint8_t* b_dst = a_dst + kABytes;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[s], L::kStageBytes);
        tma_load_4d(a_dst, &tmA, &full_bar[s], cc * kChunkK, w0 + dx, h0 + dy, n0);
        tma_load_3d(b_dst, &tmB, &full_bar[s], cc * kChunkK, nblk * BN, p.per_image_w ? n0 : tap);
        if (SPLIT) {
          tma_load_4d(a_dst + L::kOperandBytes, &tmA2, &full_bar[s], cc * kChunkK, w0 + dx, h0 + dy, n0);
          tma_load_3d(b_dst + L::kOperandBytes, &tmB2, &full_bar[s], cc * kChunkK, nblk * BN, tap);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    {
      constexpr uint32_t idesc = make_idesc_bf16(kTileM, BN, 0, 0);
      for (int it = 0; it < iters; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full_bar[s], ph, 200 + s);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
        const uint32_t b_addr = a_addr + kABytes;
        const uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(b_addr, 16, 1024);
        if (elect_one()) {
          if (SPLIT) {
            const uint64_t adesc_lo = make_smem_desc_sw128(a_addr + L::kOperandBytes, 16, 1024);
            const uint64_t bdesc_lo = make_smem_desc_sw128(b_addr + L::kOperandBytes, 16, 1024);
#pragma unroll
            for (int k = 0; k < kChunkK / 16; ++k) {
              umma_bf16(tmem_base, adesc_lo + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);   // A_lo * W_hi
              umma_bf16(tmem_base, adesc + 2 * k, bdesc_lo + 2 * k, idesc, 1);               // A_hi * W_lo
              umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, 1);                  // A_hi * W_hi
            }
          } else {
#pragma unroll
            for (int k = 0; k < kChunkK / 16; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
              umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
            }
          }
          umma_commit(&empty_bar[s]);   // frees the smem stage once these MMAs have read it
          if (it == iters - 1) umma_commit(accum_bar);   // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;      // accumulator row = pixel index inside the tile
    const int ww = r & ((1 << p.bw_log2) - 1);
    const int hh = (r >> p.bw_log2) & ((1 << p.bh_log2) - 1);
    const int nn = r >> (p.bw_log2 + p.bh_log2);
    const int n = n0 + nn, h = h0 + hh, w = w0 + ww;
    const bool valid = (n < p.nimg) && (h < p.H) && (w < p.W) && (!p.per_image_w || nn == 0);
    mbar_wait(accum_bar, 0, 300);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      conv_epilogue_store<SPLIT>(p, v, valid, n, h, w, nblk * BN + c0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, BN);
}

template <int BN, bool SPLIT>
int launch_bn(const ConvOp& op, cudaStream_t stream) {
  using L = SmemLayout<BN, SPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    STX_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  L::kTotal));
    attr_set = true;
  }
  ConvKernelParams p;
  p.H = op.H;
  p.W = op.W;
  p.nimg = op.nimg;
  p.cout = op.cout;
  p.cchunks = op.cin / kChunkK;
  p.taps = op.taps;
  p.bw_log2 = op.bw_log2;
  p.bh_log2 = op.bh_log2;
  p.tiles_w = op.tiles_w;
  p.tiles_h = op.tiles_h;
  p.n_step = op.n_step;
  p.per_image_w = op.per_image_w;
  p.relu = op.relu;
  p.pad = op.pad;
  p.bias = op.bias;
  p.addend = op.addend;
  p.mask = op.mask;
  p.out = op.out;
  p.out_lo = op.out_lo;
  p.rgb_out = nullptr;
  p.rgb_image = nullptr;
  p.rgb_preimage = 0;
  p.pool_out = nullptr;      // the per-tap kernel's row mapping does not match the fused pooling
  p.pool_out_lo = nullptr;
  p.up_out = op.up_out;
  p.up_mask = op.up_mask;
  p.bits_out = nullptr;
  p.up_bits = nullptr;
  p.active = nullptr;
  p.export_out = op.export_out;
  p.corr_scale = 1.0f;
  p.aux_out = nullptr;
  p.pool_aux = nullptr;
  STX_REQUIRE(epilogue_pointers_aligned(p), "conv: epilogue tensors must be 32-byte aligned (256-bit accesses)");
  dim3 grid(op.tiles_w * op.tiles_h * op.tiles_n, op.cout / BN, 1);
  STX_TRY(launch_pdl(conv_igemm_kernel<BN, SPLIT>, grid, dim3(kThreads), L::kTotal, stream, op.tmA, op.tmB,
                     SPLIT ? op.tmA2 : op.tmA, SPLIT ? op.tmB2 : op.tmB, p));
  count_launch();
  return STX_OK;
}

}  // namespace

int conv_prepare_split(ConvOp* op, const bf16* x_lo, const bf16* w_lo) {
  // Second (low-half) operand planes for the bf16x3 forward; geometry must already be prepared.
  STX_REQUIRE(x_lo && w_lo, "conv split: null low-half operand");
  STX_REQUIRE(!op->per_image_w && op->taps == 9, "conv split: only the 3x3 forward uses split operands");
  const int bw = 1 << op->bw_log2, bh = 1 << op->bh_log2, bni = kTileM / (bw * bh);
  const uint64_t adims[4] = {static_cast<uint64_t>(op->cin), static_cast<uint64_t>(op->W),
                             static_cast<uint64_t>(op->H), static_cast<uint64_t>(op->nimg)};
  const uint32_t abox[4] = {64u, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bni)};
  STX_TRY(encode_tmap_bf16(&op->tmA2, x_lo, 4, adims, abox));
  const uint64_t bdims[3] = {static_cast<uint64_t>(op->cin), static_cast<uint64_t>(op->cout),
                             static_cast<uint64_t>(op->taps)};
  const uint32_t bbox[3] = {64u, static_cast<uint32_t>(op->bn), 1u};
  STX_TRY(encode_tmap_bf16(&op->tmB2, w_lo, 3, bdims, bbox));
  op->split = 1;
  return STX_OK;
}

int conv_prepare(ConvOp* op, const bf16* x, int nimg, int H, int W, int cin, const bf16* wpacked, int cout,
                 int taps, int per_image_w, int force_bn) {
  STX_REQUIRE(x && wpacked, "conv: null operand");
  STX_REQUIRE(nimg > 0 && H > 0 && W > 0, "conv: empty tensor");
  STX_REQUIRE(cin % 64 == 0 && cout % 64 == 0, "conv: channels must be multiples of 64");
  STX_REQUIRE(taps == 9 || taps == 1, "conv: taps must be 9 (3x3) or 1 (1x1)");
  op->nimg = nimg;
  op->H = H;
  op->W = W;
  op->cin = cin;
  op->cout = cout;
  op->taps = taps;
  op->per_image_w = per_image_w;
  op->patch = 0;
  op->pair = 0;
  op->res = 0;
  op->mt = 1;
  op->occ = 1;
  op->gram_chunks = 0;
  op->w_tiled = nullptr;
  op->w_tiled_lo = nullptr;
  op->ksplit = 1;
  op->splitk_ws = nullptr;
  op->splitk_counters = nullptr;
  // Tile geometry: BW x BH x BNI pixels = 128 accumulator rows.
  int bw = 16;
  while (bw > 1 && (bw >> 1) >= W) bw >>= 1;       // smallest power of two >= W, capped at 16
  int bh = kTileM / bw;
  while (bh > 1 && (bh >> 1) >= H) bh >>= 1;
  int bni = kTileM / (bw * bh);
  op->bw_log2 = ilog2(bw);
  op->bh_log2 = ilog2(bh);
  op->tiles_w = (W + bw - 1) / bw;
  op->tiles_h = (H + bh - 1) / bh;
  op->n_step = per_image_w ? 1 : bni;
  op->tiles_n = (nimg + op->n_step - 1) / op->n_step;
  // N tile: 128 when that still fills the machine, otherwise 64 (more CTAs for the small deep layers).
  const int tiles_m = op->tiles_w * op->tiles_h * op->tiles_n;
  int bn = (cout % 128 == 0) ? 128 : 64;
  (void)tiles_m;
  if (force_bn == 64 || (force_bn == 128 && cout % 128 == 0)) bn = force_bn;
  op->bn = bn;

  const uint64_t adims[4] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                             static_cast<uint64_t>(nimg)};
  const uint32_t abox[4] = {64u, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bni)};
  STX_TRY(encode_tmap_bf16(&op->tmA, x, 4, adims, abox));
  const uint64_t bdims[3] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(cout),
                             static_cast<uint64_t>(per_image_w ? nimg : taps)};
  const uint32_t bbox[3] = {64u, static_cast<uint32_t>(bn), 1u};
  STX_TRY(encode_tmap_bf16(&op->tmB, wpacked, 3, bdims, bbox));
  op->bias = nullptr;
  op->addend = nullptr;
  op->mask = nullptr;
  op->out = nullptr;
  op->out_lo = nullptr;
  op->rgb_out = nullptr;
  op->rgb_image = nullptr;
  op->rgb_preimage = 0;
  op->pool_out = nullptr;
  op->pool_out_lo = nullptr;
  op->up_out = nullptr;
  op->up_mask = nullptr;
  op->bits_out = nullptr;
  op->up_bits = nullptr;
  op->active = nullptr;
  op->export_out = nullptr;
  op->corr_scale = 1.0f;
  op->aux_out = nullptr;
  op->pool_aux = nullptr;
  op->split = 0;
  op->relu = 0;
  op->pad = 1;
  return STX_OK;
}

int conv3x3_prepare(ConvOp* op, const bf16* x, const bf16* x_lo, int nimg, int H, int W, int cin, const bf16* w,
                    const bf16* w_lo, int cout, int variant) {
  const int want_pair = variant / 1000000;
  variant %= 1000000;
  const int force_igemm = variant / 100000;
  const int force_mt = (variant / 1000) % 10;
  const int force_occ = (variant / 10000) % 10;
  const int force_bn = variant % 1000;
  const int split = (x_lo != nullptr) ? 1 : 0;
  if (!force_igemm) {
    const int rc = conv_patch_prepare(op, x, nimg, H, W, cin, w, cout, split, force_bn, force_mt, force_occ, want_pair);
    if (rc == STX_OK) return split ? conv_patch_prepare_split(op, x_lo, w_lo) : STX_OK;
    if (rc != STX_ERR_UNSUPPORTED) return rc;
  }
  STX_TRY(conv_prepare(op, x, nimg, H, W, cin, w, cout, 9, 0, force_bn == 256 ? 128 : force_bn));
  return split ? conv_prepare_split(op, x_lo, w_lo) : STX_OK;
}

int conv_set_input_pad(ConvOp* op, const bf16* x, int pad) {
  STX_REQUIRE(pad >= 0 && pad <= 2, "conv: pad must be 0 (VALID), 1 (SAME) or 2 (FULL)");
  STX_REQUIRE(op->taps == 9 && !op->split && !op->per_image_w, "conv: input padding modes exist for the plain 3x3 op only");
  const int Hin = op->H + 2 - 2 * pad, Win = op->W + 2 - 2 * pad;
  STX_REQUIRE(Hin > 0 && Win > 0, "conv: input smaller than the filter");
  const uint64_t adims[4] = {static_cast<uint64_t>(op->cin), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                             static_cast<uint64_t>(op->nimg)};
  uint32_t abox[4];
  if (op->patch) {
    abox[0] = 64u; abox[1] = 10u; abox[2] = static_cast<uint32_t>(16 * op->mt + 2); abox[3] = 1u;
  } else {
    const int bw = 1 << op->bw_log2, bh = 1 << op->bh_log2;
    abox[0] = 64u; abox[1] = static_cast<uint32_t>(bw); abox[2] = static_cast<uint32_t>(bh);
    abox[3] = static_cast<uint32_t>(kTileM / (bw * bh));
  }
  STX_TRY(encode_tmap_bf16(&op->tmA, x, 4, adims, abox));
  op->pad = pad;
  return STX_OK;
}

int conv_launch(const ConvOp& op, cudaStream_t stream) {
  if (op.patch) return conv_patch_launch(op, stream);
  STX_REQUIRE(op.out != nullptr || op.up_out != nullptr, "conv: output not set");
  STX_REQUIRE(op.pool_out == nullptr, "conv: fused pooling needs the halo-patch kernel");
  if (op.split) {
    STX_REQUIRE(op.out_lo != nullptr, "conv split: low-half output not set");
    if (op.bn == 128) return launch_bn<128, true>(op, stream);
    return launch_bn<64, true>(op, stream);
  }
  if (op.bn == 128) return launch_bn<128, false>(op, stream);
  return launch_bn<64, false>(op, stream);
}

}  // namespace stx
