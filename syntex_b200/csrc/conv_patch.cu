// 3x3 SAME convolution, halo-patch variant of the tcgen05 implicit GEMM (the throughput path; conv_igemm.cu is
// the general one).
//
// conv_igemm.cu fetches one 128-pixel x 64-channel A tile per filter tap, i.e. it pulls every activation through
// L2 -> SMEM nine times; with 128x128 tiles that is 64 FLOP per L2 byte and the measured kernels sit on the L2
// bandwidth roof (~10 TB/s), not on the tensor pipe (profiles/r01_launches_fg_eval_b8_256.md). Here each
// 64-channel chunk of the (16*MT + 2) x 10 pixel halo patch around a (16*MT) x 8 output tile is loaded ONCE (one
// 4-D TMA box, zero-filled outside the image = SAME padding) and all nine taps read it in place: tcgen05
// shared-memory descriptors address any 128-byte row of a 128B-swizzled buffer (the swizzle is a function of the
// absolute shared-memory address - verified on the B200 by csrc/probe.cu, profiles/r01_descriptor_probe_*), so
// tap (r, s) is just "start address + (r*10 + s) * 128 B, stride between 8-row groups = 10 * 128 B".
// MT = 2 stacks two 16x8 output tiles (two TMEM accumulators) on one weight stream.
//
// The kernel is PERSISTENT: one CTA per SM (or two; or one CTA PAIR per TPC issuing cta_group::2 MMAs, see PatchSmem)
// walks a static round-robin list of (tile, N-block) work items (minus the images switched off, see `skipped`), so the
// prologue (barrier init, TMEM allocation, descriptor prefetch) is paid once and the whole shared memory is one deep
// TMA ring. The TMEM accumulator is double-buffered (when 2 * BN * MT <= 512 columns): eight epilogue warps drain
// tile i (tcgen05.ld -> bias / ReLU / mask / pooling -> bf16 hi/lo stores) while the MMA warp already accumulates
// tile i+1. This matters most for the 64-channel layers, whose 2.6 us of MMA work per tile used to sit between
// ~4 us of per-CTA fixed cost (profiles/r01_ncu_full_conv_patch_b8.md).
//
// Rings (all TMA -> mbarrier -> tcgen05.mma -> tcgen05.commit): halo patches (one fill per 64-channel chunk),
// weight tiles (one [BN x 64] tile per tap), accumulators (MMA -> epilogue -> MMA).
#include "conv_common.cuh"
#include "ops.h"

namespace stx {

int g_debug_resident = 0;      // 1 = resident-weight instances for the 64-channel layers (stx_debug_set key 8).
                               // Measured after the issue-loop fixes: no gain over the ring (conv1_2 forward 116 us
                               // resident with one patch stage vs 109 us ring with two; dgrad 66 us either way), so off.

namespace {

constexpr int kEpilogueWarps = 8;
constexpr int kThreads = 64 + 32 * kEpilogueWarps;   // warp 0: TMA, warp 1: MMA (+TMEM alloc), warps 2..9: epilogue
constexpr int kPatchW = 10;                          // 8 output columns + 2 halo
constexpr int kSmemBudget = 224 * 1024;              // of the 227 KB a CTA may use

constexpr int pow2_at_least(int v) {
  int p = 32;
  while (p < v) p *= 2;
  return p;
}

// RES > 0: the weights of the whole layer (9 taps x RES 64-channel chunks, N = BN = C_out) stay RESIDENT in shared
// memory for the life of the persistent CTA instead of streaming through a ring once per work item. For the
// 64-channel layers a work item is only ~1-2 us of MMA time but used to pull 72-144 KB of weights from L2 every time
// (4x the bytes of its activation patch): 148 SMs fetching the same tiles in lock step made those layers
// L2-delivery-bound (conv1_2 forward: 5.2 us per tile against 1.8 us of tensor work).
// F8 (with SPLIT): the f16 + f8 forward - the second operand pair of a stage is the e4m3 correction pair (see
// conv_common.cuh, kF8*) and accumulates into its own TMEM columns through kind::f8f6f4.
// PAIR: two CTAs of a cluster (the two SMs of a TPC) work on two M tiles against ONE N block with tcgen05
// cta_group::2: each CTA stages its own halo patch and HALF of every weight tile, the leader issues M = 256 MMAs, each
// CTA receives and drains its own 128 rows. Halves the weight bytes every SM pulls through L2 and shared memory -
// what paces the deep layers.
template <int BN, int MT, bool SPLIT, int OCC, int RES, bool F8 = false, bool PAIR = false>
struct PatchSmem {
  static constexpr int kPatchRows = (16 * MT + 2) * kPatchW;
  static constexpr int kPatchBytes = kPatchRows * 128;                        // TMA transaction bytes
  static constexpr int kPatchStride = (kPatchBytes + 1023) & ~1023;           // keep every buffer 1024-aligned
  static constexpr int kPatchStageBytes = (SPLIT ? 2 : 1) * kPatchStride;     // [hi | lo]
  static_assert(OCC == 1 || (OCC == 2 && (MT == 1 || BN == 16)), "two CTAs per SM: single-tile or N = 16 instances only");
  static constexpr int kBRows = PAIR ? BN / 2 : BN;                           // weight rows staged by this CTA
  static constexpr int kBBytes = (kBRows * 128 + 1023) & ~1023;
  static_assert(!PAIR || (MT == 1 && RES == 0 && (BN == 128 || BN == 64 || (BN == 256 && !SPLIT)) && (!SPLIT || F8)),
                "CTA pairs: single-tile, ring weights, N tile 64 / 128 (256: single-MMA only), single-MMA or f16+f8");
  static constexpr int kBStageBytes = (SPLIT ? 2 : 1) * kBBytes;              // [hi | lo]
  static constexpr int kBudget = kSmemBudget / OCC - 2048;
  // resident mode: a small ring for the per-image Gram-difference tiles of the fused Gram gradient
  static constexpr int kDStages = (RES > 0 && !SPLIT) ? (OCC == 2 ? 1 : 2) : 0;
  static constexpr int kResRoom = (kBudget - 9 * RES * kBStageBytes - kDStages * kBBytes) / kPatchStageBytes;
  // (two f16+f8 CTAs per SM - the 64-channel forward, whose tiles are epilogue-bound - have room for ONE hi+corr
  // patch stage each: the other CTA covers the refill)
  static constexpr int kPatchStages = RES > 0 ? (kResRoom > 4 ? 4 : kResRoom)
                                              : ((F8 && OCC == 2) ? 1 : ((SPLIT || MT == 2 || OCC == 2) ? 2 : 3));
  static constexpr int kBRoom = (kBudget - kPatchStages * kPatchStageBytes) / kBStageBytes;
  static constexpr int kBStages = RES > 0 ? 9 * RES : (kBRoom > 10 ? 10 : kBRoom);
  static constexpr int kBBarriers = RES > 0 ? 1 : kBStages;
  static constexpr int kBOffset = kPatchStages * kPatchStageBytes;
  static constexpr int kDOffset = kBOffset + kBStages * kBStageBytes;
  static constexpr int kBarrierOffset = kDOffset + kDStages * kBBytes;
  static constexpr int kTotal = kBarrierOffset + 512 + 1024;
  // MERGE (bf16x3 with N = 64): A_hi * [W_hi ; W_lo] is issued as ONE N = 128 MMA into two column ranges that the
  // epilogue adds. An M = 128, K = 16 MMA costs ~64 cycles however small N is (measured: ~70 cycles per N = 64 MMA,
  // ~72 per N = 128 MMA - the A operand streams out of shared memory at 64 B/clk), so two N = 64 MMAs on the same
  // A tile cost twice what one N = 128 MMA does.
  static constexpr bool kMerge = SPLIT && BN == 64 && !F8;
  static constexpr bool kTwoAcc = kMerge || (F8 && SPLIT);                    // two column ranges per tile, summed by the epilogue
  static constexpr int kAccCols = (kTwoAcc ? 2 : 1) * BN * MT;
  // F8 without SPLIT: fp16 operands, ONE MMA per K step, results stored in the f16+f8 plane format (the layers of an
  // f16+f8 plan that go without the correction product, see plan.cu)
  static_assert(!F8 || (MT == 1 && RES == 0), "f16+f8 forward: one tile, ring weights");
  static constexpr int kAccStages = (2 * kAccCols * OCC <= 512) ? 2 : 1;      // TMEM accumulator buffers
  static constexpr int kTmemCols = pow2_at_least(kAccStages * kAccCols);
  static_assert(kPatchStages >= 1, "no room for a patch stage");
  static_assert(kBStages >= 3, "weight ring too shallow");
  static_assert(2 * (kPatchStages + kBBarriers + kDStages + kAccStages) * 8 + 16 <= 512, "barrier block overflow");
  static_assert(kTotal * OCC <= 227 * 1024, "shared memory budget exceeded");
  static_assert(OCC * pow2_at_least(kAccStages * kAccCols) <= 512, "TMEM budget exceeded");
};

struct PatchSched {
  int total_items;   // tiles * N-blocks
  int nblocks;       // cout / BN
  int gram_chunks;   // fused Gram gradient: extra 1x1 K chunks (C / 64) accumulating F * Dscaled, 0 = none
  int ksplit;        // split-K factor (small grids): work item = (tile, N-block, K slice)
  float* ws;         // split-K partial tiles [tile*nblocks][ksplit][MT*128][BN] fp32
  int* counters;     // split-K arrival counters [tile*nblocks], zero between launches
  const bf16* wt;    // tiled pre-swizzled weights [tap][K/64][N/64][64][64] (null: fetch through tmB / tmB2)
  const bf16* wt_lo;
  int nblocks64;     // N / 64
};

template <int BN, int MT, bool SPLIT, int OCC, int RES, bool F8, bool PAIR>
__global__ void __launch_bounds__(kThreads, OCC)
conv_patch_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                  const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmD,
                  const ConvKernelParams p, const PatchSched sched) {
  using L = PatchSmem<BN, MT, SPLIT, OCC, RES, F8, PAIR>;
  constexpr int kPS = L::kPatchStages, kBS = L::kBStages, kAS = L::kAccStages;
  constexpr int kBB = L::kBBarriers, kDS = L::kDStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* pfull = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);
  uint64_t* pempty = pfull + kPS;
  uint64_t* bfull = pempty + kPS;
  uint64_t* bempty = bfull + kBB;
  uint64_t* dfull = bempty + kBB;      // resident mode: ring of Gram-difference tiles
  uint64_t* dempty = dfull + kDS;
  uint64_t* afull = dempty + kDS;
  uint64_t* aempty = afull + kAS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty + kAS);
  volatile int* splitk_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA pair: rank in the cluster (0 = leader: issues the MMAs, owns the barriers the TMA loads complete on)
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  // the item list is walked per cluster in pair mode (both CTAs take the same item, two M tiles of it)
  const int item_first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int item_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  // Work items of finished jobs (ConvKernelParams::active) are skipped by all three roles alike; a pair item spans two
  // tiles (possibly of two images) and is skipped when neither is wanted - both CTAs of the pair evaluate the same test.
  auto skipped = [&](int item) -> bool {
    if (p.active == nullptr) return false;
    const int per_img = p.tiles_w * p.tiles_h;
    const int t = (item / sched.ksplit) / sched.nblocks;
    if (PAIR) return !(__ldg(p.active + (2 * t) / per_img) | __ldg(p.active + (2 * t + 1) / per_img));
    return !__ldg(p.active + t / per_img);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (SPLIT) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    for (int s = 0; s < kPS; ++s) {
      mbar_init(&pfull[s], 1);
      mbar_init(&pempty[s], 1);
    }
    for (int s = 0; s < kBB; ++s) {
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], 1);
    }
    for (int s = 0; s < kDS; ++s) {
      mbar_init(&dfull[s], 1);
      mbar_init(&dempty[s], 1);
    }
    for (int s = 0; s < kAS; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], (PAIR ? 2 : 1) * kEpilogueWarps);     // pair: both CTAs' epilogue warps free the leader's
    }
    fence_barrier_init();
  }
  if (PAIR) cluster_sync_all();                  // both CTAs' barriers exist before anything remote touches them
  if (warp == 1) {
    if (PAIR) tmem_alloc_2cta(tmem_slot, L::kTmemCols);
    else tmem_alloc(tmem_slot, L::kTmemCols);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above overlapped the tail of the previous kernel in the stream; from here on we touch its results.
  pdl_launch_dependents();
  pdl_wait();

  const int chunks = p.cchunks;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // Whole warp in the loops, one elected lane issuing (see the MMA warp for why).
    {
      // Ring positions are carried as (stage, parity) pairs and advanced incrementally: a modulo by a non-power-of-two
      // stage count per tap costs a dozen dependent instructions, and the single issuing warp - not the tensor pipe -
      // paces the single-MMA kernels (measured: 83 SASS instructions and ~800 cycles per tap for 256 cycles of MMA).
      uint32_t ps = 0, pph = 0, bs = 0, bph = 0;
      constexpr uint32_t kRing = RES > 0 ? (kDS > 0 ? kDS : 1) : kBS;     // ring used by the weight / Gram tiles
      if (RES > 0) {
        // the whole layer's weights, once: tile (chunk c, tap) -> slot c * 9 + tap
        if (elect_one()) {
          mbar_arrive_expect_tx(&bfull[0], kBS * (SPLIT ? 2 : 1) * BN * 128);
          for (int t = 0; t < kBS; ++t) {
            uint8_t* bdst = smem + L::kBOffset + t * L::kBStageBytes;
            tma_load_3d(bdst, &tmB, &bfull[0], (t / 9) * 64, 0, t % 9);
            if (SPLIT) tma_load_3d(bdst + L::kBBytes, &tmB2, &bfull[0], (t / 9) * 64, 0, t % 9);
          }
        }
        __syncwarp();
      }
      for (int item = item_first; item < sched.total_items; item += item_step) {
        if (skipped(item)) continue;
        const int ks = item % sched.ksplit;
        const int rest = item / sched.ksplit;
        const int nblk = rest % sched.nblocks;
        const int tm = PAIR ? 2 * (rest / sched.nblocks) + static_cast<int>(rank) : rest / sched.nblocks;
        const int w0 = (tm % p.tiles_w) * 8;
        const int h0 = ((tm / p.tiles_w) % p.tiles_h) * (16 * MT);
        const int n = tm / tiles_per_img;
        const int c_begin = ks * chunks / sched.ksplit, c_end = (ks + 1) * chunks / sched.ksplit;
        const int gram_chunks = (ks == sched.ksplit - 1) ? sched.gram_chunks : 0;
        const int brow0 = nblk * BN + (PAIR ? static_cast<int>(rank) * (BN / 2) : 0);   // this CTA's weight rows
        for (int c = c_begin; c < c_end; ++c) {
          mbar_wait(&pempty[ps], pph ^ 1, 100 + ps);
          uint8_t* pdst = smem + ps * L::kPatchStageBytes;
          if (elect_one()) {
            if (PAIR) {
              // both CTAs' loads complete on the LEADER's barrier: it expects the bytes of both
              if (leader) mbar_arrive_expect_tx(&pfull[ps], 2 * (SPLIT ? 2 : 1) * L::kPatchBytes);
              tma_load_4d_pair(pdst, &tmA, &pfull[ps], c * 64, w0 - p.pad, h0 - p.pad, n);
              if (SPLIT) tma_load_4d_pair(pdst + L::kPatchStride, &tmA2, &pfull[ps], c * 64, w0 - p.pad, h0 - p.pad, n);
            } else {
              mbar_arrive_expect_tx(&pfull[ps], (SPLIT ? 2 : 1) * L::kPatchBytes);
              tma_load_4d(pdst, &tmA, &pfull[ps], c * 64, w0 - p.pad, h0 - p.pad, n);
              if (SPLIT) tma_load_4d(pdst + L::kPatchStride, &tmA2, &pfull[ps], c * 64, w0 - p.pad, h0 - p.pad, n);
            }
          }
          if (++ps == kPS) {
            ps = 0;
            pph ^= 1;
          }
#pragma unroll
          for (int tap = 0; RES == 0 && tap < 9; ++tap) {
            mbar_wait(&bempty[bs], bph ^ 1, 120 + bs);
            uint8_t* bdst = smem + L::kBOffset + bs * L::kBStageBytes;
            if (PAIR) {
              if (elect_one()) {
                if (leader) mbar_arrive_expect_tx(&bfull[bs], (SPLIT ? 2 : 1) * BN * 128);     // two halves
                tma_load_3d_pair(bdst, &tmB, &bfull[bs], c * 64, brow0, tap);
                if (SPLIT) tma_load_3d_pair(bdst + L::kBBytes, &tmB2, &bfull[bs], c * 64, brow0, tap);
              }
            } else if (elect_one()) {
              mbar_arrive_expect_tx(&bfull[bs], (SPLIT ? 2 : 1) * BN * 128);
              if (sched.wt != nullptr && BN >= 64) {
                // one contiguous, pre-swizzled BN*128-byte run per weight tile
                const size_t woff =
                    ((static_cast<size_t>(tap) * chunks + c) * sched.nblocks64 + nblk * (BN / 64)) * 4096;
                bulk_load_1d(bdst, sched.wt + woff, BN * 128, &bfull[bs]);
                if (SPLIT) bulk_load_1d(bdst + L::kBBytes, sched.wt_lo + woff, BN * 128, &bfull[bs]);
              } else {
                tma_load_3d(bdst, &tmB, &bfull[bs], c * 64, nblk * BN, tap);
                if (SPLIT) tma_load_3d(bdst + L::kBBytes, &tmB2, &bfull[bs], c * 64, nblk * BN, tap);
              }
            }
            if (++bs == kBS) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
        // Fused Gram gradient (utils.py:54): the features of this tile (dense 128-row K-major tiles, staged in a patch
        // buffer) times the per-image scaled Gram difference, accumulated into the same TMEM tile.
        for (int cf = 0; cf < gram_chunks; ++cf) {
          uint64_t* ring_empty = RES > 0 ? dempty : bempty;
          uint64_t* ring_full = RES > 0 ? dfull : bfull;
          mbar_wait(&pempty[ps], pph ^ 1, 140 + ps);
          mbar_wait(&ring_empty[bs], bph ^ 1, 160 + bs);
          uint8_t* pdst = smem + ps * L::kPatchStageBytes;
          uint8_t* bdst = RES > 0 ? smem + L::kDOffset + bs * L::kBBytes : smem + L::kBOffset + bs * L::kBStageBytes;
          if (PAIR) {
            if (elect_one()) {
              if (leader) {
                mbar_arrive_expect_tx(&pfull[ps], 2 * 16384);
                mbar_arrive_expect_tx(&ring_full[bs], BN * 128);
              }
              tma_load_4d_pair(pdst, &tmF, &pfull[ps], cf * 64, w0, h0, n);
              tma_load_3d_pair(bdst, &tmD, &ring_full[bs], cf * 64, brow0, n);
            }
          } else if (elect_one()) {
            mbar_arrive_expect_tx(&pfull[ps], MT * 16384);
            for (int mt = 0; mt < MT; ++mt)
              tma_load_4d(pdst + mt * 16384, &tmF, &pfull[ps], cf * 64, w0, h0 + 16 * mt, n);
            mbar_arrive_expect_tx(&ring_full[bs], BN * 128);
            tma_load_3d(bdst, &tmD, &ring_full[bs], cf * 64, nblk * BN, n);
          }
          if (++ps == kPS) {
            ps = 0;
            pph ^= 1;
          }
          if (++bs == kRing) {
            bs = 0;
            bph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (pair: the leader's, for both CTAs)
    // The whole warp walks the loops (so every value below is provably warp-uniform and lives in uniform
    // registers); one elected lane issues the tcgen05 instructions. With `if (lane == 0)` around the loops the
    // compiler wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY serialisation loop, which costs more
    // issue time than a 32-cycle N = 64 MMA takes to execute.
    if (!PAIR || leader) {
      // operand formats: the f16+f8 forward multiplies fp16 planes (and e4m3 correction planes), everything else bf16
      constexpr int kM = PAIR ? 256 : 128;     // the pair's MMA spans both CTAs' rows
      constexpr uint32_t idesc = F8 ? make_idesc(kM, BN, kFmtF16, kFmtF16, 0, 0) : make_idesc_bf16(kM, BN, 0, 0);
      constexpr uint32_t idesc_f8 = make_idesc(kM, BN, kFmtE4M3, kFmtE4M3, 0, 0);
      constexpr uint32_t idesc_gram = make_idesc_bf16(kM, BN, 0, 0);
      auto mma16 = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
        if (PAIR) umma_bf16_2cta(d, a, b, id, acc);
        else umma_bf16(d, a, b, id, acc);
      };
      auto mma8 = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
        if (PAIR) umma_f8_2cta(d, a, b, id, acc);
        else umma_f8(d, a, b, id, acc);
      };
      auto commit = [](uint64_t* bar) {
        if (PAIR) umma_commit_2cta(bar, 3);      // arrives on the barrier at this offset in BOTH CTAs
        else umma_commit(bar);
      };
      constexpr uint32_t kGroupStride = kPatchW * 128;     // bytes between consecutive 8-pixel row groups
      constexpr uint32_t kDescHiA = desc_hi_sw128(kGroupStride);   // activations: 8-row groups kGroupStride apart
      constexpr uint32_t kDescHiB = desc_hi_sw128(1024);           // weights: dense 128-byte rows
      uint32_t ps = 0, pph = 0, bs = 0, bph = 0, as = 0, aph = 0;     // ring (stage, parity) pairs, see the producer
      constexpr uint32_t kRing = RES > 0 ? (kDS > 0 ? kDS : 1) : kBS;
      const uint32_t b_ring = desc_lo_base(smem_u32(smem + L::kBOffset));
      const uint32_t d_ring = desc_lo_base(smem_u32(smem + L::kDOffset));
      if (RES > 0) {
        mbar_wait(&bfull[0], 0, 219);     // resident weights have landed
        tc_fence_after();
      }
      for (int item = item_first; item < sched.total_items; item += item_step) {
        if (skipped(item)) continue;
        mbar_wait(&aempty[as], aph ^ 1, 300 + as);    // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + as * L::kAccCols;
        const int ks = item % sched.ksplit;
        const int c_begin = ks * chunks / sched.ksplit, c_end = (ks + 1) * chunks / sched.ksplit;
        const int gram_chunks = (ks == sched.ksplit - 1) ? sched.gram_chunks : 0;
        for (int c = c_begin; c < c_end; ++c) {
          mbar_wait(&pfull[ps], pph, 200 + ps);
          // Descriptors differ between MMAs only in their 14-bit start-address field (bits 0..13, address >> 4;
          // shared memory ends below 2^18 bytes, so adding offsets never carries out of the field): keep the upper
          // word constant and do 32-bit adds on the lower one.
          const uint32_t patch = smem_u32(smem + ps * L::kPatchStageBytes);
          const uint32_t a_lo0 = desc_lo_base(patch);
          auto issue_tap = [&](int tap, uint32_t b_lo0, bool first_chunk) {
            const uint32_t tap_off16 = (((tap / 3) * kPatchW + (tap % 3)) * 128) >> 4;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t a_lo = a_lo0 + tap_off16 + ((mt * 16 * kGroupStride) >> 4);
              const uint32_t d = acc + mt * (L::kTwoAcc ? 2 : 1) * BN;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t accumulate = (first_chunk && tap == 0 && k == 0) ? 0u : 1u;
                const uint64_t ad = desc_pack(a_lo + 2 * k, kDescHiA);
                const uint64_t bd = desc_pack(b_lo0 + 2 * k, kDescHiB);
                if (F8 && SPLIT) {
                  // fp16(A) * fp16(W) into the main columns, [A8 | Alo8] * [Wlo8 ; W8] (K = 32 per instruction, the
                  // same 32 bytes of the 128-byte rows) into the correction columns
                  const uint64_t ad_c = desc_pack(a_lo + (L::kPatchStride >> 4) + 2 * k, kDescHiA);
                  const uint64_t bd_c = desc_pack(b_lo0 + (L::kBBytes >> 4) + 2 * k, kDescHiB);
                  mma16(d, ad, bd, idesc, accumulate);
                  mma8(d + BN, ad_c, bd_c, idesc_f8, accumulate);
                } else if (L::kMerge) {
                  const uint64_t ad_lo = desc_pack(a_lo + (L::kPatchStride >> 4) + 2 * k, kDescHiA);
                  // [W_hi ; W_lo] are adjacent 64-row tiles of one stage: a single N = 128 operand
                  umma_bf16(d, ad, bd, make_idesc_bf16(128, 2 * BN, 0, 0), accumulate);   // A_hi * [W_hi ; W_lo]
                  umma_bf16(d, ad_lo, bd, idesc, 1);                                      // A_lo * W_hi
                } else if (SPLIT) {
                  const uint64_t ad_lo = desc_pack(a_lo + (L::kPatchStride >> 4) + 2 * k, kDescHiA);
                  const uint64_t bd_lo = desc_pack(b_lo0 + (L::kBBytes >> 4) + 2 * k, kDescHiB);
                  umma_bf16(d, ad_lo, bd, idesc, accumulate);   // A_lo * W_hi
                  umma_bf16(d, ad, bd_lo, idesc, 1);            // A_hi * W_lo
                  umma_bf16(d, ad, bd, idesc, 1);               // A_hi * W_hi
                } else {
                  mma16(d, ad, bd, idesc, accumulate);
                }
              }
            }
          };
          if (RES > 0) {
            // resident weights: no barrier inside the chunk, nine taps issued as one straight-line block
            tc_fence_after();
            const uint32_t b_base = b_ring + c * 9 * (L::kBStageBytes >> 4);
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) issue_tap(tap, b_base + tap * (L::kBStageBytes >> 4), c == c_begin);
              umma_commit(&pempty[ps]);
            }
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&bfull[bs], bph, 220 + bs);
              tc_fence_after();
              if (elect_one()) {
                issue_tap(tap, b_ring + bs * (L::kBStageBytes >> 4), c == c_begin);
                commit(&bempty[bs]);
                if (tap == 8) commit(&pempty[ps]);
              }
              if (++bs == kBS) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
          if (++ps == kPS) {
            ps = 0;
            pph ^= 1;
          }
        }
        for (int cf = 0; cf < gram_chunks; ++cf) {
          uint64_t* ring_empty = RES > 0 ? dempty : bempty;
          uint64_t* ring_full = RES > 0 ? dfull : bfull;
          mbar_wait(&pfull[ps], pph, 240 + ps);
          mbar_wait(&ring_full[bs], bph, 260 + bs);
          tc_fence_after();
          const uint32_t f_lo = desc_lo_base(smem_u32(smem + ps * L::kPatchStageBytes));
          const uint32_t g_lo = RES > 0 ? d_ring + bs * (L::kBBytes >> 4) : b_ring + bs * (L::kBStageBytes >> 4);
          if (elect_one()) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma16(acc + mt * BN, desc_pack(f_lo + mt * (16384 >> 4) + 2 * k, kDescHiB),
                      desc_pack(g_lo + 2 * k, kDescHiB), idesc_gram, 1);
            }
            commit(&ring_empty[bs]);
            commit(&pempty[ps]);
          }
          if (++ps == kPS) {
            ps = 0;
            pph ^= 1;
          }
          if (++bs == kRing) {
            bs = 0;
            bph ^= 1;
          }
        }
        if (elect_one()) commit(&afull[as]);
        if (++as == kAS) {
          as = 0;
          aph ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;         // which half of the 32-column chunks this warp drains
    const int r = q * 32 + lane;
    const int ww = r & 7;
    int ait = 0;
    auto free_acc = [&](int s) {           // hand an accumulator back to the MMA issuer (the leader's, in a pair)
      if (PAIR) mbar_arrive_leader(&aempty[s]);
      else mbar_arrive(&aempty[s]);
    };
    for (int item = item_first; item < sched.total_items; item += item_step, ++ait) {
      if (skipped(item)) {
        --ait;            // undone by the loop increment: the accumulator ring follows processed items
        continue;
      }
      const int ks = item % sched.ksplit;
      const int rest = item / sched.ksplit;
      const int nblk = rest % sched.nblocks;
      const int tm = PAIR ? 2 * (rest / sched.nblocks) + static_cast<int>(rank) : rest / sched.nblocks;
      const int w = (tm % p.tiles_w) * 8 + ww;
      const int h0 = ((tm / p.tiles_w) % p.tiles_h) * (16 * MT);
      const int n = tm / tiles_per_img;
      const int as = ait % kAS;
      const uint32_t acc = tmem_base + as * L::kAccCols + (static_cast<uint32_t>(q * 32) << 16);
      if (BN != 16 && (p.mask || p.up_mask || p.addend)) {
        // While the tile is still accumulating: pull the epilogue's global operands (ReLU masks, addend) into L1, so
        // that the loads issued after the TMEM read hit instead of paying a DRAM/L2 round trip per chunk on the
        // critical path of the 64-channel layers (whose tiles are only ~1 us of tensor work).
        constexpr int kCPT = BN / 32, kC = MT * kCPT;
        for (int j = (kC >= 2 ? half : 0); j < kC; j += (kC >= 2 ? 2 : 1)) {
          const int h = h0 + (j / kCPT) * 16 + (r >> 3);
          if (h < p.H && w < p.W) {
            const int ch0 = nblk * BN + (j % kCPT) * 32;
            const size_t off = ((static_cast<size_t>(n) * p.H + h) * p.W + w) * p.cout + ch0;
            if (p.mask) prefetch_l1(p.mask + off);
            if (p.addend) prefetch_l1(p.addend + off);
            if (p.up_mask) {
#pragma unroll
              for (int d = 0; d < 4; ++d)
                prefetch_l1(p.up_mask + ((static_cast<size_t>(n) * 2 * p.H + 2 * h + (d >> 1)) * 2 * p.W + 2 * w + (d & 1)) * p.cout + ch0);
            }
          }
        }
      }
      mbar_wait_relaxed(&afull[as], (ait / kAS) & 1, 400 + as);
      tc_fence_after();
      if (BN == 16) {
        // conv1_1 input gradient: both stacked tiles sit in one 32-column slab (16 columns each, 3 real).
        float v[32];
        if (half == 0) {
          tmem_ld_32x32(acc, v);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) free_acc(as);
        if (half == 0) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int h = h0 + mt * 16 + (r >> 3);
            if (h < p.H && w < p.W) {
              const size_t pix = (static_cast<size_t>(n) * p.H + h) * p.W + w;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float d = v[mt * 16 + c] * 255.0f;
                if (p.rgb_preimage) {
                  const float sg = 1.0f / (1.0f + __expf(-p.rgb_image[pix * 3 + c]));
                  d *= sg * (1.0f - sg);
                }
                p.rgb_out[pix * 3 + c] = d;
              }
            }
          }
        }
      } else {
        constexpr int kChunksPerTile = BN / 32;
        constexpr int kChunks = MT * kChunksPerTile;
        constexpr int kStep = kChunks >= 2 ? 2 : 1;
        const int first = kChunks >= 2 ? half : 0;
        const int last = first + ((kChunks - 1 - first) / kStep) * kStep;   // last chunk this warp handles
        if (sched.ksplit > 1) {
          // Split-K: park the fp32 partial tile in the workspace; the last slice to arrive sums all slices in a
          // fixed order (bit-reproducible whichever CTA is last) and runs the real epilogue.
          float* wsme = sched.ws + (static_cast<size_t>(rest) * sched.ksplit + ks) * (MT * 128 * BN);
#pragma unroll 1
          for (int j = first; j < kChunks; j += kStep) {
            const int mt = j / kChunksPerTile;
            const int c0 = (j % kChunksPerTile) * 32;
            float v[32];
            tmem_ld_32x32(acc + mt * BN + c0, v);
            tmem_ld_wait();
            float4* d4 = reinterpret_cast<float4*>(wsme + (static_cast<size_t>(mt) * 128 + r) * BN + c0);
#pragma unroll
            for (int e = 0; e < 8; ++e) d4[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) free_acc(as);
          __threadfence();
          asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
          if (threadIdx.x == 64) {
            const int old = atomicAdd(&sched.counters[rest], 1);
            const int is_last = (old == sched.ksplit - 1);
            if (is_last) sched.counters[rest] = 0;
            *splitk_flag = is_last;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
          if (*splitk_flag) {
            __threadfence();
            const float* wsall = sched.ws + static_cast<size_t>(rest) * sched.ksplit * (MT * 128 * BN);
#pragma unroll 1
            for (int j = first; j < kChunks; j += kStep) {
              const int mt = j / kChunksPerTile;
              const int c0 = (j % kChunksPerTile) * 32;
              const int h = h0 + mt * 16 + (r >> 3);
              const bool valid = (h < p.H) && (w < p.W);
              // slices summed in a fixed order; two slices (16 independent 16-byte loads) in flight at a time
              float v[32];
              const float4* s4 = reinterpret_cast<const float4*>(wsall + (static_cast<size_t>(mt) * 128 + r) * BN + c0);
              const size_t slice_stride4 = static_cast<size_t>(MT) * 128 * BN / 4;
              {
                float4 t[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = __ldcg(s4 + e);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  v[4 * e] = t[e].x;
                  v[4 * e + 1] = t[e].y;
                  v[4 * e + 2] = t[e].z;
                  v[4 * e + 3] = t[e].w;
                }
              }
              int sl = 1;
              for (; sl + 1 < sched.ksplit; sl += 2) {
                float4 ta[8], tb[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  ta[e] = __ldcg(s4 + sl * slice_stride4 + e);
                  tb[e] = __ldcg(s4 + (sl + 1) * slice_stride4 + e);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  v[4 * e] = (v[4 * e] + ta[e].x) + tb[e].x;
                  v[4 * e + 1] = (v[4 * e + 1] + ta[e].y) + tb[e].y;
                  v[4 * e + 2] = (v[4 * e + 2] + ta[e].z) + tb[e].z;
                  v[4 * e + 3] = (v[4 * e + 3] + ta[e].w) + tb[e].w;
                }
              }
              if (sl < sched.ksplit) {
                float4 ta[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) ta[e] = __ldcg(s4 + sl * slice_stride4 + e);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  v[4 * e] += ta[e].x;
                  v[4 * e + 1] += ta[e].y;
                  v[4 * e + 2] += ta[e].z;
                  v[4 * e + 3] += ta[e].w;
                }
              }
              conv_epilogue_store<SPLIT>(p, v, valid, n, h, w, nblk * BN + c0);
            }
          }
          // the flag is re-written only after every warp has passed the second barrier of the NEXT item
          continue;
        }
        bool released = false;
        if (kChunks < 2 && half == 1) {       // nothing to drain for this warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) free_acc(as);
          released = true;
        }
#pragma unroll 1
        for (int j = first; j < kChunks && !released; j += kStep) {
          const int mt = j / kChunksPerTile;
          const int c0 = (j % kChunksPerTile) * 32;
          const int h = h0 + mt * 16 + (r >> 3);
          const bool valid = (h < p.H) && (w < p.W);
          float v[32];
          if (L::kTwoAcc) {
            float u[32];
            tmem_ld_32x32(acc + mt * 2 * BN + c0, v);
            tmem_ld_32x32(acc + mt * 2 * BN + BN + c0, u);
            tmem_ld_wait();
            const float cs = (F8 && SPLIT) ? p.corr_scale : 1.0f;
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = fmaf(u[e], cs, v[e]);
          } else {
            tmem_ld_32x32(acc + mt * BN + c0, v);
            tmem_ld_wait();
          }
          if (j == last) {                    // accumulator fully read by this warp: hand it back before storing
            tc_fence_before();
            __syncwarp();
            if (lane == 0) free_acc(as);
          }
          conv_epilogue_store<SPLIT, F8 ? 2 : 0>(p, v, valid, n, h, w, nblk * BN + c0);
        }
      }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();          // the peer may still be reading / signalling
  else __syncthreads();
  if (warp == 1) {
    if (PAIR) tmem_dealloc_2cta(tmem_base, L::kTmemCols);
    else tmem_dealloc(tmem_base, L::kTmemCols);
  }
}

template <int BN, int MT, bool SPLIT, int OCC, int RES = 0, bool F8 = false, bool PAIR = false>
int launch_patch(const ConvOp& op, cudaStream_t stream) {
  using L = PatchSmem<BN, MT, SPLIT, OCC, RES, F8, PAIR>;
  static bool attr_set = false;
  if (!attr_set) {
    STX_CUDA(cudaFuncSetAttribute(conv_patch_kernel<BN, MT, SPLIT, OCC, RES, F8, PAIR>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  ConvKernelParams p;
  p.H = op.H;
  p.W = op.W;
  p.nimg = op.nimg;
  p.cout = op.cout;
  p.cchunks = op.cin / 64;
  p.taps = 9;
  p.bw_log2 = 3;
  p.bh_log2 = 4;
  p.tiles_w = op.tiles_w;
  p.tiles_h = op.tiles_h;
  p.n_step = 1;
  p.per_image_w = 0;
  p.relu = op.relu;
  p.pad = op.pad;
  p.bias = op.bias;
  p.addend = op.addend;
  p.mask = op.mask;
  p.out = op.out;
  p.out_lo = op.out_lo;
  p.rgb_out = op.rgb_out;
  p.rgb_image = op.rgb_image;
  p.rgb_preimage = op.rgb_preimage;
  p.pool_out = op.pool_out;
  p.pool_out_lo = op.pool_out_lo;
  p.up_out = op.up_out;
  p.up_mask = op.up_mask;
  p.bits_out = op.bits_out;
  p.up_bits = op.up_bits;
  p.active = op.active;
  p.export_out = op.export_out;
  p.corr_scale = op.corr_scale;
  p.aux_out = op.aux_out;
  p.pool_aux = op.pool_aux;
  STX_REQUIRE(epilogue_pointers_aligned(p), "conv: epilogue tensors must be 32-byte aligned (256-bit accesses)");
  PatchSched sched;
  sched.nblocks = op.cout / BN;
  sched.ksplit = (op.ksplit > 1 && op.splitk_ws && op.splitk_counters && BN != 16 && !L::kTwoAcc) ? op.ksplit : 1;
  sched.wt = op.w_tiled;
  sched.wt_lo = SPLIT ? op.w_tiled_lo : nullptr;
  if (SPLIT && !op.w_tiled_lo) sched.wt = nullptr;
  sched.nblocks64 = op.cout / 64;
  sched.ws = op.splitk_ws;
  sched.counters = op.splitk_counters;
  sched.total_items = op.tiles_w * op.tiles_h * op.nimg * sched.nblocks * sched.ksplit;
  if (PAIR) {
    STX_REQUIRE((op.tiles_w * op.tiles_h * op.nimg) % 2 == 0 && sched.ksplit == 1, "conv: CTA pairs need an even tile count");
    sched.total_items /= 2;            // one item = two M tiles (one per CTA of the pair) against one N block
    sched.wt = nullptr;
  }
  sched.gram_chunks = (!SPLIT && op.gram_chunks > 0) ? op.gram_chunks : 0;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    STX_CUDA(cudaGetDevice(&dev));
    STX_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int slots = PAIR ? (num_sms / 2) * OCC : num_sms * OCC;       // persistent: OCC CTAs per SM / OCC pairs per TPC
  const int grid = (PAIR ? 2 : 1) * (sched.total_items < slots ? sched.total_items : slots);
  if (RES > 0) {
    sched.wt = nullptr;
    STX_REQUIRE(sched.nblocks == 1 && sched.ksplit == 1 && op.cin == 64 * RES,
                "conv: resident-weight instance on a shape it does not cover");
  }
  STX_TRY(launch_pdl_cluster(PAIR ? 2 : 1, conv_patch_kernel<BN, MT, SPLIT, OCC, RES, F8, PAIR>, dim3(grid), dim3(kThreads),
                             L::kTotal, stream, op.tmA, op.tmB, SPLIT ? op.tmA2 : op.tmA, SPLIT ? op.tmB2 : op.tmB, sched.gram_chunks ? op.tmF : op.tmA,
                             sched.gram_chunks ? op.tmD : op.tmB, p, sched));
  count_launch();
  return STX_OK;
}

}  // namespace

// Geometry + tensor maps of the patch variant. Returns STX_ERR_UNSUPPORTED (without touching op) when the shape
// does not qualify (needs W >= 8: an 8-pixel row group must be contiguous in the patch).
int conv_patch_prepare(ConvOp* op, const bf16* x, int nimg, int H, int W, int cin, const bf16* wpacked, int cout,
                       int split, int force_bn, int force_mt, int force_occ, int want_pair) {
  STX_REQUIRE(x && wpacked, "conv: null operand");
  STX_REQUIRE(cin % 64 == 0 && (cout % 64 == 0 || cout == 16), "conv: channels must be multiples of 64");
  if (W < 8) return STX_ERR_UNSUPPORTED;
  op->nimg = nimg;
  op->H = H;
  op->W = W;
  op->cin = cin;
  op->cout = cout;
  op->taps = 9;
  op->per_image_w = 0;
  op->patch = 1;
  // Tile configuration from the measured variant tables (profiles/r01_conv_variants_*_persistent.log):
  //  * N tile 128 whenever the channel count allows it (never slower than 64, even on tiny grids);
  //  * single-MMA kernels (dgrad, fast forward): one 16x8 tile per work item and TWO persistent CTAs per SM - one
  //    issuing thread cannot keep the tensor pipe fed at 4 MMAs per weight tile, two can (1025 vs 666 TFLOP/s on
  //    conv4_2 dgrad);
  //  * bf16x3 forward: one CTA per SM (12 MMAs per weight tile keep a single issuer busy; its hi/lo rings need the
  //    whole shared memory).
  int bn = 64, mt = 1, occ = split ? 1 : 2;
  if (cout == 16) {
    bn = 16;
    mt = 2;
    occ = 1;
  } else {
    if (cout % 128 == 0) bn = 128;
    if (force_bn == 64 || (force_bn == 128 && cout % 128 == 0) || (force_bn == 256 && cout % 256 == 0 && !split))
      bn = force_bn;
    if (force_mt == 1 || (force_mt == 2 && !split)) mt = force_mt;
    if (bn == 256) mt = 2;              // only the <256, 2> instance exists
    // Two CTAs per SM only pay when there are work items for them: with a grid of <= one item per SM (deep layers
    // at batch 1) the one-CTA layout wins because its 9-stage weight ring covers the L2 latency on its own.
    const int items = ((W + 7) / 8) * ((H + 16 * mt - 1) / (16 * mt)) * nimg * (cout / bn);
    if (items <= 148) occ = 1;
    // Short K (<= 2 input chunks) with N = 128: tiles are epilogue-heavy and a deeper weight ring pays; one CTA per SM
    // measured 44 us against 50 us on the conv2_2 input gradient at batch 9
    // (profiles/r01_conv_variants_b9_256_occ.log), while K >= 4 chunks prefers two (35 vs 38 us).
    if (!split && bn == 128 && cin <= 128) occ = 1;
    // Deep single-MMA layers (K >= 4 chunks, N tile 128) with plenty of work: two stacked 16x8 tiles per weight stream
    // (one CTA per SM) halve the weight traffic that paces them; measured on the input-gradient pass at 64 / 16 jobs:
    // 2.97 -> 2.85 ms / 0.79 -> 0.78 ms per evaluation (round 1 at 9 jobs: no difference).
    if (!split && bn == 128 && cin >= 256 && force_mt == 0 && force_occ == 0) {
      const int items2 = ((W + 7) / 8) * ((H + 31) / 32) * nimg * (cout / bn);
      if (items2 >= 4 * 148) mt = 2;
    }
    if (force_occ == 1 || force_occ == 2) occ = force_occ;
    if (split || mt == 2) occ = 1;      // only MT = 1 single-MMA instances exist with two CTAs per SM
  }
  // Resident weights (see PatchSmem): 64-channel outputs, one N block, at most two input chunks, enough work items
  // per CTA to amortise loading the whole layer once.
  int res = 0;
  if (g_debug_resident != 0 && bn == 64 && cout == 64 && mt == 1 && force_bn == 0 && force_mt == 0 && force_occ == 0) {
    const int items = ((W + 7) / 8) * ((H + 15) / 16) * nimg;
    if (items >= 2 * 148) {
      if (cin == 64) res = 1;
      else if (cin == 128 && !split) res = 2;
    }
    if (res == 1) occ = split ? 1 : 2;
    if (res == 2) occ = 1;
  }
  // CTA pairs (cta_group::2, see PatchSmem): asked for by the caller, N tile 64 or 128, an even number of 16x8 tiles
  // and at least one pair item per TPC. Measured at 64 jobs of 256^2 (profiles/r02_cta_pairs.log): f16+f8 forward 4.87
  // -> 4.28 ms, input gradient 2.92 -> 2.64 ms, every layer gains (the deep ones most: half the weight bytes per SM;
  // the 64-channel ones 2-3 %: they are bound by the shared-memory reads of the pixel operand), results bit-identical;
  // at 4 jobs no difference. They replace the two-stacked-tiles (MT = 2) instances wherever they apply.
  int pair = 0;
  if (want_pair && (bn == 128 || bn == 64) && res == 0 && force_mt == 0 && force_occ == 0) {
    const int tiles = ((W + 7) / 8) * ((H + 15) / 16) * nimg;
    if (tiles % 2 == 0 && (tiles / 2) * (cout / bn) >= 74) {
      pair = 1;
      mt = 1;
      if (bn == 128) occ = 1;
      // N tile 256 for the single-MMA kernels (M = 256 x N = 256 MMAs, 128 weight rows per SM): half the operand bytes
      // read from shared memory per MMA cycle again (input gradient at 64 jobs: 2.52 -> 2.43 ms)
      if (bn == 128 && !split && cout % 256 == 0 && (tiles / 2) * (cout / 256) >= 74) bn = 256;
    }
  }
  op->pair = pair;
  const int tiles_w = (W + 7) / 8;
  op->bn = bn;
  op->mt = mt;
  op->occ = occ;
  op->res = res;
  op->tiles_w = tiles_w;
  op->tiles_h = (H + 16 * mt - 1) / (16 * mt);
  op->tiles_n = nimg;
  op->n_step = 1;
  op->bw_log2 = 3;
  op->bh_log2 = 4;
  const uint64_t adims[4] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                             static_cast<uint64_t>(nimg)};
  const uint32_t abox[4] = {64u, static_cast<uint32_t>(kPatchW), static_cast<uint32_t>(16 * mt + 2), 1u};
  STX_TRY(encode_tmap_bf16(&op->tmA, x, 4, adims, abox));
  const uint64_t bdims[3] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(cout), 9u};
  const uint32_t bbox[3] = {64u, static_cast<uint32_t>(pair ? bn / 2 : bn), 1u};    // pair: each CTA loads half the rows
  STX_TRY(encode_tmap_bf16(&op->tmB, wpacked, 3, bdims, bbox));
  op->bias = nullptr;
  op->addend = nullptr;
  op->mask = nullptr;
  op->out = nullptr;
  op->out_lo = nullptr;
  op->rgb_out = nullptr;
  op->rgb_image = nullptr;
  op->rgb_preimage = 0;
  op->gram_chunks = 0;
  op->w_tiled = nullptr;
  op->w_tiled_lo = nullptr;
  op->ksplit = 1;
  op->splitk_ws = nullptr;
  op->splitk_counters = nullptr;
  op->pool_out = nullptr;
  op->pool_out_lo = nullptr;
  op->up_out = nullptr;
  op->up_mask = nullptr;
  op->bits_out = nullptr;
  op->up_bits = nullptr;
  op->active = nullptr;
  op->export_out = nullptr;
  op->split = 0;
  op->relu = 0;
  op->pad = 1;
  op->corr_scale = 1.0f;
  op->aux_out = nullptr;
  op->pool_aux = nullptr;
  return STX_OK;
}

// Split-K for grids that leave most SMs idle (deep layers at batch 1): returns the factor that fills the machine
// (<= number of 64-channel chunks), and the workspace this op then needs.
int conv_patch_plan_splitk(const ConvOp& op, size_t* ws_floats, size_t* counters) {
  *ws_floats = 0;
  *counters = 0;
  if (!op.patch || op.bn == 16 || op.pair) return 1;
  const int items = op.tiles_w * op.tiles_h * op.nimg * (op.cout / op.bn);
  const int slots = 148 * op.occ;
  const int chunks = op.cin / 64;
  int ks = 1;
  if (items * 2 <= slots) ks = slots / items;
  if (ks > 4) ks = 4;                 // the fix-up of the last slice is serial per tile: keep it short
  if (ks > chunks) ks = chunks;
  if (ks < 2) return 1;
  *ws_floats = static_cast<size_t>(items) * ks * op.mt * 128 * op.bn;
  *counters = static_cast<size_t>(items);
  return ks;
}

// Attach the fused Gram gradient to a prepared single-MMA patch op whose output lives on the layer of `feat`:
// out += feat [nimg,H,W,cout] x dscaled [nimg][cout][cout] (per image), accumulated in TMEM before the epilogue.
int conv_patch_attach_gram(ConvOp* op, const bf16* feat, const bf16* dscaled) {
  STX_REQUIRE(op->patch && !op->split && op->bn != 16, "fused Gram gradient: needs a single-MMA halo-patch op");
  STX_REQUIRE(feat && dscaled, "fused Gram gradient: null operand");
  const uint64_t fdims[4] = {static_cast<uint64_t>(op->cout), static_cast<uint64_t>(op->W),
                             static_cast<uint64_t>(op->H), static_cast<uint64_t>(op->nimg)};
  const uint32_t fbox[4] = {64u, 8u, 16u, 1u};
  STX_TRY(encode_tmap_bf16(&op->tmF, feat, 4, fdims, fbox));
  const uint64_t ddims[3] = {static_cast<uint64_t>(op->cout), static_cast<uint64_t>(op->cout),
                             static_cast<uint64_t>(op->nimg)};
  const uint32_t dbox[3] = {64u, static_cast<uint32_t>(op->pair ? op->bn / 2 : op->bn), 1u};
  STX_TRY(encode_tmap_bf16(&op->tmD, dscaled, 3, ddims, dbox));
  op->gram_chunks = op->cout / 64;
  return STX_OK;
}

int conv_patch_prepare_split(ConvOp* op, const bf16* x_lo, const bf16* w_lo) {
  STX_REQUIRE(x_lo && w_lo, "conv split: null low-half operand");
  STX_REQUIRE(op->patch && op->mt == 1, "conv split: patch geometry must be prepared with split = 1");
  const uint64_t adims[4] = {static_cast<uint64_t>(op->cin), static_cast<uint64_t>(op->W),
                             static_cast<uint64_t>(op->H), static_cast<uint64_t>(op->nimg)};
  const uint32_t abox[4] = {64u, static_cast<uint32_t>(kPatchW), 18u, 1u};
  STX_TRY(encode_tmap_bf16(&op->tmA2, x_lo, 4, adims, abox));
  const uint64_t bdims[3] = {static_cast<uint64_t>(op->cin), static_cast<uint64_t>(op->cout), 9u};
  const uint32_t bbox[3] = {64u, static_cast<uint32_t>(op->pair ? op->bn / 2 : op->bn), 1u};
  STX_TRY(encode_tmap_bf16(&op->tmB2, w_lo, 3, bdims, bbox));
  op->split = 1;
  return STX_OK;
}

int conv_patch_launch(const ConvOp& op, cudaStream_t stream) {
  if (op.bn == 16) {
    STX_REQUIRE(op.rgb_out != nullptr && op.mt == 2 && !op.split, "conv1_1 dgrad: bad configuration");
    return launch_patch<16, 2, false, 2>(op, stream);   // two CTAs per SM: the single MMA issuer is the limit
  }
  STX_REQUIRE(op.out != nullptr || op.up_out != nullptr || (op.bits_out != nullptr && op.pool_out != nullptr),
              "conv: output not set");
  STX_REQUIRE((op.bits_out == nullptr && op.up_bits == nullptr) || op.cout % 32 == 0, "conv: bit masks need 32-channel groups");
  STX_REQUIRE(op.pool_out == nullptr || (op.H % 2 == 0 && op.W % 2 == 0), "conv: fused pooling needs even H and W");
  if (op.split == 3) {      // fp16 operands, single MMA, f16+f8-format outputs
    STX_REQUIRE(op.res == 0 && op.mt == 1, "conv f16 (plain): bad configuration");
    if (op.pair && op.bn == 256) return launch_patch<256, 1, false, 1, 0, true, true>(op, stream);
    if (op.pair && op.bn == 128) return launch_patch<128, 1, false, 1, 0, true, true>(op, stream);
    STX_REQUIRE(!op.pair, "conv f16 (plain): no CTA-pair instance for this N tile");
    if (op.bn == 128) return launch_patch<128, 1, false, 1, 0, true>(op, stream);
    STX_REQUIRE(op.bn == 64, "conv f16 (plain): N tile must be 64 or 128 without pairs");
    return launch_patch<64, 1, false, 1, 0, true>(op, stream);
  }
  if (op.split == 2) {
    STX_REQUIRE(op.out_lo != nullptr || op.pool_out_lo != nullptr, "conv f16+f8: correction-plane output not set");
    STX_REQUIRE(op.res == 0 && op.mt == 1, "conv f16+f8: bad configuration");
    if (op.pair && op.bn == 128) return launch_patch<128, 1, true, 1, 0, true, true>(op, stream);
    if (op.pair) return op.occ == 2 ? launch_patch<64, 1, true, 2, 0, true, true>(op, stream)
                                    : launch_patch<64, 1, true, 1, 0, true, true>(op, stream);
    if (op.bn == 128) return launch_patch<128, 1, true, 1, 0, true>(op, stream);
    if (op.occ == 2) return launch_patch<64, 1, true, 2, 0, true>(op, stream);
    return launch_patch<64, 1, true, 1, 0, true>(op, stream);
  }
  if (op.split) {
    STX_REQUIRE(!op.pair, "conv split: no CTA-pair instance of the bf16x3 forward");
    STX_REQUIRE(op.out_lo != nullptr || op.pool_out_lo != nullptr, "conv split: low-half output not set");
    if (op.bn == 128) return launch_patch<128, 1, true, 1>(op, stream);
    if (op.res == 1) return launch_patch<64, 1, true, 1, 1>(op, stream);
    return launch_patch<64, 1, true, 1>(op, stream);
  }
  if (op.res == 1) return launch_patch<64, 1, false, 2, 1>(op, stream);
  if (op.res == 2) return launch_patch<64, 1, false, 1, 2>(op, stream);
  const bool two = op.occ == 2;
  if (op.bn == 256 && op.pair) return launch_patch<256, 1, false, 1, 0, false, true>(op, stream);
  if (op.bn == 256) return launch_patch<256, 2, false, 1>(op, stream);
  if (op.bn == 128) {
    if (op.pair) return launch_patch<128, 1, false, 1, 0, false, true>(op, stream);
    if (op.mt == 2) return launch_patch<128, 2, false, 1>(op, stream);
    return two ? launch_patch<128, 1, false, 2>(op, stream) : launch_patch<128, 1, false, 1>(op, stream);
  }
  if (op.mt == 2) return launch_patch<64, 2, false, 1>(op, stream);
  if (op.pair) return two ? launch_patch<64, 1, false, 2, 0, false, true>(op, stream)
                          : launch_patch<64, 1, false, 1, 0, false, true>(op, stream);
  return two ? launch_patch<64, 1, false, 2>(op, stream) : launch_patch<64, 1, false, 1>(op, stream);
}

}  // namespace stx
