// Bring-up probe (debug entry point, not on the product path): checks how tcgen05 shared-memory descriptors
// address an A operand whose 8-row groups start at an arbitrary 128-byte row of a larger TMA-written,
// 128B-swizzled patch (start address not 1024-byte aligned, stride between groups not a multiple of 1024).
// This is the addressing the halo-patch convolution needs (one patch load serves all nine filter taps).
#include "common.h"
#include "ptx.cuh"

namespace stx {
namespace {

__global__ void __launch_bounds__(128, 1)
patch_probe_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmB, int rows,
                   int start_row, int sbo_bytes, int base_mode, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem;                       // rows * 128 B
  uint8_t* sB = smem + 32768;               // 64 * 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, rows * 128 + 64 * 128);
    tma_load_2d(sP, &tmP, bar, 0, 0);
    tma_load_2d(sB, &tmB, bar, 0, 0);
    mbar_wait(bar, 0, 1);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sP) + start_row * 128;
    const uint32_t b_addr = smem_u32(sB);
    uint64_t adesc = make_smem_desc_sw128(a_addr, 16, sbo_bytes);
    if (base_mode == 1) adesc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    const uint64_t bdesc = make_smem_desc_sw128(b_addr, 16, 1024);
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
    umma_commit(done);
  }
  __syncwarp();
  mbar_wait(done, 0, 2);
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    float v[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 64);
}

// Bring-up probe of the CTA-pair MMA (cta_group::2): D[256 x 128] = A[256 x 64] * B[128 x 64]^T by ONE cluster of two
// CTAs. Each CTA loads its 128 rows of A and its 64 rows of B (TMA completing on the leader's barrier), the leader
// issues four M = 256 MMAs, the commit is multicast to both CTAs, each CTA reads its 128 rows of D from its own TMEM.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // 128 rows * 128 B
  uint8_t* sB = smem + 16384;               // 64 rows * 128 B
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
  uint64_t* done = full + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 1) tmem_alloc_2cta(slot, 128);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    if (rank == 0) mbar_arrive_expect_tx(full, 2 * (16384 + 8192));
    tma_load_2d_pair(sA, &tmA, full, 0, static_cast<int>(rank) * 128);
    tma_load_2d_pair(sB, &tmB, full, 0, static_cast<int>(rank) * 64);
    if (rank == 0) {
      mbar_wait(full, 0, 1);
      tc_fence_after();
      const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA), 16, 1024);
      const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB), 16, 1024);
      constexpr uint32_t idesc = make_idesc_bf16(256, 128, 0, 0);
      for (int k = 0; k < 4; ++k) umma_bf16_2cta(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
      umma_commit_2cta(done, 3);
    }
  }
  __syncwarp();
  mbar_wait(done, 0, 2);
  tc_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(rank * 128 + warp * 32 + lane) * 128 + c0 + j] = v[j];
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem, 128);
}

}  // namespace
}  // namespace stx

extern "C" int stx_debug_pair_probe(const void* a_bf16, const void* b_bf16, float* out, void* stream) {
  using namespace stx;
  STX_REQUIRE(a_bf16 && b_bf16 && out, "pair probe: null argument");
  CUtensorMap tmA, tmB;
  const uint64_t ad[2] = {64, 256};
  const uint32_t ab[2] = {64, 128};
  STX_TRY(encode_tmap_bf16(&tmA, a_bf16, 2, ad, ab));
  const uint64_t bd[2] = {64, 128};
  const uint32_t bb[2] = {64, 64};
  STX_TRY(encode_tmap_bf16(&tmB, b_bf16, 2, bd, bb));
  const int smem = 16384 + 8192 + 64 + 1024;
  STX_CUDA(cudaFuncSetAttribute(pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  pair_probe_kernel<<<2, 128, smem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, out);
  STX_CUDA(cudaGetLastError());
  return STX_OK;
}

extern "C" int stx_debug_patch_probe(const void* patch_bf16, int rows, const void* b_bf16, int start_row,
                                     int sbo_bytes, int base_mode, float* out, void* stream) {
  using namespace stx;
  STX_REQUIRE(patch_bf16 && b_bf16 && out, "probe: null argument");
  STX_REQUIRE(rows >= 128 && rows <= 256, "probe: rows must be in [128,256]");
  CUtensorMap tmP, tmB;
  const uint64_t pd[2] = {64, static_cast<uint64_t>(rows)};
  const uint32_t pb[2] = {64, static_cast<uint32_t>(rows)};
  STX_TRY(encode_tmap_bf16(&tmP, patch_bf16, 2, pd, pb));
  const uint64_t bd[2] = {64, 64};
  const uint32_t bb[2] = {64, 64};
  STX_TRY(encode_tmap_bf16(&tmB, b_bf16, 2, bd, bb));
  const int smem = 32768 + 8192 + 64 + 1024;
  STX_CUDA(cudaFuncSetAttribute(patch_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  patch_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(tmP, tmB, rows, start_row, sbo_bytes,
                                                                         base_mode, out);
  STX_CUDA(cudaGetLastError());
  return STX_OK;
}
