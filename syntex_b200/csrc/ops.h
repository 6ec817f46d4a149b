// Internal operator layer: one prepared-op struct + launcher per kernel family. The C-ABI (capi.cu) and the
// f/g plan (plan.cu) are built from these. All pointers are device pointers; all launches are asynchronous on
// the given stream and never synchronise.
#pragma once
#include "common.h"

namespace stx {

// ------------------------------------------------------------------ conv_igemm.cu
struct ConvOp {
  CUtensorMap tmA;   // activations [nimg,H,W,cin] bf16
  CUtensorMap tmB;   // packed weights [taps|nimg][cout][cin] bf16
  CUtensorMap tmA2;  // split mode: low halves of the activations
  CUtensorMap tmB2;  // split mode: low halves of the weights
  CUtensorMap tmF;   // fused Gram gradient: features of the output layer [nimg,H,W,cout]
  CUtensorMap tmD;   // fused Gram gradient: scaled Gram difference [nimg][cout][cout]
  int gram_chunks;   // cout / 64 when the Gram gradient is fused, else 0
  const bf16* w_tiled;     // patch kernel: tiled pre-swizzled weights (1-D bulk copies); null = use tmB
  const bf16* w_tiled_lo;  // split mode low halves
  int ksplit;        // split-K factor of the patch kernel (1 = off)
  float* splitk_ws;  // fp32 partial tiles (shared by all ops of a plan; launches are stream-ordered)
  int* splitk_counters;
  int split;         // 1 = bf16x3 forward (hi/lo operands, hi/lo outputs); 2 = f16 + f8 forward (patch kernel only:
                     // fp16 planes + e4m3 correction planes, conv_common.cuh kF8*); 3 = a layer of an f16 + f8 plan that
                     // runs without the correction product (fp16 operands, one MMA, same output planes)
  float corr_scale;  // split == 2: factor of the correction accumulator (1 / (512 * weight scale) of the layer)
  bf16* aux_out;     // split == 2: bf16 copy of the stored result (loss layers: the Gram gradient's A operand is bf16)
  bf16* pool_aux;    // ... and of the fused pool's result
  int patch;         // 1 = halo-patch kernel (conv_patch.cu), 0 = per-tap kernel (conv_igemm.cu)
  int mt;            // patch kernel: 16x8 output tiles stacked per CTA (1 or 2)
  int occ;           // patch kernel: persistent CTAs per SM (1 or 2)
  int pair;          // patch kernel: 1 = CTA pairs (tcgen05 cta_group::2): two M tiles share each weight tile, half per SM
  int res;           // patch kernel: 64-channel input chunks whose weights stay resident in shared memory (0 = ring)
  bf16* out_lo;
  bf16* pool_out;            // fused 2x2 average pool outputs (patch kernel only)
  bf16* pool_out_lo;
  bf16* up_out;              // fused AvgPoolGrad + ReluGrad scatter (replaces `out`)
  const bf16* up_mask;
  const int* active;         // per-image switch, nullable (conv_common.cuh)
  uint32_t* bits_out;        // one-bit ReLU mask of the result, written instead of `out` (conv_common.cuh)
  const uint32_t* up_bits;   // ... and read by the scatter instead of up_mask
  bf16* export_out;          // copy of the value after the addend (total gradient w.r.t. this layer's features)
  float* rgb_out;            // conv1_1 input-gradient mode of the patch kernel (bn == 16)
  const float* rgb_image;
  int rgb_preimage;
  int nimg, H, W, cin, cout, taps, per_image_w;
  int bw_log2, bh_log2, tiles_w, tiles_h, tiles_n, n_step, bn;
  int relu;
  int pad;             // rows/columns of (zero) padding the taps reach into: 1 = SAME (default), 0 = VALID on an input two
                       // pixels larger than the output, 2 = FULL on an input two pixels smaller (see conv_set_input_pad)
  const float* bias;
  const bf16* addend;
  const bf16* mask;
  bf16* out;
};
int conv_prepare(ConvOp* op, const bf16* x, int nimg, int H, int W, int cin, const bf16* wpacked, int cout,
                 int taps, int per_image_w, int force_bn);
int conv_prepare_split(ConvOp* op, const bf16* x_lo, const bf16* w_lo);
// Halo-patch variant (3x3 only, W >= 8); returns STX_ERR_UNSUPPORTED for shapes it does not cover.
int conv_patch_prepare(ConvOp* op, const bf16* x, int nimg, int H, int W, int cin, const bf16* wpacked, int cout,
                       int split, int force_bn, int force_mt, int force_occ, int want_pair = 0);
int conv_patch_prepare_split(ConvOp* op, const bf16* x_lo, const bf16* w_lo);
int conv_patch_attach_gram(ConvOp* op, const bf16* feat, const bf16* dscaled);
int conv_patch_plan_splitk(const ConvOp& op, size_t* ws_floats, size_t* counters);
int conv_patch_launch(const ConvOp& op, cudaStream_t stream);
// Picks the halo-patch kernel when the shape qualifies, else the per-tap kernel. x_lo / w_lo non-null = bf16x3.
// variant: 0 = automatic; otherwise bn + 1000 * mt + 10000 * occ + 100000 * (force the per-tap kernel)
// + 1000000 * (use CTA pairs where the shape qualifies).
int conv3x3_prepare(ConvOp* op, const bf16* x, const bf16* x_lo, int nimg, int H, int W, int cin, const bf16* w,
                    const bf16* w_lo, int cout, int variant);
int conv_launch(const ConvOp& op, cudaStream_t stream);
// Re-point a prepared 3x3 op (geometry = OUTPUT H x W) at an input of (H + 2 - 2 pad) x (W + 2 - 2 pad) pixels:
// pad 0 = VALID convolution of an explicitly padded input (the generator's symmetric padding), pad 2 = FULL
// correlation (the input gradient of a VALID convolution). Out-of-range taps read zeros (TMA fill).
int conv_set_input_pad(ConvOp* op, const bf16* x, int pad);

// ------------------------------------------------------------------ gram.cu
struct GramOp {
  CUtensorMap tmF;   // features viewed as [nimg][HW][C] bf16
  CUtensorMap tmU;   // cross mode: the second map (B operand)
  int cross = 0;     // 1 = A^T B of two maps, every tile computed
  int f16 = 0;       // feature planes are fp16 (f16+f8 plans) instead of bf16
  int nimg, HW, C;
  int bn;            // N tile (64 for C == 64, else 128)
  int ntiles;        // upper-triangle 128x128 tiles
  int splits;        // split-K factor over HW
  int iters_per_split;
  float* partial;    // [nimg][splits][C][C] fp32 (upper tiles written)
};
size_t gram_partial_bytes(int nimg, int HW, int C);
int gram_prepare(GramOp* op, const bf16* feat, int nimg, int HW, int C, float* partial);
size_t gram_cross_partial_bytes(int nimg, int HW, int C);
int gram_cross_prepare(GramOp* op, const bf16* a, const bf16* b, int nimg, int HW, int C, float* partial);
int gram_launch(const GramOp& op, cudaStream_t stream);
// Reduce the split-K partials into G = F^T F / (HW + eps) (mirrored to a full [nimg,C,C] matrix) and, when a
// target is given, the scaled difference Dscaled = (G - T) * dscale (bf16, [nimg][C][C]) and per-block partial
// sums of loss_scale * (G - T)^2 (one float per block; gram_finish_blocks() blocks per image).
int gram_finish_blocks(int C);
int gram_finish_launch(const GramOp& op, float eps, float* G_out /*nullable*/, const float* target /*nullable*/,
                       long long target_stride /*elements between images, 0 = broadcast*/, float dscale,
                       bf16* dscaled /*nullable*/, float loss_scale,
                       float* loss_part /*nullable, [nimg][part_stride], blocks written per image*/, int part_stride,
                       cudaStream_t stream);

// Several layers' finishes in one launch (the f/g plan: five small reductions become one).
constexpr int kMaxFinishLayers = 8;
struct GramFinishLayer {
  const float* partial;
  int C, splits, kw, block_begin;
  float norm, dscale, loss_scale;
  float* G_out;
  const float* target;
  long long target_stride;
  bf16* dscaled;
  float* loss_part;   // this layer's first column in the [nimg][part_stride] partial-loss table (nullable)
};
struct GramFinishBatch {
  GramFinishLayer l[kMaxFinishLayers];
  int n = 0, total_blocks = 0, nimg = 0, part_stride = 0;
};
int gram_finish_batch_add(GramFinishBatch* fb, const GramOp& op, float eps, float* G_out, const float* target,
                          long long target_stride, float dscale, bf16* dscaled, float loss_scale, float* loss_part);
int gram_finish_batch_launch(const GramFinishBatch& fb, cudaStream_t stream);

// ------------------------------------------------------------------ elementwise.cu
// conv1_1: fp32 image in [0,1] (optionally a pre-image passed through a sigmoid) -> x*255 - mean ->
// 3x3 SAME conv (3 -> 64) + bias + ReLU -> bf16 NHWC.   w: [27][64] fp32 ((r*3+s)*3+c major), b: [64]
// out_lo (nullable): low halves, out + out_lo carries the fp32 result to ~16 mantissa bits.
// fmt 2: out = fp16 plane, out_lo = e4m3 correction plane (conv_common.cuh, kF8*), aux (nullable) = bf16 copy.
int conv1_1_fwd_launch(const float* image, int nimg, int H, int W, int is_preimage, const float* mean3,
                       const float* w, const float* b, bf16* out, bf16* out_lo, cudaStream_t stream, int fmt = 0,
                       bf16* aux = nullptr);
// f16 + f8 forward packs of one layer: w16 [9][cout][cin] fp16, wcorr [9][cout][cin/64][128 B] e4m3 (W_lo | W), with
// the per-layer power-of-two weight scale sw (the layer's correction accumulator is then scaled by 1 / (512 sw)).
int pack_weights_f16f8_launch(const float* w_hwio, int cin, int cout, float sw, bf16* w16, bf16* wcorr,
                              cudaStream_t stream);
// conv1_1 of the f16 + f8 plans on the tensor cores (conv1_1_tc.cu): K = 27 im2col rows built in shared memory.
// wpack: 16 KB from pack_conv1_1_tc_launch (fp16 tile + e4m3 correction tile, pre-swizzled), sw its weight scale.
int pack_conv1_1_tc_launch(const float* w27x64, float sw, bf16* out16k, cudaStream_t stream);
// active: per-image switch (nullable, see ConvKernelParams)
int conv1_1_tc_launch(const float* image, int nimg, int H, int W, int is_preimage, const float* mean3, const bf16* wpack,
                      float corr_scale, const float* bias, bf16* out, bf16* out_corr, bf16* aux, cudaStream_t stream,
                      const int* active = nullptr);
// fp16 plane (+ correction plane, nullable) of a recorded layer with C channels -> fp32.
int f16f8_to_f32_launch(const bf16* hi, const bf16* corr, size_t n, int C, float* out, cudaStream_t stream);
// dgrad of conv1_1 chained with the preprocessing derivative: g [nimg,H,W,64] bf16 (already ReLU-masked) ->
// dL/d(input image or pre-image) fp32 [nimg,H,W,3].
int conv1_1_bwd_launch(const bf16* g, const float* image, int nimg, int H, int W, int is_preimage, const float* w,
                       float* grad_out, cudaStream_t stream);
// x_lo / out_lo nullable (split activations: pooled in fp32 from hi + lo, written back as hi / lo).
int avgpool2_fwd_launch(const bf16* x, const bf16* x_lo, int nimg, int H, int W, int C, bf16* out, bf16* out_lo,
                        cudaStream_t stream);
// g_in [nimg,H/2,W/2,C] -> out[h,w,c] = 0.25 * g_in[h/2,w/2,c] * (act[h,w,c] > 0)   (act nullable: no mask)
int avgpool2_bwd_mask_launch(const bf16* g_in, const bf16* act, int nimg, int H, int W, int C, bf16* out,
                             cudaStream_t stream);
int f32_to_bf16_launch(const float* x, size_t n, bf16* out, cudaStream_t stream);
int bf16_to_f32_launch(const bf16* x, size_t n, float* out, cudaStream_t stream);
int bf16pair_to_f32_launch(const bf16* hi, const bf16* lo, size_t n, float* out, cudaStream_t stream);
// HWIO fp32 [3][3][cin][cout] -> forward pack [tap][cout][cin] and dgrad pack [tap'][cin][cout] (tap' = 8 - tap)
// w_fwd_lo (nullable): low halves bf16(w - bf16(w)) in the forward layout.
int pack_weights_launch(const float* w_hwio, int cin, int cout, bf16* w_fwd, bf16* w_fwd_lo, bf16* w_dgrad,
                        cudaStream_t stream);
// OIHW fp32 [cout][cin][3][3] (torch layout) -> the same two packs with cout zero-padded to cout_p.
int pack_weights_oihw_launch(const float* w_oihw, int cin, int cout, int cout_p, bf16* w_fwd, bf16* w_dgrad,
                             cudaStream_t stream);
// Sum loss partials: out[img] = sum_k part[img*stride + k], fixed order (deterministic).
// conv1_1 weights fp32 [27][64] -> dgrad pack for the tensor-core path: [9][16][64] bf16, rows 0..2 = colours
// (taps rotated 180 degrees), rows 3..15 zero.
// Tiled + pre-swizzled packs (see elementwise.cu): forward hi / lo and dgrad, any of them nullable.
int pack_weights_tiled_launch(const float* w_hwio, int cin, int cout, bf16* f_hi, bf16* f_lo, bf16* d_hi,
                              cudaStream_t stream);
int pack_conv1_1_dgrad_launch(const float* w27x64, bf16* out, cudaStream_t stream);
int sum_partials_launch(const float* part, int nimg, int count, int stride, float* out, cudaStream_t stream);


// ------------------------------------------------------------------ wgrad.cu
// Filter gradient of a 3x3 convolution: dw_hwio fp32 [3][3][cin][cout] = sum over images/pixels of
// X[h + r - pad, w + s - pad] (x) dY[h, w]; x is [nimg, H + 2 - 2 pad, W + 2 - 2 pad, cin], dy [nimg, H, W, cout].
size_t wgrad_workspace_bytes(int nimg, int H, int W, int cin, int cout);
int wgrad_launch(const bf16* x, const bf16* dy, int nimg, int H, int W, int cin, int cout, int pad, float* dw_hwio,
                 float* workspace, cudaStream_t stream);
// Same contraction, result written in torch's OIHW layout [cout_real][cin][3][3] (output channels >= cout_real, the
// zero-padded filters of a narrow layer, are dropped).
int wgrad_oihw_launch(const bf16* x, const bf16* dy, int nimg, int H, int W, int cin, int cout, int pad, int cout_real,
                      float* dw_oihw, float* workspace, cudaStream_t stream);

// ------------------------------------------------------------------ gen_ops.cu
// mode: 0 = SYMMETRIC (edge replication), 1 = REFLECT; act: 0 none, 1 = leaky ReLU with `slope` (0 = ReLU)
int pad_rep_launch(const bf16* x, int nimg, int H, int W, int C, int mode, int act, float slope, bf16* out,
                   cudaStream_t stream);
int pad_rep_bwd_launch(const bf16* gp, const bf16* x_or_null, int nimg, int H, int W, int C, int mode, float slope,
                       bf16* dx, cudaStream_t stream);
size_t inorm_workspace_bytes(int nimg, int HW, int C);
int inorm_stats_launch(const bf16* x, int nimg, int HW, int C, float eps, float* mean, float* rstd, float* workspace,
                       cudaStream_t stream);
int inorm_apply_launch(const bf16* x, int nimg, int HW, int C, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, int relu, float slope, const bf16* addend, bf16* out, cudaStream_t stream);
int inorm_bwd_launch(const bf16* x, const bf16* dy, int nimg, int HW, int C, const float* mean, const float* rstd,
                     const float* gamma, const float* beta, int relu, float slope, float* s1, float* s2, bf16* dx,
                     float* workspace, cudaStream_t stream);

// tf.train.AdamOptimizer update over flat fp32 buffers (see gen_ops.cu); lr_dev / step_dev: device scalars.
int adam_tf_launch(float* p, const float* g, float* m, float* v, size_t n, const float* lr_dev, const float* step_dev,
                   float b1, float b2, float eps, cudaStream_t stream);

}  // namespace stx
