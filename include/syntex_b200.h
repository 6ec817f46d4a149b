/*
 * syntex_b200 — C ABI of the B200-native SynTex hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types. The reference (swift-n-brutal/syntex)
 * has no FFI of its own — its boundary is the Python surface of texture_utils / gatys.Synthesizer / po.model,
 * which builds TensorFlow graph ops. Each entry point below replaces the arithmetic of one of those call sites
 * (cited as path:line relative to the reference root). The Python mirror of the reference interface lives in
 * syntex_b200/{texture_utils,gatys,po}.py and calls these through ctypes.
 *
 * Conventions
 *   - All tensors are dense NHWC. Images/gradients are fp32, feature maps are bf16 on the device, Gram matrices
 *     and losses are fp32.
 *   - Every function returns 0 on success and a negative code on failure; stx_last_error() returns the message
 *     for the calling thread. Nothing falls back to a CPU path: without a CUDA device every compute call fails.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream). Calls are asynchronous on that stream
 *     unless documented as synchronous (the *_host variants).
 *   - Layer indices follow texture_utils/vgg19.py:11-30 (PARAM_SHAPE order up to pool4):
 *       0 conv1_1  1 conv1_2  2 pool1  3 conv2_1  4 conv2_2  5 pool2  6 conv3_1  7 conv3_2  8 conv3_3  9 conv3_4
 *       10 pool3  11 conv4_1  12 conv4_2  13 conv4_3  14 conv4_4  15 pool4
 */
#ifndef SYNTEX_B200_H_
#define SYNTEX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STX_NUM_LAYERS 16
#define STX_NUM_CONVS 12

#define STX_OK 0
#define STX_ERR_INVALID (-1)
#define STX_ERR_CUDA (-2)
#define STX_ERR_UNSUPPORTED (-3)
#define STX_ERR_STATE (-4)

typedef struct stx_vgg stx_vgg;     /* VGG19 weights up to pool4, packed for the tensor-core kernels */
typedef struct stx_plan stx_plan;   /* one f/g evaluation pipeline for a fixed (batch, height, width) */
typedef struct stx_lbfgs stx_lbfgs; /* device-resident L-BFGS(-B) state bound to a plan */

const char* stx_last_error(void);
int stx_abi_version(void);
/* Kernel launches issued by this library so far in this process (CUDA-graph replays count their kernel nodes). */
long long stx_launch_count(void);
/* Host helper, no GPU involved: CRC-32C (Castagnoli) of n bytes, continuing `crc` (0 to start) - the per-tensor checksum
 * of TensorFlow checkpoint files, which syntex_b200/po/tf_bundle.py reads and writes (reference: tensorpack's
 * ModelSaver / SaverRestore, po/progressive_model.py:354-356). */
unsigned int stx_crc32c(const void* data, size_t n, unsigned int crc);
/* Bring-up knobs; key/value meanings are internal (csrc/capi.cu: descriptor strides, kernel-variant overrides, 9 =
 * programmatic dependent launch, 10 = plain stream launches instead of the optimiser's CUDA graph, 12 = one CTA per SM
 * for the 64-channel f16+f8 forward, 13 = CUDA-core conv1_1 in f16+f8 plans, 14 = active-set kernel in the main
 * stream, 15 = CTA-pair convolution instances: bit 0 forward, bit 1 input gradient, default 3, 16 = full ReLU mask planes
 * instead of sign bits, 17 = keep evaluating the finished jobs of a batch, 18 = bit mask of the convolutions (1 =
 * conv1_2 .. 11 = conv4_4) of an f16+f8 plan that run without the correction product, -1 = the plan's own choice). Also read from the environment at load: STX_DEBUG="key=value,...". */
int stx_debug_set(int key, int value);
/* Bring-up probe of the CTA-pair MMA (tcgen05 cta_group::2): out fp32 [256][128] = a bf16 [256][64] * b bf16
 * [128][64]^T computed by one cluster of two CTAs (each loads its 128 rows of a and 64 rows of b). */
int stx_debug_pair_probe(const void* a_bf16, const void* b_bf16, float* out, void* stream);
/* Descriptor-addressing probe used while bringing up the halo-patch convolution (see csrc/probe.cu). */
int stx_debug_patch_probe(const void* patch_bf16, int rows, const void* b_bf16, int start_row, int sbo_bytes,
                          int base_mode, float* out, void* stream);

/* ---- weights: texture_utils/vgg19.py:97-130 (Vgg19.__init__) and :192-206 (get_var, constants) ------------
 * weights_hwio[i] / biases[i]: host fp32 arrays for conv layer i in network order (conv1_1 .. conv4_4), laid out
 * (3,3,Cin,Cout) and (Cout,) as stored by texture_utils/caffe2npz.py:16-18. The conv1_1 input-channel flip for
 * RGB input (vgg19.py:124-127) is the caller's job, as it is the loader's job in the reference.
 * mean3: per-channel mean subtracted after the x255 scaling (vgg19.py:142-146), in the channel order of the input. */
int stx_vgg_create(stx_vgg** out, const float* const* weights_hwio, const float* const* biases, int num_convs,
                   const float* mean3, void* stream);
void stx_vgg_destroy(stx_vgg* vgg);
/* Device pointers to the packed bf16 weights of conv layer `conv_index` (network order, 1 = conv1_2; conv1_1 runs on
 * its own fp32 kernel): forward pack [9][cout][cin] high and low halves, input-gradient pack [9][cin][cout]. Any
 * output may be NULL. Used by callers that drive stx_conv_bf16 themselves with the network's weights (the tangent
 * sweep of the PO double backward, syntex_b200/po/texture_op.py). */
int stx_vgg_weights(const stx_vgg* vgg, int conv_index, const void** fwd_hi, const void** fwd_lo, const void** dgrad);

/* ---- plan: the graph that gatys/synthesizer.py:54-107 (Synthesizer.build) and po/model.py:51-69 construct ----
 * topmost: last layer index to compute (15 = pool4, the Vgg19Extractor default, vgg19_extractor.py:5-6).
 * loss_layers/coefs: the Gram layers and weights (Synthesizer.DEFAULT_COEFS, synthesizer.py:17-23).
 * flags: 0 = accurate forward ("bf16x3": activations and weights carried as bf16 hi+lo pairs, three tensor-core
 *        MMAs per K step, forward error ~1e-5, input gradient within BASELINE.json's 1e-2);
 *        STX_PLAN_FAST_BF16 = single bf16 MMA forward (Grams within 5e-3, input gradient ~3e-2: ReLU sign flips). */
#define STX_PLAN_FAST_BF16 1
/* keep, for every loss layer k, the total gradient d sum(loss_overall) / d feat[k] of each step (what the PO models'
 * stage preparation takes with tf.gradients(loss, [feat[k]...]), po/improved_model.py:227-231) */
#define STX_PLAN_EXPORT_LAYER_GRADS 2
/* keep the low half of EVERY recorded layer. Without it the accurate forward skips the low-half store of the
 * convolutions that only feed a pooling layer (conv1_2, conv2_2, conv3_4, conv4_4): nothing on the f/g path reads it,
 * and stx_plan_copy_feature_f32 then returns those four layers rounded to bf16. Vgg19.build sets this flag. */
#define STX_PLAN_FULL_FEATURES 4
/* f16 + f8 forward: activations as fp16 planes plus e4m3 correction planes ([A8 | A_lo8] per 64-channel chunk), weights
 * as an fp16 pack plus e4m3 correction rows; per K step ONE kind::f16 MMA and ONE kind::f8f6f4 MMA (twice the MACs in
 * the same tensor-pipe time) instead of the three bf16 MMAs of the default forward - two MMAs' time per algorithmic
 * MAC at the same (bf16x3-class) accuracy of the pre-activations. Needs every convolution's map >= 8 pixels wide
 * (image side >= 64); such a plan does not serve stx_plan_feature_bf16. */
#define STX_PLAN_F16F8 8
/* With STX_PLAN_F16F8: conv4_2, conv4_3 and conv4_4 go WITHOUT the correction product (fp16 operands, one MMA per K
 * step; same plane formats). Measured against the float64 oracle at 64^2..512^2 (precision): the
 * input gradient stays at 3.3-5.8e-3 (all layers corrected: 3.2-5.6e-3; bar 1e-2), the pool4 Gram goes from 1e-4 to
 * 4.4-5.3e-4 (bar 1e-3), the forward pass of 64 jobs from 4.45 to 3.84 ms. OPT-IN (precision "f16f8-lean"): one
 * evaluation is inside the bars, but over a long optimisation the loss's rounding noise reaches the line search (1000
 * iterations: 2 of 8 jobs stop early, final loss 5-7 % above scipy's instead of 0-4 %). */
#define STX_PLAN_F16F8_LEAN 16
int stx_plan_create(stx_plan** out, const stx_vgg* vgg, int batch, int height, int width, int topmost,
                    int num_loss_layers, const int* loss_layers, const float* coefs, int flags);
void stx_plan_destroy(stx_plan* plan);
size_t stx_plan_device_bytes(const stx_plan* plan);

/* Vgg19.build (vgg19.py:132-169): image [batch,H,W,3] fp32 device, values in [0,1] (or a pre-image that is passed
 * through a sigmoid first, synthesizer.py:59). Computes every layer up to topmost; with_grams != 0 also computes
 * the Gram matrices of the loss layers (build_gram, utils.py:8-21). */
int stx_plan_forward(stx_plan* plan, const float* image, int is_preimage, int with_grams, void* stream);
/* Copy one recorded layer out as fp32 [batch,h,w,c] (device). */
int stx_plan_layer_shape(const stx_plan* plan, int layer, int* h, int* w, int* c);
int stx_plan_copy_feature_f32(stx_plan* plan, int layer, float* out, void* stream);
/* Copy the Gram matrix of loss layer k (index into loss_layers) as fp32 [batch,C,C] (device). */
int stx_plan_copy_gram(stx_plan* plan, int k, float* out, void* stream);
/* Target Grams: the buf/gram_* variables and their assign ops (synthesizer.py:66-74). Either assign the Grams of
 * the last forward, or set them from a device array [batch,C,C] (per_image = 1) / [1,C,C] (broadcast). */
int stx_plan_update_target_from_forward(stx_plan* plan, void* stream);
int stx_plan_set_target_gram(stx_plan* plan, int k, const float* gram, int per_image, void* stream);
/* A target is complete only once every loss layer has been set (or update_target_from_forward has run); until then
 * stx_plan_step / stx_lbfgs_* / stx_plan_layer_grad return STX_ERR_STATE. stx_plan_clear_target invalidates the
 * current target set (call it before re-using a plan for a new set of targets). */
int stx_plan_clear_target(stx_plan* plan);
/* Number of forward passes this plan has run. A caller that keeps a plan's activations for a later backward
 * (po/texture_op.py) records it at forward time and compares before the backward: a mismatch means another forward
 * has overwritten the activations. */
long long stx_plan_generation(const stx_plan* plan);
/* Read the target Gram of loss layer k back as fp32 [batch,C,C] (device; a broadcast target is expanded). */
int stx_plan_copy_target_gram(stx_plan* plan, int k, float* out, void* stream);

/* Synthesizer.step (synthesizer.py:119-125): one f/g evaluation.
 *   func [batch]         = loss_overall  (build_texture_loss, utils.py:44-49)
 *   grad [batch,H,W,3]   = d sum(loss_overall) / d x   (tf.gradients, synthesizer.py:79), w.r.t. the pre-image
 *                          when is_preimage.
 * x, func, grad are device pointers. layer_loss (nullable) receives the unweighted per-layer losses
 * [num_loss_layers][batch]. */
int stx_plan_step(stx_plan* plan, const float* x, int is_preimage, float* func, float* grad, float* layer_loss,
                  void* stream);
/* Same through host buffers (what sess.run does for the reference): H2D of x, f/g, D2H of func and grad.
 * Synchronous. */
int stx_plan_step_host(stx_plan* plan, const float* x_host, int is_preimage, float* func_host, float* grad_host,
                       void* stream);
/* Instrumented step for the benchmark's roofline: CUDA events around every kernel launch of one f/g evaluation.
 * ms[8]/count[8] (host) receive device milliseconds and launch counts per kernel class:
 *   0 conv3x3 forward (tcgen05)  1 conv3x3 dgrad (tcgen05)  2 Gram dF 1x1 GEMM (tcgen05)  3 Gram (tcgen05)
 *   4 Gram finish/loss  5 conv1_1 fwd+bwd  6 pooling fwd+bwd  7 other.   Synchronous. */
int stx_plan_step_timed(stx_plan* plan, const float* x, int is_preimage, float* func, float* grad, float* ms,
                        int* count, void* stream);
/* After a step on a plan created with STX_PLAN_EXPORT_LAYER_GRADS: d sum(loss_overall) / d feat[loss layer k],
 * back-propagated from every loss layer above and including k, fp32 [batch,h,w,c] (device). */
int stx_plan_copy_layer_total_grad(stx_plan* plan, int k, float* out, void* stream);
/* Per-layer gradients of build_texture_loss(calc_grad=True) (utils.py:50-55): d loss_layer[k] / d feat[k],
 * fp32 [batch,h,w,c] device, for the features of the last forward and the current targets. */
int stx_plan_layer_grad(stx_plan* plan, int k, float* out, void* stream);
/* The same in the kernels' own precision, bf16 [batch,h,w,c] (device): what the hand-written PO generator consumes
 * (syntex_b200/po/gen_ops.py), without the fp32 round trip. */
int stx_plan_layer_grad_bf16(stx_plan* plan, int k, void* out_bf16, void* stream);

/* Zero-copy view of a recorded layer of the last forward: device pointers to the bf16 planes [batch,h,w,c] the
 * plan owns (hi = bf16(value); lo = bf16(value - hi) in the accurate mode, NULL in STX_PLAN_FAST_BF16 mode; `lo` may
 * be NULL when not wanted). Valid until the next forward on this plan. */
int stx_plan_feature_bf16(stx_plan* plan, int layer, const void** hi, const void** lo);

/* Backward of the last stx_plan_forward with caller-supplied per-layer matrices (the generalisation the PO models
 * need): for every loss layer k the gradient injected at feat[k] is
 *     dL/dfeat[k] = feat[k] * M[k] + E[k]
 * M[k]: fp32 [batch,C,C] device (rounded to bf16), E[k]: bf16 [batch,h,w,C] device or NULL (E itself may be NULL);
 * the sum is back-propagated through VGG19 (ReluGrad / AvgPoolGrad / Conv2DBackpropInput, as tf.gradients does in
 * po/progressive_model.py's training graph) to grad [batch,H,W,3] fp32. With M[k] = upstream_k * 4/(C^2 n) *
 * (G_k - T_k) this is the first-order style gradient (utils.py:46-54); the second-order term of differentiating
 * through grad_layer[k] = alpha F (G - T) adds alpha/n * (F^T U + U^T F) to M[k] and alpha * U (G - T) to E[k]. */
int stx_plan_backward_custom(stx_plan* plan, const float* image, int is_preimage, const float* const* M,
                             const void* const* E, float* grad, void* stream);

/* ---- device-resident L-BFGS-B: the optimiser loop of synthesizer.py:127-151 ------------------------------
 * Replaces scipy.optimize.minimize(method="L-BFGS-B", jac=True, bounds=[0,1]^n, options={maxiter, maxcor,
 * ftol=0, gtol=0}) with the same algorithm on the GPU (subspace minimisation over the free variables by the direct
 * primal method, projected step, More'-Thuente line search along the feasible segment with lnsrlb's constants and
 * termination rules; the one omission is the breakpoint walk of the generalized Cauchy point - csrc/lbfgs.cu,
 * DESIGN.md section 6) whose state (x, history, line search) never leaves the device.
 * Each of the `batch` images of the plan is an independent problem with its own history and line search. A problem
 * stops before `maxiter` when f does not decrease (L-BFGS-B's test at factr = 0) or a line search fails twice. */
int stx_lbfgs_create(stx_lbfgs** out, stx_plan* plan, int maxcor, int bounded, float lower, float upper);
void stx_lbfgs_destroy(stx_lbfgs* opt);
/* x0: device [batch,H,W,3]. Resets the history. */
int stx_lbfgs_reset(stx_lbfgs* opt, const float* x0, int is_preimage, void* stream);
/* Run f/g evaluations until every problem has completed `maxiter` iterations (or `max_evals` evaluations were
 * spent, 0 = 3*maxiter). Asynchronous; no host round trip inside. */
int stx_lbfgs_run(stx_lbfgs* opt, int maxiter, int max_evals, void* stream);
/* Results (device -> caller's device buffers): x [batch,H,W,3], fun [batch], iterations and evaluations [batch]. */
int stx_lbfgs_result(stx_lbfgs* opt, float* x, float* fun, int* nit, int* nfev, void* stream);
/* Diagnostic: SM clock stamps (clock64) of the phases of problem 0's decision kernel in the last cycle, host array of
 * 16: entry, sums reduced, line search decided, K built, eliminated, solved, state written back. Synchronises. */
int stx_lbfgs_phase_cycles(stx_lbfgs* opt, long long* host16);
/* Whole optimisation through host buffers: H2D of x0, run, D2H of x and fun. Synchronous. This is what
 * Synthesizer.synthesize() (gatys/synthesizer.py:127-151) calls; the _f64 form returns x as float64, as the reference
 * does (scipy's result vector), widening while it copies out of the pinned staging buffer. */
int stx_lbfgs_minimize_host(stx_lbfgs* opt, const float* x0_host, int is_preimage, int maxiter, float* x_host,
                            float* fun_host, int* nit_host, int* nfev_host, void* stream);
int stx_lbfgs_minimize_host_f64(stx_lbfgs* opt, const float* x0_host, int is_preimage, int maxiter, double* x_host,
                                float* fun_host, int* nit_host, int* nfev_host, void* stream);

/* ---- stand-alone operators (building blocks; used by the Python texture_utils mirror and the tests) --------- */
/* HWIO fp32 (3,3,cin,cout) -> bf16 forward pack [9][cout][cin] and dgrad pack [9][cin][cout] (device pointers). */
int stx_pack_conv_weights(const float* w_hwio, int cin, int cout, void* w_fwd_bf16, void* w_dgrad_bf16,
                          void* stream);
/* Same plus the low halves bf16(w - bf16(w)) of the forward pack (for stx_conv_bf16x3). Any output may be NULL. */
int stx_pack_conv_weights_split(const float* w_hwio, int cin, int cout, void* w_fwd_hi, void* w_fwd_lo,
                                void* w_dgrad_bf16, void* stream);
/* bf16x3 forward: relu(conv(x_hi + x_lo, w_hi + w_lo) + bias) -> out_hi, out_lo (3x3 only). */
int stx_conv_bf16x3(const void* x_hi, const void* x_lo, int n, int H, int W, int cin, const void* w_hi,
                    const void* w_lo, int cout, const float* bias, int relu, void* out_hi, void* out_lo,
                    int force_bn, void* stream);
/* out = epilogue(conv(x, w)) with x [n,H,W,cin] bf16, w packed [taps][cout][cin] bf16, taps in {9, 1}.
 * epilogue: (+ bias[cout] fp32) (+ addend bf16) (relu) (zero where mask <= 0); any of them may be NULL/0.
 * force_bn ("variant"): 0 = automatic kernel and tile choice; otherwise bn + 1000 * mt + 100000 * per_tap + 1000000 *
 *   pair, with bn in {0, 64, 128, 256} the N tile, mt in {0, 1, 2} the stacked 16x8 output tiles of the halo-patch
 *   kernel, per_tap = 1 forcing the general per-tap kernel (conv_igemm.cu) instead of the halo-patch one
 *   (conv_patch.cu) and pair = 1 asking for the CTA-pair instance (tcgen05 cta_group::2) where the shape qualifies. */
int stx_conv_bf16(const void* x, int n, int H, int W, int cin, const void* w_packed, int cout, int taps,
                  const float* bias, const void* addend, const void* mask, int relu, void* out, int force_bn,
                  void* stream);
/* conv1_1 (+preprocessing) forward and input-gradient. w27x64/b64 fp32 device. */
int stx_conv1_1_fwd(const float* image, int n, int H, int W, int is_preimage, const float* mean3_host,
                    const float* w27x64, const float* b64, void* out_bf16, void* out_lo_bf16_or_null, void* stream);
int stx_conv1_1_bwd(const void* g_bf16, const float* image, int n, int H, int W, int is_preimage,
                    const float* w27x64, float* grad_out, void* stream);
int stx_avgpool2_fwd(const void* x_bf16, const void* x_lo_or_null, int n, int H, int W, int C, void* out_bf16,
                     void* out_lo_or_null, void* stream);
int stx_avgpool2_bwd(const void* g_bf16, const void* act_bf16_or_null, int n, int H, int W, int C, void* out_bf16,
                     void* stream);
/* build_gram (utils.py:8-21) on a bf16 feature map [n,H*W,C]: G [n,C,C] fp32. workspace: stx_gram_workspace_bytes. */
size_t stx_gram_workspace_bytes(int n, int HW, int C);
int stx_gram_bf16(const void* feat_bf16, int n, int HW, int C, float eps, float* G_out, void* workspace,
                  void* stream);
/* A^T B / (HW + eps) for two bf16 feature maps [n,HW,C] (the F^T U contraction of the double backward through
 * build_texture_loss(calc_grad=True), utils.py:50-55): out [n,C,C] fp32, not symmetric. */
size_t stx_gram_cross_workspace_bytes(int n, int HW, int C);
int stx_gram_cross_bf16(const void* a_bf16, const void* b_bf16, int n, int HW, int C, float eps, float* out,
                        void* workspace, void* stream);
/* ---- PO generator building blocks (po/progressive_model.py:77-206) -------------
 * All tensors NHWC bf16 on the device unless noted. These replace, for the generator's trainable 3x3 layers, the
 * arithmetic of tensorpack Conv2D / InstanceNorm / tf.pad(SYMMETRIC) and of tf.gradients through them. */
/* out[i] = x[i] * w[i]^T per image: x [n,H,W,cin] bf16, w [n][cout][cin] bf16 (one matrix per image), out
 * [n,H,W,cout] bf16 - the d loss_k / d F_k = alpha F (G - T) product of build_texture_loss(calc_grad=True)
 * (utils.py:50-55) and the F M products of the second-order term, for the whole batch in ONE launch. */
int stx_conv1x1_per_image_bf16(const void* x, int n, int H, int W, int cin, const void* w, int cout, void* out,
                               void* stream);
/* 3x3 convolution with an explicit padding mode. H, W are the OUTPUT size; the input is
 * (H + 2 - 2 pad) x (W + 2 - 2 pad): pad 0 = VALID on an explicitly padded input (tf.pad + Conv2D("VALID"),
 * progressive_model.py:171-175), 1 = SAME (zero padding), 2 = FULL correlation (the input gradient of the VALID
 * convolution when w_packed is the dgrad pack of stx_pack_conv_weights). Epilogue as stx_conv_bf16. */
int stx_conv_bf16_pad(const void* x, int n, int H, int W, int cin, const void* w_packed, int cout, int pad,
                      const float* bias, const void* addend, int relu, void* out, int variant, void* stream);
/* Filter gradient (what tf.gradients gives for Conv2D's kernel): dw_hwio fp32 [3][3][cin][cout] (device) =
 * sum over images and output pixels of x[h + r - pad, w + s - pad, ci] * dy[h, w, co]; x is
 * [n, H + 2 - 2 pad, W + 2 - 2 pad, cin] (pad 0: explicitly padded; pad 1: taps outside read zeros), dy [n,H,W,cout].
 * workspace: stx_conv_wgrad_workspace_bytes. Deterministic (fixed-order split-K). */
size_t stx_conv_wgrad_workspace_bytes(int n, int H, int W, int cin, int cout);
int stx_conv_wgrad_bf16(const void* x, const void* dy, int n, int H, int W, int cin, int cout, int pad,
                        float* dw_hwio, void* workspace, void* stream);
/* The same with the result in torch's OIHW layout [cout_real][cin][3][3] (cout_real <= cout: the zero-padded filters
 * of a narrow layer are dropped), so that it adds straight into an nn.Parameter's .grad. */
int stx_conv_wgrad_oihw_bf16(const void* x, const void* dy, int n, int H, int W, int cin, int cout, int pad,
                             int cout_real, float* dw_oihw, void* workspace, void* stream);
/* OIHW fp32 [cout][cin][3][3] (device) -> the forward pack [9][cout_padded][cin] and dgrad pack [9][cin][cout_padded]
 * of stx_pack_conv_weights, output channels >= cout zero. Any output may be NULL. */
int stx_pack_conv_weights_oihw(const float* w_oihw, int cin, int cout, int cout_padded, void* w_fwd_bf16,
                               void* w_dgrad_bf16, void* stream);
/* tf.pad(mode="SYMMETRIC") by one pixel (= edge replication), optionally of relu(x): out [n,H+2,W+2,C]. */
int stx_pad_replicate_bf16(const void* x, int n, int H, int W, int C, int relu, void* out, void* stream);
/* Its gradient: gp [n,H+2,W+2,C] -> dx [n,H,W,C] (ring folded onto the edge pixels); x_or_null non-NULL = the pad
 * was taken of relu(x): zero where x <= 0. */
int stx_pad_replicate_bwd_bf16(const void* gp, const void* x_or_null, int n, int H, int W, int C, void* dx,
                               void* stream);
/* General form: mode 0 = SYMMETRIC by one pixel, 1 = REFLECT (tf.pad modes of po/improved_model.py:49-55; CONSTANT
 * needs no copy: stx_conv_bf16_pad with pad = 1); act 0 = none, 1 = x > 0 ? x : slope * x applied on the way (slope 0
 * = ReLU, 0.2 = the LeakyReLU of po/improved_model.py:30-47). The backward applies the activation's derivative when
 * x_or_null is the forward input. */
int stx_pad2d_bf16(const void* x, int n, int H, int W, int C, int mode, int act, float slope, void* out, void* stream);
int stx_pad2d_bwd_bf16(const void* gp, const void* x_or_null, int n, int H, int W, int C, int mode, float slope,
                       void* dx, void* stream);
/* tensorpack InstanceNorm (per image and channel over H*W, biased variance, affine). mean/rstd fp32 [n,C]. */
size_t stx_inorm_workspace_bytes(int n, int HW, int C);
int stx_inorm_stats_bf16(const void* x, int n, int HW, int C, float eps, float* mean, float* rstd, void* workspace,
                         void* stream);
/* out = act((x - mean) * rstd * gamma + beta) [+ addend]   (addend: the residual shortcut, nullable)
 * act 0 = none, 1 = v > 0 ? v : slope * v  (slope 0 = ReLU) */
int stx_inorm_apply_bf16(const void* x, int n, int HW, int C, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, int act, float slope, const void* addend, void* out,
                         void* stream);
/* Backward of the above for upstream dy (w.r.t. the value before the addend): dx, and per image the sums
 * s1 = sum g, s2 = sum g * xhat with g = dy * act'  (d beta = sum_n s1, d gamma = sum_n s2), fp32 [n,C]. */
int stx_inorm_bwd_bf16(const void* x, const void* dy, int n, int HW, int C, const float* mean, const float* rstd,
                       const float* gamma, const float* beta, int act, float slope, float* s1, float* s2, void* dx,
                       void* workspace, void* stream);

/* tf.train.AdamOptimizer(lr, beta1, beta2, epsilon) as the PO trainers use it (po/progressive_model.py:268-271):
 *   lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t);  param -= lr_t m / (sqrt(v) + eps)     ("epsilon hat")
 * over flat fp32 device buffers of `count` elements (16-byte aligned); lr_dev and step_dev (t as a float, already
 * incremented for this step) are device scalars so that a CUDA-graph replay sees the current values. */
int stx_adam_tf(float* param, const float* grad, float* m, float* v, size_t count, const float* lr_dev,
                const float* step_dev, float beta1, float beta2, float eps, void* stream);

int stx_f32_to_bf16(const float* x, size_t count, void* out_bf16, void* stream);
int stx_bf16_to_f32(const void* x_bf16, size_t count, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SYNTEX_B200_H_ */
