// The f/g evaluation plan: every buffer, tensor map and launch of one Synthesizer.step
// (reference: gatys/synthesizer.py:54-107 builds the same thing as a TF graph) prepared once for a fixed
// (batch, height, width) and replayed without allocation or host synchronisation.
#pragma once
#include <vector>

#include "ops.h"

namespace stx {

constexpr int kNumLayers = 16;
constexpr int kNumConvs = 12;

struct LayerSpec {
  bool is_conv;
  int conv_index;   // index among conv layers, -1 for pools
  int cin, cout;
};
const LayerSpec& layer_spec(int layer);

struct Vgg {
  int num_convs = 0;
  float mean[3] = {0, 0, 0};
  float* w0 = nullptr;            // conv1_1 weights fp32 [27][64]
  bf16* wd0 = nullptr;            // conv1_1 dgrad pack [9][16][64] (tensor-core input-gradient path)
  bf16* wtc0 = nullptr;           // conv1_1 forward packs of the f16 + f8 plans (conv1_1_tc.cu), 16 KB
  float* bias[kNumConvs] = {};    // fp32 [cout]
  bf16* wf[kNumConvs] = {};       // forward pack [9][cout][cin] (conv index >= 1), high halves
  bf16* wf_lo[kNumConvs] = {};    // low halves bf16(w - bf16(w)) for the bf16x3 forward
  bf16* wd[kNumConvs] = {};       // dgrad pack   [9][cin][cout]
  bf16* wf16[kNumConvs] = {};     // f16 + f8 forward: fp16 pack [9][cout][cin] ...
  bf16* wcorr[kNumConvs] = {};    // ... and e4m3 correction rows [9][cout][cin/64][128 B] (conv_common.cuh, kF8*)
  float corr_scale[kNumConvs] = {};   // 1 / (512 * weight scale): factor of the layer's correction accumulator
  bf16* wft[kNumConvs] = {};      // tiled pre-swizzled forward pack (hi), see pack_weights_tiled_kernel
  bf16* wft_lo[kNumConvs] = {};
  bf16* wdt[kNumConvs] = {};      // tiled pre-swizzled dgrad pack
  ~Vgg();
};
int vgg_create(Vgg** out, const float* const* w, const float* const* b, int num_convs, const float* mean3,
               cudaStream_t stream);

struct BwdStep {
  enum Kind { kGramGrad, kPoolBwd, kDgrad, kConv1Bwd, kConv1BwdTensor } kind;
  ConvOp conv;             // kGramGrad, kDgrad
  const bf16* in = nullptr;
  const bf16* act = nullptr;
  bf16* out = nullptr;
  int H = 0, W = 0, C = 0;   // kPoolBwd: fine-resolution dims
  int inject_k = -1;         // loss layer whose external gradient (Plan::backward `inject`) this launch can add
};

struct LossLayer {
  int layer;
  float coef;
  int C, HW;
  GramOp gram;
  float* G = nullptr;         // [B,C,C]
  float* target = nullptr;    // [B,C,C]
  long long target_stride = 0;
  bool target_set = false;    // this layer's target has been written since the last stx_plan_clear_target
  bf16* dscaled = nullptr;    // [B,C,C]
  bf16* dfeat = nullptr;      // [B,h,w,C] total gradient dL/dF_k (only with STX_PLAN_EXPORT_LAYER_GRADS)
  int part_offset = 0;        // into loss_part rows
  int part_blocks = 0;
  ConvOp layer_grad_op;       // calc_grad path (utils.py:50-55): F * D * 4/(C^2 n) into scratch
};

// Optional per-launch timing (bench/roofline): CUDA events around every launch of a step, grouped by kernel class.
enum KernelClass { kClsConvFwd = 0, kClsConvDgrad, kClsGramGrad, kClsGram, kClsGramFinish, kClsConv1, kClsPool, kClsOther, kNumClasses };
struct Timeline {
  struct Rec { int cls; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  cudaEvent_t get();
  void reset() { recs.clear(); used = 0; }
  ~Timeline();
};

struct Plan {
  const Vgg* vgg = nullptr;
  Timeline* tl = nullptr;
  int B = 0, H = 0, W = 0, topmost = 15;
  int lh[kNumLayers], lw[kNumLayers], lc[kNumLayers];
  int export_grads = 0;            // keep dL/dF_k of every loss layer (PO stage preparation)
  int split = 1;                   // forward precision: 0 = plain bf16, 1 = bf16x3 (hi/lo bf16 activation planes),
                                   // 2 = f16 + f8 (fp16 planes + e4m3 correction planes: two MMAs' time per MAC)
  int f16_plain_layers = 0;        // f16+f8 plans: convolutions (bit = conv index) without the correction product
  int full_features = 0;           // keep the low half of every recorded layer (feature export), not only those the path reads
  bool lo_valid[kNumLayers] = {};  // act_lo[l] is written by the forward
  bool hi_valid[kNumLayers] = {};  // act[l] is written by the forward (false: the layer only leaves a one-bit ReLU mask)
  std::vector<uint32_t*> mask_bits;
  bf16* act[kNumLayers] = {};      // recorded layers (high halves; the only halves in plain mode)
  bf16* act_lo[kNumLayers] = {};   // low halves (split mode) / e4m3 correction planes (f16 + f8)
  bf16* act_bf[kNumLayers] = {};   // f16 + f8: bf16 copy of the LOSS layers (A operand of the Gram gradient)
  ConvOp fwd[kNumLayers];
  bool pool_fused[kNumLayers] = {};   // pooling layer computed in the epilogue of the convolution that feeds it
  std::vector<LossLayer> loss;
  std::vector<BwdStep> bwd;
  bf16* gbuf[2] = {nullptr, nullptr};
  float* loss_part = nullptr;      // [B][part_total]
  int part_total = 0;
  float* layer_loss = nullptr;     // [nloss][B]
  float* d_coef = nullptr;         // [nloss] coef/4
  int* d_active = nullptr;         // [B] per-image switch of the conv work loops (ConvKernelParams::active): all 1 except
                                   // while the device optimiser runs a batch whose jobs finish at different times
  int* d_part_offsets = nullptr;   // [nloss+1]
  bool have_target = false, have_grams = false;
  long long generation = 0;        // bumped by every forward(): a backward may check that its activations are still there
  // Gram + finish launches run on a side stream, forked right after their layer is produced and joined before the
  // loss reduction: they are bandwidth-bound and hide under the tensor-bound convolutions of the deeper layers.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork[kNumLayers] = {};
  cudaEvent_t ev_join = nullptr;
  // staging for the *_host entry points
  float* x_dev = nullptr;
  float* grad_dev = nullptr;
  float* func_dev = nullptr;
  float* pinned = nullptr;         // host pinned: x | grad | func
  void* arena = nullptr;
  size_t arena_bytes = 0;
  float* splitk_ws = nullptr;      // split-K workspace shared by the convolutions of this plan (small grids only)
  int* splitk_counters = nullptr;
  ~Plan();

  int forward(const float* image, int is_preimage, int with_grams, bool with_loss, cudaStream_t stream);
  // inject (nullable): per loss layer k, an external gradient [B,h,w,C] bf16 (or null) added to dL/dF_k before it is
  // back-propagated - the part of a caller's upstream gradient that is not of the form F_k * M (po/ double backward).
  int backward(const float* image, int is_preimage, float* grad, cudaStream_t stream,
               const bf16* const* inject = nullptr);
  // Backward of the last forward with caller-supplied matrices: dL/dF_k = F_k * M_k + E_k, M_k fp32 [B,C,C]
  // (device), E_k bf16 [B,h,w,C] or null; back-propagated through VGG to the image.
  int backward_custom(const float* image, int is_preimage, const float* const* M, const bf16* const* E, float* grad,
                      cudaStream_t stream);
  int step(const float* x, int is_preimage, float* func, float* grad, float* layer_loss_out, cudaStream_t stream);
  // One step with events around every launch; ms[kNumClasses] and count[kNumClasses] receive the device time and
  // the number of launches per kernel class. Synchronous.
  int step_timed(const float* x, int is_preimage, float* func, float* grad, float* ms, int* count,
                 cudaStream_t stream);
  size_t image_elems() const { return static_cast<size_t>(B) * H * W * 3; }
};
int plan_create(Plan** out, const Vgg* vgg, int B, int H, int W, int topmost, int nloss, const int* loss_layers,
                const float* coefs, int flags);

}  // namespace stx
