// extern "C" surface declared in include/syntex_b200.h. Thin: argument checks + calls into plan.cu / ops.
#include "../../include/syntex_b200.h"

#include "lbfgs.h"
#include "ops.h"
#include "plan.h"

using namespace stx;

namespace stx {
extern int g_debug_gram_lbo;
extern int g_debug_gram_sbo;
extern int g_debug_dgrad_variant;
extern int g_debug_no_gram_fusion;
extern int g_debug_no_poolbwd_fusion;
extern int g_debug_splitk;
extern int g_debug_tiled_weights;
extern int g_debug_resident;
extern int g_lbfgs_no_graph;
extern int g_debug_f8_single_cta;
extern int g_debug_conv1_1_cuda_cores;
extern int g_debug_pair;
extern int g_debug_no_mask_bits;
extern int g_lbfgs_no_skip;
extern int g_f16_plain_layers;
extern int g_lbfgs_serial_active;
}  // namespace stx

struct stx_vgg {
  Vgg* v;
};
struct stx_plan {
  Plan* p;
};
struct stx_lbfgs {
  Lbfgs* o;
};

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

const char* stx_last_error(void) { return last_error().c_str(); }
int stx_abi_version(void) { return 1; }
long long stx_launch_count(void) { return launch_count(); }

// CRC-32C (Castagnoli), slicing by 8 on the host: the checksum TensorFlow checkpoint files carry per tensor
// (syntex_b200/po/tf_bundle.py; a byte loop in Python takes about a second per megabyte).
unsigned int stx_crc32c(const void* data, size_t n, unsigned int crc) {
  static uint32_t tab[8][256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) tab[t][i] = (tab[t - 1][i] >> 8) ^ tab[0][tab[t - 1][i] & 0xFF];
    ready = true;
  }
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t c = crc ^ 0xFFFFFFFFu;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = tab[7][lo & 0xFF] ^ tab[6][(lo >> 8) & 0xFF] ^ tab[5][(lo >> 16) & 0xFF] ^ tab[4][lo >> 24] ^
        tab[3][hi & 0xFF] ^ tab[2][(hi >> 8) & 0xFF] ^ tab[1][(hi >> 16) & 0xFF] ^ tab[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

int stx_debug_set(int key, int value) {
  switch (key) {
    case 1: g_debug_gram_lbo = value; return STX_OK;
    case 2: g_debug_gram_sbo = value; return STX_OK;
    case 3: g_debug_dgrad_variant = value; return STX_OK;
    case 4: g_debug_no_gram_fusion = value; return STX_OK;
    case 5: g_debug_no_poolbwd_fusion = value; return STX_OK;
    case 6: g_debug_splitk = value; return STX_OK;
    case 7: g_debug_tiled_weights = value; return STX_OK;
    case 8: g_debug_resident = value; return STX_OK;
    case 9: g_pdl = value; return STX_OK;
    case 10: g_lbfgs_no_graph = value; return STX_OK;
    case 12: g_debug_f8_single_cta = value; return STX_OK;
    case 13: g_debug_conv1_1_cuda_cores = value; return STX_OK;
    case 14: g_lbfgs_serial_active = value; return STX_OK;
    case 15: g_debug_pair = value; return STX_OK;
    case 16: g_debug_no_mask_bits = value; return STX_OK;
    case 17: g_lbfgs_no_skip = value; return STX_OK;
    case 18: g_f16_plain_layers = value; return STX_OK;
    default: return fail(STX_ERR_INVALID, "stx_debug_set: unknown key");
  }
}

// ------------------------------------------------------------------------------------------------ weights
int stx_vgg_create(stx_vgg** out, const float* const* weights_hwio, const float* const* biases, int num_convs,
                   const float* mean3, void* stream) {
  STX_REQUIRE(out, "stx_vgg_create: null out");
  Vgg* v = nullptr;
  STX_TRY(vgg_create(&v, weights_hwio, biases, num_convs, mean3, S(stream)));
  *out = new stx_vgg{v};
  return STX_OK;
}
void stx_vgg_destroy(stx_vgg* vgg) {
  if (!vgg) return;
  delete vgg->v;
  delete vgg;
}

// ------------------------------------------------------------------------------------------------ plan
int stx_plan_create(stx_plan** out, const stx_vgg* vgg, int batch, int height, int width, int topmost,
                    int num_loss_layers, const int* loss_layers, const float* coefs, int flags) {
  STX_REQUIRE(out && vgg, "stx_plan_create: null argument");
  Plan* p = nullptr;
  STX_TRY(plan_create(&p, vgg->v, batch, height, width, topmost, num_loss_layers, loss_layers, coefs, flags));
  *out = new stx_plan{p};
  return STX_OK;
}
void stx_plan_destroy(stx_plan* plan) {
  if (!plan) return;
  delete plan->p;
  delete plan;
}
size_t stx_plan_device_bytes(const stx_plan* plan) { return plan ? plan->p->arena_bytes : 0; }

int stx_plan_forward(stx_plan* plan, const float* image, int is_preimage, int with_grams, void* stream) {
  STX_REQUIRE(plan, "stx_plan_forward: null plan");
  return plan->p->forward(image, is_preimage, with_grams, false, S(stream));
}

int stx_plan_layer_shape(const stx_plan* plan, int layer, int* h, int* w, int* c) {
  STX_REQUIRE(plan, "stx_plan_layer_shape: null plan");
  STX_REQUIRE(layer >= 0 && layer <= plan->p->topmost, "layer index out of range for this plan");
  if (h) *h = plan->p->lh[layer];
  if (w) *w = plan->p->lw[layer];
  if (c) *c = plan->p->lc[layer];
  return STX_OK;
}

int stx_plan_copy_feature_f32(stx_plan* plan, int layer, float* out, void* stream) {
  STX_REQUIRE(plan && out, "stx_plan_copy_feature_f32: null argument");
  Plan* p = plan->p;
  STX_REQUIRE(layer >= 0 && layer <= p->topmost, "layer index out of range for this plan");
  const size_t n = static_cast<size_t>(p->B) * p->lh[layer] * p->lw[layer] * p->lc[layer];
  if (!p->hi_valid[layer])
    return fail(STX_ERR_STATE, "stx_plan_copy_feature_f32: this plan keeps only the ReLU sign bits of that layer "
                               "(create it with STX_PLAN_FULL_FEATURES to record every layer)");
  if (p->split == 2)
    return f16f8_to_f32_launch(p->act[layer], p->lo_valid[layer] ? p->act_lo[layer] : nullptr, n, p->lc[layer], out,
                               S(stream));
  if (p->split && p->lo_valid[layer]) return bf16pair_to_f32_launch(p->act[layer], p->act_lo[layer], n, out, S(stream));
  return bf16_to_f32_launch(p->act[layer], n, out, S(stream));
}

int stx_plan_copy_gram(stx_plan* plan, int k, float* out, void* stream) {
  STX_REQUIRE(plan && out, "stx_plan_copy_gram: null argument");
  Plan* p = plan->p;
  STX_REQUIRE(k >= 0 && k < static_cast<int>(p->loss.size()), "loss-layer index out of range");
  if (!p->have_grams) return fail(STX_ERR_STATE, "stx_plan_copy_gram: no forward with Grams has run");
  const LossLayer& ll = p->loss[k];
  STX_CUDA(cudaMemcpyAsync(out, ll.G, static_cast<size_t>(p->B) * ll.C * ll.C * sizeof(float),
                           cudaMemcpyDeviceToDevice, S(stream)));
  return STX_OK;
}

int stx_plan_update_target_from_forward(stx_plan* plan, void* stream) {
  STX_REQUIRE(plan, "stx_plan_update_target_from_forward: null plan");
  Plan* p = plan->p;
  if (!p->have_grams) return fail(STX_ERR_STATE, "update_target: no forward with Grams has run");
  for (LossLayer& ll : p->loss) {
    STX_CUDA(cudaMemcpyAsync(ll.target, ll.G, static_cast<size_t>(p->B) * ll.C * ll.C * sizeof(float),
                             cudaMemcpyDeviceToDevice, S(stream)));
    ll.target_stride = static_cast<long long>(ll.C) * ll.C;
    ll.target_set = true;
  }
  p->have_target = true;
  return STX_OK;
}

int stx_plan_clear_target(stx_plan* plan) {
  STX_REQUIRE(plan, "stx_plan_clear_target: null plan");
  for (LossLayer& ll : plan->p->loss) ll.target_set = false;
  plan->p->have_target = false;
  return STX_OK;
}

long long stx_plan_generation(const stx_plan* plan) { return plan ? plan->p->generation : -1; }

int stx_plan_copy_target_gram(stx_plan* plan, int k, float* out, void* stream) {
  STX_REQUIRE(plan && out, "stx_plan_copy_target_gram: null argument");
  Plan* p = plan->p;
  STX_REQUIRE(k >= 0 && k < static_cast<int>(p->loss.size()), "loss-layer index out of range");
  if (!p->have_target) return fail(STX_ERR_STATE, "stx_plan_copy_target_gram: no target Grams set");
  const LossLayer& ll = p->loss[k];
  const size_t cc = static_cast<size_t>(ll.C) * ll.C;
  for (int b = 0; b < p->B; ++b)    // broadcast targets (stride 0) are expanded to [batch,C,C]
    STX_CUDA(cudaMemcpyAsync(out + b * cc, ll.target + static_cast<size_t>(b) * ll.target_stride, cc * sizeof(float),
                             cudaMemcpyDeviceToDevice, S(stream)));
  return STX_OK;
}

int stx_plan_set_target_gram(stx_plan* plan, int k, const float* gram, int per_image, void* stream) {
  STX_REQUIRE(plan && gram, "stx_plan_set_target_gram: null argument");
  Plan* p = plan->p;
  STX_REQUIRE(k >= 0 && k < static_cast<int>(p->loss.size()), "loss-layer index out of range");
  LossLayer& ll = p->loss[k];
  const size_t cc = static_cast<size_t>(ll.C) * ll.C;
  STX_CUDA(cudaMemcpyAsync(ll.target, gram, (per_image ? p->B : 1) * cc * sizeof(float), cudaMemcpyDeviceToDevice,
                           S(stream)));
  ll.target_stride = per_image ? static_cast<long long>(cc) : 0;
  // A target is complete only when EVERY loss layer has one: step / the optimiser / layer_grad refuse to run on a
  // partially set target instead of silently using zeros (or a previous target set) for the other layers.
  ll.target_set = true;
  bool all = true;
  for (const LossLayer& other : p->loss) all = all && other.target_set;
  p->have_target = all;
  return STX_OK;
}

int stx_plan_step(stx_plan* plan, const float* x, int is_preimage, float* func, float* grad, float* layer_loss,
                  void* stream) {
  STX_REQUIRE(plan, "stx_plan_step: null plan");
  return plan->p->step(x, is_preimage, func, grad, layer_loss, S(stream));
}

int stx_plan_step_timed(stx_plan* plan, const float* x, int is_preimage, float* func, float* grad, float* ms,
                        int* count, void* stream) {
  STX_REQUIRE(plan, "stx_plan_step_timed: null plan");
  return plan->p->step_timed(x, is_preimage, func, grad, ms, count, S(stream));
}

int stx_plan_step_host(stx_plan* plan, const float* x_host, int is_preimage, float* func_host, float* grad_host,
                       void* stream) {
  STX_REQUIRE(plan && x_host && func_host && grad_host, "stx_plan_step_host: null argument");
  Plan* p = plan->p;
  const size_t n = p->image_elems();
  float* px = p->pinned;
  float* pg = p->pinned + n;
  float* pf = p->pinned + 2 * n;
  memcpy(px, x_host, n * sizeof(float));
  STX_CUDA(cudaMemcpyAsync(p->x_dev, px, n * sizeof(float), cudaMemcpyHostToDevice, S(stream)));
  STX_TRY(p->step(p->x_dev, is_preimage, p->func_dev, p->grad_dev, nullptr, S(stream)));
  STX_CUDA(cudaMemcpyAsync(pg, p->grad_dev, n * sizeof(float), cudaMemcpyDeviceToHost, S(stream)));
  STX_CUDA(cudaMemcpyAsync(pf, p->func_dev, p->B * sizeof(float), cudaMemcpyDeviceToHost, S(stream)));
  STX_CUDA(cudaStreamSynchronize(S(stream)));
  memcpy(grad_host, pg, n * sizeof(float));
  memcpy(func_host, pf, p->B * sizeof(float));
  return STX_OK;
}

int stx_plan_copy_layer_total_grad(stx_plan* plan, int k, float* out, void* stream) {
  STX_REQUIRE(plan && out, "stx_plan_copy_layer_total_grad: null argument");
  Plan* p = plan->p;
  STX_REQUIRE(k >= 0 && k < static_cast<int>(p->loss.size()), "loss-layer index out of range");
  if (!p->export_grads) return fail(STX_ERR_STATE, "plan was not created with STX_PLAN_EXPORT_LAYER_GRADS");
  const LossLayer& ll = p->loss[k];
  return bf16_to_f32_launch(ll.dfeat, static_cast<size_t>(p->B) * ll.HW * ll.C, out, S(stream));
}

int stx_plan_layer_grad(stx_plan* plan, int k, float* out, void* stream) {
  STX_REQUIRE(plan && out, "stx_plan_layer_grad: null argument");
  Plan* p = plan->p;
  STX_REQUIRE(k >= 0 && k < static_cast<int>(p->loss.size()), "loss-layer index out of range");
  if (!p->have_grams || !p->have_target)
    return fail(STX_ERR_STATE, "stx_plan_layer_grad: needs a forward with Grams and target Grams");
  LossLayer& ll = p->loss[k];
  // d loss_layer / dF = 4/(C^2 n) * F (G - T)   (utils.py:46-48, :54): re-scale the difference, then one 1x1 GEMM.
  const float norm = static_cast<float>(ll.HW) + 1e-6f;
  const float dscale = 4.0f / (static_cast<float>(ll.C) * static_cast<float>(ll.C) * norm);
  STX_TRY(gram_finish_launch(ll.gram, 1e-6f, nullptr, ll.target, ll.target_stride, dscale, ll.dscaled, 0.0f, nullptr, 0,
                             S(stream)));
  STX_TRY(conv_launch(ll.layer_grad_op, S(stream)));
  const size_t n = static_cast<size_t>(p->B) * ll.HW * ll.C;
  return bf16_to_f32_launch(p->gbuf[0], n, out, S(stream));
}

int stx_plan_layer_grad_bf16(stx_plan* plan, int k, void* out_bf16, void* stream) {
  STX_REQUIRE(plan && out_bf16, "stx_plan_layer_grad_bf16: null argument");
  Plan* p = plan->p;
  STX_REQUIRE(k >= 0 && k < static_cast<int>(p->loss.size()), "loss-layer index out of range");
  if (!p->have_grams || !p->have_target)
    return fail(STX_ERR_STATE, "stx_plan_layer_grad_bf16: needs a forward with Grams and target Grams");
  LossLayer& ll = p->loss[k];
  const float norm = static_cast<float>(ll.HW) + 1e-6f;
  const float dscale = 4.0f / (static_cast<float>(ll.C) * static_cast<float>(ll.C) * norm);
  STX_TRY(gram_finish_launch(ll.gram, 1e-6f, nullptr, ll.target, ll.target_stride, dscale, ll.dscaled, 0.0f, nullptr, 0,
                             S(stream)));
  STX_TRY(conv_launch(ll.layer_grad_op, S(stream)));
  const size_t n = static_cast<size_t>(p->B) * ll.HW * ll.C;
  STX_CUDA(cudaMemcpyAsync(out_bf16, p->gbuf[0], n * sizeof(bf16), cudaMemcpyDeviceToDevice, S(stream)));
  return STX_OK;
}

int stx_vgg_weights(const stx_vgg* vgg, int conv_index, const void** fwd_hi, const void** fwd_lo,
                    const void** dgrad) {
  STX_REQUIRE(vgg, "stx_vgg_weights: null handle");
  STX_REQUIRE(conv_index >= 1 && conv_index < vgg->v->num_convs, "stx_vgg_weights: conv index out of range (1 = conv1_2)");
  if (fwd_hi) *fwd_hi = vgg->v->wf[conv_index];
  if (fwd_lo) *fwd_lo = vgg->v->wf_lo[conv_index];
  if (dgrad) *dgrad = vgg->v->wd[conv_index];
  return STX_OK;
}

int stx_plan_feature_bf16(stx_plan* plan, int layer, const void** hi, const void** lo) {
  STX_REQUIRE(plan && hi, "stx_plan_feature_bf16: null argument");
  Plan* p = plan->p;
  STX_REQUIRE(layer >= 0 && layer <= p->topmost, "stx_plan_feature_bf16: layer out of range");
  if (p->split == 2) return fail(STX_ERR_STATE, "stx_plan_feature_bf16: an f16+f8 plan keeps fp16 / e4m3 planes, not bf16");
  if (!p->hi_valid[layer])
    return fail(STX_ERR_STATE, "stx_plan_feature_bf16: this plan keeps only the ReLU sign bits of that layer");
  *hi = p->act[layer];
  if (lo) *lo = p->lo_valid[layer] ? p->act_lo[layer] : nullptr;
  return STX_OK;
}

int stx_plan_backward_custom(stx_plan* plan, const float* image, int is_preimage, const float* const* M,
                             const void* const* E, float* grad, void* stream) {
  STX_REQUIRE(plan, "stx_plan_backward_custom: null plan");
  return plan->p->backward_custom(image, is_preimage, M, reinterpret_cast<const bf16* const*>(E), grad, S(stream));
}

// ------------------------------------------------------------------------------------------------ L-BFGS
int stx_lbfgs_create(stx_lbfgs** out, stx_plan* plan, int maxcor, int bounded, float lower, float upper) {
  STX_REQUIRE(out && plan, "stx_lbfgs_create: null argument");
  Lbfgs* o = nullptr;
  STX_TRY(lbfgs_create(&o, plan->p, maxcor, bounded, lower, upper));
  *out = new stx_lbfgs{o};
  return STX_OK;
}
void stx_lbfgs_destroy(stx_lbfgs* opt) {
  if (!opt) return;
  delete opt->o;
  delete opt;
}
int stx_lbfgs_reset(stx_lbfgs* opt, const float* x0, int is_preimage, void* stream) {
  STX_REQUIRE(opt && x0, "stx_lbfgs_reset: null argument");
  return opt->o->reset(x0, is_preimage, S(stream));
}
int stx_lbfgs_run(stx_lbfgs* opt, int maxiter, int max_evals, void* stream) {
  STX_REQUIRE(opt, "stx_lbfgs_run: null optimizer");
  return opt->o->run(maxiter, max_evals, S(stream));
}
int stx_lbfgs_result(stx_lbfgs* opt, float* x, float* fun, int* nit, int* nfev, void* stream) {
  STX_REQUIRE(opt, "stx_lbfgs_result: null optimizer");
  return opt->o->result(x, fun, nit, nfev, S(stream));
}
int stx_lbfgs_phase_cycles(stx_lbfgs* opt, long long* host16) {
  STX_REQUIRE(opt && host16, "stx_lbfgs_phase_cycles: null argument");
  return opt->o->phase_cycles(host16);
}
int stx_lbfgs_minimize_host(stx_lbfgs* opt, const float* x0_host, int is_preimage, int maxiter, float* x_host,
                            float* fun_host, int* nit_host, int* nfev_host, void* stream) {
  STX_REQUIRE(opt && x0_host && x_host && fun_host, "stx_lbfgs_minimize_host: null argument");
  return opt->o->minimize_host(x0_host, is_preimage, maxiter, x_host, 0, fun_host, nit_host, nfev_host, S(stream));
}
int stx_lbfgs_minimize_host_f64(stx_lbfgs* opt, const float* x0_host, int is_preimage, int maxiter, double* x_host,
                                float* fun_host, int* nit_host, int* nfev_host, void* stream) {
  STX_REQUIRE(opt && x0_host && x_host && fun_host, "stx_lbfgs_minimize_host_f64: null argument");
  return opt->o->minimize_host(x0_host, is_preimage, maxiter, x_host, 1, fun_host, nit_host, nfev_host, S(stream));
}

// ------------------------------------------------------------------------------------------------ stand-alone ops
int stx_pack_conv_weights(const float* w_hwio, int cin, int cout, void* w_fwd_bf16, void* w_dgrad_bf16,
                          void* stream) {
  return pack_weights_launch(w_hwio, cin, cout, static_cast<bf16*>(w_fwd_bf16), nullptr,
                             static_cast<bf16*>(w_dgrad_bf16), S(stream));
}
int stx_pack_conv_weights_split(const float* w_hwio, int cin, int cout, void* w_fwd_hi, void* w_fwd_lo,
                                void* w_dgrad_bf16, void* stream) {
  return pack_weights_launch(w_hwio, cin, cout, static_cast<bf16*>(w_fwd_hi), static_cast<bf16*>(w_fwd_lo),
                             static_cast<bf16*>(w_dgrad_bf16), S(stream));
}

int stx_conv_bf16x3(const void* x_hi, const void* x_lo, int n, int H, int W, int cin, const void* w_hi,
                    const void* w_lo, int cout, const float* bias, int relu, void* out_hi, void* out_lo,
                    int force_bn, void* stream) {
  STX_REQUIRE(out_hi && out_lo, "stx_conv_bf16x3: null output");
  STX_REQUIRE(x_lo && w_lo, "stx_conv_bf16x3: null low-half operand");
  ConvOp op;
  STX_TRY(conv3x3_prepare(&op, static_cast<const bf16*>(x_hi), static_cast<const bf16*>(x_lo), n, H, W, cin,
                          static_cast<const bf16*>(w_hi), static_cast<const bf16*>(w_lo), cout, force_bn));
  op.bias = bias;
  op.relu = relu;
  op.out = static_cast<bf16*>(out_hi);
  op.out_lo = static_cast<bf16*>(out_lo);
  return conv_launch(op, S(stream));
}

int stx_conv_bf16(const void* x, int n, int H, int W, int cin, const void* w_packed, int cout, int taps,
                  const float* bias, const void* addend, const void* mask, int relu, void* out, int force_bn,
                  void* stream) {
  STX_REQUIRE(out, "stx_conv_bf16: null output");
  ConvOp op;
  if (taps == 9) {
    STX_TRY(conv3x3_prepare(&op, static_cast<const bf16*>(x), nullptr, n, H, W, cin,
                            static_cast<const bf16*>(w_packed), nullptr, cout, force_bn));
  } else {
    STX_TRY(conv_prepare(&op, static_cast<const bf16*>(x), n, H, W, cin, static_cast<const bf16*>(w_packed), cout,
                         taps, 0, force_bn % 1000));
  }
  op.bias = bias;
  op.addend = static_cast<const bf16*>(addend);
  op.mask = static_cast<const bf16*>(mask);
  op.relu = relu;
  op.out = static_cast<bf16*>(out);
  return conv_launch(op, S(stream));
}

int stx_conv1x1_per_image_bf16(const void* x, int n, int H, int W, int cin, const void* w, int cout, void* out,
                               void* stream) {
  STX_REQUIRE(x && w && out, "stx_conv1x1_per_image_bf16: null argument");
  ConvOp op;
  STX_TRY(conv_prepare(&op, static_cast<const bf16*>(x), n, H, W, cin, static_cast<const bf16*>(w), cout, 1, 1, 0));
  op.out = static_cast<bf16*>(out);
  return conv_launch(op, S(stream));
}

int stx_conv_bf16_pad(const void* x, int n, int H, int W, int cin, const void* w_packed, int cout, int pad,
                      const float* bias, const void* addend, int relu, void* out, int variant, void* stream) {
  STX_REQUIRE(out, "stx_conv_bf16_pad: null output");
  ConvOp op;
  // geometry from the OUTPUT size; the activation map is then re-pointed at the (H + 2 - 2 pad)-sized input
  STX_TRY(conv3x3_prepare(&op, static_cast<const bf16*>(x), nullptr, n, H, W, cin, static_cast<const bf16*>(w_packed),
                          nullptr, cout, variant));
  if (pad != 1) STX_TRY(conv_set_input_pad(&op, static_cast<const bf16*>(x), pad));
  op.bias = bias;
  op.addend = static_cast<const bf16*>(addend);
  op.relu = relu;
  op.out = static_cast<bf16*>(out);
  return conv_launch(op, S(stream));
}
size_t stx_conv_wgrad_workspace_bytes(int n, int H, int W, int cin, int cout) {
  return wgrad_workspace_bytes(n, H, W, cin, cout);
}
int stx_conv_wgrad_bf16(const void* x, const void* dy, int n, int H, int W, int cin, int cout, int pad,
                        float* dw_hwio, void* workspace, void* stream) {
  return wgrad_launch(static_cast<const bf16*>(x), static_cast<const bf16*>(dy), n, H, W, cin, cout, pad, dw_hwio,
                      static_cast<float*>(workspace), S(stream));
}
int stx_conv_wgrad_oihw_bf16(const void* x, const void* dy, int n, int H, int W, int cin, int cout, int pad,
                             int cout_real, float* dw_oihw, void* workspace, void* stream) {
  return wgrad_oihw_launch(static_cast<const bf16*>(x), static_cast<const bf16*>(dy), n, H, W, cin, cout, pad, cout_real,
                           dw_oihw, static_cast<float*>(workspace), S(stream));
}
int stx_pack_conv_weights_oihw(const float* w_oihw, int cin, int cout, int cout_padded, void* w_fwd_bf16,
                               void* w_dgrad_bf16, void* stream) {
  return pack_weights_oihw_launch(w_oihw, cin, cout, cout_padded, static_cast<bf16*>(w_fwd_bf16),
                                  static_cast<bf16*>(w_dgrad_bf16), S(stream));
}
int stx_pad_replicate_bf16(const void* x, int n, int H, int W, int C, int relu, void* out, void* stream) {
  return pad_rep_launch(static_cast<const bf16*>(x), n, H, W, C, 0, relu, 0.0f, static_cast<bf16*>(out), S(stream));
}
int stx_pad_replicate_bwd_bf16(const void* gp, const void* x_or_null, int n, int H, int W, int C, void* dx,
                               void* stream) {
  return pad_rep_bwd_launch(static_cast<const bf16*>(gp), static_cast<const bf16*>(x_or_null), n, H, W, C, 0, 0.0f,
                            static_cast<bf16*>(dx), S(stream));
}
int stx_pad2d_bf16(const void* x, int n, int H, int W, int C, int mode, int act, float slope, void* out, void* stream) {
  return pad_rep_launch(static_cast<const bf16*>(x), n, H, W, C, mode, act, slope, static_cast<bf16*>(out), S(stream));
}
int stx_pad2d_bwd_bf16(const void* gp, const void* x_or_null, int n, int H, int W, int C, int mode, float slope,
                       void* dx, void* stream) {
  return pad_rep_bwd_launch(static_cast<const bf16*>(gp), static_cast<const bf16*>(x_or_null), n, H, W, C, mode, slope,
                            static_cast<bf16*>(dx), S(stream));
}
size_t stx_inorm_workspace_bytes(int n, int HW, int C) { return inorm_workspace_bytes(n, HW, C); }
int stx_inorm_stats_bf16(const void* x, int n, int HW, int C, float eps, float* mean, float* rstd, void* workspace,
                         void* stream) {
  return inorm_stats_launch(static_cast<const bf16*>(x), n, HW, C, eps, mean, rstd, static_cast<float*>(workspace),
                            S(stream));
}
int stx_inorm_apply_bf16(const void* x, int n, int HW, int C, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, int act, float slope, const void* addend, void* out,
                         void* stream) {
  return inorm_apply_launch(static_cast<const bf16*>(x), n, HW, C, mean, rstd, gamma, beta, act, slope,
                            static_cast<const bf16*>(addend), static_cast<bf16*>(out), S(stream));
}
int stx_inorm_bwd_bf16(const void* x, const void* dy, int n, int HW, int C, const float* mean, const float* rstd,
                       const float* gamma, const float* beta, int act, float slope, float* s1, float* s2, void* dx,
                       void* workspace, void* stream) {
  return inorm_bwd_launch(static_cast<const bf16*>(x), static_cast<const bf16*>(dy), n, HW, C, mean, rstd, gamma, beta,
                          act, slope, s1, s2, static_cast<bf16*>(dx), static_cast<float*>(workspace), S(stream));
}

int stx_adam_tf(float* param, const float* grad, float* m, float* v, size_t count, const float* lr_dev,
                const float* step_dev, float beta1, float beta2, float eps, void* stream) {
  return adam_tf_launch(param, grad, m, v, count, lr_dev, step_dev, beta1, beta2, eps, S(stream));
}

int stx_conv1_1_fwd(const float* image, int n, int H, int W, int is_preimage, const float* mean3_host,
                    const float* w27x64, const float* b64, void* out_bf16, void* out_lo, void* stream) {
  return conv1_1_fwd_launch(image, n, H, W, is_preimage, mean3_host, w27x64, b64, static_cast<bf16*>(out_bf16),
                            static_cast<bf16*>(out_lo), S(stream));
}
int stx_conv1_1_bwd(const void* g_bf16, const float* image, int n, int H, int W, int is_preimage,
                    const float* w27x64, float* grad_out, void* stream) {
  return conv1_1_bwd_launch(static_cast<const bf16*>(g_bf16), image, n, H, W, is_preimage, w27x64, grad_out,
                            S(stream));
}
int stx_avgpool2_fwd(const void* x_bf16, const void* x_lo, int n, int H, int W, int C, void* out_bf16,
                     void* out_lo, void* stream) {
  return avgpool2_fwd_launch(static_cast<const bf16*>(x_bf16), static_cast<const bf16*>(x_lo), n, H, W, C,
                             static_cast<bf16*>(out_bf16), static_cast<bf16*>(out_lo), S(stream));
}
int stx_avgpool2_bwd(const void* g_bf16, const void* act, int n, int H, int W, int C, void* out_bf16, void* stream) {
  return avgpool2_bwd_mask_launch(static_cast<const bf16*>(g_bf16), static_cast<const bf16*>(act), n, H, W, C,
                                  static_cast<bf16*>(out_bf16), S(stream));
}

size_t stx_gram_workspace_bytes(int n, int HW, int C) {
  if (n <= 0 || HW <= 0 || C <= 0) return 0;
  return gram_partial_bytes(n, HW, C);
}
size_t stx_gram_cross_workspace_bytes(int n, int HW, int C) {
  if (n <= 0 || HW <= 0 || C <= 0) return 0;
  return gram_cross_partial_bytes(n, HW, C);
}
int stx_gram_cross_bf16(const void* a_bf16, const void* b_bf16, int n, int HW, int C, float eps, float* out,
                        void* workspace, void* stream) {
  STX_REQUIRE(out, "stx_gram_cross_bf16: null output");
  GramOp op;
  STX_TRY(gram_cross_prepare(&op, static_cast<const bf16*>(a_bf16), static_cast<const bf16*>(b_bf16), n, HW, C,
                             static_cast<float*>(workspace)));
  STX_TRY(gram_launch(op, S(stream)));
  return gram_finish_launch(op, eps, out, nullptr, 0, 0.0f, nullptr, 0.0f, nullptr, 0, S(stream));
}

int stx_gram_bf16(const void* feat_bf16, int n, int HW, int C, float eps, float* G_out, void* workspace,
                  void* stream) {
  STX_REQUIRE(G_out, "stx_gram_bf16: null output");
  GramOp op;
  STX_TRY(gram_prepare(&op, static_cast<const bf16*>(feat_bf16), n, HW, C, static_cast<float*>(workspace)));
  STX_TRY(gram_launch(op, S(stream)));
  return gram_finish_launch(op, eps, G_out, nullptr, 0, 0.0f, nullptr, 0.0f, nullptr, 0, S(stream));
}

int stx_f32_to_bf16(const float* x, size_t count, void* out_bf16, void* stream) {
  STX_REQUIRE(count == 0 || (x && out_bf16), "stx_f32_to_bf16: null argument");
  return f32_to_bf16_launch(x, count, static_cast<bf16*>(out_bf16), S(stream));
}
int stx_bf16_to_f32(const void* x_bf16, size_t count, float* out, void* stream) {
  STX_REQUIRE(count == 0 || (x_bf16 && out), "stx_bf16_to_f32: null argument");
  return bf16_to_f32_launch(static_cast<const bf16*>(x_bf16), count, out, S(stream));
}

}  // extern "C"
