#include "plan.h"

#include "ptx.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace stx {

// bring-up switches (stx_debug_set keys 3..5)
int g_debug_dgrad_variant = 0;      // variant code for the dgrad convolutions (0 = automatic)
int g_debug_no_gram_fusion = 0;     // 1 = stand-alone Gram-gradient launches
int g_debug_no_poolbwd_fusion = 0;  // 1 = stand-alone AvgPoolGrad launches
int g_f16_plain_layers = -1;        // f16+f8 plans: bit i = convolution i (1 = conv1_2 ... 11 = conv4_4) runs WITHOUT the e4m3
                                    // correction product: fp16 operands, one MMA per K step; -1 = as the plan's flags
                                    // say (STX_PLAN_F16F8_LEAN: conv4_2..conv4_4) (stx_debug_set key 18)
int g_debug_no_mask_bits = 0;       // 1 = the convolutions in front of a pool keep their full planes as ReLU masks (key 16)
int g_debug_tiled_weights = 0;      // 1 = fetch weight tiles with 1-D bulk copies from tiled pre-swizzled packs
                                    // (measured: no gain over tensor-map fetches, costs a second weight copy: off)
int g_debug_conv1_1_cuda_cores = 0;  // 1 = CUDA-core conv1_1 forward in the f16 + f8 plans too (stx_debug_set key 13)
int g_debug_pair = 3;               // bit 0: CTA pairs (tcgen05 cta_group::2) for the f16+f8 / plain-bf16 forward layers,
                                    // bit 1: for the input-gradient layers (stx_debug_set key 15)
int g_debug_f8_single_cta = 0;      // 1 = one CTA per SM for the 64-channel f16+f8 forward (stx_debug_set key 12)
int g_debug_splitk = 0;             // 1 = split-K work items for small grids (measured slower at 256^2: off)

static const LayerSpec kLayers[kNumLayers] = {
    {true, 0, 3, 64},     {true, 1, 64, 64},    {false, -1, 64, 64},   {true, 2, 64, 128},
    {true, 3, 128, 128},  {false, -1, 128, 128}, {true, 4, 128, 256},  {true, 5, 256, 256},
    {true, 6, 256, 256},  {true, 7, 256, 256},  {false, -1, 256, 256}, {true, 8, 256, 512},
    {true, 9, 512, 512},  {true, 10, 512, 512}, {true, 11, 512, 512},  {false, -1, 512, 512}};

const LayerSpec& layer_spec(int layer) { return kLayers[layer]; }

// ------------------------------------------------------------------------------------------------ weights
Vgg::~Vgg() {
  cudaFree(w0);
  cudaFree(wd0);
  cudaFree(wtc0);
  for (int i = 0; i < kNumConvs; ++i) {
    cudaFree(bias[i]);
    cudaFree(wf[i]);
    cudaFree(wf_lo[i]);
    cudaFree(wf16[i]);
    cudaFree(wcorr[i]);
    cudaFree(wd[i]);
    cudaFree(wft[i]);
    cudaFree(wft_lo[i]);
    cudaFree(wdt[i]);
  }
}

int vgg_create(Vgg** out, const float* const* w, const float* const* b, int num_convs, const float* mean3,
               cudaStream_t stream) {
  STX_REQUIRE(out && w && b && mean3, "vgg_create: null argument");
  STX_REQUIRE(num_convs >= 1 && num_convs <= kNumConvs, "vgg_create: num_convs must be in [1,12]");
  Vgg* v = new Vgg();
  v->num_convs = num_convs;
  for (int c = 0; c < 3; ++c) v->mean[c] = mean3[c];
  int ci = 0;
  int rc = STX_OK;
  for (int l = 0; l < kNumLayers && ci < num_convs && rc == STX_OK; ++l) {
    const LayerSpec& s = kLayers[l];
    if (!s.is_conv) continue;
    if (!w[ci] || !b[ci]) {
      rc = fail(STX_ERR_INVALID, "vgg_create: null weight or bias pointer");
      break;
    }
    const size_t wn = 9ull * s.cin * s.cout;
    auto chk = [&](cudaError_t e, const char* what) {
      if (e != cudaSuccess && rc == STX_OK) rc = cuda_fail(e, what);
    };
    chk(cudaMalloc(&v->bias[ci], s.cout * sizeof(float)), "cudaMalloc bias");
    if (rc == STX_OK) chk(cudaMemcpyAsync(v->bias[ci], b[ci], s.cout * sizeof(float), cudaMemcpyHostToDevice, stream), "copy bias");
    if (ci == 0) {
      chk(cudaMalloc(&v->w0, wn * sizeof(float)), "cudaMalloc w0");
      if (rc == STX_OK) chk(cudaMemcpyAsync(v->w0, w[ci], wn * sizeof(float), cudaMemcpyHostToDevice, stream), "copy w0");
      chk(cudaMalloc(&v->wd0, 9 * 16 * 64 * sizeof(bf16)), "cudaMalloc wd0");
      if (rc == STX_OK) rc = pack_conv1_1_dgrad_launch(v->w0, v->wd0, stream);
      {
        float wmax = 0.0f;
        for (size_t i = 0; i < wn; ++i) wmax = std::max(wmax, std::fabs(w[ci][i]));
        int ex = 0;
        if (wmax > 0.0f) ex = static_cast<int>(std::floor(std::log2(224.0f / wmax)));
        ex = std::max(-20, std::min(20, ex));
        const float sw = std::ldexp(1.0f, ex);
        v->corr_scale[0] = 1.0f / (512.0f * sw);
        chk(cudaMalloc(&v->wtc0, 16384), "cudaMalloc wtc0");
        if (rc == STX_OK) rc = pack_conv1_1_tc_launch(v->w0, sw, v->wtc0, stream);
      }
    } else {
      float* tmp = nullptr;
      chk(cudaMalloc(&tmp, wn * sizeof(float)), "cudaMalloc staging");
      chk(cudaMalloc(&v->wf[ci], wn * sizeof(bf16)), "cudaMalloc wf");
      chk(cudaMalloc(&v->wf_lo[ci], wn * sizeof(bf16)), "cudaMalloc wf_lo");
      chk(cudaMalloc(&v->wd[ci], wn * sizeof(bf16)), "cudaMalloc wd");
      if (rc == STX_OK) chk(cudaMemcpyAsync(tmp, w[ci], wn * sizeof(float), cudaMemcpyHostToDevice, stream), "copy w");
      if (rc == STX_OK) rc = pack_weights_launch(tmp, s.cin, s.cout, v->wf[ci], v->wf_lo[ci], v->wd[ci], stream);
      {
        // f16 + f8 packs: power-of-two weight scale from the layer's largest magnitude (e4m3 tops out at 448)
        float wmax = 0.0f;
        for (size_t i = 0; i < wn; ++i) wmax = std::max(wmax, std::fabs(w[ci][i]));
        int ex = 0;
        if (wmax > 0.0f) ex = static_cast<int>(std::floor(std::log2(224.0f / wmax)));
        ex = std::max(-20, std::min(20, ex));
        const float sw = std::ldexp(1.0f, ex);
        v->corr_scale[ci] = 1.0f / (512.0f * sw);
        chk(cudaMalloc(&v->wf16[ci], wn * sizeof(bf16)), "cudaMalloc wf16");
        chk(cudaMalloc(&v->wcorr[ci], wn * sizeof(bf16)), "cudaMalloc wcorr");
        if (rc == STX_OK) rc = pack_weights_f16f8_launch(tmp, s.cin, s.cout, sw, v->wf16[ci], v->wcorr[ci], stream);
      }
      if (g_debug_tiled_weights) {
        chk(cudaMalloc(&v->wft[ci], wn * sizeof(bf16)), "cudaMalloc wft");
        chk(cudaMalloc(&v->wft_lo[ci], wn * sizeof(bf16)), "cudaMalloc wft_lo");
        chk(cudaMalloc(&v->wdt[ci], wn * sizeof(bf16)), "cudaMalloc wdt");
        if (rc == STX_OK)
          rc = pack_weights_tiled_launch(tmp, s.cin, s.cout, v->wft[ci], v->wft_lo[ci], v->wdt[ci], stream);
      }
      if (rc == STX_OK) chk(cudaStreamSynchronize(stream), "sync after pack");
      cudaFree(tmp);
    }
    ++ci;
  }
  if (rc == STX_OK) {
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "vgg_create sync");
  }
  if (rc != STX_OK) {
    delete v;
    return rc;
  }
  *out = v;
  return STX_OK;
}

// ------------------------------------------------------------------------------------------------ plan
cudaEvent_t Timeline::get() {
  if (used == pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    pool.push_back(e);
  }
  return pool[used++];
}
Timeline::~Timeline() {
  for (cudaEvent_t e : pool) cudaEventDestroy(e);
}

Plan::~Plan() {
  cudaFree(arena);
  cudaFree(splitk_ws);
  for (uint32_t* b : mask_bits) cudaFree(b);
  cudaFree(splitk_counters);
  if (pinned) cudaFreeHost(pinned);
  if (side) cudaStreamDestroy(side);
  for (cudaEvent_t e : ev_fork)
    if (e) cudaEventDestroy(e);
  if (ev_join) cudaEventDestroy(ev_join);
}

// Launch `call`, bracketing it with events when a timeline is attached.
#define STX_LAUNCH(cls, call)                               \
  do {                                                      \
    if (tl) {                                               \
      Timeline::Rec _r{cls, tl->get(), tl->get()};          \
      cudaEventRecord(_r.a, stream);                        \
      STX_TRY(call);                                        \
      cudaEventRecord(_r.b, stream);                        \
      tl->recs.push_back(_r);                               \
    } else {                                                \
      STX_TRY(call);                                        \
    }                                                       \
  } while (0)

namespace {

// func[img] = sum_k coef_k/4 * L_k,  layer_loss[k][img] = L_k = sum of that layer's block partials (fixed order).
__global__ void __launch_bounds__(128)
loss_reduce_kernel(const float* __restrict__ part, int part_total, const int* __restrict__ offsets, int nloss,
                   const float* __restrict__ coef4, int nimg, float* __restrict__ layer_loss, float* __restrict__ func) {
  pdl_launch_dependents();   // programmatic dependent launch: see launch_pdl (common.h)
  pdl_wait();
  const int img = blockIdx.x;
  const float* p = part + static_cast<size_t>(img) * part_total;
  __shared__ float red[4];
  float total = 0.0f;
  for (int k = 0; k < nloss; ++k) {
    float s = 0.0f;
    for (int i = offsets[k] + threadIdx.x; i < offsets[k + 1]; i += 128) s += p[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    const float lk = (red[0] + red[1]) + (red[2] + red[3]);
    __syncthreads();
    if (threadIdx.x == 0 && layer_loss) layer_loss[k * nimg + img] = lk;
    total += coef4[k] * lk;
  }
  if (threadIdx.x == 0 && func) func[img] = total;
}

struct Bump {
  size_t off = 0;
  char* base = nullptr;
  template <typename T>
  T* take(size_t count) {
    off = (off + 1023) & ~size_t(1023);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

}  // namespace

static int plan_layout(Plan* p, Bump& a) {
  const int B = p->B;
  for (int l = 0; l <= p->topmost; ++l) {
    const size_t n = static_cast<size_t>(B) * p->lh[l] * p->lw[l] * p->lc[l];
    p->act[l] = a.take<bf16>(n);
    p->act_lo[l] = p->split ? a.take<bf16>(n) : nullptr;
    bool is_loss = false;
    for (const LossLayer& ll : p->loss) is_loss = is_loss || ll.layer == l;
    p->act_bf[l] = (p->split == 2 && is_loss) ? a.take<bf16>(n) : nullptr;
  }
  size_t gmax = 0;
  for (int l = 0; l <= p->topmost; ++l) gmax = std::max(gmax, static_cast<size_t>(B) * p->lh[l] * p->lw[l] * p->lc[l]);
  p->gbuf[0] = a.take<bf16>(gmax);
  p->gbuf[1] = a.take<bf16>(gmax);
  int part = 0;
  for (LossLayer& ll : p->loss) {
    const size_t cc = static_cast<size_t>(B) * ll.C * ll.C;
    ll.G = a.take<float>(cc);
    ll.target = a.take<float>(cc);
    ll.dscaled = a.take<bf16>(cc);
    ll.dfeat = p->export_grads ? a.take<bf16>(static_cast<size_t>(B) * ll.HW * ll.C) : nullptr;
    float* partial = a.take<float>(gram_partial_bytes(B, ll.HW, ll.C) / sizeof(float));
    ll.gram.partial = partial;
    ll.part_offset = part;
    ll.part_blocks = gram_finish_blocks(ll.C);
    part += ll.part_blocks;
  }
  p->part_total = part;
  p->loss_part = a.take<float>(static_cast<size_t>(B) * std::max(part, 1));
  p->layer_loss = a.take<float>(static_cast<size_t>(B) * std::max<size_t>(p->loss.size(), 1));
  p->d_coef = a.take<float>(std::max<size_t>(p->loss.size(), 1));
  p->d_part_offsets = a.take<int>(p->loss.size() + 1);
  p->d_active = a.take<int>(B);
  p->x_dev = a.take<float>(p->image_elems());
  p->grad_dev = a.take<float>(p->image_elems());
  p->func_dev = a.take<float>(B);
  return STX_OK;
}

int plan_create(Plan** out, const Vgg* vgg, int B, int H, int W, int topmost, int nloss, const int* loss_layers,
                const float* coefs, int flags) {
  STX_REQUIRE(out && vgg, "plan_create: null argument");
  STX_REQUIRE(B >= 1 && H >= 1 && W >= 1, "plan_create: batch/height/width must be positive");
  STX_REQUIRE(topmost >= 0 && topmost < kNumLayers, "plan_create: topmost out of range");
  STX_REQUIRE(nloss >= 0 && nloss <= kNumLayers, "plan_create: bad number of loss layers");
  STX_REQUIRE(nloss == 0 || (loss_layers && coefs), "plan_create: loss layers without indices/coefs");
  Plan* p = new Plan();
  p->vgg = vgg;
  p->B = B;
  p->H = H;
  p->W = W;
  p->topmost = topmost;
  p->split = (flags & STX_PLAN_FAST_BF16) ? 0 : ((flags & STX_PLAN_F16F8) ? 2 : 1);
  p->export_grads = (flags & STX_PLAN_EXPORT_LAYER_GRADS) ? 1 : 0;
  p->full_features = (flags & STX_PLAN_FULL_FEATURES) ? 1 : 0;
  p->f16_plain_layers = g_f16_plain_layers >= 0 ? g_f16_plain_layers : ((flags & STX_PLAN_F16F8_LEAN) ? 0xE00 : 0);
  for (int l = 0; l < kNumLayers; ++l) p->lo_valid[l] = p->split != 0;
  for (int l = 0; l < kNumLayers; ++l) p->hi_valid[l] = true;
  int h = H, w = W, convs_needed = 0;
  for (int l = 0; l <= topmost; ++l) {
    const LayerSpec& s = kLayers[l];
    if (!s.is_conv) {
      if ((h & 1) || (w & 1)) {
        delete p;
        return fail(STX_ERR_UNSUPPORTED, "plan_create: odd feature-map size at a pooling layer (image side must be a multiple of 16 for pool4)");
      }
      h >>= 1;
      w >>= 1;
    } else {
      convs_needed = s.conv_index + 1;
    }
    p->lh[l] = h;
    p->lw[l] = w;
    p->lc[l] = s.cout;
  }
  if (convs_needed > vgg->num_convs) {
    delete p;
    return fail(STX_ERR_INVALID, "plan_create: weights do not cover the requested topmost layer");
  }
  for (int k = 0; k < nloss; ++k) {
    const int l = loss_layers[k];
    if (l < 0 || l > topmost) {
      delete p;
      return fail(STX_ERR_INVALID, "plan_create: loss layer above topmost");
    }
    for (int j = 0; j < k; ++j)
      if (loss_layers[j] == l) {
        delete p;
        return fail(STX_ERR_INVALID, "plan_create: duplicate loss layer");
      }
    LossLayer ll;
    ll.layer = l;
    ll.coef = coefs[k];
    ll.C = p->lc[l];
    ll.HW = p->lh[l] * p->lw[l];
    p->loss.push_back(ll);
  }

  // Pass 1: size the arena; pass 2: carve it.
  Bump sizing;
  plan_layout(p, sizing);
  p->arena_bytes = sizing.off + 4096;
  cudaError_t e = cudaMalloc(&p->arena, p->arena_bytes);
  if (e != cudaSuccess) {
    delete p;
    return cuda_fail(e, "cudaMalloc(plan arena)");
  }
  e = cudaMemset(p->arena, 0, p->arena_bytes);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking);
  for (int l = 0; l < kNumLayers && e == cudaSuccess; ++l) e = cudaEventCreateWithFlags(&p->ev_fork[l], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMallocHost(&p->pinned, (2 * p->image_elems() + B) * sizeof(float));
  if (e != cudaSuccess) {
    delete p;
    return cuda_fail(e, "plan staging allocation");
  }
  Bump carve;
  carve.base = static_cast<char*>(p->arena);
  plan_layout(p, carve);

  int rc = STX_OK;
  // Forward ops.
  for (int l = 1; l <= topmost && rc == STX_OK; ++l) {
    const LayerSpec& s = kLayers[l];
    if (!s.is_conv) continue;
    ConvOp& op = p->fwd[l];
    if (p->split == 2 && ((p->f16_plain_layers >> s.conv_index) & 1)) {
      // f16 + f8 plan, a layer chosen to go without the correction product: fp16 planes against the fp16 pack, one
      // MMA per K step (the single-MMA instances: CTA pairs, N tile 256), outputs in the plan's plane format
      rc = conv3x3_prepare(&op, p->act[l - 1], nullptr, B, p->lh[l], p->lw[l], s.cin, vgg->wf16[s.conv_index], nullptr,
                           s.cout, (g_debug_pair & 1) ? 1000000 : 0);
      if (rc == STX_OK && op.patch && op.mt != 1)      // no stacked-tile instance of this flavour
        rc = conv3x3_prepare(&op, p->act[l - 1], nullptr, B, p->lh[l], p->lw[l], s.cin, vgg->wf16[s.conv_index],
                             nullptr, s.cout, 1000 + (s.cout % 128 == 0 ? 128 : 64));
      if (rc == STX_OK && !op.patch)
        rc = fail(STX_ERR_UNSUPPORTED, "plan_create: the f16+f8 forward needs feature maps at least 8 pixels wide");
      op.split = 3;
    } else if (p->split == 2) {
      // f16 + f8: fp16 planes against the fp16 pack, correction planes against the e4m3 rows (halo-patch kernel only)
      rc = conv3x3_prepare(&op, p->act[l - 1], p->act_lo[l - 1], B, p->lh[l], p->lw[l], s.cin, vgg->wf16[s.conv_index],
                           vgg->wcorr[s.conv_index], s.cout, (g_debug_pair & 1) ? 1000000 : 0);
      if (rc == STX_OK && !op.patch)
        rc = fail(STX_ERR_UNSUPPORTED, "plan_create: the f16+f8 forward needs feature maps at least 8 pixels wide");
      op.split = 2;
      op.corr_scale = vgg->corr_scale[s.conv_index];
      // 64-channel tiles are epilogue-bound (tensor pipe 27 % busy, profiles/r02_ncu_full_fwd_shallow.md): two CTAs per
      // SM double the epilogue warps; each keeps a single hi+corr patch stage
      if (op.patch && op.bn == 64 && op.tiles_w * op.tiles_h * op.nimg >= 2 * 148 && !g_debug_f8_single_cta) op.occ = 2;
    } else {
      rc = conv3x3_prepare(&op, p->act[l - 1], p->split ? p->act_lo[l - 1] : nullptr, B, p->lh[l], p->lw[l], s.cin,
                           vgg->wf[s.conv_index], p->split ? vgg->wf_lo[s.conv_index] : nullptr, s.cout,
                           (!p->split && (g_debug_pair & 1)) ? 1000000 : 0);
    }
    op.w_tiled = g_debug_tiled_weights ? vgg->wft[s.conv_index] : nullptr;
    op.w_tiled_lo = g_debug_tiled_weights ? vgg->wft_lo[s.conv_index] : nullptr;
    if (op.split == 3) op.w_tiled = op.w_tiled_lo = nullptr;     // (the tiled packs are bf16)
    op.bias = vgg->bias[s.conv_index];
    op.relu = 1;
    op.out = p->act[l];
    op.out_lo = p->act_lo[l];
    op.aux_out = p->act_bf[l];
    if (rc == STX_OK && op.patch && l + 1 <= topmost && !kLayers[l + 1].is_conv) {
      op.pool_out = p->act[l + 1];            // 2x2 average pool fused into this convolution's epilogue
      op.pool_out_lo = p->act_lo[l + 1];
      op.pool_aux = p->act_bf[l + 1];
      p->pool_fused[l + 1] = true;
      if (!p->full_features) {
        // nothing on the path reads the low half of a convolution that only feeds a fused pool: the pool's own
        // hi/lo outputs carry the precision forward, the backward mask and the Gram use the high plane
        op.out_lo = nullptr;
        p->lo_valid[l] = false;
      }
    }
  }
  // Loss-layer ops.
  std::vector<float> coef4;
  std::vector<int> offs;
  for (size_t k = 0; k < p->loss.size() && rc == STX_OK; ++k) {
    LossLayer& ll = p->loss[k];
    float* partial = ll.gram.partial;
    rc = gram_prepare(&ll.gram, p->act[ll.layer], B, ll.HW, ll.C, partial);
    ll.gram.f16 = p->split == 2 ? 1 : 0;
    coef4.push_back(ll.coef * 0.25f);
    offs.push_back(ll.part_offset);
  }
  offs.push_back(p->part_total);
  if (rc == STX_OK && !p->loss.empty()) {
    e = cudaMemcpy(p->d_coef, coef4.data(), coef4.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_part_offsets, offs.data(), offs.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = cuda_fail(e, "plan constants upload");
  }

  // Backward program (reverse-mode over the forward chain; see DESIGN.md "backward program").
  // the features of a loss layer as a bf16 plane (the Gram gradient multiplies them with bf16 matrices)
  auto feat_bf = [&](int layer) -> const bf16* { return p->split == 2 ? p->act_bf[layer] : p->act[layer]; };
  auto loss_index = [&](int layer) -> int {
    for (size_t k = 0; k < p->loss.size(); ++k)
      if (p->loss[k].layer == layer) return static_cast<int>(k);
    return -1;
  };
  int top = -1;
  for (const LossLayer& ll : p->loss) top = std::max(top, ll.layer);
  bool have = false;
  int cur = 0;
  for (int l = top; l >= 0 && rc == STX_OK; --l) {
    const LayerSpec& s = kLayers[l];
    const int k = loss_index(l);
    if (k >= 0 && have && !g_debug_no_gram_fusion && !p->bwd.empty() && p->bwd.back().kind == BwdStep::kDgrad && p->bwd.back().conv.patch &&
        !p->bwd.back().conv.split) {
      // Gram gradient of this layer fused into the dgrad launch that produces dL/d(act[l]): F * Dscaled is
      // accumulated into the same TMEM tile, then the epilogue masks (conv layers) / exports as configured.
      LossLayer& ll = p->loss[k];
      ConvOp& prod = p->bwd.back().conv;
      rc = conv_patch_attach_gram(&prod, feat_bf(l), ll.dscaled);
      prod.mask = s.is_conv ? p->act[l] : nullptr;
      prod.export_out = ll.dfeat;
      p->bwd.back().inject_k = k;
      if (rc == STX_OK) {
        rc = conv_prepare(&ll.layer_grad_op, feat_bf(l), B, p->lh[l], p->lw[l], ll.C, ll.dscaled, ll.C, 1, 1, 0);
        ll.layer_grad_op.out = p->gbuf[0];
      }
    } else if (k >= 0) {
      LossLayer& ll = p->loss[k];
      BwdStep st;
      st.kind = BwdStep::kGramGrad;
      rc = conv_prepare(&st.conv, feat_bf(l), B, p->lh[l], p->lw[l], ll.C, ll.dscaled, ll.C, 1, 1, 0);
      st.conv.addend = have ? p->gbuf[cur] : nullptr;
      st.conv.out = p->gbuf[cur];
      st.conv.mask = s.is_conv ? p->act[l] : nullptr;
      st.conv.export_out = ll.dfeat;
      st.inject_k = have ? -1 : k;   // the addend slot is free only for the topmost loss layer
      p->bwd.push_back(st);
      have = true;
      if (rc == STX_OK) {
        rc = conv_prepare(&ll.layer_grad_op, feat_bf(l), B, p->lh[l], p->lw[l], ll.C, ll.dscaled, ll.C, 1, 1, 0);
        ll.layer_grad_op.out = p->gbuf[0];
      }
    }
    if (!have || rc != STX_OK) continue;
    if (l == 0) {
      BwdStep st;
      st.kind = BwdStep::kConv1Bwd;
      st.in = p->gbuf[cur];
      // tensor-core input gradient (N tile 16) when the halo-patch kernel covers the shape, CUDA cores otherwise
      const int prc = conv_patch_prepare(&st.conv, p->gbuf[cur], B, H, W, 64, vgg->wd0, 16, 0, 0, 0, 0);
      if (prc == STX_OK) st.kind = BwdStep::kConv1BwdTensor;
      else if (prc != STX_ERR_UNSUPPORTED) rc = prc;
      p->bwd.push_back(st);
    } else if (s.is_conv) {
      BwdStep st;
      st.kind = BwdStep::kDgrad;
      rc = conv3x3_prepare(&st.conv, p->gbuf[cur], nullptr, B, p->lh[l], p->lw[l], s.cout, vgg->wd[s.conv_index],
                           nullptr, s.cin,
                           g_debug_dgrad_variant + ((g_debug_pair & 2) ? 1000000 : 0));
      st.conv.w_tiled = g_debug_tiled_weights ? vgg->wdt[s.conv_index] : nullptr;
      st.conv.out = p->gbuf[cur ^ 1];
      st.conv.mask = (kLayers[l - 1].is_conv && loss_index(l - 1) < 0) ? p->act[l - 1] : nullptr;
      p->bwd.push_back(st);
      cur ^= 1;
    } else if (loss_index(l - 1) < 0 && !g_debug_no_poolbwd_fusion && !p->bwd.empty() &&
               (p->bwd.back().kind == BwdStep::kGramGrad || p->bwd.back().kind == BwdStep::kDgrad)) {
      // AvgPoolGrad + ReluGrad fused into the epilogue of the launch that completes dL/d(pool output): it scatters
      // 0.25 * g to the four fine pixels, masked by the activations of the convolution that fed the pool.
      // Buffer choice: a dgrad launch reads its input patches (with halos, across tiles) from gbuf[cur ^ 1], so its
      // scatter must go to gbuf[cur] - the buffer its plain output would have used (unused in scatter mode). The
      // stand-alone Gram gradient reads its addend from gbuf[cur], so its scatter goes to the other buffer.
      ConvOp& prod = p->bwd.back().conv;
      if (p->bwd.back().kind == BwdStep::kDgrad) {
        prod.up_out = p->gbuf[cur];
      } else {
        prod.up_out = p->gbuf[cur ^ 1];
        cur ^= 1;
      }
      prod.up_mask = p->act[l - 1];
      // The convolution in front of a fused pool: its full-resolution result is read by this scatter alone, and only
      // for its sign - the forward writes one bit per element instead of the plane (conv_common.cuh, bits_out).
      ConvOp& fw = p->fwd[l - 1];
      if (!p->full_features && !g_debug_no_mask_bits && prod.patch && fw.patch && fw.pool_out != nullptr &&
          p->lc[l - 1] % 32 == 0) {
        const size_t words = static_cast<size_t>(B) * p->lh[l - 1] * p->lw[l - 1] * (p->lc[l - 1] / 32);
        uint32_t* bits = nullptr;
        e = cudaMalloc(&bits, words * sizeof(uint32_t));
        if (e != cudaSuccess) {
          rc = cuda_fail(e, "plan_create: ReLU bit masks");
        } else {
          p->mask_bits.push_back(bits);
          p->arena_bytes += words * sizeof(uint32_t);    // reported by stx_plan_device_bytes
          fw.bits_out = bits;
          fw.out = nullptr;
          fw.out_lo = nullptr;
          fw.aux_out = nullptr;
          prod.up_bits = bits;
          prod.up_mask = nullptr;
          p->hi_valid[l - 1] = false;
          p->lo_valid[l - 1] = false;
        }
      }
    } else {
      BwdStep st;
      st.kind = BwdStep::kPoolBwd;
      st.in = p->gbuf[cur];
      st.out = p->gbuf[cur ^ 1];
      st.act = (loss_index(l - 1) < 0) ? p->act[l - 1] : nullptr;
      st.H = p->lh[l - 1];
      st.W = p->lw[l - 1];
      st.C = p->lc[l - 1];
      p->bwd.push_back(st);
      cur ^= 1;
    }
  }
  // Per-image switch of the persistent work loops (all on; the device optimiser turns finished jobs off).
  if (rc == STX_OK) {
    std::vector<int> ones(B, 1);
    e = cudaMemcpy(p->d_active, ones.data(), ones.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = cuda_fail(e, "plan_create: active flags");
    for (int l = 1; l <= topmost; ++l)
      if (kLayers[l].is_conv) p->fwd[l].active = p->d_active;
    for (BwdStep& st : p->bwd)
      if (st.kind == BwdStep::kDgrad || st.kind == BwdStep::kGramGrad || st.kind == BwdStep::kConv1BwdTensor)
        st.conv.active = p->d_active;
  }
  // Split-K for the convolutions whose grid would leave most SMs idle (deep layers at small batch). Implemented and
  // parity-tested, but off by default: at 256^2, batch 1, the per-item fixed cost (pipeline fill, partial-tile
  // round trip through L2, serial fix-up) outweighs the extra parallelism (profiles/r01_splitk_b1.md).
  if (rc == STX_OK && g_debug_splitk) {
    std::vector<ConvOp*> ops;
    for (int l = 1; l <= topmost; ++l)
      if (kLayers[l].is_conv) ops.push_back(&p->fwd[l]);
    for (BwdStep& st : p->bwd)
      if (st.kind == BwdStep::kDgrad) ops.push_back(&st.conv);
    size_t ws_max = 0, cnt_max = 0;
    for (ConvOp* op : ops) {
      size_t ws = 0, cnt = 0;
      op->ksplit = conv_patch_plan_splitk(*op, &ws, &cnt);
      ws_max = std::max(ws_max, ws);
      cnt_max = std::max(cnt_max, cnt);
    }
    if (ws_max > 0) {
      e = cudaMalloc(&p->splitk_ws, ws_max * sizeof(float));
      if (e == cudaSuccess) e = cudaMalloc(&p->splitk_counters, cnt_max * sizeof(int));
      if (e == cudaSuccess) e = cudaMemset(p->splitk_counters, 0, cnt_max * sizeof(int));
      if (e != cudaSuccess) rc = cuda_fail(e, "split-K workspace allocation");
      for (ConvOp* op : ops) {
        op->splitk_ws = p->splitk_ws;
        op->splitk_counters = p->splitk_counters;
      }
    }
  }
  if (rc != STX_OK) {
    delete p;
    return rc;
  }
  *out = p;
  return STX_OK;
}

int Plan::forward(const float* image, int is_preimage, int with_grams, bool with_loss, cudaStream_t stream) {
  STX_REQUIRE(image, "forward: null image");
  ++generation;
  const bool grams = with_grams || with_loss;
  bool forked = false;
  // Gram + finish of loss layer `l`, on the side stream, as soon as act[l] exists.
  auto launch_gram = [&](int l) -> int {
    for (LossLayer& ll : loss) {
      if (ll.layer != l) continue;
      // (the instrumented step keeps everything on one stream so that per-class times are not polluted by waiting)
      const bool use_side = (tl == nullptr);
      if (use_side) {
        STX_CUDA(cudaEventRecord(ev_fork[l], stream));
        STX_CUDA(cudaStreamWaitEvent(side, ev_fork[l], 0));
        forked = true;
      }
      cudaStream_t main_stream = stream;
      {
        cudaStream_t stream = use_side ? side : main_stream;   // STX_LAUNCH records/launches on `stream`
        STX_LAUNCH(kClsGram, gram_launch(ll.gram, stream));
      }
    }
    return STX_OK;
  };
  // Finishes (split-K reduction, G - T, scaled difference, loss partials) in two launches on the side stream: every
  // loss layer but the deepest as soon as the last of their Grams has been issued (it hides under the convolutions
  // of the deepest block), then the deepest one alone - the only finish on the critical path to the backward pass.
  // Five separate ~10 us launches were mostly launch latency; a single launch at the very end put all of it on the
  // critical path (measured slower end to end).
  int deepest = -1;
  for (const LossLayer& ll : loss) deepest = std::max(deepest, ll.layer);
  auto launch_finish = [&](bool deepest_only) -> int {
    GramFinishBatch fb;
    fb.part_stride = part_total;
    for (LossLayer& ll : loss) {
      if ((ll.layer == deepest) != deepest_only) continue;
      // dL/dF = F * (G - T) * coef / (C^2 * n):  loss_overall = sum coef/4 * mean((G-T)^2), G = F^T F / n
      const float norm = static_cast<float>(ll.HW) + 1e-6f;
      const float dscale = ll.coef / (static_cast<float>(ll.C) * static_cast<float>(ll.C) * norm);
      const float lscale = 1.0f / (static_cast<float>(ll.C) * static_cast<float>(ll.C));
      STX_TRY(gram_finish_batch_add(&fb, ll.gram, 1e-6f, ll.G, with_loss ? ll.target : nullptr, ll.target_stride, dscale,
                                    with_loss ? ll.dscaled : nullptr, lscale,
                                    with_loss ? loss_part + ll.part_offset : nullptr));
    }
    if (fb.n == 0) return STX_OK;
    const bool use_side = (tl == nullptr) && forked;
    cudaStream_t main_stream = stream;
    {
      cudaStream_t stream = use_side ? side : main_stream;
      STX_LAUNCH(kClsGramFinish, gram_finish_batch_launch(fb, stream));
    }
    return STX_OK;
  };
  int second_deepest = -1;
  for (const LossLayer& ll : loss)
    if (ll.layer != deepest) second_deepest = std::max(second_deepest, ll.layer);
  auto after_layer = [&](int l) -> int {
    if (!grams) return STX_OK;
    STX_TRY(launch_gram(l));
    if (l == second_deepest) STX_TRY(launch_finish(false));
    if (l == deepest) STX_TRY(launch_finish(true));
    return STX_OK;
  };
  if (split == 2 && !g_debug_conv1_1_cuda_cores) {
    STX_LAUNCH(kClsConv1, conv1_1_tc_launch(image, B, H, W, is_preimage, vgg->mean, vgg->wtc0, vgg->corr_scale[0],
                                            vgg->bias[0], act[0], act_lo[0], act_bf[0], stream, d_active));
  } else {
    STX_LAUNCH(kClsConv1, conv1_1_fwd_launch(image, B, H, W, is_preimage, vgg->mean, vgg->w0, vgg->bias[0], act[0],
                                             act_lo[0], stream, split == 2 ? 2 : 0, act_bf[0]));
  }
  STX_TRY(after_layer(0));
  for (int l = 1; l <= topmost; ++l) {
    if (kLayers[l].is_conv) {
      STX_LAUNCH(kClsConvFwd, conv_launch(fwd[l], stream));
    } else if (!pool_fused[l]) {
      if (split == 2) return fail(STX_ERR_UNSUPPORTED, "forward: f16+f8 planes have no stand-alone pooling kernel");
      STX_LAUNCH(kClsPool, avgpool2_fwd_launch(act[l - 1], act_lo[l - 1], B, lh[l - 1], lw[l - 1], lc[l - 1], act[l], act_lo[l], stream));
    }
    STX_TRY(after_layer(l));
  }
  if (forked) {
    STX_CUDA(cudaEventRecord(ev_join, side));
    STX_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
  }
  if (grams) have_grams = true;
  return STX_OK;
}

int Plan::backward(const float* image, int is_preimage, float* grad, cudaStream_t stream,
                   const bf16* const* inject) {
  if (inject) {
    for (size_t k = 0; k < loss.size(); ++k) {
      if (!inject[k]) continue;
      bool found = false;
      for (const BwdStep& st : bwd) found = found || st.inject_k == static_cast<int>(k);
      if (!found) return fail(STX_ERR_UNSUPPORTED, "backward: no launch can take the external gradient of this loss layer");
    }
  }
  for (const BwdStep& st : bwd) {
    const bf16* extra = (inject && st.inject_k >= 0) ? inject[st.inject_k] : nullptr;
    switch (st.kind) {
      case BwdStep::kGramGrad:
        if (extra) {
          ConvOp op = st.conv;
          op.addend = extra;
          STX_LAUNCH(kClsGramGrad, conv_launch(op, stream));
        } else {
          STX_LAUNCH(kClsGramGrad, conv_launch(st.conv, stream));
        }
        break;
      case BwdStep::kDgrad:
        if (extra) {
          ConvOp op = st.conv;
          op.addend = extra;
          STX_LAUNCH(kClsConvDgrad, conv_launch(op, stream));
        } else {
          STX_LAUNCH(kClsConvDgrad, conv_launch(st.conv, stream));
        }
        break;
      case BwdStep::kPoolBwd:
        STX_LAUNCH(kClsPool, avgpool2_bwd_mask_launch(st.in, st.act, B, st.H, st.W, st.C, st.out, stream));
        break;
      case BwdStep::kConv1Bwd:
        STX_LAUNCH(kClsConv1, conv1_1_bwd_launch(st.in, image, B, H, W, is_preimage, vgg->w0, grad, stream));
        break;
      case BwdStep::kConv1BwdTensor: {
        ConvOp op = st.conv;
        op.rgb_out = grad;
        op.rgb_image = image;
        op.rgb_preimage = is_preimage;
        STX_LAUNCH(kClsConv1, conv_launch(op, stream));
        break;
      }
    }
  }
  return STX_OK;
}

int Plan::backward_custom(const float* image, int is_preimage, const float* const* M, const bf16* const* E,
                          float* grad, cudaStream_t stream) {
  STX_REQUIRE(image && M && grad, "backward_custom: null argument");
  if (loss.empty()) return fail(STX_ERR_STATE, "backward_custom: the plan has no loss layers");
  if (bwd.empty() || (bwd.back().kind != BwdStep::kConv1Bwd && bwd.back().kind != BwdStep::kConv1BwdTensor))
    return fail(STX_ERR_STATE, "backward_custom: backward program incomplete");
  for (size_t k = 0; k < loss.size(); ++k) {
    LossLayer& ll = loss[k];
    STX_REQUIRE(M[k], "backward_custom: null matrix");
    STX_TRY(f32_to_bf16_launch(M[k], static_cast<size_t>(B) * ll.C * ll.C, ll.dscaled, stream));
  }
  return backward(image, is_preimage, grad, stream, E);
}

int Plan::step(const float* x, int is_preimage, float* func, float* grad, float* layer_loss_out,
               cudaStream_t stream) {
  STX_REQUIRE(x && grad, "step: null x or grad");
  if (loss.empty()) return fail(STX_ERR_STATE, "step: the plan has no loss layers");
  if (!have_target) return fail(STX_ERR_STATE, "step: no target Grams set (call compute_gram(update=True) first)");
  if (bwd.empty() || (bwd.back().kind != BwdStep::kConv1Bwd && bwd.back().kind != BwdStep::kConv1BwdTensor))
    return fail(STX_ERR_STATE, "step: backward program incomplete");
  STX_TRY(forward(x, is_preimage, 1, true, stream));
  STX_TRY(launch_pdl(loss_reduce_kernel, dim3(B), dim3(128), 0, stream, loss_part, part_total, d_part_offsets,
                     static_cast<int>(loss.size()), d_coef, B, layer_loss, func ? func : func_dev));
  count_launch();
  if (layer_loss_out)
    STX_CUDA(cudaMemcpyAsync(layer_loss_out, layer_loss, loss.size() * B * sizeof(float), cudaMemcpyDeviceToDevice,
                             stream));
  STX_TRY(backward(x, is_preimage, grad, stream));
  return STX_OK;
}

int Plan::step_timed(const float* x, int is_preimage, float* func, float* grad, float* ms, int* count,
                     cudaStream_t stream) {
  STX_REQUIRE(ms && count, "step_timed: null output");
  Timeline local;
  tl = &local;
  int rc = step(x, is_preimage, func, grad, nullptr, stream);
  tl = nullptr;   // ordinary steps stay event-free
  const cudaError_t e = cudaStreamSynchronize(stream);
  if (rc == STX_OK && e != cudaSuccess) rc = cuda_fail(e, "step_timed sync");
  for (int c = 0; c < kNumClasses; ++c) {
    ms[c] = 0.f;
    count[c] = 0;
  }
  if (rc == STX_OK) {
    for (const Timeline::Rec& r : local.recs) {
      float v = 0.f;
      cudaEventElapsedTime(&v, r.a, r.b);
      ms[r.cls] += v;
      count[r.cls] += 1;
    }
  }
  return rc;
}

}  // namespace stx
