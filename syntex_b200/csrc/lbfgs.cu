// Device-resident projected L-BFGS with box bounds: the optimiser loop of gatys/synthesizer.py:127-151
// (scipy L-BFGS-B there) as a state machine that advances by exactly ONE f/g evaluation per cycle and never
// returns to the host inside a cycle. oracle/lbfgs_model.py is the line-by-line numpy model of this file.
//
//   cycle = [ plan.step(xt) -> ft, gt ]  ->  dots kernel  ->  decide kernel  ->  update kernel
//
//   dots   : one pass over x, xt, g, gt, d and the (s, y) history: every inner product the decision needs
//            (line-search slope, s'y, y'y, S'y, Y'y, S'Pg, Y'Pg, |Pg|^2) as per-block partial sums
//            (warp-shuffle + shared-memory reductions, fixed order => bit-reproducible).
//   decide : one CTA per problem: sums the partials in fp64, advances the More'-Thuente line search, and on
//            acceptance updates S'Y / Y'Y and solves the compact-form (Byrd-Nocedal-Schnabel) two-loop
//            equivalent for the coefficients of the new direction.
//   update : applies the decision: stores the new pair, moves x, builds d = -P H P g from the coefficients in
//            one pass over the history, and writes the next trial point xt = clip(x + stp d).
//
// Each image of the plan's batch is an independent problem (own history, own line search); problems that are
// finished idle through the remaining cycles.
#include "lbfgs.h"

#include "ptx.cuh"

namespace stx {
int g_lbfgs_no_graph = 0;   // 1 = launch every optimiser cycle on the stream instead of replaying a CUDA graph (key 10)
}

#include <algorithm>
#include <vector>

namespace stx {

namespace {

constexpr int kMaxM = 32;
constexpr int kVec = 4;
constexpr int kThreads = 256;
constexpr int kChunk = kThreads * kVec;     // elements per block
constexpr int kNP = 6 + 4 * kMaxM;          // partial sums per block (dots kernel)

constexpr double FTOL = 1e-3, GTOL = 0.9, XTOL = 0.1, XTRAPL = 1.1, XTRAPU = 4.0, STPMAX = 1e10, EPSMCH = 2.2e-16;
constexpr int MAX_LS = 20;

enum Mode { kStart = 0, kSearch = 1, kDone = 2 };
enum Action { kNone = 0, kAccept = 1, kRetry = 2, kRestart = 3, kStartAccept = 4, kFinalAccept = 5 };

struct LsState {
  double stx, fx, dx, sty, fy, dy, stmin, stmax, width, width1, finit, ginit, gtest, stp;
  int brackt, stage, nev;
};

struct Ctrl {
  int mode, action, push, nh, head, slot_new, nit, nfev, fresh, maxiter;
  float stp;          // step of the trial point currently in xt
  float gamma;
  float ca[kMaxM], cb[kMaxM];
  double f, pp, stp0;
  LsState ls;
  double SY[kMaxM * kMaxM], YY[kMaxM * kMaxM];   // chronological order, 0 = oldest
};

struct Vecs {
  float *x, *xt, *g, *gt, *d, *S, *Y;   // [B][n], history [B][m][n]
  float* ft;                             // [B]
  float* part;                           // [B][nblk][kNP]
  float* part2;                          // [B][nblk]  (g'd of a freshly built direction)
  Ctrl* ctrl;                            // [B]
  int n, m, nblk, bounded;
  float lo, hi;
};

__device__ __forceinline__ float clipf(float z, float lo, float hi, int bounded) {
  return bounded ? fminf(fmaxf(z, lo), hi) : z;
}
__device__ __forceinline__ bool is_free(float x, float g, float lo, float hi, int bounded) {
  return !bounded || !((x <= lo && g > 0.0f) || (x >= hi && g < 0.0f));
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------ dots
__global__ void __launch_bounds__(kThreads) lbfgs_dots_kernel(Vecs v) {
  pdl_launch_dependents();   // programmatic dependent launch: see launch_pdl (common.h)
  pdl_wait();
  const int b = blockIdx.y;
  const Ctrl& c = v.ctrl[b];
  if (c.mode == kDone) return;
  __shared__ float red[kThreads / 32][kNP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * kChunk + threadIdx.x * kVec;
  const bool in = i0 < v.n;
  const size_t off = static_cast<size_t>(b) * v.n + i0;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 x = in ? ld4(v.x + off) : z4;
  const float4 xt = in ? ld4(v.xt + off) : z4;
  const float4 g = in ? ld4(v.g + off) : z4;
  const float4 gt = in ? ld4(v.gt + off) : z4;
  const float4 d = in ? ld4(v.d + off) : z4;
  const float stp = c.stp;
  const float lo = v.lo, hi = v.hi;
  const int bd = v.bounded;

  float4 s, y, pg;
  s.x = xt.x - x.x; s.y = xt.y - x.y; s.z = xt.z - x.z; s.w = xt.w - x.w;
  y.x = gt.x - g.x; y.y = gt.y - g.y; y.z = gt.z - g.z; y.w = gt.w - g.w;
  pg.x = is_free(xt.x, gt.x, lo, hi, bd) ? gt.x : 0.f;
  pg.y = is_free(xt.y, gt.y, lo, hi, bd) ? gt.y : 0.f;
  pg.z = is_free(xt.z, gt.z, lo, hi, bd) ? gt.z : 0.f;
  pg.w = is_free(xt.w, gt.w, lo, hi, bd) ? gt.w : 0.f;
  // slope of phi(a) = f(clip(x + a d)): only coordinates that are not clipped move with a
  float dphi = 0.f;
  {
    const float zx = fmaf(stp, d.x, x.x), zy = fmaf(stp, d.y, x.y), zz = fmaf(stp, d.z, x.z), zw = fmaf(stp, d.w, x.w);
    if (!bd || (zx >= lo && zx <= hi)) dphi += gt.x * d.x;
    if (!bd || (zy >= lo && zy <= hi)) dphi += gt.y * d.y;
    if (!bd || (zz >= lo && zz <= hi)) dphi += gt.z * d.z;
    if (!bd || (zw >= lo && zw <= hi)) dphi += gt.w * d.w;
  }
  float vals[6] = {dphi, dot4(s, y), dot4(y, y), dot4(pg, pg), dot4(s, pg), dot4(y, pg)};
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float r = warp_sum(vals[k]);
    if (lane == 0) red[warp][k] = r;
  }
  const int nh = c.nh, head = c.head, m = v.m;
  for (int j = 0; j < nh; ++j) {
    const int slot = (head + j) % m;
    const size_t hoff = (static_cast<size_t>(b) * m + slot) * v.n + i0;
    const float4 sj = in ? ld4(v.S + hoff) : z4;
    const float4 yj = in ? ld4(v.Y + hoff) : z4;
    const float a0 = warp_sum(dot4(sj, y));
    const float a1 = warp_sum(dot4(yj, y));
    const float a2 = warp_sum(dot4(sj, pg));
    const float a3 = warp_sum(dot4(yj, pg));
    if (lane == 0) {
      red[warp][6 + 4 * j + 0] = a0;
      red[warp][6 + 4 * j + 1] = a1;
      red[warp][6 + 4 * j + 2] = a2;
      red[warp][6 + 4 * j + 3] = a3;
    }
  }
  __syncthreads();
  const int np = 6 + 4 * nh;
  float* dst = v.part + (static_cast<size_t>(b) * v.nblk + blockIdx.x) * kNP;
  for (int k = threadIdx.x; k < np; k += kThreads) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += red[w][k];
    dst[k] = t;
  }
}

// ------------------------------------------------------------------------------------------------ line search
__device__ double dcstep(LsState& st, double stp, double fp, double dp, double stpmin, double stpmax) {
  double stx = st.stx, fx = st.fx, dx = st.dx, sty = st.sty, fy = st.fy, dy = st.dy;
  int brackt = st.brackt;
  const double sgnd = dp * (dx / fabs(dx));
  double stpf;
  if (fp > fx) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp < stx) gamma = -gamma;
    const double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
    const double stpc = stx + r * (stp - stx);
    const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
    const double stpc = stp + r * (stx - stp);
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
    double stpc;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
      if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
      else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
      stpf = fmin(stpmax, fmax(stpmin, stpf));
    }
  } else {
    if (brackt) {
      const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      const double s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
      double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dy / s) * (dp / s)));
      if (stp > sty) gamma = -gamma;
      const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dy, r = p / q;
      stpf = stp + r * (sty - stp);
    } else if (stp > stx) {
      stpf = stpmax;
    } else {
      stpf = stpmin;
    }
  }
  if (fp > fx) {
    sty = stp; fy = fp; dy = dp;
  } else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  st.stx = stx; st.fx = fx; st.dx = dx; st.sty = sty; st.fy = fy; st.dy = dy; st.brackt = brackt;
  return stpf;
}

__device__ void ls_init(LsState& st, double f0, double g0, double stp) {
  st.brackt = 0; st.stage = 1; st.finit = f0; st.ginit = g0; st.gtest = FTOL * g0;
  st.width = STPMAX; st.width1 = 2.0 * STPMAX;
  st.stx = 0.0; st.fx = f0; st.dx = g0; st.sty = 0.0; st.fy = f0; st.dy = g0;
  st.stmin = 0.0; st.stmax = stp + XTRAPU * stp; st.stp = stp; st.nev = 0;
}

// 0 = accept, 1 = fail, 2 = evaluate again at st.stp
__device__ int ls_update(LsState& st, double f, double g) {
  const double stp = st.stp;
  st.nev += 1;
  const double ftest = st.finit + stp * st.gtest;
  if (st.stage == 1 && f <= ftest && g >= 0.0) st.stage = 2;
  bool warn = false;
  if (st.brackt && (stp <= st.stmin || stp >= st.stmax)) warn = true;
  if (st.brackt && st.stmax - st.stmin <= XTOL * st.stmax) warn = true;
  if (stp == STPMAX && f <= ftest && g <= st.gtest) warn = true;
  if (stp == 0.0 && (f > ftest || g >= st.gtest)) warn = true;
  if (f <= ftest && fabs(g) <= GTOL * (-st.ginit)) return 0;
  if (warn || st.nev >= MAX_LS) return 1;
  double nw;
  if (st.stage == 1 && f <= st.fx && f > ftest) {
    const double gt = st.gtest;
    LsState m = st;
    m.fx = st.fx - st.stx * gt; m.dx = st.dx - gt; m.fy = st.fy - st.sty * gt; m.dy = st.dy - gt;
    nw = dcstep(m, stp, f - stp * gt, g - gt, st.stmin, st.stmax);
    st.stx = m.stx; st.fx = m.fx + m.stx * gt; st.dx = m.dx + gt;
    st.sty = m.sty; st.fy = m.fy + m.sty * gt; st.dy = m.dy + gt; st.brackt = m.brackt;
  } else {
    nw = dcstep(st, stp, f, g, st.stmin, st.stmax);
  }
  if (st.brackt) {
    if (fabs(st.sty - st.stx) >= 0.66 * st.width1) nw = st.stx + 0.5 * (st.sty - st.stx);
    st.width1 = st.width;
    st.width = fabs(st.sty - st.stx);
    st.stmin = fmin(st.stx, st.sty);
    st.stmax = fmax(st.stx, st.sty);
  } else {
    st.stmin = nw + XTRAPL * (nw - st.stx);
    st.stmax = nw + XTRAPU * (nw - st.stx);
  }
  nw = fmin(STPMAX, fmax(0.0, nw));
  if (st.brackt && (nw <= st.stmin || nw >= st.stmax || st.stmax - st.stmin <= XTOL * st.stmax)) nw = st.stx;
  st.stp = nw;
  return 2;
}

// ------------------------------------------------------------------------------------------------ decide
// S'Y and Y'Y are stored by history SLOT (row / column of pair i = its slot (head + i) % m), like the vectors
// themselves: dropping the oldest pair is then a head increment, not a shift of two m x m tables.
// slot_of[i] = (head + i) % m, rebuilt whenever head moves (a runtime modulo per table access costs more than the
// shift it replaces: measured 0.425 vs 0.352 ms of optimiser time per cycle at 18 jobs).
__device__ __forceinline__ void build_slots(const Ctrl& c, int m, int* slot_of) {
  int s = c.head;
  for (int i = 0; i < m; ++i) {
    slot_of[i] = s;
    if (++s == m) s = 0;
  }
}
__device__ __forceinline__ int tab(const int* slot_of, int i, int j) { return slot_of[i] * kMaxM + slot_of[j]; }

__device__ void solve_direction(Ctrl& c, const int* so, const double* p1, const double* yp) {
  // H v = gamma v + S a + gamma Y b,  a = R^-T ((D + gamma Y'Y) R^-1 p1 - gamma Y'v),  b = -R^-1 p1
  const int k = c.nh;
  const double gamma = c.SY[tab(so, k - 1, k - 1)] / c.YY[tab(so, k - 1, k - 1)];
  double u[kMaxM], t[kMaxM], a[kMaxM];
  for (int i = k - 1; i >= 0; --i) {   // R u = p1, R upper triangular (R_ij = s_i'y_j, i <= j)
    double acc = p1[i];
    for (int j = i + 1; j < k; ++j) acc -= c.SY[tab(so, i, j)] * u[j];
    u[i] = acc / c.SY[tab(so, i, i)];
  }
  for (int i = 0; i < k; ++i) {
    double acc = c.SY[tab(so, i, i)] * u[i];
    for (int j = 0; j < k; ++j) acc += gamma * c.YY[tab(so, i, j)] * u[j];
    t[i] = acc - gamma * yp[i];
  }
  for (int i = 0; i < k; ++i) {        // R^T a = t
    double acc = t[i];
    for (int j = 0; j < i; ++j) acc -= c.SY[tab(so, j, i)] * a[j];
    a[i] = acc / c.SY[tab(so, i, i)];
  }
  c.gamma = static_cast<float>(gamma);
  for (int i = 0; i < k; ++i) {
    c.ca[i] = static_cast<float>(a[i]);
    c.cb[i] = static_cast<float>(-gamma * u[i]);
  }
}

// The serial part of a cycle's decision (line search, history bookkeeping, the m x m solves), run by ONE thread on a
// SHARED-MEMORY copy of the job's control block: on the global copy every one of the ~1500 dependent accesses to
// S'Y / Y'Y (shifting the 20 x 20 tables when the oldest pair is dropped, three triangular passes) was an L2 round
// trip, and this kernel - 57-70 us with a full history - was as long as the two streaming kernels around it.
__device__ void decide_body(Ctrl& c, const double* sums, const Vecs& v, int b, int* so) {

  const double ft = static_cast<double>(v.ft[b]);
  const double dphi = sums[0], ys = sums[1], yy = sums[2], ppn = sums[3], spg = sums[4], ypg = sums[5];
  const double dg0 = sums[kNP];
  const int m = v.m;
  c.nfev += 1;
  c.push = 0;

  auto first_step = [](double pp) { return pp > 0.0 ? fmin(1.0, 1.0 / sqrt(pp)) : 1.0; };

  if (c.mode == kStart) {
    c.f = ft;
    c.pp = ppn;
    c.nh = 0;
    c.head = 0;
    if (!(ppn > 0.0)) {              // stationary starting point: nothing to do, not resumable
      c.action = kFinalAccept;
      c.mode = kDone;
      return;
    }
    c.stp0 = first_step(ppn);
    c.stp = static_cast<float>(c.stp0);
    c.fresh = 1;
    c.mode = (c.maxiter <= 0) ? kDone : kSearch;   // budget stop keeps the pending trial point (resumable)
    c.action = kStartAccept;
    return;
  }

  int res;
  if (c.fresh) {
    if (!(dg0 < 0.0)) {
      res = 1;                       // not a descent direction (numerical breakdown): treat as a failed search
    } else {
      ls_init(c.ls, c.f, dg0, c.stp0);
      c.fresh = 0;
      res = ls_update(c.ls, ft, dphi);
    }
  } else {
    res = ls_update(c.ls, ft, dphi);
  }

  if (res == 0) {
    double p1[kMaxM], yp[kMaxM];
    int k = c.nh;
    int shift = 0;
    const bool push = ys > EPSMCH * yy;
    if (push) {
      if (k == m) {                  // drop the oldest pair: the tables are indexed by slot, nothing moves
        c.head = (c.head + 1) % m;
        k -= 1;
        shift = 1;
      }
      build_slots(c, m, so);
      for (int i = 0; i < k; ++i) {
        const int src = 6 + 4 * (i + shift);
        c.SY[tab(so, i, k)] = sums[src + 0];   // s_i . y_new
        c.YY[tab(so, i, k)] = sums[src + 1];   // y_i . y_new
        c.YY[tab(so, k, i)] = sums[src + 1];
        c.SY[tab(so, k, i)] = 0.0;             // lower triangle of S'Y is never used
      }
      c.SY[tab(so, k, k)] = ys;
      c.YY[tab(so, k, k)] = yy;
      c.slot_new = (c.head + k) % m;
      c.push = 1;
    }
    for (int i = 0; i < k; ++i) {
      const int src = 6 + 4 * (i + shift);
      p1[i] = sums[src + 2];
      yp[i] = sums[src + 3];
    }
    if (push) {
      p1[k] = spg;
      yp[k] = ypg;
      k += 1;
    }
    c.nh = k;
    c.f = ft;
    c.pp = ppn;
    c.nit += 1;
    if (!(ppn > 0.0)) {              // projected gradient vanished: converged
      c.action = kFinalAccept;
      c.mode = kDone;
      c.fresh = 0;
      return;
    }
    // On the iteration budget the next direction and trial point are still built (one extra vector pass), so a
    // later run() with a larger budget continues the same optimisation.
    if (c.nit >= c.maxiter) c.mode = kDone;
    if (k > 0) {
      solve_direction(c, so, p1, yp);
      c.stp0 = 1.0;
    } else {
      c.stp0 = first_step(ppn);
    }
    c.stp = static_cast<float>(c.stp0);
    c.fresh = 1;
    c.action = kAccept;
  } else if (res == 1) {
    if (c.nh > 0) {                  // drop the memory and restart from steepest descent at x (as L-BFGS-B does)
      c.nh = 0;
      c.head = 0;
      c.stp0 = first_step(c.pp);
      c.stp = static_cast<float>(c.stp0);
      c.fresh = 1;
      c.action = kRestart;
    } else {
      c.mode = kDone;                // line search failed along steepest descent: give up for good
      c.fresh = 0;
      c.action = kNone;
    }
  } else {
    c.stp = static_cast<float>(c.ls.stp);
    c.action = kRetry;
  }
}

__global__ void __launch_bounds__(512) lbfgs_decide_kernel(Vecs v) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  Ctrl& cg = v.ctrl[b];
  if (cg.mode == kDone) {
    if (threadIdx.x == 0) cg.action = kNone;
    return;
  }
  __shared__ double sums[kNP + 1];
  __shared__ Ctrl sc;
  static_assert(sizeof(Ctrl) % 4 == 0, "Ctrl is copied word by word");
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&cg);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sc);
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(Ctrl) / 4); i += blockDim.x) dst[i] = src[i];
  }
  const int nh = cg.nh;
  const int np = 6 + 4 * nh;
  // Sum the block partials in a fixed order: one warp per quantity, lane l takes blocks l, l+32, ... (a serial chain
  // of nblk dependent L2 round trips per quantity used to make this kernel the longest of the optimiser, ~57 us),
  // then a butterfly over the lanes. Deterministic: the summation tree depends only on nblk.
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int k = warp; k <= np; k += nwarps) {
      double t = 0.0;
      if (k < np) {
        const float* src = v.part + static_cast<size_t>(b) * v.nblk * kNP + k;
        for (int j = lane; j < v.nblk; j += 32) t += static_cast<double>(src[static_cast<size_t>(j) * kNP]);
      } else {
        const float* src = v.part2 + static_cast<size_t>(b) * v.nblk;
        for (int j = lane; j < v.nblk; j += 32) t += static_cast<double>(src[j]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) sums[k < np ? k : kNP] = t;
    }
  }
  __syncthreads();
  __shared__ int slot_of[kMaxM];
  if (threadIdx.x == 0) {
    build_slots(sc, v.m, slot_of);
    decide_body(sc, sums, v, b, slot_of);
  }
  __syncthreads();
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&sc);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&cg);
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(Ctrl) / 4); i += blockDim.x) dst[i] = src[i];
  }
}

// ------------------------------------------------------------------------------------------------ update
__global__ void __launch_bounds__(kThreads) lbfgs_update_kernel(Vecs v) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const Ctrl& c = v.ctrl[b];
  const int action = c.action;
  if (action == kNone) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * kChunk + threadIdx.x * kVec;
  const bool in = i0 < v.n;
  const size_t off = static_cast<size_t>(b) * v.n + i0;
  const float lo = v.lo, hi = v.hi;
  const int bd = v.bounded;
  const float stp = c.stp;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float gd = 0.f;   // partial g'd of a new direction
  bool new_dir = false;

  if (action == kRetry) {
    if (in) {
      const float4 x = ld4(v.x + off), d = ld4(v.d + off);
      float4 t;
      t.x = clipf(fmaf(stp, d.x, x.x), lo, hi, bd);
      t.y = clipf(fmaf(stp, d.y, x.y), lo, hi, bd);
      t.z = clipf(fmaf(stp, d.z, x.z), lo, hi, bd);
      t.w = clipf(fmaf(stp, d.w, x.w), lo, hi, bd);
      st4(v.xt + off, t);
    }
    return;
  }

  float4 x = z4, g = z4, s = z4, y = z4;
  if (action == kRestart) {
    if (in) {
      x = ld4(v.x + off);
      g = ld4(v.g + off);
    }
  } else {   // kAccept, kStartAccept, kFinalAccept: the trial point becomes the iterate
    if (in) {
      const float4 xo = ld4(v.x + off), go = ld4(v.g + off);
      x = ld4(v.xt + off);
      g = ld4(v.gt + off);
      s.x = x.x - xo.x; s.y = x.y - xo.y; s.z = x.z - xo.z; s.w = x.w - xo.w;
      y.x = g.x - go.x; y.y = g.y - go.y; y.z = g.z - go.z; y.w = g.w - go.w;
      st4(v.x + off, x);
      st4(v.g + off, g);
      if (action == kAccept && c.push) {
        const size_t hoff = (static_cast<size_t>(b) * v.m + c.slot_new) * v.n + i0;
        st4(v.S + hoff, s);
        st4(v.Y + hoff, y);
      }
    }
    if (action == kFinalAccept) return;
  }

  // New direction d = -P (gamma Pg + S a + Y b)   (a, b from the decide kernel; steepest descent without history)
  float4 pg;
  const bool f0 = is_free(x.x, g.x, lo, hi, bd), f1 = is_free(x.y, g.y, lo, hi, bd);
  const bool f2 = is_free(x.z, g.z, lo, hi, bd), f3 = is_free(x.w, g.w, lo, hi, bd);
  pg.x = f0 ? g.x : 0.f; pg.y = f1 ? g.y : 0.f; pg.z = f2 ? g.z : 0.f; pg.w = f3 ? g.w : 0.f;
  float4 d;
  const int nh = (action == kAccept) ? c.nh : 0;
  if (nh == 0) {
    d.x = -pg.x; d.y = -pg.y; d.z = -pg.z; d.w = -pg.w;
  } else {
    const float gamma = c.gamma;
    float4 acc = make_float4(gamma * pg.x, gamma * pg.y, gamma * pg.z, gamma * pg.w);
    const int nold = c.push ? nh - 1 : nh;
    for (int j = 0; j < nold; ++j) {
      const int slot = (c.head + j) % v.m;
      const size_t hoff = (static_cast<size_t>(b) * v.m + slot) * v.n + i0;
      const float4 sj = in ? ld4(v.S + hoff) : z4;
      const float4 yj = in ? ld4(v.Y + hoff) : z4;
      const float a = c.ca[j], bb = c.cb[j];
      acc.x = fmaf(a, sj.x, fmaf(bb, yj.x, acc.x));
      acc.y = fmaf(a, sj.y, fmaf(bb, yj.y, acc.y));
      acc.z = fmaf(a, sj.z, fmaf(bb, yj.z, acc.z));
      acc.w = fmaf(a, sj.w, fmaf(bb, yj.w, acc.w));
    }
    if (c.push) {
      const float a = c.ca[nh - 1], bb = c.cb[nh - 1];
      acc.x = fmaf(a, s.x, fmaf(bb, y.x, acc.x));
      acc.y = fmaf(a, s.y, fmaf(bb, y.y, acc.y));
      acc.z = fmaf(a, s.z, fmaf(bb, y.z, acc.z));
      acc.w = fmaf(a, s.w, fmaf(bb, y.w, acc.w));
    }
    d.x = f0 ? -acc.x : 0.f; d.y = f1 ? -acc.y : 0.f; d.z = f2 ? -acc.z : 0.f; d.w = f3 ? -acc.w : 0.f;
  }
  new_dir = true;
  if (in) {
    st4(v.d + off, d);
    float4 t;
    t.x = clipf(fmaf(stp, d.x, x.x), lo, hi, bd);
    t.y = clipf(fmaf(stp, d.y, x.y), lo, hi, bd);
    t.z = clipf(fmaf(stp, d.z, x.z), lo, hi, bd);
    t.w = clipf(fmaf(stp, d.w, x.w), lo, hi, bd);
    st4(v.xt + off, t);
    gd = dot4(g, d);
  }
  if (new_dir) {
    __shared__ float red[kThreads / 32];
    const float r = warp_sum(gd);
    if (lane == 0) red[warp] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) t += red[w];
      v.part2[static_cast<size_t>(b) * v.nblk + blockIdx.x] = t;
    }
  }
}

__global__ void lbfgs_init_kernel(Vecs v, const float* __restrict__ x0, int maxiter_unused) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < v.n) {
    const size_t off = static_cast<size_t>(b) * v.n + i;
    const float z = clipf(x0[off], v.lo, v.hi, v.bounded);
    v.x[off] = z;
    v.xt[off] = z;
    v.d[off] = 0.f;
    v.g[off] = 0.f;
  }
  if (i == 0) {
    Ctrl& c = v.ctrl[b];
    c.mode = kStart; c.action = kNone; c.push = 0; c.nh = 0; c.head = 0; c.slot_new = 0;
    c.nit = 0; c.nfev = 0; c.fresh = 0; c.maxiter = 0; c.stp = 0.f; c.gamma = 1.f;
    c.f = 0.0; c.pp = 0.0; c.stp0 = 1.0;
  }
}

__global__ void lbfgs_set_maxiter_kernel(Vecs v, int maxiter, int nprob) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nprob) {
    Ctrl& c = v.ctrl[b];
    c.maxiter = maxiter;
    // resume a run that stopped on its iteration budget: its next trial point is already waiting in xt
    if (c.mode == kDone && c.fresh && c.nit < maxiter) c.mode = kSearch;
  }
}

__global__ void lbfgs_export_kernel(Vecs v, float* fun, int* nit, int* nfev, int* mode, int nprob) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nprob) {
    const Ctrl& c = v.ctrl[b];
    if (fun) fun[b] = static_cast<float>(c.f);
    if (nit) nit[b] = c.nit;
    if (nfev) nfev[b] = c.nfev;
    if (mode) mode[b] = c.mode;
  }
}

}  // namespace

struct LbfgsImpl {
  Vecs v;
  int B = 0;
  float* base = nullptr;
  float* fun_dev = nullptr;
  int* status_dev = nullptr;   // [3B] nit | nfev | mode
  int* pinned_i = nullptr;     // [3B]
  float* pinned_f = nullptr;   // [n*B] x staging + [B] fun
  // The cycle is replayed as a CUDA graph. Capture is not allowed on the legacy default stream, so the optimiser
  // owns a stream and fences it against the caller's stream at entry and exit.
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  cudaGraphExec_t graph = nullptr;
  long long graph_nodes = 0;   // kernel launches per replay
  int graph_preimage = -1;
  bool started = false;
};

static LbfgsImpl* impl_of(Lbfgs* o) { return static_cast<LbfgsImpl*>(o->arena); }

Lbfgs::~Lbfgs() {
  LbfgsImpl* im = impl_of(this);
  if (im) {
    if (im->graph) cudaGraphExecDestroy(im->graph);
    if (im->stream) cudaStreamDestroy(im->stream);
    if (im->ev_in) cudaEventDestroy(im->ev_in);
    if (im->ev_out) cudaEventDestroy(im->ev_out);
    cudaFree(im->base);
    cudaFree(im->v.ctrl);
    if (im->pinned_i) cudaFreeHost(im->pinned_i);
    if (im->pinned_f) cudaFreeHost(im->pinned_f);
    delete im;
  }
  arena = nullptr;
}

int lbfgs_create(Lbfgs** out, Plan* plan, int maxcor, int bounded, float lower, float upper) {
  STX_REQUIRE(out && plan, "lbfgs_create: null argument");
  STX_REQUIRE(maxcor >= 1 && maxcor <= kMaxM, "lbfgs_create: maxcor must be in [1,32]");
  STX_REQUIRE(!bounded || lower < upper, "lbfgs_create: empty box");
  const size_t n = plan->image_elems() / plan->B;
  STX_REQUIRE(n % kVec == 0, "lbfgs_create: problem size must be a multiple of 4");
  Lbfgs* o = new Lbfgs();
  LbfgsImpl* im = new LbfgsImpl();
  o->arena = im;
  o->plan = plan;
  o->m = maxcor;
  o->bounded = bounded;
  o->lo = lower;
  o->hi = upper;
  const int B = plan->B;
  im->B = B;
  Vecs& v = im->v;
  v.n = static_cast<int>(n);
  v.m = maxcor;
  v.nblk = static_cast<int>((n + kChunk - 1) / kChunk);
  v.bounded = bounded;
  v.lo = lower;
  v.hi = upper;
  const size_t vec = static_cast<size_t>(B) * n;
  const size_t total = vec * 5 + vec * maxcor * 2 + B + static_cast<size_t>(B) * v.nblk * (kNP + 1) + B;
  cudaError_t e = cudaMalloc(&im->base, total * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(im->base, 0, total * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&v.ctrl, sizeof(Ctrl) * B + 3 * sizeof(int) * B);
  if (e == cudaSuccess) e = cudaMemset(v.ctrl, 0, sizeof(Ctrl) * B + 3 * sizeof(int) * B);
  if (e == cudaSuccess) e = cudaMallocHost(&im->pinned_i, 3 * sizeof(int) * B);
  if (e == cudaSuccess) e = cudaMallocHost(&im->pinned_f, (vec + B) * sizeof(float));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&im->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&im->ev_in, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&im->ev_out, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    delete o;
    return cuda_fail(e, "lbfgs_create allocation");
  }
  float* p = im->base;
  v.x = p; p += vec;
  v.xt = p; p += vec;
  v.g = p; p += vec;
  v.gt = p; p += vec;
  v.d = p; p += vec;
  v.S = p; p += vec * maxcor;
  v.Y = p; p += vec * maxcor;
  v.ft = p; p += B;
  v.part = p; p += static_cast<size_t>(B) * v.nblk * kNP;
  v.part2 = p; p += static_cast<size_t>(B) * v.nblk;
  im->fun_dev = p; p += B;
  im->status_dev = reinterpret_cast<int*>(v.ctrl + B);
  *out = o;
  return STX_OK;
}

static int launch_cycle(Lbfgs* o, cudaStream_t stream) {
  LbfgsImpl* im = impl_of(o);
  Vecs& v = im->v;
  STX_TRY(o->plan->step(v.xt, o->is_preimage, v.ft, v.gt, nullptr, stream));
  dim3 grid(v.nblk, im->B);
  STX_TRY(launch_pdl(lbfgs_dots_kernel, grid, dim3(kThreads), 0, stream, v));
  STX_TRY(launch_pdl(lbfgs_decide_kernel, dim3(im->B), dim3(512), 0, stream, v));
  STX_TRY(launch_pdl(lbfgs_update_kernel, grid, dim3(kThreads), 0, stream, v));
  count_launch(3);
  return STX_OK;
}

static int fence_in(LbfgsImpl* im, cudaStream_t caller) {
  STX_CUDA(cudaEventRecord(im->ev_in, caller));
  STX_CUDA(cudaStreamWaitEvent(im->stream, im->ev_in, 0));
  return STX_OK;
}
static int fence_out(LbfgsImpl* im, cudaStream_t caller) {
  STX_CUDA(cudaEventRecord(im->ev_out, im->stream));
  STX_CUDA(cudaStreamWaitEvent(caller, im->ev_out, 0));
  return STX_OK;
}

int Lbfgs::reset(const float* x0, int is_preimage_, cudaStream_t caller) {
  LbfgsImpl* im = impl_of(this);
  if (!plan->have_target) return fail(STX_ERR_STATE, "lbfgs: no target Grams set on the plan");
  is_preimage = is_preimage_;
  Vecs& v = im->v;
  STX_TRY(fence_in(im, caller));
  dim3 grid((v.n + 255) / 256, im->B);
  lbfgs_init_kernel<<<grid, 256, 0, im->stream>>>(v, x0, 0);
  STX_CUDA(cudaGetLastError());
  STX_CUDA(cudaMemsetAsync(v.part2, 0, static_cast<size_t>(im->B) * v.nblk * sizeof(float), im->stream));
  STX_TRY(fence_out(im, caller));
  im->started = true;
  return STX_OK;
}

int Lbfgs::run(int maxiter, int max_evals, cudaStream_t caller) {
  LbfgsImpl* im = impl_of(this);
  if (!im->started) return fail(STX_ERR_STATE, "lbfgs: run before reset");
  STX_REQUIRE(maxiter >= 0, "lbfgs: negative maxiter");
  Vecs& v = im->v;
  const int B = im->B;
  cudaStream_t stream = im->stream;
  if (max_evals <= 0) max_evals = 3 * maxiter + 1;
  STX_TRY(fence_in(im, caller));
  lbfgs_set_maxiter_kernel<<<(B + 127) / 128, 128, 0, stream>>>(v, maxiter, B);
  STX_CUDA(cudaGetLastError());

  if (!im->graph || im->graph_preimage != is_preimage) {
    // First use: one eager cycle (sets the kernels' shared-memory attributes, which must not happen inside a
    // capture), then capture the cycle once.
    if (im->graph) {
      cudaGraphExecDestroy(im->graph);
      im->graph = nullptr;
    }
    STX_TRY(launch_cycle(this, stream));
    max_evals -= 1;
    cudaGraph_t g = nullptr;
    STX_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    const long long before = launch_count();
    const int rc = launch_cycle(this, stream);
    im->graph_nodes = launch_count() - before;
    count_launch(-im->graph_nodes);   // captured, not launched
    const cudaError_t e = cudaStreamEndCapture(stream, &g);
    if (rc != STX_OK) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture(lbfgs cycle)");
    const cudaError_t e2 = cudaGraphInstantiate(&im->graph, g, 0);
    cudaGraphDestroy(g);
    if (e2 != cudaSuccess) return cuda_fail(e2, "cudaGraphInstantiate(lbfgs cycle)");
    im->graph_preimage = is_preimage;
  }
  // Cycles are launched in chunks sized by the iterations still missing; the only host round trips are the
  // chunk-boundary reads of the per-problem counters (a handful per run).
  int evals = 0;
  while (evals < max_evals) {
    lbfgs_export_kernel<<<(B + 127) / 128, 128, 0, stream>>>(v, nullptr, im->status_dev, im->status_dev + B,
                                                             im->status_dev + 2 * B, B);
    STX_CUDA(cudaMemcpyAsync(im->pinned_i, im->status_dev, 3 * sizeof(int) * B, cudaMemcpyDeviceToHost, stream));
    STX_CUDA(cudaStreamSynchronize(stream));
    int min_nit = maxiter;
    bool all_done = true;
    for (int b = 0; b < B; ++b)
      if (im->pinned_i[2 * B + b] != kDone) {
        all_done = false;
        min_nit = std::min(min_nit, im->pinned_i[b]);
      }
    if (all_done) break;
    int chunk = std::max(1, maxiter - min_nit);
    chunk = std::min(chunk, max_evals - evals);
    if (g_lbfgs_no_graph) {
      // plain stream launches: keeps the programmatic dependent launches between the kernels of a cycle (the host
      // enqueues a cycle in ~0.1 ms, well ahead of the device)
      for (int i = 0; i < chunk; ++i) STX_TRY(launch_cycle(this, stream));
    } else {
      for (int i = 0; i < chunk; ++i) STX_CUDA(cudaGraphLaunch(im->graph, stream));
      count_launch(im->graph_nodes * chunk);
    }
    evals += chunk;
  }
  STX_TRY(fence_out(im, caller));
  return STX_OK;
}

int Lbfgs::result(float* x, float* fun, int* nit, int* nfev, cudaStream_t caller) {
  LbfgsImpl* im = impl_of(this);
  Vecs& v = im->v;
  const int B = im->B;
  STX_TRY(fence_in(im, caller));
  if (x)
    STX_CUDA(cudaMemcpyAsync(x, v.x, static_cast<size_t>(B) * v.n * sizeof(float), cudaMemcpyDeviceToDevice,
                             im->stream));
  lbfgs_export_kernel<<<(B + 127) / 128, 128, 0, im->stream>>>(v, fun, nit, nfev, nullptr, B);
  STX_CUDA(cudaGetLastError());
  STX_TRY(fence_out(im, caller));
  return STX_OK;
}

int Lbfgs::minimize_host(const float* x0_host, int is_preimage_, int maxiter, float* x_host, float* fun_host,
                         int* nit_host, int* nfev_host, cudaStream_t caller) {
  LbfgsImpl* im = impl_of(this);
  Vecs& v = im->v;
  const int B = im->B;
  const size_t vec = static_cast<size_t>(B) * v.n;
  memcpy(im->pinned_f, x0_host, vec * sizeof(float));
  // gt doubles as the device staging buffer for x0 (it is overwritten by the first evaluation)
  STX_CUDA(cudaMemcpyAsync(v.gt, im->pinned_f, vec * sizeof(float), cudaMemcpyHostToDevice, caller));
  STX_TRY(reset(v.gt, is_preimage_, caller));
  STX_TRY(run(maxiter, 0, caller));
  STX_TRY(result(nullptr, im->fun_dev, im->status_dev, im->status_dev + B, caller));
  STX_CUDA(cudaMemcpyAsync(im->pinned_f, v.x, vec * sizeof(float), cudaMemcpyDeviceToHost, caller));
  STX_CUDA(cudaMemcpyAsync(im->pinned_f + vec, im->fun_dev, B * sizeof(float), cudaMemcpyDeviceToHost, caller));
  STX_CUDA(cudaMemcpyAsync(im->pinned_i, im->status_dev, 2 * sizeof(int) * B, cudaMemcpyDeviceToHost, caller));
  STX_CUDA(cudaStreamSynchronize(caller));
  memcpy(x_host, im->pinned_f, vec * sizeof(float));
  memcpy(fun_host, im->pinned_f + vec, B * sizeof(float));
  if (nit_host) memcpy(nit_host, im->pinned_i, B * sizeof(int));
  if (nfev_host) memcpy(nfev_host, im->pinned_i + B, B * sizeof(int));
  return STX_OK;
}

}  // namespace stx
