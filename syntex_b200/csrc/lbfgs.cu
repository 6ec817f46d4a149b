// Device-resident L-BFGS-B: the optimiser loop of gatys/synthesizer.py:127-151 (scipy.optimize.minimize,
// method "L-BFGS-B", bounds [0,1], maxcor 20, ftol = gtol = 0) as a state machine that advances by exactly ONE f/g
// evaluation per cycle and never returns to the host inside a cycle. oracle/lbfgsb_model.py (LbfgsB with the DEVICE
// options) is the numpy model of this file; it is pinned against scipy itself in tests/test_lbfgsb_model.py.
//
// Algorithm (Byrd, Lu, Nocedal, Zhu 1995; Morales, Nocedal 2011), one independent problem per image of the batch:
//   active set   A = variables at a bound whose gradient points outwards; F = the rest
//   step         subspace minimisation by the direct primal method over F:  d_F = -(Z'BZ)^-1 g_F with the compact
//                L-BFGS matrix B = theta I - W M W', W = [Y, theta S] - the inverse of the REDUCED Hessian, by
//                Sherman-Morrison-Woodbury on the 2k x 2k matrix K = M^-1 - W_F'W_F / theta - then projected onto the
//                box (version 3.0): z = clip(x + d), direction = z - x. Without history: z = clip(x - g).
//   line search  More'-Thuente dcsrch along the feasible segment with lnsrlb's constants (ftol 1e-3, gtol 0.9,
//                xtol 0.1, first trial 1, at most the largest feasible step, <= 20 evaluations); dcsrch's WARNING
//                outcomes end the search with the current point accepted, as lnsrlb does; on failure the memory is
//                dropped and the iteration restarts; the run ends when f does not decrease (factr = 0).
// The one part of L-BFGS-B NOT here is the breakpoint walk of the generalized Cauchy point (a serial scan over up to n
// breakpoints): the subspace step starts from x itself. Measured on the texture objective over 8 seeds, 100 iterations
// (tests/test_lbfgsb_model.py): final loss / scipy's = 1.008 (geometric mean) with or without the walk,
// against 1.08-1.17 for the projected L-BFGS of round 1, whose P B^-1 P step is what cost the 5 % bar.
//
//   cycle = [ plan.step(xt) -> ft, gt ] -> dots kernel -> active kernel -> decide kernel -> update kernel
//
//   dots   : one pass over x, xt, g, gt, d and the (s, y) history (fp16 with one power-of-two scale per vector, see
//            hist_t): every inner product the decision needs as per-block partial sums (packed warp-shuffle +
//            shared-memory reductions, fixed order => bit-reproducible).
//   active : W_A'W_A over the active set of the trial point (about 1 % of the variables): each block compacts its
//            slice of the variables in index order and accumulates the (2k+2)^2 Gram matrix of the gathered rows
//            in fp64. W_F'W_F = W'W - W_A'W_A then needs no masked pass over the history.
//   decide : one CTA per problem: sums the partials in fp64, advances the line search, and on acceptance updates
//            S'Y / S'S / Y'Y, builds K, solves it (Gauss-Jordan in natural order - K is quasi-definite - the whole
//            block, one barrier per step) and leaves the coefficients of the new step; a job that finishes clears
//            its image's switch in the plan, so the f/g kernels skip it while the rest of the batch runs on.
//   update : applies the decision: stores the new pair, moves x, builds the step in one pass over the history,
//            projects it, and writes the next trial point; partial sums of g'd and of the largest feasible step.
#include "lbfgs.h"

#include <cuda_fp16.h>

#include "ptx.cuh"

namespace stx {
int g_lbfgs_no_graph = 0;   // 1 = launch every optimiser cycle on the stream instead of replaying a CUDA graph (key 10)
int g_lbfgs_serial_active = 0;   // 1 = active-set kernel in the main stream after the dots kernel (key 14)
int g_lbfgs_no_skip = 0;         // 1 = finished jobs keep being evaluated until the whole batch is done (key 17)
}

#include <algorithm>
#include <thread>
#include <vector>

namespace stx {

namespace {

constexpr int kMaxM = 32;
constexpr int kVec = 4;                     // granularity of the problem size and of the active kernel's loads
constexpr int kPer = 8;                     // elements per thread in the dots / update kernels: one 16-byte load of fp16 history
constexpr int kThreads = 256;
constexpr int kChunk = kThreads * kPer;     // elements per block
constexpr int kNP0 = 8;                     // scalar partial sums per block (dots kernel)
constexpr int kNPH = 6;                     // partial sums per history pair
constexpr int kNP = kNP0 + kNPH * kMaxM;
constexpr int kQ = 2 * kMaxM + 2;           // rows of the active-set Gram: Y_0.. | S_0.. | y_new | s_new (chronological)
constexpr int kSlices = 8;                  // active kernel: blocks per problem
constexpr int kActCap = 2048;               // compacted indices held per round
constexpr int kActRows = 96;                // rows gathered per tile
constexpr int kActPairs = (kQ * (kQ + 1) / 2 + kThreads - 1) / kThreads;
constexpr int kDecideThreads = 512;
constexpr int kKN = 2 * kMaxM;              // largest reduced system

constexpr double FTOL = 1e-3, GTOL = 0.9, XTOL = 0.1, XTRAPL = 1.1, XTRAPU = 4.0, STPBIG = 1e10, EPSMCH = 2.2e-16;
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
  float theta_inv;    // 1 / theta
  float cs[kMaxM], cy[kMaxM];   // step = -g / theta + sum_j cs[j] S_j + cy[j] Y_j over F (j chronological); the
                                // coefficients carry the vectors' storage scales (they multiply the stored halves)
  float hs[kMaxM], hy[kMaxM];   // storage scales of the history vectors BY SLOT: s = hs * stored, y = hy * stored
  double f, pp, stp0, theta, stpmx;
  LsState ls;
};

// S'Y, S'S, Y'Y by history SLOT (row / column of pair i = its slot (head + i) % m), like the vectors themselves:
// dropping the oldest pair is a head increment, not a shift. SY[a][b] = s_a . y_b (both triangles are used).
struct Tables {
  double SY[kMaxM * kMaxM], SS[kMaxM * kMaxM], YY[kMaxM * kMaxM];
};

// The (s, y) history is kept as fp16 with one power-of-two scale per vector (Ctrl::hs / hy): the two passes over it
// are the optimiser's HBM time (2 x 2 GB per cycle in fp32 at 64 jobs of 256^2 with 20 pairs). The scale maps the
// vector's 2-norm just below 2^15, so no element overflows whatever their distribution, and elements down to 2^-29
// of the norm keep 11 significant bits: a relative rounding of 2^-12 per element, an order of magnitude below the
// noise the bf16 backward leaves in the gradient itself (4e-3). The S'Y / S'S / Y'Y tables are built from the pair as
// measured (fp32 differences) against the older vectors as stored; oracle/lbfgsb_model.py (hist = "f16") does the same.
typedef __half hist_t;

struct Vecs {
  float *x, *xt, *g, *gt, *d;           // [B][n]
  hist_t *S, *Y;                        // history [B][m][n]
  float* ft;                             // [B]
  float* part;                           // [B][kNP][nblk]
  float* part2;                          // [B][nblk]  g'd of a freshly built direction
  float* part3;                          // [B][nblk]  largest feasible step along it
  double* gramA;                         // [B][kSlices][kQ][kQ]
  Ctrl* ctrl;                            // [B]
  Tables* tab;                           // [B]
  int* active;                           // [B] the plan's per-image switch (null: finished jobs keep being evaluated)
  long long* dbg;                        // [16] clock64 stamps of job 0's decide kernel (diagnostic, stx_lbfgs_phase_cycles)
  int n, m, nblk, bounded;
  int hn;                                // elements between history vectors: n rounded up to 8 (16-byte aligned rows)
  float lo, hi;
};

__device__ __forceinline__ float clipf(float z, float lo, float hi, int bounded) {
  return bounded ? fminf(fmaxf(z, lo), hi) : z;
}
// cauchy's iwhere = 1 / 2: at a bound with the gradient pushing outwards (or zero)
__device__ __forceinline__ bool is_free(float x, float g, float lo, float hi, int bounded) {
  return !bounded || !((x <= lo && g >= 0.0f) || (x >= hi && g <= 0.0f));
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// eight fp32 elements of a vector, the two halves guarded separately (n is a multiple of 4, not of 8)
__device__ __forceinline__ void ld8(const float* p, bool in_a, bool in_b, float (&v)[8]) {
  const float4 a = in_a ? ld4(p) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b = in_b ? ld4(p + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, bool in_a, bool in_b, const float (&v)[8]) {
  if (in_a) st4(p, make_float4(v[0], v[1], v[2], v[3]));
  if (in_b) st4(p + 4, make_float4(v[4], v[5], v[6], v[7]));
}
__device__ __forceinline__ float dot8(const float (&a)[8], const float (&b)[8]) {
  float t = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) t = fmaf(a[e], b[e], t);
  return t;
}
// eight stored history elements (unscaled), one 16-byte load; rows are padded to 8 elements, so the load is always in
// range and the elements past n are zero (never written)
__device__ __forceinline__ void ldh8(const hist_t* p, float (&v)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
    v[2 * e] = f.x;
    v[2 * e + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 ldh8_raw(const hist_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void cvth8(const uint4& raw, float (&v)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
    v[2 * e] = f.x;
    v[2 * e + 1] = f.y;
  }
}
constexpr int kHistTrip = 4;     // history pairs per loop trip: 2 x kHistTrip 16-byte loads in flight per thread
// v * inv_scale rounded to storage; v becomes what was stored (unscaled) so that the caller can keep using it
__device__ __forceinline__ void sth8(hist_t* p, float (&v)[8], float inv_scale) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 h = __floats2half2_rn(v[2 * e] * inv_scale, v[2 * e + 1] * inv_scale);
    w[e] = *reinterpret_cast<const uint32_t*>(&h);
    const float2 f = __half22float2(h);
    v[2 * e] = f.x;
    v[2 * e + 1] = f.y;
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
// power of two that maps a vector of squared 2-norm `norm2` below 2^15 in every element (lbfgsb_model.f16_scale)
__device__ __forceinline__ float hist_scale(double norm2) {
  const double nrm = sqrt(norm2);
  if (!(nrm > 0.0) || !isfinite(nrm)) return 1.f;
  int e;
  frexp(nrm, &e);
  return ldexpf(1.f, e - 15);
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sums of EIGHT quantities over the warp with 9 shuffles instead of 8 x 5: at every step each lane hands half of its
// quantities to its partner and keeps (and accumulates) the other half. Lane 4q ends with the sum of quantity q.
// Fixed order => bit-reproducible.
__device__ __forceinline__ float warp_sum8(const float (&a)[8], int lane) {
  float b[4], c[2];
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = u16 ? a[i] : a[i + 4], keep = u16 ? a[i + 4] : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = u8 ? b[i] : b[i + 2], keep = u8 ? b[i + 2] : b[i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const float send = u4 ? c[0] : c[1], keep = u4 ? c[1] : c[0];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  d += __shfl_xor_sync(0xffffffffu, d, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;     // lane l (l % 4 == 0): quantity 4 * bit4 + 2 * bit3 + bit2 of l
}
__device__ __forceinline__ int warp_sum8_quantity(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }

// ------------------------------------------------------------------------------------------------ dots
// scalars: 0 gt'd  1 s'y  2 y'y  3 Pg'Pg  4 s'Pg  5 y'Pg  6 s's  7 -
// per pair j: S_j'y  Y_j'y  S_j'Pg  Y_j'Pg  S_j's  Y_j's        (s = xt - x, y = gt - g, Pg = gt on F of the trial point)
__global__ void __launch_bounds__(kThreads, 4) lbfgs_dots_kernel(Vecs v) {
  pdl_launch_dependents();   // programmatic dependent launch: see launch_pdl (common.h)
  pdl_wait();
  const int b = blockIdx.y;
  const Ctrl& c = v.ctrl[b];
  if (c.mode == kDone) return;
  __shared__ float red[kThreads / 32][kNP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * kChunk + threadIdx.x * kPer;
  const bool in_a = i0 < v.n, in_b = i0 + 4 < v.n;
  const size_t off = static_cast<size_t>(b) * v.n + i0;
  float x[8], xt[8], g[8], gt[8], d[8];
  ld8(v.x + off, in_a, in_b, x);
  ld8(v.xt + off, in_a, in_b, xt);
  ld8(v.g + off, in_a, in_b, g);
  ld8(v.gt + off, in_a, in_b, gt);
  ld8(v.d + off, in_a, in_b, d);
  const float lo = v.lo, hi = v.hi;
  const int bd = v.bounded;

  float s[8], y[8], pg[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const bool in = e < 4 ? in_a : in_b;
    s[e] = xt[e] - x[e];
    y[e] = gt[e] - g[e];
    pg[e] = (in && is_free(xt[e], gt[e], lo, hi, bd)) ? gt[e] : 0.f;
  }
  static_assert(kNP0 == 8 && kNPH <= 8, "warp_sum8 carries eight quantities");
  const int myq = warp_sum8_quantity(lane);
  {
    const float vals[8] = {dot8(gt, d), dot8(s, y), dot8(y, y), dot8(pg, pg), dot8(s, pg), dot8(y, pg), dot8(s, s), 0.f};
    const float r = warp_sum8(vals, lane);
    if ((lane & 3) == 0) red[warp][myq] = r;
  }
  const int nh = c.nh, head = c.head, m = v.m;
  // kHistTrip pairs per trip: all their 16-byte loads are issued (and stay raw fp16, 4 registers each) before the first
  // conversion - the kernel is bound by the bytes in flight per SM (measured: 2.6 TB/s with two 8-byte loads per trip)
  for (int j = 0; j < nh; j += kHistTrip) {
    uint4 rs[kHistTrip], ry[kHistTrip];
#pragma unroll
    for (int q = 0; q < kHistTrip; ++q) {
      rs[q] = ry[q] = make_uint4(0u, 0u, 0u, 0u);
      if (in_a && j + q < nh) {
        const size_t h = (static_cast<size_t>(b) * m + (head + j + q) % m) * v.hn + i0;
        rs[q] = ldh8_raw(v.S + h);
        ry[q] = ldh8_raw(v.Y + h);
      }
    }
#pragma unroll
    for (int q = 0; q < kHistTrip; ++q) {
      if (j + q >= nh) break;
      float sj[8], yj[8];
      cvth8(rs[q], sj);
      cvth8(ry[q], yj);
      const float vals[8] = {dot8(sj, y), dot8(yj, y), dot8(sj, pg), dot8(yj, pg), dot8(sj, s), dot8(yj, s), 0.f, 0.f};
      const float r = warp_sum8(vals, lane);
      if ((lane & 3) == 0 && myq < kNPH) {
        const int slot = (head + j + q) % m;
        // even quantities are products with S_j, odd ones with Y_j; the scales are powers of two: exact
        red[warp][kNP0 + kNPH * (j + q) + myq] = r * ((myq & 1) ? c.hy[slot] : c.hs[slot]);
      }
    }
  }
  __syncthreads();
  const int np = kNP0 + kNPH * nh;
  // quantity-major: the decision kernel reads one quantity of all blocks as consecutive floats
  float* dst = v.part + static_cast<size_t>(b) * v.nblk * kNP + blockIdx.x;
  for (int k = threadIdx.x; k < np; k += kThreads) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += red[w][k];
    dst[static_cast<size_t>(k) * v.nblk] = t;
  }
}

// ------------------------------------------------------------------------------------------------ active set
// Gram matrix of the history rows of the ACTIVE variables of the trial point (xt, gt): for i in A the row is
//   u_i = (Y_0[i] .. Y_{nh-1}[i], S_0[i] .. S_{nh-1}[i], y_new[i], s_new[i])       (chronological pair order)
// and gramA[slice] = sum over the slice's active i of u_i u_i' in fp64.
// Each block owns a slice of the variables and each thread a CONTIGUOUS run of it, so index order = thread order:
// all loads of a round are issued up front (the flags of up to kActPerThread variables live in registers), ONE block
// scan of the per-thread counts places the indices, and the rows are gathered kActRows at a time. The summation
// order depends only on the data (deterministic, independent of the batch).
constexpr int kActGatherUnroll = 8;
constexpr int kActPerThread = 128;                 // variables per thread per round (4 words of flag bits)
constexpr int kActRound = kThreads * kActPerThread;

__global__ void __launch_bounds__(kThreads) lbfgs_active_kernel(Vecs v) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const Ctrl& c = v.ctrl[b];
  if (c.mode == kDone) return;
  const int nh = c.nh, head = c.head, m = v.m;
  const int Q = 2 * nh + 2;
  const int npairs = Q * (Q + 1) / 2;
  double* out = v.gramA + (static_cast<size_t>(b) * kSlices + blockIdx.x) * kQ * kQ;
  if (!v.bounded) {   // no bounds: A is empty
    for (int p = threadIdx.x; p < Q * Q; p += kThreads) out[(p / Q) * kQ + (p % Q)] = 0.0;
    return;
  }
  __shared__ int list[kActCap];
  __shared__ float u[kActRows][kQ + 1];
  __shared__ int warp_tot[kThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this thread's (a, b) pairs, a <= b
  int pa[kActPairs], pb[kActPairs];
  double acc[kActPairs];
#pragma unroll
  for (int t = 0; t < kActPairs; ++t) {
    acc[t] = 0.0;
    int p = threadIdx.x + t * kThreads;
    pa[t] = -1;
    pb[t] = 0;
    if (p < npairs) {
      int a = 0;
      while (p >= Q - a) {   // row a holds Q - a pairs
        p -= Q - a;
        ++a;
      }
      pa[t] = a;
      pb[t] = a + p;
    }
  }
  const float lo = v.lo, hi = v.hi;
  const size_t base = static_cast<size_t>(b) * v.n;
  // slice bounds, multiples of 4
  const int per = ((v.n / kVec + kSlices - 1) / kSlices) * kVec;
  const int i_begin = blockIdx.x * per;
  const int i_end = min(v.n, i_begin + per);

  // Gather kActRows rows at a time, then accumulate. Gather: thread (rg, a) = (tid / Q, tid % Q) takes column a of rows
  // rg, rg + G, ...: kActGatherUnroll independent scattered loads in flight per thread (the history of all jobs is
  // far larger than L2: every one of them is a DRAM round trip). Accumulate: fp32 products summed per tile in fp32
  // (<= kActRows terms, four independent chains per thread), tiles added in fp64.
  const int G = kThreads / Q;                              // row groups (Q <= 66 < kThreads)
  const int ga = threadIdx.x % Q, grg = threadIdx.x / Q;
  const hist_t* gsrc;                                      // column ga of a row = gscale * gsrc[i] (history) or a difference
  float gscale = 1.f;
  {
    if (ga < nh) {
      gsrc = v.Y + (static_cast<size_t>(b) * m + (head + ga) % m) * v.hn;
      gscale = c.hy[(head + ga) % m];
    } else if (ga < 2 * nh) {
      gsrc = v.S + (static_cast<size_t>(b) * m + (head + ga - nh) % m) * v.hn;
      gscale = c.hs[(head + ga - nh) % m];
    } else {
      gsrc = nullptr;
    }
  }
  auto flush = [&](int count) {
    for (int r0 = 0; r0 < count; r0 += kActRows) {
      const int rows = min(kActRows, count - r0);
      if (grg < G) {
        for (int rb = grg; rb < rows; rb += G * kActGatherUnroll) {
          float val[kActGatherUnroll];
#pragma unroll
          for (int q = 0; q < kActGatherUnroll; ++q) {
            const int r = rb + q * G;
            val[q] = 0.f;
            if (r < rows) {
              const int i = list[r0 + r];
              if (gsrc) val[q] = __half2float(__ldg(gsrc + i)) * gscale;
              else if (ga == 2 * nh) val[q] = v.gt[base + i] - v.g[base + i];
              else val[q] = v.xt[base + i] - v.x[base + i];
            }
          }
#pragma unroll
          for (int q = 0; q < kActGatherUnroll; ++q) {
            const int r = rb + q * G;
            if (r < rows) u[r][ga] = val[q];
          }
        }
      }
      __syncthreads();
      float part[kActPairs];
#pragma unroll
      for (int t = 0; t < kActPairs; ++t) part[t] = 0.f;
      for (int r = 0; r < rows; ++r) {
#pragma unroll
        for (int t = 0; t < kActPairs; ++t)
          if (pa[t] >= 0) part[t] = fmaf(u[r][pa[t]], u[r][pb[t]], part[t]);
      }
#pragma unroll
      for (int t = 0; t < kActPairs; ++t) acc[t] += static_cast<double>(part[t]);
      __syncthreads();
    }
  };

  for (int r_begin = i_begin; r_begin < i_end; r_begin += kActRound) {
    const int r_end = min(i_end, r_begin + kActRound);
    // contiguous run of this thread inside the round, a multiple of 4 variables
    const int run = (((r_end - r_begin) / kVec + kThreads - 1) / kThreads) * kVec;
    const int t_begin = r_begin + threadIdx.x * run;
    const int t_end = min(r_end, t_begin + run);
    uint32_t bits[kActPerThread / 32] = {0u, 0u, 0u, 0u};
    int mine = 0;
#pragma unroll
    for (int q = 0; q < kActPerThread / kVec; ++q) {
      const int i0 = t_begin + q * kVec;
      if (i0 < t_end) {
        const float4 xt = ld4(v.xt + base + i0), gt = ld4(v.gt + base + i0);
        const uint32_t f = (is_free(xt.x, gt.x, lo, hi, 1) ? 0u : 1u) | (is_free(xt.y, gt.y, lo, hi, 1) ? 0u : 2u) |
                           (is_free(xt.z, gt.z, lo, hi, 1) ? 0u : 4u) | (is_free(xt.w, gt.w, lo, hi, 1) ? 0u : 8u);
        bits[q / 8] |= f << (4 * (q % 8));
        mine += __popc(f);
      }
    }
    // exclusive prefix of the per-thread counts over the block (thread order = index order)
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      if (w < warp) before += warp_tot[w];
      total += warp_tot[w];
    }
    // place the indices; when the round holds more than the list, consume it in passes of kActCap entries
    const int first = before + incl - mine;            // this thread's first position in the round's list
    for (int pass0 = 0; pass0 < total; pass0 += kActCap) {
      int pos = first;
#pragma unroll
      for (int w = 0; w < kActPerThread / 32; ++w) {
        uint32_t bw = bits[w];
        while (bw) {
          const int bit = __ffs(bw) - 1;
          bw &= bw - 1;
          if (pos >= pass0 && pos < pass0 + kActCap) list[pos - pass0] = t_begin + w * 32 + bit;
          ++pos;
        }
      }
      __syncthreads();
      flush(min(kActCap, total - pass0));
    }
    __syncthreads();      // warp_tot is rewritten by the next round
  }
#pragma unroll
  for (int t = 0; t < kActPairs; ++t) {
    if (pa[t] >= 0) {
      out[pa[t] * kQ + pb[t]] = acc[t];
      out[pb[t] * kQ + pa[t]] = acc[t];
    }
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
  st.width = STPBIG; st.width1 = 2.0 * STPBIG;
  st.stx = 0.0; st.fx = f0; st.dx = g0; st.sty = 0.0; st.fy = f0; st.dy = g0;
  st.stmin = 0.0; st.stmax = stp + XTRAPU * stp; st.stp = stp; st.nev = 0;
}

// dcsrch as lnsrlb drives it (stpmin 0, stpmax = the largest feasible step).
// 0 = the search ends with the current point accepted (CONVERGENCE, or one of dcsrch's WARNINGs: lnsrlb takes both as
// the new iterate), 1 = evaluation budget exhausted, 2 = evaluate again at st.stp
__device__ int ls_update(LsState& st, double f, double g, double stpmx) {
  const double stp = st.stp;
  st.nev += 1;
  const double ftest = st.finit + stp * st.gtest;
  if (st.stage == 1 && f <= ftest && g >= 0.0) st.stage = 2;
  bool warn = false;
  if (st.brackt && (stp <= st.stmin || stp >= st.stmax)) warn = true;
  if (st.brackt && st.stmax - st.stmin <= XTOL * st.stmax) warn = true;
  if (stp == stpmx && f <= ftest && g <= st.gtest) warn = true;
  if (stp == 0.0 && (f > ftest || g >= st.gtest)) warn = true;
  if (f <= ftest && fabs(g) <= GTOL * (-st.ginit)) return 0;
  if (warn) return 0;
  if (st.nev >= MAX_LS) return 1;
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
  nw = fmin(stpmx, fmax(0.0, nw));
  if (st.brackt && (nw <= st.stmin || nw >= st.stmax || st.stmax - st.stmin <= XTOL * st.stmax)) nw = st.stx;
  st.stp = nw;
  return 2;
}

// ------------------------------------------------------------------------------------------------ decide
__device__ __forceinline__ int slot_at(int head, int i, int m) {
  int s = head + i;
  return s >= m ? s - m : s;
}

// What thread 0 decides before the block-wide linear algebra.
struct Decision {
  int solve;      // 1 = accepted with history: update the tables, build and solve K
  int push;       // the new pair enters the history
  int shift;      // 1 = the oldest pair was dropped (chronological index i of the new numbering = i + 1 of the old)
  int k;          // pairs after the update
  int nh_old;     // pairs the dots / active kernels saw
};

// The serial part of a cycle's decision: line search and bookkeeping, run by ONE thread on a shared-memory copy of
// the job's control block.
__device__ void decide_serial(Ctrl& c, const double* sums, double dg0, double stpmx_min, const Vecs& v, int b,
                              Decision& dec) {
  const double ft = static_cast<double>(v.ft[b]);
  const double dphi = sums[0], ys = sums[1], yy = sums[2], ppn = sums[3];
  const int m = v.m;
  c.nfev += 1;
  c.push = 0;
  dec.solve = 0;
  dec.push = 0;
  dec.shift = 0;
  dec.k = c.nh;
  dec.nh_old = c.nh;

  // lnsrlb: first trial 1, except at the very first iteration of an unconstrained problem
  auto first_step = [&](double pp) {
    return (c.nit == 0 && !v.bounded && pp > 0.0) ? fmin(1.0 / sqrt(pp), STPBIG) : 1.0;
  };

  if (c.mode == kStart) {
    c.f = ft;
    c.pp = ppn;
    c.nh = 0;
    c.head = 0;
    c.theta = 1.0;
    c.theta_inv = 1.f;
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
      res = 1;                       // not a descent direction (lnsrlb info = -4): treated as a failed search
    } else {
      // largest feasible step (lnsrlb): 1 at the first iteration of a constrained problem; d = z - x is feasible at
      // step 1 by construction, so rounding must not push the bound below it
      c.stpmx = !v.bounded ? STPBIG : (c.nit == 0 ? 1.0 : fmax(fmin(stpmx_min, STPBIG), 1.0));
      ls_init(c.ls, c.f, dg0, c.stp0);
      c.fresh = 0;
      res = ls_update(c.ls, ft, dphi, c.stpmx);
    }
  } else {
    res = ls_update(c.ls, ft, dphi, c.stpmx);
  }

  if (res == 0) {
    int k = c.nh;
    // matupd's curvature test: skip the pair unless s'y > epsmch * (-g'd * stp)
    const double ddum = -c.ls.ginit * c.ls.stp;
    const bool push = ys > EPSMCH * ddum && ys > 0.0 && yy > 0.0;
    if (push) {
      if (k == m) {                  // drop the oldest pair: the tables are indexed by slot, nothing moves
        c.head = slot_at(c.head, 1, m);
        k -= 1;
        dec.shift = 1;
      }
      c.slot_new = slot_at(c.head, k, m);
      c.hs[c.slot_new] = hist_scale(sums[6]);
      c.hy[c.slot_new] = hist_scale(yy);
      c.push = 1;
      dec.push = 1;
      k += 1;
      c.theta = yy / ys;
      c.theta_inv = static_cast<float>(ys / yy);
    }
    const double fold = c.f;
    c.nh = k;
    c.f = ft;
    c.pp = ppn;
    c.nit += 1;
    dec.k = k;
    if (!(ppn > 0.0) || !(fold - ft > 0.0)) {
      // projected gradient vanished, or f did not decrease: mainlb's two stopping tests at pgtol = factr = 0
      c.action = kFinalAccept;
      c.mode = kDone;
      c.fresh = 0;
      return;
    }
    // On the iteration budget the next direction and trial point are still built (one extra vector pass), so a
    // later run() with a larger budget continues the same optimisation.
    if (c.nit >= c.maxiter) c.mode = kDone;
    c.stp0 = 1.0;
    c.stp = 1.f;
    c.fresh = 1;
    c.action = kAccept;
    dec.solve = k > 0 ? 1 : 0;
  } else if (res == 1) {
    if (c.nh > 0) {                  // drop the memory and restart the iteration at x (mainlb: "refresh the memory")
      c.nh = 0;
      c.head = 0;
      c.theta = 1.0;
      c.theta_inv = 1.f;
      c.stp0 = first_step(c.pp);
      c.stp = static_cast<float>(c.stp0);
      c.fresh = 1;
      c.action = kRestart;
    } else {
      c.mode = kDone;                // ABNORMAL_TERMINATION_IN_LNSRCH
      c.fresh = 0;
      c.action = kNone;
    }
  } else {
    c.stp = static_cast<float>(c.ls.stp);
    c.action = kRetry;
  }
}

__global__ void __launch_bounds__(kDecideThreads) lbfgs_decide_kernel(Vecs v) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  Ctrl& cg = v.ctrl[b];
  if (cg.mode == kDone) {
    if (threadIdx.x == 0) cg.action = kNone;
    return;
  }
  __shared__ double sums[kNP + 2];
  __shared__ Ctrl sc;
  __shared__ Decision dec;
  __shared__ double K[kKN][kKN + 1];     // reduced system, last column = right-hand side
  __shared__ double wv[kKN];
  __shared__ int singular;
  static_assert(sizeof(Ctrl) % 4 == 0, "Ctrl is copied word by word");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = kDecideThreads / 32;
  auto stamp = [&](int i) {
    if (b == 0 && tid == 0) v.dbg[i] = clock64();
  };
  stamp(0);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&cg);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sc);
    for (int i = tid; i < static_cast<int>(sizeof(Ctrl) / 4); i += kDecideThreads) dst[i] = src[i];
  }
  const int nh = cg.nh;
  const int np = kNP0 + kNPH * nh;
  // Sum the block partials in a fixed order: one warp per quantity, lane l takes blocks l, l+32, ..., then a
  // butterfly over the lanes. Deterministic: the summation tree depends only on nblk. A warp works on kSumIlp
  // quantities at a time: a dependent chain of fp64 adds and shuffles is ~100 cycles per link on this part, and nine
  // such chains one after the other were a third of the kernel (tests/tools/lbfgs_phases.py).
  {
    constexpr int kSumIlp = 3;
    const float* pbase = v.part + static_cast<size_t>(b) * kNP * v.nblk;
    for (int k0 = warp; k0 < np; k0 += nwarps * kSumIlp) {
      double t[kSumIlp];
#pragma unroll
      for (int q = 0; q < kSumIlp; ++q) {
        t[q] = 0.0;
        const int k = k0 + q * nwarps;
        if (k < np) {
          const float* src = pbase + static_cast<size_t>(k) * v.nblk;
          for (int j = lane; j < v.nblk; j += 32) t[q] += static_cast<double>(src[j]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < kSumIlp; ++q) t[q] += __shfl_xor_sync(0xffffffffu, t[q], o);
      }
#pragma unroll
      for (int q = 0; q < kSumIlp; ++q) {
        const int k = k0 + q * nwarps;
        if (lane == 0 && k < np) sums[k] = t[q];
      }
    }
    if (warp == nwarps - 1) {            // g'd of a freshly built direction
      double t = 0.0;
      const float* src = v.part2 + static_cast<size_t>(b) * v.nblk;
      for (int j = lane; j < v.nblk; j += 32) t += static_cast<double>(src[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) sums[kNP] = t;
    }
    if (warp == nwarps - 2) {            // largest feasible step along it
      float t = 3.0e38f;
      const float* src = v.part3 + static_cast<size_t>(b) * v.nblk;
      for (int j = lane; j < v.nblk; j += 32) t = fminf(t, src[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t = fminf(t, __shfl_xor_sync(0xffffffffu, t, o));
      if (lane == 0) sums[kNP + 1] = static_cast<double>(t);
    }
  }
  __syncthreads();
  stamp(1);                                   // partial sums reduced
  if (tid == 0) {
    singular = 0;
    decide_serial(sc, sums, sums[kNP], sums[kNP + 1], v, b, dec);
  }
  __syncthreads();
  stamp(2);                                   // line search / bookkeeping decided

  if (dec.solve) {
    Tables& T = v.tab[b];
    const int m = v.m, k = dec.k, shift = dec.shift, nho = dec.nh_old, head = sc.head;
    const int kn = dec.push ? k - 1 : k;       // index of the new pair (chronological), if any
    // ---- new row / column of the tables (slot-indexed). Old pair i (new numbering) = partial-sum index i + shift.
    if (dec.push) {
      const int sn = sc.slot_new;
      for (int i = tid; i < kn; i += kDecideThreads) {
        const int si = slot_at(head, i, m);
        const double* q = sums + kNP0 + kNPH * (i + shift);
        T.SY[si * kMaxM + sn] = q[0];      // s_i . y_new
        T.YY[si * kMaxM + sn] = q[1];
        T.YY[sn * kMaxM + si] = q[1];
        T.SS[si * kMaxM + sn] = q[4];
        T.SS[sn * kMaxM + si] = q[4];
        T.SY[sn * kMaxM + si] = q[5];      // s_new . y_i
      }
      if (tid == 0) {
        T.SY[sn * kMaxM + sn] = sums[1];
        T.YY[sn * kMaxM + sn] = sums[2];
        T.SS[sn * kMaxM + sn] = sums[6];
      }
    }
    __syncthreads();
    // ---- K = [[-D - Y_F'Y_F / theta,  L' - Y_F'S_F], [L - S_F'Y_F,  theta S_A'S_A]],  rhs = -[Y'Pg ; theta S'Pg]
    // with X_F'Z_F = X'Z - X_A'Z_A (tables minus the active-set Gram), pairs in chronological order.
    const double theta = sc.theta;
    const int N = 2 * k;
    const double* GA = v.gramA + static_cast<size_t>(b) * kSlices * kQ * kQ;
    auto gram = [&](int ra, int rb) {     // fixed-order sum over the slices
      double t = 0.0;
      for (int sl = 0; sl < kSlices; ++sl) t += GA[(static_cast<size_t>(sl) * kQ + ra) * kQ + rb];
      return t;
    };
    // Gram row of pair i: Y -> qy(i), S -> qs(i) in the numbering the active kernel used
    auto qy = [&](int i) { return (dec.push && i == kn) ? 2 * nho : i + shift; };
    auto qs = [&](int i) { return (dec.push && i == kn) ? 2 * nho + 1 : nho + i + shift; };
    for (int e = tid; e < N * (N + 1); e += kDecideThreads) {
      const int r = e / (N + 1), cc = e % (N + 1);
      double val;
      if (cc == N) {
        // right-hand side: the new pair's dots are the scalars, the old pairs' come from their partial sums
        const int i = r < k ? r : r - k;
        double ypg, spg;
        if (dec.push && i == kn) {
          spg = sums[4];
          ypg = sums[5];
        } else {
          const double* q = sums + kNP0 + kNPH * (i + shift);
          spg = q[2];
          ypg = q[3];
        }
        val = r < k ? -ypg : -theta * spg;
      } else {
        const int i = r < k ? r : r - k, j = cc < k ? cc : cc - k;
        const int si = slot_at(head, i, m), sj = slot_at(head, j, m);
        if (r < k && cc < k) {
          const double yyF = T.YY[si * kMaxM + sj] - gram(qy(i), qy(j));
          val = -yyF / theta - (i == j ? T.SY[si * kMaxM + si] : 0.0);
        } else if (r >= k && cc >= k) {
          val = theta * gram(qs(i), qs(j));
        } else {
          // (S row i, Y column j): L_ij - (S_F'Y_F)_ij, L = strictly lower triangle of S'Y; the other block is its
          // transpose
          const int is = r >= k ? i : j, jy = r >= k ? j : i;
          const int ss = slot_at(head, is, m), sy = slot_at(head, jy, m);
          const double syF = T.SY[ss * kMaxM + sy] - gram(qs(is), qy(jy));
          val = (is > jy ? T.SY[ss * kMaxM + sy] : 0.0) - syF;
        }
      }
      K[r][cc] = val;
    }
    __syncthreads();
    stamp(3);                                 // tables updated, K built
    // ---- Gauss-Jordan elimination in natural order, no pivoting, the whole block per step. K is quasi-definite: the
    // first k pivots are those of -(D + Y_F'Y_F / theta) (negative definite), the last k those of the Schur complement
    // theta S_A'S_A + B (D + Y_F'Y_F / theta)^-1 B' (positive definite) - the block LEL' factorisation formk computes -
    // so the natural order is stable and the pivot search, the row swaps and the back-substitution go away (each was a
    // chain of block barriers and fp64 shuffles: 105k of the kernel's 163k cycles, tests/tools/lbfgs_phases.py).
    // One barrier per step: every thread derives the pivot's reciprocal itself (the same shared-memory word, final
    // since the previous step's barrier), and threads map to (row group, column) without integer divisions.
    static_assert(kDecideThreads == 512 && kKN + 1 <= 2 * 64, "elimination thread map: 8 row groups x 64 columns");
    const int ecol = tid & 63, erow = tid >> 6;
    for (int p = 0; p < N; ++p) {
      const double pv = K[p][p];
      if (!(fabs(pv) > 0.0) || !isfinite(pv)) {     // uniform: every thread reads the same value
        if (tid == 0) singular = 1;
        break;
      }
      const double inv = 1.0 / pv;
      for (int j = p + 1 + ecol; j <= N; j += 64) {   // columns p+1 .. N (the right-hand side included)
        const double pj = K[p][j] * inv;
        for (int r = erow; r < N; r += 8)
          if (r != p) K[r][j] -= K[r][p] * pj;
      }
      __syncthreads();
    }
    __syncthreads();
    stamp(4);                                 // eliminated
    if (!singular)
      for (int i = tid; i < N; i += kDecideThreads) wv[i] = K[i][N] / K[i][i];
    __syncthreads();
    if (tid == 0) {
      bool ok = !singular;
      for (int i = 0; i < N && ok; ++i) ok = isfinite(wv[i]);
      if (ok) {
        // step = r / theta + W_F wv / theta^2,  W = [Y, theta S]
        for (int i = 0; i < k; ++i) {
          const int si = slot_at(head, i, m);
          sc.cy[i] = static_cast<float>(wv[i] / (theta * theta)) * sc.hy[si];
          sc.cs[i] = static_cast<float>(wv[k + i] / theta) * sc.hs[si];
        }
      } else {
        // formk / formt failure in mainlb: refresh the memory; the step falls back to the projected gradient
        sc.nh = 0;
        sc.head = 0;
        sc.push = 0;
        sc.theta = 1.0;
        sc.theta_inv = 1.f;
      }
    }
    __syncthreads();
  }
  stamp(5);                                   // solved
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&sc);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&cg);
    for (int i = tid; i < static_cast<int>(sizeof(Ctrl) / 4); i += kDecideThreads) dst[i] = src[i];
    // a job that has just finished: the f/g kernels skip its image from the next cycle on (Plan::d_active)
    if (tid == 0 && v.active && sc.mode == kDone) v.active[b] = 0;
  }
  stamp(6);
}

// ------------------------------------------------------------------------------------------------ update
__global__ void __launch_bounds__(kThreads, 3) lbfgs_update_kernel(Vecs v) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const Ctrl& c = v.ctrl[b];
  const int action = c.action;
  if (action == kNone) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * kChunk + threadIdx.x * kPer;
  const bool in_a = i0 < v.n, in_b = i0 + 4 < v.n;
  const size_t off = static_cast<size_t>(b) * v.n + i0;
  const float lo = v.lo, hi = v.hi;
  const int bd = v.bounded;
  const float stp = c.stp;

  if (action == kRetry) {
    float x[8], d[8], t[8];
    ld8(v.x + off, in_a, in_b, x);
    ld8(v.d + off, in_a, in_b, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) t[e] = clipf(fmaf(stp, d[e], x[e]), lo, hi, bd);
    st8(v.xt + off, in_a, in_b, t);
    return;
  }

  float x[8], g[8], s[8], y[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) x[e] = g[e] = s[e] = y[e] = 0.f;
  if (action == kRestart) {
    ld8(v.x + off, in_a, in_b, x);
    ld8(v.g + off, in_a, in_b, g);
  } else {   // kAccept, kStartAccept, kFinalAccept: the trial point becomes the iterate
    float xo[8], go[8];
    ld8(v.x + off, in_a, in_b, xo);
    ld8(v.g + off, in_a, in_b, go);
    ld8(v.xt + off, in_a, in_b, x);
    ld8(v.gt + off, in_a, in_b, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s[e] = x[e] - xo[e];
      y[e] = g[e] - go[e];
    }
    st8(v.x + off, in_a, in_b, x);
    st8(v.g + off, in_a, in_b, g);
    if (action == kAccept && c.push && in_a) {
      const size_t hoff = (static_cast<size_t>(b) * v.m + c.slot_new) * v.hn + i0;
      // s, y continue as the STORED (unscaled) halves: the step below multiplies them by the scaled coefficients
      sth8(v.S + hoff, s, 1.f / c.hs[c.slot_new]);
      sth8(v.Y + hoff, y, 1.f / c.hy[c.slot_new]);
    }
    if (action == kFinalAccept) return;
  }

  // New step over the free variables: -g / theta + sum_j cs_j S_j + cy_j Y_j (coefficients from the decide kernel;
  // the projected gradient step without history), projected onto the box: z = clip(x + step), d = z - x.
  const int nh = (action == kAccept) ? c.nh : 0;
  const float ti = nh > 0 ? c.theta_inv : 1.f;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = -ti * g[e];
  if (nh > 0) {
    const int nold = c.push ? nh - 1 : nh;
    for (int j = 0; j < nold; j += kHistTrip) {       // see the dots kernel
      uint4 rs[kHistTrip], ry[kHistTrip];
#pragma unroll
      for (int q = 0; q < kHistTrip; ++q) {
        rs[q] = ry[q] = make_uint4(0u, 0u, 0u, 0u);
        if (in_a && j + q < nold) {
          const size_t h = (static_cast<size_t>(b) * v.m + (c.head + j + q) % v.m) * v.hn + i0;
          rs[q] = ldh8_raw(v.S + h);
          ry[q] = ldh8_raw(v.Y + h);
        }
      }
#pragma unroll
      for (int q = 0; q < kHistTrip; ++q) {
        if (j + q >= nold) break;
        float sj[8], yj[8];
        cvth8(rs[q], sj);
        cvth8(ry[q], yj);
        const float a = c.cs[j + q], bb = c.cy[j + q];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(a, sj[e], fmaf(bb, yj[e], acc[e]));
      }
    }
    if (c.push) {
      const float a = c.cs[nh - 1], bb = c.cy[nh - 1];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(a, s[e], fmaf(bb, y[e], acc[e]));
    }
  }
  float d[8], t[8];
  float gd = 0.f, smax = 3.0e38f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const bool in = e < 4 ? in_a : in_b;
    const bool fr = is_free(x[e], g[e], lo, hi, bd);
    d[e] = fr ? clipf(x[e] + acc[e], lo, hi, bd) - x[e] : 0.f;
    t[e] = clipf(fmaf(stp, d[e], x[e]), lo, hi, bd);
    if (in) {
      gd = fmaf(g[e], d[e], gd);
      if (bd) {
        // lnsrlb: the largest step that keeps x + stp d inside the box
        const float room = d[e] > 0.f ? (hi - x[e]) / d[e] : (d[e] < 0.f ? (lo - x[e]) / d[e] : 3.0e38f);
        smax = fminf(smax, room);
      }
    }
  }
  st8(v.d + off, in_a, in_b, d);
  st8(v.xt + off, in_a, in_b, t);
  __shared__ float red[kThreads / 32], redm[kThreads / 32];
  const float r = warp_sum(gd);
  float mn = smax;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if (lane == 0) {
    red[warp] = r;
    redm[warp] = mn;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f, mm = 3.0e38f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      t += red[w];
      mm = fminf(mm, redm[w]);
    }
    v.part2[static_cast<size_t>(b) * v.nblk + blockIdx.x] = t;
    v.part3[static_cast<size_t>(b) * v.nblk + blockIdx.x] = mm;
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
    c.nit = 0; c.nfev = 0; c.fresh = 0; c.maxiter = 0; c.stp = 0.f; c.theta_inv = 1.f;
    c.f = 0.0; c.pp = 0.0; c.stp0 = 1.0; c.theta = 1.0; c.stpmx = 1.0;
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

// active (nullable): the plan's per-image switch - a finished job's image is skipped by the f/g kernels from the next
// cycle on (the jobs of a batch finish at different times; without this the tail of a run costs full cycles).
// all_on = 1: switch every image back on (end of a run: the plan is shared with other callers).
__global__ void lbfgs_export_kernel(Vecs v, float* fun, int* nit, int* nfev, int* mode, int nprob, int* active,
                                    int all_on) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nprob) {
    const Ctrl& c = v.ctrl[b];
    if (fun) fun[b] = static_cast<float>(c.f);
    if (nit) nit[b] = c.nit;
    if (nfev) nfev[b] = c.nfev;
    if (mode) mode[b] = c.mode;
    if (active) active[b] = (all_on || c.mode != kDone) ? 1 : 0;
  }
}

}  // namespace

struct LbfgsImpl {
  Vecs v;
  int B = 0;
  float* base = nullptr;
  void* hist = nullptr;        // the (s, y) history, fp16
  float* fun_dev = nullptr;
  int* status_dev = nullptr;   // [3B] nit | nfev | mode
  int* pinned_i = nullptr;     // [3B]
  float* pinned_f = nullptr;   // [n*B] x staging + [B] fun
  // The cycle is replayed as a CUDA graph. Capture is not allowed on the legacy default stream, so the optimiser
  // owns a stream and fences it against the caller's stream at entry and exit.
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  // The active-set kernel (a latency-bound gather: ~16 k scattered 4-byte reads per block, bounded by the SM's
  // outstanding requests x DRAM latency) runs on a side stream beside the bandwidth-bound dots kernel; both only read.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
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
    if (im->side) cudaStreamDestroy(im->side);
    if (im->ev_fork) cudaEventDestroy(im->ev_fork);
    if (im->ev_join) cudaEventDestroy(im->ev_join);
    if (im->ev_in) cudaEventDestroy(im->ev_in);
    if (im->ev_out) cudaEventDestroy(im->ev_out);
    cudaFree(im->base);
    cudaFree(im->hist);
    cudaFree(im->v.ctrl);
    cudaFree(im->v.tab);
    cudaFree(im->v.gramA);
    cudaFree(im->v.dbg);
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
  v.hn = static_cast<int>((n + 7) & ~static_cast<size_t>(7));
  v.active = g_lbfgs_no_skip ? nullptr : plan->d_active;
  v.m = maxcor;
  v.nblk = static_cast<int>((n + kChunk - 1) / kChunk);
  v.bounded = bounded;
  v.lo = lower;
  v.hi = upper;
  const size_t vec = static_cast<size_t>(B) * n;
  const size_t total = vec * 5 + B + static_cast<size_t>(B) * v.nblk * (kNP + 2) + B;
  const size_t hvec = static_cast<size_t>(B) * v.hn;
  const size_t hist_bytes = hvec * maxcor * 2 * sizeof(hist_t);
  const size_t gram_bytes = static_cast<size_t>(B) * kSlices * kQ * kQ * sizeof(double);
  cudaError_t e = cudaMalloc(&im->base, total * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(im->base, 0, total * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&im->hist, hist_bytes);
  if (e == cudaSuccess) e = cudaMemset(im->hist, 0, hist_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&v.ctrl, sizeof(Ctrl) * B + 3 * sizeof(int) * B);
  if (e == cudaSuccess) e = cudaMemset(v.ctrl, 0, sizeof(Ctrl) * B + 3 * sizeof(int) * B);
  if (e == cudaSuccess) e = cudaMalloc(&v.tab, sizeof(Tables) * B);
  if (e == cudaSuccess) e = cudaMemset(v.tab, 0, sizeof(Tables) * B);
  if (e == cudaSuccess) e = cudaMalloc(&v.dbg, 16 * sizeof(long long));
  if (e == cudaSuccess) e = cudaMemset(v.dbg, 0, 16 * sizeof(long long));
  if (e == cudaSuccess) e = cudaMalloc(&v.gramA, gram_bytes);
  if (e == cudaSuccess) e = cudaMemset(v.gramA, 0, gram_bytes);
  if (e == cudaSuccess) e = cudaMallocHost(&im->pinned_i, 3 * sizeof(int) * B);
  if (e == cudaSuccess) e = cudaMallocHost(&im->pinned_f, (vec + B) * sizeof(float));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&im->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&im->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&im->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&im->ev_join, cudaEventDisableTiming);
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
  v.S = static_cast<hist_t*>(im->hist);
  v.Y = v.S + hvec * maxcor;
  v.ft = p; p += B;
  v.part = p; p += static_cast<size_t>(B) * v.nblk * kNP;
  v.part2 = p; p += static_cast<size_t>(B) * v.nblk;
  v.part3 = p; p += static_cast<size_t>(B) * v.nblk;
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
  if (g_lbfgs_serial_active) {
    STX_TRY(launch_pdl(lbfgs_dots_kernel, grid, dim3(kThreads), 0, stream, v));
    STX_TRY(launch_pdl(lbfgs_active_kernel, dim3(kSlices, im->B), dim3(kThreads), 0, stream, v));
  } else {
    STX_CUDA(cudaEventRecord(im->ev_fork, stream));
    STX_CUDA(cudaStreamWaitEvent(im->side, im->ev_fork, 0));
    STX_TRY(launch_pdl(lbfgs_active_kernel, dim3(kSlices, im->B), dim3(kThreads), 0, im->side, v));
    STX_CUDA(cudaEventRecord(im->ev_join, im->side));
    STX_TRY(launch_pdl(lbfgs_dots_kernel, grid, dim3(kThreads), 0, stream, v));
    STX_CUDA(cudaStreamWaitEvent(stream, im->ev_join, 0));
  }
  STX_TRY(launch_pdl(lbfgs_decide_kernel, dim3(im->B), dim3(kDecideThreads), 0, stream, v));
  STX_TRY(launch_pdl(lbfgs_update_kernel, grid, dim3(kThreads), 0, stream, v));
  count_launch(4);
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
                                                             im->status_dev + 2 * B, B,
                                                             v.active, 0);
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
  lbfgs_export_kernel<<<(B + 127) / 128, 128, 0, stream>>>(v, nullptr, nullptr, nullptr, nullptr, B, plan->d_active, 1);
  STX_CUDA(cudaGetLastError());
  STX_TRY(fence_out(im, caller));
  return STX_OK;
}

int Lbfgs::phase_cycles(long long* host16) {
  LbfgsImpl* im = impl_of(this);
  STX_CUDA(cudaStreamSynchronize(im->stream));
  STX_CUDA(cudaMemcpy(host16, im->v.dbg, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
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
  lbfgs_export_kernel<<<(B + 127) / 128, 128, 0, im->stream>>>(v, fun, nit, nfev, nullptr, B, nullptr, 0);
  STX_CUDA(cudaGetLastError());
  STX_TRY(fence_out(im, caller));
  return STX_OK;
}

// Host-side copies between the caller's (pageable) arrays and the pinned staging buffer: 50 MB each way at 64 jobs of
// 256^2, and the caller's fresh output array is first touched here - a few threads instead of one memcpy.
template <typename F>
static void parallel_ranges(size_t n, F f) {
  const unsigned hw = std::thread::hardware_concurrency();
  size_t T = std::min<size_t>(8, std::max(1u, hw));
  if (n < (size_t(1) << 20)) T = 1;
  const size_t per = (((n + T - 1) / T) + 1023) & ~size_t(1023);
  std::vector<std::thread> th;
  for (size_t t = 1; t < T; ++t) {
    const size_t b = std::min(n, t * per), e = std::min(n, (t + 1) * per);
    if (b < e) th.emplace_back([=] { f(b, e); });
  }
  f(0, std::min(n, per));
  for (auto& t : th) t.join();
}

int Lbfgs::minimize_host(const float* x0_host, int is_preimage_, int maxiter, void* x_host, int x_is_f64,
                         float* fun_host, int* nit_host, int* nfev_host, cudaStream_t caller) {
  LbfgsImpl* im = impl_of(this);
  Vecs& v = im->v;
  const int B = im->B;
  const size_t vec = static_cast<size_t>(B) * v.n;
  float* const stage = im->pinned_f;
  parallel_ranges(vec, [=](size_t b, size_t e) { memcpy(stage + b, x0_host + b, (e - b) * sizeof(float)); });
  // gt doubles as the device staging buffer for x0 (it is overwritten by the first evaluation)
  STX_CUDA(cudaMemcpyAsync(v.gt, im->pinned_f, vec * sizeof(float), cudaMemcpyHostToDevice, caller));
  STX_TRY(reset(v.gt, is_preimage_, caller));
  STX_TRY(run(maxiter, 0, caller));
  STX_TRY(result(nullptr, im->fun_dev, im->status_dev, im->status_dev + B, caller));
  STX_CUDA(cudaMemcpyAsync(im->pinned_f, v.x, vec * sizeof(float), cudaMemcpyDeviceToHost, caller));
  STX_CUDA(cudaMemcpyAsync(im->pinned_f + vec, im->fun_dev, B * sizeof(float), cudaMemcpyDeviceToHost, caller));
  STX_CUDA(cudaMemcpyAsync(im->pinned_i, im->status_dev, 2 * sizeof(int) * B, cudaMemcpyDeviceToHost, caller));
  STX_CUDA(cudaStreamSynchronize(caller));
  if (x_is_f64) {     // the reference returns scipy's float64 vector: widen while copying out
    double* const out = static_cast<double*>(x_host);
    parallel_ranges(vec, [=](size_t b, size_t e) {
      for (size_t i = b; i < e; ++i) out[i] = static_cast<double>(stage[i]);
    });
  } else {
    float* const out = static_cast<float*>(x_host);
    parallel_ranges(vec, [=](size_t b, size_t e) { memcpy(out + b, stage + b, (e - b) * sizeof(float)); });
  }
  memcpy(fun_host, im->pinned_f + vec, B * sizeof(float));
  if (nit_host) memcpy(nit_host, im->pinned_i, B * sizeof(int));
  if (nfev_host) memcpy(nfev_host, im->pinned_i + B, B * sizeof(int));
  return STX_OK;
}

}  // namespace stx
