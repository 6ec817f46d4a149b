// Device-resident L-BFGS-B (see lbfgs.cu).
#pragma once
#include "plan.h"

namespace stx {

struct Lbfgs {
  Plan* plan = nullptr;
  int m = 20;
  int bounded = 1;
  float lo = 0.0f, hi = 1.0f;
  int is_preimage = 0;
  void* arena = nullptr;
  ~Lbfgs();
  int reset(const float* x0, int is_preimage, cudaStream_t stream);
  int run(int maxiter, int max_evals, cudaStream_t stream);
  int result(float* x, float* fun, int* nit, int* nfev, cudaStream_t stream);
  int phase_cycles(long long* host16);
  int minimize_host(const float* x0_host, int is_preimage, int maxiter, void* x_host, int x_is_f64, float* fun_host,
                    int* nit_host, int* nfev_host, cudaStream_t stream);
};
int lbfgs_create(Lbfgs** out, Plan* plan, int maxcor, int bounded, float lower, float upper);

}  // namespace stx
