// Host-side helpers shared by the launchers: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/syntex_b200.h"

namespace stx {

typedef __nv_bfloat16 bf16;

// Error convention of the C-ABI (include/syntex_b200.h): 0 = ok, negative = failure (STX_ERR_* codes), message
// kept per thread and readable through stx_last_error().
void set_error(const std::string& msg);
const std::string& last_error();
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
// Number of kernel launches issued by this library (graph replays count their kernel nodes). bench.py reports it.
void count_launch(long long n = 1);
long long launch_count();

#define STX_CUDA(expr)                                   \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return ::stx::cuda_fail(_e, #expr); \
  } while (0)

#define STX_TRY(expr)        \
  do {                       \
    int _r = (expr);         \
    if (_r != 0) return _r;  \
  } while (0)

#define STX_REQUIRE(cond, msg) \
  do {                         \
    if (!(cond)) return ::stx::fail(STX_ERR_INVALID, std::string(msg) + " [" #cond "]"); \
  } while (0)

// bf16 tensor map, 128-byte swizzle, zero fill for out-of-bounds elements (this is what implements the
// SAME zero padding of the 3x3 convolutions: halo coordinates are simply out of range).
// dims/box are fastest-first; strides are derived for a dense tensor.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint32_t* box);

static inline int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

}  // namespace stx
