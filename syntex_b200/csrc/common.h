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

// Programmatic dependent launch (PDL): a kernel launched through launch_pdl() may start while its predecessor in the
// stream is still draining - its CTAs become resident as SM resources free up, run their prologue (barrier init,
// TMEM allocation, descriptor prefetch) and then block in pdl_wait() (griddepcontrol.wait) until the predecessor
// has completed and flushed. Every kernel launched this way executes pdl_wait() before its first global-memory
// access, so ordering through memory is unchanged; what disappears is the launch gap and the idle tail between
// the ~35 dependent launches of one f/g evaluation. g_pdl = 0 (stx_debug_set key 9) restores plain launches.
extern int g_pdl;
// cluster_x > 1: launch as thread-block clusters of that many CTAs along x (grid.x must be a multiple of it).
template <typename... KArgs, typename... Args>
int launch_pdl_cluster(int cluster_x, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx");
  return 0;
}
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  return launch_pdl_cluster(1, kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

static inline int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

}  // namespace stx
