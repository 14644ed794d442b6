// Process-wide count of kernels launched by this library (bench.py reports it as "gpu_launches").
#pragma once
#include <atomic>
#include <stdint.h>
namespace hrp {
extern std::atomic<int64_t> g_launch_count;
inline void count_launch(int64_t n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

// Programmatic dependent launch (PDL).  The network executor sets this flag around a launch whose predecessor on the
// same stream is one of our kernels: the launch then carries cudaLaunchAttributeProgrammaticStreamSerialization, so
// the kernel's CTAs may start (barrier init, TMEM allocation, weight fetch) while the predecessor drains; every
// kernel calls griddepcontrol.wait (pdl_wait() in hrp_common.cuh) before touching activations.
extern thread_local bool g_pdl_launch;
struct PdlScope {
  bool prev;
  explicit PdlScope(bool on) : prev(g_pdl_launch) { g_pdl_launch = on; }
  ~PdlScope() { g_pdl_launch = prev; }
};
}  // namespace hrp

#ifdef __CUDACC__
#include <cuda_runtime.h>
#include <stdlib.h>
#include <utility>
namespace hrp {
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  static const bool force = (getenv("HRP_PDL_FORCE") != nullptr);  // probing aid: PDL on every launch
  if (g_pdl_launch || force) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
}  // namespace hrp
#endif
