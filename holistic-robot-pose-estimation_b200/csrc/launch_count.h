// Process-wide count of kernels launched by this library (bench.py reports it as "gpu_launches").
#pragma once
#include <atomic>
#include <stdint.h>
namespace hrp {
extern std::atomic<int64_t> g_launch_count;
inline void count_launch(int64_t n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }
}  // namespace hrp
