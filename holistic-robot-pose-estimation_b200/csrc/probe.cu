// Hardware probe (not on the product path): issue-to-completion time of back-to-back tcgen05.mma instructions as
// a function of (M, N), operands in shared memory (SS mode), bf16, K=16 per instruction.  Used to size the N tile
// and to decide which layers can be tensor-bound at all.
#include "../../include/hrp.h"
#include "hrp_common.cuh"

namespace hrp {

__global__ void __launch_bounds__(128) mma_rate_kernel(int M, int N, int reps, int kdistinct, long long* out) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (128 * 128 + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16((uint32_t)M, (uint32_t)N);
    const uint32_t sa = smem_u32(smem), sb = sa + 128 * 128;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const int k = r % kdistinct;  // 0..3: K=16 slices inside one 128-byte swizzled row
        umma_bf16_ss(tmem_base, make_kmajor_desc(sa + k * 32, 1024, 2), make_kmajor_desc(sb + k * 32, 1024, 2), idesc,
                     r != 0);
      }
      t1 = clock64();
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
      out[0] = t1 - t0;  // issue time
      out[1] = t2 - t0;  // completion time
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace hrp

extern "C" int hrp_probe_mma_rate(int32_t M, int32_t N, int32_t reps, int32_t kdistinct, long long* dev_out2,
                                  int32_t ctas) {
  using namespace hrp;
  HRP_REQUIRE((M == 64 || M == 128) && N >= 16 && N <= 256 && N % 16 == 0 && reps > 0 && dev_out2 != nullptr, "bad args");
  const int smem = 128 * 128 + 256 * 128 + 2048;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_rate_kernel<<<ctas, 128, smem>>>(M, N, reps, kdistinct < 1 ? 1 : kdistinct, dev_out2);
  HRP_CUDA_CHECK(cudaGetLastError());
  HRP_CUDA_CHECK(cudaDeviceSynchronize());
  return HRP_OK;
}
