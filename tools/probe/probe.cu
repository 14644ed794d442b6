// Hardware probe (not on the product path): issue-to-completion time of back-to-back tcgen05.mma instructions as
// a function of (M, N), operands in shared memory (SS mode), bf16, K=16 per instruction.  Used to size the N tile
// and to decide which layers can be tensor-bound at all.
#include "hrp_probe.h"
#include "hrp_common.cuh"

namespace hrp {

__global__ void __launch_bounds__(128) mma_rate_kernel(int M, int N, int reps, int kdistinct, int nacc, int ck, int shift,
                                                        long long* out) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_align1024(smem_dyn);
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
      // descriptors and accumulator addresses precomputed, loop unrolled by 8: the loop must not be bound by this
      // thread's own scalar latency (integer modulo, descriptor packing), only by the tcgen05 issue path
      uint64_t ad[4], bd[4];
      for (int k = 0; k < 4; ++k) {
        // ck selects the swizzle mode / row pitch (64: SW128, 32: SW64, 16: SW32); shift = A start row offset
        const uint32_t layout = (ck == 64) ? 2u : (ck == 32) ? 4u : 6u;
        const int kk = (k % kdistinct) % (ck / 16);
        ad[k] = make_kmajor_desc(sa + shift * ck * 2 + kk * 32, 8 * ck * 2, layout);
        bd[k] = make_kmajor_desc(sb + kk * 32, 8 * ck * 2, layout);
      }
      uint32_t acc[8];
      for (int i = 0; i < 8; ++i) acc[i] = tmem_base + (uint32_t)((i % nacc) * N);
      for (int i = 0; i < 8; ++i) umma_bf16_ss(acc[i], ad[i & 3], bd[i & 3], idesc, i >= nacc);
      for (int r = 8; r < reps; r += 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) umma_bf16_ss(acc[i], ad[i & 3], bd[i & 3], idesc, 1u);
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
  const int kd = (kdistinct & 0xff) < 1 ? 1 : (kdistinct & 0xff);
  int nacc = ((kdistinct >> 8) & 0xff) < 1 ? 1 : ((kdistinct >> 8) & 0xff);
  while (nacc > 1 && nacc * N > 512) --nacc;
  const int ck = ((kdistinct >> 16) & 0xff) ? ((kdistinct >> 16) & 0xff) : 64;
  const int shift = (kdistinct >> 24) & 0x7f;
  mma_rate_kernel<<<ctas, 128, smem>>>(M, N, reps, kd, nacc, ck, shift, dev_out2);
  HRP_CUDA_CHECK(cudaGetLastError());
  HRP_CUDA_CHECK(cudaDeviceSynchronize());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// Probe 2: does a K-major swizzled UMMA operand tolerate a start address shifted by an arbitrary number of ROWS
// (not a multiple of the 8-row swizzle pattern)?  The halo-tile convolution reads the nine 3x3 taps as shifted
// 128-row windows of ONE shared-memory tile, which needs exactly that.  A [rows x CK] bf16 matrix is written to
// shared memory in the TMA swizzle layout (CK=64: SW128, 32: SW64, 16: SW32), one M=128 x N=32 x K=CK product is
// issued with A starting `shift` rows into the tile, and D is copied out for the host to check.
// bo_mode: 0 -> descriptor base_offset = 0; 1 -> base_offset = (start_address >> 7) & 7.
// ------------------------------------------------------------------------------------------------------
namespace hrp {

__global__ void __launch_bounds__(128) desc_shift_kernel(int ck, int rows, int shift, int bo_mode, const uint16_t* A,
                                                         const uint16_t* Bm, float* out) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_align1024(smem_dyn);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_bytes = ck * 2;
  const int chunks = row_bytes / 16;             // 16-byte chunks per row: 8 / 4 / 2
  const int a_bytes = (rows * row_bytes + 1023) / 1024 * 1024;
  uint8_t* sB = smem + a_bytes;
  auto swz = [&](int r) { return (ck == 64) ? (r & 7) : (ck == 32) ? ((r >> 1) & 3) : ((r >> 2) & 1); };
  for (int i = threadIdx.x; i < rows * chunks; i += blockDim.x) {
    const int r = i / chunks, c = i % chunks;
    const uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)r * ck + c * 8);
    *reinterpret_cast<uint4*>(smem + (size_t)r * row_bytes + ((c ^ swz(r)) << 4)) = v;
  }
  for (int i = threadIdx.x; i < 32 * chunks; i += blockDim.x) {
    const int r = i / chunks, c = i % chunks;
    const uint4 v = *reinterpret_cast<const uint4*>(Bm + (size_t)r * ck + c * 8);
    *reinterpret_cast<uint4*>(sB + (size_t)r * row_bytes + ((c ^ swz(r)) << 4)) = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 32);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, 32);
      const uint32_t layout = (ck == 64) ? 2u : (ck == 32) ? 4u : 6u;
      const uint32_t sbo = 8 * row_bytes;
      const uint32_t a0 = smem_u32(smem) + (uint32_t)(shift * row_bytes);
      const uint32_t b0 = smem_u32(sB);
      for (int k = 0; k < ck / 16; ++k) {
        uint64_t ad = make_kmajor_desc(a0 + k * 32, sbo, layout);
        if (bo_mode == 1) ad |= (uint64_t)((a0 >> 7) & 7u) << 49;
        const uint64_t bd = make_kmajor_desc(b0 + k * 32, sbo, layout);
        umma_bf16_ss(tmem_base, ad, bd, idesc, k != 0);
      }
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t acc[32];
  tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), acc);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + lane) * 32 + j] = __uint_as_float(acc[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 32);
}

}  // namespace hrp

extern "C" int hrp_probe_desc_shift(int32_t ck, int32_t rows, int32_t shift, int32_t bo_mode, const void* A_dev,
                                    const void* B_dev, float* out_dev) {
  using namespace hrp;
  HRP_REQUIRE((ck == 16 || ck == 32 || ck == 64) && rows >= shift + 128 && rows <= 512 && shift >= 0, "bad args");
  const int smem = rows * ck * 2 + 32 * ck * 2 + 4096;
  cudaFuncSetAttribute(desc_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  desc_shift_kernel<<<1, 128, smem>>>(ck, rows, shift, bo_mode, reinterpret_cast<const uint16_t*>(A_dev),
                                      reinterpret_cast<const uint16_t*>(B_dev), out_dev);
  HRP_CUDA_CHECK(cudaGetLastError());
  HRP_CUDA_CHECK(cudaDeviceSynchronize());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// Probe 3: tcgen05.ld throughput.  W warps (4 or 8; warp w reads TMEM lane quarter w % 4) each issue `reps`
// 32x32b.x32 loads (32 lanes x 32 fp32 columns = 4 KiB per instruction), optionally waiting after each one.
// The epilogue of every conv kernel is bounded below by this rate (accumulator bytes / TMEM read bandwidth).
// ------------------------------------------------------------------------------------------------------
namespace hrp {

__global__ void __launch_bounds__(256) tmem_ld_rate_kernel(int reps, int wait_each, long long* out) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc[32];
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    tmem_ld32(taddr + (uint32_t)((r * 32) & 511), acc);
    if (wait_each) {
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) sink ^= acc[i];
    }
  }
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) sink ^= acc[i];
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) {
    out[0] = t1 - t0;
    out[1] = (long long)sink;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base_s, 512);
  }
}

}  // namespace hrp

extern "C" int hrp_probe_tmem_ld_rate(int32_t warps, int32_t reps, int32_t wait_each, long long* dev_out2) {
  using namespace hrp;
  HRP_REQUIRE((warps == 1 || warps == 4 || warps == 8) && reps > 0 && dev_out2 != nullptr, "bad args");
  tmem_ld_rate_kernel<<<1, warps * 32, 0>>>(reps, wait_each, dev_out2);
  HRP_CUDA_CHECK(cudaGetLastError());
  HRP_CUDA_CHECK(cudaDeviceSynchronize());
  return HRP_OK;
}
