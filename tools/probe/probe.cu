// Hardware probe (not on the product path): issue-to-completion time of back-to-back tcgen05.mma instructions as
// a function of (M, N), operands in shared memory (SS mode), bf16, K=16 per instruction.  Used to size the N tile
// and to decide which layers can be tensor-bound at all.
#include "hrp_probe.h"
#include "hrp_common.cuh"
#include "conv.h"

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

// ------------------------------------------------------------------------------------------------------
// Probe 4: TMA tiled-load rate of ONE SM as the convolution kernels use it.  Tensor = NHWC bf16 activations
// (B, H, W, C); every load is the 4-D box {64 channels, bw, bh, bn} (128 pixels x 128 bytes = 16 KiB, SWIZZLE_128B), i.e.
// 128 rows of 128 B whose stride in memory is C * 2 bytes.  One producer thread keeps `depth` loads in flight through a
// ring of shared-memory buffers (full / empty mbarriers); a consumer thread releases a buffer as soon as it has landed.
// Tiles are walked like the persistent kernel does: tile t = cta / share + i * (grid / share), and for each tile the
// C / 64 channel blocks in turn (share > 1: `share` neighbouring CTAs read the same tiles at the same time).
// dev_out[0] = cycles of CTA 0 for `iters` loads.
// ------------------------------------------------------------------------------------------------------
namespace hrp {
// wait flavours for the ring probe: 0 = mbar_wait (try_wait spin, the kernels' default), 1 = test_wait spin (never
// suspends), 2 = try_wait with an explicit suspend-time hint of 0 ns
__device__ __forceinline__ void probe_wait(uint64_t* bar, uint32_t parity, int mode) {
  if (mode == 0) {
    mbar_wait(bar, parity);
    return;
  }
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    if (mode == 1)
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
                   : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    else
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
                   : "=r"(done) : "r"(addr), "r"(parity), "r"(0u) : "memory");
  }
}
__global__ void __launch_bounds__(96) tma_rate_kernel(const __grid_constant__ CUtensorMap map, int depth, int iters, int kblocks,
                                                      int tiles_w, int tiles_h, int tiles_total, int bw, int bh, int bn, int share,
                                                      int rank2, int box_bytes, int producers, int wait_mode, long long* out) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_align1024(smem_dyn);
  __shared__ uint64_t full[16], empty[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&map);
  }
  __syncthreads();
  const int group = blockIdx.x / share, ngroups = gridDim.x / share;
  const long long t0 = clock64();
  const int trace0 = 600;  // first traced load (steady state)
  if (threadIdx.x == 0 || (producers == 2 && threadIdx.x == 64)) {
    // (no divisions in the loop: the coordinates advance with wrap-around counters -- an integer division costs the
    //  issuing thread more than a TMA instruction does)
    const int me = (threadIdx.x == 0) ? 0 : 1;
    int s = me;
    uint32_t par = 0;
    int kb = 0;
    int t = group % tiles_total;
    int tw = t % tiles_w, th = (t / tiles_w) % tiles_h, tn = t / (tiles_w * tiles_h);
    const int rows = box_bytes / 128;
    const int step_w = ngroups % tiles_w, step_h = (ngroups / tiles_w) % tiles_h, step_n = ngroups / (tiles_w * tiles_h);
    const int n_tiles_n = tiles_total / (tiles_w * tiles_h);
    long long tw_wait = 0, tw_exp = 0, tw_tma = 0;
    for (int i = me; i < iters; i += producers) {
      const long long c0 = clock64();
      probe_wait(&empty[s], par ^ 1, wait_mode);
      const long long c1 = clock64();
      mbar_expect_tx(&full[s], (uint32_t)box_bytes);
      const long long c2 = clock64();
      if (rank2) {
        tma_load_2d(smem + (size_t)s * box_bytes, &map, &full[s], kb * 64, t * rows);
      } else {
        tma_load_4d(smem + (size_t)s * box_bytes, &map, &full[s], kb * 64, tw * bw, th * bh, tn * bn);
      }
      const long long c3 = clock64();
      if (blockIdx.x == 0 && i >= trace0 && i < trace0 + 96) {
        out[8 + (i - trace0) * 4 + 0] = c1 - t0;
        out[8 + (i - trace0) * 4 + 1] = c3 - t0;
      }
      tw_wait += c1 - c0;
      tw_exp += c2 - c1;
      tw_tma += c3 - c2;
      for (int r = 0; r < producers; ++r)
        if (++kb == kblocks) {
          kb = 0;
          t += ngroups;
          if (t >= tiles_total) t -= tiles_total;
          tw += step_w;
          if (tw >= tiles_w) { tw -= tiles_w; ++th; }
          th += step_h;
          if (th >= tiles_h) { th -= tiles_h; ++tn; }
          tn += step_n;
          if (tn >= n_tiles_n) tn -= n_tiles_n;
        }
      s += producers;
      if (s >= depth) {
        s -= depth;
        par ^= 1;
      }
    }
    if (blockIdx.x == 0 && me == 0) {  // where the issuing thread spends its time (cycles summed over its loads)
      out[1] = tw_wait;
      out[2] = tw_exp;
      out[3] = tw_tma;
    }
  } else if (threadIdx.x == 32) {
    int s = 0;
    uint32_t par = 0;
    for (int i = 0; i < iters; ++i) {
      probe_wait(&full[s], par, wait_mode);
      if (blockIdx.x == 0 && i >= trace0 && i < trace0 + 96) out[8 + (i - trace0) * 4 + 2] = clock64() - t0;
      mbar_arrive(&empty[s]);
      if (++s == depth) {
        s = 0;
        par ^= 1;
      }
    }
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  }
}
}  // namespace hrp

// rows: 0 -> 4-D box {64, bw, bh, bn} (128 pixels); > 0 -> 2-D box {64 channels, rows pixels} on the (B*H*W, C) view
extern "C" int hrp_probe_tma_rate(const void* act, int32_t B, int32_t H, int32_t W, int32_t C, int32_t bw, int32_t bh,
                                  int32_t bn, int32_t depth, int32_t iters, int32_t share, int32_t ctas, int32_t rows,
                                  int32_t producers, long long* dev_out) {
  using namespace hrp;
  HRP_REQUIRE(act != nullptr && C % 64 == 0 && bw * bh * bn == 128 && depth >= 1 && depth <= 13 && share >= 1 &&
                  ctas % share == 0 && W % bw == 0 && H % bh == 0 && B % bn == 0 && rows >= 0 && rows <= 256 &&
                  ((producers & 0xff) == 1 || ((producers & 0xff) == 2 && depth % 2 == 0)),
              "bad args");
  CUtensorMap map;
  int rc;
  const int box_bytes = (rows > 0 ? rows : 128) * 128;
  const long long npix = (long long)B * H * W;
  if (rows > 0) {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)npix};
    uint64_t strides[1] = {(uint64_t)C * 2};
    uint32_t box[2] = {64u, (uint32_t)rows};
    rc = conv_encode_map(&map, act, 2, dims, strides, box, 64);
  } else {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64u, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
    rc = conv_encode_map(&map, act, 4, dims, strides, box, 64);
  }
  if (rc != HRP_OK) return rc;
  HRP_REQUIRE(depth * box_bytes <= 220 * 1024, "ring does not fit in shared memory");
  const int smem = depth * box_bytes + 1024;
  cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int tiles_w = W / bw, tiles_h = H / bh;
  const int tiles_total = rows > 0 ? (int)(npix / rows) : tiles_w * tiles_h * (B / bn);
  tma_rate_kernel<<<ctas, 96, smem>>>(map, depth, iters, C / 64, tiles_w, tiles_h, tiles_total, bw, bh, bn, share,
                                      rows > 0 ? 1 : 0, box_bytes, producers & 0xff, (producers >> 8) & 0xff, dev_out);
  HRP_CUDA_CHECK(cudaGetLastError());
  HRP_CUDA_CHECK(cudaDeviceSynchronize());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// Probe 5: raw issue cost of cp.async.bulk.tensor for ONE thread: n back-to-back 2-D loads of {64 ch, rows} boxes into n
// shared-memory slots, all completing on one mbarrier.  dev_out[0] = cycles until the last instruction has issued,
// dev_out[1] = cycles until all bytes have landed (CTA 0; every CTA does the same on its own tiles).
// ------------------------------------------------------------------------------------------------------
namespace hrp {
__global__ void __launch_bounds__(32) tma_issue_kernel(const __grid_constant__ CUtensorMap map, int n, int rows, int reps,
                                                       long long* out) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_align1024(smem_dyn);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&map);
    long long t_issue = 0, t_done = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      mbar_expect_tx(&bar, (uint32_t)(n * rows * 128));
#pragma unroll 1
      for (int i = 0; i < n; ++i)
        tma_load_2d(smem + (size_t)i * rows * 128, &map, &bar, 0, ((int)blockIdx.x * n + i + r * 37) * rows);
      const long long t1 = clock64();
      mbar_wait(&bar, (uint32_t)(r & 1));
      const long long t2 = clock64();
      if (r > 0) {
        t_issue += t1 - t0;
        t_done += t2 - t0;
      }
    }
    if (blockIdx.x == 0) {
      out[0] = t_issue / (reps - 1);
      out[1] = t_done / (reps - 1);
    }
  }
}
}  // namespace hrp

extern "C" int hrp_probe_tma_issue(const void* act, int64_t npix, int32_t C, int32_t rows, int32_t n, int32_t reps,
                                   int32_t ctas, long long* dev_out) {
  using namespace hrp;
  HRP_REQUIRE(act != nullptr && C % 64 == 0 && rows >= 8 && rows <= 256 && n >= 1 && n * rows * 128 <= 220 * 1024 && reps >= 2,
              "bad args");
  CUtensorMap map;
  uint64_t dims[2] = {(uint64_t)C, (uint64_t)npix};
  uint64_t strides[1] = {(uint64_t)C * 2};
  uint32_t box[2] = {64u, (uint32_t)rows};
  int rc = conv_encode_map(&map, act, 2, dims, strides, box, 64);
  if (rc != HRP_OK) return rc;
  const int smem = n * rows * 128 + 1024;
  cudaFuncSetAttribute(tma_issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  tma_issue_kernel<<<ctas, 32, smem>>>(map, n, rows, reps, dev_out);
  HRP_CUDA_CHECK(cudaGetLastError());
  HRP_CUDA_CHECK(cudaDeviceSynchronize());
  return HRP_OK;
}
