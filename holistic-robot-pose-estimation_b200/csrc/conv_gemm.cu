// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (see conv.h for the plan model).
//
// CTA = 192 threads: warp 0 = TMA producer (one elected lane), warp 1 = TMEM owner + MMA issuer (one elected
// lane issues tcgen05.mma, M=128, N=n_tile, K=16, bf16 x bf16 -> fp32 in TMEM), warps 2..5 = epilogue
// (tcgen05.ld 32 lanes x 16 columns, folded-BN scale/shift, residual / upsample-add / post-add, ReLU, bf16
// NHWC store, optional fp32 global-average-pool accumulation).
//
// Shared-memory stage = K extent 64: 64/ck sub-tiles, each a K-major [128 rows x ck] A tile (row = row pixel,
// written by ONE 4-D tiled TMA box whose out-of-bounds elements are zero-filled: that is the conv padding)
// and a [n_tile x ck] B tile (2-D TMA over the packed weight matrix).  The TMA swizzle (32/64/128 B = ck*2)
// matches the UMMA descriptor layout type, so a k-step inside a sub-tile is start address + 32 B.
#include "conv.h"
#include "launch_count.h"

#include <algorithm>
#include <mutex>
#include <vector>

namespace hrp {

// ------------------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------------------
constexpr int kNumThreads = 192;
constexpr int kStageABytes = kTileM * 64 * 2;  // 16 KiB of A per stage regardless of ck

struct __align__(16) PipeBarriers {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full;
  uint64_t res_full;
  uint32_t tmem_base;
  uint32_t pad;
};

// debug timeline: slot = (cta * 8 + tile_local - skip) * 16 + event for the first 8 CTAs and 8 consecutive tiles of each,
// starting at the CTA's tile number `skip` = tl[8 * 8 * 16] (0: pipeline fill, >= 8: steady state)
__device__ __forceinline__ void tl_stamp(long long* tl, int tile_local, int ev) {
  if (tl == nullptr || blockIdx.x >= 8) return;
  const int t = tile_local - (int)tl[8 * 8 * 16];
  if (t >= 0 && t < 8) tl[(blockIdx.x * 8 + t) * 16 + ev] = clock64();
}

// Column sums of an 8-value-per-lane tile over the 32 lanes of a warp in 9 shuffles instead of 40: at each of the first
// three butterfly levels a lane keeps the half of its values whose index bit matches its lane bit and sends the other half
// (8 -> 4 -> 2 -> 1 values), two plain levels finish.  Afterwards every lane holds the full sum of column
// ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); the tree is fixed, so the result is reproducible.
__device__ __forceinline__ float warp_colsum8(const float (&a)[8], int lane) {
  const bool h16 = (lane & 16) != 0;
  float k0 = h16 ? a[4] : a[0], k1 = h16 ? a[5] : a[1], k2 = h16 ? a[6] : a[2], k3 = h16 ? a[7] : a[3];
  const float s0 = h16 ? a[0] : a[4], s1 = h16 ? a[1] : a[5], s2 = h16 ? a[2] : a[6], s3 = h16 ? a[3] : a[7];
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  k2 += __shfl_xor_sync(0xffffffffu, s2, 16);
  k3 += __shfl_xor_sync(0xffffffffu, s3, 16);
  const bool h8 = (lane & 8) != 0;
  float m0 = h8 ? k2 : k0, m1 = h8 ? k3 : k1;
  const float t0 = h8 ? k0 : k2, t1 = h8 ? k1 : k3;
  m0 += __shfl_xor_sync(0xffffffffu, t0, 8);
  m1 += __shfl_xor_sync(0xffffffffu, t1, 8);
  const bool h4 = (lane & 4) != 0;
  float r = h4 ? m1 : m0;
  const float u = h4 ? m0 : m1;
  r += __shfl_xor_sync(0xffffffffu, u, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// Epilogue flavours (template parameter EPI): the common case carries no addends and no per-row predicates at all.
constexpr int EPI_PLAIN = 0;  // y = [relu](acc*scale + bias)
constexpr int EPI_PRE = 1;    // + up to three same-resolution addends before the ReLU (residual / fuse partials)
constexpr int EPI_FULL = 2;   // + nearest-upsampled addends, post-ReLU addend, pooled output
// (persistent kernel only) the same two flavours with EVERY addend TMA-staged in the ring entry: separate instantiations,
// because the generic kernel carrying both the staged and the gathering code paths is ~90 KB of SASS and its epilogue
// warps spent 40 % of their time in instruction-fetch stalls (ncu stall_no_inst, profiles/r02_ncu_stalls_s2fuse_b512.txt)
constexpr int EPI_PRE_ST = 5;
constexpr int EPI_FULL_ST = 6;
constexpr int EPI_RES = 3;    // + exactly one residual, TMA-prefetched into the staging tile (no per-row predicates or
                              //   global loads: the generic flavours spend ~60 issue slots per 8 channels on them)
constexpr int EPI_HEAD = 4;   // persistent kernel only: logits -> per-warp soft-argmax partials, nothing is stored
constexpr float kLog2eF = 1.4426950408889634f;

__device__ __forceinline__ float ex2_ftz_f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void add_bf16x8(float* v, const uint4& x) {
  v[0] += bf16lo_to_f32(x.x); v[1] += bf16hi_to_f32(x.x);
  v[2] += bf16lo_to_f32(x.y); v[3] += bf16hi_to_f32(x.y);
  v[4] += bf16lo_to_f32(x.z); v[5] += bf16hi_to_f32(x.z);
  v[6] += bf16lo_to_f32(x.w); v[7] += bf16hi_to_f32(x.w);
}
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <int CK, int EPI>
__global__ void __launch_bounds__(kNumThreads) conv_gemm_kernel(const __grid_constant__ ConvMaps maps,
                                                                const __grid_constant__ ConvParams p,
                                                                int stages, int bar_offset) {
  constexpr int SUB = 64 / CK;                 // sub-tiles (k-blocks) per stage
  constexpr int A_SUB_BYTES = kTileM * CK * 2;
  constexpr int KSTEPS = CK / 16;              // tcgen05.mma K=16 steps per sub-tile
  constexpr uint32_t LAYOUT = (CK == 64) ? 2u : (CK == 32) ? 4u : 6u;  // SW128 / SW64 / SW32
  constexpr uint32_t SBO = 8 * CK * 2;         // 8 rows of CK bf16

  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_align1024(smem_dyn);
  const int n_tile = p.n_tile;
  const int b_sub_bytes = n_tile * CK * 2;
  const int stage_bytes = p.vsh ? p.vsh_stage_bytes : kStageABytes + n_tile * 128;
  PipeBarriers* bars = reinterpret_cast<PipeBarriers*>(smem + bar_offset);
  float* sb_smem = reinterpret_cast<float*>(bars + 1);  // [2][n_tile] folded-BN scale / shift of this N tile

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates: the N tile is the fastest-varying block index so that CTAs sharing an A tile run together
  int t = blockIdx.x;
  int n_blk = 0;
  if (p.n_tiles > 1) {
    const int q_ = t / p.n_tiles;
    n_blk = t - q_ * p.n_tiles;
    t = q_;
  }
  const int tw = t & (p.tiles_w - 1);   // tiles_w, tiles_h, bw, bh are powers of two
  t >>= p.tw_shift;
  const int th = t & (p.tiles_h - 1);
  const int tn = t >> p.th_shift;
  const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
  const int phase = blockIdx.z;
  const int ph = phase >> 1, pw = phase & 1;  // deconv sub-pixel phase (0,0) when nphase == 1
  const int c_base = n_blk * n_tile;

  // residual (pre[0]) prefetched by TMA into a dedicated staging buffer behind the pipeline stages; without a
  // residual the staging buffer aliases the (by then idle) stages
  constexpr bool GENERIC = (EPI == EPI_PRE || EPI == EPI_FULL);
  const bool has_res = (EPI == EPI_RES) || (GENERIC && (p.pre[0] != nullptr) && (p.out != nullptr) && p.os == 1);
  uint8_t* const stag_base = has_res ? smem + (size_t)stages * stage_bytes : smem;
  const int nkb = p.ntaps * p.cpt;
  // vertical tap sharing (3x3 stride-1 convs): one iteration = (channel chunk, dw); its A buffer holds bh+2 image
  // rows and the three dh taps read 128-row windows of it (window start = dh * one image row: swizzle-aligned)
  const int n_iters = p.vsh ? 3 * p.cpt : (nkb + SUB - 1) / SUB;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->tmem_full, 1);
    mbar_init(&bars->res_full, 1);
    fence_mbar_init();
  }
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < n_tile) tmem_cols <<= 1;
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 2) {  // stage scale / shift once per CTA (overlaps the TMEM allocation)
    for (int i = threadIdx.x - 64; i < n_tile; i += 128) {
      const int c = c_base + i;
      sb_smem[i] = (c < p.Cout) ? __ldg(p.scale + c) : 0.f;
      sb_smem[n_tile + i] = (c < p.Cout) ? __ldg(p.bias + c) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_launch_dependents();
  pdl_wait();  // everything above (barriers, TMEM, scale/shift) overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      if (has_res) {  // residual tile -> dedicated staging buffer, in flight during the whole mainloop
        int nb = 0;
        for (int j = 0; j < n_tile / p.cko && c_base + j * p.cko < p.Cout; ++j) ++nb;
        mbar_expect_tx(&bars->res_full, (uint32_t)(nb * kTileM * p.cko * 2));
        for (int j = 0; j < nb; ++j)
          tma_load_4d(stag_base + (size_t)j * (kTileM * p.cko * 2), &maps.r, &bars->res_full, c_base + j * p.cko, w0, h0, n0);
      }
      // running counters instead of div/mod: this single thread's scalar latency IS the producer's throughput
      int s = 0, tap = 0, cc = 0;
      uint32_t par = 0;
      const int brow = (p.shared_phase ? 0 : phase * p.cout_pad) + c_base;
      const int tph = p.shared_phase ? 0 : ph, tpw = p.shared_phase ? 0 : pw;  // tap shift of the deconv phases
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(&bars->empty[s], par ^ 1);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        if (p.vsh) {
          const int cch = it / 3, dwi = it - cch * 3;
          mbar_expect_tx(&bars->full[s], (uint32_t)(p.vsh_a_bytes + 3 * b_sub_bytes));
          uint8_t* sb = sa + p.vsh_a_pad;
          tma_load_4d(sa, &maps.av, &bars->full[s], cch * CK, w0 + dwi - 1, h0 - 1, n0);
          for (int dhi = 0; dhi < 3; ++dhi)
            tma_load_2d(sb + dhi * b_sub_bytes, &maps.b, &bars->full[s], ((dhi * 3 + dwi) * p.cpt + cch) * CK, brow);
        } else {
          const int nsub = min(SUB, nkb - it * SUB);
          mbar_expect_tx(&bars->full[s], (uint32_t)(nsub * (A_SUB_BYTES + b_sub_bytes)));
          uint8_t* sb = sa + kStageABytes;
          for (int j = 0; j < nsub; ++j) {
            tma_load_4d(sa + j * A_SUB_BYTES, &maps.a[p.tap_map[tap]], &bars->full[s], cc * CK,
                        w0 + p.tap_dw[tap] + tpw, h0 + p.tap_dh[tap] + tph, n0);
            tma_load_2d(sb + j * b_sub_bytes, &maps.b, &bars->full[s], (it * SUB + j) * CK, brow);
            if (++cc == p.cpt) {
              cc = 0;
              ++tap;
            }
          }
        }
        if (++s == stages) {
          s = 0;
          par ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Warp-uniform control flow: descriptor words are computed by the whole warp (uniform datapath), only the
    // tcgen05 instructions are issued by one elected lane -- a divergent single-thread loop costs ~10 SASS
    // instructions (R2UR waterfalls) per MMA, more than a narrow MMA (N <= 64: 40-48 cycles) lasts.
    const uint32_t idesc = make_idesc_bf16(kTileM, (uint32_t)n_tile);
    const uint64_t dhi = make_kmajor_desc(0, SBO, LAYOUT) & 0xffffffff00000000ull;
    const uint32_t dlo = (uint32_t)make_kmajor_desc(0, SBO, LAYOUT);
    const uint32_t smem_units = dlo + (smem_u32(smem) >> 4);
    const uint32_t stage_units = (uint32_t)stage_bytes >> 4;
    const uint32_t bsub_units = (uint32_t)b_sub_bytes >> 4;
    int s = 0;
    uint32_t par = 0;
    for (int it = 0; it < n_iters; ++it) {
      mbar_wait(&bars->full[s], par);
      tc_fence_after();
      const uint32_t sa = smem_units + (uint32_t)s * stage_units;
      if (p.vsh) {
        const uint32_t sbv = sa + ((uint32_t)p.vsh_a_pad >> 4);
        const uint32_t row_units = (uint32_t)(p.bw * CK * 2) >> 4;  // one image row of the tile
        if (elect_one()) {
#pragma unroll
          for (int dhi_ = 0; dhi_ < 3; ++dhi_) {
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
              umma_bf16_ss(tmem_base, dhi | (sa + dhi_ * row_units + 2 * k), dhi | (sbv + dhi_ * bsub_units + 2 * k), idesc,
                           (it | dhi_ | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bars->empty[s]);
          if (it == n_iters - 1) umma_commit(&bars->tmem_full);
        }
      } else {
        const int nsub = min(SUB, nkb - it * SUB);
        const uint32_t sb = sa + (kStageABytes >> 4);
        if (elect_one()) {
          for (int j = 0; j < nsub; ++j) {
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
              umma_bf16_ss(tmem_base, dhi | (sa + j * (A_SUB_BYTES >> 4) + 2 * k), dhi | (sb + j * bsub_units + 2 * k), idesc,
                           (it | j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bars->empty[s]);                          // frees the smem stage when the MMAs retire
          if (it == n_iters - 1) umma_commit(&bars->tmem_full);  // accumulator complete
        }
      }
      __syncwarp();
      if (++s == stages) {
        s = 0;
        par ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 <-> TMEM lane quarters warp%4) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    // per-row addressing is only needed when addends are read or a pooled output is produced
    bool valid = true;
    size_t opix = 0;
    int n = 0, oh = 0, ow = 0;
    if (GENERIC) {
      const int wi = row & (p.bw - 1);
      const int hi = (row >> p.bw_shift) & (p.bh - 1);
      const int ni = row >> (p.bw_shift + p.bh_shift);
      n = n0 + ni;
      const int h = h0 + hi, w = w0 + wi;
      valid = (n < p.B) && (h < p.Hm) && (w < p.Wm);
      oh = h * p.os + p.oh0 + ph;
      ow = w * p.os + p.ow0 + pw;
      opix = ((size_t)n * p.Hout + oh) * p.Wout + ow;
    }
    const bool use_pre0 = GENERIC && !has_res && p.pre[0] != nullptr;  // (uniform) pre[0] read with plain loads
    const bf16* pre0 = (GENERIC && !has_res && p.pre[0] != nullptr) ? p.pre[0] + opix * p.Cout + c_base : nullptr;
    const bf16* pre1 = (GENERIC && p.pre[1] != nullptr) ? p.pre[1] + opix * p.Cout + c_base : nullptr;
    const bf16* pre2 = (GENERIC && p.pre[2] != nullptr) ? p.pre[2] + opix * p.Cout + c_base : nullptr;
    const bf16* upp[3] = {nullptr, nullptr, nullptr};
    const bf16* postp = nullptr;
    if (EPI == EPI_FULL) {
#pragma unroll
      for (int a = 0; a < 3; ++a)
        if (p.up[a] != nullptr) {
          const int sh = p.up_shift[a];
          const size_t upix = ((size_t)n * (p.Hout >> sh) + (oh >> sh)) * (p.Wout >> sh) + (ow >> sh);
          upp[a] = p.up[a] + upix * p.Cout + c_base;
        }
      if (p.post != nullptr) postp = p.post + opix * p.Cout + c_base;
    }
    const bool do_store = (EPI == EPI_RES || EPI == EPI_PLAIN) || (p.out != nullptr);
    const bool pool = (EPI == EPI_FULL) && (p.pool_out != nullptr);
    const bool relu_in_cvt = p.relu && !(EPI == EPI_FULL && (p.post != nullptr || pool));
    const bool relu_explicit = (EPI == EPI_FULL) && p.relu && !relu_in_cvt;
    const int cko = p.cko;
    const int cko_shift = (cko == 64) ? 6 : 5;
    const int sw = (cko == 64) ? (row & 7) : ((row >> 1) & 3);
    uint8_t* const stage_row = stag_base + (size_t)row * (cko * 2);
    // swizzled 16-byte chunk offsets of the four channel groups of a 32-column step (see chunk_base below)
    const uint32_t goff[4] = {(uint32_t)((0 ^ (sw & 3)) << 4), (uint32_t)((1 ^ (sw & 3)) << 4),
                              (uint32_t)((2 ^ (sw & 3)) << 4), (uint32_t)((3 ^ (sw & 3)) << 4)};
    const uint32_t blk_bytes = (uint32_t)(kTileM * cko * 2);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int c_lim = min(n_tile, p.Cout - c_base);  // channels of this N tile that exist (multiple of 32)

    mbar_wait(&bars->tmem_full, 0);
    tc_fence_after();
    if (has_res) mbar_wait(&bars->res_full, 0);

#pragma unroll 1
    for (int c0 = 0; c0 < c_lim; c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(taddr + (uint32_t)c0, acc);
      tmem_ld_wait();
      // staging address of this 32-column step: column block (cko channels x 128 rows), then the 16-byte chunk
      // (c0 % cko) / 8 + g XOR-swizzled by the row: the high chunk bit is folded in here, the low two in goff[g]
      uint8_t* const chunk_base =
          stage_row + (size_t)(c0 >> cko_shift) * blk_bytes + (((uint32_t)((c0 & (cko - 1)) >> 3) ^ (uint32_t)(sw & 4)) << 4);
#pragma unroll
      for (int g = 0; g < 4; ++g) {  // 8 channels at a time
        const int cg = c0 + g * 8;
        float v[8];
        const float4 s0 = *reinterpret_cast<const float4*>(sb_smem + cg);
        const float4 s1 = *reinterpret_cast<const float4*>(sb_smem + cg + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(sb_smem + n_tile + cg);
        const float4 b1 = *reinterpret_cast<const float4*>(sb_smem + n_tile + cg + 4);
        v[0] = fmaf(__uint_as_float(acc[g * 8 + 0]), s0.x, b0.x);
        v[1] = fmaf(__uint_as_float(acc[g * 8 + 1]), s0.y, b0.y);
        v[2] = fmaf(__uint_as_float(acc[g * 8 + 2]), s0.z, b0.z);
        v[3] = fmaf(__uint_as_float(acc[g * 8 + 3]), s0.w, b0.w);
        v[4] = fmaf(__uint_as_float(acc[g * 8 + 4]), s1.x, b1.x);
        v[5] = fmaf(__uint_as_float(acc[g * 8 + 5]), s1.y, b1.y);
        v[6] = fmaf(__uint_as_float(acc[g * 8 + 6]), s1.z, b1.z);
        v[7] = fmaf(__uint_as_float(acc[g * 8 + 7]), s1.w, b1.w);
        uint4* const sptr = reinterpret_cast<uint4*>(chunk_base + goff[g]);
        if (EPI != EPI_PLAIN && has_res) add_bf16x8(v, *sptr);  // TMA-prefetched residual (zero outside the tensor)
        if (GENERIC && valid) {
          // the tests are on kernel parameters (uniform registers): absent addends cost a uniform branch, not a
          // predicated copy of the load / unpack / add sequence
          if (use_pre0) add_bf16x8(v, __ldg(reinterpret_cast<const uint4*>(pre0 + cg)));
          if (p.pre[1] != nullptr) add_bf16x8(v, __ldg(reinterpret_cast<const uint4*>(pre1 + cg)));
          if (p.pre[2] != nullptr) add_bf16x8(v, __ldg(reinterpret_cast<const uint4*>(pre2 + cg)));
          if (EPI == EPI_FULL) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
              if (p.up[a] != nullptr) add_bf16x8(v, __ldg(reinterpret_cast<const uint4*>(upp[a] + cg)));
          }
        }
        if (relu_explicit) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (EPI == EPI_FULL && p.post != nullptr && valid) add_bf16x8(v, __ldg(reinterpret_cast<const uint4*>(postp + cg)));
        if (do_store) {
          uint4 o;
          if (relu_in_cvt) {
            o.x = pack_bf16x2_relu(v[0], v[1]);
            o.y = pack_bf16x2_relu(v[2], v[3]);
            o.z = pack_bf16x2_relu(v[4], v[5]);
            o.w = pack_bf16x2_relu(v[6], v[7]);
          } else {
            o.x = pack_bf16x2(v[0], v[1]);
            o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]);
            o.w = pack_bf16x2(v[6], v[7]);
          }
          // stage into the (now idle) pipeline buffers in the TMA-store box layout: column blocks of cko
          // channels x 128 rows, 16-byte chunks XOR-swizzled exactly as the output tensor map expects
          *sptr = o;
        }
        if (pool) {
          // global average pool: the 32 rows of a warp belong to one image (host checks bw*bh >= 32).  Bitwise
          // reproducibility: the within-warp tree is fixed, and on the path the pooled maps are 8x8 = two warps per
          // image, i.e. exactly two atomicAdd contributions per (image, channel) onto a zeroed accumulator -- fp32
          // addition is commutative, so their arrival order cannot change the result (tests: bitwise replay)
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) pv[i] = valid ? v[i] : 0.f;
          const float csum = warp_colsum8(pv, lane);
          const int n_warp = n0 + ((q * 32) >> (p.bw_shift + p.bh_shift));
          if ((lane & 3) == 0 && n_warp < p.B)
            atomicAdd(p.pool_out + (size_t)n_warp * p.Cout + c_base + cg + (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 +
                                                                            ((lane >> 2) & 1)),
                      csum * p.pool_scale);
        }
      }
    }
    tc_fence_before();
    if (do_store) {
      // generic-proxy smem writes -> visible to the async proxy, then ONE thread issues the TMA stores
      // (rows / channels outside the tensor are clipped by the TMA unit: ragged tiles need no predicates)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2 && elect_one()) {
        const int nblk = n_tile / cko;
        for (int j = 0; j < nblk; ++j) {
          const int cj = c_base + j * cko;
          if (cj >= p.Cout) break;
          tma_store_4d(stag_base + (size_t)j * (kTileM * cko * 2), &maps.o[phase], cj, w0, h0, n0);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM loops over output tiles.  The TMA producers run ahead across tile
// boundaries, the accumulators form a ring of two or four tiles in TMEM (the MMAs of the next tiles overlap the epilogue
// of tile i), the residual / addend tiles are prefetched by TMA into the output staging ring and updated in place, and
// eight epilogue warps (two per TMEM lane quarter) drain the accumulators.  Per-tile fixed costs (TMEM allocation,
// barrier init, descriptor fetch, launch) are paid once per CTA.
// ------------------------------------------------------------------------------------------------------
// warp 0 producer, warp 1 MMA, warps 2..9 epilogue, warp 10 TMA-store (+ addend fetches), warp 11 second producer,
// warp 12 second MMA issuer (flavours with PersistShape::has_m2).
// Two producers: ONE thread running the ring protocol (wait for the slot, expect_tx, cp.async.bulk.tensor, bookkeeping)
// sustains one load per ~550-650 cycles whatever the box size (tools/probe_tma.py, profiles/r02_probe_tma.txt: the raw
// instruction issues every ~75 cycles and the engine delivers > 70 B/cycle/SM, but the per-slot handshake serialises the
// thread), i.e. 16 KiB stages arrive at ~26 B/cycle/SM while the MMAs of a stage (N = 128, K = 64) take 257 cycles: every
// persistent kernel with short stages was bound by its producer THREAD.  Two threads in two warps alternate stages and
// reach ~48 B/cycle/SM.
constexpr int kThreadsP = 384;
constexpr int kEpiThreadsP = 256;
// The soft-argmax fold (EPI_HEAD) has no store warp.  History (profiles/r02_exp_final_fold.txt): with ONE (producer, ring,
// issuer) pipeline the mainloop bounded the layer (848 of 960 us with the epilogue reduced to the accumulator read; eight
// vs sixteen epilogue warps made no difference).  With the two pipelines below the mainloop alone runs in 444-550 us and the
// epilogue -- a latency chain of ~11 dependent shuffles per 32-column block, IPC 0.5 with two warps per scheduler -- became
// the limit: sixteen epilogue warps (one 32-column block per warp and tile, 96 registers) with 32 KiB stages: 990 -> 810 us.
constexpr int kEpiWarpsHead = 16;
// soft-argmax fold: warp 0 producer, warp 1 MMA issuer, warps 2..17 epilogue, warp 18 second producer, warp 19 second MMA
// issuer (no store warp).  Two issuers = two independent pipelines (PersistCfg::dual): even tiles run through producer 0
// -> first half of the stage ring -> issuer A, odd tiles through producer 1 -> second half -> issuer B; the TMEM
// accumulators alternate by tile anyway.  Each serial actor pays its ~600-900 cycles per stage handshake on every
// second tile only.
constexpr int kThreadsPHead = 64 + 32 * kEpiWarpsHead + 64;
// Flavours whose register count leaves room for a 13th warp (plain: 155 registers x 416 threads) get the second MMA
// issuer as well (PersistCfg::dual).
template <int EPI> struct PersistShape {
  static constexpr int epi_warps = (EPI == 4) ? kEpiWarpsHead : kEpiThreadsP / 32;
  static constexpr bool has_m2 = (EPI == 4 || EPI == 0 || EPI == 3 || EPI == 5 || EPI == 6);  // not the gathering flavours (164 registers)
  static constexpr int threads = (EPI == 4) ? kThreadsPHead : (has_m2 ? kThreadsP + 32 : kThreadsP);
};

struct __align__(16) PersistBarriers {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[4];
  uint64_t tmem_empty[4];
  uint64_t res_full[4];
  uint64_t stag_free[4];
  uint64_t stag_ready[4];
  uint64_t w_full;
  uint32_t tmem_base;
  uint32_t pad;
};

template <int CK, int EPI>
__global__ void __launch_bounds__(PersistShape<EPI>::threads, 1) conv_gemm_persistent(const __grid_constant__ ConvMaps maps,
                                                                     const __grid_constant__ ConvParams p,
                                                                     const PersistCfg cfg) {
  constexpr int A_SUB_BYTES = kTileM * CK * 2;
  constexpr int KSTEPS = CK / 16;
  constexpr uint32_t LAYOUT = (CK == 64) ? 2u : (CK == 32) ? 4u : 6u;
  constexpr uint32_t SBO = 8 * CK * 2;
  constexpr int WARP_P2 = (EPI == EPI_HEAD) ? 2 + kEpiWarpsHead : 11;      // second producer warp
  constexpr int WARP_M2 = (EPI == EPI_HEAD) ? 3 + kEpiWarpsHead : (PersistShape<EPI>::has_m2 ? 12 : -1);  // second MMA issuer
  const bool dual = PersistShape<EPI>::has_m2 && cfg.dual != 0;            // two independent (producer, ring, issuer) pipelines
  const int ring = dual ? cfg.stages / 2 : cfg.stages;                     // slots per pipeline
  const int SUB = cfg.sub;                     // k-blocks (CK channels of one tap) per stage: (64 / CK) x 1 or 2

  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_align1024(smem_dyn);
  const int n_tile = p.n_tile;
  const int b_sub_bytes = n_tile * CK * 2;
  const int stage_bytes = cfg.stage_bytes;      // A region (+ B region unless the weights are resident)
  const int a_region = cfg.a_region;
  const int stag_bytes = kTileM * n_tile * 2;
  const int stages = cfg.stages, nstag = cfg.nstag;  // nstag: 1, 2 or 4 output staging buffers (ring)
  const int nstag_shift = (nstag == 4) ? 2 : (nstag == 2) ? 1 : 0;
  const bool vsh = cfg.vsh != 0, wres = cfg.wres != 0;
  const int ksplit = cfg.ksplit;               // accumulators per tile (1, 2 or 4)
  // accumulator ring in TMEM: 2 or 4 tiles (4 when 4 x ksplit x n_tile <= 512 columns): the MMA issuers run up to three
  // tiles ahead of the epilogue instead of one -- with two pipelines each of them owns two accumulators
  const int nacc = cfg.nacc, nacc_shift = (nacc == 4) ? 2 : 1;
  const uint32_t ksmask = (uint32_t)ksplit - 1u;
  uint8_t* const pipe_base = smem + cfg.pipe_offset;  // resident weights (if any) live in front of the stages
  uint8_t* const stag_base = smem + cfg.stag_offset;
  PersistBarriers* bars = reinterpret_cast<PersistBarriers*>(smem + cfg.bar_offset);
  float* sb_smem = reinterpret_cast<float*>(bars + 1);  // [2][cout_pad] scale / shift of the whole layer

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkb = p.ntaps * p.cpt;
  const int n_iters = vsh ? 3 * p.cpt : (nkb + SUB - 1) / SUB;
  const int cko = p.cko;
  const int nblk_full = n_tile / cko;
  constexpr bool staged = (EPI == EPI_PRE_ST || EPI == EPI_FULL_ST);  // every addend is a TMA box in the ring entry (PersistCfg)
  constexpr bool FULLISH = (EPI == EPI_FULL || EPI == EPI_FULL_ST);
  constexpr bool GENERIC = (EPI == EPI_PRE || EPI == EPI_FULL || staged);
  constexpr bool ROLLED = staged;   // epilogue walks 8 columns at a time in a rolled loop (see the epilogue)
  const bool has_res = (EPI == EPI_RES) || staged || (GENERIC && cfg.res_tma != 0);
  const int entry_bytes = cfg.entry_bytes;          // ring entry = output / pre[0] slot (stag_bytes) + addend slots
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], PersistShape<EPI>::epi_warps);  // one arrival per epilogue warp
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars->res_full[i], 1);
      mbar_init(&bars->stag_free[i], 1);
      mbar_init(&bars->stag_ready[i], kEpiThreadsP / 32);
    }
    mbar_init(&bars->w_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, (uint32_t)cfg.tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 2 && warp < 2 + PersistShape<EPI>::epi_warps) {
    const float mul = (EPI == EPI_HEAD) ? kLog2eF : 1.0f;  // soft-argmax fold: logits in the log2 domain (ex2 below)
    for (int i = threadIdx.x - 64; i < p.cout_pad; i += 32 * PersistShape<EPI>::epi_warps) {
      sb_smem[i] = (i < p.Cout) ? __ldg(p.scale + i) * mul : 0.f;
      sb_smem[p.cout_pad + i] = (i < p.Cout) ? __ldg(p.bias + i) * mul : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_launch_dependents();
  if (!(warp == 0 && wres)) pdl_wait();  // (the producer first fetches the resident weights: not produced by a kernel)

  // tile -> (n tile, phase, w/h/n tile) with shifts: tiles_w and tiles_h are powers of two by construction
#define HRP_DECODE_TILE(tile)                                          \
  int t_ = (tile);                                                     \
  int n_blk = 0;                                                       \
  if (p.n_tiles > 1) {                                                 \
    const int q_ = t_ / p.n_tiles;                                     \
    n_blk = t_ - q_ * p.n_tiles;                                       \
    t_ = q_;                                                           \
  }                                                                    \
  int phase = 0;                                                       \
  if (p.nphase > 1) {                                                  \
    phase = t_ / tiles_m;                                              \
    t_ -= phase * tiles_m;                                             \
  }                                                                    \
  const int tw = t_ & (p.tiles_w - 1);                                 \
  t_ >>= cfg.tw_shift;                                                 \
  const int th = t_ & (p.tiles_h - 1);                                 \
  const int tn = t_ >> cfg.th_shift;                                   \
  const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;            \
  const int ph = phase >> 1, pw = phase & 1;                           \
  const int c_base = n_blk * n_tile;

  // TMA fetch of tile `tile_`'s residual / addend tiles into staging-ring entry `sbuf_` (the caller knows the entry is free)
  auto fetch_addends = [&](int tile_, int sbuf_) {
    HRP_DECODE_TILE(tile_)
    (void)th; (void)tw; (void)tn; (void)ph; (void)pw;
    int nb = 0;
    for (int j = 0; j < nblk_full && c_base + j * cko < p.Cout; ++j) ++nb;
    uint8_t* const entry = stag_base + (size_t)sbuf_ * entry_bytes;
    const int blk = kTileM * cko * 2;
    uint32_t bytes = (uint32_t)(nb * blk);
    if (staged) {
#pragma unroll 1
      for (int i = 0; i < 3; ++i)
        if (cfg.add_off[i] != 0) bytes += (uint32_t)(nb * blk);
#pragma unroll 1
      for (int a = 0; a < 3; ++a)
        if (cfg.up_off[a] != 0) bytes += (uint32_t)(nb * cfg.up_bw[a] * cfg.up_bh[a] * p.bn * cko * 2);
    }
    mbar_expect_tx(&bars->res_full[sbuf_], bytes);
    const CUtensorMap* mr = (phase == 0) ? &maps.r : &maps.rp[phase - 1];
#pragma unroll 1
    for (int j = 0; j < nb; ++j)
      tma_load_4d(entry + (size_t)j * blk, mr, &bars->res_full[sbuf_], c_base + j * cko, w0, h0, n0);
    if (staged) {
#pragma unroll 1
      for (int i = 0; i < 3; ++i)
        if (cfg.add_off[i] != 0)
#pragma unroll 1
          for (int j = 0; j < nb; ++j)
            tma_load_4d(entry + cfg.add_off[i] + (size_t)j * blk, &maps.add[i], &bars->res_full[sbuf_], c_base + j * cko, w0,
                        h0, n0);
#pragma unroll 1
      for (int a = 0; a < 3; ++a)
        if (cfg.up_off[a] != 0)
#pragma unroll 1
          for (int j = 0; j < nb; ++j)
            tma_load_4d(entry + cfg.up_off[a] + (size_t)j * cfg.up_blk[a], &maps.upm[a], &bars->res_full[sbuf_],
                        c_base + j * cko, w0 >> cfg.up_sh[a], h0 >> cfg.up_sh[a], n0);
    }
  };
  // Who fetches them: the TMA-store warp (cfg.res_store, default) -- it is the thread that learns first that an entry has
  // drained, and both producers stay free for the two operand rings --, else producer 1 (or the only producer).
  const bool res_store = has_res && cfg.res_store != 0;

  if (warp == 0 || warp == WARP_P2) {
    // ===================== TMA producers: warp 0 (+ warp WARP_P2 when cfg.nprod == 2) =====================
    // Stage g of the CTA (counted over all its tiles) is issued by producer g % nprod; both walk the same sequence and
    // skip the other one's stages.  Producer 0 also fetches the resident weights and the residual / addend tiles.
    const int me = (warp == 0) ? 0 : 1;
    if (me < cfg.nprod && elect_one()) {
      // Layers with TMA-fetched residual / addend tiles split the roles instead: producer 0 streams the operands,
      // producer 1 fetches the addend tiles -- it is the one that has to wait for a staging entry to drain, and the
      // operand stream no longer stops behind that wait.
      const bool res_role = has_res && !res_store && cfg.nprod == 2;
      const bool two = cfg.nprod == 2 && !res_role && !dual;
      const int res_owner = res_role ? 1 : 0;
      const int sbase = (dual && me == 1) ? ring : 0;   // dual: this producer's half of the ring
      uint32_t g = 0;
      if (wres && me == 0) {
        // the packed weights of this CTA's N tile stay in shared memory for the CTA's lifetime.  With several N tiles the
        // grid is a multiple of n_tiles (host), so every tile this CTA walks (blockIdx.x + i * gridDim.x) has the same
        // n_blk = blockIdx.x % n_tiles: the A tiles stream through the pipeline alone and the L2 -> SM traffic halves.
        const int wrow0 = (p.n_tiles > 1) ? (int)(blockIdx.x % (unsigned)p.n_tiles) * n_tile : 0;
        mbar_expect_tx(&bars->w_full, (uint32_t)(nkb * b_sub_bytes));
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(smem + (size_t)kb * b_sub_bytes, &maps.b, &bars->w_full, kb * CK, wrow0);
        pdl_wait();
      }
      int s = 0;
      uint32_t par = 0;
      int li = 0;
      for (int tile = blockIdx.x; tile < cfg.total_tiles; tile += gridDim.x, ++li) {
        HRP_DECODE_TILE(tile)
        (void)th; (void)tw; (void)tn;
        const int sbuf = li & (nstag - 1);
        const uint32_t spar = (uint32_t)((li >> nstag_shift) & 1);
        const int brow = (p.shared_phase ? 0 : phase * p.cout_pad) + c_base;
        const int tph = p.shared_phase ? 0 : ph, tpw = p.shared_phase ? 0 : pw;
        int tap = 0, cc = 0, dwi = 0;
        if (dual && (li & 1) != me) continue;   // the other pipeline's tile
        auto load_residual = [&]() {
          mbar_wait(&bars->stag_free[sbuf], spar ^ 1);  // the store that last used this entry has drained
          fetch_addends(tile, sbuf);
        };
        if (has_res && !res_store && nstag >= 2 && me == res_owner) load_residual();
        if (res_role && me == 1) {
          if (nstag == 1) load_residual();
          continue;
        }
        if (me == 0) tl_stamp(p.timeline, li, 0);
        for (int it = 0; it < n_iters; ++it, ++g) {
          const bool mine = !two || (int)(g & 1u) == me;
          const int sl = sbase + s;
          if (mine) {
            mbar_wait(&bars->empty[sl], par ^ 1);
            if (it == 0 && me == 0) tl_stamp(p.timeline, li, 1);
          }
          uint8_t* sa = pipe_base + (size_t)sl * stage_bytes;
          uint8_t* sb = sa + a_region;
          if (vsh) {
            // one (channel chunk, dw) per iteration: an A buffer of bh+2 image rows serves the three dh taps
            if (mine) {
              mbar_expect_tx(&bars->full[sl], (uint32_t)(cfg.a_bytes + (wres ? 0 : 3 * b_sub_bytes)));
              tma_load_4d(sa, &maps.av, &bars->full[sl], cc * CK, w0 + dwi - 1, h0 - 1, n0);
              if (!wres)
                for (int dhi = 0; dhi < 3; ++dhi)
                  tma_load_2d(sb + dhi * b_sub_bytes, &maps.b, &bars->full[sl], ((dhi * 3 + dwi) * p.cpt + cc) * CK, brow);
            }
            if (++dwi == 3) {
              dwi = 0;
              ++cc;
            }
          } else {
            const int nsub = min(SUB, nkb - it * SUB);
            if (mine) mbar_expect_tx(&bars->full[sl], (uint32_t)(nsub * (A_SUB_BYTES + (wres ? 0 : b_sub_bytes))));
            for (int j = 0; j < nsub; ++j) {
              if (mine) {
                tma_load_4d(sa + j * A_SUB_BYTES, &maps.a[p.tap_map[tap]], &bars->full[sl], cc * CK,
                            w0 + p.tap_dw[tap] + tpw, h0 + p.tap_dh[tap] + tph, n0);
                if (!wres) tma_load_2d(sb + j * b_sub_bytes, &maps.b, &bars->full[sl], (it * SUB + j) * CK, brow);
              }
              if (++cc == p.cpt) {
                cc = 0;
                ++tap;
              }
            }
          }
          if (++s == ring) {
            s = 0;
            par ^= 1;
          }
        }
        if (me == 0) tl_stamp(p.timeline, li, 2);
        if (has_res && !res_store && nstag == 1 && me == res_owner) load_residual();
      }
    }
  } else if (warp == 1 || warp == WARP_M2) {
    // ===================== MMA issuer(s): warp-uniform loop, one elected lane issues the tcgen05 instructions ==========
    const int mw = (warp == 1) ? 0 : 1;
    if (mw == 0 || dual) {
      const int sbase = (dual && mw == 1) ? ring : 0;
      const uint32_t idesc = make_idesc_bf16(kTileM, (uint32_t)n_tile);
      // descriptor = constant high word | (constant low word + smem byte address >> 4): one 32-bit add per operand
      // on the uniform datapath (shared-memory addresses stay below 2^18, so the 14-bit field never carries)
      const uint64_t dhi = make_kmajor_desc(0, SBO, LAYOUT) & 0xffffffff00000000ull;
      const uint32_t dlo = (uint32_t)make_kmajor_desc(0, SBO, LAYOUT);
      const uint32_t pipe_units = dlo + (smem_u32(pipe_base) >> 4);
      const uint32_t wres_units = dlo + (smem_u32(smem) >> 4);
      const uint32_t stage_units = (uint32_t)stage_bytes >> 4;
      const uint32_t aregion_units = (uint32_t)a_region >> 4;
      const uint32_t row_units = (uint32_t)(p.bw * CK * 2) >> 4;   // one image row of the tile, in 16-byte units
      const uint32_t bsub_units = (uint32_t)b_sub_bytes >> 4;
      if (wres) mbar_wait(&bars->w_full, 0);
      int s = 0;
      uint32_t par = 0;
      int li = 0;
      for (int tile = blockIdx.x; tile < cfg.total_tiles; tile += gridDim.x, ++li) {
        if (dual && (li & 1) != mw) continue;   // the other pipeline's tile
        const int abuf = li & (nacc - 1);
        mbar_wait(&bars->tmem_empty[abuf], (uint32_t)(((li >> nacc_shift) & 1) ^ 1));  // epilogue drained this accumulator
        tc_fence_after();
        if (lane == 0 && mw == 0) tl_stamp(p.timeline, li, 3);
        // K steps round-robin over `ksplit` accumulators (summed by the epilogue)
        const uint32_t tacc0 = tmem_base + (uint32_t)(abuf * ksplit * n_tile);
        int cc = 0, dwi = 0;
        uint32_t mi = 0;
        for (int it = 0; it < n_iters; ++it) {
          const int sl = sbase + s;
          mbar_wait(&bars->full[sl], par);
          tc_fence_after();
          if (it == 0 && lane == 0 && mw == 0) tl_stamp(p.timeline, li, 4);
          const uint32_t a_it = pipe_units + (uint32_t)sl * stage_units;
          const uint32_t b_it = a_it + aregion_units;
          if (vsh) {
            const uint32_t bw_it = wres_units + (uint32_t)(dwi * p.cpt + cc) * bsub_units;
            const uint32_t bw_step = (uint32_t)(3 * p.cpt) * bsub_units;
            if (elect_one()) {
#pragma unroll
              for (int dhi_ = 0; dhi_ < 3; ++dhi_) {
                const uint32_t ad = a_it + dhi_ * row_units;
                const uint32_t bd = wres ? bw_it + dhi_ * bw_step : b_it + dhi_ * bsub_units;
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  umma_bf16_ss(tacc0 + (mi & ksmask) * (uint32_t)n_tile, dhi | (ad + 2 * k), dhi | (bd + 2 * k), idesc,
                               mi >= (uint32_t)ksplit);
                  ++mi;
                }
              }
              umma_commit(&bars->empty[sl]);
              if (it == n_iters - 1) umma_commit(&bars->tmem_full[abuf]);
            }
            mi = (uint32_t)((it + 1) * 3 * KSTEPS);
            if (++dwi == 3) {
              dwi = 0;
              ++cc;
            }
          } else {
            const int nsub = min(SUB, nkb - it * SUB);
            const uint32_t bw_it = wres_units + (uint32_t)(it * SUB) * bsub_units;
            const uint32_t b0 = wres ? bw_it : b_it;
            if (elect_one()) {
              for (int j = 0; j < nsub; ++j) {
                const uint32_t ad = a_it + (uint32_t)j * (A_SUB_BYTES >> 4);
                const uint32_t bd = b0 + (uint32_t)j * bsub_units;
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  umma_bf16_ss(tacc0 + (mi & ksmask) * (uint32_t)n_tile, dhi | (ad + 2 * k), dhi | (bd + 2 * k), idesc,
                               mi >= (uint32_t)ksplit);
                  ++mi;
                }
              }
              umma_commit(&bars->empty[sl]);
              if (it == n_iters - 1) umma_commit(&bars->tmem_full[abuf]);
            }
            mi = (uint32_t)(min((it + 1) * SUB, nkb) * KSTEPS);
          }
          __syncwarp();
          if (it == n_iters - 1 && lane == 0 && mw == 0) tl_stamp(p.timeline, li, 5);
          if (++s == ring) {
            s = 0;
            par ^= 1;
          }
        }
      }
    }
  } else if (EPI != EPI_HEAD && warp == 10) {
    // ===================== TMA-store warp: drains finished staging buffers, never stalls the epilogue ==========
    if (p.out != nullptr && elect_one()) {
      if (res_store) {  // the first nstag tiles find their entries free
        int t0 = blockIdx.x;
        for (int b = 0; b < nstag && t0 < cfg.total_tiles; ++b, t0 += gridDim.x) fetch_addends(t0, b);
      }
      int li = 0;
      for (int tile = blockIdx.x; tile < cfg.total_tiles; tile += gridDim.x, ++li) {
        HRP_DECODE_TILE(tile)
        (void)th; (void)tw; (void)tn; (void)ph; (void)pw;
        const int sbuf = li & (nstag - 1);
        mbar_wait(&bars->stag_ready[sbuf], (uint32_t)((li >> nstag_shift) & 1));
        tl_stamp(p.timeline, li, 10);
        for (int j = 0; j < nblk_full; ++j) {
          const int cj = c_base + j * cko;
          if (cj >= p.Cout) break;
          tma_store_4d(stag_base + (size_t)sbuf * entry_bytes + (size_t)j * (kTileM * cko * 2), &maps.o[phase], cj, w0,
                       h0, n0);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        tl_stamp(p.timeline, li, 11);
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem read by the TMA unit: buffer reusable
        tl_stamp(p.timeline, li, 12);
        if (res_store) {  // the entry has drained: fetch the addends of the tile that will be built in it next
          const int nt = tile + nstag * (int)gridDim.x;
          if (nt < cfg.total_tiles) fetch_addends(nt, sbuf);
        } else {
          mbar_arrive(&bars->stag_free[sbuf]);
        }
      }
    }
  } else {
    // ===================== epilogue: warps 2..9, lane quarter = warp % 4, column half = (warp-2)/4 ==========
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int cko_shift = (cko == 64) ? 6 : 5;
    const int sw = (cko == 64) ? (row & 7) : ((row >> 1) & 3);
    const uint32_t goff[4] = {(uint32_t)((0 ^ (sw & 3)) << 4), (uint32_t)((1 ^ (sw & 3)) << 4),
                              (uint32_t)((2 ^ (sw & 3)) << 4), (uint32_t)((3 ^ (sw & 3)) << 4)};
    const uint32_t blk_bytes = (uint32_t)(kTileM * cko * 2);
    const bool do_store = (EPI == EPI_RES || EPI == EPI_PLAIN) || (p.out != nullptr);
    const bool pool = (EPI == EPI_FULL) && (p.pool_out != nullptr);   // (never together with staged addends)
    int li = 0;
    for (int tile = blockIdx.x; tile < cfg.total_tiles; tile += gridDim.x, ++li) {
      HRP_DECODE_TILE(tile)
      (void)th; (void)tw; (void)tn;
      const int abuf = li & (nacc - 1);
      const int sbuf = li & (nstag - 1);
      const uint32_t spar = (uint32_t)((li >> nstag_shift) & 1);
      bool valid = true;
      size_t opix = 0;
      int n = 0, oh = 0, ow = 0;
      if (GENERIC) {
        const int wi = row & (p.bw - 1);
        const int hi = (row >> p.bw_shift) & (p.bh - 1);
        const int ni = row >> (p.bw_shift + p.bh_shift);
        n = n0 + ni;
        const int h = h0 + hi, w = w0 + wi;
        valid = (n < p.B) && (h < p.Hm) && (w < p.Wm);
        oh = h * p.os + p.oh0 + ph;
        ow = w * p.os + p.ow0 + pw;
        opix = ((size_t)n * p.Hout + oh) * p.Wout + ow;
      }
      const bool use_pre0 = GENERIC && !has_res && p.pre[0] != nullptr;
      const bf16* pre0 = (GENERIC && !has_res && p.pre[0] != nullptr) ? p.pre[0] + opix * p.Cout + c_base : nullptr;
      const bf16* pre1 = (GENERIC && p.pre[1] != nullptr) ? p.pre[1] + opix * p.Cout + c_base : nullptr;
      const bf16* pre2 = (GENERIC && p.pre[2] != nullptr) ? p.pre[2] + opix * p.Cout + c_base : nullptr;
      const bf16* upp[3] = {nullptr, nullptr, nullptr};
      const bf16* postp = nullptr;
      if (EPI == EPI_FULL) {   // (gathering flavour only)
#pragma unroll
        for (int a = 0; a < 3; ++a)
          if (p.up[a] != nullptr) {
            const int sh = p.up_shift[a];
            const size_t upix = ((size_t)n * (p.Hout >> sh) + (oh >> sh)) * (p.Wout >> sh) + (ow >> sh);
            upp[a] = p.up[a] + upix * p.Cout + c_base;
          }
        if (p.post != nullptr) postp = p.post + opix * p.Cout + c_base;
      }
      const bool relu_in_cvt = p.relu && !(FULLISH && (p.post != nullptr || pool));
      const bool relu_explicit = FULLISH && p.relu && !relu_in_cvt;
      uint8_t* const stage_row = stag_base + (size_t)sbuf * entry_bytes + (size_t)row * (cko * 2);
      // staged nearest-upsampled addends: this thread's pixel of the low-resolution box and its swizzle phase
      const uint8_t* up_row[3] = {nullptr, nullptr, nullptr};
      uint32_t up_sw[3] = {0, 0, 0};
      if (staged) {
        const int wi = row & (p.bw - 1);
        const int hi = (row >> p.bw_shift) & (p.bh - 1);
        const int ni = row >> (p.bw_shift + p.bh_shift);
#pragma unroll
        for (int a = 0; a < 3; ++a)
          if (cfg.up_off[a] != 0) {
            const int r2 = (ni * cfg.up_bh[a] + (hi >> cfg.up_sh[a])) * cfg.up_bw[a] + (wi >> cfg.up_sh[a]);
            up_row[a] = stag_base + (size_t)sbuf * entry_bytes + cfg.up_off[a] + (size_t)r2 * (cko * 2);
            up_sw[a] = (uint32_t)((cko == 64) ? (r2 & 7) : ((r2 >> 1) & 3));
          }
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(abuf * ksplit * n_tile);
      const int c_lim = min(n_tile, p.Cout - c_base);
      const float* sc = sb_smem + c_base;
      const float* sh_ = sb_smem + p.cout_pad + c_base;

      if (warp == 2 && lane == 0) tl_stamp(p.timeline, li, 6);
      mbar_wait(&bars->tmem_full[abuf], (uint32_t)((li >> nacc_shift) & 1));
      tc_fence_after();
      if (warp == 2 && lane == 0) tl_stamp(p.timeline, li, 7);
      if (do_store) {
        if (has_res) mbar_wait(&bars->res_full[sbuf], spar);          // residual tile landed in the staging buffer
        else mbar_wait(&bars->stag_free[sbuf], spar ^ 1);            // previous store from this buffer has drained
      }
      if (warp == 2 && lane == 0) tl_stamp(p.timeline, li, 8);

#pragma unroll 1
      for (int c0 = half * 32; c0 < c_lim; c0 += 8 * PersistShape<EPI>::epi_warps) {
        uint32_t acc[ROLLED ? 8 : 32];
        if (ROLLED) {
          tmem_ld8(taddr + (uint32_t)c0, reinterpret_cast<uint32_t(&)[8]>(acc[0]));  // (rolled 8-column loop below)
        } else {
          tmem_ld32(taddr + (uint32_t)c0, reinterpret_cast<uint32_t(&)[32]>(acc[0]));
          tmem_ld_wait();
        }
        if (EPI == EPI_HEAD) {
          if (p.dbg & 4) continue;  // (ablation: accumulator read only)
          // Soft-argmax fold.  This thread owns pixel (h, w) of image n0 (tile = two image rows: bw 64, bh 2, bn 1) and the
          // 32 depth bins d0..d0+31 of keypoint kp: online-softmax partial over its 32 logits (log2 domain), then a
          // fixed-order butterfly over the warp's 32 pixels; lane 0 writes the 5-tuple.  Nothing is stored otherwise.
          const int cabs = c_base + c0;
          const int kp = cabs >> 6;
          const float d0 = (float)(cabs & 63);
          const float fw = (float)(w0 + (row & (p.bw - 1))), fh = (float)(h0 + (row >> p.bw_shift));
          float t[32];
          // four independent chains everywhere (a 32-deep dependent max / add chain costs ~4 cycles per link); scale /
          // shift come as 16-byte shared loads
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 s4 = *reinterpret_cast<const float4*>(sc + c0 + i);
            const float4 b4 = *reinterpret_cast<const float4*>(sh_ + c0 + i);
            t[i + 0] = fmaf(__uint_as_float(acc[i + 0]), s4.x, b4.x);
            t[i + 1] = fmaf(__uint_as_float(acc[i + 1]), s4.y, b4.y);
            t[i + 2] = fmaf(__uint_as_float(acc[i + 2]), s4.z, b4.z);
            t[i + 3] = fmaf(__uint_as_float(acc[i + 3]), s4.w, b4.w);
            m4[0] = fmaxf(m4[0], t[i + 0]);
            m4[1] = fmaxf(m4[1], t[i + 1]);
            m4[2] = fmaxf(m4[2], t[i + 2]);
            m4[3] = fmaxf(m4[3], t[i + 3]);
          }
          float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          // warp-wide maximum first (5 shuffles), so that every lane exponentiates against the same reference and the
          // four sums can be reduced with plain additions
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
          // sum e and sum e * i with the bin index i an immediate; the depth offset d0 is added once: sum e * (d0 + i)
          float S4[4] = {0.f, 0.f, 0.f, 0.f}, Z4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = (p.dbg & 1) ? t[i] : ex2_ftz_f(t[i] - m);  // (dbg 1: ablation without the exponentials)
            S4[i & 3] += e;
            Z4[i & 3] = fmaf(e, (float)i, Z4[i & 3]);
          }
          float S = (S4[0] + S4[1]) + (S4[2] + S4[3]);
          float Sz = fmaf(S, d0, (Z4[0] + Z4[1]) + (Z4[2] + Z4[3]));
          float Sx = S * fw, Sy = S * fh;
          // four sums over 32 lanes in 6 shuffles (instead of 20): halve the number of values carried at each of the first
          // two butterfly levels -- the lane keeps the values whose index bit matches its own lane bit and sends the others
          {
            const bool hi16 = (lane & 16) != 0;
            // level 16: keep (S, Sx) in the lower half-warp, (Sy, Sz) in the upper one
            const float a0 = hi16 ? Sy : S, a1 = hi16 ? Sz : Sx;       // kept
            const float b0 = hi16 ? S : Sy, b1 = hi16 ? Sx : Sz;       // sent
            const float r0 = a0 + __shfl_xor_sync(0xffffffffu, b0, 16);
            const float r1 = a1 + __shfl_xor_sync(0xffffffffu, b1, 16);
            const bool hi8 = (lane & 8) != 0;
            // level 8: keep the first of the pair where bit 3 is clear, the second where it is set
            const float k = hi8 ? r1 : r0, snd = hi8 ? r0 : r1;
            float v = k + __shfl_xor_sync(0xffffffffu, snd, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            // lanes 0, 8, 16, 24 now hold S, Sx, Sy, Sz of the whole warp
            S = v;
          }
          if ((lane & 7) == 0) {
            const int chunk = ((th * 4 + q) << 1) | ((cabs >> 5) & 1);
            float* dst = p.head_partials + (((size_t)n0 * p.head_chunks + chunk) * p.head_nkpt + kp) * 5;
            const int which = lane >> 3;                 // 0: S, 1: Sx, 2: Sy, 3: Sz
            dst[1 + which] = S;
            if (lane == 0) dst[0] = m;
          }
          continue;
        }
        if (!ROLLED) {
          for (int ks = 1; ks < ksplit; ++ks) {  // partial sums of the K-split accumulators
            uint32_t part[32];
            tmem_ld32(taddr + (uint32_t)(ks * n_tile + c0), part);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < (ROLLED ? 8 : 32); ++i)
              acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(part[i]));
          }
        }
        uint8_t* const chunk_base = stage_row + (size_t)(c0 >> cko_shift) * blk_bytes +
                                    (((uint32_t)((c0 & (cko - 1)) >> 3) ^ (uint32_t)(sw & 4)) << 4);
        // Eight columns at a time.  The plain / residual / gathering flavours unroll the four groups of a 32-column
        // accumulator load (the gathers want all their global loads in flight at once); the staged flavours (up to six
        // shared-memory addends per group) run them as a ROLLED loop over 8-column TMEM loads, the next one in flight while
        // the current group is processed: unrolled, their epilogue alone was ~50 KB of SASS streaming through the
        // instruction caches (ncu: 40 % of the warp stalls were instruction fetches; 32 -> 64 fuse conv 126 -> 80 us).
        uint32_t nxt[8];
#pragma unroll(ROLLED ? 1 : 4)
        for (int g = 0; g < 4; ++g) {
          const int cg = c0 + g * 8;
          const int ab = ROLLED ? 0 : g * 8;   // this group's accumulator registers
          if (ROLLED) {
            tmem_ld_wait();   // group g has landed (in acc for g == 0, in nxt otherwise): only now may it be read
            if (g > 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[i] = nxt[i];
            }
            if (g < 3) tmem_ld8(taddr + (uint32_t)(cg + 8), nxt);
            for (int ks = 1; ks < ksplit; ++ks) {  // partial sums of the K-split accumulators (probing aid)
              uint32_t part[8];
              tmem_ld8(taddr + (uint32_t)(ks * n_tile + cg), part);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(part[i]));
            }
          }
          float v[8];
          const float4 s0 = *reinterpret_cast<const float4*>(sc + cg);
          const float4 s1 = *reinterpret_cast<const float4*>(sc + cg + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(sh_ + cg);
          const float4 b1 = *reinterpret_cast<const float4*>(sh_ + cg + 4);
          v[0] = fmaf(__uint_as_float(acc[ab + 0]), s0.x, b0.x);
          v[1] = fmaf(__uint_as_float(acc[ab + 1]), s0.y, b0.y);
          v[2] = fmaf(__uint_as_float(acc[ab + 2]), s0.z, b0.z);
          v[3] = fmaf(__uint_as_float(acc[ab + 3]), s0.w, b0.w);
          v[4] = fmaf(__uint_as_float(acc[ab + 4]), s1.x, b1.x);
          v[5] = fmaf(__uint_as_float(acc[ab + 5]), s1.y, b1.y);
          v[6] = fmaf(__uint_as_float(acc[ab + 6]), s1.z, b1.z);
          v[7] = fmaf(__uint_as_float(acc[ab + 7]), s1.w, b1.w);
          uint4* const sptr = reinterpret_cast<uint4*>(chunk_base + (ROLLED ? (((uint32_t)g ^ (uint32_t)(sw & 3)) << 4) : goff[g]));
          if (EPI != EPI_PLAIN) {
            if (has_res) add_bf16x8(v, *sptr);  // residual prefetched by TMA (zero-filled outside the tensor)
            if (staged) {
              // every other addend sits in its own slot of the ring entry, same swizzled box layout: same offsets
              const uint8_t* const sp8 = reinterpret_cast<const uint8_t*>(sptr);
              if (cfg.add_off[0] != 0) add_bf16x8(v, *reinterpret_cast<const uint4*>(sp8 + cfg.add_off[0]));
              if (cfg.add_off[1] != 0) add_bf16x8(v, *reinterpret_cast<const uint4*>(sp8 + cfg.add_off[1]));
              if (FULLISH) {
                const uint32_t blk_i = (uint32_t)(c0 >> cko_shift);
                const uint32_t chunk = (uint32_t)((c0 & (cko - 1)) >> 3) + (uint32_t)g;
#pragma unroll
                for (int a = 0; a < 3; ++a)
                  if (cfg.up_off[a] != 0)
                    add_bf16x8(v, *reinterpret_cast<const uint4*>(up_row[a] + (size_t)blk_i * cfg.up_blk[a] +
                                                                  ((chunk ^ up_sw[a]) << 4)));
              }
            } else if (GENERIC && valid) {
              if (use_pre0) add_bf16x8(v, __ldg(reinterpret_cast<const uint4*>(pre0 + cg)));
              if (p.pre[1] != nullptr) add_bf16x8(v, __ldg(reinterpret_cast<const uint4*>(pre1 + cg)));
              if (p.pre[2] != nullptr) add_bf16x8(v, __ldg(reinterpret_cast<const uint4*>(pre2 + cg)));
              if (EPI == EPI_FULL) {
#pragma unroll
                for (int a = 0; a < 3; ++a)
                  if (p.up[a] != nullptr) add_bf16x8(v, __ldg(reinterpret_cast<const uint4*>(upp[a] + cg)));
              }
            }
          }
          if (relu_explicit) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (FULLISH && p.post != nullptr) {
            if (staged) add_bf16x8(v, *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(sptr) + cfg.add_off[2]));
            else if (valid) add_bf16x8(v, __ldg(reinterpret_cast<const uint4*>(postp + cg)));
          }
          if (do_store) {
            uint4 o;
            if (relu_in_cvt) {
              o.x = pack_bf16x2_relu(v[0], v[1]);
              o.y = pack_bf16x2_relu(v[2], v[3]);
              o.z = pack_bf16x2_relu(v[4], v[5]);
              o.w = pack_bf16x2_relu(v[6], v[7]);
            } else {
              o.x = pack_bf16x2(v[0], v[1]);
              o.y = pack_bf16x2(v[2], v[3]);
              o.z = pack_bf16x2(v[4], v[5]);
              o.w = pack_bf16x2(v[6], v[7]);
            }
            *sptr = o;
          }
          if (pool) {
            float pv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pv[i] = valid ? v[i] : 0.f;
            const float csum = warp_colsum8(pv, lane);
            const int n_warp = n0 + ((q * 32) >> (p.bw_shift + p.bh_shift));
            if ((lane & 3) == 0 && n_warp < p.B)
              atomicAdd(p.pool_out + (size_t)n_warp * p.Cout + c_base + cg + (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 +
                                                                              ((lane >> 2) & 1)),
                        csum * p.pool_scale);
          }
        }
      }
      // accumulator drained: hand it back to the MMA warp (one mbarrier arrival per warp: 256 per-thread arrivals
      // on one shared-memory word serialise and cost more than the epilogue arithmetic)
      tc_fence_before();
      if (do_store) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staging writes -> async proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars->tmem_empty[abuf]);
        if (do_store) mbar_arrive(&bars->stag_ready[sbuf]);  // hand the staging buffer to the store warp
      }
      if (warp == 2 && lane == 0) tl_stamp(p.timeline, li, 9);
    }
  }
#undef HRP_DECODE_TILE

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)cfg.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------
// host: geometry, weight packing, tensor maps, launch
// ------------------------------------------------------------------------------------------------------
static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static int posmod(int a, int b) { return ((a % b) + b) % b; }

int conv_geometry(const ConvLayerDesc& d, ConvParams* pp) {
  ConvParams& p = *pp;
  memset(&p, 0, sizeof(p));
  HRP_REQUIRE(d.B > 0 && d.Hin > 0 && d.Win > 0 && d.Cout > 0, "conv dims must be positive");
  HRP_REQUIRE(d.Cin == 16 || d.Cin == 32 || d.Cin % 64 == 0, "stored Cin must be 16, 32 or a multiple of 64");
  HRP_REQUIRE(d.Cout % 32 == 0, "Cout must be a multiple of 32");
  p.B = d.B;
  p.Cin = d.Cin;
  p.Cout = d.Cout;
  p.Hin = d.Hin;
  p.Win = d.Win;
  p.src_sh = p.src_sw = 1;
  p.nphase = 1;
  p.os = 1;
  p.relu = d.relu;
  p.ck = std::min(d.Cin, 64);
  p.cpt = d.Cin / p.ck;
  if (d.kind == kConv) {
    HRP_REQUIRE(d.stride == 1 || d.stride == 2, "conv stride must be 1 or 2");
    HRP_REQUIRE(d.kh * d.kw <= kMaxTaps, "too many taps");
    p.Hout = (d.Hin + 2 * d.pad - d.kh) / d.stride + 1;
    p.Wout = (d.Win + 2 * d.pad - d.kw) / d.stride + 1;
    p.Hm = p.Hout;
    p.Wm = p.Wout;
    p.ntaps = d.kh * d.kw;
    if (d.stride == 1) {
      p.Hs = d.Hin;
      p.Ws = d.Win;
      for (int i = 0; i < d.kh; ++i)
        for (int j = 0; j < d.kw; ++j) {
          p.tap_dh[i * d.kw + j] = (int8_t)(i - d.pad);
          p.tap_dw[i * d.kw + j] = (int8_t)(j - d.pad);
          p.tap_map[i * d.kw + j] = 0;
        }
    } else {
      HRP_REQUIRE(d.Hin % 2 == 0 && d.Win % 2 == 0, "stride-2 conv needs even input size");
      p.Hs = d.Hin / 2;
      p.Ws = d.Win / 2;
      p.src_sh = p.src_sw = 2;
      HRP_REQUIRE(p.Hout == p.Hs && p.Wout == p.Ws, "unsupported stride-2 geometry");
      for (int i = 0; i < d.kh; ++i)
        for (int j = 0; j < d.kw; ++j) {
          const int rh = i - d.pad, rw = j - d.pad;
          const int hp = posmod(rh, 2), wp = posmod(rw, 2);
          p.tap_dh[i * d.kw + j] = (int8_t)floordiv(rh, 2);
          p.tap_dw[i * d.kw + j] = (int8_t)floordiv(rw, 2);
          p.tap_map[i * d.kw + j] = (int8_t)(hp * 2 + wp);
        }
    }
  } else if (d.kind == kDeconvK4S2P1) {
    p.Hout = d.Hin * 2;
    p.Wout = d.Win * 2;
    p.Hm = d.Hin;
    p.Wm = d.Win;
    p.Hs = d.Hin;
    p.Ws = d.Win;
    p.os = 2;
    p.nphase = 4;
    p.ntaps = 4;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        p.tap_dh[a * 2 + b] = (int8_t)(-a);  // + ph in the kernel
        p.tap_dw[a * 2 + b] = (int8_t)(-b);  // + pw in the kernel
        p.tap_map[a * 2 + b] = 0;
      }
  } else if (d.kind == kConvUp2) {
    HRP_REQUIRE(d.kh == 1 && d.kw == 1 && d.stride == 1 && d.pad == 0, "the upsampling conv is a 1x1 conv");
    p.Hout = d.Hin * 2;
    p.Wout = d.Win * 2;
    p.Hm = d.Hin;
    p.Wm = d.Win;
    p.Hs = d.Hin;
    p.Ws = d.Win;
    p.os = 2;
    p.nphase = 4;
    p.shared_phase = 1;
    p.ntaps = 1;
    p.tap_dh[0] = p.tap_dw[0] = 0;
    p.tap_map[0] = 0;
  } else if (d.kind == kStemS2D) {
    HRP_REQUIRE(d.Cin == 16 && d.stride == 2, "s2d stem expects the 16-channel space-to-depth input");
    // d.Hin/d.Win are the s2d dims (H/2, W/2); output of the original stride-2 conv has the same size
    const int Horig = d.Hin * 2, Worig = d.Win * 2;
    p.Hout = (Horig + 2 * d.pad - d.kh) / 2 + 1;
    p.Wout = (Worig + 2 * d.pad - d.kw) / 2 + 1;
    HRP_REQUIRE(p.Hout == d.Hin && p.Wout == d.Win, "unsupported stem geometry");
    p.Hm = p.Hout;
    p.Wm = p.Wout;
    p.Hs = d.Hin;
    p.Ws = d.Win;
    const int dh_lo = floordiv(-d.pad, 2), dh_hi = floordiv(d.kh - 1 - d.pad, 2);
    const int dw_lo = floordiv(-d.pad, 2), dw_hi = floordiv(d.kw - 1 - d.pad, 2);
    const int nh = dh_hi - dh_lo + 1, nw = dw_hi - dw_lo + 1;
    HRP_REQUIRE(nh * nw <= kMaxTaps, "too many s2d taps");
    const bool wide = d.in_wpitch > 0 && (nw * 16 == 32 || nw * 16 == 64) && d.in_wpad + dw_lo >= 0 &&
                      d.in_wpad + d.Win + dw_hi <= d.in_wpitch;
    if (wide) {
      // one K block per kernel ROW: the nw horizontally adjacent s2d pixels are contiguous in the padded tensor, the
      // zero padding columns stand in for the out-of-bounds taps (vertical taps still use the TMA zero fill).  The
      // packed weight matrix is unchanged: its K order is already (row tap, column tap, 16 channels).
      p.ntaps = nh;
      p.Cin = nw * 16;
      p.ck = p.Cin;
      p.cpt = 1;
      for (int a = 0; a < nh; ++a) {
        p.tap_dh[a] = (int8_t)(dh_lo + a);
        p.tap_dw[a] = 0;
        p.tap_map[a] = 0;
      }
      p.src_pix = 16;
      p.src_row = d.in_wpitch * 16;
      p.src_img = (long long)d.Hin * d.in_wpitch * 16;
      p.src_off = (d.in_wpad + dw_lo) * 16;
    } else {
      HRP_REQUIRE(d.in_wpitch == 0, "padded s2d input does not fit this stem geometry");
      p.ntaps = nh * nw;
      for (int a = 0; a < nh; ++a)
        for (int b = 0; b < nw; ++b) {
          p.tap_dh[a * nw + b] = (int8_t)(dh_lo + a);
          p.tap_dw[a * nw + b] = (int8_t)(dw_lo + b);
          p.tap_map[a * nw + b] = 0;
        }
    }
  } else {
    set_error("unknown conv kind");
    return HRP_ERR_INVALID;
  }
  p.ktot = p.ntaps * p.Cin;
  if (p.src_pix == 0) {  // dense NHWC source (everything but the padded stems)
    p.src_pix = p.src_sw * p.Cin;
    p.src_row = p.src_sh * p.Win * p.Cin;
    p.src_img = (long long)p.Hin * p.Win * p.Cin;
    p.src_off = 0;
  }
  // M tile box
  p.bw = std::min(p.Wm, kTileM);
  // round bw down to a power of two so that bw*bh*bn == 128 exactly
  int bw = 1;
  while (bw * 2 <= p.bw) bw *= 2;
  p.bw = bw;
  int bh = 1;
  while (bh * 2 <= std::min(p.Hm, kTileM / p.bw)) bh *= 2;
  p.bh = bh;
  p.bn = kTileM / (p.bw * p.bh);
  p.tiles_w = (p.Wm + p.bw - 1) / p.bw;
  p.tiles_h = (p.Hm + p.bh - 1) / p.bh;
  p.tiles_n = (p.B + p.bn - 1) / p.bn;
  auto ilog2 = [](int v) { int sft = 0; while ((1 << sft) < v) ++sft; return sft; };
  p.tw_shift = ilog2(p.tiles_w);
  p.th_shift = ilog2(p.tiles_h);
  p.bw_shift = ilog2(p.bw);
  p.bh_shift = ilog2(p.bh);
  HRP_REQUIRE((1 << p.tw_shift) == p.tiles_w && (1 << p.th_shift) == p.tiles_h,
              "spatial sizes must give power-of-two tile counts");
  // N tile
  // N tile: 256 wide for K-heavy (tensor-bound) layers; 128 for short-K layers, whose time is the epilogue and
  // the output stream: 4 CTAs/SM fit in TMEM instead of 2 and the (small) A tile is re-read from L2
  int max_n = ((p.ktot <= 256 || d.has_residual) && p.Cout > 128 && p.Cout % 128 == 0) ? 128 : 256;
  // Small-M problems (batch-1 latency: an 8x8 or 16x16 layer is 1-2 M tiles): with 256-wide N tiles one or two CTAs
  // stream the whole weight matrix through one SM's TMA (1.2 MB for 256 -> 256 k3: ~20 us).  Narrower N tiles spread the
  // weight read and the MMAs over more SMs; the A tile they all re-read is tiny.  Only for Cout >= 128 (the 32 / 64
  // channel layers keep n_tile == Cout, which the halo kernel needs) and only while the grid stays below ~1/3 of the
  // GPU, so throughput batches are untouched.  The packed weight layout does not depend on n_tile (cout_pad == Cout).
  {
    static const bool small_m = !(getenv("HRP_CONV_SMALLM") != nullptr && getenv("HRP_CONV_SMALLM")[0] == '0');
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.nphase;
    // (measured, Panda full model, p50: batch 1 1.99 -> 1.80 ms, batch 4 2.12 -> 1.96 ms; at 16 images a 48-CTA target
    //  was 2.5 % slower than no splitting, hence the cap on the M-tile count and the 24-CTA target)
    while (small_m && !d.head_fold && p.Cout >= 128 && max_n > 32 && p.Cout % (max_n / 2) == 0 && m_tiles <= 4 &&
           m_tiles * ((p.Cout + max_n - 1) / max_n) < 24)
      max_n /= 2;
    // (Experiment, profiles/r02_exp_small_shards.txt: also halving the N tile of long-K layers whenever the grid stays
    //  within one wave -- fewer operand bytes per SM for the 8x8 layers of a 64-image shard -- was neutral to slower at
    //  every batch from 1 to 256, so the rule above stays limited to 1-4 M tiles.)
  }
  int n_tiles = (p.Cout + max_n - 1) / max_n;
  p.n_tile = (((p.Cout + n_tiles - 1) / n_tiles) + 31) / 32 * 32;
  if (d.head_fold) {
    // soft-argmax fold: every N tile holds whole keypoints (64 depth bins each): 128-wide tiles, Cout padded with
    // zero weight rows (Panda 448 -> 512, Baxter 1088 -> 1152); the epilogue skips the padding columns
    HRP_REQUIRE(d.kind == kConv && d.kh == 1 && d.kw == 1 && d.stride == 1 && p.Cout % 64 == 0,
                "the soft-argmax fold needs a 1x1 conv whose channels are keypoints x 64 depth bins");
    p.n_tile = 128;
    n_tiles = (p.Cout + 127) / 128;
  }
  p.cout_pad = n_tiles * p.n_tile;
  p.cko = (p.n_tile % 64 == 0) ? 64 : 32;  // channel block of the TMA-store epilogue
  p.pool_scale = 1.f / (float)(p.Hout * p.Wout);
  // vertical tap sharing: 3x3 stride-1 pad-1 convs whose tile lies inside one image
  const char* vs = getenv("HRP_CONV_VSH");
  // (default: only where the weights are small -- Cout <= 64 -- so a 3-tap stage still allows several CTAs per SM;
  //  HRP_CONV_VSH=0 disables it, =2 enables it for every eligible layer)
  const bool vsh_ok = (d.kind == kConv && d.stride == 1 && d.kh == 3 && d.kw == 3 && d.pad == 1 && p.bn == 1 && p.bw % 8 == 0);
  const int vsh_mode = (vs != nullptr) ? (vs[0] - '0') : 0;  // measured slower on B200 (fewer CTAs per SM): opt-in
  p.vsh = (vsh_ok && (vsh_mode == 2 || (vsh_mode == 1 && p.Cout <= 64))) ? 1 : 0;
  if (p.vsh) {
    p.vsh_a_bytes = (p.bh + 2) * p.bw * p.ck * 2;
    p.vsh_a_pad = (p.vsh_a_bytes + 1023) / 1024 * 1024;
    p.vsh_stage_bytes = p.vsh_a_pad + 3 * p.n_tile * p.ck * 2;
  }
  // pixel-pair weights for the halo kernel (conv_halo.cu): 32 -> 32 channel 3x3 stride-1 layers with an even width
  p.pair_off = 0;
  if (d.kind == kConv && d.stride == 1 && d.kh == 3 && d.kw == 3 && d.pad == 1 && d.Cin == 32 && d.Cout == 32 &&
      d.Win % 2 == 0 && p.n_tile == 32 && p.cout_pad == 32)
    p.pair_off = p.nphase * p.cout_pad * p.ktot;
  return HRP_OK;
}

size_t conv_packed_weight_elems(const ConvParams& p) {
  return (size_t)(p.shared_phase ? 1 : p.nphase) * p.cout_pad * p.ktot + (p.pair_off > 0 ? (size_t)64 * 6 * 64 : 0);
}

int conv_pack_weights(const ConvLayerDesc& d, const ConvParams& p, int cin_ref, const float* w, uint16_t* out) {
  const size_t total = conv_packed_weight_elems(p);
  memset(out, 0, total * sizeof(uint16_t));
  if (d.kind == kConv || d.kind == kConvUp2) {
    HRP_REQUIRE(cin_ref <= p.Cin, "reference Cin exceeds stored Cin");
    for (int co = 0; co < p.Cout; ++co)
      for (int ci = 0; ci < cin_ref; ++ci)
        for (int i = 0; i < d.kh; ++i)
          for (int j = 0; j < d.kw; ++j) {
            const float v = w[(((size_t)co * cin_ref + ci) * d.kh + i) * d.kw + j];
            out[(size_t)co * p.ktot + (size_t)(i * d.kw + j) * p.Cin + ci] = f32_to_bf16_bits(v);
          }
    if (p.pair_off > 0) {
      // Pixel-pair view (conv_halo.cu): two horizontally adjacent pixels form one 64-channel position, so the layer
      // becomes a 64 -> 64 channel 3x3 conv over (H, W/2) whose per-tap matrices have structured zero blocks:
      //   Wpair[dh][dj][po*32+co][pi*32+ci] = W[co][ci][dh][dw],  dw = 2*dj + pi - po  (zero when |dw| > 1)
      // dj = 0 is a full 64 x 64 tile; dj = -1 only has the block (po = 0, pi = 1) and dj = +1 only (po = 1, pi = 0),
      // which share ONE tile.  Packed as 64 rows x 6 tiles x 64 K columns: tiles 0..2 = (dh, dj = 0), tiles 3..5 = the
      // merged (dh, dj = -1 | +1) tiles.
      uint16_t* pw = out + p.pair_off;
      const size_t pitch = 6 * 64;
      for (int i = 0; i < 3; ++i)
        for (int dj = -1; dj <= 1; ++dj)
          for (int po = 0; po < 2; ++po)
            for (int pi = 0; pi < 2; ++pi) {
              const int dw = 2 * dj + pi - po;
              if (dw < -1 || dw > 1) continue;
              const int tile = (dj == 0) ? i : 3 + i;
              for (int co = 0; co < 32; ++co)
                for (int ci = 0; ci < cin_ref; ++ci)
                  pw[(size_t)(po * 32 + co) * pitch + (size_t)tile * 64 + pi * 32 + ci] =
                      f32_to_bf16_bits(w[(((size_t)co * cin_ref + ci) * 3 + i) * 3 + (dw + 1)]);
            }
    }
  } else if (d.kind == kDeconvK4S2P1) {
    HRP_REQUIRE(cin_ref <= p.Cin, "reference Cin exceeds stored Cin");
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw)
        for (int co = 0; co < p.Cout; ++co)
          for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
              const int kh = 1 - ph + 2 * a, kw = 1 - pw + 2 * b;
              uint16_t* dst = out + ((size_t)(ph * 2 + pw) * p.cout_pad + co) * p.ktot + (size_t)(a * 2 + b) * p.Cin;
              for (int ci = 0; ci < cin_ref; ++ci)
                dst[ci] = f32_to_bf16_bits(w[(((size_t)ci * p.Cout + co) * 4 + kh) * 4 + kw]);
            }
  } else if (d.kind == kStemS2D) {
    HRP_REQUIRE(cin_ref == 3, "stem conv expects 3 input channels");
    // K order = (row tap a, column tap b, 16 s2d channels), the same for the per-tap and the per-row (padded input)
    // geometries of conv_geometry
    const int dh_lo = floordiv(-d.pad, 2), dh_hi = floordiv(d.kh - 1 - d.pad, 2);
    const int dw_lo = floordiv(-d.pad, 2), dw_hi = floordiv(d.kw - 1 - d.pad, 2);
    const int nh = dh_hi - dh_lo + 1, nw = dw_hi - dw_lo + 1;
    HRP_REQUIRE(nh * nw * 16 == p.ktot, "stem geometry / packing mismatch");
    for (int co = 0; co < p.Cout; ++co)
      for (int a = 0; a < nh; ++a)
        for (int b = 0; b < nw; ++b)
          for (int hp = 0; hp < 2; ++hp)
            for (int wp = 0; wp < 2; ++wp) {
              const int i = 2 * (dh_lo + a) + hp + d.pad, j = 2 * (dw_lo + b) + wp + d.pad;
              if (i < 0 || i >= d.kh || j < 0 || j >= d.kw) continue;
              for (int ci = 0; ci < 3; ++ci) {
                const float v = w[(((size_t)co * 3 + ci) * d.kh + i) * d.kw + j];
                out[(size_t)co * p.ktot + (size_t)(a * nw + b) * 16 + (hp * 2 + wp) * 3 + ci] = f32_to_bf16_bits(v);
              }
            }
  }
  return HRP_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

int conv_encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int ck) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return HRP_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = (ck == 64)   ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (ck == 32) ? CU_TENSOR_MAP_SWIZZLE_64B
                                             : CU_TENSOR_MAP_SWIZZLE_32B;
  // (HRP_TMA_L2PROMO=0|64|128|256 overrides the L2 promotion size: probing aid, tools/probe_tma.py)
  static const int promo_env = (getenv("HRP_TMA_L2PROMO") != nullptr) ? atoi(getenv("HRP_TMA_L2PROMO")) : 256;
  const CUtensorMapL2promotion promo = (promo_env == 0)     ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                       : (promo_env == 64)  ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : (promo_env == 128) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                            : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                  reinterpret_cast<const cuuint32_t*>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return HRP_ERR_CUDA;
  }
  return HRP_OK;
}

static void set_smem_attr_once() {
  static std::once_flag once;
  std::call_once(once, [] {
  #define HRP_SET_ATTR(CKV, EPIV) \
  cudaFuncSetAttribute(conv_gemm_kernel<CKV, EPIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
    HRP_SET_ATTR(16, EPI_PLAIN); HRP_SET_ATTR(16, EPI_PRE); HRP_SET_ATTR(16, EPI_FULL); HRP_SET_ATTR(16, EPI_RES);
    HRP_SET_ATTR(32, EPI_PLAIN); HRP_SET_ATTR(32, EPI_PRE); HRP_SET_ATTR(32, EPI_FULL); HRP_SET_ATTR(32, EPI_RES);
    HRP_SET_ATTR(64, EPI_PLAIN); HRP_SET_ATTR(64, EPI_PRE); HRP_SET_ATTR(64, EPI_FULL); HRP_SET_ATTR(64, EPI_RES);
#undef HRP_SET_ATTR
#define HRP_SET_ATTR_P(CKV, EPIV) \
  cudaFuncSetAttribute(conv_gemm_persistent<CKV, EPIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
    HRP_SET_ATTR_P(16, EPI_PLAIN); HRP_SET_ATTR_P(16, EPI_PRE); HRP_SET_ATTR_P(16, EPI_FULL); HRP_SET_ATTR_P(16, EPI_RES);
    HRP_SET_ATTR_P(32, EPI_PLAIN); HRP_SET_ATTR_P(32, EPI_PRE); HRP_SET_ATTR_P(32, EPI_FULL); HRP_SET_ATTR_P(32, EPI_RES);
    HRP_SET_ATTR_P(64, EPI_PLAIN); HRP_SET_ATTR_P(64, EPI_PRE); HRP_SET_ATTR_P(64, EPI_FULL); HRP_SET_ATTR_P(64, EPI_RES);
    HRP_SET_ATTR_P(64, EPI_HEAD);
    HRP_SET_ATTR_P(16, EPI_PRE_ST); HRP_SET_ATTR_P(16, EPI_FULL_ST);
    HRP_SET_ATTR_P(32, EPI_PRE_ST); HRP_SET_ATTR_P(32, EPI_FULL_ST);
    HRP_SET_ATTR_P(64, EPI_PRE_ST); HRP_SET_ATTR_P(64, EPI_FULL_ST);
#undef HRP_SET_ATTR_P
  });
}

void conv_init() {
  set_smem_attr_once();
  conv_halo_init();
}

int conv_plan_finalize(ConvPlan* plan, const bf16* in, const bf16* w_packed, const ConvLayerDesc* desc) {
  ConvParams& p = plan->p;
  HRP_REQUIRE(in != nullptr && w_packed != nullptr, "null tensor");
  HRP_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
              "tensors must be 16-byte aligned");
  p.in = in;
  p.w = w_packed;
  if (const char* dbg = getenv("HRP_CONV_DBG")) p.dbg = atoi(dbg);
  // A maps: one per input parity for stride-2 sources, else a single map
  const int nmaps = (p.src_sh == 2) ? 4 : 1;
  for (int m = 0; m < nmaps; ++m) {
    const int hp = m >> 1, wp = m & 1;
    const bf16* base = in + p.src_off + ((size_t)hp * p.Win + wp) * p.Cin;
    uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)p.B};
    uint64_t strides[3] = {(uint64_t)p.src_pix * 2, (uint64_t)p.src_row * 2, (uint64_t)p.src_img * 2};
    uint32_t box[4] = {(uint32_t)p.ck, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int rc = conv_encode_map(&plan->maps.a[m], base, 4, dims, strides, box, p.ck);
    if (rc != HRP_OK) return rc;
  }
  for (int m = nmaps; m < 4; ++m) plan->maps.a[m] = plan->maps.a[0];
  plan->maps.av = plan->maps.a[0];
  if (p.vsh) {
    uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)p.B};
    uint64_t strides[3] = {(uint64_t)p.Cin * 2, (uint64_t)p.Win * p.Cin * 2, (uint64_t)p.Hin * p.Win * p.Cin * 2};
    uint32_t box[4] = {(uint32_t)p.ck, (uint32_t)p.bw, (uint32_t)(p.bh + 2), 1u};
    int rc = conv_encode_map(&plan->maps.av, in, 4, dims, strides, box, p.ck);
    if (rc != HRP_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.ktot, (uint64_t)(p.shared_phase ? 1 : p.nphase) * p.cout_pad};
    uint64_t strides[1] = {(uint64_t)p.ktot * 2};
    uint32_t box[2] = {(uint32_t)p.ck, (uint32_t)p.n_tile};
    int rc = conv_encode_map(&plan->maps.b, w_packed, 2, dims, strides, box, p.ck);
    if (rc != HRP_OK) return rc;
  }
  if (p.pool_out != nullptr)
    HRP_REQUIRE(p.bw * p.bh >= 32, "pooled epilogue needs >= 32 rows per image in a tile");
  if (p.out != nullptr) {
    // output maps for the TMA-store epilogue: one per deconv phase (stride-2 interleaved pixels), else one
    HRP_REQUIRE((reinterpret_cast<uintptr_t>(p.out) & 15) == 0, "output must be 16-byte aligned");
    for (int ph = 0; ph < p.nphase; ++ph) {
      const int oh = p.oh0 + (ph >> 1), ow = p.ow0 + (ph & 1);
      const bf16* base = p.out + ((size_t)oh * p.Wout + ow) * p.Cout;
      uint64_t dims[4] = {(uint64_t)p.Cout, (uint64_t)p.Wm, (uint64_t)p.Hm, (uint64_t)p.B};
      uint64_t strides[3] = {(uint64_t)p.os * p.Cout * 2, (uint64_t)p.os * p.Wout * p.Cout * 2,
                             (uint64_t)p.Hout * p.Wout * p.Cout * 2};
      uint32_t box[4] = {(uint32_t)p.cko, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
      int rc = conv_encode_map(&plan->maps.o[ph], base, 4, dims, strides, box, p.cko);
      if (rc != HRP_OK) return rc;
    }
    for (int ph = p.nphase; ph < 4; ++ph) plan->maps.o[ph] = plan->maps.o[0];
  }
  const int stage_bytes = p.vsh ? p.vsh_stage_bytes : kStageABytes + p.n_tile * 128;
  const int sub = 64 / p.ck;
  const int n_iters = p.vsh ? 3 * p.cpt : (p.ntaps * p.cpt + sub - 1) / sub;
  // shallow per-CTA pipelines, several CTAs per SM: small tiles are latency-bound, not smem-bound
  const int budget = (p.n_tile <= 64) ? 56 * 1024 : (p.n_tile <= 128) ? 72 * 1024 : 100 * 1024;
  int stages = std::max(2, budget / stage_bytes);
  if (p.vsh && 3 * stage_bytes <= 72 * 1024) stages = 3;  // the whole 3-iteration K loop in flight
  stages = std::max(1, std::min(stages, n_iters));
  plan->stages = stages;
  const int staging = (p.out != nullptr) ? kTileM * p.n_tile * 2 : 0;
  const bool res_v1 = (p.pre[0] != nullptr) && (p.out != nullptr) && p.os == 1;
  plan->bar_offset = res_v1 ? stages * stage_bytes + staging : std::max(stages * stage_bytes, staging);
  plan->smem_bytes = plan->bar_offset + (int)sizeof(PipeBarriers) + 2 * p.n_tile * (int)sizeof(float) + 1024;
  p.n_tiles = p.cout_pad / p.n_tile;
  plan->grid = dim3((unsigned)(p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles), 1u, (unsigned)p.nphase);
  const bool full = p.up[0] || p.up[1] || p.up[2] || p.post || p.pool_out;
  const bool pre = p.pre[0] || p.pre[1] || p.pre[2];
  const bool res_only = !full && p.pre[0] != nullptr && p.pre[1] == nullptr && p.pre[2] == nullptr && p.out != nullptr &&
                        p.nphase == 1 && p.os == 1;
  plan->epi = full ? EPI_FULL : (res_only ? EPI_RES : (pre ? EPI_PRE : EPI_PLAIN));
  if (p.head_partials != nullptr) {
    HRP_REQUIRE(!full && !pre && p.out == nullptr && p.ck == 64 && p.n_tile == 128 && p.nphase == 1,
                "the soft-argmax fold takes no addends and writes no tensor");
    HRP_REQUIRE(p.bw == 64 && p.bh == 2 && p.bn == 1 && p.Hm == 64 && p.Wm == 64,
                "the soft-argmax fold is specialised for 64 x 64 heatmaps (tile = two image rows)");
    HRP_REQUIRE(p.head_chunks == p.tiles_h * 8 && p.head_nkpt * 64 == p.Cout, "soft-argmax partial layout mismatch");
    plan->epi = EPI_HEAD;
  }

  // ---- persistent variant (default) ----
  {
    static int num_sms = 0;
    if (num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
      if (num_sms <= 0) num_sms = 148;
    }
    // HRP_CONV_PERSISTENT=1 forces the persistent kernel everywhere, =0 the one-tile-per-CTA kernel; by default the
    // persistent kernel is used where it measured faster on B200: wide short-K layers (the epilogue dominates and
    // eight epilogue warps + double-buffered accumulators pay off)
    const char* pe = getenv("HRP_CONV_PERSISTENT");
    if (pe != nullptr && (pe[0] == '0' || pe[0] == '1')) plan->persistent = (pe[0] == '1');
    else plan->persistent = (p.n_tile == 128 && p.ktot >= 256 && p.ktot <= 1024 && p.pre[0] == nullptr);
    if (p.head_partials != nullptr) plan->persistent = true;  // the fold lives in the persistent kernel's epilogue
    PersistCfg& c = plan->pcfg;
    const int stag_bytes = (p.out != nullptr) ? kTileM * p.n_tile * 2 : 0;
    // ---- addends: pre[0] by TMA into the output slot (in place); with `staged` all the others as well ----
    const bool generic = (plan->epi == EPI_PRE || plan->epi == EPI_FULL);
    const char* e5 = getenv("HRP_CONV_STAGED");
    // default: layers whose N tile keeps all eight epilogue warps busy (n_tile >= 64); measured on the 32 -> 64 stride-2
    // fuse convs: 113 us staged vs 125-127 us with gathers.  HRP_CONV_STAGED=1 forces it wherever it fits, =0 disables it.
    const bool staged_default = p.n_tile >= 64 && p.os == 1;
    const bool staged_on = (e5 != nullptr) ? (e5[0] == '1') : staged_default;
    bool staged = generic && p.out != nullptr && p.pre[0] != nullptr && p.pool_out == nullptr && staged_on &&
                  (p.os == 1 || (p.os == 2 && p.pre[1] == nullptr && p.pre[2] == nullptr && p.post == nullptr &&
                                 p.oh0 == 0 && p.ow0 == 0));
    memset(c.add_off, 0, sizeof(c.add_off));
    memset(c.up_off, 0, sizeof(c.up_off));
    int entry = stag_bytes;
    if (staged) {
      const bf16* same[3] = {p.pre[1], p.pre[2], p.post};
      for (int i = 0; i < 3; ++i)
        if (same[i] != nullptr) {
          c.add_off[i] = entry;
          entry += stag_bytes;
        }
      const int nblk = p.n_tile / p.cko;
      for (int a = 0; a < 3; ++a)
        if (p.up[a] != nullptr) {
          const int sh = p.up_shift[a] - (p.os == 2 ? 1 : 0);
          if (sh < 0) { staged = false; break; }
          c.up_sh[a] = sh;
          c.up_bw[a] = std::max(p.bw >> sh, 1);
          c.up_bh[a] = std::max(p.bh >> sh, 1);
          c.up_blk[a] = (c.up_bw[a] * c.up_bh[a] * p.bn * p.cko * 2 + 1023) / 1024 * 1024;
          c.up_off[a] = entry;
          entry += nblk * c.up_blk[a];
        }
    }
    const bool has_res = (p.pre[0] != nullptr) && (p.out != nullptr) && (p.os == 1 || staged);
    // (without staging, output-strided layers -- deconv, upsampling conv -- read pre[0] with plain loads)
    auto out_like_map = [&](CUtensorMap* m, const bf16* t, int ph) -> int {
      const int oh = p.oh0 + (ph >> 1), ow = p.ow0 + (ph & 1);
      const bf16* base = t + ((size_t)oh * p.Wout + ow) * p.Cout;
      uint64_t dims[4] = {(uint64_t)p.Cout, (uint64_t)p.Wm, (uint64_t)p.Hm, (uint64_t)p.B};
      uint64_t strides[3] = {(uint64_t)p.os * p.Cout * 2, (uint64_t)p.os * p.Wout * p.Cout * 2,
                             (uint64_t)p.Hout * p.Wout * p.Cout * 2};
      uint32_t box[4] = {(uint32_t)p.cko, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
      return conv_encode_map(m, base, 4, dims, strides, box, p.cko);
    };
    for (int i = 0; i < 3; ++i) plan->maps.rp[i] = plan->maps.add[i] = plan->maps.upm[i] = plan->maps.a[0];
    if (has_res) {
      int rc = out_like_map(&plan->maps.r, p.pre[0], 0);
      if (rc != HRP_OK) return rc;
      for (int ph = 1; ph < p.nphase && p.os == 2; ++ph) {
        rc = out_like_map(&plan->maps.rp[ph - 1], p.pre[0], ph);
        if (rc != HRP_OK) return rc;
      }
    } else {
      plan->maps.r = plan->maps.a[0];
    }
    if (staged) {
      const bf16* same[3] = {p.pre[1], p.pre[2], p.post};
      for (int i = 0; i < 3; ++i)
        if (same[i] != nullptr) {
          int rc = out_like_map(&plan->maps.add[i], same[i], 0);
          if (rc != HRP_OK) return rc;
        }
      for (int a = 0; a < 3; ++a)
        if (p.up[a] != nullptr) {
          const int s = p.up_shift[a];
          const uint64_t Hl = (uint64_t)(p.Hout >> s), Wl = (uint64_t)(p.Wout >> s);
          uint64_t dims[4] = {(uint64_t)p.Cout, Wl, Hl, (uint64_t)p.B};
          uint64_t strides[3] = {(uint64_t)p.Cout * 2, Wl * p.Cout * 2, Hl * Wl * p.Cout * 2};
          uint32_t box[4] = {(uint32_t)p.cko, (uint32_t)c.up_bw[a], (uint32_t)c.up_bh[a], (uint32_t)p.bn};
          int rc = conv_encode_map(&plan->maps.upm[a], p.up[a], 4, dims, strides, box, p.cko);
          if (rc != HRP_OK) return rc;
        }
    }
    const int tail = 512 + 2 * p.cout_pad * (int)sizeof(float) + 1024;  // barriers + scale/shift + alignment slack
    const int avail = 227 * 1024 - tail;
    const int b_sub = p.n_tile * p.ck * 2;
    // vertical tap sharing + resident weights: in the persistent kernel the producer thread's issue rate is the
    // limit for small tiles, so fewer TMA operations per tile is what counts (HRP_CONV_PVSH=0 / HRP_CONV_WRES=0 disable)
    const char* e1 = getenv("HRP_CONV_PVSH");
    const char* e2 = getenv("HRP_CONV_WRES");
    const bool vsh_ok = (p.nphase == 1 && p.src_sh == 1 && p.ntaps == 9 && p.tap_dh[0] == -1 && p.tap_dw[0] == -1 &&
                         p.tap_dh[8] == 1 && p.tap_dw[8] == 1 && p.bn == 1 && p.Hs == p.Hm && p.Ws == p.Wm);
    c.vsh = (vsh_ok && !(e1 != nullptr && e1[0] == '0')) ? 1 : 0;
    const int wbytes = p.ntaps * p.cpt * b_sub;
    // resident weights: one N tile always; several N tiles when the grid can be made a multiple of n_tiles without
    // idling more than ~5 % of the SMs (each CTA then owns ONE N tile for all its M tiles)
    const int grid_nt = (p.n_tiles > 1) ? (num_sms / p.n_tiles) * p.n_tiles : num_sms;
    const bool multi_ok = p.n_tiles > 1 && p.nphase == 1 && grid_nt * 20 >= num_sms * 19 &&
                          (long long)p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles >= 2LL * num_sms;
    c.wres = ((p.n_tiles == 1 || multi_ok) && p.nphase == 1 && wbytes <= 80 * 1024 && !(e2 != nullptr && e2[0] == '0')) ? 1 : 0;
    // ---- pipeline geometry -------------------------------------------------------------------------------------
    // A stage holds kmul x 64 channels' worth of k-blocks (kmul x 16 KiB of A, plus the matching weight sub-tiles unless
    // the weights are resident).  One ring handshake (wait for the slot, expect_tx, TMA burst) costs the producer thread
    // ~550-650 cycles however many bytes it moves, and the MMA issuer pays a similar price per stage it waits for
    // (tools/probe_tma.py, profiles/r02_probe_tma.txt, profiles/r02_persist_timelines.txt): with 16 KiB stages the
    // 1x1 layers ran at ~900 cycles per stage against 257 cycles of MMAs.  So: the BIGGEST stage that still leaves two
    // of them in flight (kmul = 4, 2, 1; HRP_CONV_KSTAGE pins it), two producer threads taking alternate stages.
    const char* e6 = getenv("HRP_CONV_KSTAGE");
    const char* e7 = getenv("HRP_CONV_NPROD");
    const char* e4 = getenv("HRP_CONV_NSTAG");
    const int units = (p.ntaps * p.cpt) / (64 / p.ck);   // whole 64-channel units of K per tile
    c.pipe_offset = c.wres ? (wbytes + 1023) / 1024 * 1024 : 0;
    const int pipe_avail = avail - c.pipe_offset;
    auto stage_bytes_of = [&](int kmul) {
      const int a_bytes = c.vsh ? (p.bh + 2) * p.bw * p.ck * 2 : kStageABytes * kmul;
      const int a_region = (a_bytes + 1023) / 1024 * 1024;
      return a_region + (c.wres ? 0 : (c.vsh ? 3 * b_sub : p.n_tile * 128 * kmul));
    };
    // Staging ring: the residual / addend tiles of tile i + nstag - 1 are fetched (by TMA, into the ring entry the output
    // tile will be built in) while tile i is in the epilogue, so with a residual the ring depth is the prefetch distance
    // and must cover a DRAM round trip (~2 tile periods on short-K layers): 4 entries when they fit, else 2, else 1.
    auto pick_nstag = [&](int ent, int sbytes, int min_st, int* st_out) {
      int nstag = (stag_bytes > 0 && p.n_tile <= 128) ? ((has_res && !(e4 != nullptr && e4[0] == '2')) ? 4 : 2) : 1;
      int st = (pipe_avail - nstag * ent) / sbytes;
      while (st < min_st && nstag > 1) {
        nstag >>= 1;
        st = (pipe_avail - nstag * ent) / sbytes;
      }
      *st_out = st;
      return nstag;
    };
    int kmul = 1, st = 0, nstag = 1;
    {
      // (default 16 KiB stages: measured over the whole network -- profiles/r02_exp_producer.txt -- bigger stages win on
      //  the K = 256 / 512 1x1 layers (layer4 conv3 132 -> 114 us) and lose where they leave two slots or swallow the whole
      //  K (layer2 conv3 192 -> 281 us); the sum is a wash.  HRP_CONV_KSTAGE=2|4 allows them.)
      //  The soft-argmax fold is the exception: its sixteen epilogue warps compete with the producer / issuer warps for
      //  issue slots, and halving the handshakes per tile is worth 980 -> 810 us there.
      const int kpin = (e6 != nullptr) ? atoi(e6) : (plan->epi == EPI_HEAD ? 2 : 1);
      // candidates, biggest first; a stage never spans more than the tile's K
      for (int km : {4, 2, 1}) {
        if (c.vsh && km > 1) continue;
        if (km > 1 && km > units) continue;
        if (kpin > 0 && km != kpin && !(km == 1)) continue;
        // big stages keep the staging ring they would have had with 16 KiB stages when >= 2 stages still fit beside it;
        // 16 KiB stages want >= 3
        int st_k = 0;
        int ns_ref = pick_nstag(entry, stage_bytes_of(1), 3, &st_k);
        if (km == 1) {
          kmul = 1; st = st_k; nstag = ns_ref;
          break;
        }
        const int ns_min = (ns_ref >= 2) ? 2 : ns_ref;   // (never below two entries where the small-stage plan had them)
        int ns = ns_ref;
        int st2 = (pipe_avail - ns * entry) / stage_bytes_of(km);
        while (st2 < 2 && ns > ns_min) {
          ns >>= 1;
          st2 = (pipe_avail - ns * entry) / stage_bytes_of(km);
        }
        if (st2 >= 2) {
          kmul = km; st = st2; nstag = ns;
          break;
        }
      }
    }
    if (st < 2 && staged) {
      // the addend slots do not fit beside a 2-stage pipeline: keep pre[0] by TMA, gather the others (generic loads)
      staged = false;
      memset(c.add_off, 0, sizeof(c.add_off));
      memset(c.up_off, 0, sizeof(c.up_off));
      entry = stag_bytes;
      kmul = 1;
      nstag = (stag_bytes > 0 && p.n_tile <= 128) ? 2 : 1;
      st = (pipe_avail - nstag * entry) / stage_bytes_of(1);
      while (st < 3 && nstag > 1) {
        nstag >>= 1;
        st = (pipe_avail - nstag * entry) / stage_bytes_of(1);
      }
    }
    c.sub = (64 / p.ck) * kmul;
    c.a_bytes = c.vsh ? (p.bh + 2) * p.bw * p.ck * 2 : kStageABytes * kmul;
    c.a_region = (c.a_bytes + 1023) / 1024 * 1024;
    c.stage_bytes = stage_bytes_of(kmul);
    if (c.vsh) {
      uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)p.B};
      uint64_t strides[3] = {(uint64_t)p.Cin * 2, (uint64_t)p.Win * p.Cin * 2, (uint64_t)p.Hin * p.Win * p.Cin * 2};
      uint32_t box[4] = {(uint32_t)p.ck, (uint32_t)p.bw, (uint32_t)(p.bh + 2), 1u};
      int rc = conv_encode_map(&plan->maps.av, in, 4, dims, strides, box, p.ck);
      if (rc != HRP_OK) return rc;
    }
    HRP_REQUIRE(st >= 1, "layer does not fit in shared memory");
    c.res_tma = (has_res && (p.os == 1 || staged)) ? 1 : 0;
    c.staged = staged ? 1 : 0;
    c.entry_bytes = entry;
    c.stages = std::min(8, st);
    c.nstag = nstag;
    // two producers need >= 2 stages: with ONE slot a producer revisits it every second phase and a parity wait cannot
    // tell "two phases ago" from "now" (with >= 2 slots the in-order consumer bounds the distance to one phase)
    c.nprod = (c.stages >= 2 && !(e7 != nullptr && e7[0] == '1')) ? 2 : 1;
    // soft-argmax fold: two independent pipelines (see kThreadsPHead) when each gets at least two slots
    {
      const char* e8 = getenv("HRP_CONV_DUAL");
      const char* e9 = getenv("HRP_CONV_RES_STORE");
      const bool k_has_res = plan->epi == EPI_RES || staged || c.res_tma != 0;   // (= the kernel's has_res)
      c.res_store = (k_has_res && !(e9 != nullptr && e9[0] == '0')) ? 1 : 0;
      const bool m2 = plan->epi == EPI_HEAD || plan->epi == EPI_PLAIN || plan->epi == EPI_RES || staged;
      c.dual = (m2 && (!k_has_res || c.res_store) && !c.vsh && c.nprod == 2 && c.stages >= 4 &&
                !(e8 != nullptr && e8[0] == '0')) ? 1 : 0;
      if (c.dual) c.stages &= ~1;
    }
    c.stag_offset = c.pipe_offset + c.stages * c.stage_bytes;
    c.bar_offset = c.stag_offset + nstag * entry;
    c.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles * p.nphase;
    c.tw_shift = 0;
    while ((1 << c.tw_shift) < p.tiles_w) ++c.tw_shift;
    c.th_shift = 0;
    while ((1 << c.th_shift) < p.tiles_h) ++c.th_shift;
    HRP_REQUIRE((1 << c.tw_shift) == p.tiles_w && (1 << c.th_shift) == p.tiles_h, "tile counts must be powers of two");
    if (staged) plan->persistent = true;  // (only the persistent kernel stages addends; gathers cost 2-3x the layer time)
    const char* e3 = getenv("HRP_CONV_KSPLIT");
    // K-split accumulators are off by default: back-to-back MMAs into ONE accumulator already run at the operand
    // fetch rate (tools/probe_mma.py: 40.3 / 48.3 / 64.3 cycles at N = 32 / 64 / 128 with 1 or 4 accumulators), and
    // every extra accumulator costs the epilogue another TMEM read.  HRP_CONV_KSPLIT=2|4 re-enables it for probing.
    int ks = 1;
    if (e3 != nullptr && (e3[0] == '2' || e3[0] == '4')) ks = std::min((p.n_tile <= 64) ? 4 : (p.n_tile <= 128) ? 2 : 1, e3[0] - '0');
    const int n_mma = p.ntaps * p.cpt * (p.ck / 16);
    while (ks > 1 && n_mma < 2 * ks) ks >>= 1;  // every accumulator must receive at least one MMA (no stale TMEM)
    c.ksplit = ks;
    const char* e10 = getenv("HRP_CONV_NACC");
    c.nacc = (4 * ks * p.n_tile <= 512 && !(e10 != nullptr && e10[0] == '2')) ? 4 : 2;
    int cols = 32;
    while (cols < c.nacc * ks * p.n_tile) cols <<= 1;
    c.tmem_cols = cols;
    plan->psmem = c.bar_offset + tail;
    plan->pgrid = (unsigned)std::min(c.total_tiles, (c.wres && p.n_tiles > 1) ? grid_nt : num_sms);
  }
  // ---- halo-tile variant for narrow 3x3 stride-1 layers (conv_halo.cu) ----
  plan->halo_ok = plan->halo = false;
  if (desc != nullptr && p.n_tiles == 1 && p.n_tile == p.Cout) {
    int rc = conv_halo_plan(plan, *desc, in);
    if (rc != HRP_OK) return rc;
    plan->halo = plan->halo_ok;
  }
  // HRP_CONV_VARIANT=tile|persist|halo pins the kernel (tests, profiling); halo falls back where it is not eligible
  if (const char* v = getenv("HRP_CONV_VARIANT")) {
    const std::string vs(v);
    if (vs == "tile" && plan->epi != EPI_HEAD) { plan->persistent = false; plan->halo = false; }
    else if (vs == "persist") { plan->persistent = true; plan->halo = false; }
    else if (vs == "halo") { plan->halo = plan->halo_ok; }
  }
  return HRP_OK;
}

int conv_plan_launch(const ConvPlan& plan, cudaStream_t stream) {
  set_smem_attr_once();
  const ConvParams& p = plan.p;
  if (plan.halo) return conv_halo_launch(plan, stream);
  if (plan.epi == EPI_HEAD) {
    launch_ex(conv_gemm_persistent<64, EPI_HEAD>, dim3(plan.pgrid), dim3(kThreadsPHead), (size_t)plan.psmem, stream,
              plan.maps, p, plan.pcfg);
    count_launch();
    HRP_CUDA_CHECK(cudaGetLastError());
    return HRP_OK;
  }
  if (plan.persistent) {
#define HRP_LAUNCH_P(CKV, EPIV) \
  launch_ex(conv_gemm_persistent<CKV, EPIV>, dim3(plan.pgrid), dim3(PersistShape<EPIV>::threads), (size_t)plan.psmem, stream, \
            plan.maps, p, plan.pcfg)
#define HRP_LAUNCH_P_CK(CKV)                                  \
  do {                                                        \
    if (plan.epi == EPI_PLAIN) HRP_LAUNCH_P(CKV, EPI_PLAIN);   \
    else if (plan.epi == EPI_RES) HRP_LAUNCH_P(CKV, EPI_RES);  \
    else if (plan.epi == EPI_PRE && plan.pcfg.staged) HRP_LAUNCH_P(CKV, EPI_PRE_ST);  \
    else if (plan.epi == EPI_PRE) HRP_LAUNCH_P(CKV, EPI_PRE);  \
    else if (plan.pcfg.staged) HRP_LAUNCH_P(CKV, EPI_FULL_ST); \
    else HRP_LAUNCH_P(CKV, EPI_FULL);                         \
  } while (0)
    if (p.ck == 64) HRP_LAUNCH_P_CK(64);
    else if (p.ck == 32) HRP_LAUNCH_P_CK(32);
    else HRP_LAUNCH_P_CK(16);
#undef HRP_LAUNCH_P_CK
#undef HRP_LAUNCH_P
    count_launch();
    HRP_CUDA_CHECK(cudaGetLastError());
    return HRP_OK;
  }
#define HRP_LAUNCH(CKV, EPIV)                                                                              \
  launch_ex(conv_gemm_kernel<CKV, EPIV>, plan.grid, dim3(kNumThreads), (size_t)plan.smem_bytes, stream, plan.maps, p, \
            plan.stages, plan.bar_offset)
#define HRP_LAUNCH_CK(CKV)                                \
  do {                                                    \
    if (plan.epi == EPI_PLAIN) HRP_LAUNCH(CKV, EPI_PLAIN); \
    else if (plan.epi == EPI_RES) HRP_LAUNCH(CKV, EPI_RES); \
    else if (plan.epi == EPI_PRE) HRP_LAUNCH(CKV, EPI_PRE); \
    else HRP_LAUNCH(CKV, EPI_FULL);                       \
  } while (0)
  if (p.ck == 64) HRP_LAUNCH_CK(64);
  else if (p.ck == 32) HRP_LAUNCH_CK(32);
  else HRP_LAUNCH_CK(16);
#undef HRP_LAUNCH_CK
#undef HRP_LAUNCH
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

}  // namespace hrp
