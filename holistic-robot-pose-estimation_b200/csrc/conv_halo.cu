// Halo-tile 3x3 convolution for narrow layers (Cin = Cout = 32 or 64, stride 1, pad 1) on sm_100a.
//
// The implicit-GEMM kernels in conv_gemm.cu fetch one A tile per tap: nine L2 -> shared-memory copies of (almost) the
// same pixels.  For the HRNet high-resolution branches (32 ch @ 64x64, 64 ch @ 32x32: HRnet.py:28-57,197-242) and the
// 64-channel 3x3 convs of both layer1 stacks (Resnet.py:96-135, HRnet.py:60-98) that traffic -- not the tensor pipe,
// not HBM -- is the limit (measured: ~7.7 TB/s of TMA traffic for 0.25-0.5 PFLOP/s).  Here a CTA loads a band of
// input rows ONCE, with a one-pixel zero halo produced by the TMA unit's out-of-bounds fill, as a dense
// [rows][W+2][C] tile, and the nine taps are nine row-shifted 128-row windows of that tile: the UMMA shared-memory
// descriptor's start address is simply advanced by (dh*(W+2) + dw) rows.  The hardware applies the 32/64/128-byte
// swizzle to absolute shared-memory address bits, so a start that is not a multiple of the 8-row swizzle pattern
// is legal (verified by tools/probe_desc_shift.py on B200 for all three swizzle modes).
//
// GEMM rows are "positions" of the zero-padded image in row-major order, P = h*(W+2) + w with w in [0, W+2): the
// two positions per row with w >= W are junk (3 % of the rows at W = 64) and are never stored.  A work unit is T
// consecutive 128-position tiles of one image; its valid outputs are one CONTIGUOUS range of pixels of the NHWC
// tensor, so the epilogue compacts them into shared memory and a single bulk copy (cp.async.bulk) writes them out.
//
// Warp roles (320 threads, persistent, one CTA per SM): warp 0 = TMA producer (weights once, then one box per
// unit, multi-buffered), warp 1 = MMA issuer (warp-uniform control flow, one elected lane issues; accumulators
// ring-buffered in TMEM), warps 2..9 = two epilogue groups that take alternate tiles (tcgen05.ld, folded BN,
// residual, ReLU, bf16, staging, bulk store): at N = 32 a tile's MMAs last ~720 cycles, less than one group's
// epilogue latency.
#include "conv.h"
#include "launch_count.h"

#include <algorithm>
#include <mutex>

namespace hrp {

struct __align__(16) HaloBars {
  uint64_t w_full;
  uint64_t a_full[4];
  uint64_t a_empty[4];
  uint64_t acc_full[8];
  uint64_t acc_empty[8];
  uint64_t ring_full[6];    // output / residual ring: buffer may be used by the epilogue (residual tile landed, if any)
  uint64_t ring_ready[6];   // ... holds a finished output tile (one arrival per epilogue warp of the owning group)
  uint32_t tmem_base;
  uint32_t pad;
};

// debug timeline: slot (cta * 16 + event) <- %globaltimer (ns) for the first 8 CTAs
__device__ __forceinline__ void halo_stamp(long long* tl, int ev) {
  if (tl != nullptr && blockIdx.x < 8) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    tl[blockIdx.x * 16 + ev] = (long long)g;
  }
}

constexpr int kHaloThreads = 384;  // producer, MMA issuer A, 2 x 4 epilogue warps, MMA issuer B, store warp

// CTA-order tile iterator shared by every role: unit u (T tiles of one image), tile m inside it
struct HaloTileIter {
  int u, m;
  __device__ __forceinline__ bool valid(const HaloParams& hp) const { return u < hp.total_units; }
  __device__ __forceinline__ void next(const HaloParams& hp) {
    ++m;
    const int uu = u % hp.units_per_img;
    if (m >= hp.T || uu * hp.T + m >= hp.tiles_per_img) {
      m = 0;
      u += (int)gridDim.x;
    }
  }
};

// PAIR: the 64-channel positions are pixel pairs of a 32-channel layer.  Tap column dj = -1 then only feeds the LEFT
// output pixel of a pair from the RIGHT input pixel of the neighbouring pair (and dj = +1 the mirror image), so those
// taps are N = 32, K = 32 products: the packed matrix holds three full 64 x 64 tiles (dj = 0) and three tiles in which
// the dj = -1 block (rows 0..31, K columns 32..63) and the dj = +1 block (rows 32..63, K columns 0..31) share one tile.
template <int CK, int NOUT, bool RES, bool PAIR>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                    const __grid_constant__ CUtensorMap map_b,
                                                                    const __grid_constant__ CUtensorMap map_r,
                                                                    const __grid_constant__ HaloParams hp) {
  constexpr int KSTEPS = CK / 16;
  constexpr int ROWB = CK * 2;                    // bytes per position (one swizzle span)
  constexpr uint32_t LAYOUT = (CK == 64) ? 2u : 4u;  // SWIZZLE_128B / SWIZZLE_64B
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int B_SUB = NOUT * ROWB;              // one tap of the packed weights
  constexpr int NWT = PAIR ? 6 : 9;               // weight tiles resident in shared memory
  constexpr int STAG = kTileM * NOUT * 2;         // one ring buffer (output tile / residual tile)
  constexpr int CH16 = NOUT / 8;                  // 16-byte chunks per output pixel

  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_align1024(smem_dyn);
  uint8_t* const sW = smem;
  uint8_t* const sA = smem + hp.a_offset;
  uint8_t* const sRing = smem + hp.ring_offset;
  HaloBars* bars = reinterpret_cast<HaloBars*>(smem + hp.bar_offset);
  float* sb_smem = reinterpret_cast<float*>(bars + 1);  // [2][NOUT] folded-BN scale / shift

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Wp = hp.Wp, T = hp.T, upi = hp.units_per_img;
  const int NB = hp.nring;

  if (threadIdx.x == 0) {
    halo_stamp(hp.tl, 0);
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (RES) tma_prefetch_desc(&map_r);
    mbar_init(&bars->w_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars->a_full[i], 1);
      mbar_init(&bars->a_empty[i], 2);  // both MMA issuers release a band
    }
    for (int i = 0; i < 6; ++i) {
      mbar_init(&bars->ring_full[i], 1);
      mbar_init(&bars->ring_ready[i], 4);  // one arrival per epilogue warp of the owning group
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&bars->acc_full[i], 1);
      mbar_init(&bars->acc_empty[i], 4);  // one arrival per epilogue warp of the owning group
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, (uint32_t)(hp.nacc * NOUT));
    tmem_relinquish();
  }
  if (warp >= 2 && warp < 10) {
    for (int i = threadIdx.x - 64; i < NOUT; i += 256) {
      const int c = PAIR ? (i & 31) : i;
      sb_smem[i] = __ldg(hp.scale + c);
      sb_smem[NOUT + i] = __ldg(hp.bias + c);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  if (threadIdx.x == 0) halo_stamp(hp.tl, 1);
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();  // (the producer fetches the weights first: they are not produced by a kernel)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(&bars->w_full, (uint32_t)(NWT * B_SUB));
      for (int t = 0; t < NWT; ++t) tma_load_2d(sW + t * B_SUB, &map_b, &bars->w_full, t * CK, 0);
      pdl_wait();
      int abuf = 0;
      uint32_t par = 0;
      for (int u = blockIdx.x; u < hp.total_units; u += gridDim.x) {
        mbar_wait(&bars->a_empty[abuf], par ^ 1);
        const int n = u / upi, uu = u - n * upi;
        const int P0 = uu * T * kTileM;
        const int r_lo = (int)(((uint32_t)P0 * hp.div_magic) >> 20) - 1;  // first input row of the band (may be -1)
        mbar_expect_tx(&bars->a_full[abuf], (uint32_t)(hp.NR * Wp * ROWB));
        tma_load_4d(sA + (size_t)abuf * hp.a_buf_bytes, &map_a, &bars->a_full[abuf], 0, -1, r_lo, n);
        if (++abuf == hp.n_abuf) {
          abuf = 0;
          par ^= 1;
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===================== MMA issuers: warp 1 takes the even tiles (epilogue group 0), warp 10 the odd ones ==========
    // Two issuers hide each other's per-tile bookkeeping (barrier waits, fences): with one issuer the tensor pipe
    // drained its queue between tiles and idled ~40 % of the time (measured with the debug timeline).
    // The whole warp runs the (warp-uniform) control flow so that descriptors live in uniform registers; only the
    // tcgen05 instructions themselves are issued by one elected lane.  A dense back-to-back MMA stream matters here:
    // at N = 32..64 one MMA occupies the tensor pipe for only 40-48 cycles.
    {
      const uint32_t idesc = make_idesc_bf16(kTileM, (uint32_t)NOUT);
      const uint32_t idesc_half = make_idesc_bf16(kTileM, (uint32_t)(NOUT / 2));
      const uint32_t dhi = (uint32_t)(make_kmajor_desc(0, SBO, LAYOUT) >> 32);
      const uint32_t dlo = (uint32_t)make_kmajor_desc(0, SBO, LAYOUT);
      const uint32_t w_lo = dlo + (smem_u32(sW) >> 4);
      const uint32_t a_lo = dlo + (smem_u32(sA) >> 4);
      const uint32_t abuf_stride = (uint32_t)hp.a_buf_bytes >> 4;
      const int wp_units = Wp * (ROWB >> 4);
      const uint32_t mw = (warp == 1) ? 0u : 1u;
      mbar_wait(&bars->w_full, 0);
      if (lane == 0 && mw == 0) halo_stamp(hp.tl, 2);
      int abuf = 0;
      uint32_t par = 0;
      uint32_t tc = 0;
      const bool timing = (hp.tl != nullptr);  // debug only: cycles spent in each wait / in the issue loop
      long long t_afull = 0, t_acc = 0, t_issue = 0;
      const long long t_loop0 = timing ? clock64() : 0;
      for (int u = blockIdx.x; u < hp.total_units; u += gridDim.x) {
        const int n = u / upi, uu = u - n * upi;
        const int P0 = uu * T * kTileM;
        const int r_lo = (int)(((uint32_t)P0 * hp.div_magic) >> 20) - 1;
        const int base_row = P0 - r_lo * Wp + 1;  // tile-local row of output position P0 for tap (0, 0)
        const long long ta0 = timing ? clock64() : 0;
        mbar_wait(&bars->a_full[abuf], par);
        if (timing) t_afull += clock64() - ta0;
        tc_fence_after();
        if (tc == 0 && lane == 0 && mw == 0) halo_stamp(hp.tl, 3);
        const uint32_t a_unit0 = a_lo + (uint32_t)abuf * abuf_stride + (uint32_t)(base_row * (ROWB >> 4));
        for (int m = 0; m < T; ++m) {
          if (uu * T + m >= hp.tiles_per_img) break;
          if ((tc & 1u) != mw) {
            ++tc;
            continue;
          }
          const uint32_t acc = tc & (uint32_t)(hp.nacc - 1);
          const long long tb0 = timing ? clock64() : 0;
          mbar_wait(&bars->acc_empty[acc], ((tc >> hp.nacc_shift) & 1) ^ 1);
          const long long tb1 = timing ? clock64() : 0;
          t_acc += tb1 - tb0;
          tc_fence_after();
          const uint32_t taddr = tmem_base + acc * NOUT;
          const uint32_t a_tile = a_unit0 + (uint32_t)(m * kTileM * (ROWB >> 4));
          if (elect_one()) {
            if (PAIR) {
#pragma unroll
              for (int dh = -1; dh <= 1; ++dh) {
                const uint32_t ac = a_tile + (uint32_t)(dh * wp_units);
                const uint32_t bc = w_lo + (uint32_t)((dh + 1) * (B_SUB >> 4));        // dj = 0: full 64 x 64 tile
                const uint32_t bm = w_lo + (uint32_t)((dh + 4) * (B_SUB >> 4));        // merged dj = -1 / +1 tile
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k)
                  umma_bf16_ss(taddr, ((uint64_t)dhi << 32) | (ac + 2 * k), ((uint64_t)dhi << 32) | (bc + 2 * k), idesc,
                               (dh != -1 || k != 0) ? 1u : 0u);
                // dj = -1: left output pixels (accumulator columns 0..31) <- right input pixel of pair p-1 (K steps 2, 3)
#pragma unroll
                for (int k = KSTEPS / 2; k < KSTEPS; ++k)
                  umma_bf16_ss(taddr, ((uint64_t)dhi << 32) | (ac - (uint32_t)(ROWB >> 4) + 2 * k),
                               ((uint64_t)dhi << 32) | (bm + 2 * k), idesc_half, 1u);
                // dj = +1: right output pixels (columns 32..63, weight rows 32..63) <- left input pixel of pair p+1
#pragma unroll
                for (int k = 0; k < KSTEPS / 2; ++k)
                  umma_bf16_ss(taddr + (uint32_t)(NOUT / 2), ((uint64_t)dhi << 32) | (ac + (uint32_t)(ROWB >> 4) + 2 * k),
                               ((uint64_t)dhi << 32) | (bm + (uint32_t)((NOUT / 2) * ROWB >> 4) + 2 * k), idesc_half, 1u);
              }
            } else {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int dh = tap / 3 - 1, dw = tap % 3 - 1;
                const uint32_t al = a_tile + (uint32_t)(dh * wp_units + dw * (ROWB >> 4));
                const uint32_t bl = w_lo + (uint32_t)(tap * (B_SUB >> 4));
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k)
                  umma_bf16_ss(taddr, ((uint64_t)dhi << 32) | (al + 2 * k), ((uint64_t)dhi << 32) | (bl + 2 * k), idesc,
                               (tap != 0 || k != 0) ? 1u : 0u);
              }
            }
            umma_commit(&bars->acc_full[acc]);
          }
          __syncwarp();
          if (timing) t_issue += clock64() - tb1;
          ++tc;
        }
        if (elect_one()) umma_commit(&bars->a_empty[abuf]);  // the band can be overwritten once these MMAs have read it
        __syncwarp();
        if (++abuf == hp.n_abuf) {
          abuf = 0;
          par ^= 1;
        }
      }
      if (lane == 0 && mw == 0) halo_stamp(hp.tl, 4);
      if (timing && lane == 0 && blockIdx.x < 8 && mw == 0) {
        long long* o = hp.tl + blockIdx.x * 16;
        o[10] = t_afull;
        o[11] = t_acc;
        o[12] = t_issue;
        o[13] = clock64() - t_loop0;
        o[14] = (long long)tc;
      }
    }
  } else if (warp == 11) {
    // ===================== store warp: drains the output ring, prefetches the residual tiles ==========
    // Ring buffer b = tile % NB serves tile after tile: [residual tile lands (RES)] -> epilogue group updates it in place
    // -> bulk store to global -> residual of tile + NB is fetched into it.  One thread owns every bulk store and every
    // residual load, so the epilogue warps never wait for a store to drain and the residual prefetch distance is NB tiles.
    if (elect_one()) {
      const int HW = hp.H * hp.W;
      auto tile_range = [&](const HaloTileIter& it, int* pix_lo, int* npix) -> size_t {
        const int n = it.u / upi, uu = it.u - n * upi;
        const int Pt = (uu * T + it.m) * kTileM, Pe = Pt + kTileM;
        const int rf = (int)(((uint32_t)Pt * hp.div_magic) >> 20), wf = Pt - rf * Wp;
        const int re = (int)(((uint32_t)Pe * hp.div_magic) >> 20), we = Pe - re * Wp;
        const int lo = min(rf * hp.W + min(wf, hp.W), HW), hi = min(re * hp.W + min(we, hp.W), HW);
        *pix_lo = lo;
        *npix = hi - lo;
        return (size_t)n * HW;
      };
      auto fill = [&](const HaloTileIter& it, int buf) {  // make ring buffer `buf` usable for tile `it`
        if (RES) {
          int lo, np;
          const size_t ib = tile_range(it, &lo, &np);
          mbar_expect_tx(&bars->ring_full[buf], (uint32_t)STAG);
          tma_load_2d(sRing + (size_t)buf * STAG, &map_r, &bars->ring_full[buf], 0, (int)(ib + (size_t)lo));
        } else {
          mbar_arrive(&bars->ring_full[buf]);
        }
      };
      HaloTileIter ahead{(int)blockIdx.x, 0};  // next tile whose ring buffer has not been prepared yet
      for (int b = 0; b < NB && ahead.valid(hp); ++b) {
        fill(ahead, b);
        ahead.next(hp);
      }
      // The buffer of tile t is released (and refilled with the residual of tile t + NB) as soon as its store has been
      // read out of shared memory.  Experiment (profiles/r02_halo_sweep.txt): keeping LAG = NB - 2 stores in flight before
      // releasing a buffer left the plain kernels unchanged (61.8 vs 59.9 us: the store read is not the limiter) and slowed
      // the residual kernels (68.7 -> 90.4 us, 47.0 -> 54.8 us) because the residual prefetch distance drops to 2 tiles.
      const int LAG = 0;  // (LAG = NB - 2 measured slower: the residual prefetch distance shrinks to 2 tiles -- see below)
      int buf = 0, rel_buf = 0, t_idx = 0;
      uint32_t par = 0;
      for (HaloTileIter it{(int)blockIdx.x, 0}; it.valid(hp); it.next(hp), ++t_idx) {
        mbar_wait(&bars->ring_ready[buf], par);
        int lo, np;
        const size_t ib = tile_range(it, &lo, &np);
        if (np > 0) {
          bf16* dst = hp.out + (ib + lo) * NOUT;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                       "r"(smem_u32(sRing + (size_t)buf * STAG)), "r"((uint32_t)(np * NOUT * 2))
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (t_idx >= LAG) {
          switch (LAG) {  // (the group count is an immediate)
            case 0: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
            case 1: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
            case 2: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
            case 3: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
            default: asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); break;
          }
          if (ahead.valid(hp)) {  // the store of tile t_idx - LAG has been read out: its buffer serves tile t_idx - LAG + NB
            fill(ahead, rel_buf);
            ahead.next(hp);
          }
          if (++rel_buf == NB) rel_buf = 0;
        }
        if (++buf == NB) {
          buf = 0;
          par ^= 1;
        }
      }
      halo_stamp(hp.tl, 5);
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      halo_stamp(hp.tl, 7);
    }
  } else {
    // ===================== epilogue: warps 2..9, TMEM lane quarter = warp % 4, group = (warp - 2) / 4 ==========
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int HW = hp.H * hp.W;
    HaloTileIter it{(int)blockIdx.x, 0};
    uint32_t tc = 0;
    int buf = 0;
    uint32_t rpar = 0;
    if (grp == 1 && it.valid(hp)) {  // group 1 takes the odd tiles
      it.next(hp);
      tc = 1;
      buf = 1 % NB;
      rpar = (NB == 1) ? 1u : 0u;
    }
    while (it.valid(hp)) {
      const int n = it.u / upi, uu = it.u - n * upi;
      const int Pt = (uu * T + it.m) * kTileM;
      // this thread's output position and the tile's first pixel
      const int P = Pt + row;
      const int r = (int)(((uint32_t)P * hp.div_magic) >> 20), w = P - r * Wp;
      const bool valid = (w < hp.W) && (r < hp.H);
      const int rf = (int)(((uint32_t)Pt * hp.div_magic) >> 20), wf = Pt - rf * Wp;
      const int pix_lo = min(rf * hp.W + min(wf, hp.W), HW);
      const int sr = r * hp.W + w - pix_lo;  // this thread's row in the ring buffer (valid lanes only)
      const uint32_t acc = tc & (uint32_t)(hp.nacc - 1);
      uint8_t* const rbuf = sRing + (size_t)buf * STAG;
      mbar_wait(&bars->acc_full[acc], (tc >> hp.nacc_shift) & 1);
      tc_fence_after();
      mbar_wait(&bars->ring_full[buf], rpar);  // store of tile - NB drained (and this tile's residual landed)
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * NOUT;
      const uint8_t* const rrow = rbuf + (size_t)sr * (NOUT * 2);
      const int rsw = (CH16 == 8) ? (sr & 7) : ((sr >> 1) & 3);
      uint4 o[CH16];
      // the whole accumulator row (32 or 64 columns) is fetched with one round trip: both loads in flight, one wait
      uint32_t acc_all[NOUT];
#pragma unroll
      for (int c0 = 0; c0 < NOUT; c0 += 32) tmem_ld32(taddr + (uint32_t)c0, reinterpret_cast<uint32_t(&)[32]>(acc_all[c0]));
      tmem_ld_wait();
      if (hp.dbg & 1) {  // (ablation: no arithmetic, no shared-memory traffic -- the protocol is kept)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars->acc_empty[acc]);
          mbar_arrive(&bars->ring_ready[buf]);
        }
        it.next(hp);
        if (it.valid(hp)) it.next(hp);
        tc += 2;
        buf += 2;
        if (buf >= NB) {
          buf -= NB;
          rpar ^= 1;
        }
        continue;
      }
#pragma unroll
      for (int c0 = 0; c0 < NOUT; c0 += 32) {
        const uint32_t* a = acc_all + c0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cg = c0 + g * 8;
          float v[8];
          const float4 s0 = *reinterpret_cast<const float4*>(sb_smem + cg);
          const float4 s1 = *reinterpret_cast<const float4*>(sb_smem + cg + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(sb_smem + NOUT + cg);
          const float4 b1 = *reinterpret_cast<const float4*>(sb_smem + NOUT + cg + 4);
          v[0] = fmaf(__uint_as_float(a[g * 8 + 0]), s0.x, b0.x);
          v[1] = fmaf(__uint_as_float(a[g * 8 + 1]), s0.y, b0.y);
          v[2] = fmaf(__uint_as_float(a[g * 8 + 2]), s0.z, b0.z);
          v[3] = fmaf(__uint_as_float(a[g * 8 + 3]), s0.w, b0.w);
          v[4] = fmaf(__uint_as_float(a[g * 8 + 4]), s1.x, b1.x);
          v[5] = fmaf(__uint_as_float(a[g * 8 + 5]), s1.y, b1.y);
          v[6] = fmaf(__uint_as_float(a[g * 8 + 6]), s1.z, b1.z);
          v[7] = fmaf(__uint_as_float(a[g * 8 + 7]), s1.w, b1.w);
          if (RES && valid) {
            const uint4 x = *reinterpret_cast<const uint4*>(rrow + ((((cg >> 3)) ^ rsw) << 4));
            v[0] += bf16lo_to_f32(x.x); v[1] += bf16hi_to_f32(x.x);
            v[2] += bf16lo_to_f32(x.y); v[3] += bf16hi_to_f32(x.y);
            v[4] += bf16lo_to_f32(x.z); v[5] += bf16hi_to_f32(x.z);
            v[6] += bf16lo_to_f32(x.w); v[7] += bf16hi_to_f32(x.w);
          }
          uint4& oo = o[cg >> 3];
          if (hp.relu) {
            asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(oo.x) : "f"(v[1]), "f"(v[0]));
            asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(oo.y) : "f"(v[3]), "f"(v[2]));
            asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(oo.z) : "f"(v[5]), "f"(v[4]));
            asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(oo.w) : "f"(v[7]), "f"(v[6]));
          } else {
            oo.x = pack_bf16x2(v[0], v[1]);
            oo.y = pack_bf16x2(v[2], v[3]);
            oo.z = pack_bf16x2(v[4], v[5]);
            oo.w = pack_bf16x2(v[6], v[7]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
      // The output row replaces the residual row IN PLACE: this thread has read all the chunks of its row above, and a
      // row is touched by one thread only (the TMA swizzle and the rotation below both permute chunks inside a row).
      // Ring rows are dense pixels (64 or 128 bytes apart): written chunk-by-chunk in lane order every 16-byte store
      // would hit the same 4 banks (8- / 16-way conflict).  Row sr therefore writes its chunks in the rotated order
      // (c + rot(sr)) % CH16, which spreads every quarter-warp over all 32 banks.
      if (valid) {
        const int rot = (CH16 == 8) ? (sr & 7) : ((sr >> 1) & 3);
#pragma unroll
        for (int b = 1; b < CH16; b <<= 1) {  // rotate o[] left by rot (log steps, register selects only)
          const bool on = (rot & b) != 0;
          uint4 t[CH16];
#pragma unroll
          for (int i = 0; i < CH16; ++i) {
            const uint4 x = o[(i + b) % CH16], y = o[i];
            t[i].x = on ? x.x : y.x;
            t[i].y = on ? x.y : y.y;
            t[i].z = on ? x.z : y.z;
            t[i].w = on ? x.w : y.w;
          }
#pragma unroll
          for (int i = 0; i < CH16; ++i) o[i] = t[i];
        }
        uint8_t* const srow = rbuf + (size_t)sr * (NOUT * 2);
#pragma unroll
        for (int c = 0; c < CH16; ++c) *reinterpret_cast<uint4*>(srow + (((c + rot) & (CH16 - 1)) << 4)) = o[c];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk copy
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->ring_ready[buf]);
      // this group's next tile: two tiles further in CTA order
      it.next(hp);
      if (it.valid(hp)) it.next(hp);
      tc += 2;
      buf += 2;
      if (buf >= NB) {
        buf -= NB;
        rpar ^= 1;
      }
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)(hp.nacc * NOUT));
  }
  if (threadIdx.x == 0) halo_stamp(hp.tl, 9);
}

// ------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------
static void halo_attr_once() {
  static std::once_flag once;
  std::call_once(once, [] {
#define HRP_HALO_ATTR(CKV, NV, RV, PV) \
  cudaFuncSetAttribute(conv_halo_kernel<CKV, NV, RV, PV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
    HRP_HALO_ATTR(32, 32, false, false);
    HRP_HALO_ATTR(32, 32, true, false);
    HRP_HALO_ATTR(64, 64, false, false);
    HRP_HALO_ATTR(64, 64, true, false);
    HRP_HALO_ATTR(64, 64, false, true);
    HRP_HALO_ATTR(64, 64, true, true);
#undef HRP_HALO_ATTR
  });
}

void conv_halo_init() { halo_attr_once(); }

// Decide whether the layer can run on the halo kernel and, if so, fill plan->halo_* (tensor maps, tiling, smem plan).
int conv_halo_plan(ConvPlan* plan, const ConvLayerDesc& d, const bf16* in) {
  plan->halo_ok = false;
  const ConvParams& p = plan->p;
  const char* env = getenv("HRP_CONV_HALO");
  if (env != nullptr && env[0] == '0') return HRP_OK;
  if (d.kind != kConv || d.stride != 1 || d.kh != 3 || d.kw != 3 || d.pad != 1) return HRP_OK;
  if (!((d.Cin == 32 && d.Cout == 32) || (d.Cin == 64 && d.Cout == 64))) return HRP_OK;
  if (p.out == nullptr || p.pool_out != nullptr || p.post != nullptr) return HRP_OK;
  if (p.pre[1] != nullptr || p.pre[2] != nullptr || p.up[0] != nullptr || p.up[1] != nullptr || p.up[2] != nullptr)
    return HRP_OK;
  if (p.n_tiles != 1 || p.n_tile != d.Cout || p.ck != d.Cin) return HRP_OK;
  // pixel-pair view of a 32-channel layer (weights packed behind the regular matrix by conv_pack_weights)
  const char* penv = getenv("HRP_HALO_PAIR");
  const bool pair = (p.pair_off > 0) && !(penv != nullptr && penv[0] == '0');
  const int C = pair ? 64 : d.Cin;  // channels per position (= Cin = Cout of the problem the kernel sees)
  const int H = d.Hin, W = pair ? d.Win / 2 : d.Win, Wp = W + 2;
  if (Wp > 256 || H * Wp > 60000) return HRP_OK;
  HaloParams& h = plan->hp;
  memset(&h, 0, sizeof(h));
  h.B = d.B;
  h.H = H;
  h.W = W;
  h.Cout = C;
  h.pair = pair ? 1 : 0;
  h.relu = d.relu;
  h.Wp = Wp;
  h.tiles_per_img = (H * Wp + kTileM - 1) / kTileM;
  // P / Wp as a multiply-shift, verified exhaustively over every position the kernel can form
  h.div_magic = (uint32_t)(((1u << 20) + Wp - 1) / Wp);
  const int p_max = (h.tiles_per_img + 4) * kTileM + kTileM;
  for (int P = 0; P <= p_max; ++P)
    if ((int)(((uint64_t)(uint32_t)P * h.div_magic) >> 20) != P / Wp || (uint64_t)P * h.div_magic >= (1ull << 32))
      return HRP_OK;  // (never happens for the shapes on the path; stay on the generic kernels if it does)
  const int rowb = C * 2;
  const int w_bytes = (pair ? 6 : 9) * C * rowb;   // pair view: three full taps + three merged half-tap tiles
  // Buffering.  Output / residual ring (`nring` tiles): the residual tile is fetched into the buffer the output tile is then
  // built in (in place), so the ring depth is also the residual prefetch distance.  Bands: T tiles per band, nbuf bands; the
  // MMA warps run (nbuf - 1) * T tiles ahead of the band being loaded and need ~2 tiles to cover a DRAM round trip.
  // Measured on B200 at 512 images (profiles/r02_halo_sweep.txt, us per launch, round-1 kernel -> best configuration):
  //   pixel pairs (32 ch @ 64x64)      67.6 -> 59.9  T=1 bands=3 ring=4   (T=2/b3/r2 62.5, T=3/b2/r2 65.9)
  //   pixel pairs + residual            95.8 -> 69.7  T=1 bands=3 ring=4   (ring=3 76.2, ring=2 98.4)
  //   64 ch @ 32x32                     47.2 -> 43.3  T=3 bands=2 ring=2   (T=2/b2 47.8, T=1/b4 48.6)
  //   64 ch @ 32x32 + residual          59.2 -> 47.7  T=1 bands=3 ring=3   (T=2/b2/r3 50.1, ring=4/T=1/b2 55.0)
  // i.e. a deep ring wins whenever it still leaves two tiles of band look-ahead; otherwise large bands win.
  const bool res = (p.pre[0] != nullptr);
  const int tail = (int)sizeof(HaloBars) + 2 * C * 4 + 1024;
  const int w_pad = (w_bytes + 1023) / 1024 * 1024;
  struct Cand { int T, ring; };
  static const Cand pref_pair[] = {{1, 4}, {2, 4}, {2, 3}, {2, 2}, {1, 3}, {1, 2}};
  static const Cand pref_res[] = {{1, 3}, {2, 3}, {1, 4}, {1, 2}};
  static const Cand pref_plain[] = {{3, 2}, {2, 2}, {4, 2}, {1, 2}};
  const Cand* pref = pair ? pref_pair : (res ? pref_res : pref_plain);
  const int npref = pair ? 6 : 4;
  const char* eT = getenv("HRP_HALO_T");
  const char* eN = getenv("HRP_HALO_NBUF");
  const char* eR = getenv("HRP_HALO_NRING");
  int best_T = 0, best_NR = 0, best_nbuf = 0, nring = 0;
  for (int ci = 0; ci < npref && best_T == 0; ++ci) {
    const int T = (eT != nullptr) ? atoi(eT) : pref[ci].T;
    const int ring = (eR != nullptr) ? std::max(2, std::min(6, atoi(eR))) : pref[ci].ring;
    if (T < 1 || T > 4) continue;
    const int upi = (h.tiles_per_img + T - 1) / T;
    int NR = 0;
    for (int uu = 0; uu < upi; ++uu) {
      const int P0 = uu * T * kTileM;
      const int P1 = std::min(P0 + T * kTileM, h.tiles_per_img * kTileM) - 1;
      NR = std::max(NR, P1 / Wp - P0 / Wp + 3);
    }
    if (NR > 256) continue;
    const int a_buf = (NR * Wp * rowb + 1023) / 1024 * 1024;
    const int avail = 227 * 1024 - tail - ring * kTileM * C * 2 - w_pad;
    int nbuf = (avail > 0) ? std::min(4, avail / a_buf) : 0;
    if (eN != nullptr) nbuf = std::min(nbuf, std::max(1, atoi(eN)));
    const int look = (nbuf - 1) * T;
    if (nbuf < 2 || (look < 2 && ci + 1 < npref && eT == nullptr)) continue;  // (the last candidate may run with 1 tile ahead)
    best_T = T;
    best_NR = NR;
    best_nbuf = nbuf;
    nring = ring;
  }
  const int ring_bytes = nring * kTileM * C * 2;
  if (best_T == 0) return HRP_OK;
  h.T = best_T;
  h.NR = best_NR;
  h.n_abuf = best_nbuf;
  h.nring = nring;
  h.units_per_img = (h.tiles_per_img + h.T - 1) / h.T;
  h.total_units = h.units_per_img * d.B;
  h.a_buf_bytes = (h.NR * Wp * rowb + 1023) / 1024 * 1024;
  h.a_offset = (w_bytes + 1023) / 1024 * 1024;
  h.ring_offset = h.a_offset + h.n_abuf * h.a_buf_bytes;
  h.bar_offset = h.ring_offset + ring_bytes;
  h.scale = p.scale;
  h.bias = p.bias;
  h.res = p.pre[0];
  h.out = p.out;
  // accumulator ring in TMEM: a power of two, at most 8 and at most 512 columns
  h.nacc = (8 * C <= 512) ? 8 : 4;
  h.nacc_shift = (h.nacc == 8) ? 3 : 2;
  plan->halo_smem = h.bar_offset + tail;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)d.B};
    uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {(uint32_t)C, (uint32_t)Wp, (uint32_t)h.NR, 1u};
    int rc = conv_encode_map(&plan->halo_map_a, in, 4, dims, strides, box, C);
    if (rc != HRP_OK) return rc;
  }
  plan->halo_map_r = plan->halo_map_a;
  if (p.pre[0] != nullptr) {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)d.B * H * W};
    uint64_t strides[1] = {(uint64_t)C * 2};
    uint32_t box[2] = {(uint32_t)C, (uint32_t)kTileM};
    int rc = conv_encode_map(&plan->halo_map_r, p.pre[0], 2, dims, strides, box, C);
    if (rc != HRP_OK) return rc;
  }
  if (pair) {
    uint64_t dims[2] = {(uint64_t)6 * 64, 64};
    uint64_t strides[1] = {(uint64_t)6 * 64 * 2};
    uint32_t box[2] = {64u, 64u};
    int rc = conv_encode_map(&plan->halo_map_b, p.w + p.pair_off, 2, dims, strides, box, 64);
    if (rc != HRP_OK) return rc;
  }
  int num_sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (num_sms <= 0) num_sms = 148;
  plan->halo_grid = (unsigned)std::min(h.total_units, num_sms);
  plan->halo_ok = true;
  return HRP_OK;
}

int conv_halo_launch(const ConvPlan& plan, cudaStream_t stream) {
  halo_attr_once();
  HaloParams h = plan.hp;
  h.tl = plan.p.timeline;
  h.dbg = (getenv("HRP_HALO_DBG") != nullptr) ? atoi(getenv("HRP_HALO_DBG")) : 0;
  const bool res = (h.res != nullptr);
#define HRP_HALO_LAUNCH(CKV, NV, RV, PV, MAPB) \
  launch_ex(conv_halo_kernel<CKV, NV, RV, PV>, dim3(plan.halo_grid), dim3(kHaloThreads), (size_t)plan.halo_smem, stream, \
            plan.halo_map_a, MAPB, plan.halo_map_r, h)
  if (h.pair) {
    if (res) HRP_HALO_LAUNCH(64, 64, true, true, plan.halo_map_b);
    else HRP_HALO_LAUNCH(64, 64, false, true, plan.halo_map_b);
  } else if (h.Cout == 32) {
    if (res) HRP_HALO_LAUNCH(32, 32, true, false, plan.maps.b);
    else HRP_HALO_LAUNCH(32, 32, false, false, plan.maps.b);
  } else {
    if (res) HRP_HALO_LAUNCH(64, 64, true, false, plan.maps.b);
    else HRP_HALO_LAUNCH(64, 64, false, false, plan.maps.b);
  }
#undef HRP_HALO_LAUNCH
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

}  // namespace hrp
