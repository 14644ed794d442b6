/* hrp_probe.h -- hardware probes (profiling aids, NOT part of the product ABI in include/hrp.h).
 * Compiled into libhrp_b200.so only when the library is built with HRP_BUILD_PROBES=1
 * (holistic-robot-pose-estimation_b200/build.py); used by tools/probe_*.py. */
#ifndef HRP_PROBE_H_
#define HRP_PROBE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* hardware probe (profiling aid): cycles to issue / complete `reps` back-to-back tcgen05.mma (SS mode, bf16, K=16)
 * of shape M x N on `ctas` CTAs; dev_out2 = {issue cycles, completion cycles} of CTA 0 */
int hrp_probe_mma_rate(int32_t M, int32_t N, int32_t reps, int32_t kdistinct, long long* dev_out2, int32_t ctas);
/* hardware probe: one 128x32xCK product whose swizzled K-major A operand starts `shift` rows into a [rows x ck] bf16
 * shared-memory tile (bo_mode 1 sets the descriptor base_offset field); D (128x32 fp32) is written to out_dev */
int hrp_probe_desc_shift(int32_t ck, int32_t rows, int32_t shift, int32_t bo_mode, const void* A_dev, const void* B_dev,
                         float* out_dev);
/* hardware probe: cycles for `warps` (1, 4 or 8) warps of one CTA to issue `reps` tcgen05.ld.32x32b.x32 each (4 KiB per
 * instruction), waiting after every load (wait_each = 1) or only at the end; dev_out2[0] = cycles */
int hrp_probe_tmem_ld_rate(int32_t warps, int32_t reps, int32_t wait_each, long long* dev_out2);
/* hardware probe: TMA tiled-load rate per SM for the conv kernels' A-operand boxes {64 ch, bw, bh, bn} (16 KiB, 128 rows of
 * 128 B at stride C * 2 B) of an NHWC bf16 tensor, `depth` loads in flight per CTA, `share` CTAs reading the same tiles;
 * rows > 0: 2-D box {64 ch, rows pixels} on the (B*H*W, C) view instead; producers: low byte = 1 or 2 issuing threads (two warps), second byte = wait flavour (0 try_wait spin, 1 test_wait spin, 2 try_wait with a 0 ns suspend hint); dev_out[0] = cycles CTA 0 needed for `iters` loads */
int hrp_probe_tma_rate(const void* act, int32_t B, int32_t H, int32_t W, int32_t C, int32_t bw, int32_t bh, int32_t bn,
                       int32_t depth, int32_t iters, int32_t share, int32_t ctas, int32_t rows, int32_t producers,
                       long long* dev_out);
/* hardware probe: one thread issues n back-to-back 2-D TMA loads of {64 ch, rows} boxes on one mbarrier;
 * dev_out = {cycles until the last one has issued, cycles until all bytes have landed} (mean over reps - 1 rounds) */
int hrp_probe_tma_issue(const void* act, int64_t npix, int32_t C, int32_t rows, int32_t n, int32_t reps, int32_t ctas,
                        long long* dev_out);
#ifdef __cplusplus
}
#endif
#endif
