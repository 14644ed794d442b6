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
#ifdef __cplusplus
}
#endif
#endif
