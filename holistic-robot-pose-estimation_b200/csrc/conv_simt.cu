// SIMT cross-check for the tcgen05 implicit-GEMM convolution: same plan (conv.h), no TMA / tensor cores.
// One thread per (row pixel, output channel), fp32 accumulation over bf16 operands.  Used by tests and by
// HRP_CONV_IMPL=simt debugging only -- the shipped forward never takes this path.
#include "conv.h"
#include "launch_count.h"

namespace hrp {

__global__ void conv_simt_kernel(const ConvParams p) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;  // output channel
  const size_t rowpix = blockIdx.x;
  const int phase = blockIdx.z;
  const int ph = phase >> 1, pw = phase & 1;
  if (c >= p.Cout) return;
  const int w = (int)(rowpix % p.Wm);
  const int h = (int)((rowpix / p.Wm) % p.Hm);
  const int n = (int)(rowpix / ((size_t)p.Wm * p.Hm));
  float acc = 0.f;
  const bf16* wrow = p.w + ((size_t)(p.shared_phase ? 0 : phase) * p.cout_pad + c) * p.ktot;
  for (int t = 0; t < p.ntaps; ++t) {
    const int hs = h + p.tap_dh[t] + (p.shared_phase ? 0 : ph), ws = w + p.tap_dw[t] + (p.shared_phase ? 0 : pw);
    if (hs < 0 || hs >= p.Hs || ws < 0 || ws >= p.Ws) continue;
    const int hp = p.tap_map[t] >> 1, wp = p.tap_map[t] & 1;
    const bf16* src = p.in + p.src_off + ((size_t)hp * p.Win + wp) * p.Cin + (size_t)n * p.src_img +
                      (size_t)hs * p.src_row + (size_t)ws * p.src_pix;
    const bf16* wt = wrow + (size_t)t * p.Cin;
    for (int ci = 0; ci < p.Cin; ++ci) acc = fmaf(__bfloat162float(src[ci]), __bfloat162float(wt[ci]), acc);
  }
  const int oh = h * p.os + p.oh0 + ph, ow = w * p.os + p.ow0 + pw;
  const size_t opix = ((size_t)n * p.Hout + oh) * p.Wout + ow;
  float v = fmaf(acc, p.scale[c], p.bias[c]);
  for (int a = 0; a < 3; ++a)
    if (p.pre[a] != nullptr) v += __bfloat162float(p.pre[a][opix * p.Cout + c]);
  for (int a = 0; a < 3; ++a)
    if (p.up[a] != nullptr) {
      const int sh = p.up_shift[a];
      const size_t upix = ((size_t)n * (p.Hout >> sh) + (oh >> sh)) * (p.Wout >> sh) + (ow >> sh);
      v += __bfloat162float(p.up[a][upix * p.Cout + c]);
    }
  if (p.relu) v = fmaxf(v, 0.f);
  if (p.post != nullptr) v += __bfloat162float(p.post[opix * p.Cout + c]);
  if (p.out != nullptr) p.out[opix * p.Cout + c] = __float2bfloat16_rn(v);
  if (p.pool_out != nullptr) atomicAdd(p.pool_out + (size_t)n * p.Cout + c, v * p.pool_scale);
}

int conv_plan_launch_simt(const ConvPlan& plan, cudaStream_t stream) {
  const ConvParams& p = plan.p;
  const int threads = 64;
  dim3 grid((unsigned)((size_t)p.B * p.Hm * p.Wm), (unsigned)((p.Cout + threads - 1) / threads), (unsigned)p.nphase);
  conv_simt_kernel<<<grid, threads, 0, stream>>>(p);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

}  // namespace hrp
