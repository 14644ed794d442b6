// Memory-bound layout kernels around the conv stack: input packing (fp32 NCHW -> bf16 NHWC space-to-depth),
// max-pool, NCHW<->NHWC bridges.  All are coalesced on the NHWC side and vectorised to 16 B per thread.
#include "hrp_common.cuh"
#include "launch_count.h"
#include "ops.h"

#include <algorithm>

namespace hrp {

// (B,3,H,W) fp32 -> (B,H/2,W/2,16) bf16; channel = (hp*2+wp)*3 + c; 12..15 = 0.
// One thread per s2d pixel: reads 3 channels x 2 rows x 2 adjacent floats (float2, coalesced along W),
// writes 32 contiguous bytes.
__global__ void pack_input_s2d_kernel(const float* __restrict__ x, uint4* __restrict__ out, int B, int H, int W,
                                      int out_pitch, int out_off) {
  const int Ws = W >> 1, Hs = H >> 1;
  const size_t total = (size_t)B * Hs * Ws;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ws = (int)(i % Ws);
    const int hs = (int)((i / Ws) % Hs);
    const int n = (int)(i / ((size_t)Ws * Hs));
    float v[16];
#pragma unroll
    for (int k = 12; k < 16; ++k) v[k] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int hp = 0; hp < 2; ++hp) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(x + (((size_t)n * 3 + c) * H + (2 * hs + hp)) * W + 2 * ws));
        v[(hp * 2 + 0) * 3 + c] = t.x;
        v[(hp * 2 + 1) * 3 + c] = t.y;
      }
    }
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0], v[1]);
    o0.y = pack_bf16x2(v[2], v[3]);
    o0.z = pack_bf16x2(v[4], v[5]);
    o0.w = pack_bf16x2(v[6], v[7]);
    o1.x = pack_bf16x2(v[8], v[9]);
    o1.y = pack_bf16x2(v[10], v[11]);
    o1.z = pack_bf16x2(v[12], v[13]);
    o1.w = pack_bf16x2(v[14], v[15]);
    // output rows may be padded (out_pitch pixels per row, image at column out_off): see kStemS2D in conv.h
    const size_t o = ((size_t)n * Hs + hs) * out_pitch + out_off + ws;
    out[2 * o] = o0;
    out[2 * o + 1] = o1;
  }
}

// scalar uint8 variant (any even W): one thread per s2d pixel; reads 2 bytes (uchar2) per channel-row.
__global__ void pack_input_s2d_u8_kernel(const uint8_t* __restrict__ x, uint4* __restrict__ out, int B, int H, int W,
                                         int out_pitch, int out_off) {
  const int Ws = W >> 1, Hs = H >> 1;
  const size_t total = (size_t)B * Hs * Ws;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ws = (int)(i % Ws);
    const int hs = (int)((i / Ws) % Hs);
    const int n = (int)(i / ((size_t)Ws * Hs));
    float v[16];
#pragma unroll
    for (int k = 12; k < 16; ++k) v[k] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int hp = 0; hp < 2; ++hp) {
        const uchar2 t = __ldg(reinterpret_cast<const uchar2*>(x + (((size_t)n * 3 + c) * H + (2 * hs + hp)) * W + 2 * ws));
        v[(hp * 2 + 0) * 3 + c] = (float)t.x / 255.0f;
        v[(hp * 2 + 1) * 3 + c] = (float)t.y / 255.0f;
      }
    }
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0], v[1]);
    o0.y = pack_bf16x2(v[2], v[3]);
    o0.z = pack_bf16x2(v[4], v[5]);
    o0.w = pack_bf16x2(v[6], v[7]);
    o1.x = pack_bf16x2(v[8], v[9]);
    o1.y = pack_bf16x2(v[10], v[11]);
    o1.z = pack_bf16x2(v[12], v[13]);
    o1.w = pack_bf16x2(v[14], v[15]);
    // output rows may be padded (out_pitch pixels per row, image at column out_off): see kStemS2D in conv.h
    const size_t o = ((size_t)n * Hs + hs) * out_pitch + out_off + ws;
    out[2 * o] = o0;
    out[2 * o + 1] = o1;
  }
}

// uint8 variant: fuses the caller-side `images.float() / 255.` of scripts/test.py:83-86 into the packing pass.
// One thread per FOUR consecutive s2d pixels of a row: six 8-byte loads (3 channels x 2 input rows x 8 pixels) and
// 128 contiguous output bytes (the one-pixel-per-thread version issued six 2-byte loads per 32 output bytes and ran at
// 27 % of HBM peak).  Requires W % 8 == 0 (the launcher falls back to the scalar kernel otherwise).
__global__ void __launch_bounds__(256) pack_input_s2d_u8x4_kernel(const uint8_t* __restrict__ x, uint4* __restrict__ out,
                                                                 int B, int H, int W, int out_pitch, int out_off) {
  const int Wq = W >> 3, Hs = H >> 1;
  const size_t total = (size_t)B * Hs * Wq;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;  // total < 2^31: 32-bit div / mod
    const int wq = (int)(iu % (unsigned)Wq);
    const unsigned rowi = iu / (unsigned)Wq;
    const int hs = (int)(rowi % (unsigned)Hs);
    const int n = (int)(rowi / (unsigned)Hs);
    uint2 raw[3][2];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int hp = 0; hp < 2; ++hp)
        raw[c][hp] = __ldg(reinterpret_cast<const uint2*>(x + (((size_t)n * 3 + c) * H + (2 * hs + hp)) * W + 8 * wq));
    uint4* dst = out + 2 * (((size_t)n * Hs + hs) * out_pitch + out_off + 4 * wq);
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // s2d pixel j of the group = input columns 2j, 2j+1
      float v[12];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int hp = 0; hp < 2; ++hp) {
          const uint32_t word = (j < 2) ? raw[c][hp].x : raw[c][hp].y;
          const uint32_t two = (word >> (16 * (j & 1))) & 0xffffu;
          v[(hp * 2 + 0) * 3 + c] = (float)(two & 0xffu) / 255.0f;
          v[(hp * 2 + 1) * 3 + c] = (float)(two >> 8) / 255.0f;
        }
      uint4 o0, o1;
      o0.x = pack_bf16x2(v[0], v[1]);
      o0.y = pack_bf16x2(v[2], v[3]);
      o0.z = pack_bf16x2(v[4], v[5]);
      o0.w = pack_bf16x2(v[6], v[7]);
      o1.x = pack_bf16x2(v[8], v[9]);
      o1.y = pack_bf16x2(v[10], v[11]);
      o1.z = 0u;
      o1.w = 0u;
      dst[2 * j] = o0;
      dst[2 * j + 1] = o1;
    }
  }
}

// MaxPool2d(kernel 3, stride 2, pad 1) on bf16 NHWC; thread = (output pixel, 8-channel group)
__global__ void maxpool3x3s2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W,
                                    int C8) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H >> 1, Wo = W >> 1;
  const size_t total = (size_t)B * Ho * Wo * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;  // total < 2^31 (checked by the launcher): 32-bit div / mod
    const int cg = (int)(iu % (unsigned)C8);
    const unsigned pix = iu / (unsigned)C8;
    const int wo = (int)(pix % (unsigned)Wo);
    const unsigned rowi = pix / (unsigned)Wo;
    const int ho = (int)(rowi % (unsigned)Ho);
    const int n = (int)(rowi / (unsigned)Ho);
    // max of bf16 values is exact in bf16: four packed HMNMX2 per tap instead of eight conversions + eight FMNMX
    // (this streaming kernel was issue-bound: 338 us for 1.34 GB at 512 images)
    const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
    __nv_bfloat162 m[4] = {ninf, ninf, ninf, ninf};
#pragma unroll
    for (int dh = -1; dh <= 1; ++dh) {
      const int h = 2 * ho + dh;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int dw = -1; dw <= 1; ++dw) {
        const int w = 2 * wo + dw;
        if (w < 0 || w >= W) continue;
        const uint4 t = __ldg(in + (((size_t)n * H + h) * W + w) * C8 + cg);
        const __nv_bfloat162* t2 = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], t2[k]);
      }
    }
    uint4 o;
    o.x = *reinterpret_cast<const uint32_t*>(&m[0]);
    o.y = *reinterpret_cast<const uint32_t*>(&m[1]);
    o.z = *reinterpret_cast<const uint32_t*>(&m[2]);
    o.w = *reinterpret_cast<const uint32_t*>(&m[3]);
    out[i] = o;
  }
}

// generic bridges (operator-level shims / tests): smem-tiled transpose of a (C x HW) slab per image
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int C, int HW,
                                             int Cpad) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && p < HW) ? in[((size_t)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    if (p < HW && c < Cpad) out[((size_t)n * HW + p) * Cpad + c] = __float2bfloat16_rn(tile[threadIdx.x][r]);
  }
}

__global__ void nhwc_bf16_to_nchw_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, int C, int HW,
                                             int Cpad) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < HW && c < C) ? __bfloat162float(in[((size_t)n * HW + p) * Cpad + c]) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (c < C && p < HW) out[((size_t)n * C + c) * HW + p] = tile[threadIdx.x][r];
  }
}


// HRNet fuse for the highest-resolution branch (no conv lands on it): out = relu(pre + sum_i upsample(up_i))
// (HRnet.py:254-263 with i == 0); thread = (pixel, 8-channel group)
__global__ void fuse_add_kernel(const FuseAddParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int C8 = p.C >> 3;
  const size_t total = (size_t)p.B * p.H * p.W * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    // 32-bit index arithmetic (the launcher guarantees total < 2^31): 64-bit div / mod cost ~100 instructions each and
    // made this streaming kernel ALU-bound
    const unsigned iu = (unsigned)i;
    const unsigned cg = iu % (unsigned)C8;
    const unsigned pix = iu / (unsigned)C8;
    const unsigned w = pix % (unsigned)p.W;
    const unsigned rowi = pix / (unsigned)p.W;
    const unsigned h = rowi % (unsigned)p.H;
    const unsigned n = rowi / (unsigned)p.H;
    float v[8];
    {
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(p.pre) + i);
      const uint32_t xs[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[2 * k] = bf16lo_to_f32(xs[k]);
        v[2 * k + 1] = bf16hi_to_f32(xs[k]);
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (p.up[a] == nullptr) continue;
      const int sh = p.up_shift[a];
      const size_t upix = ((size_t)n * (p.H >> sh) + (h >> sh)) * (p.W >> sh) + (w >> sh);
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(p.up[a]) + upix * C8 + cg);
      const uint32_t xs[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[2 * k] += bf16lo_to_f32(xs[k]);
        v[2 * k + 1] += bf16hi_to_f32(xs[k]);
      }
    }
    if (p.relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    reinterpret_cast<uint4*>(p.out)[i] = o;
  }
}

int launch_pack_input_s2d(const float* x, void* out, int B, int H, int W, cudaStream_t s, int out_pitch, int out_off) {
  if (out_pitch <= 0) out_pitch = W / 2;
  HRP_REQUIRE(H % 2 == 0 && W % 2 == 0, "input size must be even");
  const size_t total = (size_t)B * (H / 2) * (W / 2);
  const int threads = 256;
  const int blocks = (int)std::min<size_t>((total + threads - 1) / threads, 148 * 16);
  pack_input_s2d_kernel<<<blocks, threads, 0, s>>>(x, reinterpret_cast<uint4*>(out), B, H, W, out_pitch, out_off);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

int launch_pack_input_s2d_u8(const uint8_t* x, void* out, int B, int H, int W, cudaStream_t s, int out_pitch,
                             int out_off) {
  if (out_pitch <= 0) out_pitch = W / 2;
  HRP_REQUIRE(H % 2 == 0 && W % 2 == 0, "input size must be even");
  const size_t total = (size_t)B * (H / 2) * (W / 2);
  const int threads = 256;
  const int blocks = (int)std::min<size_t>((total + threads - 1) / threads, 148 * 16);
  if (W % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 7u) == 0 && total < (1ull << 31)) {
    const size_t total4 = total / 4;
    const int blocks4 = (int)std::min<size_t>((total4 + threads - 1) / threads, 148 * 32);
    pack_input_s2d_u8x4_kernel<<<blocks4, threads, 0, s>>>(x, reinterpret_cast<uint4*>(out), B, H, W, out_pitch, out_off);
  } else {
    pack_input_s2d_u8_kernel<<<blocks, threads, 0, s>>>(x, reinterpret_cast<uint4*>(out), B, H, W, out_pitch, out_off);
  }
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

int launch_maxpool3x3s2(const void* in, void* out, int B, int H, int W, int C, cudaStream_t s) {
  HRP_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool needs C%8==0 and even H,W");
  const size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 8);
  HRP_REQUIRE(total < (1ull << 31), "maxpool: tensor too large for 32-bit indexing");
  const int threads = 256;
  // (a strip variant -- four adjacent outputs per thread, 27 loads instead of 36 -- measured 428 us against 302 us: fewer
  //  threads in flight cost more than the saved L1 hits)
  const int blocks = (int)std::min<size_t>((total + threads - 1) / threads, 148 * 16);
  launch_ex(maxpool3x3s2_kernel, dim3(blocks), dim3(threads), 0, s, reinterpret_cast<const uint4*>(in),
            reinterpret_cast<uint4*>(out), B, H, W, C / 8);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

int launch_nchw_f32_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W, int Cpad, cudaStream_t s) {
  dim3 grid((H * W + 31) / 32, (Cpad + 31) / 32, B), block(32, 8);
  nchw_f32_to_nhwc_bf16_kernel<<<grid, block, 0, s>>>(in, reinterpret_cast<bf16*>(out), C, H * W, Cpad);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

int launch_nhwc_bf16_to_nchw_f32(const void* in, float* out, int B, int C, int H, int W, int Cpad, cudaStream_t s) {
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, B), block(32, 8);
  nhwc_bf16_to_nchw_f32_kernel<<<grid, block, 0, s>>>(reinterpret_cast<const bf16*>(in), out, C, H * W, Cpad);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

int launch_fuse_add(const FuseAddParams& p, cudaStream_t s) {
  HRP_REQUIRE(p.C % 8 == 0 && p.pre != nullptr && p.out != nullptr, "fuse_add: bad arguments");
  const size_t total = (size_t)p.B * p.H * p.W * (p.C / 8);
  HRP_REQUIRE(total < (1ull << 31), "fuse_add: tensor too large for 32-bit indexing");
  const int threads = 256;
  const int blocks = (int)std::min<size_t>((total + threads - 1) / threads, 148 * 16);
  launch_ex(fuse_add_kernel, dim3(blocks), dim3(threads), 0, s, p);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

}  // namespace hrp
