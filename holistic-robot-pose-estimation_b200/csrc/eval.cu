// Callers either side of the network, memory- / latency-bound kernels:
//   f1  crop_resize_kernel   frame + bbox -> square crop -> bilinear 256x256 uint8 NCHW + intrinsics + k_value
//                            (lib/dataset/roboutils.py:128-157, augmentations.py:165-233, geometries.py:360-402,
//                             dream.py:297-310, scripts/test.py:141-152)
//   f2  metrics_*_kernel     per-batch ADD / 2-D / joint / depth errors and the ADD / PCK AUC summary
//                            (lib/utils/metrics.py:8-162)
// Byte work (the crop) is bit-exact with the reference: every fp32 operation is written with explicit
// round-to-nearest intrinsics in the operation order torch's CPU kernels use, so nvcc cannot contract differently.
#include "eval.h"
#include "launch_count.h"

namespace hrp {

// ------------------------------------------------------------------------------------------------------
// f1
// ------------------------------------------------------------------------------------------------------
struct BilinearTap {
  int i0, i1;
  float l0, l1;
};

// torch's area_pixel_compute_source_index / guard_index_and_lambda (align_corners = False), fp32
__device__ __forceinline__ BilinearTap bilinear_tap(int dst, float scale, int in_size) {
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  src = fmaxf(src, 0.0f);
  BilinearTap t;
  t.i0 = min((int)floorf(src), in_size - 1);
  t.l1 = fminf(fmaxf(__fsub_rn(src, (float)t.i0), 0.0f), 1.0f);
  t.l0 = __fsub_rn(1.0f, t.l1);
  t.i1 = t.i0 + ((t.i0 < in_size - 1) ? 1 : 0);
  return t;
}

struct CropGeom {
  int wmin, hmin, bw, bh, side, xo, yo;
};

__device__ __forceinline__ CropGeom crop_geom(const int* bb) {
  CropGeom g;
  g.wmin = bb[0];
  g.hmin = bb[1];
  g.bw = bb[2] - bb[0];
  g.bh = bb[3] - bb[1];
  g.side = max(g.bw, g.bh);
  g.xo = (g.side - g.bw) / 2;
  g.yo = (g.side - g.bh) / 2;
  return g;
}

// one pixel of the zero-padded square crop, as uint8 / 255 in fp32 (true division, like `im.float() / 255`)
__device__ __forceinline__ float square_px(const uint8_t* frame, int frame_w, const CropGeom& g, int sy, int sx, int c) {
  const int y = sy - g.yo, x = sx - g.xo;
  if (y < 0 || y >= g.bh || x < 0 || x >= g.bw) return 0.0f;
  const uint8_t v = __ldg(frame + ((size_t)(g.hmin + y) * frame_w + (g.wmin + x)) * 3 + c);
  return __fdiv_rn((float)v, 255.0f);
}

// grid (out/4 column groups * out/rows_per_block, B); each thread produces 4 consecutive pixels of one row, all 3
// channels, and writes one uchar4 per colour plane (coalesced 128-byte rows per warp and plane).
__global__ void __launch_bounds__(256) crop_resize_kernel(CropParams p) {
  const int b = blockIdx.y;
  const int* bb = p.bbox + 4 * b;
  const CropGeom g = crop_geom(bb);
  const uint8_t* frame = p.frames + (size_t)b * p.frame_h * p.frame_w * 3;
  const int out = p.out_size;
  const int groups = out >> 2;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;

  if (t == 0) {
    // intrinsics: fp64 principal-point shift (roboutils.py:150-152), then get_K_crop_resize in fp32 torch-op order
    const double* Ki = p.K_in + 9 * b;
    float* Ko = p.K_out + 9 * b;
    float K[9];
    for (int i = 0; i < 9; ++i) K[i] = (float)Ki[i];
    K[2] = (float)(Ki[2] - (double)(g.wmin - g.xo));
    K[5] = (float)(Ki[5] - (double)(g.hmin - g.yo));
    if (g.side != out) {
      const float F = (float)out;
      // box = (x0 - w/2, y0 - h/2, x0 + w/2, y0 + h/2) with x0 = y0 = side / 2 (python floats -> fp32 tensor)
      const float b0 = (float)((double)g.side / 2 - (double)g.side / 2), b2 = (float)((double)g.side / 2 + (double)g.side / 2);
      const float cw = __fsub_rn(b2, b0);
      const float cj = __fdiv_rn(__fadd_rn(b0, b2), 2.0f);
      const float half = __fdiv_rn(__fsub_rn(cw, 1.0f), 2.0f);
      const float cx = __fsub_rn(__fadd_rn(K[2], half), cj);
      const float cy = __fsub_rn(__fadd_rn(K[5], half), cj);
      const float dx = __fsub_rn(cx, half), dy = __fsub_rn(cy, half);
      const float sc = __fdiv_rn(F, cw);
      const float fc = __fdiv_rn(__fsub_rn(F, 1.0f), 2.0f);
      K[0] = __fmul_rn(sc, K[0]);
      K[4] = __fmul_rn(sc, K[4]);
      K[2] = __fadd_rn(fc, __fmul_rn(sc, dx));
      K[5] = __fadd_rn(fc, __fmul_rn(sc, dy));
    }
    for (int i = 0; i < 9; ++i) Ko[i] = K[i];
    if (p.k_value != nullptr) {
      // scripts/test.py:141-152: sqrt(fx * fy * 1000 * 1000 / max(|x2 - x1|, |y2 - y1|)^2), fp32 left to right
      const float* kb = p.k_bbox + 4 * b;
      const float* Kk = (p.k_use_crop_K != 0) ? K : nullptr;
      const float fx = Kk ? Kk[0] : (float)Ki[0], fy = Kk ? Kk[4] : (float)Ki[4];
      const float e = fmaxf(fabsf(__fsub_rn(kb[2], kb[0])), fabsf(__fsub_rn(kb[3], kb[1])));
      const float area = __fmul_rn(e, e);
      const float num = __fmul_rn(__fmul_rn(__fmul_rn(fx, fy), 1000.0f), 1000.0f);
      p.k_value[b] = __fsqrt_rn(__fdiv_rn(num, area));
    }
  }
  if (t >= groups * out) return;
  const int oy = t / groups, ox0 = (t - oy * groups) << 2;
  uint8_t* dst = p.out_u8 + (size_t)b * 3 * out * out + (size_t)oy * out + ox0;
  uchar4 px[3];
  uint8_t* pb = reinterpret_cast<uint8_t*>(px);

  if (g.side == out) {  // augmentations.py:174-176: already the target size, the crop passes through untouched
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int y = oy - g.yo, x = ox0 + j - g.xo;
        uint8_t v = 0;
        if (y >= 0 && y < g.bh && x >= 0 && x < g.bw)
          v = __ldg(frame + ((size_t)(g.hmin + y) * p.frame_w + (g.wmin + x)) * 3 + c);
        pb[c * 4 + j] = v;
      }
  } else {
    const float scale = __fdiv_rn((float)g.side, (float)out);
    const BilinearTap th = bilinear_tap(oy, scale, g.side);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const BilinearTap tw = bilinear_tap(ox0 + j, scale, g.side);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v00 = square_px(frame, p.frame_w, g, th.i0, tw.i0, c);
        const float v01 = square_px(frame, p.frame_w, g, th.i0, tw.i1, c);
        const float v10 = square_px(frame, p.frame_w, g, th.i1, tw.i0, c);
        const float v11 = square_px(frame, p.frame_w, g, th.i1, tw.i1, c);
        // torch CPU (UpSampleKernel.cpp, Interpolate<2>::eval as compiled): fma(fma(v00,w0,v01*w1), h0, fma(v10,w0,v11*w1)*h1)
        const float r0 = __fmaf_rn(v00, tw.l0, __fmul_rn(v01, tw.l1));
        const float r1 = __fmaf_rn(v10, tw.l0, __fmul_rn(v11, tw.l1));
        const float o = __fmaf_rn(r0, th.l0, __fmul_rn(r1, th.l1));
        pb[c * 4 + j] = (uint8_t)(int)__fmul_rn(o, 255.0f);  // `(x * 255).to(torch.uint8)`: truncation
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) *reinterpret_cast<uchar4*>(dst + (size_t)c * out * out) = px[c];
}

int launch_crop_resize(const CropParams& p, cudaStream_t s) {
  const int work = (p.out_size / 4) * p.out_size;
  dim3 grid((work + 255) / 256, p.B);
  crop_resize_kernel<<<grid, 256, 0, s>>>(p);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// f2: per-batch metrics
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {  // fixed butterfly order -> deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per image, lane = keypoint (nkpt <= 32) / joint (dof <= 32)
__global__ void __launch_bounds__(128) metrics_image_kernel(MetricsParams p) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= p.B) return;
  const int N = p.nkpt, B = p.B;
  const bool on = lane < N;
  const int k = on ? lane : 0;
  const float* P = p.pred_kp3d + ((size_t)b * N + k) * 3;
  const float* G = p.gt_kp3d + ((size_t)b * N + k) * 3;
  const float* G2 = p.gt_kp2d + ((size_t)b * N + k) * 2;
  const float* K = p.K + (size_t)b * 9;
  const float px = P[0], py = P[1], pz = P[2], gx = G[0], gy = G[1], gz = G[2];
  // ADD: metrics.py:54-57
  const float dx = px - gx, dy = py - gy, dz = pz - gz;
  const float e3 = on ? sqrtf(dx * dx + dy * dy + dz * dz) : 0.0f;
  // projection with K_original (transforms.py:7-15) and the in-frame mask of metrics.py:60-66
  const float qx = K[0] * px + K[1] * py + K[2] * pz;
  const float qy = K[3] * px + K[4] * py + K[5] * pz;
  const float qz = K[6] * px + K[7] * py + K[8] * pz;
  const float u = qx / qz, v = qy / qz;
  const float g2x = G2[0], g2y = G2[1];
  const bool ok = on && (g2x <= p.frame_w) && (g2x >= 0.0f) && (g2y <= p.frame_h) && (g2y >= 0.0f);
  const float du = u - g2x, dv = v - g2y;
  const float e2 = on ? sqrtf(du * du + dv * dv) : 0.0f;
  const float e2v = ok ? e2 : 0.0f;  // `error2d_batch * valid` (a NaN error of an invalid keypoint is not reproduced)
  // depth / root-relative: metrics.py:94-112
  const float pzr = __shfl_sync(0xffffffffu, pz, p.ref_kpt), gzr = __shfl_sync(0xffffffffu, gz, p.ref_kpt);
  const float rel = (pz - pzr) - (gz - gzr);
  const float e3r = on ? sqrtf(dx * dx + dy * dy + rel * rel) : 0.0f;
  const float s3 = warp_sum(e3), s2 = warp_sum(e2v), sv = warp_sum(ok ? 1.0f : 0.0f);
  const float srel = warp_sum(on ? fabsf(rel) : 0.0f), s3r = warp_sum(e3r);
  if (on) {
    p.kp_err3d[(size_t)k * B + b] = e3;  // column-major scratch for the per-keypoint batch means
    p.kp_err2d[(size_t)k * B + b] = e2v;
    p.kp_valid[(size_t)k * B + b] = ok ? 1.0f : 0.0f;
  }
  float sj = 0.0f;
  if (p.pred_joint != nullptr) {
    const bool jon = lane < p.dof;
    const float ej = jon ? fabsf(p.gt_joint[(size_t)b * p.dof + lane] - p.pred_joint[(size_t)b * p.dof + lane]) : 0.0f;
    if (jon) p.joint_err[(size_t)lane * B + b] = ej;
    const int nj = p.drop_last_joint ? p.dof - 1 : p.dof;  // metrics.py:83-86 (Panda: finger joint excluded)
    sj = warp_sum(lane < nj ? ej : 0.0f) / (float)nj;
  }
  if (lane == 0) {
    float* o = p.per_image;
    o[0 * B + b] = s3 / (float)N;
    o[1 * B + b] = s2 / sv;
    o[2 * B + b] = sj;
    o[3 * B + b] = fabsf(pzr - gzr);
    o[4 * B + b] = srel / (float)N;
    o[5 * B + b] = s3r / (float)N;
  }
}

// block c reduces column c of the (cols, B) scratch over the batch with a fixed-order tree
__global__ void __launch_bounds__(256) metrics_column_kernel(MetricsParams p) {
  __shared__ float sh[2][256];
  const int c = blockIdx.x, N = p.nkpt, B = p.B;
  const float* num;
  const float* den = nullptr;
  float* out;
  if (c < N) {
    num = p.kp_err3d + (size_t)c * B;
    out = p.dis3d + c;
  } else if (c < 2 * N) {
    num = p.kp_err2d + (size_t)(c - N) * B;
    den = p.kp_valid + (size_t)(c - N) * B;
    out = p.dis2d + (c - N);
  } else {
    num = p.joint_err + (size_t)(c - 2 * N) * B;
    out = p.l1_jointerror + (c - 2 * N);
  }
  float a = 0.0f, d = 0.0f;
  for (int i = threadIdx.x; i < B; i += 256) {
    a += num[i];
    if (den != nullptr) d += den[i];
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = d;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0][0] / (den != nullptr ? sh[1][0] : (float)B);
}

int launch_metrics_batch(const MetricsParams& p, cudaStream_t s) {
  metrics_image_kernel<<<(p.B + 3) / 4, 128, 0, s>>>(p);
  const int cols = 2 * p.nkpt + (p.pred_joint != nullptr ? p.dof : 0);
  metrics_column_kernel<<<cols, 256, 0, s>>>(p);
  count_launch(2);
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// f2: summary (metrics.py:122-162).  Block 0 = ADD (3-D errors, metres), block 1 = PCK (2-D errors, pixels).
// ------------------------------------------------------------------------------------------------------
constexpr int kSumThreads = 1024;

__device__ double block_sum_f64(double v, double* sh) {
  __syncthreads();
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = kSumThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// k-th smallest (0-based) of n non-negative floats: 4-pass MSB radix select on the bit patterns
__device__ float block_select(const float* d, long long n, long long k, unsigned* hist, unsigned* sh_sel) {
  unsigned prefix = 0, mask = 0;
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = threadIdx.x; i < 256; i += kSumThreads) hist[i] = 0;
    __syncthreads();
    for (long long i = threadIdx.x; i < n; i += kSumThreads) {
      const unsigned key = __float_as_uint(d[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long kk = k;
      unsigned digit = 0;
      for (; digit < 256; ++digit) {
        if (kk < (long long)hist[digit]) break;
        kk -= hist[digit];
      }
      sh_sel[0] = digit;
      sh_sel[1] = (unsigned)kk;
    }
    __syncthreads();
    prefix |= sh_sel[0] << (8 * pass);
    mask |= 255u << (8 * pass);
    k = (long long)sh_sel[1];
    __syncthreads();
  }
  return __uint_as_float(prefix);
}

__global__ void __launch_bounds__(kSumThreads) metrics_summary_kernel(SummaryParams p) {
  extern __shared__ unsigned char sm_raw[];
  double* shd = reinterpret_cast<double*>(sm_raw);                            // [1024]
  unsigned* hist = reinterpret_cast<unsigned*>(shd + kSumThreads);            // [nthr + 1]
  const bool pck = (blockIdx.x == 1);
  const float* d = pck ? p.dis2d : p.dis3d;
  const long long n = p.n;
  const int nthr = pck ? p.nthr_pck : p.nthr_add;
  const double delta = pck ? 0.01 : 0.00001, limit = pck ? 20.0 : 0.1;
  unsigned* sel_hist = hist + nthr + 1;                                      // [256] + [2]
  double* out = p.out + (pck ? 11 : 0);  // [mean, median, AUC, 8 threshold fractions]

  // ---- AUC: bin index = first threshold i * delta (fp64) that is >= d (numpy compares fp32 errors with fp64 thresholds)
  for (int i = threadIdx.x; i <= nthr; i += kSumThreads) hist[i] = 0;
  __syncthreads();
  double acc = 0.0;
  unsigned tab[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned n_nan = 0;  // an image without in-frame ground-truth keypoints has error2d = 0/0 (metrics.py:67-70)
  for (long long i = threadIdx.x; i < n; i += kSumThreads) {
    const float df = d[i];
    const double x = (double)df;
    n_nan += (df != df) ? 1u : 0u;
    acc += x;
    long long j = nthr;  // above every threshold (or NaN)
    if (x <= (double)(nthr - 1) * delta) {
      j = (long long)ceil(x / delta);
      if (j < 0) j = 0;
      if (j > nthr - 1) j = nthr - 1;
      while (j > 0 && x <= (double)(j - 1) * delta) --j;
      while (!(x <= (double)j * delta)) ++j;
    }
    atomicAdd(&hist[j], 1u);
#pragma unroll
    for (int t = 0; t < 8; ++t) tab[t] += (df <= p.table_thr[(pck ? 8 : 0) + t]) ? 1u : 0u;  // fp32 compare (weak python scalar)
  }
  const double total = block_sum_f64(acc, shd);
  const bool any_nan = block_sum_f64((double)n_nan, shd) > 0.0;
  for (int t = 0; t < 8; ++t) {
    const double c = block_sum_f64((double)tab[t], shd);
    if (threadIdx.x == 0) out[3 + t] = c / (double)n;
  }
  __syncthreads();
  // inclusive prefix over the bins (single thread per 1024-chunk would do; n_thr <= 10001: serial scan by thread 0)
  if (threadIdx.x == 0) {
    unsigned run = 0;
    for (int i = 0; i <= nthr; ++i) {
      run += hist[i];
      hist[i] = run;
    }
  }
  __syncthreads();
  // trapz(y, dx=delta) = sum(delta * (y[i+1] + y[i]) / 2), y[i] = count(d <= thr_i) / n
  double part = 0.0;
  for (int i = threadIdx.x; i + 1 < nthr; i += kSumThreads) {
    const double y0 = (double)hist[i] / (double)n, y1 = (double)hist[i + 1] / (double)n;
    part += delta * (y1 + y0) / 2.0;
  }
  const double area = block_sum_f64(part, shd);
  // ---- median (np.median: mean of the two middle order statistics for even n)
  const float m_hi = block_select(d, n, n / 2, sel_hist, sel_hist + 256);
  float med = m_hi;
  if ((n & 1) == 0) {
    const float m_lo = block_select(d, n, n / 2 - 1, sel_hist, sel_hist + 256);
    med = (m_lo + m_hi) * 0.5f;  // exact unless m_lo + m_hi overflows
  }
  // np.median propagates NaN (and np.mean does through `total`); the radix select orders bit patterns of non-negative
  // finite floats only
  if (any_nan) med = __int_as_float(0x7fc00000);
  if (threadIdx.x == 0) {
    out[0] = total / (double)n;
    out[1] = (double)med;
    out[2] = area / limit;
  }
}

int launch_metrics_summary(const SummaryParams& p, cudaStream_t s) {
  const int nthr = p.nthr_add > p.nthr_pck ? p.nthr_add : p.nthr_pck;
  const size_t smem = kSumThreads * sizeof(double) + (size_t)(nthr + 1 + 256 + 2) * sizeof(unsigned);
  // (idempotent and cheap; set on every launch so that every device of a multi-device process is covered)
  cudaFuncSetAttribute(metrics_summary_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (smem > 96 * 1024) {
    set_error("metrics summary: too many thresholds");
    return HRP_ERR_INVALID;
  }
  metrics_summary_kernel<<<2, kSumThreads, smem, s>>>(p);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

}  // namespace hrp
