// Implicit-GEMM convolution on sm_100a: plan structures shared by the tcgen05 kernel (conv_gemm.cu), the
// SIMT cross-check kernel (conv_simt.cu) and the network executor (model.cu).
//
// One plan covers every conv-like layer of the reference's two backbones and deconv head
// (reference: lib/models/backbones/HRnet.py:22-98,197-242,341-429; lib/models/backbones/Resnet.py:21-29,96-135;
//  lib/models/full_net.py:194-216,78):
//   GEMM rows   = a grid (B, Hm, Wm) of "row pixels", tiled in boxes of bw x bh x bn = 128 pixels
//   GEMM K      = taps x Cin; tap t reads the source grid at (h + dh[t], w + dw[t]) through tensor map
//                 map[t] (maps differ only for stride-2 convs: one per input parity)
//   GEMM cols   = Cout (per phase; deconvs run 4 sub-pixel phases as gridDim.z)
//   epilogue    = v = acc*scale[c] + bias[c] + sum(pre[i]) + sum(upsampled up[i]);  relu;  v += post
//   output px   = (n, h*os + oh0 + ph, w*os + ow0 + pw) in a (B, Hout, Wout, Cout) bf16 NHWC tensor
#pragma once
#include "hrp_common.cuh"

namespace hrp {

constexpr int kMaxTaps = 16;
constexpr int kTileM = 128;

struct ConvParams {
  int B, Hm, Wm;          // row-pixel grid
  int bw, bh, bn;         // tile box (bw*bh*bn == 128)
  int tiles_w, tiles_h, tiles_n;
  int tw_shift, th_shift, bw_shift, bh_shift;  // log2 of tiles_w, tiles_h, bw, bh
  int Hs, Ws, Cin;        // source grid seen by the taps (per parity map for stride 2) + stored channels
  int Hout, Wout, Cout;   // output tensor
  int os, oh0, ow0;       // output pixel stride / offset
  int nphase;             // 1, or 4 for ConvTranspose2d(k4,s2,p1): phase z = (ph,pw) shifts taps and output
  int shared_phase;       // kConvUp2: the 4 phases share weight rows and taps (only the output pixel differs)
  int ntaps, ck, cpt;     // taps, channels per k-block (16/32/64), k-blocks per tap (Cin/ck)
  int n_tiles;            // cout_pad / n_tile
  int n_tile, cout_pad;   // N tile (multiple of 32, <= 256); weight rows per phase (multiple of n_tile)
  int cko;                // channels per TMA-store block of the epilogue (64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B)
  int ktot;               // ntaps*Cin
  int vsh;                // vertical tap sharing (3x3 s1 p1, tile inside one image): A buffer = bh+2 image rows per dw
  int vsh_a_bytes, vsh_a_pad, vsh_stage_bytes;
  int pair_off;           // > 0: element offset of the pixel-pair weight matrix behind the regular packing (conv_halo.cu)
  int8_t tap_dh[kMaxTaps], tap_dw[kMaxTaps], tap_map[kMaxTaps];
  // epilogue
  const float* scale;     // [Cout] folded BN scale (1 if none)
  const float* bias;      // [Cout] folded BN shift (+ conv bias)
  const bf16* pre[3];     // same-resolution addends, (B,Hout,Wout,Cout)
  const bf16* up[3];      // low-resolution addends, (B,Hout>>s,Wout>>s,Cout), nearest upsample by 2^s
  int up_shift[3];
  const bf16* post;       // added after the ReLU (HRNet cls head, HRnet.py:558-560)
  int relu;
  bf16* out;              // (B,Hout,Wout,Cout) bf16, may be null when only pooled output is wanted
  float* pool_out;        // optional (B,Cout) fp32: mean over the Hout*Wout pixels of post-activation values
  float pool_scale;       // 1/(Hout*Wout)
  // optional soft-argmax fold (final 1x1 conv of the keypoint head, full_net.py:78,296-297 + integral.py:109-135): instead
  // of writing the (B, 64*64, nkpt*64) logits, every epilogue warp reduces its 32 pixels x 32 depth bins of a keypoint to
  // an online-softmax partial (max, sum e, sum e*w, sum e*h, sum e*d in the log2 domain) and writes it to
  // head_partials[(image * head_chunks + chunk) * head_nkpt + keypoint][5]; head.cu merges the chunks per image.
  float* head_partials;
  int head_chunks, head_nkpt;
  // raw pointers for the SIMT cross-check path
  const bf16* in;         // (B,Hin,Win,Cin) source tensor base
  const bf16* w;          // (nphase*cout_pad, ktot) packed weights
  int Hin, Win, src_sh, src_sw;  // full input dims and source-grid stride (2 for parity maps)
  // element strides of the source grid seen by the A tensor maps (and the SIMT cross-check): position (n, h, w) of
  // map (hp, wp) starts at in + src_off + (hp*Win + wp)*Cin + n*src_img + h*src_row + w*src_pix
  long long src_img;
  int src_row, src_pix, src_off;
  long long* timeline;    // debug: per-role clock64 stamps of the first CTAs (null in production)
  int dbg;                // debug ablation mask (HRP_CONV_DBG, experiments only; 0 in production)
};

struct ConvMaps {
  CUtensorMap a[4];  // input: one per stride-2 parity (else a[0])
  CUtensorMap b;     // packed weights
  CUtensorMap o[4];  // output: one per deconv phase (else o[0])
  CUtensorMap av;    // input with a (bh+2)-row box for vertical tap sharing
  CUtensorMap r;     // residual (pre[0]) tile, same geometry as the output (persistent kernel)
  // staged addends of the persistent kernel (PersistCfg::staged): every addend of the epilogue is a TMA box
  CUtensorMap rp[3];   // pre[0] for output phases 1..3 of an output-strided layer (phase 0 uses r)
  CUtensorMap add[3];  // pre[1], pre[2], post: same geometry as the output
  CUtensorMap upm[3];  // up[a]: low-resolution tensor, box = the tile's footprint at that resolution
};

struct PersistCfg {
  int stages;        // smem pipeline depth
  int nstag;         // output staging buffers (1 or 2)
  int stag_offset;   // byte offset of the staging buffers
  int bar_offset;    // byte offset of the barriers (scale/shift follow them)
  int total_tiles;   // m tiles x n tiles x phases
  int tmem_cols;     // allocated TMEM columns (power of two >= nacc * ksplit * n_tile)
  int nacc;          // accumulator tiles in the TMEM ring (2 or 4)
  int tw_shift, th_shift;  // log2 of tiles_w / tiles_h (tile decode without divisions)
  int vsh;           // vertical tap sharing: one (bh+2)-row A buffer per (channel chunk, dw)
  int wres;          // packed weights of the (single) N tile resident in shared memory
  int a_bytes, a_region, stage_bytes, pipe_offset;
  int ksplit;        // accumulators per tile: K steps round-robin over them to hide the dependent-MMA latency
  int sub;           // k-blocks (ck channels of one tap) per pipeline stage
  int nprod;         // TMA producer threads (1 or 2, in two warps): stage g is issued by producer g % nprod
  int dual;          // two independent (producer, half ring, MMA issuer) pipelines, tiles alternate between them
  int res_store;     // the TMA-store warp fetches the residual / addend tiles (else a producer does)
  // Staged addends: the generic epilogue flavours (PRE / FULL) gather their addends with 16-byte loads at pixel stride --
  // ~60 issue slots and 32 L1 sectors per 8 channels and addend.  With `staged` every addend tile is fetched by TMA into
  // the staging-ring entry of its output tile (same swizzled box layout as the output / residual slot, so the epilogue
  // reads it with conflict-free 16-byte shared loads at the SAME offsets); nearest-upsampled addends are fetched at
  // their own resolution (a few rows) and indexed by (h >> s, w >> s).
  int res_tma;       // pre[0] is TMA-fetched into the output slot (updated in place)
  int staged;        // ... and so is every other addend
  int entry_bytes;   // bytes per ring entry: output / pre[0] slot + addend slots
  int add_off[3];    // byte offset in an entry of the pre[1], pre[2], post slots (0 = absent)
  int up_off[3];     // byte offset of the up[a] slot (0 = absent)
  int up_blk[3];     // bytes per channel block of an up[a] slot (1024-aligned)
  int up_bw[3], up_bh[3], up_sh[3];  // low-resolution box (pixels) and shift from the tile's row-pixel grid
};

// Halo-tile 3x3 kernel (conv_halo.cu): a unit = T consecutive 128-position tiles of one zero-padded image
struct HaloParams {
  int B, H, W, Cout, relu;
  int pair;              // pixel-pair view of a 32-channel layer: W, Wp, Cout describe the (W/2, 64-channel) problem
  int Wp;                // W + 2: padded row pitch in positions
  int NR;                // input rows per TMA box (band + halo)
  int T;                 // 128-position tiles per unit
  int tiles_per_img, units_per_img, total_units;
  int n_abuf;            // band buffers in flight
  int nacc, nacc_shift;  // accumulators in the TMEM ring (power of two)
  int nring;             // output / residual ring buffers (2..4), one 128-position tile each
  int a_buf_bytes, a_offset, ring_offset, bar_offset;
  uint32_t div_magic;    // P / Wp == (P * div_magic) >> 20
  const float* scale;
  const float* bias;
  const bf16* res;       // residual addend (pre[0]) or null
  bf16* out;
  long long* tl;         // debug timeline (globaltimer stamps of the first 8 CTAs), null in production
  int dbg;               // ablation (HRP_HALO_DBG=1): epilogue reduced to the accumulator read
};

struct ConvPlan {
  ConvParams p;
  ConvMaps maps;
  dim3 grid;
  int smem_bytes;
  int bar_offset;  // barriers + scale/shift staging live after the pipeline stages
  int stages;
  int epi;         // epilogue flavour (EPI_PLAIN / EPI_PRE / EPI_FULL)
  bool persistent; // persistent kernel (default) vs one-tile-per-CTA kernel (HRP_CONV_V1=1)
  bool halo_ok;    // the halo-tile kernel can run this layer
  bool halo;       // ... and has been selected (heuristic / autotune / HRP_CONV_VARIANT)
  HaloParams hp;
  CUtensorMap halo_map_a;
  CUtensorMap halo_map_r;  // residual as a 2-D (pixels, C) tensor, swizzled 128-row boxes
  CUtensorMap halo_map_b;  // pixel-pair weight matrix (pair mode only)
  unsigned halo_grid;
  int halo_smem;
  PersistCfg pcfg;
  unsigned pgrid;
  int psmem;
  double flops;  // 2*MACs of the reference layer (algorithmic, not padded)
};

// Layer description used to build a plan.
// kConvUp2: 1x1 conv whose output is nearest-upsampled by 2 -- every row pixel feeds its 2x2 output pixels through four
// phases that share ONE weight matrix and the unshifted tap; the epilogue's addends are indexed at the OUTPUT resolution.
// This is how the branch-0 sum of an HRNet fuse layer (HRnet.py:254-263: y0 = relu(x0 + up2(bn(conv1x1(x1))) + ...)) runs
// inside a conv epilogue instead of a separate elementwise kernel.
enum ConvKind { kConv = 0, kDeconvK4S2P1 = 1, kStemS2D = 2, kConvUp2 = 3 };

struct ConvLayerDesc {
  int kind;              // ConvKind
  int B, Hin, Win, Cin;  // input NHWC (Cin = stored channels, multiple of 16). For kStemS2D the input is the
                         // space-to-depth tensor (B, H/2, W/2, 16) and kh/kw/pad describe the ORIGINAL conv.
  int Cout;
  int kh, kw, stride, pad;
  int relu;
  int has_residual;      // a same-resolution addend (pre[0]) will be attached: keeps the N tile <= 128
  int head_fold;         // the epilogue reduces the logits to soft-argmax partials (ConvParams::head_partials): N tile 128
  // kStemS2D only: the s2d input has padded rows (in_wpitch pixels per row, the image starts at column in_wpad, the
  // padding is zero).  The horizontal taps of one kernel row are then ONE contiguous K block of nw*16 channels
  // (overlapping TMA windows, pixel stride 32 B): 4x fewer, 4x longer TMA rows than one box per tap.  0 = dense.
  int in_wpitch, in_wpad;
};

// Host-side weight packing: reference layouts -> (nphase*cout_pad, ktot) bf16 K-major.
//   kConv:          w is (Cout, Cin_ref, kh, kw) fp32 (torch Conv2d)
//   kDeconvK4S2P1:  w is (Cin_ref, Cout, 4, 4) fp32 (torch ConvTranspose2d)
//   kStemS2D:       w is (Cout, 3, kh, kw) fp32, conv stride 2; packed for the 2x2 space-to-depth input
int conv_geometry(const ConvLayerDesc& d, ConvParams* p);  // fills geometry fields (no pointers)
size_t conv_packed_weight_elems(const ConvParams& p);
int conv_pack_weights(const ConvLayerDesc& d, const ConvParams& p, int cin_ref, const float* w, uint16_t* out);

// Build tensor maps + launch config.  `in`/`w_packed` are device pointers; epilogue pointers are taken from p.
void conv_init();  // one-time kernel attribute setup (call outside stream capture)
void conv_halo_init();
int conv_encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int ck);  // bf16 tiled map, swizzle = ck*2 bytes
int conv_halo_plan(ConvPlan* plan, const ConvLayerDesc& d, const bf16* in);  // eligibility + tiling of the halo kernel
int conv_halo_launch(const ConvPlan& plan, cudaStream_t stream);
int conv_plan_finalize(ConvPlan* plan, const bf16* in, const bf16* w_packed, const ConvLayerDesc* desc = nullptr);
int conv_plan_launch(const ConvPlan& plan, cudaStream_t stream);       // tcgen05 path
int conv_plan_launch_simt(const ConvPlan& plan, cudaStream_t stream);  // cross-check path (tests only)

}  // namespace hrp
