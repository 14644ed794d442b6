// Whole-network executor: RootNetwithRegInt (ResNet-50 + deconv head + HRNet-w32 root-depth net) and the
// standalone RootNet depthnet, as a static program of planned kernels per batch-chunk size, captured once into
// a CUDA graph with one capture stream per independent lane (4 HRNet branches + the ResNet/deconv path).
//
// Reference topology restated from lib/models/full_net.py:239-397, lib/models/backbones/Resnet.py:56-135,
// lib/models/backbones/HRnet.py:101-265,341-429,499-570 (+ configs/hrnet_w32.yaml:54-93) and
// lib/models/depth_net.py:92-137.  Weights arrive by reference state-dict key (hrp_model_set_tensor).
#include "../../include/hrp.h"

#include "conv.h"
#include "head.h"
#include "launch_count.h"
#include "ops.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

namespace hrp {

namespace {

// The s2d-packed inputs of the two stems carry zero columns on both sides of every row (see ConvLayerDesc::in_wpitch):
constexpr int kS2dPad = 2;                        // left padding, pixels (7x7 stem: taps reach 2 s2d pixels left)
constexpr int kS2dPitch = 256 / 2 + 2 * kS2dPad;  // pixels per padded row
constexpr int kNumLanes = 5;  // lanes 0..3: HRNet branches (lane 0 also stem / layer1 / cls head); lane 4: ResNet path
constexpr int kLaneRN = 4;
constexpr int kImg = 256;

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

struct ConvWeights {
  bf16* w = nullptr;
  float* scale = nullptr;
  float* bias = nullptr;
};

struct Act {
  int H = 0, W = 0, C = 0;
  bf16* ptr = nullptr;
  int producer = -1;  // op index
};

struct Epi {
  int pre[3] = {-1, -1, -1};
  int up[3] = {-1, -1, -1};
  int up_shift[3] = {0, 0, 0};
  int post = -1;
};

enum OpKind { OP_CONV = 0, OP_MAXPOOL = 1, OP_FUSEADD = 2, OP_MEMSET = 3 };

struct Op {
  int kind = OP_CONV;
  int lane = 0;
  std::string name;
  ConvPlan conv;
  FuseAddParams fuse{};
  const void* mp_in = nullptr;
  void* mp_out = nullptr;
  int mp_B = 0, mp_H = 0, mp_W = 0, mp_C = 0;
  void* ms_ptr = nullptr;
  size_t ms_bytes = 0;
  std::vector<int> reads;  // activation ids
  int writes = -1;
  bool cross_lane_consumer = false;
  cudaEvent_t done = nullptr;
};

}  // namespace

struct Plan {
  int B = 0;
  std::vector<Act> acts;
  std::vector<Op> ops;
  std::map<std::string, int> taps;
  std::vector<void*> owned;
  bf16* s2d_reg = nullptr;
  bf16* s2d_root = nullptr;
  float* feat = nullptr;     // (B,2048) pooled HRNet feature
  float* xf = nullptr;       // (B,2048) pooled ResNet feature
  int heatmap = -1;          // activation id (-1 when the soft-argmax is folded into the final conv's epilogue)
  void* head_ws = nullptr;
  int head_chunks = 0;
  bool single_lane = false;        // the whole program is captured on one stream (large chunks)
  size_t arena_bytes = 0;          // activation arena (liveness-aliased when single_lane), else the sum of all tensors
  size_t act_bytes_sum = 0;        // what one-buffer-per-tensor would need
  float* fold_partials = nullptr;  // (B, fold_chunks, nkpt, 5) written by the final conv's epilogue (EPI_HEAD)
  float* fold_merged = nullptr;    // (B, nkpt, 5)
  int fold_chunks = 0;
  cudaStream_t stream = nullptr;         // launch stream of this replica (when inflight > 1)
  cudaStream_t cap[kNumLanes] = {};
  cudaEvent_t fork_ev = nullptr, join_ev[kNumLanes] = {};
  cudaEvent_t done_ev = nullptr;
  // every forward that used this plan's buffers records busy_ev on its stream when it has enqueued its last kernel; the
  // next user -- possibly on another stream -- waits on it first, so two forwards never overlap on one set of
  // activation buffers / pooled accumulators / head counters
  cudaEvent_t busy_ev = nullptr;
  bool busy_recorded = false;
  uint64_t last_use = 0;  // LRU stamp (hrp_model::use_counter)
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int n_kernels = 0;
  double flops = 0;

  ~Plan() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    for (auto& op : ops)
      if (op.done) cudaEventDestroy(op.done);
    for (int l = 0; l < kNumLanes; ++l) {
      if (cap[l]) cudaStreamDestroy(cap[l]);
      if (join_ev[l]) cudaEventDestroy(join_ev[l]);
    }
    if (fork_ev) cudaEventDestroy(fork_ev);
    if (done_ev) cudaEventDestroy(done_ev);
    if (busy_ev) {
      if (busy_recorded) cudaEventSynchronize(busy_ev);  // eviction / destruction: the last user must have finished
      cudaEventDestroy(busy_ev);
    }
    if (stream) cudaStreamDestroy(stream);
    for (void* p : owned) cudaFree(p);
  }
};

}  // namespace hrp

using namespace hrp;

struct hrp_robot;  // defined in api.cu
extern "C" const void* hrp_robot_device_table(const hrp_robot* robot);
extern "C" int hrp_robot_dims(const hrp_robot* robot, int32_t* nkpt, int32_t* dof);

struct hrp_model {
  hrp_model_desc desc;
  bool finalized = false;
  std::map<std::string, HostTensor> host;
  std::map<std::string, ConvWeights> wcache;
  // conv shape signature -> kernel variant (0 one tile per CTA, 1 persistent, 2 halo).  Filled from the committed tuning
  // table (hrp_model_set_tuning: the default, deterministic) and / or by the timed autotune (HRP_AUTOTUNE=1).
  std::map<std::string, int> tune_cache;
  bool autotune = false;
  bool pin_variants = false;                   // HRP_CONV_PERSISTENT / HRP_CONV_VARIANT pin the kernel: no table, no tuning
  std::mutex mu;                               // guards plans / tune_cache / stream_slot and serialises enqueueing
  std::map<cudaStream_t, int> stream_slot;     // caller stream -> plan replica (round-robin over desc.inflight)
  int next_slot = 0;
  uint64_t use_counter = 0;
  int max_plans = 8;                           // LRU bound on cached plans (HRP_MAX_PLANS)
  std::vector<void*> owned;
  const hrp_robot* robot = nullptr;
  RegressorTable* reg_pose = nullptr;
  RegressorTable* reg_rot = nullptr;
  float* depth_w = nullptr;
  float depth_b = 0.f;
  float* init_pose = nullptr;
  float* init_rot = nullptr;
  std::map<std::pair<int, int>, std::unique_ptr<Plan>> plans;  // (batch, replica)
  Plan* last_plan = nullptr;
  cudaEvent_t fork_ev = nullptr;
  int device = 0;
  bool use_graph = true;
  bool use_simt = false;
  // HRNet fuse, branch 0: the sum can run in an upsampling conv's epilogue (HRP_FUSE0_EPI=1; tests cover it) instead of
  // the elementwise fuse_add kernel.  Measured on B200 at 512 images it is SLOWER (160-220 us per module against 57-78 us
  // for fuse_add + 27 us for the plain 1x1 conv): the N = 32 tile leaves 4 epilogue warps per SM to gather / add 3 addend
  // streams, latency-bound (profiles/r02_exp_fuse0_epilogue.txt) -- so the elementwise kernel stays the default.
  bool fuse0_epilogue = false;
  bool head_fold = true;   // soft-argmax partials computed by the final conv's epilogue: the heatmap never reaches HBM
  std::string prefix_root;  // "rootnet_backbone." (full) or "backbone." (depthnet)

  ~hrp_model() {
    plans.clear();
    for (void* p : owned) cudaFree(p);
    if (fork_ev) cudaEventDestroy(fork_ev);
  }
};

namespace {

// ------------------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------------------
int dev_upload(hrp_model* m, const void* src, size_t bytes, void** out) {
  void* p = nullptr;
  HRP_CUDA_CHECK(cudaMalloc(&p, bytes));
  m->owned.push_back(p);
  HRP_CUDA_CHECK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
  *out = p;
  return HRP_OK;
}

const HostTensor* find(const hrp_model* m, const std::string& key) {
  auto it = m->host.find(key);
  return it == m->host.end() ? nullptr : &it->second;
}

#define HRP_NEED(ptr, key)                                                   \
  do {                                                                       \
    if ((ptr) == nullptr) {                                                  \
      set_error(std::string("missing tensor in state dict: ") + (key));      \
      return HRP_ERR_STATE;                                                  \
    }                                                                        \
  } while (0)

// fold conv bias + BatchNorm (eval) into per-channel scale / shift, pack the weights, upload (cached per conv)
int get_weights(hrp_model* m, const std::string& conv, const std::string& bn, const ConvLayerDesc& d,
                const ConvParams& geo, ConvWeights* out) {
  auto it = m->wcache.find(conv);
  if (it != m->wcache.end()) {
    *out = it->second;
    return HRP_OK;
  }
  const HostTensor* w = find(m, conv + ".weight");
  HRP_NEED(w, conv + ".weight");
  HRP_REQUIRE(w->shape.size() == 4, "conv weight must be 4-D: " + conv);
  int cin_ref, cout_ref;
  if (d.kind == kDeconvK4S2P1) {
    cin_ref = (int)w->shape[0];
    cout_ref = (int)w->shape[1];
    HRP_REQUIRE(w->shape[2] == 4 && w->shape[3] == 4, "deconv kernel must be 4x4: " + conv);
  } else {
    cout_ref = (int)w->shape[0];
    cin_ref = (int)w->shape[1];
    HRP_REQUIRE(w->shape[2] == d.kh && w->shape[3] == d.kw, "conv kernel size mismatch: " + conv);
  }
  HRP_REQUIRE(cout_ref == d.Cout, "conv Cout mismatch: " + conv);
  HRP_REQUIRE(d.kind == kStemS2D ? cin_ref == 3 : cin_ref == d.Cin, "conv Cin mismatch: " + conv);
  std::vector<uint16_t> packed(conv_packed_weight_elems(geo));
  int rc = conv_pack_weights(d, geo, cin_ref, w->data.data(), packed.data());
  if (rc != HRP_OK) return rc;
  std::vector<float> scale(d.Cout, 1.f), bias(d.Cout, 0.f);
  const HostTensor* cb = find(m, conv + ".bias");
  if (!bn.empty()) {
    const HostTensor *g = find(m, bn + ".weight"), *b = find(m, bn + ".bias"), *mu = find(m, bn + ".running_mean"),
                     *var = find(m, bn + ".running_var");
    HRP_NEED(g, bn + ".weight");
    HRP_NEED(b, bn + ".bias");
    HRP_NEED(mu, bn + ".running_mean");
    HRP_NEED(var, bn + ".running_var");
    HRP_REQUIRE((int)g->data.size() == d.Cout && (int)var->data.size() == d.Cout, "BN size mismatch: " + bn);
    for (int c = 0; c < d.Cout; ++c) {
      const double s = (double)g->data[c] / std::sqrt((double)var->data[c] + 1e-5);
      double sh = (double)b->data[c] - (double)mu->data[c] * s;
      if (cb != nullptr) sh += (double)cb->data[c] * s;
      scale[c] = (float)s;
      bias[c] = (float)sh;
    }
  } else if (cb != nullptr) {
    for (int c = 0; c < d.Cout; ++c) bias[c] = cb->data[c];
  }
  ConvWeights cw;
  rc = dev_upload(m, packed.data(), packed.size() * 2, reinterpret_cast<void**>(&cw.w));
  if (rc != HRP_OK) return rc;
  rc = dev_upload(m, scale.data(), scale.size() * 4, reinterpret_cast<void**>(&cw.scale));
  if (rc != HRP_OK) return rc;
  rc = dev_upload(m, bias.data(), bias.size() * 4, reinterpret_cast<void**>(&cw.bias));
  if (rc != HRP_OK) return rc;
  m->wcache[conv] = cw;
  *out = cw;
  return HRP_OK;
}

// Collapse one iterative regressor (three affine layers, no activation) into delta(p) = Wx*xf + A*p + c, in fp64.
int build_regressor(hrp_model* m, const std::string& fc1, const std::string& fc2, const std::string& dec, int dim,
                    RegressorTable** out) {
  const HostTensor *W1 = find(m, fc1 + ".weight"), *b1 = find(m, fc1 + ".bias"), *W2 = find(m, fc2 + ".weight"),
                   *b2 = find(m, fc2 + ".bias"), *W3 = find(m, dec + ".weight"), *b3 = find(m, dec + ".bias");
  HRP_NEED(W1, fc1 + ".weight");
  HRP_NEED(b1, fc1 + ".bias");
  HRP_NEED(W2, fc2 + ".weight");
  HRP_NEED(b2, fc2 + ".bias");
  HRP_NEED(W3, dec + ".weight");
  HRP_NEED(b3, dec + ".bias");
  const int hid = 1024, nx = 2048, in1 = nx + dim;
  HRP_REQUIRE(W1->shape.size() == 2 && W1->shape[0] == hid && W1->shape[1] == in1, "unexpected shape: " + fc1);
  HRP_REQUIRE(W2->shape.size() == 2 && W2->shape[0] == hid && W2->shape[1] == hid, "unexpected shape: " + fc2);
  HRP_REQUIRE(W3->shape.size() == 2 && W3->shape[0] == dim && W3->shape[1] == hid, "unexpected shape: " + dec);
  std::vector<double> W32((size_t)dim * hid, 0.0);
  for (int i = 0; i < dim; ++i)
    for (int k = 0; k < hid; ++k) {
      const double a = W3->data[(size_t)i * hid + k];
      const float* row = &W2->data[(size_t)k * hid];
      double* dst = &W32[(size_t)i * hid];
      for (int j = 0; j < hid; ++j) dst[j] += a * (double)row[j];
    }
  std::vector<double> Wfull((size_t)dim * in1, 0.0);
  for (int i = 0; i < dim; ++i)
    for (int k = 0; k < hid; ++k) {
      const double a = W32[(size_t)i * hid + k];
      const float* row = &W1->data[(size_t)k * in1];
      double* dst = &Wfull[(size_t)i * in1];
      for (int j = 0; j < in1; ++j) dst[j] += a * (double)row[j];
    }
  RegressorTable t;
  memset(&t, 0, sizeof(t));
  t.dim = dim;
  t.n_iter = m->desc.n_iter;
  std::vector<float> Wx((size_t)dim * nx);
  for (int i = 0; i < dim; ++i) {
    for (int j = 0; j < nx; ++j) Wx[(size_t)i * nx + j] = (float)Wfull[(size_t)i * in1 + j];
    for (int j = 0; j < dim; ++j) t.A[i][j] = (float)Wfull[(size_t)i * in1 + nx + j];
    double c = b3->data[i];
    for (int k = 0; k < hid; ++k) c += W32[(size_t)i * hid + k] * (double)b1->data[k] + (double)W3->data[(size_t)i * hid + k] * (double)b2->data[k];
    t.c[i] = (float)c;
  }
  float* wx_dev = nullptr;
  int rc = dev_upload(m, Wx.data(), Wx.size() * 4, reinterpret_cast<void**>(&wx_dev));
  if (rc != HRP_OK) return rc;
  t.Wx = wx_dev;
  return dev_upload(m, &t, sizeof(t), reinterpret_cast<void**>(out));
}

// ------------------------------------------------------------------------------------------------------
// program builder
// ------------------------------------------------------------------------------------------------------
struct Builder {
  hrp_model* m;
  Plan* pl;
  int rc = HRP_OK;
  // set before the conv() call of the final 1x1 layer: its epilogue then writes soft-argmax partials instead of logits
  float* fold_partials = nullptr;
  int fold_chunks = 0, fold_nkpt = 0;
  // Activation memory.  Pass 1 (dry): the program is traversed without touching the device -- only tensor shapes and
  // the reads / writes of every op are recorded; plan_arena() then gives every tensor an offset in ONE arena, re-using
  // the space of tensors whose last reader has already run (valid when the ops execute in program order, i.e. for
  // single-lane plans; multi-lane plans keep one region per tensor).  Pass 2 builds the real plans on those addresses.
  bool dry = false;
  char* arena = nullptr;
  std::vector<size_t> offsets;   // per activation id (pass 2)

  int new_act(int H, int W, int C) {
    Act a;
    a.H = H;
    a.W = W;
    a.C = C;
    const int id = (int)pl->acts.size();
    if (!dry && rc == HRP_OK) {
      if (id >= (int)offsets.size()) {
        set_error("internal: activation count differs between the planning passes");
        rc = HRP_ERR_STATE;
      } else {
        a.ptr = reinterpret_cast<bf16*>(arena + offsets[id]);
      }
    }
    pl->acts.push_back(a);
    return id;
  }

  // dry pass: record an op by its reads / writes only
  int dry_op(int lane, int in, const Epi* epi, int out) {
    Op op;
    op.lane = lane;
    if (in >= 0) op.reads.push_back(in);
    if (epi != nullptr) {
      for (int i = 0; i < 3; ++i) {
        if (epi->pre[i] >= 0) op.reads.push_back(epi->pre[i]);
        if (epi->up[i] >= 0) op.reads.push_back(epi->up[i]);
      }
      if (epi->post >= 0) op.reads.push_back(epi->post);
    }
    op.writes = out;
    pl->ops.push_back(op);
    if (out >= 0) pl->acts[out].producer = (int)pl->ops.size() - 1;
    return out;
  }

  // conv + (bias) + BN + addends + ReLU; returns the output activation id (or -1 when only pooled)
  int conv(int lane, const std::string& name, const std::string& bn, int in, int cout, int k, int stride, int pad,
           bool relu, const Epi& epi = Epi(), int kind = kConv, bool write_out = true, float* pool_out = nullptr,
           const bf16* raw_in = nullptr, int raw_H = 0, int raw_W = 0, int raw_C = 0) {
    if (rc != HRP_OK) return -1;
    ConvLayerDesc d;
    d.kind = kind;
    d.B = pl->B;
    d.in_wpitch = d.in_wpad = 0;
    if (raw_in != nullptr || in < 0) {
      d.Hin = raw_H;
      d.Win = raw_W;
      d.Cin = raw_C;
      if (kind == kStemS2D) {
        d.in_wpitch = kS2dPitch;
        d.in_wpad = kS2dPad;
      }
    } else {
      d.Hin = pl->acts[in].H;
      d.Win = pl->acts[in].W;
      d.Cin = pl->acts[in].C;
    }
    d.Cout = cout;
    d.kh = d.kw = k;
    d.stride = stride;
    d.pad = pad;
    d.relu = relu ? 1 : 0;
    d.has_residual = (epi.pre[0] >= 0) ? 1 : 0;
    d.head_fold = (fold_partials != nullptr) ? 1 : 0;
    Op op;
    op.kind = OP_CONV;
    op.lane = lane;
    rc = conv_geometry(d, &op.conv.p);
    if (rc != HRP_OK) return -1;
    if (dry) {
      fold_partials = nullptr;
      const int o = write_out ? new_act(op.conv.p.Hout, op.conv.p.Wout, cout) : -1;
      return dry_op(lane, in, &epi, o);
    }
    ConvWeights cw;
    rc = get_weights(m, name, bn, d, op.conv.p, &cw);
    if (rc != HRP_OK) return -1;
    ConvParams& p = op.conv.p;
    int out = -1;
    if (write_out) {
      out = new_act(p.Hout, p.Wout, cout);
      if (rc != HRP_OK) return -1;
      p.out = pl->acts[out].ptr;
    }
    p.scale = cw.scale;
    p.bias = cw.bias;
    p.pool_out = pool_out;
    if (fold_partials != nullptr) {
      p.head_partials = fold_partials;
      p.head_chunks = fold_chunks;
      p.head_nkpt = fold_nkpt;
      fold_partials = nullptr;
    }
    if (in >= 0) op.reads.push_back(in);
    for (int i = 0; i < 3; ++i) {
      if (epi.pre[i] >= 0) {
        const Act& a = pl->acts[epi.pre[i]];
        if (a.H != p.Hout || a.W != p.Wout || a.C != cout) {
          set_error("internal: pre addend shape mismatch at " + name);
          rc = HRP_ERR_STATE;
          return -1;
        }
        p.pre[i] = a.ptr;
        op.reads.push_back(epi.pre[i]);
      }
      if (epi.up[i] >= 0) {
        const Act& a = pl->acts[epi.up[i]];
        if ((a.H << epi.up_shift[i]) != p.Hout || a.C != cout) {
          set_error("internal: up addend shape mismatch at " + name);
          rc = HRP_ERR_STATE;
          return -1;
        }
        p.up[i] = a.ptr;
        p.up_shift[i] = epi.up_shift[i];
        op.reads.push_back(epi.up[i]);
      }
    }
    if (epi.post >= 0) {
      p.post = pl->acts[epi.post].ptr;
      op.reads.push_back(epi.post);
    }
    rc = conv_plan_finalize(&op.conv, raw_in != nullptr ? raw_in : pl->acts[in].ptr, cw.w, &d);
    if (rc != HRP_OK) return -1;
    // algorithmic FLOPs of the reference layer (2*MACs, unpadded)
    {
      const ConvParams& q = op.conv.p;
      double macs;
      if (kind == kDeconvK4S2P1) macs = (double)pl->B * q.Hout * q.Wout * 4.0 * q.Cin * cout;
      else if (kind == kConvUp2) macs = (double)pl->B * q.Hm * q.Wm * q.Cin * cout;  // the reference's 1x1 conv at low resolution
      else if (kind == kStemS2D) macs = (double)pl->B * q.Hout * q.Wout * (double)k * k * 3.0 * cout;
      else macs = (double)pl->B * q.Hout * q.Wout * (double)k * k * q.Cin * cout;
      op.conv.flops = 2.0 * macs;
      pl->flops += op.conv.flops;
    }
    op.writes = out;
    op.name = name;
    pl->ops.push_back(op);
    if (out >= 0) pl->acts[out].producer = (int)pl->ops.size() - 1;
    return out;
  }

  int maxpool(int lane, int in) {
    if (rc != HRP_OK) return -1;
    const Act a = pl->acts[in];
    const int out = new_act(a.H / 2, a.W / 2, a.C);
    if (rc != HRP_OK) return -1;
    if (dry) return dry_op(lane, in, nullptr, out);
    Op op;
    op.kind = OP_MAXPOOL;
    op.name = "maxpool";
    op.lane = lane;
    op.mp_in = a.ptr;
    op.mp_out = pl->acts[out].ptr;
    op.mp_B = pl->B;
    op.mp_H = a.H;
    op.mp_W = a.W;
    op.mp_C = a.C;
    op.reads.push_back(in);
    op.writes = out;
    pl->ops.push_back(op);
    pl->acts[out].producer = (int)pl->ops.size() - 1;
    return out;
  }

  int fuse_add(int lane, int pre, const Epi& epi) {
    if (rc != HRP_OK) return -1;
    const Act a = pl->acts[pre];
    const int out = new_act(a.H, a.W, a.C);
    if (rc != HRP_OK) return -1;
    if (dry) {
      Epi e2 = epi;
      for (int i = 0; i < 3; ++i) e2.pre[i] = -1;
      e2.post = -1;
      return dry_op(lane, pre, &e2, out);
    }
    Op op;
    op.kind = OP_FUSEADD;
    op.name = "fuse_add";
    op.lane = lane;
    memset(&op.fuse, 0, sizeof(op.fuse));
    op.fuse.pre = a.ptr;
    op.fuse.out = pl->acts[out].ptr;
    op.fuse.B = pl->B;
    op.fuse.H = a.H;
    op.fuse.W = a.W;
    op.fuse.C = a.C;
    op.fuse.relu = 1;
    op.reads.push_back(pre);
    for (int i = 0; i < 3; ++i)
      if (epi.up[i] >= 0) {
        op.fuse.up[i] = pl->acts[epi.up[i]].ptr;
        op.fuse.up_shift[i] = epi.up_shift[i];
        op.reads.push_back(epi.up[i]);
      }
    op.writes = out;
    pl->ops.push_back(op);
    pl->acts[out].producer = (int)pl->ops.size() - 1;
    return out;
  }

  void memset_op(int lane, void* ptr, size_t bytes) {
    if (dry) return;
    Op op;
    op.kind = OP_MEMSET;
    op.lane = lane;
    op.ms_ptr = ptr;
    op.ms_bytes = bytes;
    pl->ops.push_back(op);
  }

  // Bottleneck (Resnet.py:96-135, HRnet.py:60-98): 1x1 -> 3x3(stride) -> 1x1 (+ residual / downsample) -> ReLU
  int bottleneck(int lane, const std::string& n, int x, int planes, int stride, int post = -1, float* pool_out = nullptr) {
    int t = conv(lane, n + ".conv1", n + ".bn1", x, planes, 1, 1, 0, true);
    t = conv(lane, n + ".conv2", n + ".bn2", t, planes, 3, stride, 1, true);
    int res = x;
    if (find(m, n + ".downsample.0.weight") != nullptr || m->wcache.count(n + ".downsample.0") != 0)
      res = conv(lane, n + ".downsample.0", n + ".downsample.1", x, planes * 4, 1, stride, 0, false);
    Epi e;
    e.pre[0] = res;
    e.post = post;
    return conv(lane, n + ".conv3", n + ".bn3", t, planes * 4, 1, 1, 0, true, e, kConv, true, pool_out);
  }

  // BasicBlock (HRnet.py:28-57)
  int basic(int lane, const std::string& n, int x, int c) {
    int t = conv(lane, n + ".conv1", n + ".bn1", x, c, 3, 1, 1, true);
    Epi e;
    e.pre[0] = x;
    return conv(lane, n + ".conv2", n + ".bn2", t, c, 3, 1, 1, true, e);
  }

  void tap(const std::string& name, int act) { pl->taps[name] = act; }

  // ResNet-50 trunk (Resnet.py:56-67) on the s2d-packed input; returns the (B,8,8,2048) activation
  int resnet50(const std::string& pre, const bf16* s2d, float* pool_out) {
    const int lane = kLaneRN;
    int x = conv(lane, pre + "conv1", pre + "bn1", -1, 64, 7, 2, 3, true, Epi(), kStemS2D, true, nullptr, s2d, kImg / 2,
                 kImg / 2, 16);
    tap(pre + "stem", x);
    x = maxpool(lane, x);
    const int planes[4] = {64, 128, 256, 512}, blocks[4] = {3, 4, 6, 3};
    for (int li = 0; li < 4; ++li) {
      for (int b = 0; b < blocks[li]; ++b) {
        const bool last = (li == 3 && b == blocks[li] - 1);
        x = bottleneck(lane, pre + "layer" + std::to_string(li + 1) + "." + std::to_string(b), x, planes[li],
                       (b == 0 && li > 0) ? 2 : 1, -1, last ? pool_out : nullptr);
      }
      tap(pre + "layer" + std::to_string(li + 1), x);
    }
    return x;
  }

  // HRNet-w32 with classification head, pooled feature only (HRnet.py:499-570, generate_hm=False)
  void hrnet32(const std::string& pre, const bf16* s2d, float* pool_out) {
    const int C[4] = {32, 64, 128, 256};
    int x = conv(0, pre + "conv1", pre + "bn1", -1, 64, 3, 2, 1, true, Epi(), kStemS2D, true, nullptr, s2d, kImg / 2,
                 kImg / 2, 16);
    x = conv(0, pre + "conv2", pre + "bn2", x, 64, 3, 2, 1, true);
    for (int b = 0; b < 4; ++b) x = bottleneck(0, pre + "layer1." + std::to_string(b), x, 64, 1);
    tap(pre + "layer1", x);
    std::vector<int> ys;
    ys.push_back(conv(0, pre + "transition1.0.0", pre + "transition1.0.1", x, C[0], 3, 1, 1, true));
    ys.push_back(conv(1, pre + "transition1.1.0.0", pre + "transition1.1.0.1", x, C[1], 3, 2, 1, true));
    const int nmods[3] = {1, 4, 3};
    for (int stage = 2; stage <= 4; ++stage) {
      const int nb = stage;
      if (stage > 2) {  // HRnet.py:516-521,524-529: the new branch comes from the previous stage's LAST output
        const std::string t = pre + "transition" + std::to_string(stage - 1) + "." + std::to_string(nb - 1) + ".0";
        ys.push_back(conv(nb - 1, t + ".0", t + ".1", ys.back(), C[nb - 1], 3, 2, 1, true));
      }
      for (int mod = 0; mod < nmods[stage - 2]; ++mod) {
        const std::string mn = pre + "stage" + std::to_string(stage) + "." + std::to_string(mod);
        std::vector<int> xs(nb);
        for (int br = 0; br < nb; ++br) {
          int v = ys[br];
          for (int blk = 0; blk < 4; ++blk)
            v = basic(br, mn + ".branches." + std::to_string(br) + "." + std::to_string(blk), v, C[br]);
          xs[br] = v;
        }
        // fuse (HRnet.py:254-263): y_i = relu(sum_j f_ij(x_j)); everything is summed in fp32 in ONE epilogue
        std::vector<int> fused(nb);
        for (int i = 0; i < nb; ++i) {
          Epi e;
          int nup = 0, npre = 0;
          e.pre[npre++] = xs[i];
          if (i == 0 && m->fuse0_epilogue) {
            // y0 = relu(x0 + up2(f01(x1)) + up4(f02(x2)) + up8(f03(x3))): the j = 1 term's 1x1 conv runs as an upsampling
            // conv (4 output phases, one weight matrix) whose epilogue adds x0 and the other (low-resolution) terms and
            // applies the ReLU -- the sum never takes a separate elementwise pass
            for (int j = 2; j < nb; ++j) {
              const std::string f = mn + ".fuse_layers.0." + std::to_string(j);
              e.up[nup] = conv(j, f + ".0", f + ".1", xs[j], C[0], 1, 1, 0, false);
              e.up_shift[nup] = j;
              ++nup;
            }
            const std::string f1 = mn + ".fuse_layers.0.1";
            fused[0] = conv(0, f1 + ".0", f1 + ".1", xs[1], C[0], 1, 1, 0, true, e, kConvUp2);
            continue;
          }
          for (int j = i + 1; j < nb; ++j) {  // 1x1 conv + BN at the low resolution, upsampled in the consumer
            const std::string f = mn + ".fuse_layers." + std::to_string(i) + "." + std::to_string(j);
            e.up[nup] = conv(j, f + ".0", f + ".1", xs[j], C[i], 1, 1, 0, false);
            e.up_shift[nup] = j - i;
            ++nup;
          }
          if (i == 0) {
            // branch 0 has no conv of its own to carry the sum.  HRP_FUSE0_EPI=0: standalone elementwise kernel.
            fused[i] = fuse_add(0, xs[0], e);
            continue;
          }
          for (int j = 0; j < i; ++j) {  // chains of stride-2 3x3 convs; the last chain carries the summation
            const std::string f = mn + ".fuse_layers." + std::to_string(i) + "." + std::to_string(j);
            int t = xs[j];
            for (int k = 0; k < i - j - 1; ++k)
              t = conv(i, f + "." + std::to_string(k) + ".0", f + "." + std::to_string(k) + ".1", t, C[j], 3, 2, 1, true);
            const std::string lastc = f + "." + std::to_string(i - j - 1);
            if (j < i - 1) {
              e.pre[npre++] = conv(i, lastc + ".0", lastc + ".1", t, C[i], 3, 2, 1, false);
            } else {
              fused[i] = conv(i, lastc + ".0", lastc + ".1", t, C[i], 3, 2, 1, true, e);
            }
          }
        }
        ys = fused;
      }
      for (int i = 0; i < nb; ++i) tap(pre + "stage" + std::to_string(stage) + ".out" + std::to_string(i), ys[i]);
    }
    // classification head (HRnet.py:557-568)
    const int head_ch[4] = {32, 64, 128, 256};
    int y = bottleneck(0, pre + "incre_modules.0.0", ys[0], head_ch[0], 1);
    for (int i = 0; i < 3; ++i) {
      const std::string dn = pre + "downsamp_modules." + std::to_string(i);
      const int d = conv(0, dn + ".0", dn + ".1", y, head_ch[i + 1] * 4, 3, 2, 1, true);
      y = bottleneck(0, pre + "incre_modules." + std::to_string(i + 1) + ".0", ys[i + 1], head_ch[i + 1], 1, d);
    }
    tap(pre + "cls_y", y);
    conv(0, pre + "final_feat_layer.0", pre + "final_feat_layer.1", y, 2048, 1, 1, 0, true, Epi(), kConv, false, pool_out);
  }
};

int launch_op(const hrp_model* m, const Op& op, cudaStream_t s) {
  switch (op.kind) {
    case OP_CONV:
      return m->use_simt ? conv_plan_launch_simt(op.conv, s) : conv_plan_launch(op.conv, s);
    case OP_MAXPOOL:
      return launch_maxpool3x3s2(op.mp_in, op.mp_out, op.mp_B, op.mp_H, op.mp_W, op.mp_C, s);
    case OP_FUSEADD:
      return launch_fuse_add(op.fuse, s);
    case OP_MEMSET:
      HRP_CUDA_CHECK(cudaMemsetAsync(op.ms_ptr, 0, op.ms_bytes, s));
      return HRP_OK;
  }
  return HRP_ERR_STATE;
}

// Kernel variant per conv: the committed tuning table (hrp_model_set_tuning) decides where it has an entry -- the same
// choice on every box and every run, so results are bitwise reproducible --, the shape heuristics of conv_plan_finalize
// elsewhere.  With HRP_AUTOTUNE=1 shapes missing from the table are timed on the plan's own buffers instead (3 launches
// per variant, CUDA events) and the winners are added to the cache (hrp_model_get_tuning dumps it: that is how the table
// is produced, tools/make_tuning.py).
void tune_key(const Op& op, char* key, size_t len) {
  const ConvParams& q = op.conv.p;
  snprintf(key, len, "%d:%d:%d:%d:%d:%d:%d:%d:%d:%d:%d:%d:%d", q.B, q.Hin, q.Win, q.Cin, q.Cout, q.Hout, q.ntaps, q.nphase,
           q.src_sh, op.conv.epi, q.pre[0] != nullptr, q.pool_out != nullptr, q.out != nullptr);
}

void apply_variant(Op& op, int pick) {
  op.conv.halo = (pick == 2) && op.conv.halo_ok;
  if (pick < 2) op.conv.persistent = (pick == 1);
}

int autotune_plan(hrp_model* m, Plan* pl) {
  if (m->use_simt || m->pin_variants) return HRP_OK;
  cudaStream_t s = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  char key[256];
  int rc = HRP_OK;
  for (auto& op : pl->ops) {
    if (op.kind != OP_CONV || op.conv.p.head_partials != nullptr) continue;  // (the fold lives in one kernel only)
    if (op.conv.pcfg.staged) {  // TMA-staged addends exist in the persistent kernel only (and beat gathers 2-3x)
      op.conv.persistent = true;
      op.conv.halo = false;
      continue;
    }
    tune_key(op, key, sizeof(key));
    auto it = m->tune_cache.find(key);
    if (it != m->tune_cache.end()) {
      apply_variant(op, it->second);
      continue;
    }
    if (!m->autotune) continue;  // heuristic default of conv_plan_finalize
    if (s == nullptr) {
      HRP_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
      HRP_CUDA_CHECK(cudaEventCreate(&e0));
      HRP_CUDA_CHECK(cudaEventCreate(&e1));
    }
    float best[3] = {0.f, 0.f, 0.f};
    const int nvar = op.conv.halo_ok ? 3 : 2;
    for (int variant = 0; variant < nvar && rc == HRP_OK; ++variant) {
      op.conv.halo = (variant == 2);
      op.conv.persistent = (variant == 1);
      rc = conv_plan_launch(op.conv, s);  // warm-up (also faults in code / descriptors)
      if (rc != HRP_OK) break;
      cudaEventRecord(e0, s);
      for (int i = 0; i < 3 && rc == HRP_OK; ++i) rc = conv_plan_launch(op.conv, s);
      cudaEventRecord(e1, s);
      if (cudaEventSynchronize(e1) != cudaSuccess) {
        set_error(std::string("autotune launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        rc = HRP_ERR_CUDA;
        break;
      }
      cudaEventElapsedTime(&best[variant], e0, e1);
    }
    if (rc != HRP_OK) break;
    int pick = (best[1] < best[0]) ? 1 : 0;
    if (nvar == 3 && best[2] < best[pick]) pick = 2;
    apply_variant(op, pick);
    m->tune_cache[key] = pick;
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (s) cudaStreamDestroy(s);
  return rc;
}

int capture_plan(hrp_model* m, Plan* pl) {
  // Lanes pay off while some kernels cannot fill the GPU by themselves (the 8x8 / 16x16 layers of a 64-image chunk are
  // 32 / 128 tiles for 148 SMs).  Measured on B200, Panda full model, one stream vs five lanes: 32 images 5.28 -> 3.75 ms,
  // 64 images 6.94 -> 5.91 ms, 128 images 10.64 -> 10.28 ms, 256 images 19.26 -> 19.18 ms (a tie); at 512 images every
  // launch is a full persistent grid and lanes only contend for SMs, shared memory and L2 (12.9k img/s on one stream vs
  // 12.6k on five).  HRP_SINGLE_LANE=0/1 overrides.
  const bool single_lane = pl->single_lane;
  if (single_lane)
    for (auto& op : pl->ops) op.lane = 0;
  // cross-lane dependencies -> events
  for (size_t i = 0; i < pl->ops.size(); ++i)
    for (int a : pl->ops[i].reads) {
      const int prod = pl->acts[a].producer;
      if (prod >= 0 && pl->ops[prod].lane != pl->ops[i].lane) pl->ops[prod].cross_lane_consumer = true;
    }
  for (auto& op : pl->ops)
    if (op.cross_lane_consumer) HRP_CUDA_CHECK(cudaEventCreateWithFlags(&op.done, cudaEventDisableTiming));
  for (int l = 0; l < kNumLanes; ++l) {
    HRP_CUDA_CHECK(cudaStreamCreateWithFlags(&pl->cap[l], cudaStreamNonBlocking));
    HRP_CUDA_CHECK(cudaEventCreateWithFlags(&pl->join_ev[l], cudaEventDisableTiming));
  }
  HRP_CUDA_CHECK(cudaEventCreateWithFlags(&pl->fork_ev, cudaEventDisableTiming));
  HRP_CUDA_CHECK(cudaEventCreateWithFlags(&pl->done_ev, cudaEventDisableTiming));
  HRP_CUDA_CHECK(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
  pl->n_kernels = 0;
  for (auto& op : pl->ops)
    if (op.kind != OP_MEMSET) ++pl->n_kernels;
  if (!m->use_graph) return HRP_OK;

  const int64_t launches_before = g_launch_count.load();
  HRP_CUDA_CHECK(cudaStreamBeginCapture(pl->cap[0], cudaStreamCaptureModeThreadLocal));
  int rc = HRP_OK;
  do {
    if (cudaEventRecord(pl->fork_ev, pl->cap[0]) != cudaSuccess) { rc = HRP_ERR_CUDA; break; }
    for (int l = 1; l < kNumLanes && rc == HRP_OK; ++l)
      if (cudaStreamWaitEvent(pl->cap[l], pl->fork_ev, 0) != cudaSuccess) rc = HRP_ERR_CUDA;
    // Programmatic dependent launch between consecutive kernels of a lane (HRP_PDL=1 enables): the successor's
    // prologue overlaps the predecessor's tail and the ~5 us launch gap disappears.  Not used behind a memset
    // node or a cross-lane event wait.
    const char* pdl_env = getenv("HRP_PDL");
    // opt-in (measured no gain) and single-lane graphs only: with five lanes a batch-1 run produced a
    // last-bit different pose under PDL, i.e. some cross-lane edge is not covered by the programmatic dependency
    // Default: on for single-lane plans of <= 256 images (the kernel tails it hides are a few us each: 3-7 % of a step at
    // 64-128 images, nothing at 512 where round 1 measured -2 %); HRP_PDL=0 / 1 overrides.
    const bool pdl_want = (pdl_env != nullptr) ? (pdl_env[0] == '1') : (pl->B <= 256);
    const bool pdl_on = pdl_want && !m->use_simt && single_lane;
    bool prev_kernel[kNumLanes] = {};
    for (size_t i = 0; i < pl->ops.size() && rc == HRP_OK; ++i) {
      Op& op = pl->ops[i];
      bool waited = false;
      for (int a : op.reads) {
        const int prod = pl->acts[a].producer;
        if (prod >= 0 && pl->ops[prod].lane != op.lane) {
          waited = true;
          if (cudaStreamWaitEvent(pl->cap[op.lane], pl->ops[prod].done, 0) != cudaSuccess) rc = HRP_ERR_CUDA;
        }
      }
      if (rc != HRP_OK) break;
      {
        PdlScope pdl(pdl_on && op.kind != OP_MEMSET && prev_kernel[op.lane] && !waited);
        rc = launch_op(m, op, pl->cap[op.lane]);
      }
      prev_kernel[op.lane] = (op.kind != OP_MEMSET);
      if (rc == HRP_OK && op.cross_lane_consumer)
        if (cudaEventRecord(op.done, pl->cap[op.lane]) != cudaSuccess) rc = HRP_ERR_CUDA;
    }
    for (int l = 1; l < kNumLanes && rc == HRP_OK; ++l) {
      if (cudaEventRecord(pl->join_ev[l], pl->cap[l]) != cudaSuccess) rc = HRP_ERR_CUDA;
      if (rc == HRP_OK && cudaStreamWaitEvent(pl->cap[0], pl->join_ev[l], 0) != cudaSuccess) rc = HRP_ERR_CUDA;
    }
  } while (0);
  cudaError_t e = cudaStreamEndCapture(pl->cap[0], &pl->graph);
  g_launch_count.store(launches_before);  // captured launches are counted per replay instead
  if (rc != HRP_OK || e != cudaSuccess) {
    if (rc == HRP_OK || rc == HRP_ERR_CUDA)
      set_error(std::string("CUDA graph capture failed: ") + cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
    return HRP_ERR_CUDA;
  }
  HRP_CUDA_CHECK(cudaGraphInstantiate(&pl->exec, pl->graph, 0));
  return HRP_OK;
}

// The network program (both planning passes run exactly this): HRNet-w32 root-depth net, and for the full model the
// ResNet-50 trunk, the deconv head and the final 1x1 conv (with or without the soft-argmax fold).
void emit_program(Builder& b, hrp_model* m, Plan* pl, bool fold) {
  const bool full = (m->desc.kind == HRP_MODEL_FULL);
  b.memset_op(0, pl->feat, (size_t)pl->B * 2048 * 4);
  if (full) b.memset_op(kLaneRN, pl->xf, (size_t)pl->B * 2048 * 4);
  b.hrnet32(m->prefix_root, pl->s2d_root, pl->feat);
  if (!full || b.rc != HRP_OK) return;
  int x = b.resnet50("reg_backbone.", pl->s2d_reg, pl->xf);
  const int dc[3] = {256, 256, 256};
  for (int i = 0; i < 3 && b.rc == HRP_OK; ++i)
    x = b.conv(kLaneRN, "deconv_layers." + std::to_string(3 * i), "deconv_layers." + std::to_string(3 * i + 1), x, dc[i], 4, 2,
               1, true, Epi(), kDeconvK4S2P1);
  b.tap("deconv", x);
  if (fold) {
    // the final 1x1 conv reduces its logits to per-warp soft-argmax partials in its epilogue: 64 rows / 2 rows per tile
    // x 4 warps x 2 depth halves = 256 chunks per image; the (B, 4096, nkpt*64) heatmap is never materialised
    b.fold_partials = pl->fold_partials != nullptr ? pl->fold_partials : reinterpret_cast<float*>(16);  // (dry pass: a flag)
    b.fold_chunks = pl->fold_chunks;
    b.fold_nkpt = m->desc.nkpt;
    b.conv(kLaneRN, "final_layer", "", x, m->desc.nkpt * 64, 1, 1, 0, false, Epi(), kConv, false);
    pl->heatmap = -1;
  } else {
    x = b.conv(kLaneRN, "final_layer", "", x, m->desc.nkpt * 64, 1, 1, 0, false);
    b.tap("heatmap", x);
    pl->heatmap = x;
  }
}

// Offsets of every activation in one arena.  alias = false: one region per tensor.  alias = true (single-lane plans: the
// ops run in program order): a tensor's region is released once its last reader has been emitted and re-used, best fit,
// by later tensors.  An op's output never shares memory with its own inputs (they are released only AFTER the op);
// tensors with a tap, without a reader inside the program (the heatmap: read by the head kernel) are pinned.
size_t plan_arena(const Plan& dry, bool alias, std::vector<size_t>* offsets, size_t* sum_bytes) {
  const int n = (int)dry.acts.size();
  std::vector<size_t> bytes(n);
  std::vector<int> def(n, 0), last(n, -1);
  size_t sum = 0;
  for (int i = 0; i < n; ++i) {
    const Act& a = dry.acts[i];
    bytes[i] = ((size_t)dry.B * a.H * a.W * a.C * sizeof(bf16) + 1023) / 1024 * 1024;
    sum += bytes[i];
    def[i] = std::max(a.producer, 0);
  }
  for (size_t o = 0; o < dry.ops.size(); ++o)
    for (int a : dry.ops[o].reads) last[a] = std::max(last[a], (int)o);
  for (auto& kv : dry.taps) last[kv.second] = 1 << 30;
  if (dry.heatmap >= 0) last[dry.heatmap] = 1 << 30;
  for (int i = 0; i < n; ++i)
    if (last[i] < 0) last[i] = 1 << 30;
  *sum_bytes = sum;
  offsets->assign(n, 0);
  if (!alias) {
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
      (*offsets)[i] = off;
      off += bytes[i];
    }
    return off;
  }
  struct Blk { size_t off, size; };
  std::vector<Blk> free_list;             // sorted by offset, coalesced
  std::vector<std::pair<int, int>> live;  // (last reader op, act id)
  size_t end = 0;
  auto release = [&](size_t off, size_t size) {
    size_t i = 0;
    while (i < free_list.size() && free_list[i].off < off) ++i;
    free_list.insert(free_list.begin() + i, Blk{off, size});
    if (i + 1 < free_list.size() && free_list[i].off + free_list[i].size == free_list[i + 1].off) {
      free_list[i].size += free_list[i + 1].size;
      free_list.erase(free_list.begin() + i + 1);
    }
    if (i > 0 && free_list[i - 1].off + free_list[i - 1].size == free_list[i].off) {
      free_list[i - 1].size += free_list[i].size;
      free_list.erase(free_list.begin() + i);
    }
  };
  for (int i = 0; i < n; ++i) {  // activation ids are created in program order
    for (size_t j = 0; j < live.size();) {
      if (live[j].first < def[i]) {
        release((*offsets)[live[j].second], bytes[live[j].second]);
        live.erase(live.begin() + j);
      } else {
        ++j;
      }
    }
    int best = -1;
    for (size_t j = 0; j < free_list.size(); ++j)
      if (free_list[j].size >= bytes[i] && (best < 0 || free_list[j].size < free_list[best].size)) best = (int)j;
    if (best >= 0) {
      (*offsets)[i] = free_list[best].off;
      free_list[best].off += bytes[i];
      free_list[best].size -= bytes[i];
      if (free_list[best].size == 0) free_list.erase(free_list.begin() + best);
    } else if (!free_list.empty() && free_list.back().off + free_list.back().size == end) {
      (*offsets)[i] = free_list.back().off;  // grow the arena from the free block that touches its end
      end = free_list.back().off + bytes[i];
      free_list.pop_back();
    } else {
      (*offsets)[i] = end;
      end += bytes[i];
    }
    live.push_back(std::make_pair(last[i], i));
  }
  // self-check (cheap, ~1e5 pairs): tensors whose lifetimes [def, last reader] intersect must not share memory
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      const bool life = def[i] <= last[j] && def[j] <= last[i];
      const bool mem = (*offsets)[i] < (*offsets)[j] + bytes[j] && (*offsets)[j] < (*offsets)[i] + bytes[i];
      if (life && mem) return 0;  // (reported by the caller)
    }
  return end;
}

int build_plan(hrp_model* m, int B, Plan** out_plan, int replica) {
  std::unique_ptr<Plan> pl(new Plan());
  pl->B = B;
  // Lanes (intra-step parallelism) pay off while some kernels cannot fill the GPU by themselves (see capture_plan) AND one
  // step is in flight.  Large chunks run on one stream; so do plans of >= 32 images when the caller keeps several steps in
  // flight (desc.inflight >= 2: plan replicas on caller streams): independent steps then fill each other's gaps better
  // than lanes do, and a single-lane graph can chain its kernels by programmatic dependent launch
  // (profiles/r02_exp_small_shards.txt: 64 images 5.76 -> 5.26 ms, 128 images 9.68 -> 9.10 ms per step).
  pl->single_lane = B > 256 || (m->desc.inflight >= 2 && B >= 32);
  if (const char* sl = getenv("HRP_SINGLE_LANE")) pl->single_lane = (sl[0] == '1');
  const bool full = (m->desc.kind == HRP_MODEL_FULL);
  const bool fold = full && m->head_fold && !m->use_simt;
  auto dmalloc = [&](size_t bytes, void** p) -> int {
    HRP_CUDA_CHECK(cudaMalloc(p, bytes));
    pl->owned.push_back(*p);
    return HRP_OK;
  };
  // ---- pass 1: shapes and liveness only ----
  std::vector<size_t> offsets;
  {
    Plan dry;
    dry.B = B;
    dry.fold_chunks = 256;
    dry.s2d_root = dry.s2d_reg = reinterpret_cast<bf16*>(16);  // (never dereferenced: marks the stems' raw inputs)
    Builder bd{m, &dry};
    bd.dry = true;
    emit_program(bd, m, &dry, fold);
    if (bd.rc != HRP_OK) return bd.rc;
    const char* al = getenv("HRP_ALIAS");
    const bool alias = pl->single_lane && !(al != nullptr && al[0] == '0');
    pl->arena_bytes = plan_arena(dry, alias, &offsets, &pl->act_bytes_sum);
    if (pl->arena_bytes == 0) {
      set_error("internal: activation arena planning produced overlapping live tensors");
      return HRP_ERR_STATE;
    }
  }
  // ---- device buffers ----
  static_assert(kS2dPitch == kImg / 2 + 2 * kS2dPad, "s2d padding");
  const size_t s2d_bytes = (size_t)B * (kImg / 2) * kS2dPitch * 16 * sizeof(bf16);
  int rc = dmalloc(s2d_bytes, reinterpret_cast<void**>(&pl->s2d_root));
  if (rc != HRP_OK) return rc;
  HRP_CUDA_CHECK(cudaMemset(pl->s2d_root, 0, s2d_bytes));  // the padding columns stay zero: the pack kernel skips them
  rc = dmalloc((size_t)B * 2048 * 4, reinterpret_cast<void**>(&pl->feat));
  if (rc != HRP_OK) return rc;
  if (full) {
    rc = dmalloc(s2d_bytes, reinterpret_cast<void**>(&pl->s2d_reg));
    if (rc != HRP_OK) return rc;
    HRP_CUDA_CHECK(cudaMemset(pl->s2d_reg, 0, s2d_bytes));
    rc = dmalloc((size_t)B * 2048 * 4, reinterpret_cast<void**>(&pl->xf));
    if (rc != HRP_OK) return rc;
  }
  if (fold) {
    pl->fold_chunks = 256;
    rc = dmalloc((size_t)B * pl->fold_chunks * m->desc.nkpt * 5 * sizeof(float), reinterpret_cast<void**>(&pl->fold_partials));
    if (rc != HRP_OK) return rc;
    rc = dmalloc((size_t)B * m->desc.nkpt * 5 * sizeof(float), reinterpret_cast<void**>(&pl->fold_merged));
    if (rc != HRP_OK) return rc;
  }
  void* arena = nullptr;
  {
    cudaError_t e = cudaMalloc(&arena, std::max<size_t>(pl->arena_bytes, 1024));
    if (e != cudaSuccess) {
      set_error(std::string("activation arena allocation failed (") + std::to_string(pl->arena_bytes >> 20) + " MiB): " +
                cudaGetErrorString(e));
      return HRP_ERR_CUDA;
    }
    pl->owned.push_back(arena);
  }
  // ---- pass 2: the real program on the planned addresses ----
  Builder b{m, pl.get()};
  b.arena = reinterpret_cast<char*>(arena);
  b.offsets = offsets;
  emit_program(b, m, pl.get(), fold);
  if (b.rc != HRP_OK) return b.rc;
  if (full && pl->heatmap >= 0) {
    pl->head_chunks = head_default_chunks(B);
    const size_t ws = ((size_t)B * 4 + 255) / 256 * 256 + head_partials_elems(B, m->desc.nkpt, pl->head_chunks) * 4;
    rc = dmalloc(ws, &pl->head_ws);
    if (rc != HRP_OK) return rc;
    HRP_CUDA_CHECK(cudaMemset(pl->head_ws, 0, ws));
  }
  {
    const int64_t launches_before = g_launch_count.load();
    rc = autotune_plan(m, pl.get());
    g_launch_count.store(launches_before);
    if (rc != HRP_OK) return rc;
    if (pl->feat) HRP_CUDA_CHECK(cudaMemset(pl->feat, 0, (size_t)B * 2048 * 4));
    if (pl->xf) HRP_CUDA_CHECK(cudaMemset(pl->xf, 0, (size_t)B * 2048 * 4));
  }
  rc = capture_plan(m, pl.get());
  if (rc != HRP_OK) return rc;
  *out_plan = pl.get();
  m->plans[std::make_pair(B, replica)] = std::move(pl);
  return HRP_OK;
}

// Plans are cached per (batch, replica).  The nominal-chunk plans live for the model's lifetime; plans of other batch
// sizes (ragged tails, callers that vary B) are bounded by an LRU of max_plans entries so that a stream of distinct
// batch sizes cannot grow device memory without limit.  Caller holds m->mu.
int get_plan(hrp_model* m, int B, int replica, Plan** out) {
  auto it = m->plans.find(std::make_pair(B, replica));
  if (it != m->plans.end()) {
    it->second->last_use = ++m->use_counter;
    *out = it->second.get();
    return HRP_OK;
  }
  while ((int)m->plans.size() >= m->max_plans) {
    auto victim = m->plans.end();
    for (auto jt = m->plans.begin(); jt != m->plans.end(); ++jt) {
      if (jt->first.first == m->desc.chunk) continue;  // nominal plans are never evicted
      if (victim == m->plans.end() || jt->second->last_use < victim->second->last_use) victim = jt;
    }
    if (victim == m->plans.end()) break;
    if (m->last_plan == victim->second.get()) m->last_plan = nullptr;
    m->plans.erase(victim);  // ~Plan waits for the plan's last user (busy_ev) before freeing its buffers
  }
  int rc = build_plan(m, B, out, replica);
  if (rc == HRP_OK) (*out)->last_use = ++m->use_counter;
  return rc;
}

// Replica used by a single-chunk forward enqueued on `user`: callers that drive the model from several streams (e.g. two
// batches in flight, bench.py's strong-scaling mode) get distinct plan replicas -- up to desc.inflight of them --, so
// their forwards overlap instead of serialising on one set of buffers.  Caller holds m->mu.
int replica_of_stream(hrp_model* m, cudaStream_t user) {
  if (m->desc.inflight <= 1) return 0;
  auto it = m->stream_slot.find(user);
  if (it != m->stream_slot.end()) return it->second;
  const int slot = m->next_slot;
  m->next_slot = (m->next_slot + 1) % m->desc.inflight;
  m->stream_slot[user] = slot;
  return slot;
}

// Serialise the users of one plan: wait for the previous forward on this plan (any stream), and mark this one.
int plan_acquire(Plan* pl, cudaStream_t s) {
  if (pl->busy_ev == nullptr) HRP_CUDA_CHECK(cudaEventCreateWithFlags(&pl->busy_ev, cudaEventDisableTiming));
  if (pl->busy_recorded) HRP_CUDA_CHECK(cudaStreamWaitEvent(s, pl->busy_ev, 0));
  return HRP_OK;
}
int plan_release(Plan* pl, cudaStream_t s) {
  HRP_CUDA_CHECK(cudaEventRecord(pl->busy_ev, s));
  pl->busy_recorded = true;
  return HRP_OK;
}

int run_plan_body(hrp_model* m, Plan* pl, cudaStream_t s) {
  if (m->use_graph) {
    HRP_CUDA_CHECK(cudaGraphLaunch(pl->exec, s));
    count_launch(pl->n_kernels);
    return HRP_OK;
  }
  // eager path (HRP_NO_GRAPH=1): one stream, optional programmatic dependent launches between kernels
  const char* pdl_env = getenv("HRP_PDL");
  const bool pdl_on = (pdl_env != nullptr && pdl_env[0] == '1') && !m->use_simt;
  bool prev_kernel = false;
  for (auto& op : pl->ops) {
    PdlScope pdl(pdl_on && op.kind != OP_MEMSET && prev_kernel);
    int rc = launch_op(m, op, s);
    if (rc != HRP_OK) return rc;
    prev_kernel = (op.kind != OP_MEMSET);
  }
  return HRP_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------
extern "C" {

int hrp_model_create(const hrp_model_desc* desc, hrp_model** out) {
  HRP_REQUIRE(desc != nullptr && out != nullptr, "null argument");
  HRP_REQUIRE(desc->kind == HRP_MODEL_FULL || desc->kind == HRP_MODEL_DEPTHNET, "unknown model kind");
  if (desc->kind == HRP_MODEL_FULL) {
    HRP_REQUIRE(desc->nkpt > 0 && desc->nkpt <= kMaxKpt && desc->dof > 0 && desc->dof <= kMaxDof, "bad nkpt / dof");
    HRP_REQUIRE(desc->ref_kpt >= 0 && desc->ref_kpt < desc->nkpt, "reference keypoint out of range");
    HRP_REQUIRE(desc->n_iter >= 0 && desc->n_iter <= 16, "n_iter out of range");
  }
  HRP_REQUIRE(desc->chunk > 0 && desc->chunk <= 1024, "chunk must be in 1..1024");
  HRP_REQUIRE(desc->inflight >= 1 && desc->inflight <= 4, "inflight must be in 1..4");
  hrp_model* m = new (std::nothrow) hrp_model();
  HRP_REQUIRE(m != nullptr, "out of host memory");
  m->desc = *desc;
  m->prefix_root = (desc->kind == HRP_MODEL_FULL) ? "rootnet_backbone." : "backbone.";
  cudaGetDevice(&m->device);
  const char* e = getenv("HRP_NO_GRAPH");
  m->use_graph = !(e != nullptr && e[0] == '1');
  e = getenv("HRP_CONV_IMPL");
  m->use_simt = (e != nullptr && std::string(e) == "simt");
  e = getenv("HRP_AUTOTUNE");
  m->autotune = (e != nullptr && e[0] == '1');
  m->pin_variants = getenv("HRP_CONV_PERSISTENT") != nullptr || getenv("HRP_CONV_VARIANT") != nullptr;
  e = getenv("HRP_FUSE0_EPI");
  m->fuse0_epilogue = (e != nullptr && e[0] == '1');
  e = getenv("HRP_HEAD_FOLD");
  m->head_fold = !(e != nullptr && e[0] == '0');
  e = getenv("HRP_MAX_PLANS");
  if (e != nullptr && atoi(e) >= 1) m->max_plans = atoi(e);
  *out = m;
  return HRP_OK;
}

void hrp_model_destroy(hrp_model* model) { delete model; }

int hrp_model_set_tensor(hrp_model* model, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
  HRP_REQUIRE(model != nullptr && name != nullptr && data != nullptr && ndim >= 0 && ndim <= 8, "bad argument");
  HRP_REQUIRE(shape != nullptr || ndim == 0, "null shape");
  if (model->finalized) {
    set_error("hrp_model_set_tensor after hrp_model_finalize");
    return HRP_ERR_STATE;
  }
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    HRP_REQUIRE(shape[i] > 0, "non-positive dimension");
    t.shape.push_back(shape[i]);
    n *= (size_t)shape[i];
  }
  t.data.assign(data, data + n);
  model->host[name] = std::move(t);
  return HRP_OK;
}

int hrp_model_set_robot(hrp_model* model, const hrp_robot* robot) {
  HRP_REQUIRE(model != nullptr && robot != nullptr, "null argument");
  int32_t nkpt = 0, dof = 0;
  hrp_robot_dims(robot, &nkpt, &dof);
  HRP_REQUIRE(nkpt == model->desc.nkpt && dof == model->desc.dof, "robot table does not match the model (nkpt / dof)");
  model->robot = robot;
  return HRP_OK;
}

int hrp_model_finalize(hrp_model* m) {
  HRP_REQUIRE(m != nullptr, "null model");
  std::lock_guard<std::mutex> lock(m->mu);
  if (m->finalized) return HRP_OK;
  HRP_CUDA_CHECK(cudaSetDevice(m->device));
  const HostTensor* dw = find(m, "depth_layer.weight");
  const HostTensor* db = find(m, "depth_layer.bias");
  HRP_NEED(dw, "depth_layer.weight");
  HRP_NEED(db, "depth_layer.bias");
  HRP_REQUIRE(dw->data.size() == 2048 && db->data.size() == 1, "depth_layer must map 2048 -> 1 (multi_kp is off-path)");
  int rc = dev_upload(m, dw->data.data(), 2048 * 4, reinterpret_cast<void**>(&m->depth_w));
  if (rc != HRP_OK) return rc;
  m->depth_b = db->data[0];
  if (m->desc.kind == HRP_MODEL_FULL) {
    if (m->robot == nullptr) {
      set_error("hrp_model_set_robot must be called before hrp_model_finalize");
      return HRP_ERR_STATE;
    }
    rc = build_regressor(m, "fc_pose_1", "fc_pose_2", "decpose", m->desc.dof, &m->reg_pose);
    if (rc != HRP_OK) return rc;
    rc = build_regressor(m, "fc_rot_1", "fc_rot_2", "decrot", 6, &m->reg_rot);
    if (rc != HRP_OK) return rc;
    const HostTensor *ip = find(m, "init_pose"), *ir = find(m, "init_rot");
    HRP_NEED(ip, "init_pose");
    HRP_NEED(ir, "init_rot");
    HRP_REQUIRE((int)ip->data.size() == m->desc.dof && ir->data.size() == 6, "init_pose / init_rot size mismatch");
    rc = dev_upload(m, ip->data.data(), ip->data.size() * 4, reinterpret_cast<void**>(&m->init_pose));
    if (rc != HRP_OK) return rc;
    rc = dev_upload(m, ir->data.data(), 24, reinterpret_cast<void**>(&m->init_rot));
    if (rc != HRP_OK) return rc;
  }
  HRP_CUDA_CHECK(cudaEventCreateWithFlags(&m->fork_ev, cudaEventDisableTiming));
  conv_init();  // function attributes must be set outside of stream capture
  // build the plan(s) of the nominal chunk now: packs every weight once and validates the state dict
  for (int r = 0; r < m->desc.inflight; ++r) {
    Plan* pl = nullptr;
    rc = get_plan(m, m->desc.chunk, r, &pl);
    if (rc != HRP_OK) return rc;
  }
  m->host.clear();  // packed copies live on the device now
  m->finalized = true;
  return HRP_OK;
}

static int forward_impl(hrp_model* m, const void* x_reg_v, const void* x_root_v, int in_u8, const float* k_value,
                        const float* K, const float* init_pose, const float* init_rot, int32_t B,
                        const hrp_outputs* out, float* depth_mm, cudaStream_t user) {
  const float* x_reg = reinterpret_cast<const float*>(x_reg_v);
  const float* x_root = reinterpret_cast<const float*>(x_root_v);
  const uint8_t* x_reg8 = reinterpret_cast<const uint8_t*>(x_reg_v);
  const uint8_t* x_root8 = reinterpret_cast<const uint8_t*>(x_root_v);
  if (!m->finalized) {
    set_error("forward before hrp_model_finalize");
    return HRP_ERR_STATE;
  }
  std::lock_guard<std::mutex> lock(m->mu);  // plans, workspaces and graph launches of one handle are serialised
  HRP_CUDA_CHECK(cudaSetDevice(m->device));
  const bool full = (m->desc.kind == HRP_MODEL_FULL);
  const int chunk = m->desc.chunk;
  const int nchunks = (B + chunk - 1) / chunk;
  const bool multi = (m->desc.inflight > 1 && nchunks > 1);
  const size_t img = (size_t)3 * kImg * kImg;
  if (multi) HRP_CUDA_CHECK(cudaEventRecord(m->fork_ev, user));
  std::vector<Plan*> used;
  for (int c = 0; c < nchunks; ++c) {
    const int b0 = c * chunk;
    const int nb = std::min(chunk, B - b0);
    const int replica = multi ? (c % m->desc.inflight) : replica_of_stream(m, user);
    Plan* pl = nullptr;
    int rc = get_plan(m, nb, replica, &pl);
    if (rc != HRP_OK) return rc;
    m->last_plan = pl;
    cudaStream_t s = multi ? pl->stream : user;
    if (multi && std::find(used.begin(), used.end(), pl) == used.end()) {
      HRP_CUDA_CHECK(cudaStreamWaitEvent(s, m->fork_ev, 0));
      used.push_back(pl);
    }
    rc = plan_acquire(pl, s);
    if (rc != HRP_OK) return rc;
    rc = in_u8 ? launch_pack_input_s2d_u8(x_root8 + b0 * img, pl->s2d_root, nb, kImg, kImg, s, kS2dPitch, kS2dPad)
               : launch_pack_input_s2d(x_root + b0 * img, pl->s2d_root, nb, kImg, kImg, s, kS2dPitch, kS2dPad);
    if (rc != HRP_OK) return rc;
    if (full) {
      rc = in_u8 ? launch_pack_input_s2d_u8(x_reg8 + b0 * img, pl->s2d_reg, nb, kImg, kImg, s, kS2dPitch, kS2dPad)
                 : launch_pack_input_s2d(x_reg + b0 * img, pl->s2d_reg, nb, kImg, kImg, s, kS2dPitch, kS2dPad);
      if (rc != HRP_OK) return rc;
    }
    rc = run_plan_body(m, pl, s);
    if (rc != HRP_OK) return rc;
    if (!full) {
      rc = launch_depth(pl->feat, m->depth_w, m->depth_b, k_value + b0, depth_mm + b0, nb, 1.0f, s);
      if (rc != HRP_OK) return rc;
      rc = plan_release(pl, s);
      if (rc != HRP_OK) return rc;
      continue;
    }
    const int nk = m->desc.nkpt, dof = m->desc.dof;
    HeadParams hp;
    memset(&hp, 0, sizeof(hp));
    hp.B = nb;
    hp.nkpt = nk;
    hp.ref_kpt = m->desc.ref_kpt;
    hp.fix_root = m->desc.fix_root;
    hp.image_size = m->desc.image_size;
    hp.depth_factor = m->desc.depth_factor;
    const bool folded = (pl->fold_partials != nullptr);
    if (folded) {
      hp.partials = pl->fold_partials;
      hp.merged = pl->fold_merged;
      hp.chunks = pl->fold_chunks;
    } else {
      hp.heatmap = pl->acts[pl->heatmap].ptr;
      hp.chunks = pl->head_chunks;
      hp.counters = reinterpret_cast<unsigned int*>(pl->head_ws);
      hp.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(pl->head_ws) + ((size_t)nb * 4 + 255) / 256 * 256);
    }
    hp.K = K + (size_t)b0 * 9;
    hp.feat = pl->feat;
    hp.depth_w = m->depth_w;
    hp.depth_b = m->depth_b;
    hp.k_value = k_value + b0;
    hp.xf = pl->xf;
    hp.init_pose = (init_pose != nullptr) ? init_pose + (size_t)b0 * dof : m->init_pose;
    hp.init_rot = (init_rot != nullptr) ? init_rot + (size_t)b0 * 6 : m->init_rot;
    hp.init_batched = (init_pose != nullptr) ? 1 : 0;
    HRP_REQUIRE((init_pose == nullptr) == (init_rot == nullptr), "init_pose and init_rot must be given together");
    hp.robot = reinterpret_cast<const RobotTable*>(hrp_robot_device_table(m->robot));
    hp.reg_pose = m->reg_pose;
    hp.reg_rot = m->reg_rot;
    hp.pose = out->pose ? out->pose + (size_t)b0 * dof : nullptr;
    hp.rot = out->rot ? out->rot + (size_t)b0 * 6 : nullptr;
    hp.trans = out->trans ? out->trans + (size_t)b0 * 3 : nullptr;
    hp.root_uv = out->root_uv ? out->root_uv + (size_t)b0 * 2 : nullptr;
    hp.depth = out->depth ? out->depth + b0 : nullptr;
    hp.uvd = out->uvd ? out->uvd + (size_t)b0 * nk * 3 : nullptr;
    hp.xyz_int = out->xyz_int ? out->xyz_int + (size_t)b0 * nk * 3 : nullptr;
    hp.xyz_fk = out->xyz_fk ? out->xyz_fk + (size_t)b0 * nk * 3 : nullptr;
    hp.uv_int = out->uv_int ? out->uv_int + (size_t)b0 * nk * 2 : nullptr;
    hp.uv_fk = out->uv_fk ? out->uv_fk + (size_t)b0 * nk * 2 : nullptr;
    rc = folded ? launch_head_from_partials(hp, s) : launch_head(hp, s);
    if (rc != HRP_OK) return rc;
    rc = plan_release(pl, s);
    if (rc != HRP_OK) return rc;
  }
  if (multi)
    for (Plan* pl : used) {
      HRP_CUDA_CHECK(cudaEventRecord(pl->done_ev, pl->stream));
      HRP_CUDA_CHECK(cudaStreamWaitEvent(user, pl->done_ev, 0));
    }
  return HRP_OK;
}

int hrp_model_forward(hrp_model* model, const float* x_reg, const float* x_root, const float* k_value, const float* K,
                      const float* init_pose, const float* init_rot, int32_t B, const hrp_outputs* out, void* stream) {
  HRP_REQUIRE(model != nullptr && x_reg != nullptr && x_root != nullptr && k_value != nullptr && K != nullptr &&
                  out != nullptr && B > 0,
              "bad argument");
  HRP_REQUIRE(model->desc.kind == HRP_MODEL_FULL, "not a full model handle");
  return forward_impl(model, x_reg, x_root, 0, k_value, K, init_pose, init_rot, B, out, nullptr,
                      reinterpret_cast<cudaStream_t>(stream));
}

int hrp_model_forward_u8(hrp_model* model, const uint8_t* x_reg, const uint8_t* x_root, const float* k_value,
                         const float* K, const float* init_pose, const float* init_rot, int32_t B,
                         const hrp_outputs* out, void* stream) {
  HRP_REQUIRE(model != nullptr && x_reg != nullptr && x_root != nullptr && k_value != nullptr && K != nullptr &&
                  out != nullptr && B > 0,
              "bad argument");
  HRP_REQUIRE(model->desc.kind == HRP_MODEL_FULL, "not a full model handle");
  return forward_impl(model, x_reg, x_root, 1, k_value, K, init_pose, init_rot, B, out, nullptr,
                      reinterpret_cast<cudaStream_t>(stream));
}

int hrp_model_depthnet_forward(hrp_model* model, const float* x, const float* k_value, int32_t B, float* depth_mm,
                               void* stream) {
  HRP_REQUIRE(model != nullptr && x != nullptr && k_value != nullptr && depth_mm != nullptr && B > 0, "bad argument");
  HRP_REQUIRE(model->desc.kind == HRP_MODEL_DEPTHNET, "not a depthnet handle");
  return forward_impl(model, nullptr, x, 0, k_value, nullptr, nullptr, nullptr, B, nullptr, depth_mm,
                      reinterpret_cast<cudaStream_t>(stream));
}

int hrp_model_activation(hrp_model* model, const char* name, const void** ptr, int32_t* B, int32_t* H, int32_t* W,
                         int32_t* C) {
  HRP_REQUIRE(model != nullptr && name != nullptr && ptr != nullptr, "bad argument");
  if (model->last_plan == nullptr) {
    set_error("no forward has run yet");
    return HRP_ERR_STATE;
  }
  Plan* pl = model->last_plan;
  if (std::string(name) == "feat" || std::string(name) == "xf") {
    *ptr = (std::string(name) == "feat") ? (const void*)pl->feat : (const void*)pl->xf;
    *B = pl->B;
    *H = *W = 1;
    *C = -2048;  // negative: fp32 vector, not a bf16 NHWC tensor
    return HRP_OK;
  }
  auto it = pl->taps.find(name);
  if (it == pl->taps.end()) {
    set_error(std::string("unknown activation tap: ") + name);
    return HRP_ERR_INVALID;
  }
  const Act& a = pl->acts[it->second];
  *ptr = a.ptr;
  *B = pl->B;
  *H = a.H;
  *W = a.W;
  *C = a.C;
  return HRP_OK;
}

int hrp_model_profile(hrp_model* m, int32_t batch, int32_t iters, char* buf, int64_t buflen) {
  HRP_REQUIRE(m != nullptr && buf != nullptr && buflen > 0 && iters > 0, "bad argument");
  if (!m->finalized) {
    set_error("profile before finalize");
    return HRP_ERR_STATE;
  }
  std::lock_guard<std::mutex> lock(m->mu);
  Plan* pl = nullptr;
  int rc = get_plan(m, batch, 0, &pl);
  if (rc != HRP_OK) return rc;
  cudaStream_t s = pl->stream;
  rc = plan_acquire(pl, s);
  if (rc != HRP_OK) return rc;
  cudaEvent_t e0, e1;
  HRP_CUDA_CHECK(cudaEventCreate(&e0));
  HRP_CUDA_CHECK(cudaEventCreate(&e1));
  std::string out;
  char line[512];
  for (auto& op : pl->ops) {
    if (op.kind == OP_MEMSET) continue;
    rc = launch_op(m, op, s);  // warm
    if (rc != HRP_OK) return rc;
    HRP_CUDA_CHECK(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; ++i) {
      rc = launch_op(m, op, s);
      if (rc != HRP_OK) return rc;
    }
    HRP_CUDA_CHECK(cudaEventRecord(e1, s));
    HRP_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    HRP_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = 1e3 * ms / iters;
    if (op.kind == OP_CONV) {
      const ConvParams& q = op.conv.p;
      const double in_b = (double)q.B * q.Hin * q.Win * std::min(q.Cin, q.src_pix) * 2.0, out_b = (double)q.B * q.Hout * q.Wout * q.Cout * 2.0;
      // algorithmic bytes of the layer: input + output + every addend read by the epilogue + packed weights
      double add_b = 0.0;
      for (int a = 0; a < 3; ++a) {
        if (q.pre[a] != nullptr) add_b += out_b;
        if (q.up[a] != nullptr) add_b += out_b / (double)(1 << (2 * q.up_shift[a]));
      }
      if (q.post != nullptr) add_b += out_b;
      const double w_b = (double)q.nphase * q.cout_pad * q.ktot * 2.0;
      const double all_b = in_b + (q.out != nullptr ? out_b : 0.0) + add_b + w_b;
      snprintf(line, sizeof(line), "%s\tconv\t%d\t%dx%d\t%d\t%d\t%dx%d\t%d\t%d\t%d\t%s\t%d\t%.2f\t%.1f\t%.1f\t%.0f\t%.0f\n",
               op.name.c_str(), op.lane, q.Hin, q.Win, q.Cin, q.Cout, q.Hout, q.Wout, q.ntaps, q.n_tile, op.conv.epi,
               op.conv.halo ? "halo" : (op.conv.persistent ? "persist" : "tile"),
               op.conv.halo ? op.conv.hp.T : (op.conv.persistent ? op.conv.pcfg.stages : op.conv.stages), us,
               op.conv.flops / us * 1e-6, all_b / us * 1e-3, op.conv.flops, all_b);
    } else {
      snprintf(line, sizeof(line), "%s\tmisc\t%d\t-\t-\t-\t-\t-\t-\t-\t-\t-\t%.2f\t0\t0\t0\t0\n", op.name.c_str(), op.lane, us);
    }
    out += line;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  rc = plan_release(pl, s);
  if (rc != HRP_OK) return rc;
  if ((int64_t)out.size() + 1 > buflen) {
    set_error("profile buffer too small: need " + std::to_string(out.size() + 1));
    return HRP_ERR_INVALID;
  }
  memcpy(buf, out.c_str(), out.size() + 1);
  return HRP_OK;
}

int hrp_model_plan_memory(hrp_model* m, int32_t batch, int32_t alias, int64_t* arena_bytes, int64_t* tensor_bytes,
                          int32_t* n_tensors, int32_t* n_ops) {
  HRP_REQUIRE(m != nullptr && batch > 0 && arena_bytes != nullptr, "bad argument");
  std::lock_guard<std::mutex> lock(m->mu);
  Plan dry;
  dry.B = batch;
  dry.fold_chunks = 256;
  dry.s2d_root = dry.s2d_reg = reinterpret_cast<bf16*>(16);
  Builder bd{m, &dry};
  bd.dry = true;
  emit_program(bd, m, &dry, m->desc.kind == HRP_MODEL_FULL && m->head_fold && !m->use_simt);
  if (bd.rc != HRP_OK) return bd.rc;
  std::vector<size_t> offsets;
  size_t sum = 0;
  const size_t arena = plan_arena(dry, alias != 0, &offsets, &sum);
  if (arena == 0) {
    set_error("internal: activation arena planning produced overlapping live tensors");
    return HRP_ERR_STATE;
  }
  *arena_bytes = (int64_t)arena;
  if (tensor_bytes) *tensor_bytes = (int64_t)sum;
  if (n_tensors) *n_tensors = (int32_t)dry.acts.size();
  if (n_ops) *n_ops = (int32_t)dry.ops.size();
  return HRP_OK;
}

int hrp_model_set_tuning(hrp_model* m, const char* text) {
  HRP_REQUIRE(m != nullptr && text != nullptr, "null argument");
  std::lock_guard<std::mutex> lock(m->mu);
  std::istringstream in(text);
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ls(line);
    std::string key;
    int variant = -1;
    if (!(ls >> key >> variant) || variant < 0 || variant > 2) {
      set_error("malformed tuning line: " + line);
      return HRP_ERR_INVALID;
    }
    m->tune_cache[key] = variant;
  }
  return HRP_OK;
}

int hrp_model_get_tuning(hrp_model* m, char* buf, int64_t buflen) {
  HRP_REQUIRE(m != nullptr && buf != nullptr && buflen > 0, "bad argument");
  std::lock_guard<std::mutex> lock(m->mu);
  std::string out;
  for (auto& kv : m->tune_cache) out += kv.first + " " + std::to_string(kv.second) + "\n";
  if ((int64_t)out.size() + 1 > buflen) {
    set_error("tuning buffer too small: need " + std::to_string(out.size() + 1));
    return HRP_ERR_INVALID;
  }
  memcpy(buf, out.c_str(), out.size() + 1);
  return HRP_OK;
}

int hrp_model_stats(const hrp_model* model, int32_t batch, double* flops, int32_t* kernels, int64_t* activation_bytes) {
  HRP_REQUIRE(model != nullptr, "null model");
  for (auto& kv : model->plans)
    if (kv.first.first == batch) {
      const Plan* pl = kv.second.get();
      if (flops) *flops = pl->flops;
      if (kernels) *kernels = pl->n_kernels + ((model->desc.kind == HRP_MODEL_FULL) ? 3 : 2);
      if (activation_bytes) *activation_bytes = (int64_t)pl->arena_bytes;  // (one-buffer-per-tensor would be act_bytes_sum)
      return HRP_OK;
    }
  set_error("no plan for this batch size yet");
  return HRP_ERR_STATE;
}

}  // extern "C"
