// C ABI of libhrp_b200.so (include/hrp.h): thin extern "C" forwarding layer, no exceptions cross it.
#include "../../include/hrp.h"

#include "conv.h"
#include "launch_count.h"
#include "ops.h"

#include <new>

namespace hrp {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }
std::atomic<int64_t> g_launch_count{0};
}  // namespace hrp

using namespace hrp;

struct hrp_conv {
  ConvLayerDesc desc;
  ConvPlan plan;
};

static ConvLayerDesc to_internal(const hrp_conv_desc* d) {
  ConvLayerDesc o;
  o.kind = d->kind;
  o.B = d->B;
  o.Hin = d->Hin;
  o.Win = d->Win;
  o.Cin = d->Cin;
  o.Cout = d->Cout;
  o.kh = d->kh;
  o.kw = d->kw;
  o.stride = d->stride;
  o.pad = d->pad;
  o.relu = d->relu;
  return o;
}

extern "C" {

const char* hrp_last_error(void) { return hrp::last_error(); }
const char* hrp_version(void) { return "hrp_b200 0.1 sm_100a"; }
int64_t hrp_launch_count(void) { return g_launch_count.load(); }

int hrp_conv_packed_weight_elems(const hrp_conv_desc* desc, int64_t* elems) {
  HRP_REQUIRE(desc != nullptr && elems != nullptr, "null argument");
  ConvParams p;
  int rc = conv_geometry(to_internal(desc), &p);
  if (rc != HRP_OK) return rc;
  *elems = (int64_t)conv_packed_weight_elems(p);
  return HRP_OK;
}

int hrp_conv_pack_weights(const hrp_conv_desc* desc, int32_t cin_ref, const float* w, uint16_t* out) {
  HRP_REQUIRE(desc != nullptr && w != nullptr && out != nullptr, "null argument");
  ConvParams p;
  ConvLayerDesc d = to_internal(desc);
  int rc = conv_geometry(d, &p);
  if (rc != HRP_OK) return rc;
  return conv_pack_weights(d, p, cin_ref, w, out);
}

int hrp_conv_create(const hrp_conv_desc* desc, const void* in_dev, const void* w_packed_dev,
                    const hrp_conv_epilogue* epi, hrp_conv** out) {
  HRP_REQUIRE(desc != nullptr && epi != nullptr && out != nullptr, "null argument");
  HRP_REQUIRE(epi->scale != nullptr && epi->bias != nullptr, "scale/bias are required");
  HRP_REQUIRE(epi->out != nullptr || epi->pool_out != nullptr, "an output is required");
  hrp_conv* c = new (std::nothrow) hrp_conv();
  HRP_REQUIRE(c != nullptr, "out of host memory");
  c->desc = to_internal(desc);
  int rc = conv_geometry(c->desc, &c->plan.p);
  if (rc == HRP_OK) {
    ConvParams& p = c->plan.p;
    p.scale = epi->scale;
    p.bias = epi->bias;
    for (int i = 0; i < 3; ++i) {
      p.pre[i] = reinterpret_cast<const bf16*>(epi->pre[i]);
      p.up[i] = reinterpret_cast<const bf16*>(epi->up[i]);
      p.up_shift[i] = epi->up_shift[i];
    }
    p.post = reinterpret_cast<const bf16*>(epi->post);
    p.out = reinterpret_cast<bf16*>(epi->out);
    p.pool_out = epi->pool_out;
    rc = conv_plan_finalize(&c->plan, reinterpret_cast<const bf16*>(in_dev),
                            reinterpret_cast<const bf16*>(w_packed_dev));
  }
  if (rc != HRP_OK) {
    delete c;
    return rc;
  }
  *out = c;
  return HRP_OK;
}

int hrp_conv_out_shape(const hrp_conv* conv, int32_t* Hout, int32_t* Wout) {
  HRP_REQUIRE(conv != nullptr && Hout != nullptr && Wout != nullptr, "null argument");
  *Hout = conv->plan.p.Hout;
  *Wout = conv->plan.p.Wout;
  return HRP_OK;
}

int hrp_conv_run(hrp_conv* conv, int32_t impl, void* stream) {
  HRP_REQUIRE(conv != nullptr, "null conv handle");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (impl == HRP_IMPL_TCGEN05) return conv_plan_launch(conv->plan, s);
  if (impl == HRP_IMPL_SIMT_CHECK) return conv_plan_launch_simt(conv->plan, s);
  set_error("unknown conv impl");
  return HRP_ERR_INVALID;
}

void hrp_conv_destroy(hrp_conv* conv) { delete conv; }

int hrp_pack_input_s2d(const float* x_nchw, void* out_s2d, int32_t B, int32_t H, int32_t W, void* stream) {
  HRP_REQUIRE(x_nchw != nullptr && out_s2d != nullptr && B > 0, "bad argument");
  return launch_pack_input_s2d(x_nchw, out_s2d, B, H, W, reinterpret_cast<cudaStream_t>(stream));
}
int hrp_maxpool3x3s2(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  HRP_REQUIRE(in != nullptr && out != nullptr && B > 0, "bad argument");
  return launch_maxpool3x3s2(in, out, B, H, W, C, reinterpret_cast<cudaStream_t>(stream));
}
int hrp_nchw_f32_to_nhwc_bf16(const float* in, void* out, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Cpad,
                              void* stream) {
  HRP_REQUIRE(in != nullptr && out != nullptr && B > 0 && Cpad >= C, "bad argument");
  return launch_nchw_f32_to_nhwc_bf16(in, out, B, C, H, W, Cpad, reinterpret_cast<cudaStream_t>(stream));
}
int hrp_nhwc_bf16_to_nchw_f32(const void* in, float* out, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Cpad,
                              void* stream) {
  HRP_REQUIRE(in != nullptr && out != nullptr && B > 0 && Cpad >= C, "bad argument");
  return launch_nhwc_bf16_to_nchw_f32(in, out, B, C, H, W, Cpad, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
