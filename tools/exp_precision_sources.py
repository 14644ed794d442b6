"""CPU study: where does the bf16 error of the HRNet root-depth feature come from?  Emulates the CUDA pipeline's rounding
points inside the oracle (fp32 accumulate): weights, conv inputs, block outputs (the residual stream)."""
import sys, time
sys.path.insert(0, "/root/repo")
import torch, torch.nn.functional as F
import horopose_b200
from horopose_b200 import synth
from oracle import horopose_oracle as O
torch.set_num_threads(8)
bf = lambda t: t.to(torch.bfloat16).float()
sd = synth.depthnet_state_dict()
_, x, k, _ = synth.inputs(16, seed=101, k_range=(500., 3000.))
MODE = {}
orig_conv, orig_bn, orig_basic, orig_bott = O._Ctx.conv, O._Ctx.bn, O._basic, O._bottleneck

def conv(self, xx, name, stride=1, pad=0):
    w = self.sd[self.prefix + name + ".weight"]
    b = self.sd.get(self.prefix + name + ".bias")
    if MODE.get("w"): w = bf(w)
    if MODE.get("a_in"): xx = bf(xx)       # conv operand rounded on the fly
    return F.conv2d(xx, w, b, stride, pad)
O._Ctx.conv = conv
relu_orig = F.relu
def basic(c, xx, name):
    out = relu_orig(c.bn(c.conv(xx, name + ".conv1", 1, 1), name + ".bn1"))
    if MODE.get("inner"): out = bf(out)
    out = c.bn(c.conv(out, name + ".conv2", 1, 1), name + ".bn2")
    y = relu_orig(out + xx)
    if MODE.get("stream"): y = bf(y)
    if MODE.get("stream16"): y = y.half().float()
    return y
O._basic = basic
def run(**m):
    MODE.clear(); MODE.update(m)
    with torch.no_grad():
        return O.depthnet_forward(sd, x, k)
t0 = time.time()
ref = run()
print("ref", time.time() - t0, "s; depth mm range", float(ref.min()), float(ref.max()))
for name, m in [("weights only", dict(w=1)), ("conv inputs only (stream fp32)", dict(a_in=1)),
                ("basic-block inner act only", dict(inner=1)), ("basic-block stream only", dict(stream=1)),
                ("stream fp16 only", dict(stream16=1)),
                ("w + a_in (stream fp32 stored)", dict(w=1, a_in=1)),
                ("w + a_in + inner + stream (full bf16 emulation of basic blocks)", dict(w=1, a_in=1, inner=1, stream=1)),
                ("w + a_in + inner + stream16", dict(w=1, a_in=1, inner=1, stream16=1))]:
    out = run(**m)
    e = (out - ref).abs()
    print(f"{name:70s} max {float(e.max()):.3f} mm  rms {float((e**2).mean().sqrt()):.3f} mm")
