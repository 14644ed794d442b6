"""Per-role timeline of the persistent conv kernel (clock64 stamps; hrp_conv_set_timeline) in steady state.
    python tools/timeline_persist.py <index into bench_conv.SHAPES | s2fuse> [B] [res] [skip=8]
Prints the stamps of 8 consecutive tiles of CTA 0 (cycles relative to the first one) and where each role waits."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import horopose_b200  # noqa
from horopose_b200 import _lib, ops
from bench_conv import SHAPES

which = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
with_res = "res" in sys.argv[3:]
skip = next((int(a.split("=")[1]) for a in sys.argv[3:] if a.startswith("skip=")), 8)
bf = lambda *s: torch.randn(*s, device="cuda").to(torch.bfloat16)
if which == "s2fuse":   # fuse_layers.1.0: 3x3 s2 32 -> 64, 64x64 -> 32x32, + x1 + up2 + up4
    name = "hr fuse 32->64 k3s2 @64 + x1 + up2 + up4"
    op = ops.ConvOp(bf(B, 64, 64, 32), torch.randn(64, 32, 3, 3) * 0.05, stride=2, pad=1, relu=True,
                    pre=[bf(B, 32, 32, 64)], up=[(bf(B, 16, 16, 64), 1), (bf(B, 8, 8, 64), 2)])
else:
    name, cin, h, cout, k, stride, pad, kind = SHAPES[int(which)]
    w = torch.randn(cin, cout, 4, 4) * 0.02 if kind == ops.DECONV_K4S2P1 else torch.randn(cout, cin, k, k) * 0.02
    res = ()
    if with_res:
        ho = (h + 2 * pad - k) // stride + 1
        res = (bf(B, ho, ho, cout),)
    op = ops.ConvOp(bf(B, h, h, cin), w, kind=kind, stride=stride, pad=pad, relu=True, pre=res)
L = _lib.lib()
_lib.check(L.hrp_conv_set_variant(op.handle, C.c_int32(1)))
buf = C.create_string_buffer(256)
_lib.check(L.hrp_conv_describe(op.handle, buf, 256))
for _ in range(5):
    op.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    op.run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
tl = torch.zeros(8 * 8 * 16 + 1, dtype=torch.int64, device="cuda")
tl[-1] = skip
_lib.check(L.hrp_conv_set_timeline(op.handle, C.c_void_p(tl.data_ptr())))
op.run()
torch.cuda.synchronize()
t = tl[:-1].cpu().view(8, 8, 16)
names = ["P.start", "P.empty0", "P.issued", "M.tmemfree", "M.full0", "M.lastcommit", "E.top", "E.tmemfull", "E.stagok",
         "E.done", "S.ready", "S.issued", "S.drained"]
print(f"== {name}{' + residual' if with_res else ''}  B={B}: {us:.1f} us per launch  [{buf.value.decode()}]  tiles {skip}..{skip + 7} of CTA 0")
v = t[0]
if not (v > 0).any():
    print("   (no stamps: the CTA has fewer tiles than skip + 1)")
    sys.exit(0)
t0 = int(v[v > 0].min())
print("tile " + " ".join(f"{n:>11s}" for n in names))
for tile in range(8):
    print(f"{tile + skip:4d} " + " ".join(f"{int(v[tile, e]) - t0 if v[tile, e] > 0 else -1:11d}" for e in range(13)))
ok = [i for i in range(8) if v[i, 5] > 0]
if len(ok) >= 2:
    d = lambda a, b: float(sum(int(v[i, a]) - int(v[i, b]) for i in ok if v[i, a] > 0 and v[i, b] > 0)) / len(ok)
    period = (int(v[ok[-1], 5]) - int(v[ok[0], 5])) / (len(ok) - 1)
    print(f"tile period {period:.0f} cycles;  producer: issue span {d(2, 0):.0f} (first slot wait {d(1, 0):.0f});  "
          f"MMA: wait first stage {d(4, 3):.0f}, issue span {d(5, 4):.0f};  epilogue: wait accumulator {d(7, 6):.0f}, "
          f"wait staging/residual {d(8, 7):.0f}, math+write {d(9, 8):.0f};  store: issue {d(11, 10):.0f}, drain {d(12, 11):.0f}")
