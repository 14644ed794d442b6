"""Pipeline timeline of the persistent conv kernel (clock64 stamps per role) for one bench shape."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import horopose_b200  # noqa
from horopose_b200 import _lib, ops
from bench_conv import SHAPES

idx = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
name, cin, h, cout, k, stride, pad, kind = SHAPES[idx]
x = torch.randn(B, h, h, cin, device="cuda").to(torch.bfloat16)
w = torch.randn(cin, cout, 4, 4) * 0.02 if kind == ops.DECONV_K4S2P1 else torch.randn(cout, cin, k, k) * 0.02
res = ()
if len(sys.argv) > 3 and sys.argv[3] == "res":
    ho = (h + 2 * pad - k) // stride + 1
    res = (torch.randn(B, ho, ho, cout, device="cuda").to(torch.bfloat16),)
op = ops.ConvOp(x, w, kind=kind, stride=stride, pad=pad, relu=True, pre=res)
tl = torch.zeros(8 * 8 * 16, dtype=torch.int64, device="cuda")
for _ in range(3):
    op.run()
_lib.check(_lib.lib().hrp_conv_set_timeline(op.handle, C.c_void_p(tl.data_ptr())))
op.run()
torch.cuda.synchronize()
t = tl.cpu().view(8, 8, 16)
names = ["P.start", "P.empty0", "P.issued", "M.tmemfree", "M.full0", "M.lastcommit", "E.top", "E.tmemfull", "E.stagok",
         "E.done", "S.ready", "S.issued", "S.drained"]
print(name, "B", B)
for cta in (0,):
    t0 = int(t[cta, 0][t[cta, 0] > 0].min())
    print(f"CTA {cta}: cycles relative to first stamp")
    print("tile " + " ".join(f"{n:>11s}" for n in names))
    for tile in range(8):
        print(f"{tile:4d} " + " ".join(f"{int(t[cta, tile, e]) - t0 if t[cta, tile, e] > 0 else -1:11d}" for e in range(13)))
