"""Run ONE conv shape a few times (for ncu): python tools/prof_conv.py <index into bench_conv.SHAPES> [B] [res] [variant=N]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import horopose_b200  # noqa
from horopose_b200 import ops
from bench_conv import SHAPES

idx = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
name, cin, h, cout, k, stride, pad, kind = SHAPES[idx]
x = torch.randn(B, h, h, cin, device="cuda").to(torch.bfloat16)
w = torch.randn(cin, cout, 4, 4) * 0.02 if kind == ops.DECONV_K4S2P1 else torch.randn(cout, cin, k, k) * 0.02
res = torch.randn(B, h // stride if kind == ops.CONV else 2 * h, h // stride if kind == ops.CONV else 2 * h, cout,
                  device="cuda").to(torch.bfloat16) if "res" in sys.argv[3:] else None
op = ops.ConvOp(x, w, kind=kind, stride=stride, pad=pad, relu=True, pre=[res] if res is not None else [])
for a in sys.argv[3:]:
    if a.startswith("variant="):
        import ctypes as C
        from horopose_b200 import _lib
        _lib.check(_lib.lib().hrp_conv_set_variant(op.handle, C.c_int32(int(a.split("=")[1]))))
for _ in range(4):
    op.run()
torch.cuda.synchronize()
print("done", name)
