"""Micro-benchmark of the tcgen05 conv kernel on the dominant layer shapes."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import _lib, ops
import ctypes as C

VARIANTS = {0: 'tile', 1: 'persist', 2: 'halo'}
RES = len(sys.argv) > 2 and sys.argv[2] == 'res'

SHAPES = [
    # name, Cin, H, Cout, k, stride, pad, kind
    ("hr 32->32 k3 @64", 32, 64, 32, 3, 1, 1, ops.CONV),
    ("hr 64->64 k3 @32", 64, 32, 64, 3, 1, 1, ops.CONV),
    ("hr 128->128 k3 @16", 128, 16, 128, 3, 1, 1, ops.CONV),
    ("hr 256->256 k3 @8", 256, 8, 256, 3, 1, 1, ops.CONV),
    ("rn 64->256 k1 @64", 64, 64, 256, 1, 1, 0, ops.CONV),
    ("rn 256->64 k1 @64", 256, 64, 64, 1, 1, 0, ops.CONV),
    ("rn 64->64 k3 @64", 64, 64, 64, 3, 1, 1, ops.CONV),
    ("rn 128->128 k3 @32", 128, 32, 128, 3, 1, 1, ops.CONV),
    ("rn 256->256 k3 @16", 256, 16, 256, 3, 1, 1, ops.CONV),
    ("rn 256->1024 k1 @16", 256, 16, 1024, 1, 1, 0, ops.CONV),
    ("rn 1024->256 k1 @16", 1024, 16, 256, 1, 1, 0, ops.CONV),
    ("rn 512->512 k3 @8", 512, 8, 512, 3, 1, 1, ops.CONV),
    ("rn 512->2048 k1 @8", 512, 8, 2048, 1, 1, 0, ops.CONV),
    ("deconv 2048->256 @8", 2048, 8, 256, 4, 2, 1, ops.DECONV_K4S2P1),
    ("deconv 256->256 @32", 256, 32, 256, 4, 2, 1, ops.DECONV_K4S2P1),
    ("final 256->512 k1 @64", 256, 64, 512, 1, 1, 0, ops.CONV),
    ("hr 32->64 k3s2 @64", 32, 64, 64, 3, 2, 1, ops.CONV),
    ("hr 32->32 k3s2 @64", 32, 64, 32, 3, 2, 1, ops.CONV),
    ("hr 64->128 k3s2 @32", 64, 32, 128, 3, 2, 1, ops.CONV),
    ("rn 128->128 k3s2 @64", 128, 64, 128, 3, 2, 1, ops.CONV),
]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    print(f"B={B}")
    for name, cin, h, cout, k, stride, pad, kind in SHAPES:
        x = torch.randn(B, h, h, cin, device="cuda").to(torch.bfloat16)
        if kind == ops.DECONV_K4S2P1:
            w = torch.randn(cin, cout, 4, 4) * 0.02
            macs = B * h * h * 4 * cin * cout  # 4 taps per output pixel x 4 phases / ... = (2h)^2 * 4 * cin * cout / 4
            macs = B * (2 * h) ** 2 * 4 * cin * cout
        else:
            w = torch.randn(cout, cin, k, k) * 0.02
            ho = (h + 2 * pad - k) // stride + 1
            macs = B * ho * ho * k * k * cin * cout
        res = ()
        if RES and kind == ops.CONV and stride == 1:
            ho = (h + 2 * pad - k) // stride + 1
            res = (torch.randn(B, ho, ho, cout, device="cuda").to(torch.bfloat16),)
        op = ops.ConvOp(x, w, kind=kind, stride=stride, pad=pad, relu=True, pre=res)
        L = _lib.lib()
        line = f"{name:26s}"
        for variant in (0, 1, 2):
            if L.hrp_conv_set_variant(op.handle, C.c_int32(variant)) != 0:
                line += f"  {VARIANTS[variant]:8s}      n/a        "
                continue
            for _ in range(3):
                op.run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                op.run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            line += f"  {VARIANTS[variant]:8s}{ms*1e3:8.1f} us {2*macs/ms/1e9:7.1f} TF"
        in_b = x.numel() * 2
        out_b = op.out.numel() * 2 * (2 if res else 1)
        print(line + f"   io {(in_b + out_b) / 1e6:7.1f} MB")


def op_grid(op):
    return ""


if __name__ == "__main__":
    main()
