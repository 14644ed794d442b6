"""Time ONE conv shape with a chosen kernel variant: python tools/bench_one_conv.py C H B variant [res] [k]
Prints the minimum over 3 rounds of 40 back-to-back launches (CUDA events) and the planned launch configuration."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import _lib, ops

Cc, H, B, variant = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
res = len(sys.argv) > 5 and sys.argv[5] == "res"
k = int(sys.argv[6]) if len(sys.argv) > 6 else 3
x = torch.randn(B, H, H, Cc, device="cuda").to(torch.bfloat16)
w = torch.randn(Cc, Cc, k, k) * 0.05
pre = (torch.randn(B, H, H, Cc, device="cuda").to(torch.bfloat16),) if res else ()
op = ops.ConvOp(x, w, stride=1, pad=k // 2, relu=True, pre=pre)
_lib.check(_lib.lib().hrp_conv_set_variant(op.handle, C.c_int32(variant)))
buf = C.create_string_buffer(256)
_lib.check(_lib.lib().hrp_conv_describe(op.handle, buf, 256))
for _ in range(5):
    op.run()
torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        op.run()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 40)
print(f"C={Cc} H={H} B={B} res={int(res)}: {best * 1e3:6.1f} us  {2 * B * H * H * k * k * Cc * Cc / best / 1e9:6.1f} TFLOP/s  [{buf.value.decode()}]")
