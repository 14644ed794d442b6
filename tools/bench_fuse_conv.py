"""HRNet fuse-layer convs (generic epilogues: same-resolution + nearest-upsampled addends) at the bench batch: time per
kernel variant and, for the persistent kernel, the per-role timeline of the first tiles.
    python tools/bench_fuse_conv.py {up2|s2} B variant [timeline]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import _lib, ops

which, B, variant = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
bf = lambda *s: torch.randn(*s, device="cuda").to(torch.bfloat16)
if which == "up2":      # fuse_layers.0.1: 1x1 64 -> 32 at 32x32, output 64x64, + x0 + up4 + up8
    x = bf(B, 32, 32, 64)
    w = torch.randn(32, 64, 1, 1) * 0.1
    op = ops.ConvOp(x, w, kind=ops.CONV_UP2, relu=True, pre=[bf(B, 64, 64, 32)], up=[(bf(B, 16, 16, 32), 2), (bf(B, 8, 8, 32), 3)])
    io = B * (32 * 32 * 64 + 2 * 64 * 64 * 32) * 2
elif which == "up2pre":  # stage 2: only x0
    x = bf(B, 32, 32, 64)
    w = torch.randn(32, 64, 1, 1) * 0.1
    op = ops.ConvOp(x, w, kind=ops.CONV_UP2, relu=True, pre=[bf(B, 64, 64, 32)])
    io = B * (32 * 32 * 64 + 2 * 64 * 64 * 32) * 2
else:                   # fuse_layers.1.0: 3x3 s2 32 -> 64, 64x64 -> 32x32, + x1 + up2 + up4
    x = bf(B, 64, 64, 32)
    w = torch.randn(64, 32, 3, 3) * 0.05
    op = ops.ConvOp(x, w, stride=2, pad=1, relu=True, pre=[bf(B, 32, 32, 64)], up=[(bf(B, 16, 16, 64), 1), (bf(B, 8, 8, 64), 2)])
    io = B * (64 * 64 * 32 + 2 * 32 * 32 * 64) * 2
L = _lib.lib()
_lib.check(L.hrp_conv_set_variant(op.handle, C.c_int32(variant)))
buf = C.create_string_buffer(256)
_lib.check(L.hrp_conv_describe(op.handle, buf, 256))
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
print(f"{which} B={B}: {best * 1e3:7.1f} us  {io / best / 1e6:7.1f} GB/s  [{buf.value.decode()}]")
if len(sys.argv) > 4 and variant == 1:
    tl = torch.zeros(8 * 8 * 16 + 1, dtype=torch.int64, device="cuda")
    _lib.check(L.hrp_conv_set_timeline(op.handle, C.c_void_p(tl.data_ptr())))
    op.run()
    torch.cuda.synchronize()
    t = tl[:-1].cpu().view(8, 8, 16)
    names = ["P.start", "P.empty0", "P.issued", "M.tmemfree", "M.full0", "M.lastcommit", "E.top", "E.tmemfull", "E.stagok",
             "E.done", "S.ready", "S.issued", "S.drained"]
    t0 = int(t[0, 0][t[0, 0] > 0].min())
    print("tile " + " ".join(f"{n:>11s}" for n in names))
    for tile in range(8):
        print(f"{tile:4d} " + " ".join(f"{int(t[0, tile, e]) - t0 if t[0, tile, e] > 0 else -1:11d}" for e in range(13)))
