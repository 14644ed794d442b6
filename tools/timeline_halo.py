"""globaltimer stamps of the halo kernel over 3 back-to-back launches: python tools/timeline_halo.py C H B [res]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import _lib, ops

Cc, H, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
res = len(sys.argv) > 4 and sys.argv[4] == "res"
x = torch.randn(B, H, H, Cc, device="cuda").to(torch.bfloat16)
w = torch.randn(Cc, Cc, 3, 3) * 0.05
pre = (torch.randn(B, H, H, Cc, device="cuda").to(torch.bfloat16),) if res else ()
op = ops.ConvOp(x, w, stride=1, pad=1, relu=True, pre=pre)
_lib.check(_lib.lib().hrp_conv_set_variant(op.handle, C.c_int32(2)))
for _ in range(3):
    op.run()
torch.cuda.synchronize()
tls = [torch.zeros(8 * 16, dtype=torch.int64, device="cuda") for _ in range(3)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for tl in tls:
    _lib.check(_lib.lib().hrp_conv_set_timeline(op.handle, C.c_void_p(tl.data_ptr())))
    op.run()
e1.record()
torch.cuda.synchronize()
print(f"C={Cc} H={H} B={B} res={res}: {e0.elapsed_time(e1) / 3 * 1e3:.1f} us per launch (events)")
names = ["entry", "setup", "weights", "first band", "mma done", "stores issued", "-", "stores drained", "-", "exit"]
base = int(tls[0][0])
for i, tl in enumerate(tls):
    t = tl.cpu().view(8, 16)
    for cta in (0, 7):
        print(f"launch {i} cta {cta}: " + "  ".join(f"{n} {int(t[cta, e]) - base}" for e, n in enumerate(names)))
        print(f"      mma warp cycles: wait a_full {int(t[cta, 10])}  wait acc_empty {int(t[cta, 11])}  issue {int(t[cta, 12])}"
              f"  loop {int(t[cta, 13])}  tiles {int(t[cta, 14])}")
