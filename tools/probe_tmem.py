import ctypes as C, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import _lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
L = _lib.lib()
reps = 1024
print("warps wait_each cycles/ld(x32)  bytes/cycle (all warps)")
for warps in (1, 4, 8):
    for we in (0, 1):
        _lib.check(L.hrp_probe_tmem_ld_rate(warps, reps, we, C.c_void_p(out.data_ptr())))
        cyc = out.cpu().tolist()[0]
        print(warps, we, f"{cyc / reps:8.1f} {warps * reps * 4096 / cyc:8.1f}")
