import ctypes as C, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import _lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
L = _lib.lib()
reps = 2000
print("M N cycles/MMA(issue) cycles/MMA(complete) MACs/cycle")
for M in (128, 64):
    for N in (32, 64, 128, 192, 256):
        for nacc in (1, 4):
            if nacc * N > 512:
                continue
            _lib.check(L.hrp_probe_mma_rate(M, N, reps, 4 | (nacc << 8), C.c_void_p(out.data_ptr()), 1))
            a, b = out.cpu().tolist()
            print(M, N, f"nacc={nacc}", f"{a / reps:8.1f} {b / reps:8.1f} {M * N * 16 / (b / reps):8.0f}")
print("swizzle mode / A start shift (M=128): ck N shift cycles/MMA")
for ck in (64, 32, 16):
    for N in (32, 64, 128):
        for shift in (0, 1, 3, 67):
            _lib.check(L.hrp_probe_mma_rate(128, N, reps, 4 | (1 << 8) | (ck << 16) | (shift << 24), C.c_void_p(out.data_ptr()), 1))
            a, b = out.cpu().tolist()
            print(ck, N, shift, f"{b / reps:8.1f}")
_lib.check(L.hrp_probe_mma_rate(128, 256, reps, 4, C.c_void_p(out.data_ptr()), 148))
print("148 CTAs 128x256:", out.cpu().tolist()[1] / reps)
