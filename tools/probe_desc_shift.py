"""Hardware probe: row-shifted start addresses of swizzled K-major UMMA operands (see csrc/probe.cu, probe 2)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import _lib

L = _lib.lib()
rows = 384
for ck in (64, 32, 16):
    r = torch.arange(rows).view(-1, 1)
    k = torch.arange(ck).view(1, -1)
    A = (((r * 7 + k * 3) % 13) - 6).float()
    n = torch.arange(32).view(-1, 1)
    Bm = (((n * 5 + k) % 7) - 3).float()
    Ad, Bd = A.to(torch.bfloat16).cuda(), Bm.to(torch.bfloat16).cuda()
    out = torch.zeros(128, 32, device="cuda")
    for bo in (0, 1):
        res = []
        for shift in list(range(0, 18)) + [33, 34, 66, 67, 133, 200]:
            out.zero_()
            _lib.check(L.hrp_probe_desc_shift(ck, rows, shift, bo, C.c_void_p(Ad.data_ptr()), C.c_void_p(Bd.data_ptr()),
                                              C.c_void_p(out.data_ptr())))
            exp = A[shift:shift + 128] @ Bm.t()
            err = float((out.cpu() - exp).abs().max())
            res.append(f"{shift}:{'ok' if err == 0 else f'{err:.0f}'}")
        print(f"ck={ck} base_offset_mode={bo}: " + " ".join(res))
