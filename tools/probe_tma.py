"""TMA tiled-load rate per SM for the conv kernels' A-operand boxes (tools/probe/probe.cu, probe 4).
Build the library with HRP_BUILD_PROBES=1 first.  Prints bytes / cycle / SM for a matrix of (channels = row stride,
loads in flight, CTAs sharing a tile, tensor footprint, number of CTAs, box rank / rows)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import _lib

L = _lib.lib()
out = torch.zeros(8 + 96 * 4, dtype=torch.int64, device="cuda")
last_split = None


def rate(B, H, W, Cc, bw, bh, bn, depth, share, iters=2000, ctas=148, rows=0, producers=1):
    ctas = ctas // share * share
    x = torch.empty(B, H, W, Cc, dtype=torch.bfloat16, device="cuda").normal_()
    best = None
    for _ in range(3):
        _lib.check(L.hrp_probe_tma_rate(C.c_void_p(x.data_ptr()), B, H, W, Cc, bw, bh, bn, depth, iters, share, ctas, rows,
                                        producers, C.c_void_p(out.data_ptr())))
        cyc = int(out[0])
        if best is None or cyc < best:
            best = cyc
            n_mine = iters // (producers & 255)
            global last_split
            last_split = tuple(int(v) / n_mine for v in out[1:4])
    return (rows if rows else 128) * 128.0 * iters / best


DEPTHS = (1, 2, 4, 6)
print("TMA loads, SWIZZLE_128B, inner box = 64 bf16 channels (128 B), row stride = C x 2 B; bytes / cycle / SM "
      "(cycles per load at the deepest ring)")
print(f"{'tensor (B,H,W,C)':22s} {'MB':>6s} {'box':>13s} {'CTAs':>5s} {'thr':>3s}  " + " ".join(f"depth {d:<2d}" for d in DEPTHS))
CASES = [
    # B, H, W, C, bw, bh, bn, rows, ctas, producers | wait flavour << 8
    (32, 64, 64, 256, 64, 2, 1, 0, 148, 1),
    (32, 64, 64, 256, 64, 2, 1, 0, 148, 1 | 1 << 8),   # test_wait spin
    (32, 64, 64, 256, 64, 2, 1, 0, 148, 1 | 2 << 8),   # try_wait, 0 ns suspend hint
    (32, 64, 64, 256, 64, 2, 1, 0, 148, 2),
    (32, 64, 64, 256, 64, 2, 1, 0, 148, 2 | 1 << 8),
    (512, 64, 64, 256, 64, 2, 1, 0, 148, 1),           # DRAM
    (512, 64, 64, 256, 64, 2, 1, 0, 148, 1 | 1 << 8),
    (512, 64, 64, 256, 64, 2, 1, 0, 148, 2 | 1 << 8),
]
for (B, H, W, Cc, bw, bh, bn, rows, ctas, prod) in CASES:
    mb = B * H * W * Cc * 2 / 1e6
    rs = [rate(B, H, W, Cc, bw, bh, bn, d, 1, ctas=ctas, rows=rows, producers=prod) if d % (prod & 255) == 0 else float("nan")
          for d in DEPTHS]
    box = f"2D {{64,{rows}}}" if rows else f"{{64,{bw},{bh},{bn}}}"
    per = (rows if rows else 128) * 128 / rs[-1]
    print(f"({B},{H},{W},{Cc}){'':22s}"[:22] + f" {mb:6.0f} {box:>13s} {ctas:5d} {prod & 255:3d}  " +
          " ".join(f"{r:8.1f}" for r in rs) + f"   ({per:.0f} cycles / load, wait flavour {prod >> 8}; issuing thread per "
          f"load: wait-empty {last_split[0]:.0f}, expect_tx {last_split[1]:.0f}, TMA issue {last_split[2]:.0f})")
print()
print("trace of loads 600..: cycles since kernel start at which the producer got the slot / had issued the load / the "
      "consumer saw it land (1 producer, depth 6, 16 KiB 4-D boxes, L2-resident)")
rate(32, 64, 64, 256, 64, 2, 1, 6, 1, ctas=148, rows=0, producers=1)
tr = out[8:].cpu().view(96, 4)
prev = None
for i in range(0, 40):
    a, b, c = int(tr[i, 0]), int(tr[i, 1]), int(tr[i, 2])
    print(f"  load {600 + i}: slot got {a:8d}  issued {b:8d} (+{b - a:4d})  landed-seen {c:8d} (issue->seen {c - b:5d})"
          + (f"   d(slot got) {a - prev:5d}" if prev is not None else ""))
    prev = a
sys.exit(0)
print()
print("raw issue cost: one thread, n back-to-back 2-D loads {64 ch, rows} on one mbarrier (L2-resident 64 MiB tensor)")
x = torch.empty(32 * 64 * 64, 256, dtype=torch.bfloat16, device="cuda").normal_()
for ctas in (1, 148):
    for rows in (32, 128, 256):
        for n in (1, 2, 4, 6):
            if n * rows * 128 > 200 * 1024:
                continue
            _lib.check(L.hrp_probe_tma_issue(C.c_void_p(x.data_ptr()), C.c_int64(x.shape[0]), 256, rows, n, 50, ctas,
                                             C.c_void_p(out.data_ptr())))
            ti, td = int(out[0]), int(out[1])
            print(f"  CTAs {ctas:3d}  box {rows:3d} rows ({rows * 128 // 1024:2d} KiB)  n={n}:  issued after {ti:5d} cycles "
                  f"({ti / n:5.0f} / load), landed after {td:5d} ({n * rows * 128 / td:5.1f} B/cycle)")
