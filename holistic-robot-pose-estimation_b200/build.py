"""In-tree build of libhrp_b200.so (hand-written sm_100a CUDA behind the C ABI of include/hrp.h).

`python -m horopose_b200.build` or `__graft_entry__.build()` compiles every .cu under csrc/ (plus the hardware probes
of tools/probe/ when HRP_BUILD_PROBES=1) with
nvcc -gencode arch=compute_100a,code=sm_100a and links one shared library next to this file.  The build is
incremental (object files under build/, rebuilt when a source or header is newer).  nvcc cross-compiles
without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
REPO = PKG_DIR.parent
LIB_PATH = PKG_DIR / "libhrp_b200.so"
BUILD_DIR = PKG_DIR / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-I", str(REPO / "include"),
    "-I", str(CSRC),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newest_header_mtime() -> float:
    hdrs = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((REPO / "include").glob("*.h"))
    return max(h.stat().st_mtime for h in hdrs)


def build(verbose: bool = False, force: bool = False) -> Path:
    BUILD_DIR.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    probes = os.environ.get("HRP_BUILD_PROBES") == "1"
    if probes:  # hardware probes (tools/probe/): profiling aids, not part of the product library by default
        srcs.append(REPO / "tools" / "probe" / "probe.cu")
    stale_probe = BUILD_DIR / "probe.o"
    if not probes and stale_probe.exists():
        stale_probe.unlink()
        force_link = True
    else:
        force_link = False
    hdr_m = _newest_header_mtime()
    jobs = []
    objs = []
    for src in srcs:
        obj = BUILD_DIR / (src.stem + ".o")
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_m):
            cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get("HRP_NVCC_EXTRA", "").split(), "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stdout + r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(run, jobs):
                if verbose and out.strip():
                    print(out)
    if jobs or not LIB_PATH.exists() or force or force_link:
        cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs), "-gencode",
               "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        out = run(cmd)
        if verbose and out.strip():
            print(out)
    return LIB_PATH


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
