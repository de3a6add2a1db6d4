#!/usr/bin/env python
"""Kernel-variant builds of libkrep_b200.so for A/B timing on the GPU box (selected with KREP_B200_LIB=...).
Only scan_literal.cu is recompiled per variant; the other objects are the in-tree ones.
Usage: python scripts/build_variants.py  ->  build/variants/libkrep_b200_<name>.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from krep_b200 import build as kb  # noqa: E402

VARIANTS = {   # default build: W4_NX=0, W4_WARP_EMIT=0, WARP_EMIT=1 (aligned-word kernel only)
    "nx2": ["-DKREP_B200_W4_NX=2"],
    "nx0_warp_emit": ["-DKREP_B200_W4_WARP_EMIT=1"],
    "nx2_warp_emit": ["-DKREP_B200_W4_NX=2", "-DKREP_B200_W4_WARP_EMIT=1"],
    "aligned_lane_emit": ["-DKREP_B200_WARP_EMIT=0"],
    "count_minb2": ["scan_count.cu", "-DKREP_B200_COUNT_MINB=2"],
}


def main():
    kb.build()
    out_dir = os.path.join(ROOT, "build", "variants")
    os.makedirs(out_dir, exist_ok=True)
    objdir = os.path.join(kb.HERE, "build")
    for name, flags in VARIANTS.items():
        src = "scan_literal.cu"
        if flags and flags[0].endswith(".cu"):
            src, flags = flags[0], flags[1:]
        others = [os.path.join(objdir, s.rsplit(".", 1)[0] + ".o") for s in kb.SOURCES if s != src]
        obj = os.path.join(out_dir, f"{src[:-3]}_{name}.o")
        subprocess.run([kb.NVCC, *[f for f in kb.FLAGS if f not in ("-Xptxas", "-v")], *flags, "-c", os.path.join(kb.CSRC, src), "-o", obj], check=True)
        so = os.path.join(out_dir, f"libkrep_b200_{name}.so")
        subprocess.run([kb.NVCC, "-shared", "-o", so, obj, *others, "-Xcompiler", "-fopenmp", "-lgomp"], check=True)
        print(so)


if __name__ == "__main__":
    main()
