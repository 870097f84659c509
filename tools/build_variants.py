#!/usr/bin/env python
"""Build launch-shape variants of the library under memilio_b200/csrc/variants/ for A/B timing on the GPU box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from memilio_b200 import build as b  # noqa: E402

VARIANTS = {
    "i128x2": ["MIO_TUNE_ONLY_G6", "MIO_INTEGRATE_THREADS=128", "MIO_INTEGRATE_MINBLOCKS=2"],
    "i128x3": ["MIO_TUNE_ONLY_G6", "MIO_INTEGRATE_THREADS=128", "MIO_INTEGRATE_MINBLOCKS=3"],
    "i128x4": ["MIO_TUNE_ONLY_G6", "MIO_INTEGRATE_THREADS=128", "MIO_INTEGRATE_MINBLOCKS=4"],
    "i256x1": ["MIO_TUNE_ONLY_G6", "MIO_INTEGRATE_THREADS=256", "MIO_INTEGRATE_MINBLOCKS=1"],
    "i160x3": ["MIO_TUNE_ONLY_G6", "MIO_INTEGRATE_THREADS=160", "MIO_INTEGRATE_MINBLOCKS=3"],
    "i192x2": ["MIO_TUNE_ONLY_G6", "MIO_INTEGRATE_THREADS=192", "MIO_INTEGRATE_MINBLOCKS=2"],
    "p10": ["MIO_TUNE_ONLY_G6", "MIO_INTEGRATE_ITEM_PASSES=10"],
    "p20": ["MIO_TUNE_ONLY_G6", "MIO_INTEGRATE_ITEM_PASSES=20"],
    "p2": ["MIO_TUNE_ONLY_G6", "MIO_INTEGRATE_ITEM_PASSES=2"],
    "xf8": ["MIO_TUNE_ONLY_G6", "MIO_EXACT_MINBLOCKS_FACT=8"],
    "xf12": ["MIO_TUNE_ONLY_G6", "MIO_EXACT_MINBLOCKS_FACT=12"],
    "s384": ["MIO_TUNE_ONLY_G6", "MIO_SELECT_THREADS=384"],
    "s640": ["MIO_TUNE_ONLY_G6", "MIO_SELECT_THREADS=640"],
    "s768": ["MIO_TUNE_ONLY_G6", "MIO_SELECT_THREADS=768"],
    "s512": ["MIO_TUNE_ONLY_G6", "MIO_SELECT_THREADS=512"],
    "s1024": ["MIO_TUNE_ONLY_G6", "MIO_SELECT_THREADS=1024"],
    "r128x4": ["MIO_TUNE_ONLY_G6", "MIO_RETURN_TMA_THREADS=128", "MIO_RETURN_TMA_MINBLOCKS=4"],
    "r512x1": ["MIO_TUNE_ONLY_G6", "MIO_RETURN_TMA_THREADS=512", "MIO_RETURN_TMA_MINBLOCKS=1"],
    "x8": ["MIO_TUNE_ONLY_G6", "MIO_EXACT_MINBLOCKS=8"],
    "x10": ["MIO_TUNE_ONLY_G6", "MIO_EXACT_MINBLOCKS=10"],
    "x12": ["MIO_TUNE_ONLY_G6", "MIO_EXACT_MINBLOCKS=12"],
    "i64x4": ["MIO_TUNE_ONLY_G6", "MIO_INTEGRATE_THREADS=64", "MIO_INTEGRATE_MINBLOCKS=4"],
    "i64x6": ["MIO_TUNE_ONLY_G6", "MIO_INTEGRATE_THREADS=64", "MIO_INTEGRATE_MINBLOCKS=6"],
    "e256x2": ["MIO_DEPART_MINBLOCKS=2", "MIO_RETURN_MINBLOCKS=3"],
    "e256x4": ["MIO_DEPART_MINBLOCKS=4", "MIO_RETURN_MINBLOCKS=4"],
    "e128x6": ["MIO_EDGE_THREADS=128", "MIO_DEPART_MINBLOCKS=6", "MIO_RETURN_MINBLOCKS=8"],
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for name in names:
        d = os.path.join(b.CSRC, "variants", name)
        out = os.path.join(b.CSRC, "variants", f"libmio_cuda_{name}.so")
        b.build(force=True, defines=VARIANTS[name], out=out, objdir=d)
        print(out)
