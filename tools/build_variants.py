#!/usr/bin/env python
"""Build tuning variants of the CUDA library: tools/build_variants.py name=DEF1,DEF2=VAL ... -> memilio_b200/csrc/variants/<name>/libmio_cuda.so
(select one at run time with MIO_CUDA_LIB=<path>). Variant builds compile the G = 6 kernels only."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from memilio_b200 import build as b  # noqa: E402


def one(spec):
    name, _, defs = spec.partition("=")
    defines = [d for d in defs.split(",") if d]
    d = os.path.join(b.CSRC, "variants", name)
    os.makedirs(d, exist_ok=True)
    out = os.path.join(d, "libmio_cuda.so")
    b.build(force=True, defines=defines, out=out, objdir=d)
    log = open(os.path.join(d, "build", "ptxas.log")).read().splitlines()
    for i, ln in enumerate(log):
        if "integrate_nodes_kernelILi6ELi1ELi0" in ln and "Compiling" in ln:
            print(name, "|", log[i + 2].strip(), "|", log[i + 3].strip()[:60])
    return out


if __name__ == "__main__":
    with ThreadPoolExecutor(max_workers=2) as ex:
        for o in ex.map(one, sys.argv[1:]):
            print(os.path.relpath(o, ROOT))
