"""Build the CUDA library (sm_100a only) in-tree: memilio_b200/csrc/libmio_cuda.so."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libmio_cuda.so")
CU_SOURCES = ["integrate.cu", "edges.cu", "reduce.cu", "capi.cu", "peak.cu", "comm.cu"]
CXX_SOURCES = ["study.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler",
              "-fPIC,-fvisibility=hidden,-O2", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "mio_cuda.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB, objdir: str = CSRC) -> str:
    """Compile every CUDA / C++ source for sm_100a and link the C-ABI library. `defines` (e.g. launch-shape macros of
    kernels.h) and `out` let tools/build_variants.py produce tuning variants next to the default library."""
    if out == LIB and not force and not needs_build():
        return LIB
    os.makedirs(objdir, exist_ok=True)
    from concurrent.futures import ThreadPoolExecutor

    incremental = not force and not defines and not os.environ.get("MIO_NVCC_EXTRA")

    def deps_time(src):
        """newest mtime of a source and the local headers it includes (transitively)"""
        import re
        seen, todo = set(), [os.path.join(CSRC, src)]
        while todo:
            f = os.path.normpath(todo.pop())
            if f in seen or not os.path.exists(f):
                continue
            seen.add(f)
            with open(f) as fh:
                for inc in re.findall(r'^\s*#include\s+"([^"]+)"', fh.read(), flags=re.M):
                    todo.append(os.path.join(os.path.dirname(f), inc))
        return max(os.path.getmtime(f) for f in seen)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        if incremental and os.path.exists(obj) and os.path.getmtime(obj) > deps_time(src):
            return obj, None  # up to date (object newer than its source and the headers it includes)
        extra = os.environ.get("MIO_NVCC_EXTRA", "").split()  # tuning experiments only
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed for " + src)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, CU_SOURCES + CXX_SOURCES))
    objs = [o for o, _ in results]
    logs = [l for _, l in results if l is not None]
    if incremental and len(logs) < len(results):
        logs.append("(objects not listed above were up to date; `--force` rebuilds everything)")
    cmd = [_nvcc(), "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    os.makedirs(os.path.join(objdir, "build"), exist_ok=True)
    with open(os.path.join(objdir, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
