"""The C-ABI shared library loads without a GPU and exports every symbol include/mio_cuda.h declares; calls that need
a device fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from memilio_b200 import build, capi, cdefs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not fn.endswith(".h"):
            continue
        with open(os.path.join(ROOT, "include", fn)) as f:
            text = f.read()
        names += re.findall(r"MIO_API\s+[\w\s\*]+?\b(mio_\w+)\s*\(", text)
    return names


def test_header_symbols_are_exported():
    lib_path = build.build()
    L = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/ but not exported by {lib_path}"
    # the python mirror binds the same set
    assert sorted(set(names)) == sorted(set(capi.EXPORTS))


def test_size_queries_need_no_device():
    L = capi.lib()
    assert L.mio_cuda_version() >= 100
    assert L.mio_num_compartments(0) == 10 and L.mio_num_compartments(1) == 27
    assert L.mio_pset_size(0, 6) == 8 + 16 * 6
    assert L.mio_ctable_size(6, 1, 2) == 1 + 1 + (2 + 4 * 36)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import scenarios
    p = scenarios.example_ode_secir_graph()
    with pytest.raises(capi.MioCudaError) as e:
        capi.GraphSimulation(p)
    assert e.value.code == 1 and "no CPU fallback" in str(e.value)
