"""The C++ facade (include/memilio_b200/) keeps MEmilio's API surface for the path: the reference's own example
programs compile against it UNCHANGED (checked here when /root/reference is present; it is not on the GPU box), and
the repository's restated examples build and link against the C-ABI library."""
import os
import shutil
import subprocess

import pytest

from memilio_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXAMPLES = "/root/reference/cpp/examples"
CXX = shutil.which("g++")


@pytest.mark.skipif(CXX is None, reason="no host compiler")
@pytest.mark.parametrize("example", ["ode_secir.cpp", "ode_secir_graph.cpp", "ode_secir_parameter_study_graph.cpp",
                                     "ode_secirvvs.cpp"])
def test_reference_example_compiles_unchanged(example):
    src = os.path.join(REF_EXAMPLES, example)
    if not os.path.exists(src):
        pytest.skip("reference tree not present")
    r = subprocess.run([CXX, "-std=c++20", "-fsyntax-only", "-I" + os.path.join(ROOT, "include/memilio_b200/compat"),
                        "-I" + os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.skipif(CXX is None, reason="no host compiler")
def test_repository_examples_build_and_link():
    build.build()
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "examples")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    for prog in ("secir_single", "graph_two_nodes", "parameter_study", "graph_vvs"):
        assert os.path.exists(os.path.join(ROOT, "examples", "_build", prog))
