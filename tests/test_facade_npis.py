"""The facade's dynamic NPI host logic (include/memilio_b200/npis.hpp) needs no GPU: the exact cases of the reference's
cpp/tests/test_dynamic_npis.cpp (implement, get_active_damping, get_damping_indices) are compiled against the facade
and run on the CPU. (Simulation::advance with NPIs integrates on the GPU: tests/test_gpu_facade.py.)"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r'''
#include "memilio_b200/mio.hpp"
#include <cstdio>
#include <cstdlib>
#define CHECK(c) do { if (!(c)) { std::printf("FAILED line %d: %s\n", __LINE__, #c); std::exit(1); } } while (0)
using T = mio::SimulationTime<double>;
using L = mio::DampingLevel;
using Y = mio::DampingType;
static bool is(const mio::ContactMatrix<double>::RawDamping& d, double v, int l, int y, double t)
{
    for (Eigen::Index k = 0; k < d.m.size(); ++k) {
        if (std::abs(d.m.data()[k] - v) > 1e-14) return false;
    }
    return d.level == l && d.type == y && std::abs(d.time - t) < 1e-14;
}
int main()
{
    {   // TEST(DynamicNPIs, get_damping_indices), get_active_damping
        mio::ContactMatrix<double> e(Eigen::MatrixX<double>::Constant(1, 1, 1.0));
        e.add_damping(0.3, L(0), Y(0), T(0.3)); e.add_damping(0.4, L(0), Y(0), T(0.4));
        e.add_damping(0.6, L(1), Y(1), T(0.6)); e.add_damping(0.5, L(0), Y(1), T(0.5));
        e.add_damping(0.4, L(0), Y(0), T(0.7)); e.add_damping(0.5, L(0), Y(0), T(0.9));
        auto idx = mio::get_damping_indices(e, L(0), Y(0), T(0.3), T(0.9));
        CHECK(idx.size() == 2 && idx[0] == 1 && idx[1] == 4);
        CHECK(mio::get_active_damping(e, L(0), Y(0), T(0.6))(0, 0) == 0.4);
        CHECK(mio::get_active_damping(e, L(1), Y(1), T(0.6))(0, 0) == 0.6);
        CHECK(mio::get_active_damping(e, L(1), Y(1), T(0.4))(0, 0) == 0.0);
        CHECK(mio::get_active_damping(e, L(0), Y(0), T(0.8))(0, 0) == 0.4);
    }
    {   // TEST(DynamicNPIs, implement): 3 x 1 expressions, the mask is the identity
        mio::ContactMatrixGroup<double> g(2, 1);
        g[0] = mio::ContactMatrix<double>(Eigen::MatrixX<double>::Constant(3, 1, 1.0));
        g[1] = g[0];
        auto mask = [](const Eigen::VectorX<double>& w) {
            Eigen::MatrixX<double> m(3, 1);
            for (int i = 0; i < 3; ++i) m(i, 0) = w[i];
            return m;
        };
        g[0].add_damping(0.15, L(0), Y(0), T(0.3)); g[0].add_damping(0.2, L(0), Y(0), T(0.4));
        g[0].add_damping(0.25, L(0), Y(1), T(0.5)); g[0].add_damping(0.35, L(0), Y(0), T(0.7));
        g[1].add_damping(0.25, L(1), Y(1), T(0.5)); g[1].add_damping(0.45, L(0), Y(0), T(0.9));
        std::vector<mio::DampingSampling<double>> a{mio::DampingSampling<double>(
            0.4, L(0), Y(0), T(0), {0, 1}, Eigen::VectorX<double>::Constant(3, 1.0))};
        mio::implement_dynamic_npis<double>(g, a, T(0.45), T(0.6), mask);
        const auto& d0 = g[0].get_dampings();
        CHECK(d0.size() == 6 && is(d0[0], 0.15, 0, 0, 0.3) && is(d0[1], 0.2, 0, 0, 0.4) && is(d0[2], 0.4, 0, 0, 0.45) &&
              is(d0[3], 0.25, 0, 1, 0.5) && is(d0[4], 0.2, 0, 0, 0.6) && is(d0[5], 0.35, 0, 0, 0.7));
        const auto& d1 = g[1].get_dampings();
        CHECK(d1.size() == 4 && is(d1[0], 0.4, 0, 0, 0.45) && is(d1[1], 0.25, 1, 1, 0.5) && is(d1[2], 0.0, 0, 0, 0.6) &&
              is(d1[3], 0.45, 0, 0, 0.9));
        std::vector<mio::DampingSampling<double>> b{mio::DampingSampling<double>(
            0.3, L(0), Y(0), T(0), {0}, Eigen::VectorX<double>::Constant(3, 1.0))};
        mio::implement_dynamic_npis<double>(g, b, T(0.3), T(0.9), mask);
        const auto& e0 = g[0].get_dampings();
        CHECK(e0.size() == 5 && is(e0[0], 0.3, 0, 0, 0.3) && is(e0[1], 0.4, 0, 0, 0.45) && is(e0[2], 0.25, 0, 1, 0.5) &&
              is(e0[3], 0.3, 0, 0, 0.6) && is(e0[4], 0.35, 0, 0, 0.7));
        CHECK(g[1].get_dampings().size() == 4);
    }
    {   // TEST(DynamicNPIs, set_threshold / get_threshold): sorted descending, biggest exceeded threshold
        mio::DynamicNPIs<double> n;
        n.set_threshold(0.5, {}); n.set_threshold(1.0, {}); n.set_threshold(0.75, {});
        CHECK(n.get_thresholds()[0].first == 1.0 && n.get_thresholds()[1].first == 0.75 && n.get_thresholds()[2].first == 0.5);
        CHECK(n.get_max_exceeded_threshold(2.0)->first == 1.0);
        CHECK(n.get_max_exceeded_threshold(0.8)->first == 0.75);
        CHECK(n.get_max_exceeded_threshold(0.4) == n.get_thresholds().end());
    }
    std::printf("ok\n");
    return 0;
}
'''


def test_facade_dynamic_npi_host_logic(tmp_path):
    from memilio_b200 import build
    lib = build.build()
    src = tmp_path / "npis_host.cpp"
    src.write_text(PROGRAM)
    exe = tmp_path / "npis_host"
    csrc = os.path.dirname(lib)
    subprocess.run(["g++", "-O1", "-std=c++20", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L" + csrc, "-lmio_cuda", "-Wl,-rpath," + csrc], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr
