"""the HDF5 result files of mio::save_result (cpp/memilio/io/result_io.cpp:33-100) written without an HDF5
library (include/memilio_b200/h5lite.hpp). No HDF5 library or tool exists in this image, so the check is a reader of
our own (tests/h5lite.py) that (a) reads a result file the REFERENCE wrote with libhdf5 (tests/golden/
results_osecirvvs.h5 = cpp/tests/data/results_osecirvvs.h5) and (b) reads back what the facade writes; the structures
both files use (superblock version 0, symbol table groups, version 1 object headers, the datatype message bytes) are
compared field by field."""
import os
import struct
import subprocess

import numpy as np

import h5lite

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "results_osecirvvs.h5")

PROGRAM = r'''
#include "memilio_b200/mio.hpp"
int main(int argc, char** argv)
{
    // three regions, 2 age groups x 3 compartments, 5 time points; ids out of order and more regions than the default
    // symbol table node holds in a second file
    const int G = 2, C = 3;
    auto series = [&](int id) {
        mio::TimeSeries<double> ts(G * C);
        for (int t = 0; t < 5; ++t) {
            Eigen::VectorX<double> v(G * C);
            for (int e = 0; e < G * C; ++e) {
                v[e] = id + 0.25 * t + 10.0 * e;
            }
            ts.add_time_point(0.5 * t, v);
        }
        return ts;
    };
    std::vector<mio::TimeSeries<double>> res{series(1001), series(5), series(1002)};
    if (!mio::save_result(res, {1001, 5, 1002}, G, std::string(argv[1]) + "/three.h5")) {
        return 1;
    }
    std::vector<mio::TimeSeries<double>> many;
    std::vector<int> ids;
    for (int k = 0; k < 400; ++k) {
        many.push_back(series(1000 + k));
        ids.push_back(1000 + k);
    }
    if (!mio::save_result(many, ids, G, std::string(argv[1]) + "/many.h5")) {
        return 2;
    }
    // mio::read_result (io/result_io.cpp:110-230): our own file and the one the reference wrote with libhdf5
    auto back = mio::read_result(std::string(argv[1]) + "/three.h5");
    if (!back || back.value().size() != 3) {
        return 3;
    }
    const auto& r5 = back.value()[0]; // regions come back in ascending id order: 5, 1001, 1002
    if (r5.get_groups().get_num_time_points() != 5 || r5.get_groups().get_num_elements() != G * C ||
        r5.get_groups()[2][4] != 5 + 0.25 * 2 + 10.0 * 4 || r5.get_totals()[4][1] != 2 * (5 + 0.25 * 4) + 10.0 * (1 + 4) ||
        r5.get_totals().get_time(3) != 1.5) {
        return 4;
    }
    auto ref = mio::read_result(argv[2]);
    if (!ref || ref.value().size() != 1) {
        return 5;
    }
    const auto& rr = ref.value()[0];
    std::printf("ref %td %td %td %.17g %.17g\n", rr.get_groups().get_num_time_points(), rr.get_groups().get_num_elements(),
                rr.get_totals().get_num_elements(), rr.get_totals()[0][0], rr.get_groups()[30][80]);
    // save_edges (io/mobility_io.cpp:206-306): two start nodes, three edges
    std::vector<mio::TimeSeries<double>> edges{series(1), series(2), series(3)};
    if (!mio::save_edges(edges, {{1001, 1002}, {1001, 1003}, {1002, 1001}}, std::string(argv[1]) + "/edges.h5")) {
        return 6;
    }
    return mio::read_result(std::string(argv[1]) + "/missing.h5") ? 7 : 0; // a missing file is an error, not a crash
}
'''


def test_reader_reads_the_reference_file():
    r = h5lite.read_result(GOLDEN)
    assert list(r) == [0]
    g = r[0]
    assert sorted(g) == ["Group1", "Group2", "Group3", "Time", "Total"]
    assert g["Time"].shape == (31,) and g["Total"].shape == (31, 27)
    np.testing.assert_array_equal(g["Time"], np.arange(31.0))
    np.testing.assert_allclose(g["Group1"] + g["Group2"] + g["Group3"], g["Total"], rtol=1e-14)


def _expected(id_, G=2, Cn=3):
    t = np.arange(5)
    e = np.arange(G * Cn)
    return id_ + 0.25 * t[:, None] + 10.0 * e[None, :]


def test_facade_save_result_round_trip_and_structure(tmp_path):
    from memilio_b200 import build
    lib = build.build()
    src = tmp_path / "save.cpp"
    src.write_text(PROGRAM)
    exe = tmp_path / "save"
    csrc = os.path.dirname(lib)
    subprocess.run(["g++", "-O1", "-std=c++20", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L" + csrc, "-lmio_cuda", "-Wl,-rpath," + csrc], check=True)
    r = subprocess.run([str(exe), str(tmp_path), GOLDEN], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    # the facade's reader on the file libhdf5 wrote agrees with tests/h5lite.py
    gold = h5lite.read_result(GOLDEN)[0]
    w = r.stdout.split()
    assert w[0] == "ref" and [int(x) for x in w[1:4]] == [31, 81, 27]
    assert float(w[4]) == gold["Total"][0][0] and float(w[5]) == gold["Group3"][30][26]
    edges = h5lite.read_result(str(tmp_path / "edges.h5"))
    assert sorted(edges) == [1001, 1002]
    assert sorted(edges[1001]) == ["End1002", "End1003", "Time", "Total"] and sorted(edges[1002]) == ["End1001", "Time", "Total"]
    np.testing.assert_array_equal(edges[1001]["End1003"], _expected(2))
    np.testing.assert_array_equal(edges[1001]["Total"], _expected(1) + _expected(2))
    r = h5lite.read_result(str(tmp_path / "three.h5"))
    assert sorted(r) == [5, 1001, 1002]
    for id_, g in r.items():
        v = _expected(id_)
        np.testing.assert_array_equal(g["Time"], 0.5 * np.arange(5))
        np.testing.assert_array_equal(g["Group1"], v[:, :3])
        np.testing.assert_array_equal(g["Group2"], v[:, 3:])
        np.testing.assert_array_equal(g["Total"], v[:, :3] + v[:, 3:])
    many = h5lite.read_result(str(tmp_path / "many.h5"))
    assert sorted(many) == list(range(1000, 1400))
    np.testing.assert_array_equal(many[1399]["Group2"], _expected(1399)[:, 3:])

    # structure, field by field against the file libhdf5 wrote
    ours, ref = h5lite.H5File(str(tmp_path / "three.h5")), h5lite.H5File(GOLDEN)
    assert ours.b[:16] == ref.b[:16]  # signature, versions, sizes of offsets and lengths
    assert ours.internal_k == ref.internal_k == 16
    for f in (ours, ref):
        assert f.root["cache"] == 1  # the root entry caches its B-tree and heap
        assert [(t, len(d)) for t, _, d in f.messages(f.root["ohdr"])] == [(0x11, 16)]
        assert f.b[f.root["btree"]:f.root["btree"] + 6] == b"TREE\x00\x00"
        assert f.b[f.root["heap"]:f.root["heap"] + 8] == b"HEAP\x00\x00\x00\x00"
    def dataset_messages(f):
        region = list(f.listdir().values())[0]
        ds = f.listdir(region)["Group1"]
        return {t: (fl, d) for t, fl, d in f.messages(ds["ohdr"])}
    mo, mr = dataset_messages(ours), dataset_messages(ref)
    assert mo[0x03] == mr[0x03]  # datatype: IEEE little-endian double, byte for byte, same flags
    assert mo[0x05] == mr[0x05]  # fill value message
    assert mo[0x08][1][:2] == mr[0x08][1][:2] == b"\x03\x01"  # layout version 3, contiguous
    assert mo[0x01][1][0] == mr[0x01][1][0] == 1 and mo[0x01][1][1] == mr[0x01][1][1] == 2  # dataspace v1, rank 2
    assert struct.unpack_from("<QQ", mo[0x01][1], 8) == (5, 3)
