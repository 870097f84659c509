"""The oracle's restatement of cpp/memilio/data/analyze_result.h against the exact small cases of the reference's own
tests (cpp/tests/test_analyze_result.cpp), and the literal per-run study loop (draw_sample + advance + interpolate)
against the pieces it is made of."""
import numpy as np

import oracle_lib
from memilio_b200 import study, synthetic


def test_interpolate_days_add_when_within_tolerance():
    # TestInterpolateTimeSeries.aFewSimplePoints / timePointsCanBeChosen (test_analyze_result.cpp:121-173)
    t = [0.1, 0.9, 1.4, 2.5, 3.0]
    v = [[0.0], [1.0], [2.0], [3.0], [4.0]]
    ot, ov = oracle_lib.interpolate(t, v, tps=[0.2, 0.6, 1.0, 1.3, 2.4])
    np.testing.assert_array_equal(ot, [0.2, 0.6, 1.0, 1.3, 2.4])
    np.testing.assert_allclose(ov[:, 0], [1 / 8, 5 / 8, 1.0 + 1 / 5, 1.0 + 4 / 5, 2.0 + 10 / 11], rtol=1e-14)
    # integer days between ceil(t0 - tol) and floor(tmax + tol)
    ot, ov = oracle_lib.interpolate(t, v)
    np.testing.assert_array_equal(ot, [1.0, 2.0, 3.0])
    np.testing.assert_allclose(ov[:, 0], [1.2, 2.0 + 6 / 11, 4.0], rtol=1e-14)
    # first / last point within tolerance of a whole day are kept (analyze_result.h:194-197, :229-232, :258-261)
    ot, ov = oracle_lib.interpolate([1e-12, 0.5, 2.0 - 1e-12], [[0.0], [1.0], [2.0]])
    np.testing.assert_array_equal(ot, [0.0, 1.0, 2.0])
    np.testing.assert_array_equal(ov[[0, 2], 0], [0.0, 2.0])


def test_interpolate_ensemble_basic():
    # TestInterpolateEnsemble.basic (test_analyze_result.cpp:196-216)
    _, a = oracle_lib.interpolate([0.0, 0.5, 2.0], [[0.0], [1.0], [2.0]])
    _, b = oracle_lib.interpolate([0.0, 1.5, 2.0], [[0.0], [1.0], [2.0]])
    np.testing.assert_allclose(a[:, 0], [0.0, 1.0 + 1 / 3, 2.0], rtol=1e-15)
    np.testing.assert_allclose(b[:, 0], [0.0, 2 / 3, 2.0], rtol=1e-15)


def test_ensemble_sum_mean_percentile_small_cases():
    # TestEnsembleSum.basic (:218-264)
    runs = np.array([[[0.0, 1.0, 2.0], [3.0, 4.0, 5.0], [6.0, 7.0, 8.0]],
                     [[0.5, 2.0, 5.0], [7.5, 9.5, 1.0], [1.5, 2.5, 3.0]]])[..., None]
    r = oracle_lib.ensemble_reduce(runs, want_sum=True)
    np.testing.assert_array_equal(r["sum"][:, :, 0], [[9.0, 12.0, 15.0], [9.5, 14.0, 9.0]])
    # TestEnsembleMean.basic (:266-298)
    runs = np.array([[[0.0, 1.0, 2.0], [0.0, 1.0, 2.0]], [[0.5, 3.0, 0.0], [1.5, 0.5, 1.0]]])[..., None]
    r = oracle_lib.ensemble_reduce(runs, want_mean=True)
    np.testing.assert_array_equal(r["mean"][:, :, 0], [[0.25, 2.0, 1.0], [0.75, 0.75, 1.5]])
    # TestEnsemblePercentile.basic (:300-371): 4 runs, 2 nodes, 3 time points, 2 elements
    runs = np.zeros((4, 2, 3, 2))
    runs[0, 1, 0, 0] = 0.2
    runs[1, 0, 1, 1], runs[1, 1, 0, 0] = 1.0, 0.1
    runs[2, 0, 1, 1], runs[2, 1, 0, 0] = 2.0, 0.3
    runs[3, 0, 1, 1] = 3.0
    q = oracle_lib.ensemble_reduce(runs, p=[0.2, 0.4, 0.7, 0.9])["percentile"]
    assert [q[i, 0, 1, 1] for i in range(4)] == [0.0, 1.0, 2.0, 3.0]
    assert [q[i, 1, 0, 0] for i in range(4)] == [0.0, 0.1, 0.2, 0.3]


def test_study_run_block_is_sample_plus_simulate_plus_interpolate():
    g = synthetic.covid_study_graph(N=8, max_out_degree=3, n_seeded=2, damping_day=2.0)
    ps = study.ParameterStudy(g, 0.0, 3.0, 0.5, 5)
    p = ps.sample(0, 5)  # the product's sampler: bit-identical to the oracle's (tests/test_sampling.py)
    blk = oracle_lib.study_run_block(ps, p, 1, 3, 3.0, 4, want_daily=True)
    for i, run in enumerate((1, 2, 3)):
        ref = oracle_lib.graph_simulate_daily(p, run, 3.0, 4)
        np.testing.assert_array_equal(blk["state"][i], ref["state"])
        np.testing.assert_array_equal(blk["counters"][i], ref["counters"])
        np.testing.assert_array_equal(blk["daily"][i], ref["daily"])
        # a day that coincides with the last time point of a node: y1 + 1.0 * (y2 - y1), equal to y2 up to rounding
        np.testing.assert_allclose(ref["daily"][:, 3], ref["state"], rtol=1e-14, atol=1e-12)
        np.testing.assert_array_equal(ref["daily"][:, 0], p.y0[run])
    assert np.isclose(blk["checksum"], blk["daily"].sum(), rtol=1e-12)
