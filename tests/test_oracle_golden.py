"""Pins the CPU oracle against every golden vector the reference holds for the single-node osecir path
(SURVEY.md section 4 / 8c). CPU only."""
import numpy as np
import pytest

import oracle_lib
import scenarios


@pytest.mark.parametrize("name", list(scenarios.GOLDEN_CASES))
def test_oracle_reproduces_reference_golden_csv(name):
    s = scenarios.golden_case(name)
    t, y, dt, cnt = oracle_lib.simulate(s["G"], s["pset"], s["local"], s["ct"], s["ct_max_loc"], s["ct_max_pts"],
                                        s["y0"], s["t0"], s["tmax"], s["dt0"], s["integrator"])
    ref = s["csv"]
    # same number of time points = same accepted adaptive steps (ASSERT_EQ in the reference tests)
    assert len(t) == ref.shape[0], (len(t), ref.shape[0])
    assert cnt[0] == len(t) - 1
    np.testing.assert_allclose(t, ref[:, 0], rtol=0, atol=1e-10)
    ys = scenarios.age_sum(y, s["G"])
    assert ys.shape[1] == ref.shape[1] - 1
    np.testing.assert_allclose(ys, ref[:, 1:], rtol=0, atol=s["tol"])


def test_oracle_adaptive_controller_is_exercised():
    # the strongest controller pin: steps grow 0.1 -> 0.45 (cap 0.9*5), shrink around the damping kink t in [29,31]
    s = scenarios.golden_case("py_default_g3")
    t, y, dt, cnt = oracle_lib.simulate(s["G"], s["pset"], s["local"], s["ct"], 1, 2, s["y0"], 0.0, 50.0, 0.1,
                                        s["integrator"])
    steps = np.diff(t)
    assert abs(steps[0] - 0.1) < 1e-12 and abs(steps[1] - 0.45) < 1e-12
    assert cnt[1] > 0  # some attempts were rejected
    assert steps.max() > 2.0 and steps[(t[1:] > 29) & (t[1:] < 31.5)].min() < 0.6


def test_vvs_oracle_reproduces_reference_golden_csv():
    # osecirvvs (27 compartments, day-chunk advance of osecirvvs::Simulation) against
    # pycode/memilio-simulation/tests/data/ode-secirvvs-compare.csv, tolerance of the reference test (1e-10)
    s = scenarios.vvs_golden_case()
    t, y, dt, cnt = oracle_lib.simulate_vvs(s["G"], s["pset"], s["local"], s["ct"], 1, 2, s["y0"], 0.0, 50.0, 0.1,
                                            vacc_partial=s["vacc_partial"], vacc_full=s["vacc_full"])
    ref = s["csv"]
    assert len(t) == ref.shape[0], (len(t), ref.shape[0])
    np.testing.assert_allclose(t, ref[:, 0], rtol=0, atol=1e-10)
    assert y.shape[1] == ref.shape[1] - 1 == 27
    np.testing.assert_allclose(y, ref[:, 1:], rtol=0, atol=1e-10)
    # the day chunks force a time point on every integer day (model.h:752)
    assert set(np.arange(1, 51)).issubset(set(np.round(t[np.abs(t - np.round(t)) < 1e-12]).astype(int)))
