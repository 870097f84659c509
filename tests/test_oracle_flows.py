"""The oracle's mio::FlowSimulation restatement (compartments/flow_simulation.h:37-97, flow_simulation_base.h:40-120),
pinned by the reference's own tests: TestOdeSecir.testDamping (cpp/tests/test_odesecir.cpp:604-676, exact relations
between the S->E flows of an Euler step) and TestGraphSimulation.consistencyFlowMobility (see test_oracle_golden.py)."""
import numpy as np

import oracle_lib
import scenarios
from memilio_b200 import cdefs, flat


def _euler_flows(damping, dt=1.0, tmax=1.0):
    s = scenarios.damping_flow_case(damping)
    return oracle_lib.simulate_flows(s["G"], s["pset"], s["local"], s["ct"], s["ct_max_loc"], s["ct_max_pts"], s["y0"],
                                     0.0, tmax, dt, cdefs.integrator_desc(kind=flat.INTEGRATOR_EULER))


def test_damping_relations_of_the_reference_test():
    # EXPECT_EQ(2 * result_b[1].get_last_value()[0], result_a[1].get_last_value()[0]) etc.
    a = _euler_flows(None)[2][-1][0]
    b = _euler_flows(0.5)[2][-1][0]
    c = _euler_flows(1.0)[2][-1][0]
    d = _euler_flows(-1.0)[2][-1][0]
    assert a > 0
    assert 2 * b == a
    assert c == 0.0
    assert 2 * a == d


def test_flow_simulation_agrees_with_simulation():
    # TestFlows / TestCompartmentSimulation: the populations derived from the integrated flows describe the same
    # solution as the integrated populations (different integrands, so only up to the integrator's tolerance), the
    # flows are non-decreasing, and population(t) = y0 + get_derivatives(flows(t)) holds exactly
    s = scenarios.golden_case("py_default_g3")
    t_f, y_f, f_f, _, cnt = oracle_lib.simulate_flows(s["G"], s["pset"], s["local"], s["ct"], s["ct_max_loc"],
                                                      s["ct_max_pts"], s["y0"], 0.0, 50.0, 0.1, s["integrator"])
    t_s, y_s, _, _ = oracle_lib.simulate(s["G"], s["pset"], s["local"], s["ct"], s["ct_max_loc"], s["ct_max_pts"],
                                         s["y0"], 0.0, 50.0, 0.1, s["integrator"])
    assert t_f[-1] == t_s[-1] == 50.0
    np.testing.assert_allclose(y_f[-1], y_s[-1], rtol=2e-3, atol=1e-3)
    assert np.all(np.diff(f_f, axis=0) >= 0)
    assert np.all(f_f[0] == 0)
    np.testing.assert_allclose(y_f.sum(axis=1), y_f[0].sum(), rtol=1e-12)
    assert cnt[0] == len(t_f) - 1
