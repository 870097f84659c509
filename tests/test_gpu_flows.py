"""mio::FlowSimulation nodes (cpp/memilio/compartments/flow_simulation.h:37-97) on the GPU, through the
C-ABI, against the oracle's literal restatement and the exact relations of the reference's own test
(TestOdeSecir.testDamping, cpp/tests/test_odesecir.cpp:604-676). Tolerance as for the populations: 1e-8 relative,
step counts equal."""
import numpy as np
import pytest

import oracle_lib
import scenarios
from memilio_b200 import capi, cdefs, flat, synthetic
from memilio_b200.problem import single_node_problem

pytestmark = pytest.mark.gpu


def assert_close(a, b, rtol=1e-8, atol=1e-9, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert np.all(err <= 0), f"{what}: max violation {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"


def _flow_problem(s, integrator, dt0):
    p = single_node_problem(s["G"], s["pset"], s["local"], s["ct"], s["y0"], ct_max_loc=s["ct_max_loc"],
                            ct_max_pts=s["ct_max_pts"], integrator=integrator, t0=0.0, dt0=dt0)
    p.flow_simulation = 1
    return p


@pytest.mark.parametrize("name", ["py_default_g1", "py_default_g3", "cpp_ck54_g3"])
def test_simulate_flows_matches_oracle(name):
    s = scenarios.golden_case(name)
    t_o, y_o, f_o, dt_o, cnt_o = oracle_lib.simulate_flows(s["G"], s["pset"], s["local"], s["ct"], s["ct_max_loc"],
                                                           s["ct_max_pts"], s["y0"], 0.0, s["tmax"], s["dt0"],
                                                           s["integrator"])
    with capi.GraphSimulation(_flow_problem(s, s["integrator"], s["dt0"]), traj_capacity=256) as sim:
        sim.advance(s["tmax"])
        npts, times, values = sim.trajectory()
        ftraj = sim.flow_trajectory()
        flows = sim.flows()
        cnt = sim.counters()[0, 0]
        dt = sim.node_dt()[0, 0]
        state = sim.state()
    n = int(npts[0, 0])
    assert n == len(t_o)
    np.testing.assert_array_equal(cnt[:3], cnt_o[:3])
    assert_close(times[0, 0, :n], t_o, rtol=1e-8, atol=1e-10, what="time points")
    assert_close(ftraj[0, 0, :n], f_o, what="flow series")
    assert_close(values[0, 0, :n], y_o, what="populations derived from the flows")
    assert_close(flows[0, 0], f_o[-1], what="get_flows().get_last_value()")
    assert_close(state[0, 0], y_o[-1], what="get_result().get_last_value()")
    assert abs(dt - dt_o) <= 1e-7 * abs(dt_o)


def test_simulate_flows_in_several_advance_calls():
    # every advance call restarts from the last population / flow time point (flow_simulation.h:63-92)
    s = scenarios.golden_case("py_default_g3")
    with capi.GraphSimulation(_flow_problem(s, s["integrator"], s["dt0"]), traj_capacity=256) as sim:
        for t in (7.0, 7.25, 30.0):
            sim.advance(t)
        flows = sim.flows()[0, 0]
        state = sim.state()[0, 0]
        cnt = sim.counters()[0, 0]
    t_o, y_o, f_o, _, cnt_o = oracle_lib.simulate_flows_stepwise(s["G"], s["pset"], s["local"], s["ct"],
                                                                  s["ct_max_loc"], s["ct_max_pts"], s["y0"], 0.0,
                                                                  [7.0, 7.25, 30.0], s["dt0"], s["integrator"])
    assert_close(flows, f_o[-1], what="flows")
    assert_close(state, y_o[-1], what="population")
    np.testing.assert_array_equal(cnt[:3], cnt_o[:3])


def test_damping_relations_with_euler():
    # TestOdeSecir.testDamping: one Euler step of dt = 1; the S->E flows are exactly halved / zero / doubled
    res = {}
    integ = cdefs.integrator_desc(kind=flat.INTEGRATOR_EULER)
    for key, damp in (("a", None), ("b", 0.5), ("c", 1.0), ("d", -1.0)):
        s = scenarios.damping_flow_case(damp)
        with capi.GraphSimulation(_flow_problem(s, integ, 1.0), traj_capacity=8) as sim:
            sim.advance(1.0)
            res[key] = sim.flows()[0, 0]
        t_o, y_o, f_o, _, _ = oracle_lib.simulate_flows(1, s["pset"], s["local"], s["ct"], 1, 2, s["y0"], 0.0, 1.0,
                                                        1.0, integ)
        assert_close(res[key], f_o[-1], rtol=1e-12, atol=0, what=f"flows of model_{key}")
    assert res["a"][0] > 0
    assert 2 * res["b"][0] == res["a"][0]
    assert res["c"][0] == 0.0
    assert 2 * res["a"][0] == res["d"][0]


def test_flow_simulation_nodes_in_a_graph():
    # Graph<SimulationNode<FlowSimulation<Model>>, MobilityEdge>: the edges change the populations, the flows run on
    runs, N, G = 3, 30, 6
    p = synthetic.county_problem(N=N, runs=runs, G=G, max_out_degree=6, n_seeded=3, damping_day=4.0, perturb=0.1)
    p.flow_simulation = 1
    tmax = 9.0
    with capi.GraphSimulation(p) as sim:
        sim.advance(tmax)
        state = sim.state()
        flows = sim.flows()
        cnt = sim.counters()
    for r in range(runs):
        ref = oracle_lib.graph_simulate_flows(p, r, tmax)
        assert_close(state[r], ref["state"], what=f"run {r} populations")
        assert_close(flows[r], ref["flows"], what=f"run {r} flows")
        np.testing.assert_array_equal(cnt[r][:, :3], ref["counters"][:, :3])
        np.testing.assert_array_equal(cnt[r][:, 3], ref["counters"][:, 3])


def test_flow_simulation_rejects_unsupported_combinations():
    s = scenarios.golden_case("cpp_rkf45_g1")
    p = _flow_problem(s, s["integrator"], 0.3)
    with pytest.raises(capi.MioCudaError):
        capi.GraphSimulation(p)
