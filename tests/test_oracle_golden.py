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


@pytest.mark.parametrize("flow_sim", [0, 1])
def test_graph_mechanics_seir_golden(flow_sim):
    # TestGraphSimulation.consistencyFlowMobility (cpp/tests/test_graph_simulation.cpp:276-339): node 0 of a fully
    # connected 3-node ODE-SEIR graph, 1000 graph steps of dt 0.001 with commuter departure / return every other step,
    # against cpp/tests/data/graphsimulation-compare.csv at the reference's 1e-10. Pins the oracle's GraphSim,
    # SimulationNode, MobilityEdge::apply_mobility (incl. the Euler return step) and the default integrator on a graph.
    import ctypes as C
    import os
    import numpy as np
    import oracle_lib
    from memilio_b200 import cdefs
    ref = np.loadtxt(os.path.join(os.path.dirname(__file__), "golden", "graphsimulation-compare.csv"), skiprows=1)
    cap = 2048
    t = np.zeros(cap)
    v = np.zeros((cap, 4))
    L = oracle_lib.lib()
    # flow_sim = 1: the nodes are mio::FlowSimulation (the variant the golden file was written from, :307-338); the
    # reference test requires both variants to agree with it: pins the oracle's FlowSimulation restatement as well
    L.orc_seir_graph_golden_ex.restype = C.c_int
    n = L.orc_seir_graph_golden_ex(flow_sim, cap, cdefs.dptr(t), cdefs.dptr(v))
    assert n == ref.shape[0] == 1001
    np.testing.assert_allclose(t[:n], ref[:, 0], rtol=0, atol=1e-10)
    np.testing.assert_allclose(v[:n], ref[:, 1:5], rtol=0, atol=1e-10)
