"""Oracle graph mechanics against the numbers the reference's own tests hold (cpp/tests/test_mobility.cpp,
test_graph_simulation.cpp). CPU only."""
import numpy as np

import oracle_lib
import scenarios
from memilio_b200 import flat
from memilio_b200.flat import ContactLocation


def _edge_apply_model():
    # TestMobility.edgeApplyMobility setup, cpp/tests/test_mobility.cpp:154-176
    G = 1
    pset = flat.osecir_pset(G, TransmissionProbabilityOnContact=1., RiskOfInfectionFromSymptomatic=1.,
                            RelativeTransmissionNoSymptoms=1., SeverePerInfectedSymptoms=0.5, TimeExposed=1.,
                            TimeInfectedNoSymptoms=1.)
    local = np.array([flat.DBL_MAX, flat.DBL_MAX])
    flat.osecir_apply_constraints(pset, G, local)
    ct = flat.contact_table(G, [ContactLocation(np.full((1, 1), 5.0))], 1, 1)
    y0 = np.zeros(10)
    y0[flat.ISY] = 10
    y0[flat.S] = 1000 - 10
    return G, pset, local, ct, y0


def test_edge_apply_mobility_reference_numbers():
    G, pset, local, ct, y0 = _edge_apply_model()
    out = oracle_lib.edge_apply_sequence(G, pset, local, ct, 1, 1, y0, np.full(10, 0.1), 3.125, 0.5)
    # forward mobility: exact values of test_mobility.cpp:181-185
    np.testing.assert_array_equal(out[0, 0], [990 - 99, 0, 0, 0, 10 - 1, 0, 0, 0, 0, 0])
    np.testing.assert_array_equal(out[0, 1], [990 + 99, 0, 0, 0, 10 + 1, 0, 0, 0, 0, 0])
    # returns: test_mobility.cpp:187-204
    v = out[1, 0]
    assert abs(v.sum() - 1000) <= 4 * np.spacing(1000.0)  # EXPECT_DOUBLE_EQ = 4 ulp
    assert v[0] < 990 and abs(v[0] - 990) < 50
    assert 0 < v[1] < 50
    assert 0 < v[2] < 5
    assert abs(v[4] - 10) < 5
    assert 0 < v[6] < 5
    assert abs(out[1, 1].sum() - 1000) <= 4 * np.spacing(1000.0)
    # change node again: test_mobility.cpp:206-214
    assert abs(out[2, 0].sum() - 900) <= 4 * np.spacing(900.0)
    assert abs(out[2, 1].sum() - 1100) <= 4 * np.spacing(1100.0)


def test_graph_no_mobility_equals_single_simulations():
    # TestMobility.compareNoMobilityWithSingleIntegration idea (cpp/tests/test_mobility.cpp:30-127):
    # zero coefficients => every node evolves like a stand-alone simulation advanced in dt_graph chunks
    p = scenarios.example_ode_secir_graph()
    p.edge_coeff[:] = 0.0
    g = oracle_lib.graph_simulate(p, 0, 10.0)
    assert g["steps"] == 20
    assert np.all(g["counters"][:, 3] == 0)
    # node 1 started with S only: nothing may happen
    np.testing.assert_array_equal(g["state"][1], p.y0[0, 1])
    assert abs(g["state"][0].sum() - p.y0[0, 0].sum()) < 1e-9


def test_graph_example_conserves_population_and_moves_infection():
    p = scenarios.example_ode_secir_graph()
    g = oracle_lib.graph_simulate(p, 0, 30.0, halfsteps=60)
    assert g["steps"] == 60
    tot0 = p.y0[0].sum()
    for k in range(60):
        assert abs(g["halfsteps"][k].sum() - tot0) < 1e-8
    assert g["state"][1][flat.R] > 1.0  # infection reached node 1 through commuters
    np.testing.assert_array_equal(g["halfsteps"][-1], g["state"])
