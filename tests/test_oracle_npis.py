"""The oracle's dynamic NPI restatement (oracle/mio_oracle_npis.hpp) pinned by the exact cases of the
reference's own tests, cpp/tests/test_dynamic_npis.cpp (restated here, the expected values are the reference's)."""
import numpy as np

import oracle_lib
from memilio_b200 import flat
from memilio_b200.flat import ContactLocation


def test_get_damping_indices():
    # TEST(DynamicNPIs, get_damping_indices), :86-103 (the time period is an open interval)
    d = [(0.0, 0, 0, 0.3), (0.0, 0, 0, 0.4), (0.0, 1, 1, 0.6), (0.0, 0, 1, 0.5), (0.0, 0, 0, 0.7), (0.0, 0, 0, 0.9)]
    assert oracle_lib.npi_damping_indices(d, 0, 0, 0.3, 0.9) == [1, 4]


def test_get_active_damping():
    # TEST(DynamicNPIs, get_active_damping / get_active_damping_empty), :105-143
    d = [(0.3, 0, 0, 0.3), (0.4, 0, 0, 0.4), (0.6, 1, 1, 0.6), (0.5, 0, 1, 0.5), (0.4, 0, 0, 0.7), (0.5, 0, 0, 0.9)]
    assert oracle_lib.npi_active_damping(d, 0, 0, 0.6) == 0.4
    assert oracle_lib.npi_active_damping(d, 1, 1, 0.6) == 0.6
    assert oracle_lib.npi_active_damping(d, 1, 1, 0.4) == 0.0
    assert oracle_lib.npi_active_damping(d, 0, 0, 0.8) == 0.4
    assert oracle_lib.npi_active_damping([], 0, 0, 0.6) == 0.0


def _same(got, expected, rows):
    assert len(got) == len(expected)
    for (t, lvl, typ, coeffs), (v, elvl, etyp, et) in zip(got, expected):
        assert (lvl, typ) == (elvl, etyp)
        np.testing.assert_allclose(t, et, rtol=1e-15)
        np.testing.assert_allclose(coeffs, np.full(rows, v), rtol=1e-12, atol=1e-15)


def test_implement_empty():
    # TEST(DynamicNPIs, implement_empty), :145-170
    npi = dict(value=0.8, level=0, type=0, matrix_indices=[0, 1], group_weights=np.ones(2))
    out = oracle_lib.npi_implement([[], []], 2, [(npi, 0.45, 0.6)])
    for m in range(2):
        _same(out[m], [(0.8, 0, 0, 0.45), (0.0, 0, 0, 0.6)], 2)


def test_implement():
    # TEST(DynamicNPIs, implement), :172-260
    mats = [[(0.15, 0, 0, 0.3), (0.2, 0, 0, 0.4), (0.25, 0, 1, 0.5), (0.35, 0, 0, 0.7)],
            [(0.25, 1, 1, 0.5), (0.45, 0, 0, 0.9)]]
    npi1 = dict(value=0.4, level=0, type=0, matrix_indices=[0, 1], group_weights=np.ones(3))
    out = oracle_lib.npi_implement(mats, 3, [(npi1, 0.45, 0.6)])
    _same(out[0], [(0.15, 0, 0, 0.3), (0.2, 0, 0, 0.4), (0.4, 0, 0, 0.45), (0.25, 0, 1, 0.5), (0.2, 0, 0, 0.6),
                   (0.35, 0, 0, 0.7)], 3)
    _same(out[1], [(0.4, 0, 0, 0.45), (0.25, 1, 1, 0.5), (0.0, 0, 0, 0.6), (0.45, 0, 0, 0.9)], 3)
    npi2 = dict(value=0.3, level=0, type=0, matrix_indices=[0], group_weights=np.ones(3))
    out = oracle_lib.npi_implement(mats, 3, [(npi1, 0.45, 0.6), (npi2, 0.3, 0.9)])
    _same(out[0], [(0.3, 0, 0, 0.3), (0.4, 0, 0, 0.45), (0.25, 0, 1, 0.5), (0.3, 0, 0, 0.6), (0.35, 0, 0, 0.7)], 3)
    _same(out[1], [(0.4, 0, 0, 0.45), (0.25, 1, 1, 0.5), (0.0, 0, 0, 0.6), (0.45, 0, 0, 0.9)], 3)


def _mock_case(contact=1.0):
    # model(1): 10 InfectedSymptoms of 100, contact matrix constant
    G = 1
    y0 = np.zeros(10)
    y0[flat.ISY] = 10.0
    y0[flat.S] = 90.0
    pset = flat.osecir_pset(G)
    local = np.array([flat.DBL_MAX, flat.DBL_MAX])
    ct = flat.contact_table(G, [ContactLocation(np.full((1, 1), contact), None, [], [])], 1, 2)
    return dict(G=G, pset=pset, local=local, ct=ct, y0=y0)


def _sim(npi, t0, tmax, query, frozen=True, y_isy=10.0):
    s = _mock_case()
    s["y0"][flat.ISY] = y_isy
    s["y0"][flat.S] = 100.0 - y_isy
    return oracle_lib.npi_simulate(1, s["pset"], s["local"], s["ct"], 1, 2, s["y0"], t0, [tmax], 0.1, npi,
                                   frozen=frozen, query=query)


def _npi(value, **kw):
    return oracle_lib.npi_desc([(0.05 * 50000, [dict(value=value, level=0, type=0, matrix_indices=[0],
                                                       group_weights=np.ones(1))])], duration=5.0, base=50000.0, **kw)


def test_secir_threshold_safe_and_exceeded():
    # TEST(DynamicNPIs, secir_threshold_safe), :421-450: 1 of 100 stays below 5 %
    npi = oracle_lib.npi_desc([(0.05 * 23000, [dict(value=1.0, level=0, type=0, matrix_indices=[0],
                                                    group_weights=np.ones(1))])], duration=5.0, base=23000.0)
    assert _sim(npi, 0.0, 3.0, [], y_isy=1.0)["num_dampings"][0] == 0
    # TEST(DynamicNPIs, secir_threshold_exceeded), :452-482
    assert _sim(_npi(1.0), 0.0, 3.0, [])["num_dampings"][0] == 2


def test_secir_delayed_implementation():
    # TEST(DynamicNPIs, secir_delayed_implementation), :484-539
    r = _sim(_npi(0.5, delay=3.0), 0.0, 3.0, [0.0, 1.0, 3.0])
    np.testing.assert_array_equal(r["contact"][:, 0, 0], [0.5, 0.5, 0.5])
    r = _sim(_npi(0.5, delay=2.0), 1.0, 5.0, [3.0, 4.0])
    np.testing.assert_array_equal(r["contact"][:, 0, 0], [1.0, 0.5])
    r = _sim(_npi(0.5, delay=10.0), 1.0, 4.0, [11.0, 12.0])
    np.testing.assert_array_equal(r["contact"][:, 0, 0], [1.0, 0.5])


def test_secir_implementation_with_directives():
    # TEST(DynamicNPIs, secir_implementation_with_directives), :541-640
    r = _sim(_npi(0.5, directive_begin=5.0), 0.0, 3.0, [0.0, 1.0, 3.0])
    np.testing.assert_array_equal(r["contact"][:, 0, 0], [1.0, 1.0, 1.0])
    r = _sim(_npi(0.5, delay=2.0, directive_begin=0.0), 0.0, 3.0, [0.0, 4.0, 6.0])
    np.testing.assert_array_equal(r["contact"][:, 0, 0], [0.5, 0.5, 1.0])
    r = _sim(_npi(0.5, delay=2.0, directive_begin=0.0, directive_end=3.0), 0.0, 5.0, [0.0, 2.0, 4.0])
    np.testing.assert_array_equal(r["contact"][:, 0, 0], [0.5, 0.5, 1.0])
    r = _sim(_npi(0.5, delay=2.0, directive_begin=0.0, directive_end=1000000.0), 1.0, 5.0, [3.0, 4.0, 8.0, 9.0])
    np.testing.assert_array_equal(r["contact"][:, 0, 0], [1.0, 0.5, 0.5, 1.0])
    r = _sim(_npi(0.5, delay=2.0, directive_begin=0.0, directive_end=4.0), 1.0, 6.0, [1.0, 4.0, 5.0])
    np.testing.assert_array_equal(r["contact"][:, 0, 0], [1.0, 0.5, 1.0])


def test_real_simulation_with_npi_slows_the_epidemic():
    # not frozen: the contact reduction shows in the trajectory; day chunks end at integer days and at the NPI's end
    s = _mock_case(contact=10.0)
    npi = oracle_lib.npi_desc([(1000.0, [dict(value=0.8, level=0, type=0, matrix_indices=[0],
                                              group_weights=np.ones(1))])], duration=4.5, base=50000.0)
    none = oracle_lib.npi_desc([], duration=4.5, base=50000.0)
    a = oracle_lib.npi_simulate(1, s["pset"], s["local"], s["ct"], 1, 2, s["y0"], 0.0, [12.0], 0.1, npi)
    b = oracle_lib.npi_simulate(1, s["pset"], s["local"], s["ct"], 1, 2, s["y0"], 0.0, [12.0], 0.1, none)
    assert a["counters"][3] >= 2  # renewed after it expired
    assert a["y"][-1][flat.S] > b["y"][-1][flat.S] + 1.0
    # chunks: 1, 2, 3, 4, 4.5 (the NPI's end; renewed until 9), 5.5, ..., 8.5, 9 (its end again), 10, 11, 12
    assert {1.0, 4.0, 4.5, 5.5, 8.5, 9.0, 10.0, 12.0}.issubset(set(a["t"])) and 5.0 not in a["t"]
