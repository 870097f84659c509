"""Parity of the CUDA path (through the C-ABI) against the CPU oracle and the reference's golden vectors.
Tolerance: compartment trajectories within 1e-8 relative (FP64, same tableau and step control); accepted / rejected
step counts equal (BASELINE.json north_star)."""
import numpy as np
import pytest

import oracle_lib
import scenarios
from memilio_b200 import capi, flat, synthetic

pytestmark = pytest.mark.gpu

RTOL = 1e-8


def assert_close(a, b, rtol=RTOL, atol=1e-9, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert np.all(err <= 0), f"{what}: max violation {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"


@pytest.mark.parametrize("name", list(scenarios.GOLDEN_CASES))
def test_single_node_trajectories_match_golden_and_oracle(name):
    s = scenarios.golden_case(name)
    p = scenarios.golden_problem(name)
    t_o, y_o, dt_o, cnt_o = oracle_lib.simulate(s["G"], s["pset"], s["local"], s["ct"], s["ct_max_loc"],
                                                s["ct_max_pts"], s["y0"], s["t0"], s["tmax"], s["dt0"],
                                                s["integrator"])
    with capi.GraphSimulation(p, traj_capacity=128) as sim:
        sim.advance(s["tmax"])
        npts, times, values = sim.trajectory()
        cnt = sim.counters()[0, 0]
        dt = sim.node_dt()[0, 0]
    n = int(npts[0, 0])
    ref = s["csv"]
    # identical accepted step count = identical number of time points as the reference's golden file
    assert n == ref.shape[0] == len(t_o)
    assert cnt[0] == cnt_o[0] and cnt[1] == cnt_o[1] and cnt[2] == cnt_o[2]
    # The adaptive time points depend on err^(-1/5) where err is a cancellation-prone difference of two RK solutions
    # (|xerr| ~ 1e-6 |dt k|): ulp-level differences (FMA contraction, libdevice vs glibc sin/cos/pow) are amplified
    # to ~1e-10 relative in dt. Time points and values are therefore held to the north-star tolerance (1e-8 relative),
    # the number of accepted/rejected steps to equality.
    assert_close(times[0, 0, :n], ref[:, 0], rtol=1e-8, atol=1e-10, what="time points vs golden")
    ys = scenarios.age_sum(values[0, 0, :n], s["G"])
    assert_close(ys, ref[:, 1:], rtol=1e-8, atol=1e-8, what="trajectory vs golden")
    assert_close(values[0, 0, :n], y_o, what="trajectory vs oracle")
    assert_close(times[0, 0, :n], t_o, rtol=1e-8, atol=1e-10, what="time points vs oracle")
    print(name, "max |dt| vs oracle", np.abs(times[0, 0, :n] - t_o).max(), "max rel dy",
          (np.abs(values[0, 0, :n] - y_o) / (np.abs(y_o) + 1e-9)).max())
    assert abs(dt - dt_o) <= 1e-7 * abs(dt_o)


def test_example_ode_secir_graph_every_half_day():
    # BASELINE config 2: cpp/examples/ode_secir_graph.cpp
    p = scenarios.example_ode_secir_graph()
    ref = oracle_lib.graph_simulate(p, 0, 30.0, halfsteps=60)
    with capi.GraphSimulation(p) as sim:
        for k in range(60):
            sim.advance(0.5 * (k + 1))
            assert_close(sim.state()[0], ref["halfsteps"][k], what=f"state after graph step {k}")
        cnt = sim.counters()[0]
        assert sim.t == 30.0
        np.testing.assert_array_equal(cnt[:, :3], ref["counters"][:, :3])
        np.testing.assert_array_equal(cnt[:, 3], ref["counters"][:, 3])
        assert_close(sim.node_dt()[0], ref["dt"], rtol=1e-8, what="node dt")


def test_example_graph_single_advance_equals_stepwise():
    p = scenarios.example_ode_secir_graph()
    ref = oracle_lib.graph_simulate(p, 0, 30.0)
    with capi.GraphSimulation(p) as sim:
        sim.advance(30.0)
        assert_close(sim.state()[0], ref["state"], what="final state")


@pytest.mark.parametrize("G,N", [(6, 40), (3, 25), (1, 37), (2, 16)])
def test_county_graph_ensemble_matches_oracle(G, N):
    # a small Germany-shaped graph, several runs with different parameters, seeded so that the first days hit the
    # return-phase underflow fix-up (order dependent path) and the contact damping ramp at day 5
    runs = 5
    p = synthetic.county_problem(N=N, runs=runs, G=G, max_out_degree=8, n_seeded=3, damping_day=5.0, perturb=0.1)
    tmax = 12.0
    with capi.GraphSimulation(p, record_daily=True, max_days=13) as sim:
        sim.advance(tmax)
        state = sim.state()
        cnt = sim.counters()
        dts = sim.node_dt()
        daily = sim.daily()
        st = sim.stats()
    assert daily.shape == (runs, N, 13, p.n)
    total_underflow = 0
    for r in range(runs):
        ref = oracle_lib.graph_simulate(p, r, tmax, halfsteps=24)
        assert_close(state[r], ref["state"], what=f"run {r} final state")
        np.testing.assert_array_equal(cnt[r][:, :3], ref["counters"][:, :3])
        # fix-ups are applied (and counted) by the literal replay of flagged steps: per node, equal to the oracle's
        np.testing.assert_array_equal(cnt[r][:, 3], ref["counters"][:, 3])
        assert_close(dts[r], ref["dt"], rtol=1e-7, what="node dt")
        # daily record = state after the edge phase at integer days (what interpolate_simulation_result returns)
        np.testing.assert_array_equal(daily[r][:, 0], p.y0[r])
        for d in range(1, 13):
            assert_close(daily[r][:, d], ref["halfsteps"][2 * d - 1], what=f"run {r} day {d}")
        total_underflow += int(ref["counters"][:, 3].sum())
        # population is conserved by construction of the exchange
        assert abs(state[r].sum() - p.y0[r].sum()) < 1e-6 * p.y0[r].sum()
    if G == 6:
        assert total_underflow > 0 and st["exact_replays"] > 0  # the order dependent path was exercised


def _variant(name):
    G, N, runs = 6, 30, 3
    p = synthetic.county_problem(N=N, runs=runs, G=G, max_out_degree=6, n_seeded=4, damping_day=4.0, perturb=0.1)
    rng = np.random.default_rng(7)
    if name == "dense_coefficients":
        # arbitrary per-compartment coefficients: not of the form w[e][age] * mask
        p.edge_coeff *= rng.uniform(0.5, 1.5, p.edge_coeff.shape)
        p.edge_coeff[:, flat.ISEV::10] = 1e-3 * rng.uniform(0.5, 1.5, (p.num_edges, G))
    elif name == "commuter_testing":
        # test_commuters active on days [2, 7): ode_secir/model.h:717-742
        p.psets[:, :, flat.PH["commuter_nondetection"]] = 0.4
        p.psets[:, :, flat.PH["start_commuter_detection"]] = 2.0
        p.psets[:, :, flat.PH["end_commuter_detection"]] = 7.0
        p.y0[:, :, flat.INS::10] += 5.0
        p.y0[:, :, flat.ISY::10] += 3.0
    elif name == "mobility_damping":
        # dampings on the mobility coefficients (parameter_space.h:241-248), change points at t=3 and t=6.5
        p.mob_max_pts = 3
        tabs = [flat.mobility_table(p.n, [3.0, 6.5], [np.full(p.n, 0.5 + 0.1 * r), rng.uniform(0.1, 0.9, p.n)], 3)
                for r in range(runs)]
        p.mobtables = np.stack(tabs)
    elif name == "two_parameter_sets":
        ps2 = p.psets.copy()
        ps2[:, :, flat.PSET_HEADER:] *= rng.uniform(0.8, 1.2, ps2[:, :, flat.PSET_HEADER:].shape)
        for r in range(runs):
            flat.osecir_apply_constraints(ps2[r, 0], G)
        p.psets = np.ascontiguousarray(np.concatenate([p.psets, ps2], axis=1))
        p.node_pset = (np.arange(N) % 2).astype(np.int32)
        ct2 = p.ctables.copy()
        p.ctables = np.ascontiguousarray(np.concatenate([p.ctables, ct2], axis=1))
        p.node_ctable = ((np.arange(N) // 2) % 2).astype(np.int32)
    elif name == "icu_and_tnt_saturated":
        # small capacities: the cosine smoothers of ICU shortage and test-and-trace are inside / beyond their ramps
        p.node_local[:, :, 0] = p.y0.sum(axis=2) * 2e-6
        p.node_local[:, :, 1] = p.y0.sum(axis=2) * 1e-6
        p.y0[:, :, flat.ICR::10] += p.y0[:, :, flat.S::10] * 1e-6
        p.y0[:, :, flat.INS::10] += p.y0[:, :, flat.S::10] * 2e-6
    p.__post_init__()
    return p


@pytest.mark.parametrize("name", ["dense_coefficients", "commuter_testing", "mobility_damping", "two_parameter_sets",
                                  "icu_and_tnt_saturated"])
def test_edge_phase_variants_match_oracle(name):
    p = _variant(name)
    tmax = 9.0
    with capi.GraphSimulation(p) as sim:
        sim.advance(tmax)
        state = sim.state()
        cnt = sim.counters()
    for r in range(p.num_runs):
        ref = oracle_lib.graph_simulate(p, r, tmax)
        assert_close(state[r], ref["state"], what=f"{name} run {r}")
        np.testing.assert_array_equal(cnt[r][:, :3], ref["counters"][:, :3])


def test_full_size_county_graph_single_run():
    # BASELINE config 3: 400 counties, ~10k edges, one run, 100 days
    p = synthetic.county_problem(N=400, runs=1)
    assert p.num_edges == 10000
    ref = oracle_lib.graph_simulate(p, 0, 100.0)
    with capi.GraphSimulation(p) as sim:
        sim.advance(100.0)
        state = sim.state()[0]
        cnt = sim.counters()[0]
    assert_close(state, ref["state"], what="config 3 final state")
    mism = int((cnt[:, :3] != ref["counters"][:, :3]).any(axis=1).sum())
    assert mism == 0, f"{mism} nodes with different accepted/rejected/forced step counts"
    np.testing.assert_array_equal(cnt[:, 3], ref["counters"][:, 3])  # underflow fix-ups per node


def test_error_behaviour():
    p = scenarios.example_ode_secir_graph()
    with capi.GraphSimulation(p) as sim:
        with pytest.raises(capi.MioCudaError) as e:
            sim.set_runs(1, p, 0, 1)  # slot out of range
        assert e.value.code == 2
    p.edge_start[:] = [1, 0]
    p.edge_end[:] = [0, 1]
    with pytest.raises(capi.MioCudaError) as e:
        capi.GraphSimulation(p)  # edges not in reference order
    assert e.value.code == 3


@pytest.mark.parametrize("G,N", [(6, 24), (1, 2)])
def test_mobility_results_match_oracle(G, N):
    # MobilityEdge::get_mobility_results (instant.h:421-445): per edge and graph step the sums of the mobile population
    # over the saved index groups and its total, departure and return (after the underflow fix-up) alike
    if N == 2:
        p = scenarios.example_ode_secir_graph(runs=2)
        tmax = 30.0
    else:
        p = synthetic.county_problem(N=N, runs=3, G=G, max_out_degree=5, n_seeded=3, damping_day=3.0, perturb=0.1)
        tmax = 8.0
    save = [[a * 10 + flat.INS for a in range(G)] + [a * 10 + flat.INSC for a in range(G)],
            [a * 10 + flat.ISY for a in range(G)] + [a * 10 + flat.ISYC for a in range(G)]]
    steps = int(round(tmax / 0.5))
    with capi.GraphSimulation(p) as sim:
        sim.enable_mobility_results(save, steps + 2)
        sim.advance(tmax / 2)
        sim.enable_mobility_results(save, 2 * steps + 5)  # growing the capacity keeps what was recorded
        sim.advance(tmax)
        npts, times, values = sim.mobility_results()
        state = sim.state()
        cnt = sim.counters()
    fixups = 0
    for r in range(p.num_runs):
        o_npts, o_times, o_values, o_state = oracle_lib.graph_simulate_mobility(p, r, tmax, save, cap=2 * steps + 5)
        np.testing.assert_array_equal(npts[r], o_npts)
        assert (o_npts == steps).all()
        np.testing.assert_array_equal(times[:steps], o_times[0, :steps])
        assert_close(values[r][:, :steps], o_values[:, :steps], rtol=1e-8, atol=1e-9, what=f"run {r} edge results")
        assert_close(state[r], o_state, what="node states with every return step replayed literally")
        fixups += int(cnt[r][:, 3].sum())
    if G == 6:
        assert fixups > 0  # returns with the order dependent fix-up were recorded


def test_block_parallel_exact_replay_matches_oracle():
    # MIO_EXACT_REPLAY=blocks: the literal replay parallel over the edges of one source node
    # (return_exact_blocks_kernel, not the default: see DESIGN.md section 9) is bit-identical in its fix-up decisions
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, 'tests'); import oracle_lib\n"
        "from memilio_b200 import capi, synthetic\n"
        "p = synthetic.county_problem(N=40, runs=5, G=6, max_out_degree=8, n_seeded=3, damping_day=5.0, perturb=0.1)\n"
        "sim = capi.GraphSimulation(p); sim.advance(12.0); st = sim.state(); cnt = sim.counters(); s = sim.stats()\n"
        "tot = 0\n"
        "for r in range(5):\n"
        "    ref = oracle_lib.graph_simulate(p, r, 12.0)\n"
        "    np.testing.assert_allclose(st[r], ref['state'], rtol=1e-8, atol=1e-9)\n"
        "    np.testing.assert_array_equal(cnt[r], ref['counters'])\n"
        "    tot += int(ref['counters'][:, 3].sum())\n"
        "assert tot > 0 and s['exact_replays'] > 0\n"
        "print('ok', tot, s['exact_replays'])\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, MIO_EXACT_REPLAY="blocks"),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert r.stdout.startswith("ok")
