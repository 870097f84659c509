"""Loaders (memilio_b200/loaders.py): the reference's real contact matrix files (copied data fixtures,
tests/golden/contacts = data/Germany/contacts of the reference) and set_edges / set_node_uncertainty against the
formulas of cpp/memilio/mobility/graph.h:346-446; the loaded graph samples and (on the GPU tests) simulates."""
import os

import numpy as np
import pytest

from memilio_b200 import flat, loaders, study, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
CONTACTS = os.path.join(HERE, "golden", "contacts")


def test_contact_matrices_of_the_reference_data_directory():
    base, mini = loaders.read_contact_matrices(CONTACTS)
    assert base.shape == mini.shape == (4, 6, 6)
    # first row of baseline_home.txt / values the reference's Germany scenarios are built on
    np.testing.assert_array_equal(base[0, 0], [0.4413, 0.4504, 1.2383, 0.8033, 0.0494, 0.0017])
    assert (base >= 0).all() and (mini >= 0).all()


def test_read_mobility_plain_rejects_non_square(tmp_path):
    f = tmp_path / "m.txt"
    f.write_text("0 1 2\n3 0 4\n")
    with pytest.raises(ValueError, match="Not a square matrix"):
        loaders.read_mobility_plain(str(f))
    f.write_text("0 10.5\n3 0\n")
    np.testing.assert_array_equal(loaders.read_mobility_plain(str(f)), [[0, 10.5], [3, 0]])


def test_set_edges_threshold_weights_and_order():
    G, C, N = 6, 10, 4
    rng = np.random.default_rng(3)
    pop = np.zeros((N, G, C))
    pop[:, :, flat.S] = rng.uniform(1e3, 5e4, (N, G))
    w = np.array([0.0, 0.0, 1.0, 1.0, 0.33, 0.0])
    mob = rng.uniform(0, 50, (N, N))
    mob[0, 3] = 1e-3  # below the threshold: no edge
    np.fill_diagonal(mob, 0.0)
    es, ee, coeff = loaders.set_edges(pop, mob, [flat.S, flat.E, flat.INS, flat.ISY, flat.R], w)
    assert list(zip(es, ee)) == sorted(zip(es, ee)) and (0, 3) not in set(zip(es, ee))
    for e, (i, j) in enumerate(zip(es, ee)):
        working = sum(pop[i, a].sum() * w[a] for a in range(G))
        c = mob[i, j] / working
        assert c > 4e-5
        v = coeff[e].reshape(G, C)
        np.testing.assert_allclose(v[:, flat.S], c * w, rtol=1e-15)
        assert (v[:, [flat.INSC, flat.ISYC, flat.ISEV, flat.ICR, flat.D]] == 0).all()
        np.testing.assert_array_equal(v[:, flat.E], v[:, flat.R])


def test_loaded_graph_samples_like_any_study_graph(tmp_path):
    G, C, N = 6, 10, 5
    pop, _, _, _ = synthetic.county_graph(N, 7, 4, G)
    y = np.zeros((N, G, C))
    y[:, :, flat.S] = pop
    y[:, :, flat.E] = 10.0
    rng = np.random.default_rng(1)
    mob = np.round(rng.uniform(0, 400, (N, N)))
    np.fill_diagonal(mob, 0)
    mfile = tmp_path / "commuter_mobility.txt"
    mfile.write_text("\n".join(" ".join(f"{v:.1f}" for v in row) for row in mob) + "\n")
    ref = synthetic.covid_study_graph(N=N, max_out_degree=4)
    g = loaders.study_graph_from_files(y, str(mfile), CONTACTS, ref.header, ref.age_params, np.full(N, 1e9), 7.5e-5)
    assert g.baseline.shape == (4, G, G) and g.mobility_location == 2
    ps = study.ParameterStudy(g, 0.0, 5.0, 0.5, 3)
    p = ps.sample(0, 3)
    assert p.num_edges == len(g.edge_start) and p.ct_max_loc == 4
    # TestAndTraceCapacity ~ U(0.8, 1.2) * 7.5e-5 * population; compartments +-10 %
    cap = y.reshape(N, -1).sum(axis=1) * 7.5e-5
    assert ((p.node_local[:, :, 1] >= 0.8 * cap - 1e-12) & (p.node_local[:, :, 1] <= 1.2 * cap + 1e-12)).all()
    tot = p.y0.reshape(3, N, G, C).sum(axis=3)
    np.testing.assert_allclose(tot, np.broadcast_to(y.sum(axis=2), tot.shape), rtol=1e-12)  # group totals survive


@pytest.mark.gpu
def test_loaded_germany_style_graph_simulates_like_the_oracle(tmp_path):
    # SURVEY 8d config 1 variant / 8f-4: the four Germany contact locations (L = 4) with their minimum matrices, a
    # commuter matrix file -> edges, sampled by ParameterStudy, on the GPU against the CPU oracle
    import oracle_lib
    from memilio_b200 import capi
    G, C, N = 6, 10, 6
    pop, _, _, _ = synthetic.county_graph(N, 11, 4, G)
    y = np.zeros((N, G, C))
    y[:, :, flat.S] = pop
    y[:2, :, flat.E] = 1e-3 * pop[:2]
    rng = np.random.default_rng(5)
    mob = np.round(rng.uniform(0, 3000, (N, N)))
    np.fill_diagonal(mob, 0)
    mfile = tmp_path / "commuter_mobility.txt"
    mfile.write_text("\n".join(" ".join(f"{v:.1f}" for v in row) for row in mob) + "\n")
    ref = synthetic.covid_study_graph(N=N, max_out_degree=4)
    dv = study.uv_array(1)
    study.set_uniform(dv, 0, 0.3, 0.6)
    g = loaders.study_graph_from_files(y, str(mfile), CONTACTS, ref.header, ref.age_params, np.full(N, 1e9), 7.5e-5,
                                       dampings=[study.Damping(value=dv, time=2.0, matrices=(1, 2))])
    ps = study.ParameterStudy(g, 0.0, 6.0, 0.5, 3)
    p = ps.sample(0, 3)
    with capi.GraphSimulation(p) as sim:
        sim.advance(6.0)
        state, cnt = sim.state(), sim.counters()
    for r in range(3):
        o = oracle_lib.graph_simulate(p, r, 6.0)
        np.testing.assert_allclose(state[r], o["state"], rtol=1e-8, atol=1e-9)
        np.testing.assert_array_equal(cnt[r][:, :3], o["counters"][:, :3])
