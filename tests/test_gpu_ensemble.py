"""Ensemble post-processing on the device (sum_nodes / ensemble_mean / ensemble_percentile,
cpp/memilio/data/analyze_result.h:267-343) against numpy on the same recorded values, and against the exact small
cases of the reference's own tests (cpp/tests/test_analyze_result.cpp:196-371). Percentiles are order statistics of
recorded values: bit-exact. Mean: 1e-13 relative (summation order differs from the reference's run order)."""
import numpy as np
import pytest

from memilio_b200 import capi, ensemble, study, synthetic

pytestmark = pytest.mark.gpu


def test_reference_percentile_small_case():
    # TestEnsemblePercentile.basic, cpp/tests/test_analyze_result.cpp:300-371: 4 runs, 2 nodes, 3 time points, 2 elements
    import torch
    runs = np.zeros((4, 2, 3, 2))
    runs[0, 1, 0, 0] = 0.2
    runs[1, 0, 1, 1], runs[1, 1, 0, 0] = 1.0, 0.1
    runs[2, 0, 1, 1], runs[2, 1, 0, 0] = 2.0, 0.3
    runs[3, 0, 1, 1] = 3.0
    cols = torch.tensor(runs.reshape(4, -1).T.copy(), device="cuda")  # [column][run]
    pct, mean = ensemble.select_cuda([cols], [4], [0.2, 0.4, 0.7, 0.9], True, 0)
    q = pct.reshape(4, 2, 3, 2)
    assert [q[i, 0, 1, 1] for i in range(4)] == [0.0, 1.0, 2.0, 3.0]
    assert [q[i, 1, 0, 0] for i in range(4)] == [0.0, 0.1, 0.2, 0.3]
    np.testing.assert_allclose(mean, runs.reshape(4, -1).mean(axis=0), rtol=1e-15)


def test_reference_mean_small_case():
    # TestEnsembleMean.basic, cpp/tests/test_analyze_result.cpp:266-298
    import torch
    runs = np.array([[[0.0, 1.0, 2.0], [0.0, 1.0, 2.0]], [[0.5, 3.0, 0.0], [1.5, 0.5, 1.0]]])
    cols = torch.tensor(runs.reshape(2, -1).T.copy(), device="cuda")
    _, mean = ensemble.select_cuda([cols], [2], [0.5], True, 0)
    np.testing.assert_array_equal(mean.reshape(2, 3), [[0.25, 2.0, 1.0], [0.75, 0.75, 1.5]])


def test_select_handles_negative_values_ties_and_segments():
    import torch
    rng = np.random.default_rng(0)
    R, ncol = 1000, 37
    data = rng.normal(0, 1, (ncol, R))
    data[:, ::7] = 0.0
    data[3] = -2.5
    data[4, :500] = -0.0
    segs = [torch.tensor(np.ascontiguousarray(data[:, a:b]), device="cuda") for a, b in ((0, 100), (100, 101), (101, 1000))]
    p = [0.05, 0.25, 0.5, 0.75, 0.95, 0.999]
    pct, mean = ensemble.select_cuda(segs, [100, 1, 899], p, True, 0)
    srt = np.sort(data, axis=1)
    for i, q in enumerate(p):
        np.testing.assert_array_equal(pct[i], srt[:, int(R * q)])
    np.testing.assert_allclose(mean, data.mean(axis=1), rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("mode", [ensemble.ENS_AGE_SUM, ensemble.ENS_FULL])
def test_study_record_and_reduce(mode):
    G, N, runs, T = 6, 14, 37, 6
    g = synthetic.covid_study_graph(N=N, max_out_degree=5, n_seeded=3, damping_day=3.0)
    ps = study.ParameterStudy(g, 0.0, float(T), 0.5, runs)
    p = ps.sample(0, runs)
    rec = ensemble.EnsembleRecord(G, N, T + 1, capacity_runs=40, mode=mode)
    # two batches of different size through the same handle, like ParameterStudy.run does
    daily = np.zeros((runs, N, T + 1, p.n))
    with capi.GraphSimulation(p, max_runs=20, record_daily=True, max_days=T + 1, upload=False) as sim:
        for run0, cnt in ((0, 20), (20, 17)):
            sim.reset()
            sim.set_active_runs(cnt)
            rec.attach(sim, run0)
            sim.set_runs(0, p, run0, cnt)
            sim.advance(float(T))
            daily[run0:run0 + cnt] = sim.daily(0, cnt)
    info = rec.info()
    assert info["num_runs"] == runs and info["num_columns"] == N * (T + 1) * rec.El + (T + 1) * p.n
    pct, mean = rec.reduce()
    node_pct, sum_pct = rec.split(pct)
    node_mean, sum_mean = rec.split(mean)
    # expected from the (oracle-verified) daily record with the reference's arithmetic order
    d = daily.reshape(runs, N, T + 1, G, 10)
    if mode == ensemble.ENS_AGE_SUM:
        per_node = np.zeros((runs, N, T + 1, 10))
        for a in range(G):
            per_node = per_node + d[:, :, :, a, :]
    else:
        per_node = daily
    node_sum = np.zeros((runs, T + 1, p.n))
    for k in range(N):
        node_sum = node_sum + daily[:, k]
    for i, q in enumerate(ensemble.DEFAULT_PERCENTILES):
        idx = int(runs * q)
        np.testing.assert_array_equal(node_pct[i], np.sort(per_node, axis=0)[idx])
        np.testing.assert_array_equal(sum_pct[i], np.sort(node_sum, axis=0)[idx])
    np.testing.assert_allclose(node_mean, per_node.mean(axis=0), rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(sum_mean, node_sum.mean(axis=0), rtol=1e-13)
    # single-rank path of the sharded reduction gives the same answer
    pct2, mean2 = ensemble.reduce_distributed(rec.as_torch(), runs, 0, 1)
    np.testing.assert_array_equal(pct2, pct)
    np.testing.assert_array_equal(mean2, mean)
    rec.close()


def test_parameter_study_run_matches_manual_batches():
    g = synthetic.covid_study_graph(N=10, max_out_degree=4, n_seeded=2, damping_day=2.0)
    ps = study.ParameterStudy(g, 0.0, 4.0, 0.5, 11)
    res = ps.run(batch=4)
    assert len(res) == 11 and res[0].shape == (10, 5, 60)
    assert ps.rng_counter == 11 << 32
    ps2 = study.ParameterStudy(g, 0.0, 4.0, 0.5, 11)
    p = ps2.sample(0, 11)
    with capi.GraphSimulation(p, record_daily=True, max_days=5) as sim:
        sim.advance(4.0)
        daily = sim.daily()
    for r in range(11):
        np.testing.assert_array_equal(res[r], daily[r])
    # a second run() continues the random sequence (parameter_studies.h:339)
    res2 = ps.run(batch=11)
    assert not np.array_equal(res2[0], res[0])
    # ranks take contiguous blocks of the same global run sequence
    ps3 = study.ParameterStudy(g, 0.0, 4.0, 0.5, 11)
    part = ps3.run(batch=8, rank=1, world=2)
    assert len(part) == 5
    for r in range(5):
        np.testing.assert_array_equal(part[r], daily[6 + r])


def test_run_ensemble_matches_record_and_reduce():
    # the pipelined public call (sampling thread, two handles, on-device reduction) against one handle + reduce
    G, N, runs, T = 6, 12, 29, 5
    g = synthetic.covid_study_graph(N=N, max_out_degree=4, n_seeded=3, damping_day=2.0)
    ps = study.ParameterStudy(g, 0.0, float(T), 0.5, runs)
    pct, mean = ps.run_ensemble(num_batches=3)
    assert ps.rng_counter == runs << 32
    ps2 = study.ParameterStudy(g, 0.0, float(T), 0.5, runs)
    p = ps2.sample(0, runs)
    rec = ensemble.EnsembleRecord(G, N, T + 1, capacity_runs=runs, mode=ensemble.ENS_AGE_SUM)
    with capi.GraphSimulation(p, upload=False) as sim:
        rec.attach(sim, 0)
        sim.set_runs(0, p, 0, runs)
        sim.advance(float(T))
    pct2, mean2 = rec.reduce()
    np.testing.assert_array_equal(pct, pct2)
    np.testing.assert_array_equal(mean, mean2)
    # a second call continues the random sequence and reuses the cached handles
    pct3, _ = ps.run_ensemble(num_batches=3)
    assert not np.array_equal(pct3, pct)
    ps.close()
    rec.close()


def test_full_size_ensemble_properties():
    # BASELINE config 4 at full size (10 000 runs x 400 counties x 100 days): size independent properties, and eight
    # of the runs against the CPU oracle (states 1e-8, step counts equal)
    import oracle_lib
    R, N, T = 10000, 400, 100.0
    g = synthetic.covid_study_graph(N=N)
    ps = study.ParameterStudy(g, 0.0, T, 0.5, R)
    p = ps.sample(0, R)
    total0 = p.y0.sum(axis=(1, 2))
    with capi.GraphSimulation(p) as sim:
        sim.advance(T)
        state = sim.state()
        cnt = sim.counters()
    # the population of a run is conserved by every flow and by the commuter exchange (Dead included)
    np.testing.assert_allclose(state.sum(axis=(1, 2)), total0, rtol=1e-11)
    assert np.isfinite(state).all() and state.min() >= -1e-9
    # every node took at least the 200 graph steps; nothing had to be forced
    assert cnt[:, :, 0].min() >= 200 and cnt[:, :, 2].sum() == 0
    # a run does not depend on the ensemble it is computed in: the same runs alone, bit for bit
    pick = [0, 1, 2, 3, 4999, 5000, 9998, 9999]
    import copy
    q = copy.copy(p)
    for name in ("psets", "ctables", "node_local", "y0", "mobtables"):
        a = getattr(p, name)
        if a is not None:
            setattr(q, name, np.ascontiguousarray(a[pick]))
    with capi.GraphSimulation(q) as sim:
        sim.advance(T)
        alone = sim.state()
        cnt_alone = sim.counters()
    np.testing.assert_array_equal(alone, state[pick])
    np.testing.assert_array_equal(cnt_alone, cnt[pick])
    # and all eight against the CPU oracle over the full 100 days
    from concurrent.futures import ThreadPoolExecutor
    oracle_lib.lib()
    with ThreadPoolExecutor(max_workers=8) as ex:  # ctypes releases the GIL
        refs = list(ex.map(lambda r: oracle_lib.graph_simulate(q, r, T), range(len(pick))))
    for r, ref in enumerate(refs):
        np.testing.assert_allclose(alone[r], ref["state"], rtol=1e-8, atol=1e-9)
        np.testing.assert_array_equal(cnt_alone[r][:, :3], ref["counters"][:, :3])


def test_sharded_select_equals_single_gpu_reduce():
    # the N-GPU reduction (runs -> columns exchange, then multi-segment selection per column block, gather) emulated on
    # one device: the record of 3 "ranks" with uneven run counts, every rank's column block selected over the three run
    # segments exactly as ensemble.reduce_distributed does after exchange_columns; bit-equal to the one-GPU reduce
    import torch
    G, N, runs, T = 6, 12, 41, 6
    g = synthetic.covid_study_graph(N=N, max_out_degree=4, n_seeded=3, damping_day=2.0)
    ps = study.ParameterStudy(g, 0.0, float(T), 0.5, runs)
    p = ps.sample(0, runs)
    rec = ensemble.EnsembleRecord(G, N, T + 1, capacity_runs=runs, mode=ensemble.ENS_AGE_SUM)
    with capi.GraphSimulation(p, upload=False) as sim:
        rec.attach(sim, 0)
        sim.set_runs(0, p, 0, runs)
        sim.advance(float(T))
    pct1, mean1 = rec.reduce()
    full = rec.as_torch()[:, :runs]
    ncol = full.shape[0]
    world = 3
    blocks = [study.distribute_runs(runs, world, r) for r in range(world)]
    local = [full[:, a:a + c].contiguous() for a, c in blocks]  # what each rank recorded
    pct = np.zeros_like(pct1)
    mean = np.zeros_like(mean1)
    for rank, (c0, cn) in enumerate(ensemble.column_partition(ncol, world)):
        segs = [local[s][c0:c0 + cn].contiguous() for s in range(world)]  # received from rank s
        pr, mr = ensemble.select_cuda(segs, [c for _, c in blocks], ensemble.DEFAULT_PERCENTILES, True, 0)
        pct[:, c0:c0 + cn] = pr
        mean[c0:c0 + cn] = mr
    np.testing.assert_array_equal(pct, pct1)
    np.testing.assert_allclose(mean, mean1, rtol=1e-14, atol=1e-300)
    rec.close()


def test_windowed_record_equals_full_record():
    # a record sized for 3 days, advanced and reduced window by window (what a study does whose full record does not
    # fit the device), against the record of all days: every percentile and the mean, bit for bit
    G, N, runs, T = 6, 12, 23, 7
    g = synthetic.covid_study_graph(N=N, max_out_degree=4, n_seeded=3, damping_day=2.0)
    ps = study.ParameterStudy(g, 0.0, float(T), 0.5, runs)
    pct, mean = ps.run_ensemble(num_batches=2)
    ps.close()
    for blk in (3, 8, 1):
        ps2 = study.ParameterStudy(g, 0.0, float(T), 0.5, runs)
        pct2, mean2 = ps2.run_ensemble(day_block=blk)
        np.testing.assert_array_equal(pct2, pct)
        np.testing.assert_array_equal(mean2, mean)
        # a second call reuses the handle and the record and continues the random sequence
        pct3, _ = ps2.run_ensemble(day_block=blk)
        assert not np.array_equal(pct3, pct)
        ps2.close()
