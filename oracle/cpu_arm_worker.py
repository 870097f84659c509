#!/usr/bin/env python
"""cpu_arm_worker.py -- one process ("MPI rank") of the CPU reference arm (test infrastructure, NOT product code).

The reference's ParameterStudy::run shards the runs over MPI ranks with distribute_runs
(cpp/memilio/compartments/parameter_studies.h:163-213,354-367); MPI is absent here, so bench.py starts one of these
processes per host core. Each times its own literal run_impl loop on the CPU oracle (oracle/): draw_sample(graph) +
GraphSimulation::advance + interpolate_simulation_result for every run of its block, and writes the final states and
step counters (the parity vectors) plus its loop time to an .npz file.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_study(model, nodes, days, sim_kind, total_runs, flow_sim=0):
    from memilio_b200 import flat, study, synthetic
    if model == "osecirvvs":
        graph = synthetic.vvs_study_graph(N=nodes, days=int(days) + 1,
                                          sim_kind=flat.SIM_MODEL if sim_kind == "model" else flat.SIM_BASE)
    else:
        graph = synthetic.covid_study_graph(N=nodes)
    ps = study.ParameterStudy(graph, 0.0, float(days), 0.5, total_runs)
    ps.flow_simulation = int(flow_sim)  # the nodes are mio::FlowSimulation (compartments/flow_simulation.h:37-97)
    return ps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="osecir")
    ap.add_argument("--sim-kind", default="model")
    ap.add_argument("--flow-sim", type=int, default=0)
    ap.add_argument("--nodes", type=int, default=400)
    ap.add_argument("--days", type=float, default=100.0)
    ap.add_argument("--sample-runs", type=int, required=True, help="runs of the whole CPU sample (all workers)")
    ap.add_argument("--study-runs", type=int, required=True, help="num_runs of the study the sample is the head of")
    ap.add_argument("--workers", type=int, required=True)
    ap.add_argument("--rank", type=int, required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--start-file", default=None, help="the timed loop starts when this file exists (common start)")
    a = ap.parse_args()
    import numpy as np
    import oracle_lib
    ps = build_study(a.model, a.nodes, a.days, a.sim_kind, a.study_runs, a.flow_sim)
    problem = oracle_lib.sample_problem(ps, 0, 1)  # graph topology and capacities for the oracle's descriptors
    # distribute_runs(sample_runs, workers, rank), parameter_studies.h:354-367
    base, rem = divmod(a.sample_runs, a.workers)
    count = base + (1 if a.rank < rem else 0)
    run0 = a.rank * base + min(a.rank, rem)
    num_days = int(a.days) + 1
    np.savez(a.out + ".ready.npz", ready=1)
    if a.start_file:
        while not os.path.exists(a.start_file):
            time.sleep(0.005)
    t0 = time.perf_counter()
    res = oracle_lib.study_run_block(ps, problem, run0, count, float(a.days), num_days) if count > 0 else None
    dt = time.perf_counter() - t0
    if res is None:
        n = problem.n
        res = dict(state=np.zeros((0, a.nodes, n)), counters=np.zeros((0, a.nodes, 4), dtype=np.int64), checksum=0.0)
    np.savez(a.out, state=res["state"], counters=res["counters"], checksum=res["checksum"], seconds=dt, run0=run0,
             count=count)


if __name__ == "__main__":
    main()
