#!/usr/bin/env python
"""bench.py -- Graph-ODE-SECIR ensemble throughput on B200 (BASELINE.json metric: node-days/s = runs x counties x days).

One "step" = one pass of the hot path over one batch: R runs of the synthetic 400-county ODE-SECIR graph (6 age groups,
~10k commuter edges, dt 0.5) advanced over T days with GraphSimulation::advance semantics (adaptive Cash-Karp node
integration + commuter departure/return). `value` times the device work with the inputs resident in HBM; `e2e` times the
same step through the C-ABI with host buffers (H2D of the sampled runs, advance, D2H of the result).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

For N > 1 launch under torchrun (one rank per GPU); runs shard over ranks with no data-path collective ("weak").
`--impl reference` times the reference's CPU path (the oracle port of it, see oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work per unit, see DESIGN.md "Roofline"
# SURVEY.md 8d. osecir: 6 RHS (650) + stage axpys (1800) + solution (720) + error (720) + norm (540);
# osecirvvs: 6 RHS (1.9k) + 4860 + 1944 + 1944 + 1460
FLOP_PER_CK54_ATTEMPT = {"osecir": 7.7e3, "osecirvvs": 21.6e3}
STATE_DOUBLES = {"osecir": 60, "osecirvvs": 162}


def edge_bytes_per_run_day(model, N, E):
    """SURVEY.md 8d, explicit traveler state: departure writes and return reads E*n*8 bytes, each phase reads two node
    records and read-modify-writes a third: 2*E*n*8 + 2*N*3*n*8 (10.8 MB for osecir at N=400, E=10000)."""
    n = STATE_DOUBLES[model]
    return 2.0 * E * n * 8 + 2.0 * N * 3 * n * 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--runs", type=int, default=10000, help="ensemble runs per GPU (weak scaling)")
    ap.add_argument("--days", type=float, default=100.0)
    ap.add_argument("--nodes", type=int, default=400)
    ap.add_argument("--cpu-sample-runs", type=int, default=0, help="runs of the CPU baseline sample (0 = one per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-batches", type=int, default=2, help="batches of the end-to-end leg (sampling overlaps GPU)")
    ap.add_argument("--model", default="osecir", choices=["osecir", "osecirvvs"],
                    help="osecir = BASELINE configs[3] (the default bench line); osecirvvs = configs[4]'s model on one GPU")
    ap.add_argument("--sim-kind", default="model", choices=["model", "base"],
                    help="osecirvvs node type: osecirvvs::Simulation (vaccination, ADL mobility hooks) or the plain "
                         "mio::Simulation the reference's graph benchmark instantiates")
    return ap.parse_args()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c in sm if c > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(args):
    if args.model == "osecirvvs":
        return {"workload": f"ensemble of the synthetic {args.nodes}-county ODE-SECIRVVS graph (BASELINE configs[4] "
                            f"model, one GPU's share): {args.runs} runs/GPU x {args.nodes} counties x {args.days:g} days",
                "runs_per_gpu": args.runs, "counties": args.nodes, "edges": None, "age_groups": 6, "days": args.days,
                "dt_graph": 0.5, "integrator": "cash_karp54 abs 1e-10 rel 1e-5",
                "node_simulation": "osecirvvs::Simulation (day chunks, vaccination)" if args.sim_kind == "model"
                else "mio::Simulation<osecirvvs::Model> (as cpp/benchmarks/graph_simulation.cpp)",
                "parallelism": f"runs sharded x{args.gpus}",
                "l2_policy": "inputs larger than L2 (state 4 x %.2f GB per GPU)" % (args.runs * args.nodes * 1440 / 1e9),
                "sampling": "osecirvvs::draw_sample(graph) semantics, threefry2x32 seeds {1..6}, uniform +-10% "
                            "distributions around the parameters of cpp/benchmarks/graph_simulation.cpp",
                "outputs": "per-node daily p05/p25/p50/p75/p95 + mean of age-summed compartments and of the "
                           "node-summed state"}
    return {"workload": f"ParameterStudy ensemble, synthetic {args.nodes}-county ODE-SECIR graph "
                        f"(BASELINE configs[3]): {args.runs} runs/GPU x {args.nodes} counties x {args.days:g} days",
            "runs_per_gpu": args.runs, "counties": args.nodes, "edges": None, "age_groups": 6, "days": args.days,
            "dt_graph": 0.5, "integrator": "cash_karp54 abs 1e-10 rel 1e-5", "parallelism": f"runs sharded x{args.gpus}",
            "l2_policy": "inputs larger than L2 (state 4 x %.2f GB per GPU)" % (args.runs * args.nodes * 480 / 1e9),
            "sampling": "osecir::draw_sample(graph) semantics, threefry2x32 seeds {1..6}, uniform ranges of set_covid_parameters",
            "outputs": "per-node daily p05/p25/p50/p75/p95 + mean of age-summed compartments and of the node-summed state"}


def cpu_baseline(problem, days, sample_runs, threads):
    """The reference's CPU path (oracle port) on the host cores: `sample_runs` runs partitioned over `threads`
    workers like distribute_runs (parameter_studies.h:354-367). Returns (node-days/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from concurrent.futures import ThreadPoolExecutor
    oracle_lib.lib()
    runs = [r % problem.num_runs for r in range(sample_runs)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:  # ctypes releases the GIL during the call
        list(ex.map(lambda r: oracle_lib.graph_simulate(problem, r, days)["steps"], runs))
    dt = time.perf_counter() - t0
    return sample_runs * problem.num_nodes * days / dt, dt


METRIC = {"osecir": "graph_ode_secir_node_days_per_s", "osecirvvs": "graph_ode_secirvvs_node_days_per_s"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference itself
    cannot be built here) on all host threads. Rank 0 only."""
    if rank != 0:
        return
    from memilio_b200 import synthetic
    cores = os.cpu_count() or 1
    sample = args.cpu_sample_runs or 4 * cores
    from memilio_b200 import study
    if args.model == "osecirvvs":
        from memilio_b200 import flat
        graph = synthetic.vvs_study_graph(N=args.nodes, days=int(args.days) + 1,
                                          sim_kind=flat.SIM_MODEL if args.sim_kind == "model" else flat.SIM_BASE)
    else:
        graph = synthetic.covid_study_graph(N=args.nodes)
    ps = study.ParameterStudy(graph, 0.0, args.days, 0.5, max(sample, 1))
    problem = ps.sample(0, max(sample, 1))
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(problem, min(args.days, 2.0), cores, cores)
    vals, secs = [], []
    for _ in range(args.steps):
        v, s = cpu_baseline(problem, args.days, sample, cores)
        vals.append(v)
        secs.append(s)
    value = float(np.sum(sample * args.nodes * args.days * len(vals)) / np.sum(secs))
    cfg = workload_config(args)
    cfg["edges"] = problem.num_edges
    sample_txt = f"{sample} runs x {args.nodes} counties x {args.days:g} days per step over {cores} threads"
    emit({
        "impl": "reference", "metric": METRIC[args.model], "value": value, "unit": "node-days/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "node-days/s", "cores": cores, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": "node-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


# the one JSON line goes to the real stdout; everything else that libraries print there (the NCCL version banner of
# the first communicator, ...) is sent to stderr
_REAL_STDOUT = None


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from concurrent.futures import ThreadPoolExecutor
    from memilio_b200 import capi, ensemble, study, synthetic
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION") == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout; this script prints ONE json line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    R, N, T = args.runs, args.nodes, args.days
    days = int(T) + 1
    # ParameterStudy over the synthetic county graph; weak scaling: R runs per GPU, rank r owns the contiguous block
    # distribute_runs(world*R, world, r) of ONE global study (seeds {1..6}), sampled bit-exactly like the reference
    from memilio_b200 import flat
    model = flat.MODEL_OSECIRVVS if args.model == "osecirvvs" else flat.MODEL_OSECIR
    run0, count = study.distribute_runs(world * R, world, rank)
    assert count == R
    t_s = time.perf_counter()
    if args.model == "osecirvvs":
        graph = synthetic.vvs_study_graph(N=N, days=int(T) + 1,
                                          sim_kind=flat.SIM_MODEL if args.sim_kind == "model" else flat.SIM_BASE)
    else:
        graph = synthetic.covid_study_graph(N=N)
    ps = study.ParameterStudy(graph, 0.0, T, 0.5, world * R)
    problem = ps.sample(run0, R)
    num_groups = graph.num_groups
    host_sampling_s = time.perf_counter() - t_s
    # pinned host staging
    pinned_names = [n_ for n_ in ("y0", "node_local", "psets", "ctables", "mobtables")
                    if getattr(problem, n_) is not None]
    for name in pinned_names:
        a = getattr(problem, name)
        tns = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        tns.numpy()[...] = a
        setattr(problem, name, tns.numpy())
        setattr(problem, "_pin_" + name, tns)
    h2d = sum(getattr(problem, n_).nbytes for n_ in pinned_names)

    sim = capi.GraphSimulation(problem, device=local_rank, max_runs=R, upload=False)
    rec = ensemble.EnsembleRecord(num_groups, N, days, capacity_runs=R, mode=ensemble.ENS_AGE_SUM,
                                  device=local_rank, model=model)
    rec_t = rec.as_torch()
    ncol = rec.info()["num_columns"]
    d2h = (len(ensemble.DEFAULT_PERCENTILES) + 1) * ncol * 8
    stream = torch.cuda.ExternalStream(sim.stream_ptr(), device=torch.device("cuda", local_rank))

    def fresh():
        """rewind the handle to t0 and upload the runs (H2D) -- untimed for `value`"""
        sim.reset()
        rec.attach(sim, 0)
        sim.set_runs(0, problem, 0, R)

    def reduce():
        return ensemble.reduce_distributed(rec_t, R, rank, world)

    def step_device(profile=False):
        fresh()
        sim.set_profiling(profile)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        sim.advance(T)  # synchronises the handle's stream before it returns
        e1.record(stream)
        reduce()        # synchronous: exchange (N>1), selection kernels, gather on rank 0
        e2.record(stream)
        e2.synchronize()
        return e0.elapsed_time(e2), e1.elapsed_time(e2)

    for _ in range(args.warmup):
        step_device()
    # ---- timed: device leg (inputs resident in HBM) -----------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    ms_dev, ms_red = [], []
    prof = None
    launches = 0
    attempts = 0
    for k in range(args.steps):
        a, b = step_device(profile=True)
        ms_dev.append(a)
        ms_red.append(b)
        p = sim.profile()
        prof = p if prof is None else {c: (prof[c][0] + p[c][0], prof[c][1] + p[c][1]) for c in p}
        launches += sum(v[1] for v in p.values()) + 1  # + the selection kernel
        cnt = sim.counters()
        attempts += int(cnt[:, :, 0].sum() + cnt[:, :, 1].sum())
    barrier()
    clock_info = clocks.stop()
    replays = sim.stats()["exact_replays"]
    sim.set_profiling(False)

    # ---- timed: end-to-end leg: what a user of ParameterStudy.run pays, from distributions to percentiles ------------
    # host sampling (C library) of batch b+1 overlaps the GPU work of batch b; H2D of every batch and D2H of the
    # reduced result are inside the timed region.
    nb = max(1, min(args.e2e_batches, R))
    bounds = [(i * R // nb, (i + 1) * R // nb) for i in range(nb)]
    bmax = max(b - a for a, b in bounds)
    sim2 = None

    def step_e2e():
        t0 = time.perf_counter()
        # the public call: ParameterStudy.run_ensemble = sampling (host threads), H2D from pinned buffers, advance,
        # on-device percentile reduction, D2H of the reduced result
        out = ps.run_ensemble(num_batches=nb, device=local_rank, rank=rank, world=world)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, out

    step_e2e()
    barrier()
    e2e_runs = [step_e2e() for _ in range(args.steps)]
    ms_e2e = [x[0] for x in e2e_runs]
    barrier()
    if rank == 0 and e2e_runs[-1][1][0] is not None:
        pct = e2e_runs[-1][1][0]
        assert np.isfinite(pct).all()

    t_dev = torch.tensor([float(np.sum(ms_dev)), float(np.sum(ms_e2e)), float(np.sum(ms_red)), host_sampling_s],
                         device="cuda", dtype=torch.float64)
    agg = torch.tensor([float(attempts), float(launches)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    tot_dev_ms, tot_e2e_ms, tot_red_ms, samp_s = [float(x) for x in t_dev.tolist()]
    units = world * R * N * T * args.steps
    value = units / (tot_dev_ms * 1e-3)
    e2e = units / (tot_e2e_ms * 1e-3)

    if rank == 0:
        peaks = {}
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        fp64_peak = capi.measure_fp64_peak(local_rank, 0.5)
        kernel_ms = {c: prof[c][0] / args.steps for c in prof}
        kernel_ms["select"] = tot_red_ms / args.steps
        dominant = max(kernel_ms, key=kernel_ms.get)
        integ_s = prof["integrate"][0] * 1e-3
        flop_attempt = FLOP_PER_CK54_ATTEMPT[args.model]
        edge_bytes = edge_bytes_per_run_day(args.model, N, problem.num_edges)
        fp64_achieved = attempts * flop_attempt / integ_s / 1e12 if integ_s > 0 else 0.0
        edge_s = (prof["depart"][0] + prof["return"][0] + prof["return_exact"][0]) * 1e-3
        edge_gbs = edge_bytes * R * T * args.steps / edge_s / 1e9 if edge_s > 0 else 0.0
        traffic = ncu_traffic(R) if args.model == "osecir" else {}
        roof_fp64 = {"kernel": "integrate_nodes_kernel", "bound": "fp64", "achieved": fp64_achieved, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak if fp64_peak else None,
                     "traffic": traffic.get("integrate"),
                     "peak_source": "measured in this run (DFMA chains on all SMs, mio_cuda_measure_fp64_peak); "
                                    "MEASURED_PEAKS.json has no FP64 figure",
                     "attempts_per_launch": attempts / max(prof["integrate"][1], 1),
                     "flop_per_attempt": flop_attempt}
        roof_edge = {"kernel": "depart_kernel+return_kernel(+return_exact)", "bound": "hbm", "achieved": edge_gbs,
                     "peak": hbm_peak, "unit": "GB/s", "frac": edge_gbs / hbm_peak,
                     "traffic": (traffic.get("depart", 0) + traffic.get("return", 0)) or None,
                     "peak_source": hbm_src, "algorithmic_bytes_per_run_day": edge_bytes,
                     "note": "algorithmic bytes are SURVEY 8d's explicit-traveler-state figure; the aggregated "
                             "formulation moves fewer DRAM bytes (see traffic), so achieved can exceed DRAM throughput"}
        line = {
            "metric": METRIC[args.model], "value": value, "unit": "node-days/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args), edges=int(problem.num_edges)),
            "e2e": {"value": e2e, "unit": "node-days/s",
                    "h2d_bytes_per_step": int(getattr(ps, "last_h2d_bytes", h2d)), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": tot_e2e_ms / args.steps, "batches": nb,
                    "includes": "ParameterStudy.run_ensemble: host sampling (C library threads, overlapped), "
                                "H2D of the runs' parameters and initial values from pinned memory, advance, "
                                "percentile reduction, D2H"},
            "gpu_launches": int(agg[1].item()),
            "clocks": clock_info,
            "roofline": roof_edge if dominant in ("depart", "return", "return_exact") else roof_fp64,
            "roofline_node_phase": roof_fp64, "roofline_edge_phase": roof_edge,
            "kernel_ms_per_step": kernel_ms, "dominant_kernel": dominant,
            "rk_attempts_per_node_day": float(agg[0].item()) / units,
            "exact_return_replays": int(replays),
            "host_sampling_s_per_step": samp_s,
        }
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = args.cpu_sample_runs or 4 * cores
            v, s = cpu_baseline(problem, T, sample, cores)
            line["cpu_baseline"] = {"value": v, "unit": "node-days/s", "cores": cores, "kind": "port",
                                    "sample": f"{sample} runs x {N} counties x {T:g} days over {cores} threads "
                                              f"({s:.1f} s)"}
        emit(line)
    sim.close()
    if sim2 is not None:
        sim2.close()
    rec.close()
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(R):
    """DRAM bytes per launch of each kernel from the committed ncu --set full capture (profiles/ncu_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum at the captured run count), scaled linearly to this run count."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        t = json.load(f)
    return {k: v["bytes_per_launch"] * R / v["runs"] for k, v in t["kernels"].items()}


if __name__ == "__main__":
    main()
