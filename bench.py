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


DEVICE_STATE_DOUBLES = {"osecir": 60, "osecirvvs": 180}  # W lanes x 10 slots (device_common.cuh)
LANES = {"osecir": 6, "osecirvvs": 18}


def edge_bytes_per_run_day_explicit(model, N, E):
    """SURVEY.md 8d, explicit traveler state (what the reference moves): departure writes and return reads E*n*8 bytes,
    each phase reads two node records and read-modify-writes a third: 2*E*n*8 + 2*N*3*n*8 (10.8 MB for osecir at
    N=400, E=10000). Reported as a notional figure only: the kernels do not move these bytes."""
    n = STATE_DOUBLES[model]
    return 2.0 * E * n * 8 + 2.0 * N * 3 * n * 8


def edge_bytes_per_run_day(model, N):
    """Algorithmic bytes of the aggregated formulation the kernels implement, per run and day (one
    departure + one return): departure reads u and y and writes y and the visitor vector (4 node records), return
    reads y, the visitor vector and u and writes y (4 node records), and reads the return coefficients (32 B per
    lane): 8*N*n*8 + N*W*32 (1.61 MB for osecir at N=400). The run-invariant graph (CSR, weights) is L2 resident and
    not counted."""
    n = DEVICE_STATE_DOUBLES[model]
    return 8.0 * N * n * 8 + N * LANES[model] * 32.0


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
    ap.add_argument("--day-block", type=int, default=0,
                    help="size the ensemble record for this many days and advance / reduce window by window "
                         "(0 = one record for all days; needed when runs x columns does not fit the device)")
    ap.add_argument("--model", default="osecir", choices=["osecir", "osecirvvs"],
                    help="osecir = BASELINE configs[3] (the default bench line); osecirvvs = configs[4]'s model on one GPU")
    ap.add_argument("--flow-sim", action="store_true",
                    help="the nodes are mio::FlowSimulation (SURVEY 8f-1): the cumulative flows are integrated (osecir)")
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
            "node_simulation": "mio::FlowSimulation (cumulative flows integrated, error control on the flows)"
            if args.flow_sim else "osecir::Simulation",
            "l2_policy": "inputs larger than L2 (state 4 x %.2f GB per GPU)" % (args.runs * args.nodes * 480 / 1e9),
            "sampling": "osecir::draw_sample(graph) semantics, threefry2x32 seeds {1..6}, uniform ranges of set_covid_parameters",
            "outputs": "per-node daily p05/p25/p50/p75/p95 + mean of age-summed compartments and of the node-summed state"}


def cpu_baseline(args, sample_runs, workers, study_runs):
    """The reference's CPU path (oracle port, built -O3) on the host cores: the first `sample_runs` runs of the study,
    partitioned over `workers` PROCESSES like distribute_runs over MPI ranks (parameter_studies.h:163-213,354-367).
    Every process times its own literal run_impl loop -- draw_sample(graph) + GraphSimulation::advance +
    interpolate_simulation_result per run (oracle/cpu_arm_worker.py) -- after a common start; the arm's time is the
    slowest process. Returns dict(value node-days/s, seconds, state[sample][N][n], counters[sample][N][4])."""
    import tempfile
    tmp = tempfile.mkdtemp(prefix="mio_cpu_arm_")
    start = os.path.join(tmp, "start")
    worker = os.path.join(ROOT, "oracle", "cpu_arm_worker.py")
    procs, outs = [], []
    for r in range(workers):
        out = os.path.join(tmp, f"w{r}.npz")
        outs.append(out)
        procs.append(subprocess.Popen(
            [sys.executable, worker, "--model", args.model, "--sim-kind", args.sim_kind, "--flow-sim", str(int(args.flow_sim)),
             "--nodes", str(args.nodes),
             "--days", str(args.days), "--sample-runs", str(sample_runs), "--study-runs", str(study_runs),
             "--workers", str(workers), "--rank", str(r), "--out", out, "--start-file", start],
            stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="1")))
    # common start once every process has built its study (setup is not part of the reference's run() either)
    deadline = time.time() + 600
    while not all(os.path.exists(o + ".ready.npz") for o in outs):
        if any(p.poll() not in (None, 0) for p in procs) or time.time() > deadline:
            for p in procs:
                p.kill()
            raise RuntimeError("cpu_baseline: a worker failed during setup")
        time.sleep(0.01)
    open(start, "w").close()
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("cpu_baseline: a worker failed")
    res = [np.load(o) for o in outs]
    secs = max(float(r["seconds"]) for r in res)
    state = np.concatenate([r["state"] for r in res], axis=0)
    counters = np.concatenate([r["counters"] for r in res], axis=0)
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return dict(value=sample_runs * args.nodes * args.days / secs, seconds=secs, state=state, counters=counters)


def parity_against_oracle(gpu_state, gpu_counters, cpu):
    """The runs of the CPU sample are the first runs of the GPU ensemble (same study, same seeds): final states within
    1e-8 relative (floor 1e-9 absolute: compartments that hold less than a billionth of a person), accepted / rejected
    / forced step counts per node equal (north star). Mismatches are reported, not hidden."""
    k = cpu["state"].shape[0]
    gs, gc = gpu_state[:k], gpu_counters[:k]
    denom = np.maximum(np.abs(cpu["state"]), 1e-1)  # 1e-9 absolute floor at 1e-8 relative
    rel = np.abs(gs - cpu["state"]) / denom
    mism = (gc[:, :, :3] != cpu["counters"][:, :, :3]).any(axis=2)
    fix_gpu, fix_cpu = int(gc[:, :, 3].sum()), int(cpu["counters"][:, :, 3].sum())
    return {"runs": int(k), "nodes_compared": int(mism.size), "max_rel_err": float(rel.max()) if rel.size else 0.0,
            "tolerance": 1e-8, "within_tolerance": bool((rel <= 1e-8).all()),
            "step_count_mismatches": int(mism.sum()),
            "attempts_compared": int(cpu["counters"][:, :, 0].sum() + cpu["counters"][:, :, 1].sum()),
            "underflow_fixups_gpu": fix_gpu, "underflow_fixups_oracle": fix_cpu}


METRIC = {"osecir": "graph_ode_secir_node_days_per_s", "osecirvvs": "graph_ode_secirvvs_node_days_per_s"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference itself
    cannot be built here) on all host cores, one process per core. Rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample_runs or 4 * cores
    study_runs = world * args.runs
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(args, cores, cores, study_runs)
    vals, secs = [], []
    for _ in range(args.steps):
        r = cpu_baseline(args, sample, cores, study_runs)
        vals.append(r["value"])
        secs.append(r["seconds"])
    value = float(sample * args.nodes * args.days * len(vals) / np.sum(secs))
    cfg = workload_config(args)
    cfg["runs_timed_per_step"] = sample
    cfg["extrapolated_from_runs"] = sample
    cfg["note"] = ("the CPU arm times a bounded sample (the first %d runs of the same study) and reports node-days/s; "
                   "throughput is linear in runs, so the figure compares with the GPU arm's at runs_per_gpu" % sample)
    sample_txt = (f"{sample} runs x {args.nodes} counties x {args.days:g} days per step over {cores} processes "
                  f"(oracle -O3; draw_sample + advance + interpolate per run; slowest process)")
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
    ps.flow_simulation = int(args.flow_sim)
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
    blk = args.day_block if args.day_block > 0 else days
    rec = ensemble.EnsembleRecord(num_groups, N, min(blk, days), capacity_runs=R, mode=ensemble.ENS_AGE_SUM,
                                  device=local_rank, model=model)
    rec_t = rec.as_torch()
    ncol = N * days * rec.El + days * rec.n  # columns of the whole period
    d2h = (len(ensemble.DEFAULT_PERCENTILES) + 1) * ncol * 8
    stream = torch.cuda.ExternalStream(sim.stream_ptr(), device=torch.device("cuda", local_rank))

    def fresh():
        """rewind the handle to t0 and upload the runs (H2D) -- untimed for `value`"""
        sim.reset()
        rec.set_window(0)
        rec.attach(sim, 0)
        sim.set_runs(0, problem, 0, R)

    def reduce():
        return ensemble.reduce_distributed(rec_t, R, rank, world, rec=rec)

    def step_device(profile=False):
        fresh()
        sim.set_profiling(profile)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        if blk >= days:
            sim.advance(T)  # synchronises the handle's stream before it returns
            e1.record(stream)
            red = reduce()  # synchronous: exchange (N>1), selection kernels, gather on rank 0
            e2.record(stream)
            e2.synchronize()
            return e0.elapsed_time(e2), e1.elapsed_time(e2), red
        # windowed record: advance and reduce window by window (all calls are synchronous); the reduction time is the
        # sum over the windows
        marks = []

        def timer(kind):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            marks.append((kind, ev))
        full = ensemble.windowed_advance_reduce(sim, rec, rec_t, R, 0.0, T, days, blk, rank, world,
                                                ensemble.DEFAULT_PERCENTILES, True, timer)
        e2.record(stream)
        e2.synchronize()
        ms_reduce = sum(a[1].elapsed_time(b[1]) for a, b in zip(marks[:-1], marks[1:]) if a[0] == "reduce")
        red = (full[:-1], full[-1]) if full is not None else (None, None)
        return e0.elapsed_time(e2), ms_reduce, red

    for _ in range(args.warmup):
        step_device()
    replays0 = sim.stats()["exact_replays"]
    # ---- timed: device leg (inputs resident in HBM) -----------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    ms_dev, ms_red = [], []
    prof = None
    launches = 0
    attempts = 0
    red_dev = None
    for k in range(args.steps):
        a, b, red_dev = step_device(profile=True)
        ms_dev.append(a)
        ms_red.append(b)
        p = sim.profile()
        prof = p if prof is None else {c: (prof[c][0] + p[c][0], prof[c][1] + p[c][1]) for c in p}
        launches += sum(v[1] for v in p.values()) + 1  # + the selection kernel
        cnt = sim.counters()
        attempts += int(cnt[:, :, 0].sum() + cnt[:, :, 1].sum())
    barrier()
    clock_info = clocks.stop()
    replays = (sim.stats()["exact_replays"] - replays0) / max(args.steps, 1)
    sim.set_profiling(False)
    # the first runs of rank 0's block are the runs the CPU arm simulates: keep their GPU results for the parity check
    cores = os.cpu_count() or 1
    cpu_sample = min(args.cpu_sample_runs or 4 * cores, R)
    gpu_state = gpu_cnt = None
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline  # the CPU leg runs at N=1 only
    if want_cpu:
        gpu_state, gpu_cnt = sim.state(0, cpu_sample), sim.counters(0, cpu_sample)

    # ---- N > 1: the sharded reduction (exchange + multi-segment select + gather) against directly gathered columns
    multi_gpu_check = None
    if world > 1:
        # the record holds the last window (all days when there is one window): columns of that window are compared
        nblk = min(blk, days)
        d0_last = ((days - 1) // nblk) * nblk
        len_last = days - d0_last
        ncol_rec = int(rec_t.shape[0])
        gen = torch.Generator().manual_seed(1234)
        cand = torch.randperm(ncol_rec, generator=gen)[:4000].sort().values.numpy()
        El, nn = rec.El, rec.n
        node_cols = N * nblk * El
        full_idx, keep = [], []
        for c in cand:
            if c < node_cols:
                k_, rem = divmod(int(c), nblk * El)
                dd, e_ = divmod(rem, El)
                if dd < len_last:
                    full_idx.append((k_ * days + d0_last + dd) * El + e_)
                    keep.append(int(c))
            else:
                dd, j_ = divmod(int(c) - node_cols, nn)
                if dd < len_last:
                    full_idx.append(N * days * El + (d0_last + dd) * nn + j_)
                    keep.append(int(c))
        keep, full_idx = keep[:1000], np.array(full_idx[:1000])
        cols = torch.tensor(keep, device=rec_t.device)
        ncheck = len(keep)
        mine = rec_t[cols, :R].contiguous()
        parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        if rank == 0:
            allruns = torch.cat(parts, dim=1)  # [ncheck][world*R], rank-major = run order
            srt = torch.sort(allruns, dim=1).values
            total = allruns.shape[1]
            pct_dev = red_dev[0]
            bad = 0
            for i, q in enumerate(ensemble.DEFAULT_PERCENTILES):
                want = srt[:, int(total * q)].cpu().numpy()
                bad += int((pct_dev[i][full_idx] != want).sum())
            mean_ref = allruns.mean(dim=1).cpu().numpy()
            mean_err = float(np.max(np.abs(red_dev[1][full_idx] - mean_ref) / np.maximum(np.abs(mean_ref), 1e-300)))
            multi_gpu_check = {"columns": int(ncheck), "runs": int(total), "percentiles": len(ensemble.DEFAULT_PERCENTILES),
                               "percentile_mismatches": bad, "mean_max_rel_err": mean_err,
                               "days_checked": [int(d0_last), int(days - 1)],
                               "method": "columns gathered from all ranks, torch.sort, index size_t(runs*p)"}
            assert bad == 0, f"sharded percentiles differ from directly sorted columns in {bad} entries"

    # ---- timed: end-to-end leg: what a user of ParameterStudy.run pays, from distributions to percentiles ------------
    # host sampling (C library) of batch b+1 overlaps the GPU work of batch b; H2D of every batch and D2H of the
    # reduced result are inside the timed region.
    nb = max(1, min(args.e2e_batches, R))
    if blk < days:
        nb = 1  # a windowed record keeps all runs of the rank resident
        # the device leg's handle and record make room for the public call's own (same sizes)
        sim.close()
        rec.close()
        del rec_t
        torch.cuda.empty_cache()
    sim2 = None

    def step_e2e():
        t0 = time.perf_counter()
        # the public call: ParameterStudy.run_ensemble = sampling (host threads), H2D from pinned buffers, advance,
        # on-device percentile reduction, D2H of the reduced result
        out = ps.run_ensemble(num_batches=nb, device=local_rank, rank=rank, world=world,
                              day_block=args.day_block if args.day_block > 0 else None)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, out

    # the first (warm-up) call samples the study from RNG counter 0: the same runs as the device leg, so its result
    # has to reproduce the device leg's bit for bit (sampling, batching, two handles, upload -- nothing may change a
    # value); later calls continue the random sequence (parameter_studies.h:339)
    _, first = step_e2e()
    e2e_check = None
    if rank == 0:
        same_p = bool(np.array_equal(first[0], red_dev[0]))
        same_m = bool(np.array_equal(first[1], red_dev[1]))
        e2e_check = {"percentiles_bit_equal_to_device_leg": same_p, "mean_bit_equal_to_device_leg": same_m,
                     "columns": int(ncol)}
        assert same_p, "run_ensemble percentiles differ from the device-resident leg on the same runs"
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
        edge_bytes = edge_bytes_per_run_day(args.model, N)
        edge_bytes_explicit = edge_bytes_per_run_day_explicit(args.model, N, problem.num_edges)
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
        edge_traffic = (traffic.get("depart", 0) + traffic.get("return", 0)) or None
        edge_launches = max(prof["depart"][1] + prof["return"][1], 1)
        alg_per_launch = edge_bytes * R * T * args.steps / edge_launches  # one departure or return of R runs
        roof_edge = {"kernel": "depart + return (+ return_exact)", "bound": "hbm", "achieved": edge_gbs,
                     "peak": hbm_peak, "unit": "GB/s", "frac": edge_gbs / hbm_peak,
                     "traffic": edge_traffic / 2 if edge_traffic else None,
                     "traffic_over_algorithmic": (edge_traffic / 2) / alg_per_launch if edge_traffic else None,
                     "peak_source": hbm_src, "algorithmic_bytes_per_run_day": edge_bytes,
                     "algorithmic_bytes_per_launch": alg_per_launch,
                     "explicit_traveler_bytes_per_run_day": edge_bytes_explicit,
                     "note": "algorithmic bytes are those of the aggregated formulation the kernels implement "
                             "(4 node records per phase + return coefficients); SURVEY 8d's explicit-traveler-state "
                             "figure is listed for reference only, those bytes are never moved. time includes "
                             "return_exact. traffic = measured DRAM bytes per launch (mean of depart and return)"}
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
            "exact_return_replays_per_step": float(replays),
            "host_sampling_s_per_step": samp_s,
            "e2e_check": e2e_check,
        }
        if multi_gpu_check is not None:
            line["multi_gpu_check"] = multi_gpu_check
        if want_cpu:
            cpu = cpu_baseline(args, cpu_sample, cores, world * R)
            line["cpu_baseline"] = {"value": cpu["value"], "unit": "node-days/s", "cores": cores, "kind": "port",
                                    "sample": f"{cpu_sample} runs x {N} counties x {T:g} days over {cores} processes "
                                              f"(oracle -O3; draw_sample + advance + interpolate per run; "
                                              f"{cpu['seconds']:.1f} s)"}
            # the same runs on both sides: the CPU arm's results are the parity vectors
            line["parity_check"] = parity_against_oracle(gpu_state, gpu_cnt, cpu)
        emit(line)
    ps.close()
    sim.close()
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
