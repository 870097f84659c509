#!/usr/bin/env python
"""Turn the CSV pages exported by tools/gpu_final.sh (ncu --set full, raw + source pages; launch list) into the text
summaries committed under profiles/. Usage: summarize_profiles.py <gpurun_out dir> <profiles dir> <round tag>"""
import collections
import csv
import gzip
import json
import os
import sys

src, dst, tag = sys.argv[1], sys.argv[2], sys.argv[3]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(fn):
    rows = list(csv.reader(open(fn)))
    return dict(zip(rows[0], zip(rows[2], rows[1])))


def source(fn):
    rows = list(csv.reader(gzip.open(fn, "rt")))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    ins = [r for r in rows[2:] if len(r) >= len(hdr) and r[0] != "Address"]
    ops, samples = collections.Counter(), collections.Counter()
    stall = collections.Counter()
    cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in ins:
        t = r[ix["Source"]].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += int(r[ix["Instructions Executed"]] or 0)
        samples[op] += int(r[ix["# Samples"]] or 0)
        for c in cols:
            stall[c] += int(r[ix[c]] or 0)
    return rows[0][1] if len(rows[0]) > 1 else "", ops, samples, stall


traffic = {}
for k in ["integrate_nodes", "depart_smem", "return_tma", "return_exact", "select_percentiles"]:
    rfn = os.path.join(src, f"prof_{k}_raw.csv")
    sfn = os.path.join(src, f"prof_{k}_source.csv.gz")
    if not os.path.exists(rfn):
        continue
    d = raw(rfn)
    name, ops, samples, stall = source(sfn)
    out = [f"# ncu --set full --clock-control none --import-source on -k regex:{k}, one launch of",
           "# python bench.py --runs 10000 --days 5 --steps 1 --warmup 1 --no-cpu-baseline  (B200, " + tag + ")",
           "kernel: " + name[:150]]
    for key in KEYS:
        if key in d:
            out.append(f"  {key:74s} {d[key][0]} {d[key][1]}")
    tot = sum(ops.values())
    out.append(f"# SASS opcode histogram (warp instructions executed, total {tot}) and sampled stall reasons")
    for op, n in ops.most_common(18):
        out.append(f"  {op:10s} {n:14d} {100 * n / max(tot, 1):6.2f}%  samples {samples[op]}")
    ts = sum(stall.values())
    for c, n in stall.most_common(8):
        out.append(f"  {c:28s} {100 * n / max(ts, 1):6.2f}%")
    open(os.path.join(dst, f"{tag}_ncu_{k}.txt"), "w").write("\n".join(out) + "\n")
    traffic[k] = float(d["dram__bytes_read.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[
        d["dram__bytes_read.sum"][1]] + float(d["dram__bytes_write.sum"][0]) * {
            "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][1]]

# launch list: per-kernel totals and per-launch average DRAM traffic
rows = list(csv.reader(open(os.path.join(src, "launches.csv"))))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr):
        continue
    per.setdefault((r[ix["ID"]], r[ix["Kernel Name"]]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]])
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for (_, name), m in per.items():
    short = name.split("(")[0].replace("void ", "").replace("miocu::", "")
    a = agg[short]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tot_ns = sum(a[1] for a in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120",
       "# python bench.py --runs 10000 --days 5 --steps 1 --warmup 1 --no-cpu-baseline (first 120 launches; " + tag + ")",
       "# per-launch times are cold-cache and serialised: compare the kernels' SHARES with bench.py's kernel_ms_per_step",
       f"{'kernel':44s} launches   total ms   share   avg ms   avg DRAM MB/launch"]
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{name[:44]:44s} {a[0]:8d} {a[1] / 1e6:10.2f} {100 * a[1] / tot_ns:6.1f}% {a[1] / 1e6 / a[0]:8.3f} {a[2] / 1e6 / a[0]:12.1f}")
open(os.path.join(dst, f"{tag}_launches_summary.txt"), "w").write("\n".join(out) + "\n")
# DRAM traffic per launch for bench.py's roofline.traffic (launch-list averages: both variants of the node phase)
names = {"integrate": "integrate_nodes_kernel", "depart": "depart_smem_kernel", "return": "return_tma_kernel"}
kern = {}
for key, nm in names.items():
    for name, a in agg.items():
        if name.startswith(nm):
            kern[key] = {"runs": 10000, "bytes_per_launch": a[2] / a[0]}
json.dump({"source": "ncu launch list (dram__bytes_read.sum + dram__bytes_write.sum, average per launch), bench.py --runs "
                     "10000 --days 5 (" + tag + ", profiles/" + tag + "_launches_summary.txt)", "kernels": kern},
          open(os.path.join(dst, "ncu_traffic.json"), "w"), indent=2)
print("\n".join(out))
