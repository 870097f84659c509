#!/bin/bash
# ncu --set full of selected kernels (regex list in $@) on a short bench run; reports under gpurun_out/
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --days 30 2> gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
for k in "$@"; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_$k \
    python bench.py --runs 2000 --days 5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1
tail -1 gpurun_out/ncu_full_$k.log
done
