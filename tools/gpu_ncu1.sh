#!/bin/bash
# ncu --set full of one launch of a kernel (regex $1, skip $2 launches) of a short bench run; csv pages under gpurun_out/
set -u
mkdir -p gpurun_out
k=$1; skip=${2:-40}; tag=${3:-$k}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o /tmp/prof_$tag \
    python bench.py --runs 10000 --days ${DAYS:-12} --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log | cut -c1-150
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_source.csv 2>/dev/null
gzip -f gpurun_out/prof_${tag}_source.csv
ls -la gpurun_out/prof_${tag}*
