#!/bin/bash
# GPU box: ncu launch list + full captures (with source) of the main kernels on a short bench. The reports are
# exported to CSV pages on the box (raw + source) and removed: gpurun_out/ is limited to 64 MiB.
set -u
mkdir -p gpurun_out
ARGS="${BENCH_ARGS:---runs 2000 --days 5 --steps 1 --warmup 1 --no-cpu-baseline}"
if [ "${SKIP_LIST:-0}" != "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/ncu_launch_bench.log 2>&1
tail -2 gpurun_out/ncu_launch_bench.log | cut -c1-300
fi
for k in ${KERNELS:-integrate_nodes depart_smem return_kernel}; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-5} -c ${COUNT:-1} -f -o /tmp/prof_$k \
    python bench.py $ARGS > gpurun_out/ncu_full_$k.log 2>&1
tail -1 gpurun_out/ncu_full_$k.log | cut -c1-200
ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_${k}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$k.ncu-rep --page source --csv > gpurun_out/prof_${k}_source.csv 2>/dev/null
gzip -f gpurun_out/prof_${k}_source.csv
done
ls -la gpurun_out
