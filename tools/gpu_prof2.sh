#!/bin/bash
# ncu full capture of one kernel for several library variants; CSV pages exported on the box
set -u
mkdir -p gpurun_out
ARGS="${BENCH_ARGS:---runs 10000 --days 5 --steps 1 --warmup 1 --no-cpu-baseline}"
k=${KERNEL:-integrate_nodes}
for v in "$@"; do
  if [ "$v" = default ]; then unset MIO_CUDA_LIB; else export MIO_CUDA_LIB=$PWD/memilio_b200/csrc/variants/libmio_cuda_$v.so; fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-5} -c 1 -f -o /tmp/prof_${k}_$v \
      python bench.py $ARGS > gpurun_out/ncu_full_${k}_$v.log 2>&1
  tail -1 gpurun_out/ncu_full_${k}_$v.log | cut -c1-200
  ncu -i /tmp/prof_${k}_$v.ncu-rep --page raw --csv > gpurun_out/prof_${k}_${v}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_${k}_$v.ncu-rep --page source --csv > gpurun_out/prof_${k}_${v}_source.csv 2>/dev/null
  gzip -f gpurun_out/prof_${k}_${v}_source.csv
done
