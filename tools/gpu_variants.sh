#!/bin/bash
# 30-day bench per variant library (kernel-class times only). usage: gpu_variants.sh name ...
set -u
mkdir -p gpurun_out
for name in "$@"; do
  lib=$PWD/memilio_b200/csrc/variants/$name/libmio_cuda.so
  [ "$name" = default ] && lib=$PWD/memilio_b200/csrc/libmio_cuda.so
  MIO_CUDA_LIB=$lib timeout 100 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --days ${DAYS:-30} 2> gpurun_out/var_$name.err | tee gpurun_out/var_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', 'ms %.1f'%d['ms_per_step'], {k:round(v,1) for k,v in d['kernel_ms_per_step'].items()}, 'fp64 frac %.3f'%d['roofline_node_phase']['frac'])" || tail -3 gpurun_out/var_$name.err
done
