#!/bin/bash
# quick check on the GPU box: parity tests (edge + single node + vvs), then a 30-day bench with per-kernel-class times.
# usage: gpu_check.sh [variant-lib ...]   (no argument: the default library)
set -u
mkdir -p gpurun_out
run_one() {
  local tag=$1
  echo "== $tag"
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vvs.py tests/test_gpu_facade.py -x -q --timeout 90 2>&1 | tail -4
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --days 30 2> gpurun_out/bench_$tag.err | tee gpurun_out/bench_$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g ms %.1f'%(d['value'], d['ms_per_step']), {k:round(v,1) for k,v in d['kernel_ms_per_step'].items()}, 'e2e ms %.1f'%d['e2e']['ms_per_step'], 'fp64 frac %.3f'%d['roofline_node_phase']['frac'])"
  tail -2 gpurun_out/bench_$tag.err
}
if [ $# -eq 0 ]; then
  run_one default
else
  for lib in "$@"; do
    MIO_CUDA_LIB=$PWD/$lib run_one $(basename $(dirname $lib))
  done
fi
