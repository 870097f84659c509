#!/bin/bash
# A/B timing of launch-shape variants (tools/build_variants.py) on the GPU box: per-kernel-class ms from bench.py
set -u
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
fi
for v in warmup default "$@"; do
  if [ "$v" = default ] || [ "$v" = warmup ]; then unset MIO_CUDA_LIB; else export MIO_CUDA_LIB=$PWD/memilio_b200/csrc/variants/libmio_cuda_$v.so; fi
  echo "== variant $v"
  timeout 600 python bench.py --runs 10000 --days ${DAYS:-30} --steps 1 --warmup 1 --no-cpu-baseline 2>gpurun_out/var_$v.err | tee gpurun_out/var_$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms_per_step'])"
  tail -2 gpurun_out/var_$v.err
done
