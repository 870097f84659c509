#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== bench (short)"
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench_short.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e'])"
tail -3 gpurun_out/bench.err
