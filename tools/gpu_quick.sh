#!/bin/bash
# quick A/B: GPU parity tests (edge + ensemble) and a 30-day bench with the per-kernel-class times
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --days 30 2> gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e']['ms_per_step'])"
tail -2 gpurun_out/bench.err
