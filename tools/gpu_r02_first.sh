#!/bin/bash
# round 2, first call: GPU tests, default bench line with the new parity keys, reference arm
set -u
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
tail -3 gpurun_out/bench_r02a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02a.json'))
for k in ('value','ms_per_step','kernel_ms_per_step','parity_check','e2e_check','cpu_baseline','roofline_edge_phase','exact_return_replays_per_step'):
    print(k, d.get(k))
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_r02a_ref.json 2> gpurun_out/bench_r02a_ref.err
cat gpurun_out/bench_r02a_ref.json | cut -c1-600
