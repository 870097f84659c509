#!/bin/bash
# GPU box: parity tests, smoke, both bench arms, osecirvvs bench line. Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench"
timeout 1500 python bench.py --steps 2 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
echo "== bench vvs"
timeout 1500 python bench.py --model osecirvvs --runs 2000 --steps 1 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_vvs.err | tee gpurun_out/bench_vvs.json
tail -5 gpurun_out/bench_vvs.err
echo "== bench reference"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json
tail -3 gpurun_out/bench_ref.err
