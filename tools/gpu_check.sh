#!/bin/bash
# Run on the GPU box via gpurun: parity tests, bench, ncu launch list + full captures. Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pytest -m gpu" 
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench"
timeout 1500 python bench.py --steps 3 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --runs 2000 --days 5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
tail -2 gpurun_out/ncu_launch_bench.log
echo "== ncu full"
for k in integrate_nodes depart_kernel return_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -f -o gpurun_out/prof_$k \
    python bench.py --runs 2000 --days 5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1
tail -1 gpurun_out/ncu_full_$k.log
done
fi
ls -la gpurun_out
