#!/bin/bash
# Round-end evidence run on the GPU box: parity tests, smoke, both bench arms, vvs line, ncu launch list and full captures.
set -u
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench"
timeout 1500 python bench.py --steps 3 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200
tail -3 gpurun_out/bench.err
echo "== bench vvs"
timeout 1500 python bench.py --model osecirvvs --runs 2000 --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_vvs.err | tee gpurun_out/bench_vvs.json | cut -c1-200
echo "== bench reference"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json | cut -c1-200
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --runs 10000 --days 5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
tail -1 gpurun_out/ncu_launch_bench.log | cut -c1-120
for k in integrate_nodes depart_smem return_tma return_exact select_percentiles; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-3} -c 1 -f -o /tmp/prof_$k \
    python bench.py --runs 10000 --days 5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1
tail -1 gpurun_out/ncu_full_$k.log | cut -c1-120
ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_${k}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$k.ncu-rep --page source --csv > gpurun_out/prof_${k}_source.csv 2>/dev/null
gzip -f gpurun_out/prof_${k}_source.csv
done
fi
ls -la gpurun_out | head -40
