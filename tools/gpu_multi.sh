#!/bin/bash
# multi-GPU bench lines: gpu_multi.sh N tag [bench args...]
set -u
N=$1; tag=$2; shift 2
mkdir -p gpurun_out
timeout ${TMO:-800} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N "$@" > gpurun_out/bench_${tag}_N$N.json 2> gpurun_out/bench_${tag}_N$N.err
tail -3 gpurun_out/bench_${tag}_N$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${tag}_N$N.json"))
    print("$tag N=$N value %.4g ms %.1f e2e %.4g"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()})
    print(d.get("multi_gpu_check")); print(d.get("e2e_check"))
except Exception as e:
    print("no json", e)
PY
