#!/bin/bash
# 2-GPU check of a round's final state: data-parallel equality test, then the weak line with the flat and the bucketed all-reduce
# and the strong line, all under the bench watchdog
mkdir -p gpurun_out
TAG=${1:-n2c}
timeout 200 python -m pytest tests/test_gpu_zz_fit.py -m gpu -q -k "data_parallel" 2>&1 | tail -3
for MODE in "weak off" "weak on" "strong off"; do
  set -- $MODE
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 --no-decode --no-cpu-baseline --scaling $1 --overlap-allreduce $2 --watchdog 100 > gpurun_out/${TAG}_bench_$1_$2.json 2> gpurun_out/${TAG}_bench_$1_$2.err
  echo "rc=$?"
  tail -c 300 gpurun_out/${TAG}_bench_$1_$2.err | grep -v "OMP_NUM_THREADS\|\*\*\*\*" 
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_$1_$2.json').read().strip().splitlines()[-1])
    print('$MODE', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, 'e2e', d['e2e']['value'], d['config'].get('global_batch'))
except Exception as e: print('$MODE', 'no line', e)
PY
done
