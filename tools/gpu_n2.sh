#!/bin/bash
# 2-GPU session: data-parallel equality test, trainer schedule test, weak / strong bench lines
mkdir -p gpurun_out
TAG=${1:-n2}
timeout 900 python -m pytest tests/test_gpu_zz_fit.py tests/test_gpu_parity.py -m gpu -q -k "data_parallel or transfer or schedule or subject" 2>&1 | tail -8
for MODE in weak strong; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-decode --no-cpu-baseline --scaling $MODE > gpurun_out/${TAG}_bench_${MODE}.json 2> gpurun_out/${TAG}_bench_${MODE}.err
  tail -c 400 gpurun_out/${TAG}_bench_${MODE}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${MODE}.json').read().strip().splitlines()[-1])
print('$MODE', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'], d['config'].get('global_batch'))
PY
done
