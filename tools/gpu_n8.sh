#!/bin/bash
# N-GPU session (gpurun --gpus N): config-4 schedule, weak / strong bench lines, data-parallel equality
mkdir -p gpurun_out
TAG=${1:-n8}
N=${2:-8}
[ -n "$SKIP_C4" ] || timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/config4_run.py --out gpurun_out/${TAG}_config4_${N}gpu.json 2>&1 | tail -3
for MODE in weak strong; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-decode --no-cpu-baseline --scaling $MODE > gpurun_out/${TAG}_bench_${N}gpu_${MODE}.json 2> gpurun_out/${TAG}_bench_${N}gpu_${MODE}.err
  tail -c 300 gpurun_out/${TAG}_bench_${N}gpu_${MODE}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${N}gpu_${MODE}.json').read().strip().splitlines()[-1])
print('$MODE', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'], d['config'].get('global_batch'))
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 --no-decode --no-cpu-baseline --overlap-allreduce off > gpurun_out/${TAG}_bench_${N}gpu_flat.json 2> gpurun_out/${TAG}_bench_${N}gpu_flat.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${N}gpu_flat.json').read().strip().splitlines()[-1])
print("flat", {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'])
PY
[ -n "$SKIP_C4" ] || timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/dp_equality.py --steps 4 --out gpurun_out/${TAG}_dp_equality_4gpu.json 2>&1 | tail -2
