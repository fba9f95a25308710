#!/bin/bash
# 2-GPU check of the side-stream backward under both all-reduce modes + rank equality
mkdir -p gpurun_out
TAG=${1:-n2b}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dp_equality.py --steps 4 --out gpurun_out/${TAG}_dp_equality_2gpu.json 2>&1 | tail -2
for OV in off on; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-decode --no-cpu-baseline --overlap-allreduce $OV > gpurun_out/${TAG}_bench_2gpu_ov${OV}.json 2> gpurun_out/${TAG}_bench_2gpu_ov${OV}.err
  tail -c 300 gpurun_out/${TAG}_bench_2gpu_ov${OV}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_2gpu_ov${OV}.json').read().strip().splitlines()[-1])
print('overlap=$OV', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'])
PY
done
timeout 900 python -m pytest tests/test_gpu_zz_fit.py -m gpu -q 2>&1 | tail -3
