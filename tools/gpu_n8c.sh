#!/bin/bash
# 8-GPU: overlapped all-reduce with a capped number of NCCL channels (fewer SMs taken from the critical path)
mkdir -p gpurun_out
TAG=${1:-n8c}
for CH in 4 8; do
  NCCL_MAX_NCHANNELS=$CH timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 --no-decode --no-cpu-baseline --overlap-allreduce on > gpurun_out/${TAG}_ov_ch${CH}.json 2> gpurun_out/${TAG}_ov_ch${CH}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_ov_ch${CH}.json').read().strip().splitlines()[-1])
print('overlap channels=$CH', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
done
NCCL_MAX_NCHANNELS=8 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 --no-decode --no-cpu-baseline --overlap-allreduce off > gpurun_out/${TAG}_flat_ch8.json 2> gpurun_out/${TAG}_flat_ch8.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_flat_ch8.json').read().strip().splitlines()[-1])
print('flat channels=8', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
