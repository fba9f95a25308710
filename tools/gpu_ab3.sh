#!/bin/bash
# one gpurun call: the full GPU suite, then the bench line under several environment settings (same box, alternating).
# PYTEST_K="expr" restricts the suite.
# usage: gpu_ab3.sh TAG "ENV1=a ENV2=b" "ENV3=c" ...   (each argument is one setting; "-" = defaults)
mkdir -p gpurun_out
TAG=${1:-ab3}; shift
timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for rep in 1 2; do
i=0
for SET in "$@"; do
  i=$((i+1))
  if [ "$SET" = "-" ]; then ENVS=""; else ENVS="$SET"; fi
  env $ENVS timeout 600 python bench.py --steps 30 --warmup 5 --no-decode --no-cpu-baseline --breakdown gpurun_out/${TAG}_${i}_breakdown.txt > gpurun_out/${TAG}_${i}_bench.json 2> gpurun_out/${TAG}_${i}_bench.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_${i}_bench.json').read().strip().splitlines()[-1])
print('[$SET]', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['roofline']['us_per_recurrent_step'])
PY
done
done
