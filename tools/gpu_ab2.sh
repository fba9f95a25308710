#!/bin/bash
# A/B of an environment switch in ONE session (same box): bench only, alternating.  usage: gpu_ab2.sh TAG VAR "v1 v2"
mkdir -p gpurun_out
TAG=${1:-ab2}
VAR=${2:-E2T_REC_DBGSKIP}
for rep in 1 2; do
for V in ${3:-0 8}; do
  if [ "$V" = "unset" ]; then unset $VAR; else export $VAR=$V; fi
  timeout 600 python bench.py --steps 30 --warmup 5 --no-decode --no-cpu-baseline --breakdown gpurun_out/${TAG}_${V}_breakdown.txt > gpurun_out/${TAG}_${V}_bench.json 2> gpurun_out/${TAG}_${V}_bench.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_${V}_bench.json').read().strip().splitlines()[-1])
print('$VAR=$V', {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['us_per_recurrent_step'], d['roofline']['us_per_launch'])
PY
done
done
