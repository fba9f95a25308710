#!/bin/bash
# A/B of the forward recurrence variants: E2T_REC_DUAL=0/1
mkdir -p gpurun_out
TAG=${1:-ab}
for DUAL in 0 1; do
  export E2T_REC_DUAL=$DUAL
  echo "===== E2T_REC_DUAL=$DUAL"
  timeout 600 python -m pytest tests -m gpu -x -q -k "full_width or config2_train_step_matches or deterministic" 2>&1 | tail -2
  E2T_REC_DEBUG=1 timeout 300 python tools/one_step.py 1 > gpurun_out/${TAG}_dual${DUAL}_timeline.txt 2>&1
  grep -A10 "rec fwd16\]" gpurun_out/${TAG}_dual${DUAL}_timeline.txt | head -13 | cut -c1-160
  timeout 600 python bench.py --steps 20 --warmup 5 --no-decode --no-cpu-baseline --breakdown gpurun_out/${TAG}_dual${DUAL}_breakdown.txt > gpurun_out/${TAG}_dual${DUAL}_bench.json 2> gpurun_out/${TAG}_dual${DUAL}_bench.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_dual${DUAL}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['us_per_recurrent_step'])
PY
done
