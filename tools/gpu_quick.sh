#!/bin/bash
# quick A/B: recurrent timelines + event breakdown only
mkdir -p gpurun_out
TAG=${1:-q}
E2T_REC_DEBUG=1 timeout 300 python tools/one_step.py 1 > gpurun_out/${TAG}_rec_timeline.txt 2>&1
grep -A8 "rec fwd16\]" gpurun_out/${TAG}_rec_timeline.txt | head -12 | cut -c1-150
grep -A8 "rec bptt2\]" gpurun_out/${TAG}_rec_timeline.txt | head -12 | cut -c1-150
timeout 600 python bench.py --steps 20 --warmup 5 --no-decode --no-cpu-baseline --breakdown gpurun_out/${TAG}_breakdown_events.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 300 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['us_per_recurrent_step'], d['roofline']['frac'])
PY
head -4 gpurun_out/${TAG}_breakdown_events.txt
