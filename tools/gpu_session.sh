#!/bin/bash
# one gpurun call: GPU tests (writes gpurun_out/parity_r2.json), recurrent timelines, the bench line with the event breakdown
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
E2T_REC_DEBUG=2 timeout 300 python tools/one_step.py 1 > gpurun_out/${TAG}_rec_timeline.txt 2>&1
head -60 gpurun_out/${TAG}_rec_timeline.txt | cut -c1-150
timeout 600 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/${TAG}_breakdown_events.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['us_per_recurrent_step'], d['roofline']['frac'], d['decode']['greedy_ms_per_utt_batch1_host'])
PY
head -12 gpurun_out/${TAG}_breakdown_events.txt
