#!/bin/bash
# quick check of the recurrent kernels: the parity tests that exercise them, a timeline, a short bench with the breakdown
mkdir -p gpurun_out
TAG=${1:-ab}
E2T_REC_TRAPINFO=1 timeout 600 python -m pytest tests -m gpu -x -q -k "full_width or config2_train_step_matches or deterministic" 2>&1 | tail -15
echo "== forced re-pulls"
E2T_REC_DBGSKIP=4 E2T_REC_TRAPINFO=1 timeout 600 python -m pytest tests -m gpu -x -q -k "full_width or config2_train_step_matches" 2>&1 | tail -3
E2T_REC_DEBUG=1 timeout 300 python tools/one_step.py 1 > gpurun_out/${TAG}_timeline.txt 2>&1
grep -A10 "rec fwd16\]" gpurun_out/${TAG}_timeline.txt | head -13 | cut -c1-200
grep -A8 "rec bptt3\]" gpurun_out/${TAG}_timeline.txt | head -12 | cut -c1-200
grep -A11 "dec fwd16\]" gpurun_out/${TAG}_timeline.txt | head -13 | cut -c1-200
grep -A11 "dec bwd16\]" gpurun_out/${TAG}_timeline.txt | head -13 | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --no-decode --no-cpu-baseline --breakdown gpurun_out/${TAG}_breakdown.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['us_per_recurrent_step'])
PY
head -12 gpurun_out/${TAG}_breakdown.txt
