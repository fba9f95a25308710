#!/bin/bash
# timing experiments of the forward recurrence (E2T_REC_DBGSKIP bits): timelines only
mkdir -p gpurun_out
TAG=${1:-sk}
for SK in ${2:-1 2 3}; do
  echo "===== E2T_REC_DBGSKIP=$SK"
  E2T_REC_DBGSKIP=$SK E2T_REC_DEBUG=1 timeout 300 python tools/one_step.py 1 > gpurun_out/${TAG}_skip${SK}_timeline.txt 2>&1
  grep -A8 "rec fwd16\]" gpurun_out/${TAG}_skip${SK}_timeline.txt | head -10 | cut -c1-230
done
