#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-ab}
for SK in 8; do
  echo "===== E2T_REC_DBGSKIP=$SK"
  E2T_REC_DBGSKIP=$SK E2T_REC_DEBUG=1 timeout 300 python tools/one_step.py 1 > gpurun_out/${TAG}_skip${SK}_timeline.txt 2>&1
  grep -A12 "rec fwd16\]" gpurun_out/${TAG}_skip${SK}_timeline.txt | head -14 | cut -c1-230
done
