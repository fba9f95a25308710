#!/bin/bash
# last GPU call of a round: the full GPU suite, the default bench line with the event breakdown, the ncu launch list
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --breakdown gpurun_out/${TAG}_breakdown_events.txt --watchdog 180 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/${TAG}_bench.err
cut -c1-400 gpurun_out/${TAG}_bench.json
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -s 98 -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/one_step.py 2 > gpurun_out/${TAG}_launches.log 2>&1
tail -2 gpurun_out/${TAG}_launches.log
