#!/bin/bash
# evidence session of a round (1 GPU): launch list, --set full captures of the four persistent recurrent kernels, timelines,
# the default bench line with the event breakdown.  Outputs under gpurun_out/${TAG}_*; summaries are copied to profiles/.
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 98 -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/one_step.py 2 > gpurun_out/${TAG}_launches.log 2>&1
tail -2 gpurun_out/${TAG}_launches.log
for K in k_lstm_fwd16 k_lstm_bptt3 k_dec_fwd16 k_dec_bwd16; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/${TAG}_prof_$K -f python tools/one_step.py 1 > gpurun_out/${TAG}_prof_$K.log 2>&1
  ncu -i gpurun_out/${TAG}_prof_$K.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/${TAG}_ncu_full_$K.txt 2>&1
  head -30 gpurun_out/${TAG}_ncu_full_$K.txt
done
E2T_REC_DEBUG=1 timeout 300 python tools/one_step.py 1 > gpurun_out/${TAG}_rec_timeline.txt 2>&1
timeout 900 python bench.py --breakdown gpurun_out/${TAG}_breakdown_events.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 400 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | cut -c1-1500
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_reference.json | cut -c1-600
