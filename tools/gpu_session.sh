#!/bin/bash
# one gpurun call: GPU tests (writes gpurun_out/parity_r2.json), GEMM / cuBLAS peaks, the bench line with the event breakdown
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_gemm.py gpurun_out/r2_gemm_peaks.json > gpurun_out/r2_gemm_peaks.log 2>&1
tail -3 gpurun_out/r2_gemm_peaks.log
timeout 600 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/r2a_breakdown_events.txt > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 1500 gpurun_out/r2a_bench.json
