#!/usr/bin/env python
"""Time the tcgen05 GEMM at the shapes of one training step (diagnostic). usage: python tools/bench_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecog2txt_b200 import Engine, EngineConfig
eng = Engine(EngineConfig(subnet_ids=(7,), subnet_C=(8,), subnet_W=(4,), E=8, H=(16,), D=8, Hd=32, V=16, max_B=2, max_T=8, max_L=4))
SH = [(8704, 1600, 800, 0, 0.0), (8704, 800, 1600, 0, 0.0), (8704, 800, 1600, 0, 1.0), (8704, 1600, 100, 0, 0.0),
      (8704, 100, 1600, 0, 1.0), (256, 3200, 800, 0, 1.0), (256, 800, 3200, 0, 0.0), (2816, 1806, 800, 0, 0.0),
      (800, 1600, 8704, 1, 0.0), (400, 1600, 8448, 1, 0.0), (100, 1600, 8704, 1, 0.0), (1806, 800, 2816, 1, 0.0),
      (8192, 8192, 8192, 0, 0.0)]
for M, N, K, tn, beta in SH:
    ms = eng.bench_gemm(M, N, K, bool(tn), beta, 10)
    print(f"{'TN' if tn else 'NT'} [{M:5d},{N:5d},{K:5d}] beta={beta:.0f}: {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
