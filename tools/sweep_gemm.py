#!/usr/bin/env python
"""Sweep tile width / split-K of the tcgen05 GEMM for the step's awkward shapes (diagnostic; E2T_GEMM_BN / E2T_GEMM_KSPLIT)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecog2txt_b200 import Engine, EngineConfig
eng = Engine(EngineConfig(subnet_ids=(7,), subnet_C=(8,), subnet_W=(4,), E=8, H=(16,), D=8, Hd=32, V=16, max_B=2, max_T=8, max_L=4))
SHAPES = [(256, 3200, 800, 0), (256, 800, 3200, 0), (400, 1600, 8448, 1), (800, 1600, 8704, 1), (800, 3200, 2816, 1),
          (1806, 800, 2816, 1), (2816, 1806, 800, 0), (2816, 800, 1806, 0), (8704, 1600, 800, 0)]
for M, N, K, tn in SHAPES:
    res = []
    for bn in (0, 64, 128, 160, 192, 256):
        if tn and bn % 32: continue
        for ks in (0, 1, 2, 3, 4, 6, 8):
            os.environ["E2T_GEMM_BN"] = str(bn); os.environ["E2T_GEMM_KSPLIT"] = str(ks)
            try:
                ms = eng.bench_gemm(M, N, K, bool(tn), 0.0, 10)
            except Exception as e:
                print("fail", M, N, K, bn, ks, e); continue
            res.append((ms * 1e3, bn, ks))
    res.sort()
    auto = [r for r in res if r[1] == 0 and r[2] == 0][0][0]
    print(f"{'TN' if tn else 'NT'} [{M},{N},{K}] auto {auto:.1f} us; best: " + ", ".join(f"{us:.1f}us(BN={bn or 'a'},ks={ks or 'a'})" for us, bn, ks in res[:5]), flush=True)
