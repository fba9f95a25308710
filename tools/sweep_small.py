#!/usr/bin/env python
"""Tile width / split-K sweep of the per-step decoder GEMM shapes (M = batch): E2T_GEMM_BN / E2T_GEMM_KSPLIT overrides."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecog2txt_b200 import Engine, EngineConfig
eng = Engine(EngineConfig(subnet_ids=(7,), subnet_C=(8,), subnet_W=(4,), E=8, H=(16,), D=8, Hd=32, V=16, max_B=2, max_T=8, max_L=4))
for M, N, K, tn in [(256, 3200, 952, 0), (256, 800, 3200, 0), (256, 1806, 800, 0)]:
    res = []
    for two in (os.environ.get("E2T_GEMM_NO_2SM", ""),):      # read once per process by the library: run the tool twice
        for bn in (0, 32, 64, 96, 128, 160, 192, 256):
            for ks in (0, 1, 2, 3, 4, 6, 8):
                if (bn == 0) != (ks == 0):
                    continue
                os.environ["E2T_GEMM_BN"] = str(bn); os.environ["E2T_GEMM_KSPLIT"] = str(ks)
                try:
                    ms = eng.bench_gemm(M, N, K, bool(tn), 0.0, 20)
                except Exception as e:
                    continue
                res.append((ms * 1e3, bn, ks, two))
    res.sort()
    auto = [r for r in res if r[1] == 0][0][0]
    print(f"[{M},{N},{K}] auto {auto:.1f} us; best: " + ", ".join(f"{us:.1f}us(BN={bn},ks={ks},no2sm={two or 0})" for us, bn, ks, two in res[:8]), flush=True)
