#!/usr/bin/env python
"""Time the tcgen05 GEMM at the shapes of one training step and at 8192^3, next to cuBLAS tf32 / bf16 through torch
(the MEASURED tensor peaks bench.py's roofline uses).  Operands are uniform random (E2T_BENCH_ZERO_OPERANDS=1 restores the
all-zero operands of round 1, for the record).  usage: python tools/bench_gemm.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ecog2txt_b200 import Engine, EngineConfig
eng = Engine(EngineConfig(subnet_ids=(7,), subnet_C=(8,), subnet_W=(4,), E=8, H=(16,), D=8, Hd=32, V=16, max_B=2, max_T=8, max_L=4))
SH = [(8704, 1600, 800, 0, 0.0), (8704, 800, 1600, 0, 0.0), (8704, 800, 1600, 0, 1.0), (8704, 1600, 100, 0, 0.0),
      (8704, 100, 1600, 0, 1.0), (256, 3200, 800, 0, 1.0), (256, 800, 3200, 0, 0.0), (2816, 1806, 800, 0, 0.0),
      (800, 1600, 8704, 1, 0.0), (400, 1600, 8448, 1, 0.0), (100, 1600, 8704, 1, 0.0), (1806, 800, 2816, 1, 0.0),
      (8192, 8192, 8192, 0, 0.0), (8192, 8192, 8192, 1, 0.0)]
out = {"own_tcgen05_tf32": {}, "operands": "zeros" if os.environ.get("E2T_BENCH_ZERO_OPERANDS") else "uniform(-0.5, 0.5)"}
for M, N, K, tn, beta in SH:
    ms = eng.bench_gemm(M, N, K, bool(tn), beta, 20)
    tf = 2.0 * M * N * K / ms / 1e9
    out["own_tcgen05_tf32"][f"{'TN' if tn else 'NT'}[{M},{N},{K}]beta{beta:.0f}"] = {"us": ms * 1e3, "tflops": tf}
    print(f"{'TN' if tn else 'NT'} [{M:5d},{N:5d},{K:5d}] beta={beta:.0f}: {ms * 1e3:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)


def cublas(dtype, tf32, n=8192, iters=30):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.rand(n, n, device="cuda", dtype=dtype) - 0.5
    b = torch.rand(n, n, device="cuda", dtype=dtype) - 0.5
    for _ in range(5):
        a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    return 2.0 * n ** 3 * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12


out["cublas_tf32_8192_tflops"] = cublas(torch.float32, True)
out["cublas_bf16_8192_tflops"] = cublas(torch.bfloat16, False)
out["cublas_fp16_8192_tflops"] = cublas(torch.float16, False)
print(json.dumps({k: v for k, v in out.items() if k != "own_tcgen05_tf32"}))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
