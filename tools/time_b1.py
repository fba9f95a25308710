"""Online-predictor latency (N3): one host utterance in, tokens out, wall clock per call at the config-2 geometry,
through the CUDA-graph replay; E2T_NO_SMALL_DECODE=1 runs the batched decode step instead of the small-batch kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ecog2txt_b200 import Engine, EngineConfig
from ecog2txt_b200.params import init_engine

geo = dict(subnet_ids=(400,), subnet_C=(256,), subnet_W=(12,), E=100, H=(400, 400, 400), D=150, Hd=800, V=1806)
for B in (1, 4, 8):
    eng = Engine(EngineConfig(**geo, max_B=B, max_T=400, max_L=20))
    init_engine(eng, seed=1)
    x = np.random.RandomState(0).randn(B, 400, 256).astype(np.float32)
    for _ in range(4):
        eng.greedy_decode(x, None, max_len=20, want_logp=False)
    t0 = time.perf_counter()
    for _ in range(50):
        eng.greedy_decode(x, None, max_len=20, want_logp=False)
    dt = (time.perf_counter() - t0) / 50
    print(f"B={B}: {1e3 * dt:.3f} ms per call ({1e3 * dt / B:.3f} ms per utterance), graph replays {eng.counter('decode_graph_replays')}, "
          f"small-batch kernels {'off' if os.environ.get('E2T_NO_SMALL_DECODE') else 'on'}", flush=True)
    eng.close()
