#!/usr/bin/env python
"""Per-step timeline of the persistent recurrent kernels at the bench shape (diagnostic; needs a GPU).
E2T_REC_DEBUG=<n> python tools/rec_timeline.py   -> stderr tables for the first n fwd and n bwd launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ecog2txt_b200 import Engine, EngineConfig
from ecog2txt_b200.params import init_engine
B = int(os.environ.get("B", 256))
eng = Engine(EngineConfig(max_B=B, max_T=400, max_L=20, ff_dropout=0.1, rnn_dropout=0.5))
init_engine(eng, 1)
rs = np.random.RandomState(0)
x = rs.randn(B, 400, 256).astype(np.float32)
y = rs.randint(3, 1806, size=(B, 11)).astype(np.int32); y[:, -1] = 1
for i in range(int(os.environ.get("STEPS", 3))):
    print("loss", eng.train_step_grads(x, None, y, seed=i))
