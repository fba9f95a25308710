#!/usr/bin/env python
"""Run N training steps of the bench workload (device-resident inputs); for use under ncu / timing experiments.
usage: python tools/one_step.py [steps] [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from ecog2txt_b200 import Engine, EngineConfig
from ecog2txt_b200.params import init_engine
from ecog2txt_b200.synthetic import SyntheticCorpus, load_vocab

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
eng = Engine(EngineConfig(**Bn.GEO, max_B=B, max_T=Bn.T_FRAMES, max_L=20, max_beam=8, ff_dropout=Bn.FF_DROPOUT,
                          rnn_dropout=Bn.RNN_DROPOUT))
_st = torch.cuda.Stream(priority=-1); torch.cuda.set_stream(_st)
eng.set_stream(_st.cuda_stream)
init_engine(eng, seed=1)
corpus = SyntheticCorpus(load_vocab(size=Bn.GEO["V"]), T=Bn.T_FRAMES, C=256, seed=0)
b = corpus.batch(B, seed=0, L=Bn.L_TGT)
x, y = torch.from_numpy(b["encoder_inputs"]).cuda(), torch.from_numpy(b["decoder_targets"]).cuda()
ntok = float((y != 0).sum())
torch.cuda.synchronize()
for i in range(steps):
    t0 = time.perf_counter()
    eng.train_step_grads(x, None, y, seed=i, want_loss=False)
    eng.adam_ema_step(1.0 / ntok)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"step {i}: host enqueue {1e3 * (t1 - t0):.2f} ms, +sync {1e3 * (t2 - t1):.2f} ms, launches so far {eng.launch_counts()[0]}", flush=True)
