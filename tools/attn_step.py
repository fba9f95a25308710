#!/usr/bin/env python
"""One training step of the config-2 model with an attention module (for ncu captures of the attention kernels).
usage: python tools/attn_step.py [luong|bahdanau] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from ecog2txt_b200 import Engine, EngineConfig
from ecog2txt_b200.params import init_engine
from ecog2txt_b200.synthetic import SyntheticCorpus, load_vocab

kind = sys.argv[1] if len(sys.argv) > 1 else "luong"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B = 256
eng = Engine(EngineConfig(**Bn.GEO, max_B=B, max_T=Bn.T_FRAMES, max_L=20, max_beam=1, ff_dropout=Bn.FF_DROPOUT,
                          rnn_dropout=Bn.RNN_DROPOUT, attention=kind))
init_engine(eng, seed=1)
corpus = SyntheticCorpus(load_vocab(size=Bn.GEO["V"]), T=Bn.T_FRAMES, C=256, seed=0)
b = corpus.batch(B, seed=0, L=Bn.L_TGT)
x, y = torch.from_numpy(b["encoder_inputs"]).cuda(), torch.from_numpy(b["decoder_targets"]).cuda()
for i in range(steps):
    eng.train_step_grads(x, None, y, seed=i, want_loss=False)
    eng.adam_ema_step(1.0 / float((y != 0).sum()))
eng.sync()
print("done", eng.launch_counts())
