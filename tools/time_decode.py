#!/usr/bin/env python
"""Per-kernel CUDA-event breakdown of one greedy and one beam-8 decode of 256 utterances (config 2).
usage: python tools/time_decode.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from ecog2txt_b200 import Engine, EngineConfig
from ecog2txt_b200.params import init_engine
from ecog2txt_b200.synthetic import SyntheticCorpus, load_vocab

B = 256
eng = Engine(EngineConfig(**Bn.GEO, max_B=B, max_T=Bn.T_FRAMES, max_L=20, max_beam=8))
stream = torch.cuda.current_stream()
eng.set_stream(stream.cuda_stream)
init_engine(eng, seed=1)
corpus = SyntheticCorpus(load_vocab(size=Bn.GEO["V"]), T=Bn.T_FRAMES, C=256, seed=0)
x = torch.from_numpy(corpus.batch(B, seed=0, L=Bn.L_TGT)["encoder_inputs"]).cuda()
for name, fn in (("greedy", lambda: eng.greedy_decode(x, None, max_len=20, want_logp=False)),
                 ("beam8", lambda: eng.beam_decode(x, None, beam=8, max_len=20))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 5:.3f} ms per batch of {B}")
    eng.profile_enable(True)
    fn()
    rep = eng.profile_report()
    eng.profile_enable(False)
    for k, (n, t) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"   {t:8.3f} ms  n={n:4d}  {1e3 * t / n:8.1f} us  {k}")
