#!/usr/bin/env python
"""Device time of the optional rows at config-2 size (B = 256): training step with / without the encoder-targets head
(encoder_1_projection = [225], 13 targets), with Luong / Bahdanau attention, and one e2t_input_saliency call.
usage: python tools/time_optional.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench as Bn
from ecog2txt_b200 import Engine, EngineConfig
from ecog2txt_b200.params import init_engine
from ecog2txt_b200.synthetic import SyntheticCorpus, load_vocab

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = 256
corpus = SyntheticCorpus(load_vocab(size=Bn.GEO["V"]), T=Bn.T_FRAMES, C=256, seed=0)
b = corpus.batch(B, seed=0, L=Bn.L_TGT)
hx, hy = b["encoder_inputs"], b["decoder_targets"]
x, y = torch.from_numpy(hx).cuda(), torch.from_numpy(hy).cuda()
aux = torch.randn(B, Bn.T_FRAMES, 13, device="cuda")
ntok = float((y != 0).sum())
stream = torch.cuda.current_stream()


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, kw in (("plain", {}), ("aux_head_225_13", dict(aux_layer=1, aux_hidden=225, aux_F=13, aux_kind="gaussian", aux_penalty=0.1)),
                 ("luong", dict(attention="luong")), ("bahdanau", dict(attention="bahdanau"))):
    eng = Engine(EngineConfig(**Bn.GEO, max_B=B, max_T=Bn.T_FRAMES, max_L=20, max_beam=1, ff_dropout=Bn.FF_DROPOUT,
                              rnn_dropout=Bn.RNN_DROPOUT, **kw))
    eng.set_stream(stream.cuda_stream)
    init_engine(eng, seed=1)

    def step():
        if "aux_F" in kw:
            eng.set_encoder_targets(aux)
        eng.train_step_grads(x, None, y, seed=1, want_loss=False)
        eng.adam_ema_step(1.0 / ntok)

    ms = timed(step, iters)
    line = f"{name:18s} train step {ms:7.3f} ms  ({B / ms * 1e3:8.0f} utt/s)"
    if name in ("plain", "aux_head_225_13"):
        def sal():
            if "aux_F" in kw:
                eng.set_encoder_targets(aux)
            eng.input_saliency(hx, None, hy, want_dx=False, want_norms=True)
        line += f"   saliency (host in, norms out) {timed(sal, 5):7.3f} ms"
    print(line, flush=True)
    if "attention" in kw:
        for env in ("", "1"):
            if env:
                os.environ["E2T_ATTN_BLOCK"] = "1"
            else:
                os.environ.pop("E2T_ATTN_BLOCK", None)
            ms_v = timed(step, iters)
            eng.profile_enable(True)
            for _ in range(2):
                step()
            rep = eng.profile_report()
            eng.profile_enable(False)
            att = {k: round(1e3 * t / n, 1) for k, (n, t) in rep.items() if "attn" in k or "tanh" in k}
            print(f"    {'block-per-row kernels' if env else 'warp-per-row kernels '}: step {ms_v:7.3f} ms; us/launch {att}", flush=True)
        os.environ.pop("E2T_ATTN_BLOCK", None)
    eng.close()
