"""diagnostic: per-tensor gradient errors of the encoder-targets head cases on the tensor-core backend"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import parity_common as pc
from ecog2txt_b200 import _lib
from oracle import seq2seq_oracle as O

lib = _lib.load()
for env in ("", "1"):
    if env:
        os.environ["E2T_AUX_FP32"] = "1"
    for name, geo, B, T, L, ff, rnn in (("gauss", pc.MEDIUM_AUX, 16, 96, 6, 0, 0), ("gauss_drop", pc.MEDIUM_AUX, 16, 96, 6, .1, .5),
                                        ("cat", pc.MEDIUM_AUX_CAT, 16, 96, 6, 0, 0), ("noaux", pc.MEDIUM, 16, 96, 6, 0, 0)):
        ocfg = O.OracleConfig(**geo)
        P = pc.make_params(ocfg)
        eng = pc.engine_for(geo, lib, B, T, L, ff_dropout=ff, rnn_dropout=rnn, gemm_backend="auto")
        eng.set_all({k: v.numpy() for k, v in P.items()})
        x, lens, y = pc.make_batch(ocfg, B, T, L)
        T2 = -(-T // 12)
        masks = O.make_masks(ocfg, 3, B, T2, L, ff, rnn, torch.float32) if (ff or rnn) else None
        aux = pc.make_aux_targets(ocfg, lens, T) if ocfg.aux_F > 0 and ocfg.aux_layer >= 0 else None
        lo, no, g, acts = O.loss_and_grads(ocfg, P, torch.from_numpy(x), None, torch.from_numpy(y).long(), masks=masks,
                                           aux_targets=None if aux is None else torch.from_numpy(aux))
        if aux is not None:
            eng.set_encoder_targets(aux)
        loss, ntok = eng.train_step_grads(x, None, y, seed=3)
        G = eng.get_all(_lib.GRAD)
        errs = {k.split("seq2seq/")[1][:40]: round(pc.rel_err(v, g[k].numpy()), 4) for k, v in G.items()}
        print(f"fp32={env!r} {name}: loss {loss:.4f} vs {lo:.4f}  worst {max(errs.values())}")
        print("   ", {k: v for k, v in errs.items() if v > 5e-3})
        eng.close()
