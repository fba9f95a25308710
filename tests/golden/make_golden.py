#!/usr/bin/env python
"""Generates tests/golden/seq2seq_tiny.npz from the in-repo oracle (oracle/seq2seq_oracle.py, fp64 then cast).

PARITY UNPINNED: the reference's own implementation (machine_learning on TF1.15) cannot be imported here
(SURVEY.md section 8c), so these vectors pin the ORACLE (and through it every engine build) against regressions --
they are not outputs of the reference.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_common as pc  # noqa: E402
from oracle import seq2seq_oracle as O  # noqa: E402


def main():
    geo = pc.TINY
    ocfg = O.OracleConfig(**geo)
    P32 = pc.make_params(ocfg, eos_bias=-1.0)
    P = {k: v.double() for k, v in P32.items()}
    B, T, L = 5, 19, 5
    x, lens, y = pc.make_batch(ocfg, B, T, L, seed=11)
    xt, yt = torch.from_numpy(x).double(), torch.from_numpy(y).long()
    loss, ntok, g, acts = O.loss_and_grads(ocfg, P, xt, None, yt)
    out = {"x": x, "lens": lens, "y": y, "loss": np.float64(loss), "ntok": np.int64(ntok),
           "final_h": acts["final_h"].numpy(), "final_c": acts["final_c"].numpy(), "conv_out": acts["conv_out"].numpy()}
    for k, v in P32.items():
        out["P|" + k.replace("/", "|")] = v.numpy()
    for k, v in g.items():
        out["G|" + k.replace("/", "|")] = v.numpy()
    opt = O.AdamEMA(ocfg, P)
    P2 = dict(P)
    opt.step(P2, g, 1.0 / ntok)
    for k in P2:
        out["W1|" + k.replace("/", "|")] = P2[k].numpy()
        out["S1|" + k.replace("/", "|")] = opt.ema[k].numpy()
    toks, logp, _ = O.greedy_decode(ocfg, P, xt, None, max_len=6, temperature=0.7)
    out["greedy_tokens"], out["greedy_logp"] = toks.numpy(), logp.numpy()
    bt, bs = O.beam_decode(ocfg, P, xt, None, beam=3, max_len=6, temperature=0.7)
    out["beam_tokens"], out["beam_scores"] = bt.numpy(), bs.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "seq2seq_tiny.npz"), **out)
    print("wrote seq2seq_tiny.npz: loss", loss, "ntok", ntok)
    make_optional()
    make_hidden_projection()


def make_optional():
    """seq2seq_tiny_optional.npz: the optional rows together -- encoder-targets head (A6, Gaussian, hidden layer) +
    Bahdanau attention (A7) + input saliency under both penalties (A13)."""
    geo = dict(pc.TINY_AUX, attention="bahdanau")
    ocfg = O.OracleConfig(**geo)
    P32 = pc.make_params(ocfg, eos_bias=-1.0)
    P = {k: v.double() for k, v in P32.items()}
    B, T, L = 4, 19, 5
    x, lens, y = pc.make_batch(ocfg, B, T, L, seed=12)
    aux = pc.make_aux_targets(ocfg, lens, T, seed=6)
    xt, yt, at = torch.from_numpy(x).double(), torch.from_numpy(y).long(), torch.from_numpy(aux).double()
    loss, ntok, g, acts = O.loss_and_grads(ocfg, P, xt, None, yt, aux_targets=at)
    out = {"x": x, "lens": lens, "y": y, "aux": aux, "loss": np.float64(loss), "ntok": np.int64(ntok),
           "decoder_loss": np.float64(acts["decoder_loss"]), "aux_loss": np.float64(acts["aux_loss"]),
           "aux_frames": np.int64(acts["aux_frames"]),
           "dx": O.input_gradients(ocfg, P, xt, None, yt, aux_targets=at).numpy()}
    for k, v in P32.items():
        out["P|" + k.replace("/", "|")] = v.numpy()
    for k, v in g.items():
        out["G|" + k.replace("/", "|")] = v.numpy()
    toks, logp, _ = O.greedy_decode(ocfg, P, xt, None, max_len=6, temperature=0.7)
    out["greedy_tokens"], out["greedy_logp"] = toks.numpy(), logp.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "seq2seq_tiny_optional.npz"), **out)
    print("wrote seq2seq_tiny_optional.npz: loss", loss, "=", acts["decoder_loss"], "+", acts["aux_loss"])


def make_hidden_projection(only=True):
    """seq2seq_tiny_proj.npz: the model with one hidden decoder_projection layer (layer_sizes['decoder_projection'] = [7],
    mochastar_word_sequence.yaml:65): loss, every gradient, one Adam+EMA step, greedy and beam decode."""
    ocfg = O.OracleConfig(**pc.TINY_PROJ)
    P32 = pc.make_params(ocfg, eos_bias=-1.0)
    P = {k: v.double() for k, v in P32.items()}
    B, T, L = 5, 19, 5
    x, lens, y = pc.make_batch(ocfg, B, T, L, seed=13)
    xt, yt = torch.from_numpy(x).double(), torch.from_numpy(y).long()
    loss, ntok, g, acts = O.loss_and_grads(ocfg, P, xt, None, yt)
    out = {"x": x, "lens": lens, "y": y, "loss": np.float64(loss), "ntok": np.int64(ntok)}
    for k, v in P32.items():
        out["P|" + k.replace("/", "|")] = v.numpy()
    for k, v in g.items():
        out["G|" + k.replace("/", "|")] = v.numpy()
    opt = O.AdamEMA(ocfg, P)
    P2 = dict(P)
    opt.step(P2, g, 1.0 / ntok)
    for k in P2:
        out["W1|" + k.replace("/", "|")] = P2[k].numpy()
        out["S1|" + k.replace("/", "|")] = opt.ema[k].numpy()
    toks, logp, _ = O.greedy_decode(ocfg, P, xt, None, max_len=6, temperature=0.7)
    out["greedy_tokens"], out["greedy_logp"] = toks.numpy(), logp.numpy()
    bt, bs = O.beam_decode(ocfg, P, xt, None, beam=3, max_len=6, temperature=0.7)
    out["beam_tokens"], out["beam_scores"] = bt.numpy(), bs.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "seq2seq_tiny_proj.npz"), **out)
    print("wrote seq2seq_tiny_proj.npz: loss", loss, "ntok", ntok)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "proj":      # add the newer fixture without touching the committed ones
        make_hidden_projection()
    else:
        main()
