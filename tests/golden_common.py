"""Checks an engine build (emulated on CPU, CUDA on the GPU) against tests/golden/seq2seq_tiny.npz."""
import os

import numpy as np

import parity_common as pc
from ecog2txt_b200 import _lib

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seq2seq_tiny.npz")


def load():
    z = np.load(GOLD)
    P = {k[2:].replace("|", "/"): z[k] for k in z.files if k.startswith("P|")}
    return z, P


def check_engine_against_golden(lib, backend="simt", tol=2e-4):
    z, P = load()
    B, T, L = z["x"].shape[0], z["x"].shape[1], z["y"].shape[1]
    eng = pc.engine_for(pc.TINY, lib, B, T, 6, max_beam=3, gemm_backend=backend)
    eng.set_all(P)
    loss, ntok = eng.train_step_grads(z["x"], None, z["y"], seed=0)
    assert ntok == int(z["ntok"])
    assert abs(loss - float(z["loss"])) <= tol * abs(float(z["loss"]))
    assert (eng.activation("lens", (B,), np.int32) == z["lens"]).all()
    assert pc.rel_err(eng.activation("final_h", z["final_h"].shape), z["final_h"]) <= tol
    assert pc.rel_err(eng.activation("final_c", z["final_c"].shape), z["final_c"]) <= tol
    for k, v in eng.get_all(_lib.GRAD).items():
        assert pc.rel_err(v, z["G|" + k.replace("/", "|")]) <= 5 * tol, k
    eng.adam_ema_step(1.0 / ntok)
    for k, v in eng.get_all(_lib.VALUE).items():
        assert np.allclose(v, z["W1|" + k.replace("/", "|")], rtol=1e-4, atol=1e-6), k
    for k, v in eng.get_all(_lib.EMA).items():
        assert np.allclose(v, z["S1|" + k.replace("/", "|")], rtol=1e-4, atol=1e-6), k
    # decoding with the ORIGINAL weights
    eng.set_all(P)
    toks, logp = eng.greedy_decode(z["x"], None, max_len=6, temperature=0.7)
    assert (toks == z["greedy_tokens"]).all()
    assert np.abs(logp - z["greedy_logp"]).max() < 2e-3
    bt, bs = eng.beam_decode(z["x"], None, beam=3, max_len=6, temperature=0.7)
    assert np.abs(bs - z["beam_scores"]).max() < 5e-3
    assert (bt[:, 0] == z["beam_tokens"][:, 0]).all()
    eng.close()


GOLD_OPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seq2seq_tiny_optional.npz")
GEO_OPT = dict(pc.TINY_AUX, attention="bahdanau")


def load_optional():
    z = np.load(GOLD_OPT)
    P = {k[2:].replace("|", "/"): z[k] for k in z.files if k.startswith("P|")}
    return z, P


def check_engine_against_optional_golden(lib, backend="simt", tol=2e-4):
    """Encoder-targets head + Bahdanau attention + saliency against tests/golden/seq2seq_tiny_optional.npz."""
    z, P = load_optional()
    B, T, L = z["x"].shape[0], z["x"].shape[1], z["y"].shape[1]
    eng = pc.engine_for(GEO_OPT, lib, B, T, 6, gemm_backend=backend)
    eng.set_all(P)
    eng.set_encoder_targets(z["aux"])
    loss, ntok = eng.train_step_grads(z["x"], None, z["y"], seed=0)
    ld, nt, la, nf = eng.last_losses()
    assert ntok == int(z["ntok"]) and nf == int(z["aux_frames"])
    for got, key in ((loss, "loss"), (ld, "decoder_loss"), (la, "aux_loss")):
        assert abs(got - float(z[key])) <= tol * abs(float(z[key])), key
    for k, v in eng.get_all(_lib.GRAD).items():
        assert pc.rel_err(v, z["G|" + k.replace("/", "|")]) <= 5 * tol, k
    eng.set_encoder_targets(z["aux"])
    dx, sq = eng.input_saliency(z["x"], None, z["y"])
    assert pc.rel_err(dx, z["dx"]) <= 5 * tol
    assert pc.rel_err(sq, (z["dx"].astype(np.float64) ** 2).sum(1)) <= 10 * tol
    toks, logp = eng.greedy_decode(z["x"], None, max_len=6, temperature=0.7)
    assert (toks == z["greedy_tokens"]).all()
    assert np.abs(logp - z["greedy_logp"]).max() < 2e-3
    eng.close()


GOLD_PROJ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seq2seq_tiny_proj.npz")


def load_proj():
    z = np.load(GOLD_PROJ)
    P = {k[2:].replace("|", "/"): z[k] for k in z.files if k.startswith("P|")}
    return z, P


def check_engine_against_proj_golden(lib, backend="simt", tol=2e-4):
    """The model with a hidden decoder_projection layer against tests/golden/seq2seq_tiny_proj.npz."""
    z, P = load_proj()
    B, T = z["x"].shape[0], z["x"].shape[1]
    eng = pc.engine_for(pc.TINY_PROJ, lib, B, T, 6, max_beam=3, gemm_backend=backend)
    eng.set_all(P)
    loss, ntok = eng.train_step_grads(z["x"], None, z["y"], seed=0)
    assert ntok == int(z["ntok"])
    assert abs(loss - float(z["loss"])) <= tol * abs(float(z["loss"]))
    for k, v in eng.get_all(_lib.GRAD).items():
        assert pc.rel_err(v, z["G|" + k.replace("/", "|")]) <= 5 * tol, k
    eng.adam_ema_step(1.0 / ntok)
    for k, v in eng.get_all(_lib.VALUE).items():
        assert np.allclose(v, z["W1|" + k.replace("/", "|")], rtol=1e-4, atol=1e-6), k
    for k, v in eng.get_all(_lib.EMA).items():
        assert np.allclose(v, z["S1|" + k.replace("/", "|")], rtol=1e-4, atol=1e-6), k
    eng.set_all(P)
    toks, logp = eng.greedy_decode(z["x"], None, max_len=6, temperature=0.7)
    assert (toks == z["greedy_tokens"]).all()
    assert np.abs(logp - z["greedy_logp"]).max() < 2e-3
    bt, bs = eng.beam_decode(z["x"], None, beam=3, max_len=6, temperature=0.7)
    assert np.abs(bs - z["beam_scores"]).max() < 5e-3
    assert (bt[:, 0] == z["beam_tokens"][:, 0]).all()
    eng.close()
