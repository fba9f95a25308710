"""Parity checks shared by the CPU (kernel-emulation) and GPU test files: the engine under test is
driven through the C-ABI and compared with the oracle on the same seeded inputs."""
import numpy as np
import torch

from ecog2txt_b200 import Engine, EngineConfig, _lib
from oracle import seq2seq_oracle as O

TINY = dict(subnet_ids=(7,), subnet_C=(6,), subnet_W=(4,), E=5, H=(8, 8), D=6, Hd=16, V=11)
SMALL = dict(subnet_ids=(400,), subnet_C=(32,), subnet_W=(12,), E=20, H=(32, 32, 32), D=12, Hd=64, V=60)
# aligned shapes (multiples of 4 / 16-byte rows) so that the tcgen05 GEMM path is eligible on the GPU
MEDIUM = dict(subnet_ids=(400,), subnet_C=(64,), subnet_W=(12,), E=32, H=(64, 64), D=24, Hd=128, V=200)
# full-width recurrent layers (H = 400 per direction, decoder 800) on a narrow input: exercises the
# persistent tcgen05 recurrent kernels at the config-2 layer shape while the oracle still runs in seconds
WIDE = dict(subnet_ids=(400,), subnet_C=(32,), subnet_W=(12,), E=100, H=(400, 400), D=24, Hd=800, V=200)
# optional Luong attention (A7) on top of the same geometries
# one hidden layer in the vocabulary projection (layer_sizes['decoder_projection'], yaml:65)
TINY_PROJ = dict(TINY, proj_hidden=7)
MEDIUM_PROJ = dict(MEDIUM, proj_hidden=96)
TINY_ATTN = dict(TINY, attention="luong")
MEDIUM_ATTN = dict(MEDIUM, attention="luong")
TINY_BAH = dict(TINY, attention="bahdanau")
MEDIUM_BAH = dict(MEDIUM, attention="bahdanau")
# A6 encoder-targets head on layer 1 (hidden projection + output layer) / straight on the top layer, both kinds
TINY_AUX = dict(TINY, aux_layer=1, aux_hidden=7, aux_F=3, aux_kind="gaussian", aux_penalty=0.7)
TINY_AUX_CAT = dict(TINY, aux_layer=0, aux_hidden=0, aux_F=5, aux_kind="categorical", aux_penalty=1.3)
MEDIUM_AUX = dict(MEDIUM, aux_layer=1, aux_hidden=36, aux_F=13, aux_kind="gaussian", aux_penalty=0.5)
MEDIUM_AUX_CAT = dict(MEDIUM, aux_layer=1, aux_hidden=36, aux_F=42, aux_kind="categorical", aux_penalty=1.0)
TWO_SUBJ = dict(subnet_ids=(400, 401), subnet_C=(6, 10), subnet_W=(4, 3), E=5, H=(8,), D=6, Hd=16, V=11)
# BASELINE.json config 2 exactly (mochastar_word_sequence.yaml:62-75,84-85,89): the geometry bench.py is quoted on
FULL = dict(subnet_ids=(400,), subnet_C=(256,), subnet_W=(12,), E=100, H=(400, 400, 400), D=150, Hd=800, V=1806)

# Tolerances = <= 10x the errors measured on B200 (profiles/parity_r2.json), relative to each tensor's largest entry:
#   fp32 CUDA-core backend ("simt"): measured loss <= 1.2e-7, state <= 5.5e-7, gradients <= 6e-7 (attention: 4.1e-5 on the
#   Bahdanau query weights with dropout .1/.5, 7.9e-6 without)
#   tensor-core backend ("auto"; 11-bit-significand operands, fp32 accumulate), small geometries: loss <= 1.0e-5,
#   state <= 1.3e-3, gradients <= 3.0e-3; config 2 (3 layers x 34 steps, K up to 3072): see FULL_TOL
SIMT_TOL = dict(loss=1e-6, state=5e-6, grad=2e-4)
TC_TOL = dict(loss=1e-4, state=1e-2, grad=3e-2)
# (log-probabilities of the TRAINED model: logits an order of magnitude larger than with random weights, measured 8.1e-3)
# (beam-8 scores of the trained model -- sums of up to 12 such log-probabilities: 3.3e-2 measured, all beams identical)
FULL_TOL = dict(loss=2e-5, state=1e-2, grad=3e-2, logp=5e-3, logp_trained=3e-2, beam_score=1e-2, beam_score_trained=1e-1)


def tols(backend):
    return SIMT_TOL if backend == "simt" else TC_TOL

# ---- achieved-error record -------------------------------------------------------------------------------------
# Every oracle comparison appends its measured errors here; tests/test_gpu_zz_fit.py's last test (or the session
# teardown in conftest.py) writes the table to gpurun_out/parity_r2.json, which is committed as profiles/parity_r2.json.
# Tolerances in the tests are <= 10x the errors recorded there.
PARITY_RECORD = {}


def record(name, **errs):
    PARITY_RECORD.setdefault(name, {}).update({k: (float(v) if not isinstance(v, (dict, list, str)) else v) for k, v in errs.items()})


def dump_record(path):
    import json
    import os
    if not PARITY_RECORD:
        return
    os.makedirs(os.path.dirname(path), exist_ok=True)
    old = {}
    if os.path.exists(path):
        try:
            old = json.load(open(path))
        except Exception:      # noqa: BLE001 -- a truncated file from a killed run is simply replaced
            old = {}
    old.update(PARITY_RECORD)
    with open(path, "w") as f:
        json.dump(old, f, indent=1, sort_keys=True)


def make_params(ocfg, seed=1, bias_scale=0.1, eos_bias=None):
    P = O.init_params(ocfg, seed)
    g = torch.Generator().manual_seed(seed + 100)
    for k in P:
        if P[k].ndim == 1:
            P[k] = (torch.randn(P[k].shape, generator=g) * bias_scale).float()
    if eos_bias is not None:  # make greedy hypotheses longer than one token
        pb = O.proj_names(ocfg)[1] + "/biases"
        P[pb][ocfg.eos_id] = eos_bias
        P[pb][ocfg.pad_id] = -20.0
    return P


def make_batch(ocfg, B, T, L, subnet=0, seed=0, ragged=True):
    rs = np.random.RandomState(seed)
    C = ocfg.subnet_C[subnet]
    x = rs.randn(B, T, C).astype(np.float32)
    lens = rs.randint(max(1, T // 3), T + 1, size=B) if ragged else np.full(B, T)
    lens[0] = T
    for b in range(B):
        x[b, lens[b]:] = 0.0
    y = np.zeros((B, L), np.int32)
    for b in range(B):
        n = rs.randint(1, L) if ragged else L - 1
        y[b, :n] = rs.randint(3, ocfg.V, size=n)
        y[b, n] = ocfg.eos_id
    return x, lens.astype(np.int32), y


def make_aux_targets(ocfg, lens, T, seed=5):
    """Encoder targets at the input frame rate, zero / pad padded past each utterance's length."""
    rs = np.random.RandomState(seed)
    B = len(lens)
    if ocfg.aux_kind == "gaussian":
        a = rs.randn(B, T, ocfg.aux_F).astype(np.float32)
    else:
        a = rs.randint(0, ocfg.aux_F, size=(B, T)).astype(np.int32)   # class 0 = pad: masked frames inside the length too
    for b in range(B):
        a[b, lens[b]:] = 0
    return a


def engine_for(geo, lib, B, T, L, **kw):
    ecfg = EngineConfig(**geo, max_B=B, max_T=T, max_L=L, **kw)
    return Engine(ecfg, lib=lib)


def rel_err(a, ref):
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-6))


def check_train_step(lib, geo, B, T, L, ff=0.0, rnn=0.0, seed=3, tol=None, backend="simt", give_lens=False,
                     subnet=0, grad_tol=None, name=None, batch=None, flip_margin=1e-3, loss_tol=None):
    """One training step through the C-ABI against O.loss_and_grads.  Bounds (defaults from tols(backend)): loss_tol on the
    relative error of the loss, tol on the final encoder state, grad_tol on every gradient tensor relative to that tensor's
    largest entry.  The achieved errors are recorded under `name` (PARITY_RECORD)."""
    t0 = tols(backend)
    loss_tol = t0["loss"] if loss_tol is None else loss_tol
    grad_tol = t0["grad"] if grad_tol is None else grad_tol
    tol = t0["state"] if tol is None else tol
    ocfg = O.OracleConfig(**geo)
    P = make_params(ocfg)
    eng = engine_for(geo, lib, B, T, L, ff_dropout=ff, rnn_dropout=rnn, gemm_backend=backend)
    eng.set_all({k: v.numpy() for k, v in P.items()})
    x, lens, y = batch if batch is not None else make_batch(ocfg, B, T, L, subnet=subnet)
    W = ocfg.subnet_W[subnet]
    T2 = -(-T // W)
    masks = O.make_masks(ocfg, seed, B, T2, L, ff, rnn, torch.float32) if (ff > 0 or rnn > 0) else None
    aux = make_aux_targets(ocfg, lens, T) if ocfg.aux_layer >= 0 and ocfg.aux_F > 0 else None
    if aux is not None:
        eng.set_encoder_targets(aux)
    loss, ntok = eng.train_step_grads(x, lens if give_lens else None, y, subnet=subnet, seed=seed)
    # the ReLU pattern of the conv output as the engine computed it ([T',B,E] -> [B,T',E]; dropped units are zero in both
    # and irrelevant).  The oracle differentiates the SAME piecewise-linear branch (see O.temporal_conv); the patterns
    # themselves must agree except where the oracle's pre-activation is within `flip_margin` of zero.
    conv_eng = eng.activation("conv_out", (T2, B, ocfg.E)).transpose(1, 0, 2)
    relu_mask = torch.from_numpy(np.ascontiguousarray(conv_eng != 0)) if ocfg.conv_act == "relu" else None
    lo, no, g, acts = O.loss_and_grads(ocfg, P, torch.from_numpy(x), None, torch.from_numpy(y).long(),
                                       subnet=subnet, masks=masks,
                                       aux_targets=None if aux is None else torch.from_numpy(aux), conv_relu_mask=relu_mask)
    n_flips, flip_z = 0, 0.0
    if relu_mask is not None:
        z = acts["conv_preact"].numpy()
        kept = np.ones_like(z, bool) if masks is None or "conv" not in masks else (masks["conv"].numpy() != 0)
        valid = (np.arange(T2)[None, :] < acts["lens2"].numpy()[:, None])[:, :, None]      # frames past len' feed nothing
        flips = ((z > 0) != relu_mask.numpy()) & kept & valid
        n_flips = int(flips.sum())
        flip_z = float(np.abs(z[flips]).max()) if n_flips else 0.0
        assert flip_z <= flip_margin, (n_flips, flip_z)
        assert n_flips <= max(4, 1e-4 * z.size), n_flips
    if aux is not None:
        ld, nt, la, nf = eng.last_losses()
        assert nf == acts["aux_frames"] and nt == no
        assert abs(la - acts["aux_loss"]) <= 10 * loss_tol * max(abs(acts["aux_loss"]), 1.0), (la, acts["aux_loss"])
        assert abs(ld - acts["decoder_loss"]) <= loss_tol * max(abs(acts["decoder_loss"]), 1.0)
    assert ntok == no
    e_loss = abs(loss - lo) / max(abs(lo), 1.0)
    assert (eng.activation("lens", (B,), np.int32) == lens).all()
    e_h = rel_err(eng.activation("final_h", (B, ocfg.Hd)), acts["final_h"].numpy())
    e_c = rel_err(eng.activation("final_c", (B, ocfg.Hd)), acts["final_c"].numpy()) if "final_c" in acts else 0.0
    G = eng.get_all(_lib.GRAD)
    errs = {k: rel_err(v, g[k].numpy()) for k, v in G.items()}
    worst = max(errs.values())
    opt = ("" if ocfg.attention == "none" else f"_{ocfg.attention}") + \
          ("" if not (ocfg.aux_layer >= 0 and ocfg.aux_F > 0) else f"_aux{ocfg.aux_kind[:3]}{ocfg.aux_hidden}")
    rec_name = name or f"train_step/{backend}/E{ocfg.E}_H{'x'.join(map(str, ocfg.H))}_V{ocfg.V}_C{ocfg.subnet_C[subnet]}{opt}" \
                       f"/B{B}_T{T}_L{L}_ff{ff}_rnn{rnn}"
    record(rec_name, loss=e_loss, final_h=e_h, final_c=e_c, worst_grad=worst, grads=errs, tol=tol, grad_tol=grad_tol, loss_tol=loss_tol,
           conv_relu_flips=n_flips, conv_relu_flip_max_abs_preact=flip_z)
    assert e_loss <= loss_tol, (loss, lo)
    assert e_h <= tol and e_c <= tol, (e_h, e_c)
    for k, e in errs.items():
        assert e <= grad_tol, (k, e)
    eng._last_counters = {k: eng.counter(k) for k in ("launches", "tcgen05_launches", "persistent_rnn_launches")}
    counters = eng._last_counters
    eng.close()
    check_train_step.last_counters = counters
    return worst


def check_saliency(lib, geo, B, T, L, tol=2e-4, backend="simt", which="decoder", use_ema=False):
    """A13: input gradient with one penalty switched on (get_saliencies, /root/reference/ecog2txt/trainers.py:703-732)."""
    import dataclasses
    ocfg = O.OracleConfig(**geo)
    P = make_params(ocfg)
    eng = engine_for(geo, lib, B, T, L, ff_dropout=0.1, rnn_dropout=0.5, gemm_backend=backend)   # dropout must stay off
    eng.set_all({k: v.numpy() for k, v in P.items()}, _lib.EMA if use_ema else _lib.VALUE)
    x, lens, y = make_batch(ocfg, B, T, L)
    has_aux = ocfg.aux_layer >= 0 and ocfg.aux_F > 0
    aux = make_aux_targets(ocfg, lens, T) if has_aux else None
    pd, pa = (1.0, 0.0) if which == "decoder" else (0.0, 1.0)
    ocfg2 = dataclasses.replace(ocfg, penalty_scale=pd, aux_penalty=pa)
    ref = O.input_gradients(ocfg2, P, torch.from_numpy(x), None, torch.from_numpy(y).long(),
                            aux_targets=None if aux is None else torch.from_numpy(aux)).numpy()
    if aux is not None:
        eng.set_encoder_targets(aux)
    dx, sq = eng.input_saliency(x, None, y, use_ema=use_ema, decoder_penalty=pd, aux_penalty=pa)
    assert np.abs(ref).max() > 0
    W = ocfg.subnet_W[0]
    for b in range(B):   # only the zero frames that complete the last window see a gradient
        assert (dx[b, -(-lens[b] // W) * W:] == 0).all(), "no gradient past the last window"
    e = rel_err(dx, ref)
    assert e <= 5 * tol, e
    assert rel_err(sq, (ref ** 2).sum(1)) <= 10 * tol
    eng.close()
    return e


def check_decode(lib, geo, B, T, max_len, beam=0, backend="simt", temperature=0.7, use_ema=False, margin=1e-3,
                 logp_tol=2e-3, score_tol=5e-3, name=None, x=None, eng=None, require_separated=True):
    """Greedy (beam = 0) or beam decode through the C-ABI against the oracle.  Margin-aware: rows whose oracle top-2
    logit gap at some live step is below `margin` may legitimately pick the other token; all others must be IDENTICAL
    (north_star: "decoded token sequences identical under greedy decode").  Achieved errors are recorded under `name`."""
    ocfg = O.OracleConfig(**geo)
    P = make_params(ocfg, eos_bias=-1.0)
    own = eng is None
    if own:
        eng = engine_for(geo, lib, B, T, max_len, max_beam=max(beam, 1), gemm_backend=backend)
    eng.set_all({k: v.numpy() for k, v in P.items()}, _lib.EMA if use_ema else _lib.VALUE)
    if x is None:
        x, lens, _ = make_batch(ocfg, B, T, 4)
    xt = torch.from_numpy(x)
    rec_name = name or f"decode/{backend}/beam{beam}/E{ocfg.E}_H{'x'.join(map(str, ocfg.H))}_V{ocfg.V}/B{B}_T{T}_len{max_len}"
    if beam == 0:
        t_ref, lp_ref, logits = O.greedy_decode(ocfg, P, xt, None, max_len=max_len, temperature=temperature)
        toks, logp = eng.greedy_decode(x, None, max_len=max_len, temperature=temperature, use_ema=use_ema)
        # margin-aware: rows whose oracle top-2 logit gap is tiny at some live step may legitimately differ
        top2 = logits.topk(2, dim=2).values
        gap = (top2[..., 0] - top2[..., 1]).numpy()
        live = np.ones_like(gap, bool)
        tr = t_ref.numpy()
        for b in range(B):
            ends = np.where(tr[b] == ocfg.eos_id)[0]
            if len(ends):
                live[b, ends[0] + 1:] = False
        safe = np.all((gap > margin) | ~live, axis=1)
        e_logp = float(np.abs(logp[safe] - lp_ref.numpy()[safe]).max()) if safe.any() else 0.0
        record(rec_name, safe_rows=float(safe.mean()), rows_identical=float((toks == tr).all(1).mean()),
               safe_rows_identical=float((toks[safe] == tr[safe]).all(1).mean()) if safe.any() else 1.0,
               logp_abs=e_logp, min_live_gap=float(gap[live].min()), logp_tol=logp_tol, margin=margin)
        assert safe.mean() > 0.5
        assert (toks[safe] == tr[safe]).all()
        assert e_logp < logp_tol
        assert (tr != ocfg.pad_id).sum(1).max() > 1, "test params should give multi-token hypotheses"
    else:
        t_ref, s_ref = O.beam_decode(ocfg, P, xt, None, beam=beam, max_len=max_len, temperature=temperature)
        toks, scores = eng.beam_decode(x, None, beam=beam, max_len=max_len, temperature=temperature, use_ema=use_ema)
        s_ref = s_ref.numpy()
        tscale = max(1.0, 0.7 / temperature)       # logit errors enter the scores divided by the temperature
        e_score = float(np.abs(scores - s_ref).max())
        # beams whose score is well separated from their neighbours must hold identical tokens
        sep = np.ones_like(s_ref, bool)
        d = np.abs(np.diff(s_ref, axis=1)) > 10 * margin * tscale
        sep[:, 1:] &= d
        sep[:, :-1] &= d
        record(rec_name, score_abs=e_score, separated_beams=float(sep.mean()),
               beams_identical=float((toks == t_ref.numpy()).all(2).mean()), score_tol=score_tol * tscale)
        assert e_score < score_tol * tscale
        assert sep.any() or not require_separated
        assert (toks[sep] == t_ref.numpy()[sep]).all()
        assert (np.diff(scores, axis=1) <= 1e-6).all(), "beams must be best-first"
    if own:
        eng.close()
