"""CPU oracle for the ECoG->text seq2seq hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the arithmetic of the reference's hot path lives in the un-vendored,
un-pinned third-party package `machine_learning` (jgmakin/machine_learning @ HEAD, on
TensorFlow 1.15.5; /root/reference/README.md:7-23) which is absent here, and the reference
ships no tests, golden vectors or checkpoints.  This file restates the algorithm from the
reference's own call sites and checkpoint-name parser:

  * forward-graph fragment            /root/reference/ecog2txt/trainers.py:773-859
  * variable naming / weight shapes   /root/reference/ecog2txt/trainers.py:444-554
  * conv kernel axes (1,W,C,E)        /root/reference/ecog2txt/plotters.py:508-531
  * final-state axes                  /root/reference/ecog2txt/plotters.py:1388
  * stream manifests (pad/EOS/OOV)    /root/reference/ecog2txt/subjects.py:274-404
  * schedules / scopes                /root/reference/ecog2txt/trainers.py:303-374
  * decode outputs -> strings         /root/reference/ecog2txt/trainers.py:925-963
  * hyper-parameters                  /root/reference/ecog2txt/auxiliary/EFC/mochastar_word_sequence.yaml

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (ecog2txt_b200) never does.

Everything is plain torch on CPU; `dtype` selects fp32 (parity tolerance) or fp64 (to measure
how far fp32 itself is from exact).  Loops are explicit (one LSTM step per python iteration) so
that each line can be read against Appendix D.  The CPU baseline bench.py quotes is oracle/speed_mode.py: the
same model with the recurrences on torch.nn.LSTM (oneDNN), checked against this file in tests/test_golden.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

PAD_ID, EOS_ID, OOV_ID = 0, 1, 2  # /root/reference/ecog2txt/auxiliary/vocab.mocha-timit.1806 lines 1-3


# ------------------------------------------------------------------------------------------------
# configuration + parameter naming (Appendix B)
# ------------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """Model geometry.  Defaults = mochastar_word_sequence.yaml:62-75,84-85,89 (config 2)."""
    subnet_ids: Sequence[int] = (400,)
    subnet_C: Sequence[int] = (256,)          # grid_size 16x16, yaml:57-59
    subnet_W: Sequence[int] = (12,)           # round(200/16.5), subjects.py:144-153
    E: int = 100                              # encoder_embedding
    H: Sequence[int] = (400, 400, 400)        # encoder_rnn (per direction)
    D: int = 150                              # decoder_embedding
    Hd: int = 800                             # decoder_rnn  (== 2*H[-1], the bridge)
    V: int = 1806                             # vocab.mocha-timit.1806
    conv_act: str = "relu"                    # App. D item 3 [CHOICE]
    emb_act: str = "relu"                     # App. D item 6 [CHOICE]
    start_id: int = EOS_ID                    # App. D item 6 [CHOICE]
    pad_id: int = PAD_ID
    eos_id: int = EOS_ID
    penalty_scale: float = 1.0                # subjects.py:289
    lr: float = 5e-4                          # App. D item 8
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    ema_decay: float = 0.99                   # yaml:3
    attention: str = "none"                   # "luong" / "bahdanau": optional A7 module [CHOICE; absent from the reference, SURVEY 0.5]
    # A6 encoder-targets head (App. D item 12): FF head on the outputs of encoder layer `aux_layer`
    # ("encoder_1_targets" -> 1, trainers.py:798-799; yaml:54,68-69,81); aux_layer < 0 = no head (README.md:61)
    aux_layer: int = -1
    aux_hidden: int = 0                       # encoder_1_projection = [225]; 0 = straight to the output layer
    aux_F: int = 0                            # num_features of the encoder targets (13 MFCC-type / 42 phoneme classes)
    aux_kind: str = "gaussian"                # "gaussian": float targets, squared error; "categorical": int targets, CE
    aux_penalty: float = 1.0                  # encoder_1_targets_penalty_scale, yaml:54
    # layer_sizes['decoder_projection'] (yaml:65, empty in every shipped manifest): one optional hidden FF layer between the
    # decoder state and the vocabulary projection -- relu + FF dropout like the other FF layers [CHOICE]
    proj_hidden: int = 0

    def __post_init__(self):
        assert self.Hd == 2 * self.H[-1], "bridge needs decoder_rnn == 2*encoder_rnn[-1]"


def param_shapes(cfg: OracleConfig) -> "Dict[str, Tuple[int, ...]]":
    """Canonical (TF-checkpoint) names and shapes, in flat-buffer order:
    subject-private tensors first (seq2seq/subnet_<id>/...), then the shared ones
    (trainers.py:337-338,470-478)."""
    shapes: Dict[str, Tuple[int, ...]] = {}
    for sid, C, W in zip(cfg.subnet_ids, cfg.subnet_C, cfg.subnet_W):
        base = f"seq2seq/subnet_{sid}/encoder_embedding_{C}_{cfg.E}_0"
        shapes[base + "/weights"] = (1, W, C, cfg.E)       # plotters.py:511-514
        shapes[base + "/biases"] = (cfg.E,)
    n_in = cfg.E
    for l, H in enumerate(cfg.H):
        for d in ("fw", "bw"):
            base = f"seq2seq/encoder_rnn_{l}/bidirectional_rnn/{d}/multi_rnn_cell/cell_0/lstm_cell"
            shapes[base + "/kernel"] = (n_in + H, 4 * H)   # trainers.py:527-529
            shapes[base + "/bias"] = (4 * H,)
        n_in = 2 * H
    base = f"seq2seq/decoder_embedding_{cfg.V}_{cfg.D}_0"
    shapes[base + "/weights"] = (cfg.V, cfg.D)
    shapes[base + "/biases"] = (cfg.D,)
    base = "seq2seq/decoder_rnn/multi_rnn_cell/cell_0/lstm_cell"
    shapes[base + "/kernel"] = (cfg.D + cfg.Hd, 4 * cfg.Hd)
    shapes[base + "/bias"] = (4 * cfg.Hd,)
    hid, out = proj_names(cfg)
    if hid is not None:                                    # '<x>_projection' scopes number their layers, trainers.py:488-520
        shapes[hid + "/weights"] = (cfg.Hd, cfg.proj_hidden)
        shapes[hid + "/biases"] = (cfg.proj_hidden,)
    shapes[out + "/weights"] = (cfg.V, cfg.proj_hidden or cfg.Hd)   # final layer: transposed, trainers.py:513-520
    shapes[out + "/biases"] = (cfg.V,)
    if cfg.attention in ("luong", "bahdanau"):
        # stored [out, in] like the projection; q = h Wq^T, h~ = tanh(Wc [ctx; h] + bc)
        shapes["seq2seq/decoder_attention/query/weights"] = (cfg.Hd, cfg.Hd)
        if cfg.attention == "bahdanau":
            shapes["seq2seq/decoder_attention/keys/weights"] = (cfg.Hd, cfg.Hd)
            shapes["seq2seq/decoder_attention/score/weights"] = (1, cfg.Hd)
        shapes["seq2seq/decoder_attention/combine/weights"] = (cfg.Hd, 2 * cfg.Hd)
        shapes["seq2seq/decoder_attention/combine/biases"] = (cfg.Hd,)
    if cfg.aux_layer >= 0:
        n_in = 2 * cfg.H[cfg.aux_layer]
        k = 0
        if cfg.aux_hidden > 0:
            base = f"seq2seq/encoder_{cfg.aux_layer}_projection_{n_in}_{cfg.aux_hidden}_0"
            shapes[base + "/weights"] = (n_in, cfg.aux_hidden)
            shapes[base + "/biases"] = (cfg.aux_hidden,)
            n_in, k = cfg.aux_hidden, 1
        base = f"seq2seq/encoder_{cfg.aux_layer}_projection_{n_in}_{cfg.aux_F}_{k}"
        shapes[base + "/weights"] = (cfg.aux_F, n_in)      # final layer of a *_projection: transposed, trainers.py:513-520
        shapes[base + "/biases"] = (cfg.aux_F,)
    return shapes


def proj_names(cfg: OracleConfig):
    """(hidden-layer base name or None, output-layer base name) of the decoder projection."""
    if cfg.proj_hidden > 0:
        return (f"seq2seq/decoder_projection_{cfg.Hd}_{cfg.proj_hidden}_0",
                f"seq2seq/decoder_projection_{cfg.proj_hidden}_{cfg.V}_1")
    return None, f"seq2seq/decoder_projection_{cfg.Hd}_{cfg.V}_0"


def aux_names(cfg: OracleConfig):
    """(hidden-layer base name or None, output-layer base name) of the encoder-targets head."""
    n_in = 2 * cfg.H[cfg.aux_layer]
    if cfg.aux_hidden > 0:
        return (f"seq2seq/encoder_{cfg.aux_layer}_projection_{n_in}_{cfg.aux_hidden}_0",
                f"seq2seq/encoder_{cfg.aux_layer}_projection_{cfg.aux_hidden}_{cfg.aux_F}_1")
    return None, f"seq2seq/encoder_{cfg.aux_layer}_projection_{n_in}_{cfg.aux_F}_0"


def init_params(cfg: OracleConfig, seed: int = 1, dtype=torch.float32) -> "Dict[str, torch.Tensor]":
    """Glorot-uniform matrices, zero biases [CHOICE]; numpy RandomState so that the product's
    host-side initialiser (which must not import this file) can reproduce it bit-for-bit."""
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        if len(shape) == 1:
            a = np.zeros(shape, np.float32)
        else:
            if len(shape) == 4:
                fan_in, fan_out = shape[1] * shape[2], shape[3]
            elif name.endswith("decoder_projection_%d_%d_0/weights" % (cfg.Hd, cfg.V)) or (
                    cfg.aux_layer >= 0 and name == aux_names(cfg)[1] + "/weights"):
                fan_in, fan_out = shape[1], shape[0]
            else:
                fan_in, fan_out = shape[0], shape[1]
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            a = rs.uniform(-lim, lim, size=shape).astype(np.float32)
        out[name] = torch.from_numpy(a).to(dtype)
    return out


# ------------------------------------------------------------------------------------------------
# dropout masks: counter-based hash shared bit-for-bit with the CUDA kernels
# ------------------------------------------------------------------------------------------------
def _mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def dropout_keep(seed: int, stream: int, n: int, p: float) -> np.ndarray:
    """keep[i] = hash(seed, stream, i) >= p*2^32 ; same integer recipe as csrc (e2t_keep)."""
    if p <= 0.0:
        return np.ones(n, np.bool_)
    assert n < 2 ** 32
    idx = np.arange(n, dtype=np.uint64)
    s0 = (int(seed) * 0x9E3779B1 + int(stream) * 0x85EBCA77 + 0x165667B1) & 0xFFFFFFFF
    s = _mix32(np.full(1, s0, np.uint64))[0]
    h = _mix32(idx ^ s)
    h = _mix32((h + np.uint64(0x27D4EB2F)) & np.uint64(0xFFFFFFFF))
    p32 = float(np.float32(p))                      # csrc holds p as fp32
    thresh = np.uint64(min(int(p32 * 4294967296.0), 0xFFFFFFFF))
    return h >= thresh


# stream ids (must match csrc): conv output 0 ; encoder layer l output 1+l ; decoder embedding 64
STREAM_CONV, STREAM_ENC0, STREAM_DEMB, STREAM_AUX, STREAM_PROJ = 0, 1, 64, 96, 112


# ------------------------------------------------------------------------------------------------
# forward pieces
# ------------------------------------------------------------------------------------------------
def infer_lengths(x: torch.Tensor) -> torch.Tensor:
    """App. D item 1: len_b = 1 + index of last frame with any non-zero feature
    (nn.sequences_tools(tfh.hide_shape(x)), trainers.py:806-807)."""
    nz = (x != 0).any(dim=2)                                  # [B,T]
    T = x.shape[1]
    idx = torch.arange(1, T + 1).unsqueeze(0) * nz
    return idx.max(dim=1).values.to(torch.int64)


def reverse_within_length(x: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
    """tf.reverse_sequence(x, lens, seq_axis=1, batch_axis=0), trainers.py:808-810."""
    out = x.clone()
    for b in range(x.shape[0]):
        n = int(lens[b])
        out[b, :n] = x[b, :n].flip(0)
    return out


def _act(z, kind):
    if kind == "relu":
        return torch.relu(z)
    assert kind == "linear"
    return z


def temporal_conv(xr, lens, w4, bias, act, relu_mask=None, want_preact=False):
    """App. D item 3: kernel (1,W,C,E), stride W, zero-pad tail to a multiple of W.
    Returns y [B,T',E] and lens' = ceil(len/W) (trainers.py:813-818,534-541).
    relu_mask [B,T',E] (tests only): impose this activation pattern instead of (z > 0).  A ReLU is discontinuous in its
    derivative: ONE pre-activation within rounding distance of zero that lands on the other side changes the whole
    weight gradient by a few per cent at config 2 (measured: the fp32 and the fp64 oracle differ by 3.8e-2 of the
    largest entry through a single flip among 217 600), so gradient parity is stated for a common pattern, and the
    patterns themselves are compared separately (they may differ only where |z| is within rounding distance of 0)."""
    B, T, C = xr.shape
    _, W, C2, E = w4.shape
    assert C == C2
    T2 = -(-T // W)
    xp = torch.zeros(B, T2 * W, C, dtype=xr.dtype)
    xp[:, :T] = xr
    a = xp.reshape(B, T2, W * C)
    z = a @ w4.reshape(W * C, E) + bias
    y = z * relu_mask.to(z.dtype) if (relu_mask is not None and act == "relu") else _act(z, act)
    lens2 = (lens + W - 1) // W
    return (y, lens2, z) if want_preact else (y, lens2)


def lstm_cell(z, c_prev):
    """TF1 LSTMCell: gate order i, j, f, o; forget_bias = 1 (App. D item 4)."""
    H = c_prev.shape[-1]
    i, j, f, o = z[..., :H], z[..., H:2 * H], z[..., 2 * H:3 * H], z[..., 3 * H:]
    c = torch.sigmoid(f + 1.0) * c_prev + torch.sigmoid(i) * torch.tanh(j)
    h = torch.sigmoid(o) * torch.tanh(c)
    return h, c


def lstm_direction(x, lens, K, bias, reverse):
    """One direction of tf.nn.(bidirectional_)dynamic_rnn: x [B,T,In]; past len the state is
    frozen and the output is zero; the backward direction runs on the sequence reversed
    within its own length and its outputs are reversed back (== masked scan from the tail)."""
    B, T, n_in = x.shape
    H = K.shape[1] // 4
    Wx, Wh = K[:n_in], K[n_in:]
    h = torch.zeros(B, H, dtype=x.dtype)
    c = torch.zeros(B, H, dtype=x.dtype)
    out = torch.zeros(B, T, H, dtype=x.dtype)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        valid = (t < lens).to(x.dtype).unsqueeze(1)
        z = x[:, t] @ Wx + h @ Wh + bias
        hn, cn = lstm_cell(z, c)
        h = valid * hn + (1 - valid) * h
        c = valid * cn + (1 - valid) * c
        out[:, t] = valid * hn
    return out, h, c


def encoder(cfg, P, x, lens, subnet, train_masks=None, conv_relu_mask=None):
    """A2-A5: lengths -> reverse -> conv -> stacked BiLSTM.  Returns dict of activations."""
    sid, C, W = cfg.subnet_ids[subnet], cfg.subnet_C[subnet], cfg.subnet_W[subnet]
    if lens is None:
        lens = infer_lengths(x)
    xr = reverse_within_length(x, lens)
    base = f"seq2seq/subnet_{sid}/encoder_embedding_{C}_{cfg.E}_0"
    y, lens2, z = temporal_conv(xr, lens, P[base + "/weights"], P[base + "/biases"], cfg.conv_act, conv_relu_mask, True)
    acts = {"lens": lens, "lens2": lens2, "conv_out_nodrop": y, "conv_preact": z}
    if train_masks is not None and "conv" in train_masks:
        y = y * train_masks["conv"]
    acts["conv_out"] = y
    inp = y
    for l, H in enumerate(cfg.H):
        outs, hs, cs = [], [], []
        for d, rev in (("fw", False), ("bw", True)):
            b = f"seq2seq/encoder_rnn_{l}/bidirectional_rnn/{d}/multi_rnn_cell/cell_0/lstm_cell"
            o, h, c = lstm_direction(inp, lens2, P[b + "/kernel"], P[b + "/bias"], rev)
            outs.append(o), hs.append(h), cs.append(c)
        out = torch.cat(outs, dim=2)
        acts[f"enc{l}_out"] = out
        if train_masks is not None and f"enc{l}" in train_masks:
            out = out * train_masks[f"enc{l}"]
        inp = out
    acts["final_h"] = torch.cat(hs, dim=1)                     # App. D item 5 (bridge)
    acts["final_c"] = torch.cat(cs, dim=1)
    return acts


def luong_attention(cfg, P, h, enc, lens2):
    """A7 [CHOICE, not in the reference]: Luong 'general' attention on the decoder OUTPUT (no input feeding, so the
    recurrence is untouched): q = h Wq^T; score_s = q . enc_s for s < lens2 (else masked); alpha = softmax(score);
    ctx = sum_s alpha_s enc_s; h~ = tanh(Wc [ctx; h] + bc).  enc [B,T',Hd] = outputs of the last encoder layer."""
    q = h @ P["seq2seq/decoder_attention/query/weights"].T
    score = torch.einsum("bf,bsf->bs", q, enc)
    mask = torch.arange(enc.shape[1]).unsqueeze(0) < lens2.unsqueeze(1)
    score = score.masked_fill(~mask, -1e30)
    alpha = torch.softmax(score, dim=1) * mask.to(h.dtype)
    ctx = torch.einsum("bs,bsf->bf", alpha, enc)
    return torch.tanh(torch.cat([ctx, h], dim=1) @ P["seq2seq/decoder_attention/combine/weights"].T
                      + P["seq2seq/decoder_attention/combine/biases"])


def bahdanau_attention(cfg, P, h, enc, lens2):
    """A7, additive variant [CHOICE, not in the reference]: score_s = v . tanh(Wq h + Wk enc_s) for s < lens2; the
    softmax, the context and the combine layer h~ = tanh(Wc [ctx; h] + bc) are those of `luong_attention` (attention on the
    decoder OUTPUT, no input feeding, so the recurrence is untouched)."""
    q = h @ P["seq2seq/decoder_attention/query/weights"].T                     # [B,A]
    kp = enc @ P["seq2seq/decoder_attention/keys/weights"].T                   # [B,T',A]
    v = P["seq2seq/decoder_attention/score/weights"][0]
    score = torch.tanh(q.unsqueeze(1) + kp) @ v                                # [B,T']
    mask = torch.arange(enc.shape[1]).unsqueeze(0) < lens2.unsqueeze(1)
    score = score.masked_fill(~mask, -1e30)
    alpha = torch.softmax(score, dim=1) * mask.to(h.dtype)
    ctx = torch.einsum("bs,bsf->bf", alpha, enc)
    return torch.tanh(torch.cat([ctx, h], dim=1) @ P["seq2seq/decoder_attention/combine/weights"].T
                      + P["seq2seq/decoder_attention/combine/biases"])


def decoder_step(cfg, P, y_prev, h, c, emb_mask=None, enc=None, lens2=None, proj_mask=None):
    """One decoder step: Emb[y_prev] (+bias, act) -> LSTM(Hd) [-> attention] [-> relu(. W1 + b1), FF dropout] ->
    logits = . Wp^T + b."""
    eb = f"seq2seq/decoder_embedding_{cfg.V}_{cfg.D}_0"
    e = _act(P[eb + "/weights"][y_prev] + P[eb + "/biases"], cfg.emb_act)
    if emb_mask is not None:
        e = e * emb_mask
    rb = "seq2seq/decoder_rnn/multi_rnn_cell/cell_0/lstm_cell"
    K, bias = P[rb + "/kernel"], P[rb + "/bias"]
    z = e @ K[:cfg.D] + h @ K[cfg.D:] + bias
    h, c = lstm_cell(z, c)
    hid, pb = proj_names(cfg)
    ho = h
    if cfg.attention == "luong":
        ho = luong_attention(cfg, P, h, enc, lens2)
    elif cfg.attention == "bahdanau":
        ho = bahdanau_attention(cfg, P, h, enc, lens2)
    if hid is not None:
        ho = torch.relu(ho @ P[hid + "/weights"] + P[hid + "/biases"])
        if proj_mask is not None:
            ho = ho * proj_mask
    logits = ho @ P[pb + "/weights"].T + P[pb + "/biases"]
    return logits, h, c


def make_masks(cfg, seed, B, T2, L, ff_p, rnn_p, dtype):
    """Dropout masks laid out exactly like the CUDA buffers (time-major [T',B,F])."""
    masks = {}
    def mk(stream, shape, p):
        keep = dropout_keep(seed, stream, int(np.prod(shape)), p).reshape(shape)
        return torch.from_numpy(keep.astype(np.float64) / (1.0 - p)).to(dtype)
    if ff_p > 0:
        masks["conv"] = mk(STREAM_CONV, (T2, B, cfg.E), ff_p).permute(1, 0, 2)
        masks["demb"] = mk(STREAM_DEMB, (L, B, cfg.D), ff_p).permute(1, 0, 2)
        if cfg.proj_hidden > 0:
            masks["proj"] = mk(STREAM_PROJ, (L, B, cfg.proj_hidden), ff_p).permute(1, 0, 2)
        if cfg.aux_layer >= 0 and cfg.aux_hidden > 0:
            masks["aux"] = mk(STREAM_AUX, (T2, B, cfg.aux_hidden), ff_p).permute(1, 0, 2)
    if rnn_p > 0:
        for l, H in enumerate(cfg.H[:-1]):
            masks[f"enc{l}"] = mk(STREAM_ENC0 + l, (T2, B, 2 * H), rnn_p).permute(1, 0, 2)
    return masks


def prepare_encoder_targets(tgt: torch.Tensor, lens: torch.Tensor, W: int) -> torch.Tensor:
    """A6 targets (trainers.py:791-795): reverse within length, then keep every W-th frame, `[:, 0::W]`.
    tgt [B,T,F] float or [B,T] int at the input's frame rate -> [B,T',...].  The reference infers the
    length from the targets' own zero padding; here the utterance length is used (same trial, same clock)."""
    r = reverse_within_length(tgt, lens)
    return r[:, 0::W]


def aux_head(cfg, P, acts, aux_targets, subnet, masks=None):
    """A6 (App. D item 12): FF head on the (un-dropped) outputs of encoder layer `aux_layer`:
    [hidden = act(x W1 + b1), FF dropout] -> out = hidden W2^T + b2 (W2 stored transposed, trainers.py:513-520).
    gaussian: 0.5 * sum_f (out - tgt)^2 over frames t' < len'; categorical: CE vs the class index, frames whose
    target is the pad index 0 or t' >= len' are masked.  Returns (aux_penalty * sum, number of unmasked frames)."""
    W = cfg.subnet_W[subnet]
    x = acts[f"enc{cfg.aux_layer}_out"]                          # [B,T',2H]
    hb, ob = aux_names(cfg)
    if hb is not None:
        x = _act(x @ P[hb + "/weights"] + P[hb + "/biases"], cfg.conv_act)
        if masks is not None and "aux" in masks:
            x = x * masks["aux"]
    out = x @ P[ob + "/weights"].T + P[ob + "/biases"]           # [B,T',F]
    tg = prepare_encoder_targets(aux_targets, acts["lens"], W)
    T2 = out.shape[1]
    valid = torch.arange(T2).unsqueeze(0) < acts["lens2"].unsqueeze(1)
    acts["aux_out"] = out
    if cfg.aux_kind == "gaussian":
        m = valid.to(out.dtype).unsqueeze(2)
        loss = 0.5 * (((out - tg.to(out.dtype)) ** 2) * m).sum()
    else:
        tg = tg.long()
        valid = valid & (tg != 0)
        lp = torch.log_softmax(out, dim=2)
        loss = -(lp.gather(2, tg.unsqueeze(2)).squeeze(2) * valid.to(out.dtype)).sum()
    return loss * cfg.aux_penalty, int(valid.sum())


def train_loss(cfg, P, x, lens, y, subnet=0, masks=None, aux_targets=None, conv_relu_mask=None):
    """Teacher-forced forward + masked CE (App. D items 6-7).  y [B,L] int64 with EOS appended and
    pad_id after it.  Returns (sum of token losses * penalty_scale [+ aux_penalty * sum of the
    encoder-targets losses], n_unmasked_tokens, acts); the caller divides by the *global* token count
    (one common normaliser for both penalties [CHOICE], so that data-parallel training needs one scalar)."""
    acts = encoder(cfg, P, x, lens, subnet, masks, conv_relu_mask)
    h, c = acts["final_h"], acts["final_c"]
    B, L = y.shape
    prev = torch.full((B,), cfg.start_id, dtype=torch.int64)
    loss = torch.zeros((), dtype=x.dtype)
    logits_all = []
    for k in range(L):
        em = None if masks is None or "demb" not in masks else masks["demb"][:, k]
        pm = None if masks is None or "proj" not in masks else masks["proj"][:, k]
        logits, h, c = decoder_step(cfg, P, prev, h, c, em, acts[f"enc{len(cfg.H) - 1}_out"], acts["lens2"], pm)
        logits_all.append(logits)
        lp = torch.log_softmax(logits, dim=1)
        m = (y[:, k] != cfg.pad_id).to(x.dtype)
        loss = loss - (lp.gather(1, y[:, k:k + 1]).squeeze(1) * m).sum()
        prev = y[:, k]
    acts["logits"] = torch.stack(logits_all, dim=1)
    ntok = int((y != cfg.pad_id).sum())
    loss = loss * cfg.penalty_scale
    acts["decoder_loss"] = float(loss.detach())
    if cfg.aux_layer >= 0 and aux_targets is not None:
        la, nf = aux_head(cfg, P, acts, aux_targets, subnet, masks)
        acts["aux_loss"], acts["aux_frames"] = float(la.detach()), nf
        loss = loss + la
    return loss, ntok, acts


def input_gradients(cfg, P, x, lens, y, subnet=0, aux_targets=None):
    """A13 saliency (trainers.py:703-732): d(loss)/d(encoder_inputs) [B,T,C] with dropout off; the caller picks the
    penalty under study by setting the others to 0 in `cfg` (get_saliencies zeroes every *_targets penalty but one)."""
    xg = x.clone().requires_grad_(True)
    if lens is None:
        lens = infer_lengths(x)
    loss, _, _ = train_loss(cfg, P, xg, lens, y, subnet, None, aux_targets)
    loss.backward()
    return xg.grad


def loss_and_grads(cfg, P, x, lens, y, subnet=0, masks=None, aux_targets=None, conv_relu_mask=None):
    """Reference gradients by autograd of `train_loss` (sum, not yet / ntok)."""
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    loss, ntok, acts = train_loss(cfg, Pg, x, lens, y, subnet, masks, aux_targets, conv_relu_mask)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Pg.items()}
    return float(loss.detach()), ntok, grads, {k: (v.detach() if torch.is_tensor(v) else v) for k, v in acts.items()}


# ------------------------------------------------------------------------------------------------
# optimiser (App. D item 8): TF1 AdamOptimizer + ExponentialMovingAverage
# ------------------------------------------------------------------------------------------------
class AdamEMA:
    def __init__(self, cfg, P):
        self.cfg, self.t = cfg, 0
        self.m = {k: torch.zeros_like(v) for k, v in P.items()}
        self.v = {k: torch.zeros_like(v) for k, v in P.items()}
        self.ema = {k: v.clone() for k, v in P.items()}       # shadow starts at the initial value

    def step(self, P, grads, grad_scale, trainable=None):
        cfg = self.cfg
        self.t += 1
        lr_t = cfg.lr * math.sqrt(1 - cfg.beta2 ** self.t) / (1 - cfg.beta1 ** self.t)
        for k in P:
            if trainable is not None and not trainable(k):
                continue
            g = grads[k] * grad_scale
            self.m[k] = cfg.beta1 * self.m[k] + (1 - cfg.beta1) * g
            self.v[k] = cfg.beta2 * self.v[k] + (1 - cfg.beta2) * g * g
            P[k] = P[k] - lr_t * self.m[k] / (self.v[k].sqrt() + cfg.eps)
            self.ema[k] = cfg.ema_decay * self.ema[k] + (1 - cfg.ema_decay) * P[k]


# ------------------------------------------------------------------------------------------------
# decoding (App. D item 9)
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def greedy_decode(cfg, P, x, lens, max_len=20, subnet=0, temperature=1.0):
    """argmax decode; stop at EOS (EOS kept, tail = pad).  Returns tokens [B,max_len] int64,
    per-step log-prob of the emitted token under softmax(logits/temperature) [B,max_len]
    (0 where padded) and the raw logits [B,max_len,V] (for margin-aware comparisons)."""
    acts = encoder(cfg, P, x, lens, subnet)
    h, c = acts["final_h"], acts["final_c"]
    B = x.shape[0]
    prev = torch.full((B,), cfg.start_id, dtype=torch.int64)
    done = torch.zeros(B, dtype=torch.bool)
    toks = torch.full((B, max_len), cfg.pad_id, dtype=torch.int64)
    logp = torch.zeros(B, max_len, dtype=x.dtype)
    all_logits = torch.zeros(B, max_len, cfg.V, dtype=x.dtype)
    for k in range(max_len):
        logits, h, c = decoder_step(cfg, P, prev, h, c, None, acts[f"enc{len(cfg.H) - 1}_out"], acts["lens2"])
        all_logits[:, k] = logits
        nxt = logits.argmax(dim=1)
        lp = torch.log_softmax(logits / temperature, dim=1).gather(1, nxt[:, None]).squeeze(1)
        toks[:, k] = torch.where(done, torch.full_like(nxt, cfg.pad_id), nxt)
        logp[:, k] = torch.where(done, torch.zeros_like(lp), lp)
        prev = torch.where(done, prev, nxt)
        done = done | (nxt == cfg.eos_id)
    return toks, logp, all_logits


@torch.no_grad()
def beam_decode(cfg, P, x, lens, beam=8, max_len=20, subnet=0, temperature=1.0):
    """Beam search, additive log-softmax(logits/temperature), no length normalisation; finished
    beams are frozen (only a pad continuation with score +0).  Returns tokens [B,beam,max_len]
    best-first (trainers.py:952-963) and scores [B,beam]."""
    acts = encoder(cfg, P, x, lens, subnet)
    B, V = x.shape[0], cfg.V
    NEG = -1e30
    toks = torch.full((B, beam, max_len), cfg.pad_id, dtype=torch.int64)
    scores = torch.full((B, beam), NEG, dtype=x.dtype)
    scores[:, 0] = 0.0
    h = acts["final_h"][:, None].repeat(1, beam, 1)
    c = acts["final_c"][:, None].repeat(1, beam, 1)
    prev = torch.full((B, beam), cfg.start_id, dtype=torch.int64)
    done = torch.zeros(B, beam, dtype=torch.bool)
    for k in range(max_len):
        enc_top = acts[f"enc{len(cfg.H) - 1}_out"]
        logits, hn, cn = decoder_step(cfg, P, prev.reshape(-1), h.reshape(B * beam, -1), c.reshape(B * beam, -1), None,
                                      enc_top.repeat_interleave(beam, dim=0), acts["lens2"].repeat_interleave(beam))
        lp = torch.log_softmax(logits / temperature, dim=1).reshape(B, beam, V)
        hn, cn = hn.reshape(B, beam, -1), cn.reshape(B, beam, -1)
        cand = scores[:, :, None] + lp
        frozen = torch.full((B, beam, V), NEG, dtype=x.dtype)
        frozen[:, :, cfg.pad_id] = scores
        cand = torch.where(done[:, :, None], frozen, cand)
        # stable best-first: sort by (-score, flat index) so ties resolve to the lowest index
        flat = cand.reshape(B, beam * V)
        order = torch.argsort(-flat, dim=1, stable=True)[:, :beam]
        top = flat.gather(1, order)
        src, tok = order // V, order % V
        toks = toks.gather(1, src[:, :, None].expand(-1, -1, max_len)).clone()
        was_done = done.gather(1, src)
        toks[:, :, k] = torch.where(was_done, torch.full_like(tok, cfg.pad_id), tok)
        h = torch.where(was_done[:, :, None], h.gather(1, src[:, :, None].expand(-1, -1, h.shape[2])),
                        hn.gather(1, src[:, :, None].expand(-1, -1, hn.shape[2])))
        c = torch.where(was_done[:, :, None], c.gather(1, src[:, :, None].expand(-1, -1, c.shape[2])),
                        cn.gather(1, src[:, :, None].expand(-1, -1, cn.shape[2])))
        prev = torch.where(was_done, prev.gather(1, src), tok)
        done = was_done | (tok == cfg.eos_id)
        scores = top
    return toks, scores


# ------------------------------------------------------------------------------------------------
# strings + metrics (App. D item 10)
# ------------------------------------------------------------------------------------------------
def inds_to_sentence(inds, tokens_list, pad_token="<pad>", eos_token="<EOS>"):
    """trainers.py:957-962."""
    return "".join(tokens_list[i] for i in inds).replace("_", " ").replace(pad_token, "").replace(
        eos_token, "").rstrip()


def word_error_rate(ref_words: List[str], hyp_words: List[str]) -> float:
    """word-level Levenshtein / reference length (utils_jgm.toolbox.wer_vector; subjects.py:546-549)."""
    n, m = len(ref_words), len(hyp_words)
    d = list(range(m + 1))
    for i in range(1, n + 1):
        prev, d[0] = d[0], i
        for j in range(1, m + 1):
            cur = d[j]
            d[j] = min(d[j] + 1, d[j - 1] + 1, prev + (ref_words[i - 1] != hyp_words[j - 1]))
            prev = cur
    return d[m] / max(n, 1)
