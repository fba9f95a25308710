"""CPU oracle, SPEED MODE: the same model as seq2seq_oracle.py with every recurrence handed to torch.nn.LSTM (oneDNN
fused cell on CPU) and packed sequences for the ragged lengths.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

PARITY UNPINNED like the module it mirrors (the reference's arithmetic lives in the un-vendored `machine_learning`
package on TF1.15, /root/reference/README.md:7-23).  What this file is for: BASELINE.md section 3 asks for the CPU number
of "the oracle in speed mode ... oneDNN LSTM allowed" -- the fastest honest CPU restatement of
/root/reference/ecog2txt/trainers.py:773-859 (reverse -> strided conv -> stacked BiLSTM -> LSTM decoder -> masked CE) this
image can run, so that the GPU/CPU ratio is not quoted against a Python-loop port.  tests/test_golden.py checks it against
the explicit-loop oracle (loss and every gradient, fp32 tolerance) before bench.py is allowed to time it.

Only tests/ and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

from . import seq2seq_oracle as O


def _tf_to_torch_gates(m: torch.Tensor, H: int) -> torch.Tensor:
    """TF1 LSTMCell packs (i, j, f, o) along the last axis (trainers.py:527-529); torch.nn.LSTM wants (i, f, g, o) rows."""
    i, j, f, o = m[..., :H], m[..., H:2 * H], m[..., 2 * H:3 * H], m[..., 3 * H:]
    return torch.cat([i, f, j, o], dim=-1)


def _load_lstm(lstm: nn.LSTM, layer_suffix: str, K: torch.Tensor, b: torch.Tensor, n_in: int, H: int):
    Kt = _tf_to_torch_gates(K, H)
    bt = _tf_to_torch_gates(b.clone(), H)
    bt[H:2 * H] += 1.0                                          # forget_bias = 1 folded into the bias (App. D item 4)
    with torch.no_grad():
        getattr(lstm, "weight_ih_" + layer_suffix).copy_(Kt[:n_in].T)
        getattr(lstm, "weight_hh_" + layer_suffix).copy_(Kt[n_in:].T)
        getattr(lstm, "bias_ih_" + layer_suffix).copy_(bt)
        getattr(lstm, "bias_hh_" + layer_suffix).zero_()
    getattr(lstm, "bias_hh_" + layer_suffix).requires_grad_(False)      # TF has ONE bias per cell


class SpeedModel(nn.Module):
    """Parameters live in torch.nn.LSTM layout (converted once from the canonical TF-named dict)."""

    def __init__(self, cfg: O.OracleConfig, P: Dict[str, torch.Tensor], subnet: int = 0):
        super().__init__()
        assert cfg.attention == "none" and cfg.aux_layer < 0 and cfg.proj_hidden == 0, \
            "speed mode covers the reference model (no optional rows)"
        self.cfg, self.subnet = cfg, subnet
        sid, C, W = cfg.subnet_ids[subnet], cfg.subnet_C[subnet], cfg.subnet_W[subnet]
        self.W, self.C = W, C
        base = f"seq2seq/subnet_{sid}/encoder_embedding_{C}_{cfg.E}_0"
        self.conv_w = nn.Parameter(P[base + "/weights"].reshape(W * C, cfg.E).clone())
        self.conv_b = nn.Parameter(P[base + "/biases"].clone())
        self.enc = nn.ModuleList()
        n_in = cfg.E
        for l, H in enumerate(cfg.H):
            lstm = nn.LSTM(n_in, H, batch_first=True, bidirectional=True)
            for d, suf in (("fw", "l0"), ("bw", "l0_reverse")):
                b = f"seq2seq/encoder_rnn_{l}/bidirectional_rnn/{d}/multi_rnn_cell/cell_0/lstm_cell"
                _load_lstm(lstm, suf, P[b + "/kernel"], P[b + "/bias"], n_in, H)
            self.enc.append(lstm)
            n_in = 2 * H
        eb = f"seq2seq/decoder_embedding_{cfg.V}_{cfg.D}_0"
        self.emb_w = nn.Parameter(P[eb + "/weights"].clone())
        self.emb_b = nn.Parameter(P[eb + "/biases"].clone())
        rb = "seq2seq/decoder_rnn/multi_rnn_cell/cell_0/lstm_cell"
        self.dec = nn.LSTM(cfg.D, cfg.Hd, batch_first=True)
        _load_lstm(self.dec, "l0", P[rb + "/kernel"], P[rb + "/bias"], cfg.D, cfg.Hd)
        pb = f"seq2seq/decoder_projection_{cfg.Hd}_{cfg.V}_0"
        self.proj_w = nn.Parameter(P[pb + "/weights"].clone())
        self.proj_b = nn.Parameter(P[pb + "/biases"].clone())

    def forward(self, x, y, ff_p: float = 0.0, rnn_p: float = 0.0):
        """x [B,T,C] zero padded, y [B,L] int64 -> (summed masked CE * penalty_scale, token count)."""
        cfg = self.cfg
        lens = O.infer_lengths(x)
        B, T, C = x.shape
        # reverse within length (trainers.py:808-810) as one gather
        t = torch.arange(T).unsqueeze(0)
        src = torch.where(t < lens.unsqueeze(1), lens.unsqueeze(1) - 1 - t, t)
        xr = torch.gather(x, 1, src.unsqueeze(2).expand(B, T, C))
        T2 = -(-T // self.W)
        if T2 * self.W != T:
            xr = F.pad(xr, (0, 0, 0, T2 * self.W - T))
        h = O._act(xr.reshape(B, T2, self.W * C) @ self.conv_w + self.conv_b, cfg.conv_act)
        h = F.dropout(h, ff_p, self.training)
        lens2 = ((lens + self.W - 1) // self.W).clamp(min=1)    # an empty utterance still needs one packed step
        empty = (lens == 0)
        hn = cn = None
        full = bool((lens2 == T2).all())          # fixed-length batches (the benchmark regime): dense path, oneDNN-fused
        for lstm in self.enc:
            if full:
                h, (hn, cn) = lstm(h)
            else:
                packed = pack_padded_sequence(h, lens2.cpu(), batch_first=True, enforce_sorted=False)
                out, (hn, cn) = lstm(packed)
                h, _ = pad_packed_sequence(out, batch_first=True, total_length=T2)
            h = F.dropout(h, rnn_p, self.training)
        h0 = torch.cat([hn[0], hn[1]], dim=1)
        c0 = torch.cat([cn[0], cn[1]], dim=1)
        if empty.any():
            keep = (~empty).to(x.dtype).unsqueeze(1)
            h0, c0 = h0 * keep, c0 * keep
        prev = torch.cat([torch.full((B, 1), cfg.start_id, dtype=torch.int64), y[:, :-1]], dim=1)
        e = O._act(self.emb_w[prev] + self.emb_b, cfg.emb_act)
        e = F.dropout(e, ff_p, self.training)
        hd, _ = self.dec(e, (h0.unsqueeze(0).contiguous(), c0.unsqueeze(0).contiguous()))
        logits = hd @ self.proj_w.T + self.proj_b
        loss = F.cross_entropy(logits.reshape(-1, cfg.V), y.reshape(-1), ignore_index=cfg.pad_id, reduction="sum")
        return loss * cfg.penalty_scale, int((y != cfg.pad_id).sum())

    def canonical_grads(self) -> Dict[str, torch.Tensor]:
        """Gradients under the TF names / layouts of seq2seq_oracle.param_shapes (for the parity test)."""
        cfg = self.cfg
        sid, C, W = cfg.subnet_ids[self.subnet], self.C, self.W

        def lstm_grads(lstm, suf, n_in, H):
            gk = torch.cat([getattr(lstm, "weight_ih_" + suf).grad.T, getattr(lstm, "weight_hh_" + suf).grad.T], dim=0)
            gb = getattr(lstm, "bias_ih_" + suf).grad          # bias_hh receives the same gradient; TF has one bias
            inv = lambda m: torch.cat([m[..., :H], m[..., 2 * H:3 * H], m[..., H:2 * H], m[..., 3 * H:]], dim=-1)  # noqa: E731
            return inv(gk), inv(gb)
        out = {}
        base = f"seq2seq/subnet_{sid}/encoder_embedding_{C}_{cfg.E}_0"
        out[base + "/weights"] = self.conv_w.grad.reshape(1, W, C, cfg.E)
        out[base + "/biases"] = self.conv_b.grad
        n_in = cfg.E
        for l, H in enumerate(cfg.H):
            for d, suf in (("fw", "l0"), ("bw", "l0_reverse")):
                b = f"seq2seq/encoder_rnn_{l}/bidirectional_rnn/{d}/multi_rnn_cell/cell_0/lstm_cell"
                out[b + "/kernel"], out[b + "/bias"] = lstm_grads(self.enc[l], suf, n_in, H)
            n_in = 2 * H
        eb = f"seq2seq/decoder_embedding_{cfg.V}_{cfg.D}_0"
        out[eb + "/weights"], out[eb + "/biases"] = self.emb_w.grad, self.emb_b.grad
        rb = "seq2seq/decoder_rnn/multi_rnn_cell/cell_0/lstm_cell"
        out[rb + "/kernel"], out[rb + "/bias"] = lstm_grads(self.dec, "l0", cfg.D, cfg.Hd)
        pb = f"seq2seq/decoder_projection_{cfg.Hd}_{cfg.V}_0"
        out[pb + "/weights"], out[pb + "/biases"] = self.proj_w.grad, self.proj_b.grad
        return out


class SpeedTrainer:
    """forward + backward + TF1-style Adam (bias-corrected step size, eps outside the root) + EMA shadows."""

    def __init__(self, cfg: O.OracleConfig, P: Dict[str, torch.Tensor]):
        self.cfg, self.model, self.t = cfg, SpeedModel(cfg, P), 0
        self.model.train()
        self.params = [p for p in self.model.parameters()]
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.ema = [p.detach().clone() for p in self.params]

    def step(self, x, y, ff_p, rnn_p):
        cfg = self.cfg
        for p in self.params:
            p.grad = None
        loss, ntok = self.model(x, y, ff_p, rnn_p)
        loss.backward()
        self.t += 1
        lr_t = cfg.lr * math.sqrt(1 - cfg.beta2 ** self.t) / (1 - cfg.beta1 ** self.t)
        with torch.no_grad():
            for p, m, v, s in zip(self.params, self.m, self.v, self.ema):
                if p.grad is None:
                    continue
                g = p.grad / max(ntok, 1)
                m.mul_(cfg.beta1).add_(g, alpha=1 - cfg.beta1)
                v.mul_(cfg.beta2).addcmul_(g, g, value=1 - cfg.beta2)
                p.sub_(lr_t * m / (v.sqrt() + cfg.eps))
                s.mul_(cfg.ema_decay).add_(p, alpha=1 - cfg.ema_decay)
        return float(loss.detach()), ntok
