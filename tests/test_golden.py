"""CPU: the oracle and the kernel-emulation build against the committed golden vectors; the oracle's LSTM cell
against an independent implementation (torch.nn.LSTM after gate re-ordering); the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import golden_common as gc
import parity_common as pc
from oracle import seq2seq_oracle as O


def test_oracle_reproduces_golden_vectors():
    z, P = gc.load()
    ocfg = O.OracleConfig(**pc.TINY)
    Pt = {k: torch.from_numpy(v) for k, v in P.items()}
    loss, ntok, g, acts = O.loss_and_grads(ocfg, Pt, torch.from_numpy(z["x"]), None, torch.from_numpy(z["y"]).long())
    assert ntok == int(z["ntok"]) and abs(loss - float(z["loss"])) < 1e-4 * abs(float(z["loss"]))
    for k, v in g.items():
        assert pc.rel_err(v.numpy(), z["G|" + k.replace("/", "|")]) < 1e-4, k
    toks, _, _ = O.greedy_decode(ocfg, Pt, torch.from_numpy(z["x"]), None, max_len=6, temperature=0.7)
    assert (toks.numpy() == z["greedy_tokens"]).all()


def test_emulated_engine_reproduces_golden_vectors(emu_lib):
    gc.check_engine_against_golden(emu_lib)


def test_hidden_projection_reproduces_golden_vectors(emu_lib):
    """layer_sizes['decoder_projection'] = [7]: the oracle (fp32) and the emulated engine against seq2seq_tiny_proj.npz (fp64)."""
    z, P = gc.load_proj()
    ocfg = O.OracleConfig(**pc.TINY_PROJ)
    Pt = {k: torch.from_numpy(v) for k, v in P.items()}
    loss, ntok, g, acts = O.loss_and_grads(ocfg, Pt, torch.from_numpy(z["x"]), None, torch.from_numpy(z["y"]).long())
    assert ntok == int(z["ntok"]) and abs(loss - float(z["loss"])) < 1e-4 * abs(float(z["loss"]))
    for k, v in g.items():
        assert pc.rel_err(v.numpy(), z["G|" + k.replace("/", "|")]) < 1e-4, k
    toks, _, _ = O.greedy_decode(ocfg, Pt, torch.from_numpy(z["x"]), None, max_len=6, temperature=0.7)
    assert (toks.numpy() == z["greedy_tokens"]).all()
    gc.check_engine_against_proj_golden(emu_lib)


def test_optional_rows_reproduce_golden_vectors(emu_lib):
    """A6 + A7 (Bahdanau) + A13: the oracle (fp32) and the emulated engine against seq2seq_tiny_optional.npz (fp64)."""
    z, P = gc.load_optional()
    ocfg = O.OracleConfig(**gc.GEO_OPT)
    Pt = {k: torch.from_numpy(v) for k, v in P.items()}
    xt, yt, at = torch.from_numpy(z["x"]), torch.from_numpy(z["y"]).long(), torch.from_numpy(z["aux"])
    loss, ntok, g, acts = O.loss_and_grads(ocfg, Pt, xt, None, yt, aux_targets=at)
    assert ntok == int(z["ntok"]) and abs(loss - float(z["loss"])) < 1e-4 * abs(float(z["loss"]))
    assert acts["aux_frames"] == int(z["aux_frames"])
    for k, v in g.items():
        assert pc.rel_err(v.numpy(), z["G|" + k.replace("/", "|")]) < 1e-4, k
    assert pc.rel_err(O.input_gradients(ocfg, Pt, xt, None, yt, aux_targets=at).numpy(), z["dx"]) < 1e-4
    gc.check_engine_against_optional_golden(emu_lib)


def test_oracle_lstm_matches_torch_nn_lstm():
    """TF1 LSTMCell (gates i,j,f,o; forget_bias 1; kernel [In+H,4H]) == torch.nn.LSTM (gates i,f,g,o;
    weight_ih [4H,In]) after re-ordering: an implementation of the cell the oracle did not write."""
    torch.manual_seed(0)
    In, H, B, T = 5, 7, 3, 9
    K = torch.randn(In + H, 4 * H) * 0.3
    bias = torch.randn(4 * H) * 0.1
    x = torch.randn(B, T, In)
    lens = torch.tensor([T, T, T])
    out, h, c = O.lstm_direction(x, lens, K, bias, reverse=False)
    lstm = torch.nn.LSTM(In, H, batch_first=True)
    order = [0, 2, 1, 3]                     # torch (i,f,g,o) <- TF (i,j,f,o) block indices
    Kb = K.reshape(In + H, 4, H)[:, order].reshape(In + H, 4 * H)
    bb = bias.clone().reshape(4, H)
    bb[2] += 1.0                             # forget_bias
    bb = bb[order].reshape(-1)
    with torch.no_grad():
        lstm.weight_ih_l0.copy_(Kb[:In].T)
        lstm.weight_hh_l0.copy_(Kb[In:].T)
        lstm.bias_ih_l0.copy_(bb)
        lstm.bias_hh_l0.zero_()
        ref, (hn, cn) = lstm(x)
    assert torch.allclose(out, ref, atol=1e-5)
    assert torch.allclose(h, hn[0], atol=1e-5) and torch.allclose(c, cn[0], atol=1e-5)


def test_oracle_ragged_equals_per_utterance():
    """dynamic_rnn semantics: a zero-padded batch gives exactly what each utterance gives alone."""
    ocfg = O.OracleConfig(**pc.TINY)
    P = pc.make_params(ocfg)
    x, lens, y = pc.make_batch(ocfg, 4, 19, 5)
    full = O.encoder(ocfg, P, torch.from_numpy(x), None, 0)
    for b in range(4):
        one = O.encoder(ocfg, P, torch.from_numpy(x[b:b + 1, :lens[b]]), None, 0)
        assert torch.allclose(full["final_h"][b], one["final_h"][0], atol=1e-6)
        assert torch.allclose(full["final_c"][b], one["final_c"][0], atol=1e-6)


def test_cuda_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports exactly what include/e2t.h declares."""
    import __graft_entry__ as ge
    from ecog2txt_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "e2t.h")).read()
    declared = set(re.findall(r"\b(e2t_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    path = ge.build_cuda()
    lib = ctypes.CDLL(path)
    for sym in declared:
        getattr(lib, sym)
    lib.e2t_abi_version.restype = ctypes.c_int
    assert lib.e2t_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback_in_product_path():
    """Without a CUDA device e2t_create must fail loudly (no CPU fallback)."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ecog2txt_b200 import Engine, EngineConfig, E2TError
    with pytest.raises(E2TError, match="no CUDA device|CPU fallback"):
        Engine(EngineConfig())


def test_product_package_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, files in os.walk(os.path.join(root, "ecog2txt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_oracle_conv_matches_torch_conv2d():
    """App. D items 2-3 against an independent implementation: reverse within length, then tf.nn.conv2d semantics with a
    (1, W, C, E) kernel at stride W and the tail zero-padded to a multiple of W == F.conv2d over [B, C, 1, T]."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    B, T, C, W, E = 3, 19, 6, 4, 5
    x = torch.randn(B, T, C)
    lens = torch.tensor([19, 7, 12])
    for b in range(B):
        x[b, lens[b]:] = 0
    w4, bias = torch.randn(1, W, C, E) * 0.3, torch.randn(E) * 0.1
    y, lens2 = O.temporal_conv(O.reverse_within_length(x, lens), lens, w4, bias, "relu")
    xr = torch.stack([torch.cat([x[b, :lens[b]].flip(0), x[b, lens[b]:]]) for b in range(B)])
    T2 = -(-T // W)
    xp = F.pad(xr, (0, 0, 0, T2 * W - T)).permute(0, 2, 1).unsqueeze(2)            # [B, C, 1, T2*W]
    ref = F.conv2d(xp, w4.permute(3, 2, 0, 1), bias, stride=(1, W)).squeeze(2).permute(0, 2, 1)
    assert torch.allclose(y, torch.relu(ref), atol=1e-5)
    assert lens2.tolist() == [5, 2, 3]


def test_oracle_bilstm_matches_packed_torch_lstm():
    """dynamic_rnn semantics on ragged batches (state frozen and outputs zero past the length, the backward direction reversed
    within each length) == torch.nn.LSTM(bidirectional=True) on a packed sequence, after the TF -> torch gate re-ordering."""
    from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
    torch.manual_seed(1)
    In, H, B, T = 5, 7, 4, 9
    lens = torch.tensor([9, 3, 6, 1])
    x = torch.randn(B, T, In)
    for b in range(B):
        x[b, lens[b]:] = 0
    Ks = [torch.randn(In + H, 4 * H) * 0.3 for _ in range(2)]
    bs = [torch.randn(4 * H) * 0.1 for _ in range(2)]
    outs, hs, cs = zip(*(O.lstm_direction(x, lens, Ks[d], bs[d], reverse=bool(d)) for d in range(2)))
    lstm = torch.nn.LSTM(In, H, batch_first=True, bidirectional=True)
    order = [0, 2, 1, 3]                     # torch (i,f,g,o) <- TF (i,j,f,o) block indices
    with torch.no_grad():
        for d, sfx in enumerate(("", "_reverse")):
            Kb = Ks[d].reshape(In + H, 4, H)[:, order].reshape(In + H, 4 * H)
            bb = bs[d].clone().reshape(4, H)
            bb[2] += 1.0                     # forget_bias
            getattr(lstm, "weight_ih_l0" + sfx).copy_(Kb[:In].T)
            getattr(lstm, "weight_hh_l0" + sfx).copy_(Kb[In:].T)
            getattr(lstm, "bias_ih_l0" + sfx).copy_(bb[order].reshape(-1))
            getattr(lstm, "bias_hh_l0" + sfx).zero_()
        packed = pack_padded_sequence(x, lens, batch_first=True, enforce_sorted=False)
        ref, (hn, cn) = lstm(packed)
        ref, _ = pad_packed_sequence(ref, batch_first=True, total_length=T)
    assert torch.allclose(torch.cat(outs, dim=2), ref, atol=1e-5)
    for d in range(2):
        assert torch.allclose(hs[d], hn[d], atol=1e-5) and torch.allclose(cs[d], cn[d], atol=1e-5)


def test_oracle_loss_matches_torch_cross_entropy():
    """App. D item 7: the masked, summed CE of train_loss == F.cross_entropy(ignore_index=pad, reduction='sum') on its logits."""
    import torch.nn.functional as F
    ocfg = O.OracleConfig(**pc.TINY)
    P = pc.make_params(ocfg)
    x, lens, y = pc.make_batch(ocfg, 4, 19, 5)
    yt = torch.from_numpy(y).long()
    loss, ntok, acts = O.train_loss(ocfg, P, torch.from_numpy(x), None, yt)
    ref = F.cross_entropy(acts["logits"].reshape(-1, ocfg.V), yt.reshape(-1), ignore_index=ocfg.pad_id, reduction="sum")
    assert abs(float(loss) - float(ref)) < 1e-4 * float(ref) and ntok == int((y != 0).sum())


def test_speed_mode_oracle_matches_explicit_loops():
    """oracle/speed_mode.py (torch.nn.LSTM / oneDNN, packed sequences -- the CPU baseline bench.py times) against the
    explicit-loop oracle on a ragged batch that includes an empty utterance: loss and every gradient, fp32 tolerance."""
    import parity_common as pc
    from oracle import speed_mode as S
    ocfg = O.OracleConfig(**pc.SMALL)
    P = pc.make_params(ocfg)
    x, lens, y = pc.make_batch(ocfg, 9, 50, 6)
    x[4] = 0.0
    xt, yt = torch.from_numpy(x), torch.from_numpy(y).long()
    lo, no, g, _ = O.loss_and_grads(ocfg, P, xt, None, yt)
    m = S.SpeedModel(ocfg, P)
    m.eval()
    loss, ntok = m(xt, yt)
    loss.backward()
    assert ntok == no and abs(float(loss) - lo) <= 1e-5 * abs(lo)
    gs = m.canonical_grads()
    assert set(gs) == set(g)
    for k in g:
        assert pc.rel_err(gs[k].numpy(), g[k].numpy()) <= 2e-4, k
    # one optimiser step of the trainer = the explicit oracle's Adam + EMA
    tr = S.SpeedTrainer(ocfg, P)
    tr.model.eval()
    tr.step(xt, yt, 0.0, 0.0)
    opt = O.AdamEMA(ocfg, dict(P))
    P2 = dict(P)
    opt.step(P2, g, 1.0 / no)
    pb = f"seq2seq/decoder_projection_{ocfg.Hd}_{ocfg.V}_0/weights"
    assert pc.rel_err(tr.model.proj_w.detach().numpy(), P2[pb].numpy()) <= 1e-5
