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
