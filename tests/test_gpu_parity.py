"""-m gpu: the CUDA path (through the C-ABI) against the oracle on a real B200."""
import numpy as np
import pytest

import torch

import parity_common as pc
from ecog2txt_b200 import Engine, EngineConfig, _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("backend", ["simt", "auto"])
@pytest.mark.parametrize("geo,B,T,L", [(pc.TINY, 3, 19, 5), (pc.SMALL, 8, 50, 6), (pc.MEDIUM, 32, 120, 7)])
def test_train_step_matches_oracle(gpu_lib, backend, geo, B, T, L):
    # tolerances: parity_common.SIMT_TOL / TC_TOL (<= 10x the errors measured on B200, profiles/parity_r2.json)
    pc.check_train_step(gpu_lib, geo, B, T, L, backend=backend)


@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_train_step_with_dropout(gpu_lib, backend):
    pc.check_train_step(gpu_lib, pc.MEDIUM, 16, 96, 6, ff=0.1, rnn=0.5, backend=backend)


def test_train_step_explicit_lengths_and_second_subject(gpu_lib):
    pc.check_train_step(gpu_lib, pc.TWO_SUBJ, 4, 17, 5, give_lens=True, subnet=1)


@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_greedy_decode(gpu_lib, backend):
    pc.check_decode(gpu_lib, pc.MEDIUM, 24, 100, 8, backend=backend)


@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_beam_decode(gpu_lib, backend):
    pc.check_decode(gpu_lib, pc.MEDIUM, 6, 60, 6, beam=4, backend=backend)


@pytest.mark.parametrize("B,T,ff,rnn", [(40, 100, 0.0, 0.0), (160, 60, 0.1, 0.5), (129, 30, 0.0, 0.0)])
def test_persistent_recurrent_kernels_full_width(gpu_lib, B, T, ff, rnn):
    """H=400 BiLSTM layers through the whole-sequence tcgen05 kernels (1-2 batch tiles, ragged lengths,
    dropout copies), against the oracle; tolerance = tf32 operands (10-bit mantissa), fp32 accumulate."""
    pc.check_train_step(gpu_lib, pc.WIDE, B, T, 5, ff=ff, rnn=rnn, backend="auto")
    c = pc.check_train_step.last_counters
    assert c["persistent_rnn_launches"] == 6, c   # 2 layers x (forward + backward) + the decoder (Hd = 800) forward + backward


def test_persistent_kernels_recover_from_incomplete_tiles(gpu_lib):
    """The persistent recurrent kernels pull a tile when one word per producer warp has arrived and detect a piece that was
    not there yet by the NaN it leaves in the accumulator (or by its tag); the step is then pulled again.  That path almost
    never runs by itself (once in ~3000 steps), so E2T_REC_DBGSKIP=4 forces a re-pull on every fifth step of every kernel
    (forward, both BPTT hand-offs, decoder forward / backward): the results must still match the oracle.  The switch is read
    once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, 'tests'); import parity_common as pc; from ecog2txt_b200 import _lib; "
            "lib = _lib.load(); pc.check_train_step(lib, pc.WIDE, 160, 60, 5, ff=0.1, rnn=0.5, backend='auto'); "
            "pc.check_train_step(lib, pc.WIDE, 40, 100, 5, backend='auto'); print('REPULL_OK')")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, E2T_REC_DBGSKIP="4")
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "REPULL_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_persistent_kernels_are_deterministic(gpu_lib):
    import numpy as np
    from ecog2txt_b200 import _lib
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.WIDE)
    P = pc.make_params(ocfg)
    x, lens, y = pc.make_batch(ocfg, 64, 80, 5)
    outs = []
    for _ in range(2):
        eng = pc.engine_for(pc.WIDE, gpu_lib, 64, 80, 5, gemm_backend="auto")
        eng.set_all({k: v.numpy() for k, v in P.items()})
        eng.train_step_grads(x, None, y, seed=1)
        outs.append(eng.get_all(_lib.GRAD))
        eng.close()
    for k in outs[0]:
        if "decoder_embedding" in k and k.endswith("weights"):
            continue   # atomicAdd scatter: order-dependent rounding
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_side_stream_steps_equal_single_stream(gpu_lib):
    """K optimiser steps with everything the side stream takes (weight gradients, re-packs, the upper tensors' Adam step,
    deferred column sums; joined lazily: DESIGN.md 3.2b) against the same K steps in profiling mode, which keeps every
    launch on ONE stream: weights, EMA shadows and Adam moments must be BIT-identical, read both through the tensor API
    and right behind the last optimiser step (param_join)."""
    import numpy as np
    from ecog2txt_b200 import _lib
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.WIDE)
    P = pc.make_params(ocfg)
    batches = [pc.make_batch(ocfg, 64, 80, 5, seed=s) for s in range(3)]
    res = []
    for single_stream in (False, True):
        eng = pc.engine_for(pc.WIDE, gpu_lib, 64, 80, 5, gemm_backend="auto")
        eng.set_all({k: v.numpy() for k, v in P.items()})
        if single_stream:
            eng.profile_enable(True)
        for i, (x, lens, y) in enumerate(batches):
            _, ntok = eng.train_step_grads(x, None, y, seed=i)
            eng.adam_ema_step(1.0 / ntok)
        out = {w: eng.get_all(w) for w in (_lib.VALUE, _lib.EMA, _lib.ADAM_M, _lib.ADAM_V)}   # right behind the step
        # ... and a decode right behind another step sees the updated weights too
        x, lens, y = batches[0]
        _, ntok = eng.train_step_grads(x, None, y, seed=7)
        eng.adam_ema_step(1.0 / ntok)
        toks, _ = eng.greedy_decode(x[:8], None, max_len=5, use_ema=False)
        res.append((out, toks))
        eng.close()
    for w in res[0][0]:
        for k in res[0][0][w]:
            if "decoder_embedding" in k and k.endswith("weights"):
                continue   # atomicAdd scatter: order-dependent rounding
            assert np.array_equal(res[0][0][w][k], res[1][0][w][k]), (w, k)
    assert np.array_equal(res[0][1], res[1][1])


@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_small_batch_greedy_decode(gpu_lib, backend):
    """N3: k_dec_small_cell / k_dec_small_pick (at most 4 utterances) against the oracle at the decoder width of config 2
    (Hd = 800; batches whose oracle top-2 gaps are >= 6e-3 at every live step), and -- fp32 backend, where both paths are
    fp32 -- against the batched decode step on the same utterances."""
    import numpy as np
    from oracle import seq2seq_oracle as O
    for B in (1, 3):
        pc.check_decode(gpu_lib, pc.WIDE, B, 100, 8, backend=backend, name=f"decode/{backend}/small_batch/B{B}")
    if backend != "simt":
        return
    ocfg = O.OracleConfig(**pc.WIDE)
    P = pc.make_params(ocfg, eos_bias=-1.0)
    x, _, _ = pc.make_batch(ocfg, 5, 100, 4)
    eng = pc.engine_for(pc.WIDE, gpu_lib, 5, 100, 8, gemm_backend=backend)
    eng.set_all({k: v.numpy() for k, v in P.items()})
    t5, lp5 = eng.greedy_decode(x, None, max_len=8, temperature=0.5)
    t4, lp4 = eng.greedy_decode(np.ascontiguousarray(x[:4]), None, max_len=8, temperature=0.5)
    eng.close()
    d = float(np.abs(lp4 - lp5[:4]).max())
    pc.record(f"decode/{backend}/small_batch_vs_batched", logp_abs=d, rows_identical=float((t4 == t5[:4]).all(1).mean()))
    assert (t4 == t5[:4]).all()
    assert d < 1e-4


@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_hidden_decoder_projection_layer(gpu_lib, backend):
    """layer_sizes['decoder_projection'] = [P] (mochastar_word_sequence.yaml:65): the optional relu + FF-dropout layer in front
    of the vocabulary projection, forward / backward / decode against the oracle on both GEMM backends (measured on B200,
    profiles/parity_r2.json: worst gradient tensor 9.5e-3 of its largest entry on the tensor-core backend -- the layer's own
    W1 / b1, whose relu sees tf32 products -- against the backend's usual 3e-2 bound)."""
    pc.check_train_step(gpu_lib, pc.MEDIUM_PROJ, 16, 96, 6, backend=backend,
                        name="train_step/auto/hidden_projection" if backend == "auto" else None)
    if backend == "simt":
        pc.check_train_step(gpu_lib, pc.MEDIUM_PROJ, 16, 96, 6, ff=0.1, rnn=0.5, backend=backend)
    pc.check_decode(gpu_lib, pc.MEDIUM_PROJ, 8, 96, 6, backend=backend)
    pc.check_decode(gpu_lib, pc.MEDIUM_PROJ, 4, 96, 6, beam=4, backend=backend)


@pytest.mark.parametrize("backend,tol", [("simt", 2e-4), ("auto", 1e-2)])
def test_attention_train_and_decode(gpu_lib, backend, tol):
    """A7 (optional Luong attention): training step (loss, every gradient incl. the attention tensors and the encoder
    path through the attention), greedy and beam decode, CUDA-core and tensor-core GEMM backends."""
    pc.check_train_step(gpu_lib, pc.MEDIUM_ATTN, 16, 96, 6, backend=backend)
    pc.check_train_step(gpu_lib, pc.MEDIUM_ATTN, 16, 96, 6, ff=0.1, rnn=0.5, backend=backend)
    pc.check_decode(gpu_lib, pc.MEDIUM_ATTN, 8, 96, 6, backend=backend)
    pc.check_decode(gpu_lib, pc.MEDIUM_ATTN, 4, 96, 6, beam=4, backend=backend)


@pytest.mark.parametrize("backend,tol", [("simt", 2e-4), ("auto", 1e-2)])
def test_bahdanau_attention_train_and_decode(gpu_lib, backend, tol):
    """A7, additive (Bahdanau) score: same coverage as the Luong module."""
    pc.check_train_step(gpu_lib, pc.MEDIUM_BAH, 16, 96, 6, backend=backend)
    pc.check_train_step(gpu_lib, pc.MEDIUM_BAH, 16, 96, 6, ff=0.1, rnn=0.5, backend=backend)
    pc.check_decode(gpu_lib, pc.MEDIUM_BAH, 8, 96, 6, backend=backend)
    # temperature 0.2: at 0.7 no two beam scores of these random weights are 1e-2 apart and the token check would be vacuous
    pc.check_decode(gpu_lib, pc.MEDIUM_BAH, 4, 96, 6, beam=4, backend=backend, temperature=0.2)


@pytest.mark.parametrize("backend,tol", [("simt", 2e-4), ("auto", 2e-4)])
def test_golden_vectors(gpu_lib, backend, tol):
    """The committed golden vectors (tests/golden/make_golden.py) through the CUDA path; TINY shapes stay on the fp32
    CUDA-core kernels even with backend=auto (tensor-core eligibility needs 64^3 work), hence the fp32 tolerance."""
    import golden_common as gc
    gc.check_engine_against_golden(gpu_lib, backend=backend, tol=tol)


def test_hidden_projection_golden_vectors(gpu_lib):
    """The model with a hidden decoder_projection layer against tests/golden/seq2seq_tiny_proj.npz through the CUDA path
    (TINY shapes stay on the fp32 CUDA cores)."""
    import golden_common as gc
    gc.check_engine_against_proj_golden(gpu_lib, backend="simt")


def test_optional_rows_golden_vectors(gpu_lib):
    """A6 + A7 (Bahdanau) + A13 against tests/golden/seq2seq_tiny_optional.npz through the CUDA path."""
    import golden_common as gc
    gc.check_engine_against_optional_golden(gpu_lib, backend="simt")
    gc.check_engine_against_optional_golden(gpu_lib, backend="auto")


def test_online_predictor_graph_replay(gpu_lib):
    """N3: one-utterance greedy decodes of a fixed shape are replayed as a CUDA graph from the third call on and return
    exactly what the eager path returns (also after the weights change: the graph reads the live buffers)."""
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.WIDE)
    P = pc.make_params(ocfg, eos_bias=-1.0)
    eng = pc.engine_for(pc.WIDE, gpu_lib, 1, 100, 8, gemm_backend="auto")
    eng.set_all({k: v.numpy() for k, v in P.items()})
    x, _, _ = pc.make_batch(ocfg, 1, 100, 4)
    outs = [eng.greedy_decode(x, None, max_len=8, temperature=0.5) for _ in range(4)]
    assert eng.counter("decode_graph_replays") >= 2
    for t, lp in outs[1:]:
        assert (t == outs[0][0]).all() and np.array_equal(lp, outs[0][1])
    x2, _, _ = pc.make_batch(ocfg, 1, 100, 4, seed=5)
    t_graph, _ = eng.greedy_decode(x2, None, max_len=8, temperature=0.5)
    eng2 = pc.engine_for(pc.WIDE, gpu_lib, 1, 100, 8, gemm_backend="auto")
    eng2.set_all({k: v.numpy() for k, v in P.items()})
    t_eager, _ = eng2.greedy_decode(x2, None, max_len=8, temperature=0.5)
    assert (t_graph == t_eager).all()
    eng.close(); eng2.close()


def _full_size_engine(gpu_lib, B, T=400, **kw):
    """BASELINE.json config 2 at full size: 256-channel ECoG, 3x400 BiLSTM, 800 LSTM decoder, V = 1806."""
    from ecog2txt_b200.params import init_engine
    geo = dict(subnet_ids=(400,), subnet_C=(256,), subnet_W=(12,), E=100, H=(400, 400, 400), D=150, Hd=800, V=1806)
    eng = Engine(EngineConfig(**geo, max_B=B, max_T=T, max_L=12, **kw), lib=gpu_lib)
    init_engine(eng, seed=1)
    return eng


def _full_size_batch(B, T=400, L=11, seed=0, ragged=True):
    rs = np.random.RandomState(seed)
    x = rs.randn(B, T, 256).astype(np.float32)
    lens = rs.randint(T // 2, T + 1, size=B) if ragged else np.full(B, T)
    lens[0] = T
    for b in range(B):
        x[b, lens[b]:] = 0.0
    y = np.zeros((B, L), np.int32)
    for b in range(B):
        n = rs.randint(3, L)
        y[b, :n] = rs.randint(3, 1806, size=n)
        y[b, n] = 1
    return x, lens.astype(np.int32), y


def test_full_size_properties(gpu_lib):
    """Config-2 shapes are far beyond what the oracle finishes in seconds, so the full-size path is pinned through
    size-independent properties: (1) bit-identical gradients run to run, (2) zero-padding invariance (T = 400 vs the
    same utterances padded to T = 436: lengths are inferred from the padding), (3) batch additivity (the gradient of the
    summed loss over 256 utterances = the sum over its two halves), (4) the forward-only loss equals the training loss
    without dropout."""
    B = 256
    x, lens, y = _full_size_batch(B)
    eng = _full_size_engine(gpu_lib, B, T=436)
    loss, ntok = eng.train_step_grads(x, None, y, seed=3)
    g1 = eng.get_all(_lib.GRAD)
    loss_b, _ = eng.train_step_grads(x, None, y, seed=3)
    g2 = eng.get_all(_lib.GRAD)
    assert loss == loss_b and np.isfinite(loss) and ntok == int((y != 0).sum())
    emb = "seq2seq/decoder_embedding_1806_150_0/weights"      # atomicAdd scatter: order-dependent rounding
    for k in g1:
        if k != emb:
            assert np.array_equal(g1[k], g2[k]), k
    le, ne = eng.eval_loss(x, None, y)
    assert ne == ntok and abs(le - loss) <= 1e-5 * abs(loss)
    # (2) padding invariance
    xp = np.zeros((B, 436, 256), np.float32)
    xp[:, :400] = x
    loss_p, ntok_p = eng.train_step_grads(xp, None, y, seed=3)
    gp = eng.get_all(_lib.GRAD)
    assert ntok_p == ntok and abs(loss_p - loss) <= 1e-5 * abs(loss)
    for k in g1:
        assert pc.rel_err(gp[k], g1[k]) <= 1e-4, k
    # (3) batch additivity: different batch tiling (one 128-row tile instead of two) and summation order
    acc = None
    ltot = 0.0
    for lo in (0, 128):
        l_h, _ = eng.train_step_grads(np.ascontiguousarray(x[lo:lo + 128]), None, np.ascontiguousarray(y[lo:lo + 128]), seed=3)
        gh = eng.get_all(_lib.GRAD)
        ltot += l_h
        acc = gh if acc is None else {k: acc[k] + gh[k] for k in gh}
    assert abs(ltot - loss) <= 2e-4 * abs(loss)
    for k in g1:
        assert pc.rel_err(acc[k], g1[k]) <= 5e-3, (k, pc.rel_err(acc[k], g1[k]))
    eng.close()


def test_degenerate_inputs(gpu_lib):
    """Edge cases of the reference's data: an all-zero (empty) utterance inside a batch, a single frame, a single
    utterance, targets that are EOS only -- against the oracle, tensor-core backend."""
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.MEDIUM)
    P = pc.make_params(ocfg)
    x, lens, y = pc.make_batch(ocfg, 5, 40, 4)
    x[2] = 0.0                      # empty utterance: length 0, encoder states zero
    x[3, 1:] = 0.0                  # one frame
    y[4] = 0
    y[4, 0] = ocfg.eos_id           # EOS-only target
    eng = pc.engine_for(pc.MEDIUM, gpu_lib, 5, 40, 4, gemm_backend="auto")
    eng.set_all({k: v.numpy() for k, v in P.items()})
    lo, no, g, _ = O.loss_and_grads(ocfg, P, torch.from_numpy(x), None, torch.from_numpy(y).long())
    loss, ntok = eng.train_step_grads(x, None, y, seed=0)
    assert ntok == no and abs(loss - lo) <= 1e-2 * abs(lo)
    G = eng.get_all(_lib.GRAD)
    for k, v in G.items():
        assert pc.rel_err(v, g[k].numpy()) <= 5e-2, k
    # B = 1 decodes against the oracle, every utterance of the batch on its own (incl. the empty and the one-frame one)
    for b in range(5):
        xb = np.ascontiguousarray(x[b:b + 1])
        t_ref, lp_ref, logits = O.greedy_decode(ocfg, P, torch.from_numpy(xb), None, max_len=4)
        toks, logp = eng.greedy_decode(xb, None, max_len=4)
        assert toks.shape == (1, 4)
        top2 = logits.topk(2, dim=2).values
        if float((top2[..., 0] - top2[..., 1]).min()) > 1e-3:
            assert (toks == t_ref.numpy()).all(), (b, toks, t_ref)
            assert np.abs(logp - lp_ref.numpy()).max() < 2e-3
    eng.close()


def test_staged_inputs_match_host_path(gpu_lib):
    """e2t_stage_inputs (the prefetch pipeline bench.py's e2e leg and SequenceNetwork.fit use): staging batch i+1 while batch i
    trains gives exactly the loss / gradients of the plain host path, in both slots, and guards against shape mix-ups."""
    from oracle import seq2seq_oracle as O
    from ecog2txt_b200 import E2TError
    ocfg = O.OracleConfig(**pc.MEDIUM)
    P = pc.make_params(ocfg)
    eng = pc.engine_for(pc.MEDIUM, gpu_lib, 16, 96, 6, gemm_backend="auto")
    eng.set_all({k: v.numpy() for k, v in P.items()})
    batches = [pc.make_batch(ocfg, 16, 96, 6, seed=s) for s in range(3)]
    ref = []
    for x, _, y in batches:
        loss, ntok = eng.train_step_grads(x, None, y, seed=7)
        ref.append((loss, ntok, eng.get_all(_lib.GRAD)))
    eng.stage_inputs(0, batches[0][0], None, batches[0][2])
    for i, (x, _, y) in enumerate(batches):
        if i + 1 < len(batches):
            eng.stage_inputs((i + 1) & 1, batches[i + 1][0], None, batches[i + 1][2])     # overlaps the step below
        loss, ntok = eng.train_step_grads_staged(i & 1, seed=7)
        assert (loss, ntok) == ref[i][:2]
        g = eng.get_all(_lib.GRAD)
        for k in g:
            if "decoder_embedding" in k and k.endswith("weights"):
                continue   # atomicAdd scatter
            assert np.array_equal(g[k], ref[i][2][k]), k
    with pytest.raises(E2TError):
        eng._staged_shape[0] = (8, 96, 6, 0)        # a slot holds what was staged, not what the caller claims
        eng.train_step_grads_staged(0, seed=7)
    eng.close()


@pytest.mark.parametrize("backend,tol", [("simt", 2e-4), ("auto", 1e-2)])
def test_encoder_targets_head(gpu_lib, backend, tol):
    """A6: the encoder-targets head (FF 2H -> hidden -> F on encoder layer 1; Gaussian and categorical targets), loss and
    every gradient -- the head's own tensors and the extra gradient that reaches layers 0-1 and the conv through it."""
    pc.check_train_step(gpu_lib, pc.MEDIUM_AUX, 16, 96, 6, backend=backend)
    pc.check_train_step(gpu_lib, pc.MEDIUM_AUX, 16, 96, 6, ff=0.1, rnn=0.5, backend=backend)
    pc.check_train_step(gpu_lib, pc.MEDIUM_AUX_CAT, 16, 96, 6, backend=backend)
    pc.check_train_step(gpu_lib, pc.TINY_AUX_CAT, 4, 21, 5, backend=backend)


@pytest.mark.parametrize("backend,tol", [("simt", 2e-4), ("auto", 1e-2)])
def test_input_saliency(gpu_lib, backend, tol):
    """A13: d(loss)/d(encoder_inputs) against autograd of the oracle, decoder penalty / encoder-targets penalty, EMA weights."""
    pc.check_saliency(gpu_lib, pc.MEDIUM, 16, 96, 6, backend=backend, tol=tol)
    pc.check_saliency(gpu_lib, pc.MEDIUM_AUX, 16, 96, 6, backend=backend, tol=tol, which="aux")
    pc.check_saliency(gpu_lib, pc.MEDIUM_AUX_CAT, 16, 96, 6, backend=backend, tol=tol, which="decoder", use_ema=True)
    pc.check_saliency(gpu_lib, pc.WIDE, 40, 100, 5, backend=backend, tol=tol)


def test_full_size_saliency_and_encoder_targets(gpu_lib):
    """Config-2 shapes with the reference's head (encoder_1_projection = [225], 13 MFCC-type targets, yaml:68-69,81):
    (1) linearity -- the saliency under both penalties = the sum of the two single-penalty saliencies; (2) a directional
    finite difference of the forward-only loss along the saliency agrees with |dx|^2; (3) padding frames beyond the last
    conv window carry no gradient; (4) the training loss is the sum of its two terms and both are finite."""
    B = 64
    x, lens, y = _full_size_batch(B)
    eng = _full_size_engine(gpu_lib, B, aux_layer=1, aux_hidden=225, aux_F=13, aux_kind="gaussian", aux_penalty=0.1)
    rs = np.random.RandomState(4)
    aux = rs.randn(B, 400, 13).astype(np.float32)
    for b in range(B):
        aux[b, lens[b]:] = 0
    eng.set_encoder_targets(aux)
    loss, ntok = eng.train_step_grads(x, None, y, seed=3)
    ld, nt, la, nf = eng.last_losses()
    assert np.isfinite(loss) and abs(loss - (ld + la)) <= 1e-5 * abs(loss) and nt == ntok
    assert nf == int(np.ceil(lens / 12).sum()) and la > 0
    sal = {}
    for name, (pd, pa) in dict(dec=(1.0, 0.0), aux=(0.0, 0.1), both=(1.0, 0.1)).items():
        eng.set_encoder_targets(aux)
        sal[name], _ = eng.input_saliency(x, None, y, decoder_penalty=pd, aux_penalty=pa, want_norms=False)
    assert np.abs(sal["dec"]).max() > 0 and np.abs(sal["aux"]).max() > 0
    assert pc.rel_err(sal["dec"] + sal["aux"], sal["both"]) <= 1e-3
    for b in range(B):
        assert (sal["both"][b, -(-lens[b] // 12) * 12:] == 0).all()
    # the same model on the fp32 CUDA-core backend: its saliency agrees with the tensor-core one, and a central finite
    # difference of ITS forward-only loss along the saliency direction reproduces <dx, d> (tf32 rounding of the loss
    # would swamp the difference, hence fp32 for this leg)
    eng_s = _full_size_engine(gpu_lib, B, gemm_backend="simt", aux_layer=1, aux_hidden=225, aux_F=13, aux_kind="gaussian",
                              aux_penalty=0.1)
    eng_s.set_encoder_targets(aux)
    sal_s, _ = eng_s.input_saliency(x, None, y, want_norms=False)
    assert pc.rel_err(sal["both"], sal_s) <= 5e-2, pc.rel_err(sal["both"], sal_s)
    d = sal_s.copy()
    for b in range(B):
        d[b, lens[b]:] = 0        # keep the inferred lengths unchanged
    d /= np.sqrt((d.astype(np.float64) ** 2).sum())
    g_dir = float((sal_s.astype(np.float64) * d).sum())
    eps = 5e-2
    vals = []
    for sgn in (1.0, -1.0):
        eng_s.set_encoder_targets(aux)
        vals.append(eng_s.eval_loss((x + sgn * eps * d).astype(np.float32), lens, y)[0])
    fd = (vals[0] - vals[1]) / (2 * eps)
    assert abs(fd - g_dir) <= 5e-2 * abs(g_dir), (fd, g_dir)
    eng_s.close()
    eng.close()


def test_gradient_buckets_and_device_side_scale(gpu_lib):
    """e2t_set_grad_buckets on the tensor-core path (WIDE: persistent recurrent kernels, permuted gate order un-permuted per
    bucket): bit-identical gradients, buckets cover the buffer, a side stream can wait on every bucket event; and
    e2t_adam_ema_step_dev (token count read on the device) updates exactly like the host-scale call."""
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.WIDE)
    P = pc.make_params(ocfg)
    x, lens, y = pc.make_batch(ocfg, 64, 80, 5)
    ntok_ref = int((y != 0).sum())
    out = []
    for on in (False, True):
        eng = pc.engine_for(pc.WIDE, gpu_lib, 64, 80, 5, gemm_backend="auto", ff_dropout=0.1, rnn_dropout=0.5)
        eng.set_all({k: v.numpy() for k, v in P.items()})
        eng.set_grad_buckets(on)
        loss, ntok = eng.train_step_grads(x, None, y, seed=1)
        assert ntok == ntok_ref
        side = torch.cuda.Stream()
        b = eng.grad_buckets()
        for i in range(len(b)):
            eng.grad_bucket_wait(i, side.cuda_stream)
        side.synchronize()
        g = eng.get_all(_lib.GRAD)
        if on:
            eng.adam_ema_step_dev(torch.tensor([float(ntok)], device="cuda"))
        else:
            eng.adam_ema_step(1.0 / ntok)
        out.append((loss, g, eng.get_all(_lib.VALUE), eng.get_all(_lib.EMA), b, eng.flat_buffer(_lib.GRAD)[1]))
        eng.close()
    (l0, g0, w0, s0, b0, total), (l1, g1, w1, s1, b1, _) = out
    assert l0 == l1 and b0 == [(0, total)]
    assert len(b1) == 1 + 2 + 1 and sum(n for _, n in b1) == total
    for k in g0:
        if "decoder_embedding" in k and k.endswith("weights"):
            assert pc.rel_err(g1[k], g0[k]) < 1e-5      # atomicAdd scatter: order-dependent rounding
            continue
        assert np.array_equal(g0[k], g1[k]), k
        assert np.allclose(w0[k], w1[k], rtol=1e-6, atol=1e-8) and np.allclose(s0[k], s1[k], rtol=1e-6, atol=1e-8), k
