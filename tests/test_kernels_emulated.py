"""CPU: the kernel source of csrc/ executed by the fiber emulation (tests/emu) against the oracle.
Catches indexing / sequencing bugs of the SIMT path and of the host orchestration without a GPU."""
import numpy as np
import pytest

import parity_common as pc


@pytest.mark.parametrize("geo,B,T,L", [(pc.TINY, 3, 19, 5), (pc.SMALL, 4, 50, 6)])
def test_train_step_matches_oracle(emu_lib, geo, B, T, L):
    pc.check_train_step(emu_lib, geo, B, T, L)


def test_train_step_with_dropout(emu_lib):
    pc.check_train_step(emu_lib, pc.TINY, 3, 19, 5, ff=0.1, rnn=0.5)


def test_train_step_explicit_lengths_and_second_subject(emu_lib):
    pc.check_train_step(emu_lib, pc.TWO_SUBJ, 4, 17, 5, give_lens=True, subnet=1)


def test_single_frame_and_unit_batch(emu_lib):
    # T < W (one partial window), B = 1, L = 1 (EOS only)
    pc.check_train_step(emu_lib, pc.TINY, 1, 3, 2)


def test_greedy_decode(emu_lib):
    pc.check_decode(emu_lib, pc.TINY, 6, 21, 6)


def test_beam_decode(emu_lib):
    pc.check_decode(emu_lib, pc.TINY, 4, 21, 6, beam=4)


def test_beam_width_one_equals_greedy(emu_lib):
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.TINY)
    P = pc.make_params(ocfg, eos_bias=-1.0)
    eng = pc.engine_for(pc.TINY, emu_lib, 4, 21, 6, gemm_backend="simt")
    eng.set_all({k: v.numpy() for k, v in P.items()})
    x, _, _ = pc.make_batch(ocfg, 4, 21, 4)
    g, lp = eng.greedy_decode(x, None, max_len=6, temperature=0.5)
    b, sc = eng.beam_decode(x, None, beam=1, max_len=6, temperature=0.5)
    assert (b[:, 0] == g).all()
    assert np.allclose(sc[:, 0], lp.sum(1), atol=1e-5)


def test_attention_train_step(emu_lib):
    """A7: fused score / masked softmax / context kernels + their backward against autograd of the oracle."""
    pc.check_train_step(emu_lib, pc.TINY_ATTN, 3, 19, 5)
    pc.check_train_step(emu_lib, pc.TINY_ATTN, 3, 19, 5, ff=0.1, rnn=0.5)


def test_attention_decode(emu_lib):
    pc.check_decode(emu_lib, pc.TINY_ATTN, 6, 21, 6)
    pc.check_decode(emu_lib, pc.TINY_ATTN, 4, 21, 6, beam=4)


def test_bahdanau_attention(emu_lib):
    """A7, additive score v . tanh(Wq h + Wk enc_s): forward, every gradient (query / keys / score vector / combine, and
    the encoder path through both the context and the keys), greedy and beam decode."""
    pc.check_train_step(emu_lib, pc.TINY_BAH, 3, 19, 5)
    pc.check_train_step(emu_lib, pc.TINY_BAH, 3, 19, 5, ff=0.1, rnn=0.5)
    pc.check_decode(emu_lib, pc.TINY_BAH, 6, 21, 6, margin=1e-4)     # fp32 emulation: a tighter tie margin is enough
    pc.check_decode(emu_lib, pc.TINY_BAH, 4, 21, 6, beam=4, margin=1e-5)


def test_encoder_targets_head(emu_lib):
    """A6: FF head on an encoder layer's outputs + gaussian / categorical loss, forward and backward."""
    pc.check_train_step(emu_lib, pc.TINY_AUX, 3, 19, 5)
    pc.check_train_step(emu_lib, pc.TINY_AUX, 3, 19, 5, ff=0.1, rnn=0.5)
    pc.check_train_step(emu_lib, pc.TINY_AUX_CAT, 4, 21, 5)
    pc.check_train_step(emu_lib, pc.TINY_AUX_CAT, 4, 21, 5, ff=0.1, rnn=0.5, give_lens=True)


def test_encoder_targets_are_optional_per_step(emu_lib):
    """a step without targets skips the head: zero gradients on its parameters, loss = decoder loss only"""
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.TINY_AUX)
    P = pc.make_params(ocfg)
    eng = pc.engine_for(pc.TINY_AUX, emu_lib, 3, 19, 5, gemm_backend="simt")
    eng.set_all({k: v.numpy() for k, v in P.items()})
    x, lens, y = pc.make_batch(ocfg, 3, 19, 5)
    import torch
    lo, no, g, _ = O.loss_and_grads(ocfg, P, torch.from_numpy(x), None, torch.from_numpy(y).long())
    loss, ntok = eng.train_step_grads(x, None, y)
    assert abs(loss - lo) <= 2e-4 * abs(lo)
    G = eng.get_all(1)
    for k, v in G.items():
        if "_projection" in k and "encoder_" in k:
            assert (v == 0).all(), k
    assert eng.last_losses()[2:] == (0.0, 0)
    eng.close()


def test_input_saliency(emu_lib):
    """A13: d(loss)/d(inputs) for the decoder penalty and for the encoder-targets penalty, value and EMA weights."""
    pc.check_saliency(emu_lib, pc.TINY, 3, 19, 5)
    pc.check_saliency(emu_lib, pc.TINY_AUX, 3, 19, 5, which="aux")
    pc.check_saliency(emu_lib, pc.TINY_AUX_CAT, 4, 21, 5, which="decoder", use_ema=True)


def test_gradient_buckets_cover_the_buffer_and_change_nothing(emu_lib):
    """e2t_set_grad_buckets: the staged flush gives bit-identical gradients; the buckets are disjoint, cover every tensor and
    come in completion order (decoder side, encoder layers top to bottom, head, conv)."""
    import numpy as np
    from ecog2txt_b200 import _lib
    from oracle import seq2seq_oracle as O
    geo = dict(pc.TINY_AUX, attention="luong")
    ocfg = O.OracleConfig(**geo)
    P = pc.make_params(ocfg)
    x, lens, y = pc.make_batch(ocfg, 3, 19, 5)
    aux = pc.make_aux_targets(ocfg, lens, 19)
    grads, buckets = [], []
    for on in (False, True):
        eng = pc.engine_for(geo, emu_lib, 3, 19, 5, gemm_backend="simt", ff_dropout=0.1, rnn_dropout=0.5)
        eng.set_all({k: v.numpy() for k, v in P.items()})
        eng.set_grad_buckets(on)
        eng.set_encoder_targets(aux)
        eng.train_step_grads(x, None, y, seed=3)
        grads.append(eng.get_all(_lib.GRAD))
        buckets.append(eng.grad_buckets())
        total, tensors = eng.flat_buffer(_lib.GRAD)[1], eng.tensors()
        eng.close()
    for k in grads[0]:
        assert np.array_equal(grads[0][k], grads[1][k]), k
    assert buckets[0] == [(0, total)]
    b = buckets[1]
    assert len(b) == 1 + 2 + 1 + 1                      # decoder side, 2 encoder layers, head, conv
    assert sorted(b)[0][0] == 0 and sum(n for _, n in b) == total
    ends = sorted((o, o + n) for o, n in b)
    assert all(ends[i][1] == ends[i + 1][0] for i in range(len(ends) - 1))
    name_at = {off: name for name, (_, off) in tensors.items()}
    assert "decoder_embedding" in name_at[b[0][0]] and "encoder_rnn_1" in name_at[b[1][0]]
    assert "encoder_1_projection" in name_at[b[2][0]] and "encoder_rnn_0" in name_at[b[3][0]] and b[4][0] == 0


def test_attention_general_kernels(emu_lib, monkeypatch):
    """the block-per-row attention kernels (used when the decoder is wider than a warp's registers hold: Hd > 1024)"""
    monkeypatch.setenv("E2T_ATTN_BLOCK", "1")
    pc.check_train_step(emu_lib, pc.TINY_ATTN, 3, 19, 5, ff=0.1, rnn=0.5)
    pc.check_train_step(emu_lib, pc.TINY_BAH, 3, 19, 5)
    pc.check_decode(emu_lib, pc.TINY_ATTN, 4, 21, 6, beam=4)


def test_attention_multi_chunk_staging(emu_lib, monkeypatch):
    """utterances longer than one shared-memory tile of encoder rows: the staged sweeps run chunk by chunk"""
    monkeypatch.setenv("E2T_ATTN_TILE_ROWS", "2")
    pc.check_train_step(emu_lib, pc.TINY_ATTN, 3, 19, 5, ff=0.1, rnn=0.5)
    pc.check_train_step(emu_lib, pc.TINY_BAH, 3, 19, 5)
    pc.check_decode(emu_lib, pc.TINY_ATTN, 6, 21, 6)
    pc.check_decode(emu_lib, pc.TINY_BAH, 4, 21, 6, beam=4, margin=1e-5)


def test_tall_column_sums(emu_lib):
    """T' * B >= 512 rows: the bias gradients go through the two-pass column sums, whose first pass reads 16 bytes per thread
    where the rows are 16-byte aligned (full groups of 4 columns, ragged right edges, and the scalar variant for odd strides)"""
    pc.check_train_step(emu_lib, pc.TINY_AUX, 40, 60, 5)
    pc.check_train_step(emu_lib, pc.TINY_AUX_CAT, 36, 60, 5, ff=0.1, rnn=0.5)


def test_api_error_paths(emu_lib):
    """error behaviour of the optional entry points: every misuse is an int error code + message, never a crash"""
    import numpy as np
    import ctypes as C
    from ecog2txt_b200 import E2TError
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.TINY)
    x, lens, y = pc.make_batch(ocfg, 3, 19, 5)
    plain = pc.engine_for(pc.TINY, emu_lib, 3, 19, 5, gemm_backend="simt")
    with pytest.raises(E2TError, match="no encoder-targets head"):
        plain.set_encoder_targets(np.zeros((3, 19, 3), np.float32))
    with pytest.raises(E2TError, match="bucket index"):
        plain._ck(plain._lib.e2t_grad_bucket_info(plain._h, 0, None, None))          # no step has run yet
    with pytest.raises(E2TError, match="slot must be"):
        plain.wait_staged(2)
    with pytest.raises(E2TError, match="saliency needs decoder targets"):
        plain._ck(plain._lib.e2t_input_saliency(plain._h, 0, x.ctypes.data_as(C.c_void_p), None, None, 0, 3, 19, 5, 0, 1.0, 0.0,
                                                None, None))
    with pytest.raises(E2TError, match="exceeds max_B"):
        plain.input_saliency(np.zeros((4, 19, 6), np.float32), None, np.ones((4, 5), np.int32))
    plain.close()
    head = pc.engine_for(pc.TINY_AUX, emu_lib, 3, 19, 5, gemm_backend="simt")
    with pytest.raises(TypeError):
        head.set_encoder_targets(np.zeros((3, 19, 3), np.float64))                    # gaussian targets are float32
    with pytest.raises(E2TError, match="exceed the capacities"):
        head.set_encoder_targets(np.zeros((3, 40, 3), np.float32))
    head.set_encoder_targets(np.zeros((2, 19, 3), np.float32))                         # set for B = 2 ...
    with pytest.raises(E2TError, match="another batch shape"):
        head.train_step_grads(x, None, y)                                              # ... consumed by a B = 3 step
    head.close()
    with pytest.raises(E2TError, match="aux_layer"):
        pc.engine_for(dict(pc.TINY, aux_layer=5, aux_F=3), emu_lib, 3, 19, 5)
    with pytest.raises(E2TError, match="categorical head"):
        pc.engine_for(dict(pc.TINY, aux_layer=0, aux_F=1, aux_kind="categorical"), emu_lib, 3, 19, 5)


def test_beam_topk_variants(emu_lib, monkeypatch):
    """the one-pass top-k at a vocabulary that needs the 64-per-lane instantiation, with more beams than warps, and the
    general block-wide kernel (vocabularies > 2048) on the same inputs"""
    big_v = dict(pc.TINY, V=300)
    pc.check_decode(emu_lib, big_v, 4, 21, 6, beam=4, margin=1e-5)
    pc.check_decode(emu_lib, big_v, 3, 21, 12, beam=10, margin=1e-5)      # 30 state rows fit max_L * max_B = 36
    monkeypatch.setenv("E2T_BEAM_BLOCK", "1")
    pc.check_decode(emu_lib, big_v, 4, 21, 6, beam=4, margin=1e-5)


def test_maximum_utterance_length(emu_lib):
    """the longest trial the reference's generators emit (max_samples = 1250 frames, data_generators.py:35-42) next to a short
    one in the same batch: 313 recurrent steps per layer and direction; and the decode of such a batch"""
    pc.check_train_step(emu_lib, pc.TINY, 2, 1250, 3)
    pc.check_decode(emu_lib, pc.TINY, 2, 1250, 4, margin=1e-4)


def test_staged_inputs_and_host_buffers(emu_lib):
    """e2t_stage_inputs / E2T_STAGED* (synchronous in the emulation build, same slots and bookkeeping) from page-locked-style
    host buffers (e2t_host_alloc): same loss / gradients as the plain host path, shape mix-ups are errors"""
    import numpy as np
    from ecog2txt_b200 import E2TError, _lib
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.TINY)
    P = pc.make_params(ocfg)
    eng = pc.engine_for(pc.TINY, emu_lib, 4, 19, 5, gemm_backend="simt")
    eng.set_all({k: v.numpy() for k, v in P.items()})
    batches = [pc.make_batch(ocfg, 4, 19, 5, seed=s) for s in range(3)]
    ref = []
    for x, _, y in batches:
        loss, ntok = eng.train_step_grads(x, None, y, seed=7)
        ref.append((loss, ntok, eng.get_all(_lib.GRAD)))
    ring = [eng.host_buffer((4, 19, 6), np.float32) for _ in range(3)]
    with pytest.raises(E2TError, match="never staged"):
        eng._staged_shape = {1: (4, 19, 5, 0)}
        eng.train_step_grads_staged(1)
    for i, (x, _, y) in enumerate(batches):
        ring[i % 3][...] = x
        eng.stage_inputs(i & 1, ring[i % 3], None, y)
        eng.train_step_grads_staged(i & 1, seed=7, want_loss=False)
        ld, nt, la, nf = eng.last_losses()
        assert (ld, nt, la, nf) == (ref[i][0], ref[i][1], 0.0, 0)
        eng.post_losses(i & 3)                         # the pipelined read returns the same numbers
        assert eng.fetch_losses(i & 3) == (ld, nt, la, nf)
        g = eng.get_all(_lib.GRAD)
        for k in g:
            assert np.array_equal(g[k], ref[i][2][k]), k
    with pytest.raises(E2TError, match="another shape"):
        eng._staged_shape[0] = (3, 19, 5, 0)
        eng.train_step_grads_staged(0, seed=7)
    with pytest.raises(E2TError, match="slot must be"):
        eng.post_losses(4)
    eng.close()


def test_attention_long_targets(emu_lib):
    """more decoder steps than the encoder-gradient kernel keeps in registers at once (12) and than a block has warps (16):
    the chunked accumulation over decoder steps and the row-group loop of the attention kernels"""
    pc.check_train_step(emu_lib, pc.TINY_ATTN, 2, 19, 14)
    pc.check_train_step(emu_lib, pc.TINY_BAH, 2, 19, 14)
    pc.check_train_step(emu_lib, pc.TINY_BAH, 2, 19, 19, ff=0.1, rnn=0.5)


def test_attention_odd_width(emu_lib):
    """decoder width not a multiple of 4 (scalar staging of the encoder rows, ragged last register of a lane)"""
    odd = dict(pc.TINY, H=(8, 7), Hd=14)
    pc.check_train_step(emu_lib, dict(odd, attention="luong"), 3, 19, 5)
    pc.check_train_step(emu_lib, dict(odd, attention="bahdanau"), 3, 19, 5, ff=0.1, rnn=0.5)
    pc.check_decode(emu_lib, dict(odd, attention="bahdanau"), 4, 21, 6, beam=3, margin=1e-5)


def test_input_saliency_device_buffers(emu_lib):
    """loc = E2T_DEVICE: inputs used in place, dx / sq_norms written straight into the caller's device buffers (plain host
    memory in the emulation build) -- same numbers as the host-buffer call"""
    import ctypes as C
    import numpy as np
    from ecog2txt_b200 import _lib
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.TINY)
    P = pc.make_params(ocfg)
    eng = pc.engine_for(pc.TINY, emu_lib, 3, 19, 5, gemm_backend="simt")
    eng.set_all({k: v.numpy() for k, v in P.items()})
    x, lens, y = pc.make_batch(ocfg, 3, 19, 5)
    dx_h, sq_h = eng.input_saliency(x, None, y)
    dx_d, sq_d = np.full_like(dx_h, np.nan), np.full_like(sq_h, np.nan)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    eng._ck(eng._lib.e2t_input_saliency(eng._h, 0, vp(x), None, vp(y), _lib.DEVICE, 3, 19, 5, 0, 1.0, 0.0, vp(dx_d), vp(sq_d)))
    eng.sync()
    assert np.array_equal(dx_d, dx_h) and np.array_equal(sq_d, sq_h)
    eng._ck(eng._lib.e2t_input_saliency(eng._h, 0, vp(x), vp(lens), vp(y), _lib.DEVICE, 3, 19, 5, 0, 1.0, 0.0, None, vp(sq_d)))
    eng.sync()
    assert np.array_equal(sq_d, sq_h)             # explicit lengths, norms only
    eng.close()


def test_size_independent_properties(emu_lib):
    """the properties the full-size GPU test pins, on the emulated kernels: zero-padding invariance (lengths are inferred from
    the padding), batch additivity of the summed-loss gradient, eval loss == training loss without dropout, determinism"""
    import numpy as np
    from ecog2txt_b200 import _lib
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.SMALL)
    P = pc.make_params(ocfg)
    x, lens, y = pc.make_batch(ocfg, 6, 50, 6)
    eng = pc.engine_for(pc.SMALL, emu_lib, 6, 62, 6, gemm_backend="simt")
    eng.set_all({k: v.numpy() for k, v in P.items()})
    loss, ntok = eng.train_step_grads(x, None, y, seed=3)
    g1 = eng.get_all(_lib.GRAD)
    loss_b, _ = eng.train_step_grads(x, None, y, seed=3)
    g2 = eng.get_all(_lib.GRAD)
    assert loss == loss_b and all(np.array_equal(g1[k], g2[k]) for k in g1)
    le, ne = eng.eval_loss(x, None, y)
    assert ne == ntok and abs(le - loss) <= 1e-5 * abs(loss)
    xp = np.zeros((6, 62, 32), np.float32)
    xp[:, :50] = x
    loss_p, _ = eng.train_step_grads(xp, None, y, seed=3)
    gp = eng.get_all(_lib.GRAD)
    assert abs(loss_p - loss) <= 1e-5 * abs(loss) and all(pc.rel_err(gp[k], g1[k]) <= 1e-4 for k in g1)
    acc, ltot = None, 0.0
    for lo in (0, 3):
        l_h, _ = eng.train_step_grads(np.ascontiguousarray(x[lo:lo + 3]), None, np.ascontiguousarray(y[lo:lo + 3]), seed=3)
        gh = eng.get_all(_lib.GRAD)
        ltot += l_h
        acc = gh if acc is None else {k: acc[k] + gh[k] for k in gh}
    assert abs(ltot - loss) <= 1e-4 * abs(loss) and all(pc.rel_err(acc[k], g1[k]) <= 1e-4 for k in g1)
    eng.close()


@pytest.mark.parametrize("geo,B,T", [(pc.TINY, 1, 21), (pc.TINY, 3, 21), (pc.SMALL, 4, 50)])
def test_small_batch_greedy_decode(emu_lib, geo, B, T):
    """N3 (online predictor): with at most 4 utterances the decode steps run as k_dec_small_cell / k_dec_small_pick (two
    matrix-vector launches per step, arg-max merged by the last block to arrive) -- against the oracle, one utterance, the
    4-row dispatch boundary, a vocabulary that is not a multiple of the 8 rows per block, D not a multiple of 4."""
    pc.check_decode(emu_lib, geo, B, T, 6, name=f"emu/small_decode/B{B}")
    pc.check_decode(emu_lib, geo, B, T, 6, temperature=0.3, use_ema=True, name=f"emu/small_decode/B{B}_ema")


def test_small_batch_decode_equals_batched_path(emu_lib):
    """The same utterances through the small-batch kernels (B = 4) and the batched decode step (B = 5: GEMM + cell +
    projection + k_greedy_pick): identical tokens, log-probabilities to fp32 rounding."""
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.SMALL)
    P = pc.make_params(ocfg, eos_bias=-1.0)
    x, _, _ = pc.make_batch(ocfg, 5, 50, 4)
    eng = pc.engine_for(pc.SMALL, emu_lib, 5, 50, 6, gemm_backend="simt")
    eng.set_all({k: v.numpy() for k, v in P.items()})
    t5, lp5 = eng.greedy_decode(x, None, max_len=6, temperature=0.5)
    t4, lp4 = eng.greedy_decode(np.ascontiguousarray(x[:4]), None, max_len=6, temperature=0.5)
    eng.close()
    assert (t4 == t5[:4]).all()
    assert np.abs(lp4 - lp5[:4]).max() < 1e-5


def test_hidden_decoder_projection_layer(emu_lib):
    """layer_sizes['decoder_projection'] = [P] (mochastar_word_sequence.yaml:65): a relu + FF-dropout layer between the decoder
    output and the vocabulary projection -- loss, every gradient (incl. the layer's own tensors under their numbered
    '<x>_projection' names, trainers.py:488-520) and greedy / beam decode against the oracle."""
    pc.check_train_step(emu_lib, pc.TINY_PROJ, 3, 19, 5)
    pc.check_train_step(emu_lib, pc.TINY_PROJ, 3, 19, 5, ff=0.1, rnn=0.5)
    pc.check_decode(emu_lib, pc.TINY_PROJ, 6, 21, 6)
    pc.check_decode(emu_lib, pc.TINY_PROJ, 4, 21, 6, beam=4)
    eng = pc.engine_for(pc.TINY_PROJ, emu_lib, 3, 19, 5)
    names = eng.tensors()
    assert names["seq2seq/decoder_projection_16_7_0/weights"][0] == (16, 7)
    assert names["seq2seq/decoder_projection_7_11_1/weights"][0] == (11, 7)      # last layer: stored transposed
    assert "seq2seq/decoder_projection_16_11_0/weights" not in names
    eng.close()
