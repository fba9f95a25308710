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
