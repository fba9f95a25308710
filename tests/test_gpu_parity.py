"""-m gpu: the CUDA path (through the C-ABI) against the oracle on a real B200."""
import numpy as np
import pytest

import parity_common as pc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("backend", ["simt", "auto"])
@pytest.mark.parametrize("geo,B,T,L", [(pc.TINY, 3, 19, 5), (pc.SMALL, 8, 50, 6), (pc.MEDIUM, 32, 120, 7)])
def test_train_step_matches_oracle(gpu_lib, backend, geo, B, T, L):
    # fp32 SIMT: 2e-4 relative; tcgen05 kind::tf32 (10-bit mantissa operands, fp32 accumulate): 1e-2
    tol = 2e-4 if backend == "simt" else 1e-2
    pc.check_train_step(gpu_lib, geo, B, T, L, backend=backend, tol=tol)


@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_train_step_with_dropout(gpu_lib, backend):
    tol = 2e-4 if backend == "simt" else 1e-2
    pc.check_train_step(gpu_lib, pc.MEDIUM, 16, 96, 6, ff=0.1, rnn=0.5, backend=backend, tol=tol)


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
    pc.check_train_step(gpu_lib, pc.WIDE, B, T, 5, ff=ff, rnn=rnn, backend="auto", tol=1e-2)
    c = pc.check_train_step.last_counters
    assert c["persistent_rnn_launches"] == 4, c   # 2 layers x (forward + backward)


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


@pytest.mark.parametrize("backend,tol", [("simt", 2e-4), ("auto", 1e-2)])
def test_attention_train_and_decode(gpu_lib, backend, tol):
    """A7 (optional Luong attention): training step (loss, every gradient incl. the attention tensors and the encoder
    path through the attention), greedy and beam decode, CUDA-core and tensor-core GEMM backends."""
    pc.check_train_step(gpu_lib, pc.MEDIUM_ATTN, 16, 96, 6, backend=backend, tol=tol)
    pc.check_train_step(gpu_lib, pc.MEDIUM_ATTN, 16, 96, 6, ff=0.1, rnn=0.5, backend=backend, tol=tol)
    pc.check_decode(gpu_lib, pc.MEDIUM_ATTN, 8, 96, 6, backend=backend)
    pc.check_decode(gpu_lib, pc.MEDIUM_ATTN, 4, 96, 6, beam=4, backend=backend)


@pytest.mark.parametrize("backend,tol", [("simt", 2e-4), ("auto", 2e-4)])
def test_golden_vectors(gpu_lib, backend, tol):
    """The committed golden vectors (tests/golden/make_golden.py) through the CUDA path; TINY shapes stay on the fp32
    CUDA-core kernels even with backend=auto (tensor-core eligibility needs 64^3 work), hence the fp32 tolerance."""
    import golden_common as gc
    gc.check_engine_against_golden(gpu_lib, backend=backend, tol=tol)


def test_sequence_network_fit_on_gpu(gpu_lib, tmp_path):
    """TFRecords -> SequenceNetwork.fit -> checkpoint -> restore_and_assess through libe2t.so (tensor-core path)."""
    from ecog2txt_b200 import SequenceNetwork
    from ecog2txt_b200.subjects import make_synthetic_subject
    vocab = ["<pad>", "<EOS>", "<OOV>"] + [f"w{i:02d}_" for i in range(37)]
    s = make_synthetic_subject(400, vocab, str(tmp_path / "tf"), n_train_blocks=3, n_valid_blocks=1,
                               utterances_per_block=32, T=96, C=64, n_sentences=10, ragged=True, seed=0)
    s.data_generator.corpus.max_words = 6
    s.write_tf_records_maybe()
    manifest = {"layer_sizes": {"encoder_embedding": [32], "encoder_rnn": [64, 64], "decoder_embedding": [24],
                                "decoder_rnn": [128], "decoder_projection": []},
                "FF_dropout": 0.1, "RNN_dropout": 0.3, "TEMPORALLY_CONVOLVE": True, "EMA_decay": 0.9, "N_epochs": 60,
                "beam_width": 1, "temperature": 0.384, "assessment_epoch_interval": 20}
    net = SequenceNetwork(manifest, VERBOSE=False, N_cases=32, max_hyp_length=8, learning_rate=5e-3)
    net.checkpoint_path = str(tmp_path / "ckpt" / "model.ckpt")
    a = net.fit([s])
    assert net._engine.counter("persistent_rnn_launches") > 0
    wer = a["training"].decoder_word_error_rates
    assert wer[-1] < 0.2, wer                     # 10 fixed sentences are learnable: WER -> ~0 on the training set
    vwer = a["validation"].decoder_word_error_rates
    assert vwer[-1] < 0.75, vwer   # 32 held-out utterances vs 96 training ones: generalises, loosely (chance is ~1.0)
    res = net.restore_and_assess([s], 60)
    assert abs(res["training"].word_error_rate - wer[-1]) < 1e-9
    net.beam_width = 4
    res_b = net.restore_and_assess([s], 60)
    assert res_b["training"].word_error_rate <= wer[-1] + 0.05


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
