"""CPU: the SequenceNetwork drop-in class end to end (TFRecords -> fit -> checkpoints -> restore_and_assess) with the
arithmetic executed by the kernel-emulation build of csrc (TEST-ONLY); the product path runs the same host code on
libe2t.so (see the -m gpu twin in test_gpu_sequence_network.py)."""
import os

import numpy as np
import pytest

from ecog2txt_b200 import SequenceNetwork
from ecog2txt_b200 import params as prm
from ecog2txt_b200.subjects import make_synthetic_subject

VOCAB = ["<pad>", "<EOS>", "<OOV>"] + [f"w{i}_" for i in range(8)]
MANIFEST = {   # the keys SequenceNetwork pulls from the experiment manifest (mochastar_word_sequence.yaml:3-5,11-12,25,29,62-75,88)
    "layer_sizes": {"encoder_embedding": [5], "encoder_rnn": [8, 8], "decoder_embedding": [6], "decoder_rnn": [16],
                    "decoder_projection": []},
    "FF_dropout": 0.0, "RNN_dropout": 0.0, "TEMPORALLY_CONVOLVE": True, "EMA_decay": 0.5, "N_epochs": 40,
    "beam_width": 1, "temperature": 0.384, "assessment_epoch_interval": 10, "tf_summaries_dir": None,
}


def _subject(tmp_path, sid=400, seed=0, **kw):
    s = make_synthetic_subject(sid, VOCAB, str(tmp_path / f"tf_{sid}"), n_train_blocks=2, n_valid_blocks=1,
                               utterances_per_block=6, T=16, C=6, n_sentences=3, ragged=True, seed=seed, **kw)
    s.sampling_rate_decimated = 50.0      # decimation_factor = round(200 / 50) = 4
    s.data_generator.corpus.max_words = 3
    return s


def test_subject_protocol_and_tfrecords(tmp_path):
    s = _subject(tmp_path)
    assert s.subnet_id == 400 and s.decimation_factor == 4
    assert s.block_ids == {"training": {1, 2}, "validation": {3}, "testing": set()}
    classes = s.write_tf_records_maybe()
    assert all(os.path.exists(s.tf_record_partial_path.format(b)) for b in (1, 2, 3))
    assert set(classes) <= set(VOCAB[3:])
    assert s.data_manifests["encoder_inputs"].num_features == 6
    assert s.data_manifests["decoder_targets"].num_features == len(VOCAB)
    s2 = _subject(tmp_path, sid=401, pretrain_all_blocks=True)
    assert s2.block_ids["training"] == {1, 2, 3}        # subjects.py:123-126


def test_fit_checkpoint_restore_assess(tmp_path, emu_lib):
    s = _subject(tmp_path)
    s.write_tf_records_maybe()
    net = SequenceNetwork(MANIFEST, training_GPUs=[0], TARGETS_ARE_SEQUENCES=True, VERBOSE=False, N_cases=6,
                          max_hyp_length=5, learning_rate=2e-2, lib=emu_lib, gemm_backend="simt")
    net.checkpoint_path = str(tmp_path / "ckpt" / "model.ckpt")
    a = net.fit([s])
    tr = a["training"]
    assert tr.decoder_word_error_rates.shape == (4,) and tr.decoder_accuracies.shape == (4,)
    assert tr.losses[-1] < 0.5 * tr.losses[0], tr.losses            # it learns
    assert tr.decoder_word_error_rates[-1] <= tr.decoder_word_error_rates[0]
    # checkpoints under the TF naming convention the reference parses (trainers.py:240-252,444-554)
    for ep in (10, 20, 30, 40):
        assert os.path.exists(f"{net.checkpoint_path}-{ep}.index")
    shapes = prm.variable_to_shape_map(net.checkpoint_path, 40)
    assert shapes["seq2seq/subnet_400/encoder_embedding_6_5_0/weights"] == [1, 4, 6, 5]
    assert shapes["seq2seq/encoder_rnn_1/bidirectional_rnn/bw/multi_rnn_cell/cell_0/lstm_cell/kernel"] == [24, 32]
    assert shapes[f"seq2seq/decoder_projection_16_{len(VOCAB)}_0/weights/ExponentialMovingAverage"] == [len(VOCAB), 16]
    w = net.get_weights_as_numpy_array("seq2seq/decoder_rnn/multi_rnn_cell/cell_0/lstm_cell/kernel", 40)
    assert w.shape == (6 + 16, 64)
    # a fresh net restores and reproduces the last assessment
    net2 = SequenceNetwork(MANIFEST, VERBOSE=False, N_cases=6, max_hyp_length=5, lib=emu_lib, gemm_backend="simt")
    net2.checkpoint_path = net.checkpoint_path
    res = net2.restore_and_assess([s], 40)
    assert abs(res["validation"].word_error_rate - a["validation"].decoder_word_error_rates[-1]) < 1e-9
    assert abs(res["training"].accuracy - tr.decoder_accuracies[-1]) < 1e-9
    sent = net2.predict(net2._load_partition(s, "validation")[0][0])
    assert isinstance(sent, str)
    # test-time occlusion (plotters.py:603-640): an empty list changes nothing, silencing channels changes the hypotheses
    net2.inputs_to_occlude = []
    assert net2.restore_and_assess([s], 40)["training"].hypotheses == res["training"].hypotheses
    net2.inputs_to_occlude = [0, 1, 2, 3, 4]
    occ = net2.restore_and_assess([s], 40)["training"]
    assert occ.hypotheses != res["training"].hypotheses and occ.word_error_rate >= res["training"].word_error_rate
    net2.inputs_to_occlude = None


def test_transfer_learning_scopes(tmp_path, emu_lib):
    """sequential_transfer_learn semantics (trainers.py:329-374): train subject A, then restore only the shared
    scope for subject B and train only its private subnet: the shared tensors must stay bit-identical."""
    a_subj, b_subj = _subject(tmp_path, 400, seed=0), _subject(tmp_path, 401, seed=1)
    for s in (a_subj, b_subj):
        s.write_tf_records_maybe()
    m = dict(MANIFEST, N_epochs=10)
    net = SequenceNetwork(m, VERBOSE=False, N_cases=6, max_hyp_length=5, learning_rate=1e-2, lib=emu_lib, gemm_backend="simt")
    net.checkpoint_path = str(tmp_path / "tl" / "model.ckpt")
    net.fit([a_subj], train_vars_scope="seq2seq", reuse_vars_scope=None)
    shared_name = "seq2seq/encoder_rnn_0/bidirectional_rnn/fw/multi_rnn_cell/cell_0/lstm_cell/kernel"
    shared_before = net.get_weights_as_numpy_array(shared_name, 10)
    net.fit([b_subj], train_vars_scope="seq2seq/subnet", reuse_vars_scope="seq2seq/(?!subnet)", _restore_epoch=10)
    shared_after = net.get_weights_as_numpy_array(shared_name, 20)
    assert np.array_equal(shared_before, shared_after)
    conv_b = net.get_weights_as_numpy_array("seq2seq/subnet_401/encoder_embedding_6_5_0/weights", 20)
    assert conv_b.shape == (1, 4, 6, 5)


@pytest.mark.parametrize("stream,F", [("audio_sequence", 4), ("phoneme_sequence", 5)])
def test_encoder_targets_and_saliencies(tmp_path, emu_lib, stream, F):
    """A6 + A13 through the class: an 'encoder_1_targets' stream (audio -> Gaussian, phonemes -> categorical) adds the FF
    head under the reference's '<x>_projection' names; restore_and_get_saliencies with the penalties the caller set
    (MultiSubjectTrainer.get_saliencies zeroes all but one, /root/reference/ecog2txt/trainers.py:703-732)."""
    s = _subject(tmp_path, encoder_targets=stream, encoder_targets_features=F, encoder_targets_penalty_scale=0.5)
    s.write_tf_records_maybe()
    m = dict(MANIFEST, N_epochs=10)
    m["layer_sizes"] = dict(m["layer_sizes"], encoder_1_projection=[7])
    net = SequenceNetwork(m, VERBOSE=False, N_cases=6, max_hyp_length=5, learning_rate=2e-2, lib=emu_lib, gemm_backend="simt")
    net.checkpoint_path = str(tmp_path / "aux" / "model.ckpt")
    net.fit([s])
    shapes = prm.variable_to_shape_map(net.checkpoint_path, 10)
    assert shapes["seq2seq/encoder_1_projection_16_7_0/weights"] == [16, 7]
    assert shapes[f"seq2seq/encoder_1_projection_7_{F}_1/weights"] == [F, 7]          # final layer transposed
    trained = net.get_weights_as_numpy_array(f"seq2seq/encoder_1_projection_7_{F}_1/weights", 10)
    net0 = SequenceNetwork(m, VERBOSE=False, N_cases=6, max_hyp_length=5, lib=emu_lib, gemm_backend="simt")
    eng0 = net0._get_engine([s], 16, 5)
    assert np.abs(trained - eng0.get(f"seq2seq/encoder_1_projection_7_{F}_1/weights")).max() > 1e-3   # the head trains
    # saliencies: decoder penalty only, then the encoder-targets penalty only
    key = "encoder_1_targets"
    old = s.data_manifests[key].penalty_scale
    s.data_manifests[key].penalty_scale = 0.0
    dec = net.restore_and_get_saliencies([s], 10, data_partition="validation", assessment_type="norms")
    s.data_manifests[key].penalty_scale, s.data_manifests["decoder_targets"].penalty_scale = 1.0, 0.0
    aux = net.restore_and_get_saliencies([s], 10, data_partition="validation", assessment_type="norms")
    seqs = net.restore_and_get_saliencies([s], 10, data_partition="validation", assessment_type="sequences")
    s.data_manifests[key].penalty_scale, s.data_manifests["decoder_targets"].penalty_scale = old, 1.0
    assert dec.shape == (6,) and aux.shape == (6,) and (dec > 0).all() and (aux > 0).all()
    assert not np.allclose(dec, aux)
    ex = net._load_partition(s, "validation")
    assert len(seqs) == len(ex) and all(g.shape == e[0].shape for g, e in zip(seqs, ex))
    ref = np.mean([np.sqrt((g ** 2).sum(0)) for g in seqs], axis=0)
    # 'norms' also counts the gradient on the zero frames that complete each trial's last conv window
    assert (ref <= aux * (1 + 1e-5)).all() and np.allclose(ref, aux, rtol=0.2)


def test_hidden_decoder_projection_from_the_manifest(tmp_path, emu_lib):
    """layer_sizes['decoder_projection'] = [P] (mochastar_word_sequence.yaml:65 ships it empty): the hidden layer is built,
    trained and check-pointed under the numbered '<x>_projection' names recover_model_sizes parses (trainers.py:488-520)."""
    s = _subject(tmp_path)
    s.write_tf_records_maybe()
    man = dict(MANIFEST, layer_sizes=dict(MANIFEST["layer_sizes"], decoder_projection=[9]), N_epochs=20)
    net = SequenceNetwork(man, VERBOSE=False, N_cases=6, max_hyp_length=5, learning_rate=2e-2, lib=emu_lib, gemm_backend="simt")
    net.checkpoint_path = str(tmp_path / "ckpt_proj" / "model.ckpt")
    tr = net.fit([s])["training"]
    assert tr.losses[-1] < tr.losses[0], tr.losses
    shapes = prm.variable_to_shape_map(net.checkpoint_path, 20)
    assert shapes["seq2seq/decoder_projection_16_9_0/weights"] == [16, 9]
    assert shapes[f"seq2seq/decoder_projection_9_{len(VOCAB)}_1/weights"] == [len(VOCAB), 9]      # last layer: transposed
    assert f"seq2seq/decoder_projection_16_{len(VOCAB)}_0/weights" not in shapes
    two = dict(MANIFEST, layer_sizes=dict(MANIFEST["layer_sizes"], decoder_projection=[9, 9]))
    net2 = SequenceNetwork(two, VERBOSE=False, N_cases=6, max_hyp_length=5, lib=emu_lib, gemm_backend="simt")
    net2.checkpoint_path = str(tmp_path / "ckpt_proj2" / "model.ckpt")
    with pytest.raises(NotImplementedError):
        net2.fit([s])
