"""CPU: MultiSubjectTrainer (the caller of the hot path) on the kernel-emulation build:
sequential / parallel transfer-learning schedules (trainers.py:303-374), checkpoint discovery (:240-252),
recover_model_sizes on our checkpoints (:444-554) and the online predictor (:925-963)."""
import os

import numpy as np

from ecog2txt_b200 import MultiSubjectTrainer
from ecog2txt_b200 import params as prm
from test_sequence_network import MANIFEST, VOCAB, _subject


def _trainer(tmp_path, emu_lib, ids=(400, 401), subject_kw=None, **sn):
    subjects = [_subject(tmp_path, sid, seed=i, **(subject_kw or {})) for i, sid in enumerate(ids)]
    for s in subjects:
        s.write_tf_records_maybe()
    manifest = {sid: dict(MANIFEST, token_type="word_sequence", decoder_targets_penalty_scale=1.0) for sid in ids}
    sn_kwargs = dict(N_cases=6, max_hyp_length=5, learning_rate=1e-2, lib=emu_lib, gemm_backend="simt", **sn)
    return MultiSubjectTrainer(manifest, list(ids), checkpoint_dir=str(tmp_path / "ckpt"), SN_kwargs=sn_kwargs,
                               VERBOSE=False, subjects=subjects)


def test_sequential_transfer_learn(tmp_path, emu_lib):
    tr = _trainer(tmp_path, emu_lib, assessment_epoch_interval=5)
    assert tr.ecog_subjects[0].pretrain_all_blocks and not tr.ecog_subjects[1].pretrain_all_blocks
    assert tr.ecog_subjects[0].block_ids["training"] == {1, 2, 3}      # first subject pre-trains on all its blocks
    assert tr.restore_epoch is None
    a = tr.sequential_transfer_learn(pretraining_epochs=5, training_epochs=10, posttraining_epochs=5)
    # epochs: subject 400: 10 ; subject 401: 5 (subnet only) + 10 + 5
    assert tr.restore_epoch == 30
    assert a["training"].decoder_word_error_rates.shape == (3,)
    tr._restore_epoch = None
    assert tr.restore_epoch == 30                                       # discovered from model.ckpt-<epoch>.index
    # the pre-training phase of subject 401 (epochs 10 -> 15) must not have touched the shared tensors
    shared = "seq2seq/decoder_rnn/multi_rnn_cell/cell_0/lstm_cell/kernel"
    w10 = tr.net.get_weights_as_numpy_array(shared, 10)
    w15 = tr.net.get_weights_as_numpy_array(shared, 15)
    w30 = tr.net.get_weights_as_numpy_array(shared, 30)
    assert np.array_equal(w10, w15) and not np.array_equal(w15, w30)
    # recover_model_sizes reads our checkpoints with the reference's parsing rules
    layer_sizes, data_sizes, strides, EMA = tr.recover_model_sizes()
    assert layer_sizes["encoder_embedding"] == [5] and layer_sizes["encoder_rnn"] == [8, 8]
    assert layer_sizes["decoder_embedding"] == [6] and layer_sizes["decoder_rnn"] == [16]
    assert layer_sizes["decoder_projection"] == []
    assert strides[401] == [4] and data_sizes[401]["encoder_inputs"] == 6
    assert data_sizes[None]["decoder_targets"] == len(VOCAB) and EMA == 1
    res = tr.assess_saved_model()
    assert abs(res["training"].word_error_rate - a["training"].decoder_word_error_rates[-1]) < 1e-9
    # online predictor: one utterance at a time -> sentence
    predict = tr.construct_online_predictor()
    x = tr.net._load_partition(tr.ecog_subjects[-1], "validation")[0][0]
    s1 = predict(x)
    assert isinstance(s1, str) and s1 == predict(x)


def test_parallel_transfer_learn_and_resume(tmp_path, emu_lib):
    tr = _trainer(tmp_path, emu_lib, N_epochs=10, assessment_epoch_interval=5)
    a = tr.parallel_transfer_learn()
    assert a["training"].decoder_word_error_rates.shape == (2,)
    assert os.path.exists(os.path.join(tr.checkpoint_dir, "model.ckpt-10.index"))
    shapes = prm.variable_to_shape_map(tr.net.checkpoint_path, 10)
    assert "seq2seq/subnet_400/encoder_embedding_6_5_0/weights" in shapes     # both private subnets live in one model
    assert "seq2seq/subnet_401/encoder_embedding_6_5_0/weights" in shapes
    tr.parallel_transfer_learn(RESUME=True)                                   # last subject only, from the latest epoch
    assert tr.restore_epoch == 20


def test_get_saliencies_and_projection_sizes(tmp_path, emu_lib):
    """get_saliencies (trainers.py:703-732) zeroes every *_targets penalty but the one named by contrib_method and restores
    them afterwards; recover_model_sizes splits an 'encoder_1_projection' into hidden sizes + the (transposed) output size."""
    tr = _trainer(tmp_path, emu_lib, ids=(400,), subject_kw=dict(encoder_targets="audio_sequence", encoder_targets_features=4,
                                                                   encoder_targets_penalty_scale=0.3),
                  N_epochs=5, assessment_epoch_interval=5)
    tr.net.layer_sizes = dict(tr.net.layer_sizes, encoder_1_projection=[7])
    tr.parallel_transfer_learn()
    layer_sizes, data_sizes, _, _ = tr.recover_model_sizes()
    assert layer_sizes["encoder_1_projection"] == [7] and data_sizes[None]["encoder_1_targets"] == 4
    assert layer_sizes["decoder_projection"] == [] and data_sizes[None]["decoder_targets"] == len(VOCAB)
    dec = tr.get_saliencies("decoder_saliency_map")
    aux = tr.get_saliencies("encoder_1_saliency_map")
    assert dec.shape == (6,) and aux.shape == (6,) and not np.allclose(dec, aux)
    mans = tr.ecog_subjects[-1].data_manifests
    assert mans["encoder_1_targets"].penalty_scale == 0.3 and mans["decoder_targets"].penalty_scale == 1.0


def test_introspection_callers(tmp_path, emu_lib):
    """_retrieve_layer_weights / get_encoder_embedding (trainers.py:680-751), get_internal_activations (:757-859) against the
    oracle's encoder on the restored EMA weights, tf_record_to_numpy_data (:861-922)."""
    import torch
    from oracle import seq2seq_oracle as O
    tr = _trainer(tmp_path, emu_lib, ids=(400,), subject_kw=dict(encoder_targets="audio_sequence", encoder_targets_features=4),
                  N_epochs=5, assessment_epoch_interval=5)
    tr.net.layer_sizes = dict(tr.net.layer_sizes, encoder_1_projection=[7])
    tr.parallel_transfer_learn()
    s = tr.ecog_subjects[-1]
    Wc = tr.get_encoder_embedding()
    assert Wc.shape == (1, 4, 6, 5)
    assert np.array_equal(Wc, tr._retrieve_layer_weights("encoder_embedding"))
    assert tr._retrieve_layer_weights("decoder_embedding").shape == (len(VOCAB), 6)
    acts = tr.get_internal_activations()
    ex = tr.net._load_partition(s, "validation")
    N, T = len(ex), max(e[0].shape[0] for e in ex)
    T2 = -(-T // 4)
    assert acts.convolved_inputs.shape == (N, T2, 5) and acts.reversed_inputs.shape == (N, T, 6)
    assert acts.final_RNN_state.shape == (2, 1, N, 16) and acts.decimated_reversed_targets.shape == (N, T2, 4)
    # oracle on the EMA weights of the checkpoint
    shapes = prm.variable_to_shape_map(tr.net.checkpoint_path, tr.restore_epoch)
    ema = "/ExponentialMovingAverage"
    P = {k[:-len(ema)]: torch.from_numpy(tr.net.get_weights_as_numpy_array(k, tr.restore_epoch)) for k in shapes if k.endswith(ema)}
    ocfg = O.OracleConfig(subnet_ids=(400,), subnet_C=(6,), subnet_W=(4,), E=5, H=(8, 8), D=6, Hd=16, V=len(VOCAB),
                          aux_layer=1, aux_hidden=7, aux_F=4)
    x = np.zeros((N, T, 6), np.float32)
    for i, e in enumerate(ex):
        x[i, :e[0].shape[0]] = e[0]
    ref = O.encoder(ocfg, P, torch.from_numpy(x), None, 0)
    assert np.allclose(acts.convolved_inputs, ref["conv_out"].numpy(), atol=1e-5)
    assert np.allclose(acts.final_RNN_state[1, 0], ref["final_h"].numpy(), atol=1e-5)
    assert np.allclose(acts.final_RNN_state[0, 0], ref["final_c"].numpy(), atol=1e-5)
    assert np.array_equal(acts.lengths, ref["lens"].numpy())
    assert np.array_equal(acts.reversed_inputs, O.reverse_within_length(torch.from_numpy(x), ref["lens"]).numpy())
    aux = np.zeros((N, T, 4), np.float32)
    for i, e in enumerate(ex):
        aux[i, :e[2].shape[0]] = e[2]
    assert np.array_equal(acts.decimated_reversed_targets, O.prepare_encoder_targets(torch.from_numpy(aux), ref["lens"], 4).numpy())
    # raw records
    recs = list(tr.tf_record_to_numpy_data(400, 3))
    assert len(recs) == 6 and recs[0]["encoder_inputs"].shape[1] == 6 and recs[0]["encoder_1_targets"].shape[1] == 4
    assert recs[0]["decoder_targets"].dtype == object and recs[0]["decoder_targets"][0, 0].decode().endswith("_")
    with np.testing.assert_raises(ValueError):
        next(tr.tf_record_to_numpy_data(999, 3))


def test_recover_model_sizes_with_hidden_decoder_projection(tmp_path, emu_lib):
    """A checkpoint of a model with layer_sizes['decoder_projection'] = [9] parses back to that list and the vocabulary
    size under the reference's rules (numbered '<x>_projection' layers, the last one transposed: trainers.py:488-520)."""
    tr = _trainer(tmp_path, emu_lib, ids=(400,), N_epochs=5, assessment_epoch_interval=5)
    tr.net.layer_sizes = dict(tr.net.layer_sizes, decoder_projection=[9])
    tr.parallel_transfer_learn()
    layer_sizes, data_sizes, _, _ = tr.recover_model_sizes()
    assert layer_sizes["decoder_projection"] == [9] and data_sizes[None]["decoder_targets"] == len(VOCAB)
    assert layer_sizes["decoder_rnn"] == [16]
