"""-m gpu: the SequenceNetwork drop-in end to end on a real B200 (TFRecords -> fit with the staged input pipeline ->
checkpoint -> restore_and_assess).  Kept in its own file, after test_gpu_parity.py in collection order: it exercises the host-side
pipeline on top of kernels whose parity the other file has already established."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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
