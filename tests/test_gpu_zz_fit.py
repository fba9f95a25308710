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


def test_fit_device_cache_equals_staged_pipeline(gpu_lib, tmp_path):
    """fit() with the training set resident on the device (minibatches gathered there, only indices cross PCIe) must train
    exactly like the staged host pipeline (native padding threads -> page-locked ring -> copy stream): same minibatches,
    same dropout seeds -> the same weights up to the one atomic reduction (embedding scatter)."""
    from ecog2txt_b200 import SequenceNetwork, _lib
    from ecog2txt_b200.subjects import make_synthetic_subject
    vocab = ["<pad>", "<EOS>", "<OOV>"] + [f"w{i:02d}_" for i in range(37)]
    s = make_synthetic_subject(400, vocab, str(tmp_path / "tf"), n_train_blocks=2, n_valid_blocks=1,
                               utterances_per_block=32, T=96, C=64, n_sentences=10, ragged=True, seed=0)
    s.data_generator.corpus.max_words = 6
    s.write_tf_records_maybe()
    manifest = {"layer_sizes": {"encoder_embedding": [32], "encoder_rnn": [64, 64], "decoder_embedding": [24],
                                "decoder_rnn": [128], "decoder_projection": []},
                "FF_dropout": 0.1, "RNN_dropout": 0.3, "TEMPORALLY_CONVOLVE": True, "EMA_decay": 0.9, "N_epochs": 3,
                "beam_width": 1, "temperature": 0.384, "assessment_epoch_interval": 3}
    out = []
    for cache_bytes in (16 << 30, 0):
        net = SequenceNetwork(manifest, VERBOSE=False, N_cases=32, max_hyp_length=8, learning_rate=5e-3,
                              device_cache_bytes=cache_bytes)
        net.checkpoint_path = None
        net.fit([s])
        out.append(net._engine.get_all(_lib.VALUE))
    for k in out[0]:
        np.testing.assert_allclose(out[0][k], out[1][k], rtol=2e-4, atol=2e-6, err_msg=k)


def _gpu_trainer(tmp_path, ids=(400, 401), C=(64, 48), **sn):
    from ecog2txt_b200 import MultiSubjectTrainer
    from ecog2txt_b200.subjects import make_synthetic_subject
    vocab = ["<pad>", "<EOS>", "<OOV>"] + [f"w{i:02d}_" for i in range(37)]
    subjects = []
    for i, (sid, c) in enumerate(zip(ids, C)):
        s = make_synthetic_subject(sid, vocab, str(tmp_path / "tf"), n_train_blocks=2, n_valid_blocks=1,
                                   utterances_per_block=32, T=96, C=c, n_sentences=10, ragged=True, seed=i)
        s.data_generator.corpus.max_words = 6
        s.write_tf_records_maybe()
        subjects.append(s)
    base = {"layer_sizes": {"encoder_embedding": [32], "encoder_rnn": [64, 64], "decoder_embedding": [24],
                            "decoder_rnn": [128], "decoder_projection": []},
            "FF_dropout": 0.1, "RNN_dropout": 0.3, "TEMPORALLY_CONVOLVE": True, "EMA_decay": 0.9, "N_epochs": 4,
            "beam_width": 1, "temperature": 0.384, "assessment_epoch_interval": 2, "token_type": "word_sequence"}
    manifest = {sid: dict(base) for sid in ids}
    sn_kwargs = dict(N_cases=32, max_hyp_length=8, learning_rate=5e-3, **sn)
    return MultiSubjectTrainer(manifest, list(ids), checkpoint_dir=str(tmp_path / "ckpt"), SN_kwargs=sn_kwargs, VERBOSE=False,
                               subjects=subjects)


def test_transfer_learning_schedules_on_gpu(gpu_lib, tmp_path):
    """N2 on hardware (/root/reference/ecog2txt/trainers.py:329-374): sequential transfer learning of two synthetic subjects
    with different electrode counts on the CUDA path.  The pre-training phase of the second subject trains
    'seq2seq/subnet' only: every shared tensor (value, EMA shadow, Adam slots) must come out of it BIT-identical, while the
    new subject's conv moves; the following full-model phase moves the shared tensors again.  Then the joint schedule
    (parallel_transfer_learn): a minibatch of one subject must leave the other subject's private tensors untouched."""
    from ecog2txt_b200 import _lib, params as prm
    tr = _gpu_trainer(tmp_path)
    tr.net.max_to_keep = None        # the test reads the epoch-6 checkpoint back after later ones were written
    tr.sequential_transfer_learn(pretraining_epochs=4, training_epochs=6, posttraining_epochs=2)
    assert tr.restore_epoch == 6 + 4 + 8
    assert tr.net._engine.emulated or tr.net._engine.counter("tcgen05_launches") > 0
    z6, z10, z18 = (np.load(f"{tr.net.checkpoint_path}-{e}.npz") for e in (6, 10, 18))
    shared = [k for k in z10.files if k.startswith("seq2seq/") and not k.startswith("seq2seq/subnet")]
    assert len(shared) > 20
    for k in shared:
        assert np.array_equal(z6[k], z10[k]), k                  # frozen during 'seq2seq/subnet' pre-training: bit-identical
    kern = "seq2seq/decoder_rnn/multi_rnn_cell/cell_0/lstm_cell/kernel"
    assert not np.array_equal(z10[kern], z18[kern])              # and trained again afterwards
    conv401 = "seq2seq/subnet_401/encoder_embedding_48_32_0/weights"
    assert conv401 in z10.files and "seq2seq/subnet_400/encoder_embedding_64_32_0/weights" not in z10.files
    fresh = prm.glorot_init({conv401: z10[conv401].shape}, tr.net.seed)[conv401]
    assert not np.array_equal(z10[conv401], fresh)               # the private conv did train
    res = tr.assess_saved_model()
    assert 0.0 <= res["validation"].word_error_rate <= 1.5
    # joint schedule: both subjects in one model
    tr2 = _gpu_trainer(tmp_path / "joint", N_epochs=2, assessment_epoch_interval=2)
    a = tr2.parallel_transfer_learn()
    assert a["validation"].decoder_confusions.shape == (40, 40) and a["validation"].decoder_confusions.sum() > 0
    eng = tr2.net._engine
    before = {k: (eng.get(k, _lib.VALUE), eng.get(k, _lib.ADAM_M)) for k in eng.tensors() if "subnet_400" in k}
    ex = tr2.net._load_partition(tr2.ecog_subjects[1], "training")
    x, y = tr2.net._batch(ex, np.arange(32), eng.cfg.max_T, eng.cfg.max_L, eng.cfg.pad_id)
    _, ntok = eng.train_step_grads(x, None, y, subnet=1, seed=5)
    eng.adam_ema_step(1.0 / ntok, subnet=1)
    for k, (v, m) in before.items():
        assert np.array_equal(eng.get(k, _lib.VALUE), v) and np.array_equal(eng.get(k, _lib.ADAM_M), m), k


def test_data_parallel_ranks_stay_identical_on_gpus(gpu_lib, tmp_path):
    """SURVEY.md section 4 ("distributed"): K data-parallel steps over NCCL -- every rank must hold bit-identical weights,
    equal (to reduction-order tolerance) to the 1-GPU run on the concatenated batch.  Needs >= 2 GPUs (gpurun --gpus 2);
    the result of the most recent multi-GPU run is committed as profiles/r2_dp_equality_2gpu.json."""
    import json
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2); see profiles/r2_dp_equality_*.json for the recorded run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    world = 2 if n < 4 else 4
    out = tmp_path / "dp.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(root, "tools", "dp_equality.py"), "--steps", "5", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    assert res["ranks_bit_identical"] and res["world"] == world
    assert res["token_count"][0] == res["token_count"][1]
    assert res["grad_max_rel_diff_vs_1gpu"] <= 1e-4, res
    assert res["weights_l2_rel_diff_vs_1gpu"] <= 1e-3, res
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(root, "gpurun_out", f"r2_dp_equality_{world}gpu.json"), "w"), indent=1)
