#!/usr/bin/env python
"""BASELINE.json config 4: the multi-subject transfer-learning schedule of trainers.py (pretrain the new subject's private conv
with the shared network frozen, then train everything) on N GPUs, data-parallel (every minibatch sharded over the ranks, one
NCCL all-reduce per step), at the config-2 network geometry with two synthetic subjects of different electrode counts.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/config4_run.py --out F.json

Reference anchor: /root/reference/ecog2txt/trainers.py:329-374 (sequential_transfer_learn)."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--blocks", type=int, default=2, help="training blocks per subject")
    ap.add_argument("--per-block", type=int, default=256, help="utterances per block")
    ap.add_argument("--pre", type=int, default=2)
    ap.add_argument("--train", type=int, default=3)
    ap.add_argument("--post", type=int, default=1)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ecog2txt_b200 import MultiSubjectTrainer
    from ecog2txt_b200.subjects import make_synthetic_subject
    from ecog2txt_b200.synthetic import load_vocab
    tmp = os.path.join(tempfile.gettempdir(), "e2t_config4")          # one copy of the records, written by rank 0
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(tmp)
    if world > 1:
        dist.barrier()
    vocab = load_vocab(size=1806)
    ids, chans = (400, 401), (256, 128)
    subjects = []
    for i, (sid, c) in enumerate(zip(ids, chans)):
        s = make_synthetic_subject(sid, vocab, os.path.join(tmp, "tf"), n_train_blocks=a.blocks, n_valid_blocks=1,
                                   utterances_per_block=a.per_block, T=400, C=c, n_sentences=50, ragged=False, seed=i)
        s.data_generator.corpus.max_words = 10
        if rank == 0:
            s.write_tf_records_maybe()
        subjects.append(s)
    if world > 1:
        dist.barrier()
    base = {"layer_sizes": {"encoder_embedding": [100], "encoder_rnn": [400, 400, 400], "decoder_embedding": [150],
                            "decoder_rnn": [800], "decoder_projection": []},
            "FF_dropout": 0.1, "RNN_dropout": 0.5, "TEMPORALLY_CONVOLVE": True, "EMA_decay": 0.99, "N_epochs": 4,
            "beam_width": 1, "temperature": 0.384, "assessment_epoch_interval": 100, "token_type": "word_sequence"}
    manifest = {sid: dict(base) for sid in ids}
    tr = MultiSubjectTrainer(manifest, list(ids), checkpoint_dir=os.path.join(tmp, f"ckpt{rank if world == 1 else ''}"),
                             SN_kwargs=dict(N_cases=256, max_hyp_length=12, learning_rate=5e-4), VERBOSE=False, subjects=subjects)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.sequential_transfer_learn(pretraining_epochs=a.pre, training_epochs=a.train, posttraining_epochs=a.post)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    n_train = a.blocks * a.per_block
    # epochs run: subject 0 trains `train`; subject 1 pre-trains `pre`, then trains `train`; post-training on both
    utt = n_train * (a.train + a.pre + a.train + 2 * a.post)
    res = {"config": "BASELINE.json configs[3]: sequential_transfer_learn, 2 synthetic subjects (256 / 128 electrodes), config-2 geometry",
           "world": world, "global_minibatch": 256, "per_rank_minibatch": 256 // world, "training_utterances_per_subject": n_train,
           "epochs": {"pretraining": a.pre, "training": a.train, "posttraining": a.post}, "utterances_trained": utt,
           "wall_s_incl_assessments_and_checkpoints": dt, "utt_per_s_wall": utt / dt,
           "persistent_rnn_launches": int(tr.net._engine.counter("persistent_rnn_launches")), "restore_epoch": int(tr.restore_epoch)}
    if rank == 0:
        print(json.dumps(res))
        if a.out:
            json.dump(res, open(a.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
