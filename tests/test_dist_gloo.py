"""CPU, world_size 2, gloo: the data-parallel plumbing of SURVEY.md section 8e -- shard the minibatch, ONE all-reduce of
the engine's flat gradient buffer (+ the token count), identical Adam+EMA on every rank -- must reproduce the
single-process step on the whole batch.  Engines run on the kernel-emulation build (TEST-ONLY)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_common as pc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import ctypes
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import __graft_entry__ as ge
    from ecog2txt_b200 import _lib
    from ecog2txt_b200.dist import allreduce_grads, flat_tensor, shard_range
    from oracle import seq2seq_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _lib.bind(ctypes.CDLL(ge.EMU_LIB))
    geo = pc.TINY
    ocfg = O.OracleConfig(**geo)
    P = pc.make_params(ocfg)
    B, T, L = 7, 19, 5          # 7 utterances over 2 ranks: shards of 4 and 3
    x, lens, y = pc.make_batch(ocfg, B, T, L)
    eng = pc.engine_for(geo, lib, B, T, L, gemm_backend="simt")
    eng.set_all({k: v.numpy() for k, v in P.items()})
    lo, hi = shard_range(B, rank, world)
    loss, ntok = eng.train_step_grads(np.ascontiguousarray(x[lo:hi]), None, np.ascontiguousarray(y[lo:hi]), seed=0)
    g = flat_tensor(eng, _lib.GRAD)
    if os.environ.get("E2T_TEST_BUCKETED") == "1":
        # bucket-by-bucket all-reduce (sequential here: no streams in the emulation) + token count kept in "device" memory
        from ecog2txt_b200.dist import BucketedAllReduce
        ar = BucketedAllReduce(eng)
        eng.train_step_grads(np.ascontiguousarray(x[lo:hi]), None, np.ascontiguousarray(y[lo:hi]), seed=0)   # bucketing now on
        assert len(eng.grad_buckets()) == 1 + len(geo["H"]) + 1
        nt = torch.tensor([float(ntok)], dtype=torch.float32)
        ar.reduce_async(nt)
        ntok_g = float(nt.item())
        eng.adam_ema_step_dev(nt.numpy())
    else:
        # the product path (SequenceNetwork.fit, bench.py): ONE collective -- gradients with the token count in the tail --
        # and the optimiser reads the all-reduced count on the "device"
        from ecog2txt_b200.dist import allreduce_step
        gc = flat_tensor(eng, _lib.GRAD_AND_COUNT)
        assert gc.numel() == g.numel() + 4 and float(gc[-4]) == float(ntok) and gc.data_ptr() == g.data_ptr()
        allreduce_step(gc)
        ntok_g = float(gc[-4])
        eng.adam_ema_step_dev(None)
        acc = eng.read_loss_accumulators(reset=True)
        assert acc[1] == ntok and abs(acc[0] - loss) <= 1e-6 * abs(loss)
        assert eng.read_loss_accumulators()[1] == 0
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ntok=ntok_g, lo=lo, hi=hi,
             **{k.replace("/", "|"): v for k, v in eng.get_all(_lib.VALUE).items()},
             **{"G|" + k.replace("/", "|"): v for k, v in eng.get_all(_lib.GRAD).items()})
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("bucketed", ["0", "1"])
def test_two_rank_step_equals_single_process(tmp_path, emu_lib, bucketed, monkeypatch):
    monkeypatch.setenv("E2T_TEST_BUCKETED", bucketed)
    from ecog2txt_b200 import _lib
    from ecog2txt_b200.dist import shard_range
    from oracle import seq2seq_oracle as O
    assert [shard_range(7, r, 2) for r in range(2)] == [(0, 4), (4, 7)]
    assert [shard_range(5, r, 8) for r in range(8)] == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 5), (5, 5), (5, 5)]
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # single process, whole batch
    geo = pc.TINY
    ocfg = O.OracleConfig(**geo)
    P = pc.make_params(ocfg)
    x, lens, y = pc.make_batch(ocfg, 7, 19, 5)
    eng = pc.engine_for(geo, emu_lib, 7, 19, 5, gemm_backend="simt")
    eng.set_all({k: v.numpy() for k, v in P.items()})
    loss, ntok = eng.train_step_grads(x, None, y, seed=0)
    G = eng.get_all(_lib.GRAD)
    eng.adam_ema_step(1.0 / ntok)
    W = eng.get_all(_lib.VALUE)
    assert float(r0["ntok"]) == float(r1["ntok"]) == float(ntok)
    for k in W:
        kk = k.replace("/", "|")
        assert np.array_equal(r0[kk], r1[kk]), k                      # replicas stay bit-identical
        assert np.allclose(r0["G|" + kk], G[k], rtol=1e-5, atol=1e-6), k   # summed shard grads == full-batch grads
        assert np.allclose(r0[kk], W[k], rtol=1e-5, atol=1e-6), k
    eng.close()


def _fit_worker(rank, world, port, tmp):
    import ctypes
    import pickle
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import __graft_entry__ as ge
    from ecog2txt_b200 import SequenceNetwork, _lib
    from test_sequence_network import MANIFEST, _subject
    from pathlib import Path
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _lib.bind(ctypes.CDLL(ge.EMU_LIB))
    s = _subject(Path(tmp))                      # TFRecords were written by the parent
    m = dict(MANIFEST, N_epochs=10, assessment_epoch_interval=10)
    net = SequenceNetwork(m, VERBOSE=False, N_cases=4, max_hyp_length=5, learning_rate=2e-2, lib=lib, gemm_backend="simt")
    net.checkpoint_path = os.path.join(tmp, "ckpt", "model.ckpt")
    a = net.fit([s])
    res = net.restore_and_assess([s], 10)        # sharded decode: every rank ends up with every hypothesis
    with open(os.path.join(tmp, f"fit_rank{rank}.pkl"), "wb") as f:
        pickle.dump({"wer": a["training"].decoder_word_error_rates, "acc": a["training"].decoder_accuracies,
                     "hyps": res["training"].hypotheses, "refs": res["training"].references,
                     "vwer": res["validation"].word_error_rate}, f)
    dist.destroy_process_group()


def test_two_rank_fit_and_sharded_assessment(tmp_path, emu_lib):
    """SequenceNetwork.fit under torch.distributed (world 2, gloo): every minibatch sharded over the ranks, one all-reduce
    per step; assessment decodes every world-th minibatch per rank and gathers the strings.  Both ranks must report the
    same numbers, and the sharded assessment must equal a single-process assessment of the same checkpoint."""
    import pickle
    from ecog2txt_b200 import SequenceNetwork
    from test_sequence_network import MANIFEST, _subject
    s = _subject(tmp_path)
    s.write_tf_records_maybe()
    port = _free_port()
    mp.spawn(_fit_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (pickle.load(open(tmp_path / f"fit_rank{r}.pkl", "rb")) for r in range(2))
    assert np.array_equal(r0["wer"], r1["wer"]) and np.array_equal(r0["acc"], r1["acc"])
    assert r0["hyps"] == r1["hyps"] and r0["refs"] == r1["refs"] and r0["vwer"] == r1["vwer"]
    m = dict(MANIFEST, N_epochs=10, assessment_epoch_interval=10)
    net = SequenceNetwork(m, VERBOSE=False, N_cases=4, max_hyp_length=5, lib=emu_lib, gemm_backend="simt")
    net.checkpoint_path = str(tmp_path / "ckpt" / "model.ckpt")
    res = net.restore_and_assess([s], 10)
    assert res["training"].hypotheses == r0["hyps"] and res["training"].references == r0["refs"]
    assert res["validation"].word_error_rate == r0["vwer"]
