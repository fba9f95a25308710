#!/usr/bin/env python
"""N-rank data-parallel equality check on real GPUs (SURVEY.md section 4 "distributed").

torchrun --nproc-per-node N tools/dp_equality.py --steps K --out res.json
Every rank trains K steps on its shard of the same global minibatches (config-2 geometry, tensor-core path, dropout off
so that the 1-GPU run on the concatenated batch is the same function), ONE NCCL all-reduce per step (gradients + token
count), Adam+EMA with the count read on the device.  Checks: (1) all ranks hold BIT-identical E2T_VALUE / E2T_EMA buffers
after K steps; (2) rank 0 re-runs the K steps alone on the whole batches: the weights agree to reduction-order tolerance.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--global-batch", type=int, default=256)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from ecog2txt_b200 import Engine, EngineConfig, _lib
    from ecog2txt_b200.dist import allreduce_step, flat_tensor, shard_range
    from ecog2txt_b200.params import init_engine
    from ecog2txt_b200.synthetic import SyntheticCorpus, load_vocab
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    GEO = dict(subnet_ids=(400,), subnet_C=(256,), subnet_W=(12,), E=100, H=(400, 400, 400), D=150, Hd=800, V=1806)
    B = a.global_batch
    corpus = SyntheticCorpus(load_vocab(size=1806), T=400, C=256, seed=0)
    batches = [corpus.batch(B, seed=s, L=11) for s in range(a.steps)]

    def run(per_rank, reduce):
        eng = Engine(EngineConfig(**GEO, max_B=B, max_T=400, max_L=12, device=local, lr=1e-3))
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
        init_engine(eng, seed=1)
        gc = flat_tensor(eng, _lib.GRAD_AND_COUNT)
        g0 = None
        for s, b in enumerate(batches):
            lo, hi = shard_range(B, rank, world) if per_rank else (0, B)
            x = np.ascontiguousarray(b["encoder_inputs"][lo:hi])
            y = np.ascontiguousarray(b["decoder_targets"][lo:hi])
            eng.train_step_grads(x, None, y, seed=s, want_loss=False)
            if reduce:
                allreduce_step(gc)
            if s == 0:
                g0 = gc.clone()          # first-step gradient (+ token count): the quantity the all-reduce must reproduce
            eng.adam_ema_step_dev(None)
        w = flat_tensor(eng, _lib.VALUE).clone()
        e = flat_tensor(eng, _lib.EMA).clone()
        eng.close()
        return w, e, g0

    w, e, g = run(True, True)
    ws = [torch.empty_like(w) for _ in range(world)]
    es = [torch.empty_like(e) for _ in range(world)]
    dist.all_gather(ws, w)
    dist.all_gather(es, e)
    identical = all(torch.equal(ws[0], t) for t in ws[1:]) and all(torch.equal(es[0], t) for t in es[1:])
    res = None
    if rank == 0:
        w1, e1, g1 = run(False, False)
        d = float((w - w1).abs().max() / w1.abs().max())
        res = {"world": world, "steps": a.steps, "global_batch": B, "ranks_bit_identical": bool(identical),
               # all-reduced gradient of step 0 vs the 1-GPU gradient of the concatenated batch, relative to its largest entry
               "grad_max_rel_diff_vs_1gpu": float((g[:-4] - g1[:-4]).abs().max() / g1[:-4].abs().max()),
               "token_count": [float(g[-4]), float(g1[-4])],
               # weights after K Adam steps (Adam's m / sqrt(v) turns a sign flip of a near-zero gradient entry into a
               # full-size step, so the max is loose by construction; the L2 figure is the meaningful one)
               "weights_l2_rel_diff_vs_1gpu": float((w - w1).norm() / w1.norm()),
               "max_rel_diff_vs_1gpu": d, "max_rel_diff_ema_vs_1gpu": float((e - e1).abs().max() / e1.abs().max()),
               "geometry": "config 2 (C=256, 3x400, 800, V=1806), T=400, L=11, tensor-core backend, dropout off"}
        print(json.dumps(res), flush=True)
        if a.out:
            json.dump(res, open(a.out, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
