#!/usr/bin/env python
"""bench.py -- training throughput (utterances/s) and greedy-decode latency of the seq2seq hot path.

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N>1) prints ONE JSON line
on rank 0.  A "step" = one full training step (forward + backward + all-reduce (N>1) + Adam + EMA)
over one synthetic batch of the config-2 workload of BASELINE.json: T=400 frames x C=256 channels,
3x400 BiLSTM encoder, 800 LSTM decoder, V=1806, per-GPU batch 256 (weak scaling).
  value : whole-job utterances/s with the batch pool already resident in HBM (pool > L2)
  e2e   : same metric through the public API with HOST (pinned) buffers: H2D of x/y and D2H of the loss
          inside the timed region, every step
  roofline     : recurrent-step kernels (the dominant category), algorithmic FLOPs / CUDA-event time
  cpu_baseline : the oracle (torch CPU port; the reference's TF1.15 path cannot run here) on a bounded sample
`--impl reference` times that CPU port alone on the same config / metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

GEO = dict(subnet_ids=(400,), subnet_C=(256,), subnet_W=(12,), E=100, H=(400, 400, 400), D=150, Hd=800, V=1806)
T_FRAMES, L_TGT = 400, 11
FF_DROPOUT, RNN_DROPOUT = 0.1, 0.5   # mochastar_word_sequence.yaml:4,11
METRIC, UNIT = "train_utterances_per_sec", "utt/s"


def flops_per_utt(L):
    """BASELINE.md section 4 (multiply-add = 2 FLOP)."""
    W, C, E, H, D, Hd, V, T2 = 12, 256, 100, 400, 150, 800, 1806, 34
    conv = 2 * T2 * W * C * E
    enc = 2 * 2 * T2 * (E + H) * 4 * H + 2 * 2 * 2 * T2 * (3 * H) * 4 * H
    dec = (2 * (D + Hd) * 4 * Hd + 2 * Hd * V) * L
    fwd = conv + enc + dec
    rec_fwd = 3 * 2 * 2 * T2 * H * 4 * H + L * 2 * Hd * 4 * Hd     # h Wh products only
    return dict(forward=fwd, train=3 * fwd, recurrent_train=2 * rec_fwd)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


WORKLOAD = "config2: T=400 C=256 3x400 BiLSTM + 800 LSTM decoder V=1806, L=11, dropout .1/.5, Adam+EMA"


def _host_threads():
    # every host thread this process may use -- explicitly, because torchrun exports OMP_NUM_THREADS=1 to its workers and
    # the CPU arm would otherwise run single-threaded under the N > 1 launch
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port_utt_per_s(B, steps, threads=None, warmup=1, mode="speed"):
    """One full training step (forward, backward, Adam+EMA) of the CPU oracle on the host cores, masks / dropout included.
    mode "speed": oracle/speed_mode.py -- recurrences on torch.nn.LSTM (oneDNN), the fastest CPU restatement available here
    (BASELINE.md section 3); mode "loops": oracle/seq2seq_oracle.py -- one LSTM step per Python iteration + autograd."""
    import torch
    from oracle import seq2seq_oracle as O
    from ecog2txt_b200.synthetic import SyntheticCorpus, load_vocab
    torch.set_num_threads(threads or _host_threads())
    ocfg = O.OracleConfig(**GEO)
    P = O.init_params(ocfg, 1)
    corpus = SyntheticCorpus(load_vocab(size=GEO["V"]), T=T_FRAMES, C=256, seed=0)
    if mode == "speed":
        from oracle import speed_mode as S
        trainer = S.SpeedTrainer(ocfg, P)
    else:
        opt = O.AdamEMA(ocfg, P)
    times = []
    for s in range(warmup + steps):
        b = corpus.batch(B, seed=s, L=L_TGT)
        x, y = torch.from_numpy(b["encoder_inputs"]), torch.from_numpy(b["decoder_targets"]).long()
        t0 = time.perf_counter()
        if mode == "speed":
            trainer.step(x, y, FF_DROPOUT, RNN_DROPOUT)
        else:
            masks = O.make_masks(ocfg, s, B, 34, L_TGT, FF_DROPOUT, RNN_DROPOUT, torch.float32)
            _, ntok, g, _ = O.loss_and_grads(ocfg, P, x, None, y, masks=masks)
            opt.step(P, g, 1.0 / ntok)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    return B / float(np.mean(times)), torch.get_num_threads(), float(np.mean(times))


def run_reference(args, rank, world):
    """The reference arm: the reference's own TF1.15 path cannot be installed here, so this times the CPU oracle
    in speed mode -- same workload, metric, steps and warm-up as the GPU arm, every host thread."""
    if rank != 0:
        return
    B = args.ref_batch
    warm = max(args.warmup, 3)
    v, cores, dt = cpu_port_utt_per_s(B, args.steps, warmup=warm, mode="speed")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": B},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} training steps of {B} utterances (oracle speed mode: torch CPU fp32, "
                                   "torch.nn.LSTM / oneDNN; the reference's TF1.15 + machine_learning path is not installable here)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def measure_cublas_tflops(torch, dtype, tf32, n=8192, iters=20):
    """Dense n^3 GEMM through cuBLAS (torch.matmul), random operands, CUDA events -- a measured tensor peak for the roofline."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    try:
        a = torch.rand(n, n, device="cuda", dtype=dtype) - 0.5
        b = torch.rand(n, n, device="cuda", dtype=dtype) - 0.5
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        return 2.0 * n ** 3 * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="utterances per GPU per step (weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch utterances per GPU (global batch grows with N); strong: --global-batch utterances per "
                         "step in total, sharded over the N GPUs (north_star: the per-subject minibatch sharded across the GPUs)")
    ap.add_argument("--global-batch", type=int, default=256, help="utterances per step over all GPUs (strong scaling)")
    ap.add_argument("--ref-batch", type=int, default=256, help="utterances per CPU step (the GPU arm's per-GPU batch)")
    ap.add_argument("--backend", default="auto")
    ap.add_argument("--cpu-baseline-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", default=None, help="write the per-kernel event-time breakdown of a step to this file")
    ap.add_argument("--overlap-allreduce", default="auto", nargs="?", const="on", choices=["auto", "on", "off"],
                    help="N > 1: bucketed all-reduce on a side stream, overlapped with the backward pass, instead of one flat "
                         "all-reduce after it.  Measured: N = 2 113.7 k vs 114.3 k utt/s flat (the per-bucket flushes cost what the "
                         "overlap hides); N = 8 before the side stream took the weight-gradient GEMMs 572.2 k vs 561.9 k, with it "
                         "618.8 k vs 620.8 k flat (profiles/r2_bench_8gpu_*.json: the buckets now complete on the side stream, late) "
                         "-- auto = flat")
    ap.add_argument("--no-decode", action="store_true", help="skip the greedy / beam-8 decode legs")
    ap.add_argument("--decode-utterances", type=int, default=10000, help="utterances decoded per leg, sharded over the ranks")
    ap.add_argument("--watchdog", type=int, default=int(os.environ.get("E2T_BENCH_WATCHDOG", "900")),
                    help="seconds after which a run that has not finished dumps every thread's Python stack to stderr and "
                         "exits non-zero (0 = off): a wedged GPU or collective must not hang the caller")
    args = ap.parse_args()
    if args.watchdog > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from ecog2txt_b200 import Engine, EngineConfig, _lib
    from ecog2txt_b200.params import init_engine
    from ecog2txt_b200.synthetic import SyntheticCorpus, load_vocab
    from ecog2txt_b200.dist import flat_tensor

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU port"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.scaling == "strong":
        assert args.global_batch % world == 0, "--global-batch must be divisible by the number of GPUs"
        B = args.global_batch // world
    else:
        B = args.batch
    eng = Engine(EngineConfig(**GEO, max_B=B, max_T=T_FRAMES, max_L=20, max_beam=8, ff_dropout=FF_DROPOUT,
                              rnn_dropout=RNN_DROPOUT, gemm_backend=args.backend, device=local))
    # everything runs on ONE high-priority torch stream (the library's kernels, NCCL, the timing events): the library hands the
    # weight-gradient GEMMs of encoder layer l to its own lowest-priority side stream while layer l - 1 runs its BPTT, and the
    # block scheduler must prefer the main stream's kernels (torch's default stream has the LOWEST priority)
    stream = torch.cuda.Stream(priority=0 if os.environ.get("E2T_BENCH_PRIO0") else -1)
    stream.wait_stream(torch.cuda.current_stream())
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    init_engine(eng, seed=1)
    grads = flat_tensor(eng, _lib.GRAD_AND_COUNT)      # gradients + token count: ONE collective per step

    # ---- synthetic pool: 4 distinct batches per rank, 105 MB each -> 420 MB > 126 MB L2
    corpus = SyntheticCorpus(load_vocab(size=GEO["V"]), T=T_FRAMES, C=256, seed=0)
    NPOOL = 4
    host, dev = [], []
    for i in range(NPOOL):
        b = corpus.batch(B, seed=1000 * rank + i, L=L_TGT)
        hx = torch.from_numpy(b["encoder_inputs"]).pin_memory()
        hy = torch.from_numpy(b["decoder_targets"]).pin_memory()
        host.append((hx, hy))
        dev.append((hx.cuda(), hy.cuda()))
    ntok_t = torch.zeros(1, device="cuda")
    # the token count of a step sits behind the gradients in the same buffer (written by the backward pass)

    # N > 1: ONE flat NCCL all-reduce of the 57 MB gradient buffer per step; --overlap-allreduce runs it bucket by
    # bucket on a side stream while the backward pass is still going (ecog2txt_b200/dist.py: BucketedAllReduce).  Either way
    # the global token count stays on the device (e2t_adam_ema_step_dev): the timed loop has no host synchronisation
    from ecog2txt_b200.dist import BucketedAllReduce
    overlap = args.overlap_allreduce == "on"
    ar = BucketedAllReduce(eng) if world > 1 and overlap else None
    ntok_dev = [(y != 0).sum().float().reshape(1) for _, y in dev]
    ntok_cache = [float((hy != 0).sum()) for _, hy in host]

    def reduce_and_step(ntok_local_t):
        if ar is not None:
            ntok_t.copy_(ntok_local_t)
            ar.reduce_async(ntok_t)
            eng.adam_ema_step_dev(ntok_t)
            return
        if world > 1:
            dist.all_reduce(grads)          # 57.4 MB of gradients + the token count in the tail: one NCCL all-reduce
        eng.adam_ema_step_dev(None)         # 1 / global token count read from the tail, on the device

    def step_device(i):
        x, y = dev[i % NPOOL]
        eng.train_step_grads(x, None, y, seed=i, want_loss=False)
        reduce_and_step(ntok_dev[i % NPOOL])

    staged = {"next": None, "posted": None, "loss": None}

    def step_host(i):
        # end to end through the public API with HOST buffers: the H2D copy of THIS step's batch was started (pinned
        # memory, library copy stream) while the previous step computed; every step copies its inputs and reads its loss
        if staged["next"] != i:
            hx, hy = host[i % NPOOL]
            eng.stage_inputs(i & 1, hx.numpy(), None, hy.numpy())
        hx, hy = host[(i + 1) % NPOOL]
        eng.stage_inputs((i + 1) & 1, hx.numpy(), None, hy.numpy())
        staged["next"] = i + 1
        eng.train_step_grads_staged(i & 1, seed=i, want_loss=False)
        reduce_and_step(ntok_dev[i % NPOOL])
        # D2H of this step's loss + token count: posted behind the step (page-locked ring), read by the host one step
        # later -- every step's result reaches the host inside the timed region, the device is never left idle for it
        eng.post_losses(i & 3)
        if staged["posted"] is not None:
            staged["loss"] = eng.fetch_losses(staged["posted"])[0]
        staged["posted"] = i & 3
        return staged["loss"]

    def drain_host():
        if staged["posted"] is not None:
            staged["loss"] = eng.fetch_losses(staged["posted"])[0]
            staged["posted"] = None

    def timed(fn, steps, warmup, drain=None):
        for i in range(warmup):
            fn(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(warmup + i)
        if drain is not None:
            drain()                          # the last step's loss is on the host before the clock stops
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    l0, _ = eng.launch_counts()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_device, args.steps, max(args.warmup, 3))
    l1, tc1 = eng.launch_counts()
    launches = (l1 - l0) * args.steps // (args.steps + max(args.warmup, 3))
    ms_e2e = timed(step_host, args.steps, 3, drain_host)
    clocks = sampler.stop() if rank == 0 else None     # sampled over both timed regions
    value = B * world * args.steps / (ms * 1e-3)
    e2e = B * world * args.steps / (ms_e2e * 1e-3)

    # ---- roofline leg: recurrent-step kernels, CUDA events around every launch of the category
    fl = flops_per_utt(L_TGT)
    roof = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1590.0 * 0.88)
        psrc = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback"
        eng.profile_enable(True)
        # one untimed profiled step: profiling mode keeps everything on one stream, so GEMMs that normally run short-lived
        # on the side stream launch their persistent variant here for the first time (module load: seen as one 570 ms launch)
        eng.train_step_grads(dev[0][0], None, dev[0][1], seed=0, want_loss=False)
        eng.adam_ema_step(1.0 / ntok_cache[0])
        eng.profile_enable(True)   # drops the records of that step
        nprof = 3
        for i in range(nprof):     # rank-local steps: no collective may be issued by rank 0 alone
            x, y = dev[i % NPOOL]
            eng.train_step_grads(x, None, y, seed=i, want_loss=False)
            eng.adam_ema_step(1.0 / ntok_cache[i % NPOOL])
        cat_ms = {c: eng.profile_read(c) for c in range(6)}
        if args.breakdown:
            rep = eng.profile_report()
            with open(args.breakdown, "w") as f:
                f.write("# per-kernel CUDA-event time over %d training steps (ms total, launches, us/launch)\n" % nprof)
                for k, (n, t) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
                    f.write(f"{t / nprof:9.4f} ms/step  n/step={n // nprof:4d}  {1e3 * t / n:9.2f} us  {k}\n")
        eng.profile_enable(False)
        # (after the profiled steps: twenty 8192^3 GEMMs heat the part and would slow the kernels measured right behind them)
        # measured tf32 peak of THIS box: cuBLAS tf32 at 8192^3 with random operands (the analogue of the driver's cuBLAS
        # bf16 figure); the larger of it and bf16_sustained / 2 is the denominator, so no kernel of ours can beat its "peak"
        try:
            tf32_meas = measure_cublas_tflops(torch, torch.float32, True)
        except Exception as e:      # noqa: BLE001 -- a box without a usable cuBLAS still gets a bench line
            print(f"[bench] cuBLAS tf32 measurement failed: {e}", file=sys.stderr)
            tf32_meas = 0.0
        # ... and our own tcgen05 GEMM at the same size (random operands): it outruns cuBLAS tf32 on this part
        # (profiles/r2_gemm_peaks.json: 1037 vs 745 TFLOP/s), and a "peak" must not be beaten by the kernels measured against it
        try:
            own_ms = eng.bench_gemm(8192, 8192, 8192, False, 0.0, 10)
            tf32_own = 2.0 * 8192 ** 3 / own_ms / 1e9
        except Exception as e:      # noqa: BLE001
            print(f"[bench] own tf32 GEMM measurement failed: {e}", file=sys.stderr)
            tf32_own = 0.0
        # dominant kernels: the whole-layer persistent recurrent kernels (forward + BPTT) -- a latency chain of T' = 34
        # dependent steps per launch, so the tensor roofline is an upper bound they cannot approach
        T2, H = 34, 400
        flops_launch = 2 * 2 * T2 * B * H * 4 * H            # both directions, h Wh (or dz Wh^T) products of one layer
        n_f, n_b = cat_ms[4][1], cat_ms[5][1]
        rec_ms = cat_ms[4][0] + cat_ms[5][0]
        tot = sum(v[0] for v in cat_ms.values())
        peak_tf32 = max(peak_tf / 2.0, tf32_meas, tf32_own)
        achieved = flops_launch * (n_f + n_b) / (rec_ms * 1e-3) / 1e12 if rec_ms > 0 else 0.0
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            # per launch like `achieved`: mean of the forward and the BPTT kernel of one layer (ncu --set full, one launch each)
            traffic = (tj["k_lstm_fwd16"]["dram_bytes_per_launch"] + tj["k_lstm_bptt3"]["dram_bytes_per_launch"]) // 2
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        conv_bytes = B * T_FRAMES * 256 * 4
        conv_gbs = 2 * conv_bytes * nprof / (cat_ms[2][0] * 1e-3) / 1e9 if cat_ms[2][0] > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "k_lstm_fwd16 + k_lstm_bptt3 (persistent whole-layer recurrent kernels of the encoder, tcgen05 "
                                            "kind::f16 on fp16 operands with the 11-bit significand of tf32, fp32 accumulate)",
                "achieved": achieved, "peak": peak_tf32, "unit": "TFLOP/s", "frac": achieved / peak_tf32, "traffic": traffic,
                "peak_source": f"max(cuBLAS tf32 8192^3 measured live with random operands = {tf32_meas:.1f}, this repo's tcgen05 tf32 GEMM at "
                               f"8192^3 measured live = {tf32_own:.1f}, {psrc} / 2 = {peak_tf / 2:.1f})",
                "algorithmic_flops_per_launch": flops_launch, "launches_per_step": (n_f + n_b) // nprof,
                "avg_launch_us": 1e3 * rec_ms / max(n_f + n_b, 1), "share_of_step": rec_ms / tot if tot > 0 else None,
                "us_per_launch": {"fwd": 1e3 * cat_ms[4][0] / max(n_f, 1), "bptt": 1e3 * cat_ms[5][0] / max(n_b, 1)},
                "us_per_recurrent_step": {"fwd": 1e3 * cat_ms[4][0] / max(n_f, 1) / T2, "bptt": 1e3 * cat_ms[5][0] / max(n_b, 1) / T2},
                "category_ms_per_step": {k: cat_ms[i][0] / nprof for i, k in enumerate(
                    ("decoder_per_step_kernels", "bulk_gemm", "conv", "other", "persistent_rnn_fwd", "persistent_rnn_bwd"))},
                "hbm_kernel": {"kernel": "k_conv_tc fwd + bwd (gather-GEMM over the ECoG tensor)", "bound": "hbm",
                               "achieved": conv_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": conv_gbs / hbm_peak,
                               "algorithmic_bytes_per_launch": conv_bytes},
                "whole_step_frac_of_tf32_peak": (fl["train"] * value / world / 1e12) / peak_tf32}

    # ---- decode legs (BASELINE.json: greedy-decode ms/utterance; config 5: beam-8 throughput over 10 000 utterances sharded
    # over the ranks).  Inference is embarrassingly parallel: rank r decodes its contiguous share, no collective on the data
    # path; device-resident inputs from the same > L2 pool, CUDA-event time on the launching stream, max over ranks.
    decode = None
    if not args.no_decode:
        from ecog2txt_b200.dist import shard_range
        n_total = args.decode_utterances
        lo, hi = shard_range(n_total, rank, world)
        mine = hi - lo

        def run_decode(beam, bsz, n_utt):
            done, i = 0, 0
            while done < n_utt:
                n = min(bsz, n_utt - done)
                x = dev[i % NPOOL][0]
                xb = x if n == x.shape[0] else x[:n]
                if beam:
                    eng.beam_decode(xb, None, beam=beam, max_len=20)
                else:
                    eng.greedy_decode(xb, None, max_len=20, want_logp=False)
                done += n
                i += 1

        def timed_decode(beam, bsz, n_utt):
            run_decode(beam, bsz, min(2 * bsz, n_utt))          # warm-up
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run_decode(beam, bsz, n_utt)
            e1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()) * 1e-3

        t_greedy = timed_decode(0, B, mine)
        # beam-8: pick the minibatch size on a short probe (same choice on every rank: rank 0's)
        probe = {}
        for bsz in sorted({min(32, B), min(128, B), B}):
            try:
                probe[bsz] = timed_decode(8, bsz, 2 * bsz) / (2 * bsz)
            except Exception as e:   # a size the beam workspace cannot hold: skip it, keep the rest of the run
                print(f"[bench] beam-8 probe at batch {bsz} failed: {e}", file=sys.stderr)
        best = torch.tensor([min(probe, key=probe.get)], device="cuda")
        if world > 1:
            dist.broadcast(best, 0)
        beam_bsz = int(best.item())
        t_beam = timed_decode(8, beam_bsz, mine)
        decode = {"utterances": n_total, "sharding": f"{world} x {mine} (contiguous shares, no collective)",
                  "greedy_utt_per_s": n_total / t_greedy, "greedy_ms_per_utt": 1e3 * t_greedy / n_total, "greedy_batch": B,
                  "beam8_utt_per_s": n_total / t_beam, "beam8_ms_per_utt": 1e3 * t_beam / n_total, "beam8_batch": beam_bsz,
                  "beam8_probe_ms_per_utt": {str(k): 1e3 * v for k, v in probe.items()}, "max_len": 20}
        if rank == 0:
            # online-predictor path (trainers.py:925-949): ONE host utterance in, tokens out, wall clock per call
            x1 = np.ascontiguousarray(host[0][0][:1].numpy())
            for _ in range(3):
                eng.greedy_decode(x1, None, max_len=20, want_logp=False)
            t0 = time.perf_counter()
            for _ in range(20):
                eng.greedy_decode(x1, None, max_len=20, want_logp=False)
            decode["greedy_ms_per_utt_batch1_host"] = 1e3 * (time.perf_counter() - t0) / 20
            decode["batch1_graph_replays"] = eng.counter("decode_graph_replays")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, dt = cpu_port_utt_per_s(args.ref_batch, args.cpu_baseline_steps, warmup=2, mode="speed")
        vl, _, dtl = cpu_port_utt_per_s(args.ref_batch, 3, warmup=1, mode="loops")
        v1, _, dt1 = cpu_port_utt_per_s(64, 2, threads=1, warmup=1, mode="speed")      # also a 1-thread figure
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_baseline_steps} training steps of {args.ref_batch} utterances of the same workload "
                         f"({dt:.2f} s/step; oracle speed mode = torch CPU fp32 with torch.nn.LSTM / oneDNN; the reference's "
                         "TF1.15 path is not runnable here)",
               "value_explicit_loops": vl, "sample_explicit_loops": f"3 steps of {args.ref_batch} utterances, one LSTM step per "
                                                                    f"Python iteration + autograd ({dtl:.2f} s/step)",
               "value_1thread": v1, "sample_1thread": f"2 steps of 64 utterances on one thread, speed mode ({dt1:.2f} s/step)"}

    if rank == 0:
        hx, hy = host[0]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "tf32" if eng.launch_counts()[1] > 0 else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B,
                       "global_batch": B * world, "cache": "pool of 4 batches x 105 MB per rank (> 126 MB L2)",
                       "parallelism": f"dp{world}", "gemm_backend": args.backend,
                       "allreduce": None if world == 1 else ("ONE flat NCCL all-reduce per step (57.4 MB of gradients + the token count in its tail)" if ar is None else
                                                              f"{len(eng.grad_buckets())} buckets on a side stream, overlapped with the backward pass")},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(hx.numel() * 4 + hy.numel() * 4),
                    "d2h_bytes_per_step": 16, "ms_per_step": ms_e2e / args.steps,
                    "d2h": "loss + token count of every step, copied into page-locked memory behind the step (e2t_post_losses) and "
                           "read by the host one step later (e2t_fetch_losses); the last one before the clock stops",
                    "h2d_gbs_per_rank": (hx.numel() * 4 + hy.numel() * 4) / (ms_e2e / args.steps * 1e-3) / 1e9},
            "gpu_launches": int(launches), "tcgen05_launches_total": int(tc1),
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        if decode:
            line["decode"] = decode
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
