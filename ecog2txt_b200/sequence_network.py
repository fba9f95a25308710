"""``SequenceNetwork`` -- the drop-in class for the slot the reference fills with
``machine_learning.neural_networks.sequence_networks.SequenceNetwork`` (imported at
/root/reference/ecog2txt/trainers.py:33).  Same constructor keywords, attributes and methods as observed from
the reference's call sites:

* ctor                         trainers.py:126-135  (manifest-defaulted keys: yaml :3-5,11-12,25,29,62-75,88)
* ``fit``                      trainers.py:309-318,341-367 (train_vars_scope / reuse_vars_scope / _restore_epoch)
* ``restore_and_assess``       trainers.py:379-380 ; plotters.py:631-636
* ``get_weights_as_numpy_array`` trainers.py:699-700,750-751
* ``restore_and_get_saliencies`` trainers.py:722-725
* attributes                   checkpoint_path, N_epochs, layer_sizes, TEMPORALLY_CONVOLVE, EMA_decay, FF_dropout,
                               RNN_dropout, assessment_epoch_interval, beam_width, temperature,
                               inputs_to_occlude (plotters.py:607,630,639: channels silenced at assessment)

All arithmetic happens in libe2t.so (CUDA, sm_100a) behind the C-ABI; this file is the host-side epoch loop
(the reference's tfh.GraphBuilder train/assess cadence, trainers.py:852-859), the TFRecord batch assembly and
the host metrics.  Under torch.distributed every minibatch is sharded across ranks with one all-reduce of the
flat gradient buffer per step.
"""
from __future__ import annotations

import os
import re
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import EOS_token as _EOS, OOV_token as _OOV, pad_token as _PAD, _lib
from . import params as prm
from . import tfrecord
from .dist import allreduce_step, flat_tensor, shard_range
from .engine import Engine, EngineConfig
from .metrics import confusion_counts, target_inds_to_sequences, wer_vector

_MANIFEST_KEYS = dict(layer_sizes=None, FF_dropout=0.0, RNN_dropout=0.0, TEMPORALLY_CONVOLVE=True, EMA_decay=0.99,
                      N_epochs=800, beam_width=1, temperature=1.0, assessment_epoch_interval=10,
                      tf_summaries_dir=None)


class SequenceNetwork:
    def __init__(self, manifest, EOS_token=_EOS, pad_token=_PAD, OOV_token=_OOV, training_GPUs=(0,),
                 TARGETS_ARE_SEQUENCES=True, VERBOSE=True, N_cases=256, max_hyp_length=20, learning_rate=5e-4,
                 seed=1, gemm_backend="auto", attention="none", lib=None, loader_threads=None, device_cache_bytes=16 << 30,
                 **kwargs):
        # utils_jgm.auto_attribute(CHECK_MANIFEST=True): a keyword wins, else manifest[key] (README.md:42)
        for key, default in _MANIFEST_KEYS.items():
            if key in kwargs and kwargs[key] is not None:
                val = kwargs.pop(key)
            elif manifest is not None and key in manifest:
                val = manifest[key]
            else:
                val = default
            setattr(self, key, val)
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        if self.layer_sizes is None:
            raise ValueError("layer_sizes must come from the manifest or a keyword")
        # TEMPORALLY_CONVOLVE=False (README.md:80-83, trainers.py:386): the encoder embedding degenerates to a per-frame
        # dense layer = a temporal convolution of width 1 / stride 1, which the same kernels run (W = 1) [CHOICE: the
        # un-vendored graph builder is not visible; every shipped manifest sets it to true]
        self.EOS_token, self.pad_token, self.OOV_token = EOS_token, pad_token, OOV_token
        self.training_GPUs = list(training_GPUs)
        self.assessment_GPU = self.training_GPUs[0]
        self.TARGETS_ARE_SEQUENCES = TARGETS_ARE_SEQUENCES
        self.VERBOSE = VERBOSE
        self.N_cases, self.max_hyp_length, self.learning_rate, self.seed = N_cases, max_hyp_length, learning_rate, seed
        self.gemm_backend = gemm_backend
        # native threads assembling a minibatch (the tf.data num_parallel_calls of the reference, subjects.py:618)
        self.loader_threads = int(loader_threads) if loader_threads else max(1, min(8, (os.cpu_count() or 2) // 2))
        self.attention = attention       # "luong": optional A7 module (default "none" = the reference model)
        self.checkpoint_path: Optional[str] = None
        self.max_to_keep = 5             # checkpoints kept on disk (tf.train.Saver default)
        # fit(): keep the padded training set of every subject on the device (uploaded once per fit) when it fits this many
        # bytes; a step then moves only its example indices over PCIe instead of 105 MB of ECoG.  0 disables.
        self.device_cache_bytes = int(device_cache_bytes or 0)
        self.inputs_to_occlude = None
        self._lib = lib
        self._engine: Optional[Engine] = None
        self._engine_key = None

    # ------------------------------------------------------------------------------------------
    def vprint(self, *a, **k):
        if self.VERBOSE:
            print(*a, **k)

    def _geometry(self, subnets_params):
        ls = self.layer_sizes
        first = subnets_params[0].data_manifests
        V = int(first['decoder_targets'].num_features)
        flist = first['decoder_targets'].get_feature_list()
        geo = dict(
            subnet_ids=tuple(int(s.subnet_id) for s in subnets_params),
            subnet_C=tuple(int(s.data_manifests['encoder_inputs'].num_features) for s in subnets_params),
            subnet_W=tuple(int(s.decimation_factor) if self.TEMPORALLY_CONVOLVE else 1 for s in subnets_params),
            E=int(ls['encoder_embedding'][0]), H=tuple(int(h) for h in ls['encoder_rnn']),
            D=int(ls['decoder_embedding'][0]), Hd=int(ls['decoder_rnn'][0]), V=V,
            pad_id=flist.index(self.pad_token) if self.pad_token in flist else 0,
            eos_id=flist.index(self.EOS_token), start_id=flist.index(self.EOS_token),
        )
        # hidden layers of the vocabulary projection (mochastar_word_sequence.yaml:65; empty in every shipped manifest)
        proj_hidden = list(ls.get('decoder_projection') or [])
        if len(proj_hidden) > 1:
            raise NotImplementedError("at most one hidden layer in decoder_projection")
        if proj_hidden:
            geo['proj_hidden'] = int(proj_hidden[0])
        # A6: an 'encoder_<n>_targets' stream puts an FF head on encoder layer n (trainers.py:791-799; yaml:54,68-69)
        self._aux_key = None
        for key, man in first.items():
            m = re.fullmatch(r'encoder_(\d+)_targets', key)
            if m:
                if self._aux_key is not None:
                    raise NotImplementedError("one encoder-targets stream per model")
                n = int(m.group(1))
                hidden = list(ls.get(f'encoder_{n}_projection', []) or [])
                if len(hidden) > 1:
                    raise NotImplementedError("at most one hidden layer in encoder_<n>_projection")
                categorical = str(man.distribution).lower() == 'categorical'
                geo.update(aux_layer=n, aux_hidden=int(hidden[0]) if hidden else 0, aux_F=int(man.num_features),
                           aux_kind='categorical' if categorical else 'gaussian', aux_penalty=float(man.penalty_scale))
                self._aux_key = key
        geo['penalty_scale'] = float(first['decoder_targets'].penalty_scale)
        return geo, flist

    def _device(self) -> int:
        """CUDA ordinal of this process: one process per GPU under torchrun -> LOCAL_RANK (training_GPUs = [0] is what the
        reference hard-codes, trainers.py:131, and would put every rank on cuda:0); otherwise training_GPUs[0]."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and "LOCAL_RANK" in os.environ:
            return int(os.environ["LOCAL_RANK"])
        return int(self.training_GPUs[0])

    def _get_engine(self, subnets_params, max_T, max_L) -> Engine:
        geo, flist = self._geometry(subnets_params)
        key = (tuple(sorted(geo.items())), self.attention, self.FF_dropout, self.RNN_dropout, self.EMA_decay, int(self.beam_width), self.N_cases)
        e = self._engine
        if e is None or self._engine_key != key or e.cfg.max_T < max_T or e.cfg.max_L < max_L:
            if e is not None:
                e.close()
            dev = self._device()
            cfg = EngineConfig(**geo, max_B=self.N_cases, max_T=max_T, max_L=max(max_L, self.max_hyp_length),
                               max_beam=max(int(self.beam_width), 1), ff_dropout=float(self.FF_dropout),
                               rnn_dropout=float(self.RNN_dropout), lr=self.learning_rate,
                               ema_decay=float(self.EMA_decay), gemm_backend=self.gemm_backend, device=dev,
                               attention=self.attention)
            self._engine = Engine(cfg, lib=self._lib)
            prm.init_engine(self._engine, self.seed)
            self._engine_key = key
            self._bind_stream(self._engine)
        self._targets_list = flist
        return self._engine

    def _bind_stream(self, eng):
        """Data-parallel runs: the library must enqueue on the stream torch.distributed orders its NCCL kernels against
        (torch's current stream of this rank's device), or the all-reduce would race the backward pass."""
        import torch.distributed as dist
        if getattr(eng, "emulated", False) or not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        import torch
        torch.cuda.set_device(eng.cfg.device)
        eng.set_stream(torch.cuda.current_stream().cuda_stream)

    # -- data ----------------------------------------------------------------------------------
    @staticmethod
    def _load_partition(subject, partition):
        """[(x [T,C] fp32, y [L] int32, encoder targets or None)] for every trial of the partition's blocks
        (trainers.py:891-901)."""
        keys = ['encoder_inputs', 'decoder_targets'] + [k for k in subject.data_manifests if re.fullmatch(r'encoder_\d+_targets', k)]
        mans = {k: subject.data_manifests[k] for k in keys}
        paths = [subject.tf_record_partial_path.format(b) for b in sorted(subject.block_ids[partition])]
        aux = keys[2] if len(keys) > 2 else None
        return [(ex['encoder_inputs'], ex['decoder_targets'], ex[aux] if aux else None)
                for ex in tfrecord.read_examples(paths, mans)]

    @staticmethod
    def _batch(examples, idx, T_pad, L_pad, pad_id, threads=1):
        x = tfrecord.pad_batch_f32([examples[i][0] for i in idx], T_pad, threads=threads)
        y = np.full((len(idx), L_pad), pad_id, np.int32)
        for r, i in enumerate(idx):
            t = examples[i][1]
            y[r, :len(t)] = t
        return x, y

    def _ring_view(self, eng, k, B, T_pad, Cs):
        """[B, T_pad, Cs] view of slot k % 3 of the page-locked staging ring (allocated with the engine; three slots: the batch
        being consumed by the running step, the one being copied, the one being assembled).  Falls back to pageable memory
        if page-locked memory cannot be had."""
        if getattr(self, "_ring_owner", None) is not eng:
            self._ring_owner, self._ring = eng, None
            n = self.N_cases * eng.cfg.max_T * max(eng.cfg.subnet_C)
            try:
                self._ring = [eng.host_buffer((n,), np.float32) for _ in range(3)]
            except Exception as e:      # noqa: BLE001 -- any allocation failure: pageable staging still works, only slower
                self.vprint(f"page-locked staging ring unavailable ({e}); using pageable buffers")
        if self._ring is None:
            return None
        return self._ring[k % 3][:B * T_pad * Cs].reshape(B, T_pad, Cs)

    @staticmethod
    def _aux_batch(examples, idx, T_pad):
        """Encoder targets of a minibatch: fp32 [B,T,F] (zero padded) or int32 [B,T] class indices (pad index 0)."""
        a0 = examples[idx[0]][2]
        if a0 is None:
            return None
        if a0.dtype == np.float32 or a0.dtype == np.float64:
            return tfrecord.pad_batch_f32([np.asarray(examples[i][2], np.float32)[:T_pad] for i in idx], T_pad)
        out = np.zeros((len(idx), T_pad), np.int32)
        for r, i in enumerate(idx):
            t = np.asarray(examples[i][2]).reshape(-1)[:T_pad]
            out[r, :len(t)] = t
        return out

    # -- fit -----------------------------------------------------------------------------------
    def fit(self, subnets_params, train_vars_scope='seq2seq', reuse_vars_scope=None, _restore_epoch=None):
        """Train N_epochs on the subjects jointly (one subject per minibatch, App. D item 11); every
        assessment_epoch_interval epochs decode training / validation data with the EMA weights and checkpoint.
        Returns {'training'|'validation': struct(.decoder_accuracies, .decoder_word_error_rates, ...)}."""
        import torch.distributed as dist
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)
        data = {s.subnet_id: {p: self._load_partition(s, p) for p in ('training', 'validation')} for s in subnets_params}
        max_T = max(e[0].shape[0] for d in data.values() for part in d.values() for e in part)
        max_L = max(len(e[1]) for d in data.values() for part in d.values() for e in part)
        eng = self._get_engine(subnets_params, max_T, max_L)
        pad_id = eng.cfg.pad_id
        start_epoch, restored = 0, []
        if _restore_epoch is not None and _restore_epoch > 0:
            restored = prm.load_checkpoint(eng, self.checkpoint_path, _restore_epoch, reuse_vars_scope=reuse_vars_scope)
            start_epoch = _restore_epoch
        # everything NOT restored starts fresh (a cached engine of the same geometry may hold the previous fit's weights)
        prm.init_engine(eng, self.seed, keep=restored)
        if not restored:
            eng.step = 0
        pat = re.compile(train_vars_scope or 'seq2seq')
        for name in eng.tensors():
            eng.set_trainable(name, bool(pat.match(name)))
        # data-parallel: ONE collective per step -- the gradient buffer with the token count in its tail (dist.allreduce_step)
        grads = flat_tensor(eng, _lib.GRAD_AND_COUNT) if world > 1 else None
        rs = np.random.RandomState(self.seed + 17 + start_epoch)
        n_assess = self.N_epochs // self.assessment_epoch_interval
        assessments = {p: SimpleNamespace(decoder_accuracies=np.zeros(n_assess), decoder_word_error_rates=np.zeros(n_assess),
                                          decoder_confusions=None, epochs=np.zeros(n_assess, int), losses=np.zeros(n_assess))
                       for p in ('training', 'validation')}
        step = eng.step
        eng.read_loss_accumulators(reset=True)
        cache = self._device_cache(eng, subnets_params, data, max_T, max_L, pad_id)
        if cache is not None:
            import torch
        for epoch in range(start_epoch, start_epoch + self.N_epochs):
            # minibatches: (subject index, example indices); all ranks draw the same order, then shard each batch
            plan = []
            for si, s in enumerate(subnets_params):
                order = rs.permutation(len(data[s.subnet_id]['training']))
                plan += [(si, order[i:i + self.N_cases]) for i in range(0, len(order), self.N_cases)]
            order = rs.permutation(len(plan))
            # this rank's shard of every minibatch of the epoch.  Input pipeline (the tf.data prefetch of the reference,
            # trainers.py:891-901): minibatch k+1 is assembled on the host (native threads, into a page-locked ring buffer) and
            # its host->device copy is started (e2t_stage_inputs, library copy stream) WHILE minibatch k trains.  The step
            # itself is enqueue-only -- forward/backward, the all-reduce, Adam+EMA with the token count read on the device
            # (the loop bench.py times); the host runs at most ~3 minibatches ahead (wait_staged) and reads the summed
            # epoch loss once per epoch.
            shards = []
            for pi in order:
                si, idx = plan[pi]
                lo, hi = shard_range(len(idx), rank, world)
                shards.append((si, idx[lo:hi]))

            def host_batch(k):
                si, ids = shards[k]
                if not len(ids):
                    return None
                ex = data[subnets_params[si].subnet_id]['training']
                Cs = int(ex[ids[0]][0].shape[1])
                out = self._ring_view(eng, k, len(ids), max_T, Cs)
                x = tfrecord.pad_batch_f32([ex[i][0] for i in ids], max_T, out=out, threads=self.loader_threads)
                y = np.full((len(ids), max_L), pad_id, np.int32)
                for r, i in enumerate(ids):
                    y[r, :len(ex[i][1])] = ex[i][1]
                return x, y, self._aux_batch(ex, ids, max_T)

            if cache is not None:
                # device-resident path: gather the minibatch out of the cached training set on the device
                for k, (si, ids) in enumerate(shards):
                    if len(ids):
                        cx, cy = cache[subnets_params[si].subnet_id]
                        it = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64)).to(cx.device, non_blocking=True)
                        eng.train_step_grads(cx.index_select(0, it), None, cy.index_select(0, it), subnet=si, seed=step,
                                             want_loss=False)
                    else:
                        flat_tensor(eng, _lib.GRAD_AND_COUNT).zero_()
                    allreduce_step(grads)
                    eng.adam_ema_step_dev(None, subnet=si)
                    step += 1
                shards = []
            nxt = host_batch(0) if shards else None
            if nxt is not None:
                eng.stage_inputs(0, nxt[0], None, nxt[1], subnet=shards[0][0])
            for k, (si, ids) in enumerate(shards):
                cur = nxt
                if cur is not None:
                    if cur[2] is not None:   # A6: encoder targets of this step (copied on the compute stream)
                        eng.set_encoder_targets(cur[2])
                    eng.train_step_grads_staged(k & 1, seed=step, want_loss=False)      # enqueue only
                else:   # fewer utterances than ranks: contribute zero gradients and a zero count
                    flat_tensor(eng, _lib.GRAD_AND_COUNT).zero_()
                # assemble and stage the next minibatch while this one computes; ring slot (k+1) % 3 was last read by the
                # copy of minibatch k-2, which has finished once the copy of minibatch k-1 has (same copy stream)
                if k >= 1:
                    eng.wait_staged((k - 1) & 1)
                nxt = host_batch(k + 1) if k + 1 < len(shards) else None
                if nxt is not None:
                    eng.stage_inputs((k + 1) & 1, nxt[0], None, nxt[1], subnet=shards[k + 1][0])
                allreduce_step(grads)
                eng.adam_ema_step_dev(None, subnet=si)
                step += 1
            ep_loss, ep_tok, ep_aux, _ = eng.read_loss_accumulators(reset=True)      # the epoch's one host synchronisation
            ep_loss += ep_aux
            done = epoch + 1 - start_epoch
            if done % self.assessment_epoch_interval == 0:
                k = done // self.assessment_epoch_interval - 1
                for part in ('training', 'validation'):
                    res = self._assess(eng, subnets_params, data, part, max_T, max_L)
                    a = assessments[part]
                    a.decoder_accuracies[k], a.decoder_word_error_rates[k] = res.accuracy, res.word_error_rate
                    a.decoder_confusions = res.decoder_confusions      # of the latest assessment (trainers.py:604-611)
                    a.epochs[k], a.losses[k] = epoch + 1, ep_loss / max(ep_tok, 1)
                self.vprint(f"epoch {epoch + 1}: train loss/token {ep_loss / max(ep_tok, 1):.4f}  "
                            f"WER train {assessments['training'].decoder_word_error_rates[k]:.3f} "
                            f"valid {assessments['validation'].decoder_word_error_rates[k]:.3f}")
                if self.checkpoint_path and rank == 0:
                    prm.save_checkpoint(eng, self.checkpoint_path, epoch + 1, max_to_keep=self.max_to_keep)
        if self.checkpoint_path and rank == 0 and self.N_epochs % self.assessment_epoch_interval:
            prm.save_checkpoint(eng, self.checkpoint_path, start_epoch + self.N_epochs, max_to_keep=self.max_to_keep)
        if world > 1:
            dist.barrier()      # rank 0's checkpoint is complete before any rank goes on to restore it
        return assessments

    def _device_cache(self, eng, subnets_params, data, max_T, max_L, pad_id):
        """{subnet_id: (x [N, max_T, C] fp32, y [N, max_L] int32)} as CUDA tensors, or None (emulation build, encoder targets
        in play, or the padded training sets exceed device_cache_bytes).  The library is bound to torch's current stream so
        that the gathers and the library's kernels are ordered."""
        if getattr(eng, "emulated", False) or not self.device_cache_bytes:
            return None
        total = 0
        for s in subnets_params:
            ex = data[s.subnet_id]['training']
            if not len(ex) or any(e[2] is not None for e in ex[:1]):
                return None
            total += len(ex) * max_T * int(ex[0][0].shape[1]) * 4
        if total > self.device_cache_bytes:
            return None
        import torch
        dev = torch.device("cuda", eng.cfg.device)
        torch.cuda.set_device(dev)
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
        cache = {}
        for s in subnets_params:
            ex = data[s.subnet_id]['training']
            x = tfrecord.pad_batch_f32([e[0] for e in ex], max_T, threads=self.loader_threads)
            y = np.full((len(ex), max_L), pad_id, np.int32)
            for r, e in enumerate(ex):
                y[r, :len(e[1])] = e[1]
            cache[s.subnet_id] = (torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev))
        return cache

    # -- assessment ----------------------------------------------------------------------------
    def _assess(self, eng, subnets_params, data, partition, max_T, max_L):
        """Decode (greedy if beam_width == 1, else beam) with the EMA weights; WER over the last subject's trials
        like the reference's assessor (one subnet's validation data, trainers.py:838-849)."""
        import torch.distributed as dist
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)
        s = subnets_params[-1]
        si = len(subnets_params) - 1
        examples = data[s.subnet_id][partition]
        pad_id, eos_id = eng.cfg.pad_id, eng.cfg.eos_id
        refs, hyps, n_ok, n_tok = [], [], 0, 0
        V = int(eng.cfg.V)
        conf = np.zeros((V, V), np.int64)
        Lh = max(self.max_hyp_length, max_L)
        # inference is embarrassingly parallel over utterances: rank r decodes every world-th minibatch,
        # the decoded strings are gathered at the end (no collective on the data path)
        for bi, i in enumerate(range(0, len(examples), self.N_cases)):
            if bi % world != rank:
                continue
            idx = np.arange(i, min(i + self.N_cases, len(examples)))
            x, y = self._batch(examples, idx, max_T, max_L, pad_id, threads=self.loader_threads)
            if self.inputs_to_occlude is not None and len(self.inputs_to_occlude):
                # test-time occlusion (plotters.py:603-640): the listed input channels are silenced for this assessment
                x[:, :, np.asarray(self.inputs_to_occlude, int)] = 0.0
            if int(self.beam_width) > 1:
                toks, _ = eng.beam_decode(x, None, beam=int(self.beam_width), max_len=Lh, subnet=si, use_ema=True,
                                          temperature=float(self.temperature))
            else:
                t, _ = eng.greedy_decode(x, None, max_len=Lh, subnet=si, use_ema=True,
                                         temperature=float(self.temperature), want_logp=False)
                toks = t[:, None, :]
            for r in range(len(idx)):
                refs.append(target_inds_to_sequences(y[r][None, None, :], self._targets_list)[0])
                hyps.append(target_inds_to_sequences(toks[r:r + 1], self._targets_list)[0])
                m = y[r] != pad_id
                n_ok += int((toks[r, 0, :max_L][m] == y[r][m]).sum())
                n_tok += int(m.sum())
            conf += confusion_counts(y, toks[:, 0, :max_L], V, pad_id)
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, (refs, hyps, n_ok, n_tok, conf))
            conf = sum(p[4] for p in parts)
            # rank-major order -> original minibatch order (rank r holds minibatches r, r + world, ...)
            nb = -(-len(examples) // self.N_cases)
            refs, hyps, cursor = [], [], [0] * world
            for bi in range(nb):
                r = bi % world
                n = min(self.N_cases, len(examples) - bi * self.N_cases)
                refs += parts[r][0][cursor[r]:cursor[r] + n]
                hyps += parts[r][1][cursor[r]:cursor[r] + n]
                cursor[r] += n
            n_ok, n_tok = sum(p[2] for p in parts), sum(p[3] for p in parts)
        wers = wer_vector(refs, hyps) if refs else np.zeros(0)
        return SimpleNamespace(word_error_rate=float(wers.mean()) if len(wers) else float('nan'),
                               accuracy=n_ok / max(n_tok, 1), references=refs, hypotheses=hyps,
                               word_error_rates=wers, decoder_confusions=conf)

    def restore_and_assess(self, subnets_params, restore_epoch, WRITE=False, data_partitions=('training', 'validation')):
        data = {s.subnet_id: {p: self._load_partition(s, p) for p in data_partitions} for s in subnets_params}
        max_T = max(e[0].shape[0] for d in data.values() for part in d.values() for e in part)
        max_L = max(len(e[1]) for d in data.values() for part in d.values() for e in part)
        eng = self._get_engine(subnets_params, max_T, max_L)
        prm.load_checkpoint(eng, self.checkpoint_path, restore_epoch, reuse_vars_scope='seq2seq')
        return {p: self._assess(eng, subnets_params, data, p, max_T, max_L) for p in data_partitions}

    def restore_and_get_saliencies(self, subnets_params, restore_epoch, data_partition='validation', assessment_type='norms'):
        """d(loss)/d(encoder_inputs) of the last subject's `data_partition` trials with the EMA weights of checkpoint
        `restore_epoch` (trainers.py:722-725).  The penalties in force are the manifests' `penalty_scale`s: the caller
        (MultiSubjectTrainer.get_saliencies, trainers.py:703-732) zeroes all but the one under study.
        assessment_type 'norms': [C] = mean over trials of the per-electrode L2 norm over time of the input gradient;
        'sequences': list of [T_i, C] gradients, one per trial [CHOICE: the upstream definition is not in the tree]."""
        if assessment_type not in ('norms', 'sequences'):
            raise ValueError("assessment_type must be 'norms' or 'sequences'")
        s = subnets_params[-1]
        si = len(subnets_params) - 1
        examples = self._load_partition(s, data_partition)
        max_T = max(e[0].shape[0] for e in examples)
        max_L = max(len(e[1]) for e in examples)
        eng = self._get_engine(subnets_params, max_T, max_L)
        prm.load_checkpoint(eng, self.checkpoint_path, restore_epoch, reuse_vars_scope='seq2seq')
        mans = s.data_manifests
        pd = float(mans['decoder_targets'].penalty_scale)
        pa = float(mans[self._aux_key].penalty_scale) if self._aux_key else 0.0
        norms, seqs = [], []
        for i in range(0, len(examples), self.N_cases):
            idx = np.arange(i, min(i + self.N_cases, len(examples)))
            x, y = self._batch(examples, idx, max_T, max_L, eng.cfg.pad_id, threads=self.loader_threads)
            aux = self._aux_batch(examples, idx, max_T)
            if aux is not None:
                eng.set_encoder_targets(aux)
            dx, sq = eng.input_saliency(x, None, y, subnet=si, use_ema=float(self.EMA_decay) > 0, decoder_penalty=pd,
                                        aux_penalty=pa, want_dx=assessment_type == 'sequences')
            norms.append(np.sqrt(sq))
            if dx is not None:
                seqs += [dx[r, :examples[j][0].shape[0]].copy() for r, j in enumerate(idx)]
        return np.concatenate(norms).mean(0) if assessment_type == 'norms' else seqs

    def restore_and_get_activations(self, subnets_params, restore_epoch, data_partition='validation'):
        """Forward pass (no dropout, EMA weights of checkpoint `restore_epoch`) over the last subject's `data_partition` trials;
        returns what MultiSubjectTrainer.get_internal_activations assembles in the reference (trainers.py:757-859):
        convolved_inputs [N,T',E], reversed_inputs [N,T,C], decimated_reversed_targets [N,T',F] (None without an encoder-targets
        stream), final_RNN_state [2 (c, h), 1 (last encoder layer), N, decoder units] (axes of plotters.py:1388), lengths [N]."""
        s = subnets_params[-1]
        si = len(subnets_params) - 1
        W = int(s.decimation_factor)
        examples = self._load_partition(s, data_partition)
        max_T = max(e[0].shape[0] for e in examples)
        max_L = max(len(e[1]) for e in examples)
        eng = self._get_engine(subnets_params, max_T, max_L)
        prm.load_checkpoint(eng, self.checkpoint_path, restore_epoch, reuse_vars_scope='seq2seq')
        T2 = -(-max_T // W)
        conv, rev, tgt, fin, lens_all = [], [], [], [], []
        for i in range(0, len(examples), self.N_cases):
            idx = np.arange(i, min(i + self.N_cases, len(examples)))
            B = len(idx)
            x, y = self._batch(examples, idx, max_T, max_L, eng.cfg.pad_id, threads=self.loader_threads)
            eng.eval_loss(x, None, y, subnet=si, use_ema=float(self.EMA_decay) > 0)
            lens = eng.activation("lens", (B,), np.int32)
            conv.append(eng.activation("conv_out", (T2, B, eng.cfg.E)).transpose(1, 0, 2))
            fin.append(np.stack([eng.activation("final_c", (B, eng.cfg.Hd)), eng.activation("final_h", (B, eng.cfg.Hd))])[:, None])
            r = x.copy()
            for b in range(B):
                r[b, :lens[b]] = x[b, :lens[b]][::-1]          # tf.reverse_sequence, trainers.py:808-810
            rev.append(r)
            aux = self._aux_batch(examples, idx, max_T)
            if aux is not None:
                ra = aux.copy()
                for b in range(B):
                    ra[b, :lens[b]] = aux[b, :lens[b]][::-1]
                tgt.append(ra[:, 0::W])                        # trainers.py:791-795
            lens_all.append(lens)
        return SimpleNamespace(convolved_inputs=np.concatenate(conv), reversed_inputs=np.concatenate(rev),
                               decimated_reversed_targets=np.concatenate(tgt) if tgt else None,
                               final_RNN_state=np.concatenate(fin, axis=2), lengths=np.concatenate(lens_all))

    def get_weights_as_numpy_array(self, full_var_name, restore_epoch):
        """The stored variable (or its `/ExponentialMovingAverage` shadow) from checkpoint `restore_epoch`."""
        with np.load(f"{self.checkpoint_path}-{restore_epoch}.npz") as z:
            return np.array(z[full_var_name])

    # -- online predictor (construct_online_predictor, trainers.py:925-949) ---------------------------
    def prepare_for_prediction(self, subnets_params, restore_epoch, max_T: int = 1250):
        """Build the engine for these subjects and restore checkpoint `restore_epoch` (EMA weights are used to decode).
        max_T: longest utterance the predictor accepts (data_generators.py:38,157 clips trials at 1250 samples)."""
        eng = self._get_engine(subnets_params, max_T, self.max_hyp_length)
        prm.load_checkpoint(eng, self.checkpoint_path, restore_epoch, reuse_vars_scope='seq2seq')
        return eng

    def predict_tokens(self, inputs: np.ndarray, subnet: int = 0) -> np.ndarray:
        """One utterance [T, C] -> token indices [max_hyp_length] (greedy, EMA weights).  Calls of one shape are replayed
        as a single CUDA graph by the library from the second call on (e2t.h: e2t_greedy_decode)."""
        eng = self._engine
        if eng is None:
            raise RuntimeError("fit, restore_and_assess or prepare_for_prediction first")
        x = np.ascontiguousarray(inputs[None], np.float32)
        toks, _ = eng.greedy_decode(x, None, max_len=self.max_hyp_length, subnet=subnet, use_ema=True,
                                    temperature=float(self.temperature), want_logp=False)
        return toks[0]

    def predict(self, inputs: np.ndarray, subnet: int = 0) -> str:
        """One utterance [T, C] -> decoded sentence with the EMA weights (greedy)."""
        return target_inds_to_sequences(self.predict_tokens(inputs, subnet)[None, None, :], self._targets_list)[0]
