"""``MultiSubjectTrainer`` -- the direct caller of the hot path, re-stated thinly.

Mirrors /root/reference/ecog2txt/trainers.py:41-408,444-554,925-963 for everything that drives ``SequenceNetwork``:

* ctor                               trainers.py:42-141   (subjects, last subject's manifest -> SequenceNetwork, EOS on targets)
* ``checkpoint_dir`` / ``restore_epoch``   trainers.py:213-252   (``model.ckpt-<epoch>.index`` discovery)
* ``parallel_transfer_learn``        trainers.py:303-327  (all subjects in one fit; RESUME)
* ``sequential_transfer_learn``      trainers.py:329-374  (per subject: pre-train the private subnet with the shared scope
                                                           frozen and restored, then train everything)
* ``assess_saved_model`` / ``update_net_from_saved_model`` / ``recover_model_sizes``   trainers.py:376-554
* ``construct_online_predictor`` / ``target_inds_to_sequences``   trainers.py:925-963
* ``_retrieve_layer_weights`` / ``get_encoder_embedding`` / ``get_saliencies`` / ``get_internal_activations`` /
  ``tf_record_to_numpy_data``        trainers.py:680-922

Out of scope here exactly as in SURVEY.md section 2: result files for the plotters, tf summaries, saliency plots,
``ECoGDataGenerator`` subclasses (lab-private data): subjects are passed in ready-made (``subjects.ECoGSubject``;
``subjects.make_synthetic_subject`` builds synthetic ones), or looked up through ``subject_factory``.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import EOS_token, OOV_token, pad_token, TOKEN_TYPES
from . import params as prm
from .metrics import target_inds_to_sequences
from .sequence_network import SequenceNetwork


# variable-scope regexes that split the model into its two halves (the reference's transfer-learning contract,
# trainers.py:337-338): tensors under seq2seq/subnet_<id>/ belong to one subject, everything else is shared
SUBJECT_SCOPE = 'seq2seq/subnet'
SHARED_SCOPE = 'seq2seq/(?!subnet)'
WHOLE_MODEL_SCOPE = 'seq2seq'


@dataclass
class FitPhase:
    """One `SequenceNetwork.fit` call of a transfer-learning schedule."""
    subject_index: int
    n_epochs: int
    train_scope: str                 # what Adam updates
    restore_scope: Optional[str]     # what is read back from the checkpoint of `restore_from` (None: nothing, fresh start)
    restore_from: int                # epoch of the checkpoint to start from (0: none)
    report: bool                     # whether the phase's assessments are kept as a result


def sequential_schedule(n_subjects: int, pretraining_epochs: int, training_epochs: int, posttraining_epochs: int) -> List[FitPhase]:
    """The reference's sequential transfer learning (trainers.py:329-374) as data.  Subject 0 trains the whole model from
    scratch; every later subject first fits ONLY its private input layer for `pretraining_epochs` on top of the restored
    (frozen) shared layers, then the whole model for `training_epochs`; the last subject trains `posttraining_epochs` longer.
    Epoch numbers accumulate across phases: each phase starts from the checkpoint the previous one ended with."""
    phases, epoch = [], 0
    for i in range(n_subjects):
        if i > 0:
            phases.append(FitPhase(i, pretraining_epochs, SUBJECT_SCOPE, SHARED_SCOPE, epoch, report=False))
            epoch += pretraining_epochs
        n = training_epochs + (posttraining_epochs if i == n_subjects - 1 else 0)
        phases.append(FitPhase(i, n, WHOLE_MODEL_SCOPE, WHOLE_MODEL_SCOPE if i > 0 else None, epoch, report=True))
        epoch += n
    return phases


class MultiSubjectTrainer:
    def __init__(self, experiment_manifest, subject_ids: Sequence[int], checkpoint_dir: str = '.', restore_epoch=None,
                 SN_kwargs=(), ES_kwargs=(), VERBOSE=True, subjects: Optional[Sequence] = None,
                 subject_factory: Optional[Callable] = None, **kwargs):
        """experiment_manifest: {subject_id: manifest dict} (the parsed YAML, trainers.py:60-61) or a path to it."""
        SN_kwargs = dict(SN_kwargs)
        if isinstance(experiment_manifest, str):
            import yaml
            with open(experiment_manifest) as f:
                experiment_manifest = yaml.safe_load(f)
        self.experiment_manifest = experiment_manifest
        last = experiment_manifest[subject_ids[-1]]
        token_type = last.get('token_type', 'word_sequence')
        if token_type not in TOKEN_TYPES:                                            # trainers.py:64-65
            raise ValueError(f"token_type {token_type!r} is not one of {sorted(TOKEN_TYPES)}")
        self._token_type = token_type
        # subjects: every subject but the last pre-trains on all of its blocks (trainers.py:72-82; subjects.py:123-126)
        if subjects is None:
            assert subject_factory is not None, "pass ready-made `subjects` or a `subject_factory(manifest, id, **kw)`"
            subjects = [subject_factory(experiment_manifest[sid], sid, pretrain_all_blocks=(sid != subject_ids[-1]),
                                        **dict(ES_kwargs)) for sid in subject_ids]
        else:
            subjects = list(subjects)
            assert [s.subnet_id for s in subjects] == list(subject_ids)
            for s in subjects:
                s.pretrain_all_blocks = s.subnet_id != subject_ids[-1]
        self.ecog_subjects: List = subjects
        self.VERBOSE = VERBOSE
        self._restore_epoch = restore_epoch
        # data manifests: EOS on sequence targets, per-stream penalty scales from the manifest (trainers.py:93-103)
        for subject in self.ecog_subjects:
            for data_key, man in subject.data_manifests.items():
                if data_key == 'decoder_targets' and 'sequence' in token_type:
                    man.APPEND_EOS = True
                try:
                    man.penalty_scale = experiment_manifest[subject.subnet_id][data_key + '_penalty_scale']
                except KeyError:
                    pass
        # the reference pins training_GPUs=[0] (trainers.py:131); here SN_kwargs may override it, and under torchrun the
        # network binds to this rank's LOCAL_RANK device (SequenceNetwork._device)
        SN_kwargs.setdefault('training_GPUs', [0])
        self.net = SequenceNetwork(last, EOS_token=EOS_token, pad_token=pad_token, OOV_token=OOV_token,
                                   TARGETS_ARE_SEQUENCES='sequence' in token_type, VERBOSE=VERBOSE, **SN_kwargs)
        self.checkpoint_dir = checkpoint_dir
        self.results: List[dict] = []

    # ---- checkpoint location and discovery (trainers.py:213-252) -----------------------------------
    @property
    def checkpoint_dir(self) -> str:
        return self._checkpoint_dir

    @checkpoint_dir.setter
    def checkpoint_dir(self, path: str):
        """The network always writes <checkpoint_dir>/model.ckpt-<epoch>.*; moving the directory re-points it."""
        self._checkpoint_dir = path
        self.net.checkpoint_path = os.path.join(path, 'model.ckpt')

    @property
    def restore_epoch(self) -> Optional[int]:
        """The epoch set explicitly, else the newest checkpoint found in checkpoint_dir, else None."""
        if self._restore_epoch is not None:
            return self._restore_epoch
        found = prm.checkpoint_epochs(self.net.checkpoint_path)
        return found[-1] if found else None

    @restore_epoch.setter
    def restore_epoch(self, epoch: Optional[int]):
        self._restore_epoch = epoch

    def vprint(self, *a, **k):
        if self.VERBOSE:
            print(*a, **k)

    def _save_results(self, assessments):
        self.results.append(assessments)

    # ---- trainers.py:303-327 ---------------------------------------------------------------------
    def parallel_transfer_learn(self, RESUME=False, fit_kwargs=()):
        """All subjects jointly in one fit (one subject per minibatch, shared layers see everyone's data)."""
        fit_kwargs = dict(fit_kwargs)
        if RESUME:
            # continue from the newest checkpoint with the LAST subject only, whole model restored and trainable
            fit_kwargs.setdefault('_restore_epoch', self.restore_epoch)
            fit_kwargs.update(train_vars_scope=WHOLE_MODEL_SCOPE, reuse_vars_scope=WHOLE_MODEL_SCOPE)
            self.ecog_subjects = self.ecog_subjects[-1:]
        assessments = self.net.fit(self.ecog_subjects, **fit_kwargs)
        self._save_results(assessments)
        if self._restore_epoch is not None:
            self.restore_epoch = self.restore_epoch + self.net.N_epochs if RESUME else self.net.N_epochs
        return assessments

    # ---- sequential transfer learning (trainers.py:329-374) ------------------------------------------
    def sequential_transfer_learn(self, pretraining_epochs=60, training_epochs=200, posttraining_epochs=340):
        """Subjects one after the other; see `sequential_schedule` for the phases.  Returns the last phase's assessments."""
        assessments = None
        phases = sequential_schedule(len(self.ecog_subjects), pretraining_epochs, training_epochs, posttraining_epochs)
        for ph in phases:
            self.net.N_epochs = ph.n_epochs
            out = self.net.fit([self.ecog_subjects[ph.subject_index]], train_vars_scope=ph.train_scope,
                               reuse_vars_scope=ph.restore_scope, _restore_epoch=ph.restore_from or None)
            if ph.report:
                assessments = out
                self._save_results(out)
        self.restore_epoch = phases[-1].restore_from + phases[-1].n_epochs
        return assessments

    # ---- trainers.py:376-408 ---------------------------------------------------------------------
    def assess_saved_model(self):
        self.update_net_from_saved_model()
        return self.net.restore_and_assess(self.ecog_subjects, self.restore_epoch)

    def update_net_from_saved_model(self):
        """Make the network and the subjects' manifests agree with what the checkpoint holds (sizes are read back from
        the stored variable names / shapes, so a model can be assessed without its original manifest)."""
        layer_sizes, data_sizes, strides, has_ema = self.recover_model_sizes()
        net = self.net
        net.layer_sizes = layer_sizes
        net.TEMPORALLY_CONVOLVE = len(strides)          # truthy iff conv kernels were found
        net.EMA_decay = 0.99 * has_ema
        shared_sizes = data_sizes.get(None, {})
        for subject in self.ecog_subjects:
            own_sizes = data_sizes.get(subject.subnet_id, {})
            for key, man in subject.data_manifests.items():
                if key in own_sizes or key in shared_sizes:
                    man.num_features = shared_sizes.get(key, own_sizes.get(key))
            subject_strides = strides.get(subject.subnet_id)
            if subject_strides:
                subject.decimation_factor = int(np.prod(subject_strides))

    # ---- trainers.py:444-554 ---------------------------------------------------------------------
    def recover_model_sizes(self):
        """layer_sizes, data_sizes, strides, EMA flag from the variable names / shapes of the latest checkpoint, by the
        same rules as the reference: `subnet_<id>` scoping, `<subsubnet>_<Nin>_<Nout>_<layer>/weights`, 4-D conv kernels
        (stride = shape[1], inputs = shape[-2]), 4-gate LSTM kernels (size = shape[-1] // 4), transposed final projection."""
        var_shapes = prm.variable_to_shape_map(self.net.checkpoint_path, self.restore_epoch)
        EMA = int(any(name.endswith('/ExponentialMovingAverage') for name in var_shapes))
        layer_sizes: Dict[str, list] = {}
        data_sizes: Dict = {}
        strides: Dict = {}
        rnn = {}
        proj: Dict = {}
        for name, shape in var_shapes.items():
            parts = name.split('/')
            if parts[0] != 'seq2seq' or parts[-1] not in ('weights', 'kernel') or 'Adam' in name:
                continue
            subnet_id = None
            scope = parts[1:]
            if scope[0].startswith('subnet_'):
                subnet_id = int(scope[0].split('_')[1])
                scope = scope[1:]
            if any(re.fullmatch(r'cell_\d+', p) for p in scope):                     # trainers.py:480-485,527-529
                m = re.fullmatch(r'(\w+?)(?:_(\d+))?', scope[0])
                key = scope[0] if not re.search(r'_\d+$', scope[0]) else scope[0].rsplit('_', 1)[0]
                idx = int(scope[0].rsplit('_', 1)[1]) if re.search(r'_\d+$', scope[0]) else 0
                rnn.setdefault(key, {})[idx] = shape[-1] // 4
                continue
            m = re.fullmatch(r'(\w+)_(\d+)_(\d+)_(\d+)', scope[0])                   # "three numbers appended", :488-491
            if not m:
                continue
            subsub, layer = m.group(1), int(m.group(4))
            if len(shape) == 4:                                                      # conv: trainers.py:534-541
                strides.setdefault(subnet_id, []).append(shape[1])
                data_sizes.setdefault(subnet_id, {})['encoder_inputs'] = shape[-2]
                layer_sizes.setdefault(subsub, []).append(shape[-1])
            elif subsub.endswith('_projection'):                                     # resolved below (needs every layer)
                proj.setdefault((subnet_id, subsub), {})[layer] = shape
            else:
                if subsub == 'decoder_embedding':
                    data_sizes.setdefault(subnet_id, {})['decoder_targets'] = shape[0]
                layer_sizes.setdefault(subsub, []).append(shape[-1])
        for (subnet_id, subsub), layers in proj.items():
            # the LAST layer of a *_projection is stored transposed and only tells the output size (trainers.py:513-520);
            # earlier layers are ordinary hidden layers
            layer_sizes[subsub] = [layers[k][-1] for k in sorted(layers)[:-1]]
            data_sizes.setdefault(subnet_id, {})[subsub.replace('_projection', '_targets')] = layers[max(layers)][0]
        for key, d in rnn.items():                                                   # encoder_rnn_<n> -> one list, :543-552
            layer_sizes[key] = [d[i] for i in sorted(d)]
        return layer_sizes, data_sizes, strides, EMA

    # ---- trainers.py:676-701,734-751 -------------------------------------------------------------
    def _retrieve_layer_weights(self, weights_name):
        """The EMA copy of the first layer's weights of the sub-network called `weights_name` (e.g. 'decoder_embedding',
        'encoder_embedding') from the restore_epoch checkpoint, found by name like the reference does."""
        stored = prm.variable_to_shape_map(self.net.checkpoint_path, self.restore_epoch)
        wanted = re.compile(rf'.*{weights_name}.*0/weights{re.escape(prm.EMA_SUFFIX)}')
        matches = [name for name in sorted(stored) if wanted.match(name)]
        if not matches:
            raise KeyError(f"checkpoint {self.restore_epoch} holds no first-layer EMA weights for {weights_name!r}")
        return self.net.get_weights_as_numpy_array(matches[-1], self.restore_epoch)

    def get_encoder_embedding(self):
        """The last subject's temporal-conv kernel [1, W, C, E] (EMA), named from the recovered model sizes."""
        layer_sizes, data_sizes, _, _ = self.recover_model_sizes()
        subj_id = self.ecog_subjects[-1].subnet_id
        name = 'seq2seq/subnet_{0}/encoder_embedding_{1}_{2}_0/weights/ExponentialMovingAverage'.format(
            subj_id, data_sizes[subj_id]['encoder_inputs'], layer_sizes['encoder_embedding'][0])
        return self.net.get_weights_as_numpy_array(name, self.restore_epoch)

    # ---- trainers.py:757-859 ---------------------------------------------------------------------
    def get_internal_activations(self):
        """convolved_inputs, reversed_inputs, decimated_reversed_targets, final_RNN_state of the last subject's validation data
        with the restored (EMA) weights."""
        return self.net.restore_and_get_activations(self.ecog_subjects, self.restore_epoch, data_partition='validation')

    # ---- trainers.py:861-922 ---------------------------------------------------------------------
    def tf_record_to_numpy_data(self, subj_id, block_id):
        """Yields one dict per trial of the block's TFRecord: float streams reshaped to [T, num_features_raw], string streams
        as raw byte strings [T, 1] (no index substitution) -- for inspecting the content of the records."""
        from . import tfrecord
        from .subjects import SequenceDataManifest
        for subject in self.ecog_subjects:
            if subject.subj_id == subj_id:
                break
        else:
            raise ValueError('Requested subject not in this trainer')
        raw = {}
        for key, man in subject.data_manifests.items():
            is_float = man.get_feature_list is None
            raw[key] = SequenceDataManifest(man.sequence_type,
                                            num_features_raw=man.num_features_raw if is_float else 1)
        yield from tfrecord.read_examples([subject.tf_record_partial_path.format(block_id)], raw)

    # ---- trainers.py:703-732 ---------------------------------------------------------------------
    def get_saliencies(self, contrib_method, assessment_type='norms'):
        """Average "saliency" of the input electrodes: error gradients of ONE output back-propagated into the inputs.
        contrib_method = '<data_key minus _targets>_saliency_map', e.g. 'decoder_saliency_map' / 'encoder_1_saliency_map':
        every *_targets penalty is set to 0 except that one (set to 1), as in the reference."""
        subject = self.ecog_subjects[-1]
        target_streams = {k: m for k, m in subject.data_manifests.items() if '_targets' in k}
        studied = contrib_method.replace('saliency_map', 'targets')
        if studied not in target_streams:
            raise KeyError(f"{contrib_method!r} names no target stream of subject {subject.subnet_id}")
        saved = {k: m.penalty_scale for k, m in target_streams.items()}
        for k, m in target_streams.items():
            m.penalty_scale = 1.0 if k == studied else 0.0
        try:
            return self.net.restore_and_get_saliencies(self.ecog_subjects, self.restore_epoch, data_partition='validation',
                                                       assessment_type=assessment_type)
        finally:
            for k, m in target_streams.items():
                m.penalty_scale = saved[k]

    # ---- trainers.py:925-963 ---------------------------------------------------------------------
    def construct_online_predictor(self, subject_index: int = -1):
        """predict(inputs [T, C]) -> sentence with the restored EMA weights: the B = 1 greedy-decode latency path."""
        subject = self.ecog_subjects[subject_index]
        subjects = self.ecog_subjects
        si = subjects.index(subject)
        self.net.prepare_for_prediction(subjects, self.restore_epoch)
        targets_list = subject.data_manifests['decoder_targets'].get_feature_list()

        def predict(inputs: np.ndarray) -> str:
            toks = self.net.predict_tokens(inputs, subnet=si)
            return target_inds_to_sequences(toks[None, None, :], targets_list)[0]
        return predict
