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
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import EOS_token, OOV_token, pad_token, TOKEN_TYPES
from . import params as prm
from .metrics import target_inds_to_sequences
from .sequence_network import SequenceNetwork


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
        assert token_type in TOKEN_TYPES, 'Unrecognized token_type!! -- jgm'          # trainers.py:64-65
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
        self.net = SequenceNetwork(last, EOS_token=EOS_token, pad_token=pad_token, OOV_token=OOV_token, training_GPUs=[0],
                                   TARGETS_ARE_SEQUENCES='sequence' in token_type, VERBOSE=VERBOSE, **SN_kwargs)
        self.checkpoint_dir = checkpoint_dir
        self.results: List[dict] = []

    # ---- trainers.py:213-252 ---------------------------------------------------------------------
    @property
    def checkpoint_dir(self):
        try:
            self.net.checkpoint_path = os.path.join(self._checkpoint_dir, 'model.ckpt')
        except AttributeError:
            pass
        return self._checkpoint_dir

    @checkpoint_dir.setter
    def checkpoint_dir(self, checkpoint_dir):
        self._checkpoint_dir = checkpoint_dir
        self.checkpoint_dir

    @property
    def restore_epoch(self):
        if self._restore_epoch is not None:
            return self._restore_epoch
        model_name = 'model.ckpt'
        if not os.path.isdir(self.checkpoint_dir):
            return None
        epochs = sorted(int(name.split('-')[1].split('.')[0]) for name in os.listdir(self.checkpoint_dir)
                        if name.split('-')[0] == model_name and name.split('.')[-1] == 'index')
        return epochs[-1] if epochs else None

    @restore_epoch.setter
    def restore_epoch(self, restore_epoch):
        self._restore_epoch = restore_epoch

    def vprint(self, *a, **k):
        if self.VERBOSE:
            print(*a, **k)

    def _save_results(self, assessments):
        self.results.append(assessments)

    # ---- trainers.py:303-327 ---------------------------------------------------------------------
    def parallel_transfer_learn(self, RESUME=False, fit_kwargs=()):
        """All subjects jointly in one fit (one subject per minibatch, shared layers see everyone's data)."""
        if RESUME:
            fit_kwargs = {'_restore_epoch': self.restore_epoch, **dict(fit_kwargs),
                          'train_vars_scope': 'seq2seq', 'reuse_vars_scope': 'seq2seq'}
            self.ecog_subjects = [self.ecog_subjects[-1]]
        assessments = self.net.fit(self.ecog_subjects, **dict(fit_kwargs))
        self._save_results(assessments)
        if self._restore_epoch is not None:
            self.restore_epoch = self.restore_epoch + self.net.N_epochs if RESUME else self.net.N_epochs
        return assessments

    # ---- trainers.py:329-374 ---------------------------------------------------------------------
    def sequential_transfer_learn(self, pretraining_epochs=60, training_epochs=200, posttraining_epochs=340):
        proprietary_scopes = 'seq2seq/subnet'
        reusable_scopes = 'seq2seq/(?!subnet)'  # negative lookahead
        fit_kwargs: Dict = {}
        latest_epoch = 0
        assessments = None
        for subject in self.ecog_subjects:
            if subject is self.ecog_subjects[0]:
                latest_epoch = 0
                fit_kwargs['reuse_vars_scope'] = None
            else:
                # first acquire this subject's encoder embedding with everything shared frozen
                self.net.N_epochs = pretraining_epochs
                fit_kwargs['train_vars_scope'] = proprietary_scopes
                fit_kwargs['reuse_vars_scope'] = reusable_scopes
                fit_kwargs['_restore_epoch'] = latest_epoch
                self.net.fit([subject], **fit_kwargs)
                latest_epoch += self.net.N_epochs
                fit_kwargs['_restore_epoch'] = latest_epoch
                fit_kwargs['reuse_vars_scope'] = 'seq2seq'
            if subject is self.ecog_subjects[-1]:
                training_epochs += posttraining_epochs
            self.net.N_epochs = training_epochs
            fit_kwargs['train_vars_scope'] = 'seq2seq'
            assessments = self.net.fit([subject], **fit_kwargs)
            latest_epoch += self.net.N_epochs
            self._save_results(assessments)
        self.restore_epoch = latest_epoch
        return assessments

    # ---- trainers.py:376-408 ---------------------------------------------------------------------
    def assess_saved_model(self):
        self.update_net_from_saved_model()
        return self.net.restore_and_assess(self.ecog_subjects, self.restore_epoch)

    def update_net_from_saved_model(self):
        self.net.layer_sizes, data_sizes, strides, EMA = self.recover_model_sizes()
        self.net.TEMPORALLY_CONVOLVE = len(strides)
        self.net.EMA_decay = 0.99 * EMA
        for subject in self.ecog_subjects:
            s_id = subject.subnet_id
            manifests = subject.data_manifests
            for key, data_size in data_sizes.get(s_id, {}).items():
                if key in manifests:
                    manifests[key].num_features = data_size
            for key, data_size in data_sizes.get(None, {}).items():
                if key in manifests:
                    manifests[key].num_features = data_size
            if strides.get(s_id):
                subject.decimation_factor = int(np.prod(strides[s_id]))

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
        var_to_shape = prm.variable_to_shape_map(self.net.checkpoint_path, self.restore_epoch)
        weights_full_name = None
        for key in sorted(var_to_shape):
            if re.match('.*{0}.*0/weights/ExponentialMovingAverage'.format(weights_name), key):
                weights_full_name = key
        assert weights_full_name, "Uh-oh, no such weights found! -- jgm"
        return self.net.get_weights_as_numpy_array(weights_full_name, self.restore_epoch)

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
        old_penalties = {}
        for key, manifest in subject.data_manifests.items():
            if '_targets' in key:
                old_penalties[key] = manifest.penalty_scale
                manifest.penalty_scale = 0.0
        key = contrib_method.replace('saliency_map', 'targets')
        subject.data_manifests[key].penalty_scale = 1.0
        try:
            return self.net.restore_and_get_saliencies([subject] if len(self.ecog_subjects) == 1 else self.ecog_subjects,
                                                       self.restore_epoch, data_partition='validation',
                                                       assessment_type=assessment_type)
        finally:
            for key, manifest in subject.data_manifests.items():
                if '_targets' in key:
                    manifest.penalty_scale = old_penalties[key]

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
