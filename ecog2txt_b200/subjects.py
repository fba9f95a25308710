"""The "subnet_params" protocol the hot path reads from its callers, restated as plain classes.

* ``SequenceDataManifest`` -- /root/reference/ecog2txt/subjects.py:274-404 (same attribute names and defaulting
  rules: num_features / num_features_raw / transform / padding_value / distribution / penalty_scale / APPEND_EOS).
* ``ECoGSubject``          -- the attributes a SequenceNetwork consumes (/root/reference/ecog2txt/subjects.py:49-68):
  subnet_id, block_ids, decimation_factor, tf_record_partial_path, data_manifests; block partitioning follows
  subjects.py:110-138 (default_dataset + block_types, pretrain_all_blocks).
* ``SyntheticDataGenerator`` plays the role of ``ECoGDataGenerator`` for this path only: it writes one TFRecord
  per block (data_generators.py:317-326,382-425) from the seeded synthetic corpus -- the reference's real
  loaders are lab-private and its default generator yields nothing (data_generators.py:515-530).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Set

import numpy as np

from . import DATA_PARTITIONS, EOS_token, OOV_token, pad_token
from . import tfrecord
from .synthetic import SyntheticCorpus


class SequenceDataManifest:
    def __init__(self, sequence_type, num_features=None, num_features_raw=None, transform=None, padding_value=None,
                 penalty_scale=1.0, distribution=None, mask=None, get_feature_list=None, APPEND_EOS=False):
        self.sequence_type = sequence_type
        self._num_features = num_features
        self._num_features_raw = num_features_raw
        self._transform = transform
        self._padding_value = padding_value
        self.penalty_scale = penalty_scale
        self._distribution = distribution
        self.mask = mask                      # object with .inds (channel sub-selection), subjects.py:342-343
        self.get_feature_list = get_feature_list
        self.APPEND_EOS = APPEND_EOS

    # -- sizes (subjects.py:304-336)
    @property
    def num_features(self):
        if self.mask is not None:
            return len(self.mask.inds)
        if self.get_feature_list is not None:
            return len(self.get_feature_list())
        return self._num_features

    @num_features.setter
    def num_features(self, v):
        self._num_features = v

    @property
    def num_features_raw(self):
        if self._num_features_raw is not None:
            return self._num_features_raw
        if self.mask is not None:
            return self._num_features
        if self.get_feature_list is not None:
            return 1
        return self.num_features

    @num_features_raw.setter
    def num_features_raw(self, v):
        self._num_features_raw = v

    # -- parse-time transforms (subjects.py:338-367)
    @property
    def OOV_id(self) -> int:
        fl = self.get_feature_list()
        return fl.index(OOV_token) if OOV_token in fl else 2

    @property
    def EOS_id(self) -> int:
        return self.get_feature_list().index(EOS_token)

    def apply_transform(self, seq: np.ndarray) -> np.ndarray:
        """float streams: user transform, else channel mask, else identity."""
        if self._transform is not None:
            return self._transform(seq)
        if self.mask is not None:
            return seq[:, np.asarray(self.mask.inds)]
        return seq

    @property
    def distribution(self):
        if self._distribution is not None:
            return self._distribution
        if self.sequence_type == 'ecog_sequence':
            return 'Rayleigh'
        if self.sequence_type == 'audio_sequence':
            return 'Gaussian'
        return 'categorical'

    @distribution.setter
    def distribution(self, v):
        self._distribution = v

    @property
    def padding_value(self):
        if self._padding_value is None:
            if self.get_feature_list is None:
                return 0.0
            fl = self.get_feature_list()
            return fl.index(pad_token) if pad_token in fl else 0
        return self._padding_value

    @padding_value.setter
    def padding_value(self, v):
        self._padding_value = v


def sentence_tokenize(token_list: Sequence[str], token_type: str = 'word_sequence') -> List[bytes]:
    """Words of one trial -> the byte strings stored in the TFRecord (ECoGDataGenerator._sentence_tokenize,
    /root/reference/ecog2txt/data_generators.py:445-473): lower-cased, underscore-postfixed, UTF-8; a 'trial' token type joins
    the words of the sentence into ONE token.  ('word_piece_sequence' needs tensor2tensor's SubwordTextEncoder: out of scope.)"""
    if token_type == 'word_piece_sequence':
        raise NotImplementedError("word-piece tokenisation needs tensor2tensor's SubwordTextEncoder (out of scope)")
    if token_type == 'trial':
        return [' '.join(token.lower() + '_' for token in token_list).encode('utf-8')]
    return [(token.lower() + '_').encode('utf-8') for token in token_list]


def get_class_list(vocab_file_path: str) -> List[str]:
    """The class list of a text stream = the whitespace-separated entries of its vocabulary file
    (ECoGDataGenerator.get_class_list, data_generators.py:428-435; e.g. auxiliary/vocab.mocha-timit.1806)."""
    with open(vocab_file_path, 'r') as f:
        return f.read().split()


class SyntheticDataGenerator:
    """Writes `<tf_record_partial_path>.format(block)` for each block from a SyntheticCorpus
    (ECoGDataGenerator.write_to_Protobuf_maybe, data_generators.py:382-425)."""

    sampling_rate = 200          # mochastar_word_sequence.yaml:84

    def __init__(self, corpus: SyntheticCorpus, tf_record_partial_path: str, utterances_per_block: int = 50,
                 seed: int = 0, audio_features: int = 0, n_phonemes: int = 0):
        self.corpus = corpus
        self.tf_record_partial_path = tf_record_partial_path
        self.utterances_per_block = utterances_per_block
        self.seed = seed
        self.num_ECoG_channels = corpus.C
        # optional frame-rate streams for the encoder-targets head (canonical keys audio_sequence / phoneme_sequence,
        # data_generators.py:515-530): fixed random read-outs of the ECoG frame, so that the head is learnable
        self.audio_features, self.n_phonemes = int(audio_features), int(n_phonemes)
        rs = np.random.RandomState(seed + 7919)
        self._audio_map = rs.randn(corpus.C, max(self.audio_features, 1)).astype(np.float32) / np.sqrt(corpus.C)
        self._phoneme_map = rs.randn(corpus.C, max(self.n_phonemes, 1)).astype(np.float32)

    @property
    def phoneme_list(self) -> List[str]:
        """class list of the phoneme stream; index 0 is the pad class"""
        return [pad_token] + [f"ph{i}" for i in range(1, self.n_phonemes)]

    def _ecog_token_generator(self, block: int):
        rs = np.random.RandomState(self.seed * 100003 + int(block))
        for _ in range(self.utterances_per_block):
            s = int(rs.randint(0, self.corpus.n_sentences))
            x, _n = self.corpus.utterance(s, rs)
            ex = {'ecog_sequence': x, 'text_sequence': [w.encode('utf-8') for w in self.corpus.words(s)]}
            if self.audio_features:
                ex['audio_sequence'] = (x @ self._audio_map).astype(np.float32)
            if self.n_phonemes:
                cls = 1 + np.argmax(x @ self._phoneme_map[:, 1:], axis=1)
                ex['phoneme_sequence'] = [self.phoneme_list[c].encode('utf-8') for c in cls]
            yield ex

    def _write_to_Protobuf(self, block: int):
        path = self.tf_record_partial_path.format(block)
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with tfrecord.TFRecordWriter(path) as w:
            for example_dict in self._ecog_token_generator(block):
                w.write_example(example_dict)

    def write_to_Protobuf_maybe(self, sequence_type: str, block_set: Iterable[int]) -> List[str]:
        man = SequenceDataManifest(sequence_type, num_features_raw=1)
        targets: Set[str] = set()
        for block in block_set:
            path = self.tf_record_partial_path.format(block)
            if not os.path.exists(path):
                self._write_to_Protobuf(block)
            for ex in tfrecord.read_examples([path], {'seq': man}):
                targets.update(w.decode('utf-8') for w in ex['seq'][:, 0])
        return list(targets)


class ECoGSubject:
    """Attribute bag read by SequenceNetwork.fit / restore_and_assess (subjects.py:49-68)."""

    def __init__(self, subj_id: int, data_generator, block_dict: Dict[int, dict], data_manifests: Dict[str, SequenceDataManifest],
                 block_types: Optional[Dict[str, Set[str]]] = None, pretrain_all_blocks: bool = False,
                 decimation_factor: Optional[int] = None, sampling_rate_decimated: float = 16.5,
                 block_ids: Optional[Dict[str, Set[int]]] = None):
        self.subj_id = subj_id
        self.data_generator = data_generator
        self._block_dict = block_dict
        self.block_types = block_types or {p: {'mocha'} for p in DATA_PARTITIONS}
        self.pretrain_all_blocks = pretrain_all_blocks
        self._decimation_factor = decimation_factor
        self.sampling_rate_decimated = sampling_rate_decimated
        self._block_ids = block_ids
        self._data_manifests = data_manifests
        self.target_specs = {}
        self.input_mask = None

    @property
    def subnet_id(self):
        return self.subj_id

    @property
    def block_ids(self) -> Dict[str, Set[int]]:
        if self._block_ids:
            return self._block_ids
        ids = {part: {blk for blk, info in self._block_dict.items()
                      if info['default_dataset'] == part and info['type'] in self.block_types[part]}
               for part in DATA_PARTITIONS}
        if self.pretrain_all_blocks:
            ids['training'] = {blk for blks in ids.values() for blk in blks}
        return ids

    @block_ids.setter
    def block_ids(self, v):
        self._block_ids = v

    @property
    def tf_record_partial_path(self):
        return self.data_generator.tf_record_partial_path

    @property
    def decimation_factor(self) -> int:
        if self._decimation_factor is None:
            return int(np.round(self.data_generator.sampling_rate / self.sampling_rate_decimated))
        return self._decimation_factor

    @decimation_factor.setter
    def decimation_factor(self, v):
        self._decimation_factor = v

    @property
    def data_manifests(self):
        for man in self._data_manifests.values():
            if man.sequence_type == 'ecog_sequence':
                man.num_features = self.data_generator.num_ECoG_channels
        return self._data_manifests

    def write_tf_records_maybe(self, sequence_type=None, data_partitions=DATA_PARTITIONS):
        if sequence_type is None:
            sequence_type = self.data_manifests['decoder_targets'].sequence_type
        class_list = []
        for part in data_partitions:
            class_list = self.data_generator.write_to_Protobuf_maybe(sequence_type, self.block_ids[part])
        return class_list


def make_synthetic_subject(subj_id: int, vocab: Sequence[str], out_dir: str, n_train_blocks: int = 4,
                           n_valid_blocks: int = 1, utterances_per_block: int = 50, T: int = 400, C: int = 256,
                           n_sentences: int = 50, ragged: bool = True, seed: int = 0,
                           pretrain_all_blocks: bool = False, encoder_targets: Optional[str] = None,
                           encoder_targets_features: int = 13, encoder_targets_layer: int = 1,
                           encoder_targets_penalty_scale: float = 1.0) -> ECoGSubject:
    """A subject in the shape of block_breakdowns.json + the minimal data_mapping of the README
    ({'decoder_targets': 'text_sequence', 'encoder_inputs': 'ecog_sequence'}, /root/reference/README.md:61)."""
    corpus = SyntheticCorpus(list(vocab), n_sentences=n_sentences, T=T, C=C, ragged=ragged, seed=seed)
    path = os.path.join(out_dir, f"EFC{subj_id}_B{{0}}.tfrecord")   # mochastar_word_sequence.yaml:90
    if encoder_targets not in (None, 'audio_sequence', 'phoneme_sequence'):
        raise ValueError("encoder_targets must be None, 'audio_sequence' or 'phoneme_sequence'")
    gen = SyntheticDataGenerator(corpus, path, utterances_per_block, seed=seed,
                                 audio_features=encoder_targets_features if encoder_targets == 'audio_sequence' else 0,
                                 n_phonemes=encoder_targets_features if encoder_targets == 'phoneme_sequence' else 0)
    blocks = {b + 1: {'type': 'mocha', 'default_dataset': 'training'} for b in range(n_train_blocks)}
    for b in range(n_valid_blocks):
        blocks[n_train_blocks + b + 1] = {'type': 'mocha', 'default_dataset': 'validation'}
    vocab_list = list(vocab)
    manifests = {
        'encoder_inputs': SequenceDataManifest('ecog_sequence', num_features=C),
        'decoder_targets': SequenceDataManifest('text_sequence', get_feature_list=lambda: vocab_list, APPEND_EOS=True),
    }
    key = f'encoder_{encoder_targets_layer}_targets'      # 'encoder_1_targets' in the reference (trainers.py:791-799)
    if encoder_targets == 'audio_sequence':
        manifests[key] = SequenceDataManifest('audio_sequence', num_features=encoder_targets_features,
                                              num_features_raw=encoder_targets_features,
                                              penalty_scale=encoder_targets_penalty_scale)
    elif encoder_targets == 'phoneme_sequence':
        plist = gen.phoneme_list
        manifests[key] = SequenceDataManifest('phoneme_sequence', get_feature_list=lambda: plist,
                                              penalty_scale=encoder_targets_penalty_scale)
    return ECoGSubject(subj_id, gen, blocks, manifests, pretrain_all_blocks=pretrain_all_blocks)
