"""Seeded synthetic stand-in for ECoGDataGenerator._ecog_token_generator
(/root/reference/ecog2txt/data_generators.py:515-530 yields nothing; real loaders are lab-private).

Each of `n_sentences` fixed sentences (4-10 words drawn from the vocabulary, as in the 50-sentence
MOCHA-TIMIT setting, mocha-1_word_sequence.yaml:74) owns a random [T, C] template; an utterance is
template + unit Gaussian noise, so the task is learnable.  Frames are never exactly zero inside
the valid length, so length inference from zero padding (trainers.py:806-807) is well defined.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

from . import EOS_token, OOV_token, pad_token

AUX_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "auxiliary")


def load_vocab(path: Optional[str] = None, size: int = 1806) -> List[str]:
    """One token per line, <pad>,<EOS>,<OOV> first (get_class_list, data_generators.py:427-444).
    Without a file, a synthetic vocabulary of the same shape ('w0000_' ...) is produced."""
    if path is not None:
        with open(path) as f:
            return f.read().split()
    return [pad_token, EOS_token, OOV_token] + [f"w{i:04d}_" for i in range(size - 3)]


@dataclass
class SyntheticCorpus:
    vocab: List[str]
    n_sentences: int = 50
    T: int = 400
    C: int = 256
    ragged: bool = False
    seed: int = 0
    min_words: int = 4
    max_words: int = 10
    template_scale: float = 1.0

    def __post_init__(self):
        rs = np.random.RandomState(self.seed)
        V = len(self.vocab)
        self.sentences: List[np.ndarray] = []
        for _ in range(self.n_sentences):
            n = rs.randint(self.min_words, self.max_words + 1)
            self.sentences.append(rs.randint(3, V, size=n).astype(np.int32))
        self._tseed = rs.randint(0, 2 ** 31 - 1, size=self.n_sentences)
        self.max_target_len = self.max_words + 1  # + <EOS>

    def template(self, s: int) -> np.ndarray:
        return np.random.RandomState(self._tseed[s]).randn(self.T, self.C).astype(np.float32) * self.template_scale

    def utterance(self, s: int, rs: np.random.RandomState) -> Tuple[np.ndarray, int]:
        T = self.T
        n = int(rs.randint(T // 2, T * 3 // 2 + 1)) if self.ragged else T
        n = min(n, 1250)  # max_samples = 200 Hz * 6.25 s (data_generators.py:35-42)
        base = self.template(s)
        if n <= T:
            x = base[:n].copy()
        else:
            x = np.concatenate([base, base[: n - T]], axis=0)
        x += rs.randn(n, self.C).astype(np.float32)
        x[x == 0.0] = 1e-6
        return x, n

    def words(self, s: int) -> List[str]:
        return [self.vocab[i] for i in self.sentences[s]]

    def batch(self, B: int, seed: int, L: Optional[int] = None, T_pad: Optional[int] = None):
        """Padded batch dict like the parsed TFRecords: encoder_inputs [B,Tmax,C] f32 (0-padded),
        decoder_targets [B,L] int32 (EOS appended, pad after), lengths [B], sentence ids [B]."""
        rs = np.random.RandomState(seed)
        ids = rs.randint(0, self.n_sentences, size=B)
        xs, ns = zip(*(self.utterance(int(s), rs) for s in ids))
        Tmax = T_pad or max(ns)
        x = np.zeros((B, Tmax, self.C), np.float32)
        for b, (xb, n) in enumerate(zip(xs, ns)):
            x[b, :n] = xb
        L = L or self.max_target_len
        y = np.zeros((B, L), np.int32)
        for b, s in enumerate(ids):
            w = self.sentences[int(s)]
            y[b, : len(w)] = w
            y[b, len(w)] = 1
        return {"encoder_inputs": x, "decoder_targets": y, "lengths": np.asarray(ns, np.int32), "sentence_ids": ids}
