"""Host-side assessment metrics (A12 of SURVEY.md section 8a): decoded indices -> strings -> word error rate.

* ``target_inds_to_sequences`` -- /root/reference/ecog2txt/trainers.py:952-963
* ``wer_vector``               -- utils_jgm.toolbox.wer_vector as used at /root/reference/ecog2txt/subjects.py:546-549
* ``confusion_counts``         -- the `.decoder_confusions` the reference reads at /root/reference/ecog2txt/trainers.py:604-611
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import EOS_token, pad_token


def target_inds_to_sequences(hypotheses, targets_list: Sequence[str], iExample: int = 0) -> List[str]:
    """hypotheses [N, beam, L] -> one string per beam of example iExample: tokens joined, '_' -> ' ',
    <pad>/<EOS> stripped, right-stripped."""
    out = []
    for hyp in hypotheses[iExample]:
        s = ''.join(targets_list[int(i)] for i in hyp)
        out.append(s.replace('_', ' ').replace(pad_token, '').replace(EOS_token, '').rstrip())
    return out


def word_error_rate(ref_words: Sequence[str], hyp_words: Sequence[str]) -> float:
    """word-level Levenshtein distance / reference length."""
    n, m = len(ref_words), len(hyp_words)
    d = list(range(m + 1))
    for i in range(1, n + 1):
        prev, d[0] = d[0], i
        for j in range(1, m + 1):
            cur = d[j]
            d[j] = min(d[j] + 1, d[j - 1] + 1, prev + (ref_words[i - 1] != hyp_words[j - 1]))
            prev = cur
    return d[m] / max(n, 1)


def wer_vector(references: Sequence[str], hypotheses: Sequence[str]) -> np.ndarray:
    return np.asarray([word_error_rate(r.split(), h.split()) for r, h in zip(references, hypotheses)], np.float64)


def confusion_counts(targets: np.ndarray, predictions: np.ndarray, num_classes: int, pad_id: int = 0) -> np.ndarray:
    """[V, V] int64 counts: entry (i, j) = how often target class i was decoded as class j, over the unmasked (non-pad)
    target positions (rows = true token, columns = decoded token -- the axes the reference labels with the class list
    on both sides, trainers.py:606-616).  targets / predictions: integer arrays of equal shape, position-aligned."""
    t = np.asarray(targets).reshape(-1).astype(np.int64)
    p = np.asarray(predictions).reshape(-1).astype(np.int64)
    keep = (t != pad_id) & (p >= 0) & (p < num_classes)
    return np.bincount(t[keep] * num_classes + p[keep], minlength=num_classes * num_classes).reshape(num_classes, num_classes)
