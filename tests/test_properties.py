"""CPU: property tests (hypothesis) of host-side invariants around the hot path."""
import numpy as np
from hypothesis import given, settings, strategies as st

from ecog2txt_b200 import tfrecord
from ecog2txt_b200.dist import shard_range
from ecog2txt_b200.metrics import wer_vector


@given(st.integers(0, 5000), st.integers(1, 64))
def test_shard_range_partitions_the_items(n, world):
    """contiguous, disjoint, covering shards whose sizes differ by at most one (data-parallel minibatch / decode sharding)"""
    parts = [shard_range(n, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    sizes = [hi - lo for lo, hi in parts]
    assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes


@settings(max_examples=30, deadline=None)
@given(st.lists(st.integers(0, 12), min_size=1, max_size=9), st.integers(1, 5), st.integers(1, 6))
def test_padded_batch_assembly(lens, C, threads):
    """pad_batch_f32 (single- and multi-threaded): utterance i lands in out[i, :len_i], everything after it is zero"""
    rs = np.random.RandomState(sum(lens) + C)
    seqs = [rs.randn(n, C).astype(np.float32) for n in lens]
    T = max(max(lens), 1)
    out = tfrecord.pad_batch_f32(seqs, T, threads=threads)
    assert out.shape == (len(lens), T, C)
    for i, s in enumerate(seqs):
        assert np.array_equal(out[i, :len(s)], s) and not out[i, len(s):].any()


def _edit_distance(a, b):
    d = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        prev, d[0] = d[0], i
        for j, y in enumerate(b, 1):
            prev, d[j] = d[j], min(d[j] + 1, d[j - 1] + 1, prev + (x != y))
    return d[-1]


_words = st.lists(st.sampled_from(["the", "birch", "canoe", "slid", "on", "smooth", "planks"]), min_size=1, max_size=8)


@settings(max_examples=50, deadline=None)
@given(st.lists(st.tuples(_words, st.lists(st.sampled_from(["the", "birch", "canoe", "glue", "sheet"]), max_size=8)),
                min_size=1, max_size=5))
def test_wer_vector_is_word_edit_distance_over_reference_length(pairs):
    """utils_jgm.toolbox.wer_vector as used at subjects.py:546-549: word-level Levenshtein distance / reference length"""
    refs = [" ".join(r) for r, _ in pairs]
    hyps = [" ".join(h) for _, h in pairs]
    got = wer_vector(refs, hyps)
    want = [_edit_distance(r, h) / len(r) for r, h in pairs]
    assert np.allclose(got, want)


@given(st.binary(max_size=200))
def test_masked_crc32c_matches_the_definition(data):
    """TFRecord framing: masked crc = ((crc >> 15) | (crc << 17)) + 0xa282ead8 over CRC-32C (Castagnoli), bitwise reference"""
    crc = 0xFFFFFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    crc ^= 0xFFFFFFFF
    masked = (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF
    assert tfrecord.masked_crc32c(data) == masked
