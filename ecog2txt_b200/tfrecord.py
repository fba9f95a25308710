"""TFRecord / tf.train.Example plumbing of the hot path's input contract, bound to the
native library ``libe2t_io.so`` (include/e2t_io.h, csrc/tfrecord_io.cpp) through ctypes.

Mirrors, for this path only, what the reference reaches through TensorFlow and ``tf_helpers``:

* ``TFRecordWriter`` + ``make_feature_example``  -- /root/reference/ecog2txt/data_generators.py:317-326
* ``parse_protobuf_seq2seq_example``            -- /root/reference/ecog2txt/subjects.py:297-302,616-618,
                                                   /root/reference/ecog2txt/trainers.py:891-901
* ``string_seq_to_index_seq``                   -- /root/reference/ecog2txt/subjects.py:344-361

One ``tf.train.Example`` per trial; float streams are stored flattened row-major ``[T, num_features_raw]``
and reshaped on parse; text streams are lists of UTF-8 byte strings (``word_`` tokens).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, Iterator, List, Optional, Sequence

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libe2t_io.so")
ABI_VERSION = 2
_P = C.c_void_p
_U8P = C.POINTER(C.c_uint8)
_SIGNATURES = {
    "e2t_io_last_error": (C.c_char_p, []),
    "e2t_io_abi_version": (C.c_int, []),
    "e2t_io_masked_crc32c": (C.c_uint32, [_P, C.c_uint64]),
    "e2t_tfr_reader_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "e2t_tfr_reader_next": (C.c_int, [_P, C.POINTER(_U8P), C.POINTER(C.c_uint64)]),
    "e2t_tfr_reader_close": (C.c_int, [_P]),
    "e2t_example_find": (C.c_int, [_U8P, C.c_uint64, C.c_char_p, C.POINTER(C.c_int), C.POINTER(_U8P),
                                   C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "e2t_bytes_list_next": (C.c_int, [_U8P, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(_U8P), C.POINTER(C.c_uint64)]),
    "e2t_int64_list_copy": (C.c_int, [_U8P, C.c_uint64, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "e2t_tokens_to_indices": (C.c_int, [_U8P, C.c_uint64, C.POINTER(C.c_char_p), _P, C.c_int32, C.c_int32, C.c_int32,
                                        _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "e2t_tfr_writer_open": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "e2t_tfr_writer_write": (C.c_int, [_P, _P, C.c_uint64]),
    "e2t_tfr_writer_close": (C.c_int, [_P]),
    "e2t_example_builder_new": (C.c_int, [C.POINTER(_P)]),
    "e2t_example_builder_free": (C.c_int, [_P]),
    "e2t_example_builder_reset": (C.c_int, [_P]),
    "e2t_example_builder_add_floats": (C.c_int, [_P, C.c_char_p, _P, C.c_uint64]),
    "e2t_example_builder_add_bytes": (C.c_int, [_P, C.c_char_p, _P, _P, C.c_uint64]),
    "e2t_example_builder_add_int64s": (C.c_int, [_P, C.c_char_p, _P, C.c_uint64]),
    "e2t_example_builder_finish": (C.c_int, [_P, C.POINTER(_U8P), C.POINTER(C.c_uint64)]),
    "e2t_pad_batch_f32": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.POINTER(_P), _P]),
    "e2t_pad_batch_f32_mt": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.POINTER(_P), _P, C.c_int]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


class TFRecordError(RuntimeError):
    pass


def load() -> C.CDLL:
    """The native IO library; raises if it has not been built (no pure-Python fallback in the product)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: run `python __graft_entry__.py` (build) first")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.e2t_io_abi_version() != ABI_VERSION:
            raise RuntimeError("libe2t_io ABI mismatch; rebuild")
        _lib = lib
    return _lib


def _ck(rc):
    if rc < 0:
        raise TFRecordError(load().e2t_io_last_error().decode())
    return rc


def masked_crc32c(data: bytes) -> int:
    return int(load().e2t_io_masked_crc32c(data, len(data)))


# ------------------------------------------------------------------------------------------------
# writing
# ------------------------------------------------------------------------------------------------
class TFRecordWriter:
    """``with TFRecordWriter(path) as w: w.write_example({'ecog_sequence': x[T,C], 'text_sequence': [b'the_', ...]})``"""

    def __init__(self, path: str):
        self._lib = load()
        self._w = _P()
        _ck(self._lib.e2t_tfr_writer_open(os.fsencode(path), C.byref(self._w)))
        self._b = _P()
        _ck(self._lib.e2t_example_builder_new(C.byref(self._b)))

    def write(self, record: bytes):
        _ck(self._lib.e2t_tfr_writer_write(self._w, record, len(record)))

    def write_example(self, example_dict: Dict[str, object]):
        """tfh.make_feature_example(example_dict).SerializeToString() + write: every key becomes one
        feature; float arrays are flattened row-major, string sequences become a BytesList."""
        lib = self._lib
        _ck(lib.e2t_example_builder_reset(self._b))
        for key, val in example_dict.items():
            k = key.encode()
            arr = np.asarray(val)
            if arr.dtype.kind == "f":
                a = np.ascontiguousarray(arr, np.float32).reshape(-1)
                _ck(lib.e2t_example_builder_add_floats(self._b, k, a.ctypes.data_as(_P), a.size))
            elif arr.dtype.kind in "iu":
                a = np.ascontiguousarray(arr, np.int64).reshape(-1)
                _ck(lib.e2t_example_builder_add_int64s(self._b, k, a.ctypes.data_as(_P), a.size))
            else:
                strs = [s if isinstance(s, bytes) else str(s).encode("utf-8") for s in np.asarray(val, object).reshape(-1)]
                blob = b"".join(strs)
                lens = np.asarray([len(s) for s in strs], np.uint64)
                _ck(lib.e2t_example_builder_add_bytes(self._b, k, blob, lens.ctypes.data_as(_P), len(strs)))
        out, n = _U8P(), C.c_uint64()
        _ck(lib.e2t_example_builder_finish(self._b, C.byref(out), C.byref(n)))
        _ck(lib.e2t_tfr_writer_write(self._w, out, n.value))

    def close(self):
        if self._w:
            _ck(self._lib.e2t_tfr_writer_close(self._w))
            self._w = _P()
        if self._b:
            self._lib.e2t_example_builder_free(self._b)
            self._b = _P()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ------------------------------------------------------------------------------------------------
# reading
# ------------------------------------------------------------------------------------------------
class _Vocab:
    """Sorted view of a class list for the native string -> index lookup."""

    def __init__(self, feature_list: Sequence[str]):
        self.source = list(feature_list)      # own copy: keeps the content the table was built from
        enc = [t.encode("utf-8") for t in feature_list]
        order = sorted(range(len(enc)), key=lambda i: enc[i])
        self.n = len(enc)
        self.sorted = (C.c_char_p * self.n)(*[enc[i] for i in order])
        self.ids = np.asarray(order, np.int32)


def iter_records(path: str, check_crc: bool = True) -> Iterator[bytes]:
    """Raw record payloads of one TFRecord file (tf.data.TFRecordDataset)."""
    lib = load()
    r = _P()
    _ck(lib.e2t_tfr_reader_open(os.fsencode(path), int(check_crc), C.byref(r)))
    try:
        data, n = _U8P(), C.c_uint64()
        while _ck(lib.e2t_tfr_reader_next(r, C.byref(data), C.byref(n))) == 1:
            yield C.string_at(data, n.value)
    finally:
        lib.e2t_tfr_reader_close(r)


def parse_example(record: bytes, manifests: Dict[str, "object"], _vocab_cache: Optional[dict] = None) -> Dict[str, np.ndarray]:
    """parse_protobuf_seq2seq_example(example_proto, {data_key: SequenceDataManifest}): for each data_key
    pull feature ``manifest.sequence_type``; floats -> [T, num_features_raw] fp32 (channel mask applied when the
    manifest carries one); strings -> int32 indices [T'] through the manifest's class list (EOS appended when
    APPEND_EOS, OOV -> index of <OOV>, fallback 2), or the raw byte strings when the manifest has no class list."""
    lib = load()
    buf = (C.c_uint8 * len(record)).from_buffer_copy(record)
    out = {}
    cache = _vocab_cache if _vocab_cache is not None else {}
    for data_key, man in manifests.items():
        kind, payload, plen, count = C.c_int(), _U8P(), C.c_uint64(), C.c_uint64()
        _ck(lib.e2t_example_find(buf, len(record), man.sequence_type.encode(), C.byref(kind), C.byref(payload),
                                 C.byref(plen), C.byref(count)))
        if kind.value == 0:
            raise TFRecordError(f"feature '{man.sequence_type}' not present in the Example")
        if kind.value == 2:
            a = np.ctypeslib.as_array(C.cast(payload, C.POINTER(C.c_float)), shape=(count.value,)).copy() \
                if count.value else np.zeros(0, np.float32)
            nraw = int(man.num_features_raw or 1)
            if a.size % nraw:
                raise TFRecordError(f"{man.sequence_type}: {a.size} floats is not a multiple of num_features_raw={nraw}")
            a = a.reshape(-1, nraw)
            out[data_key] = man.apply_transform(a)
        elif kind.value == 1:
            flist = man.get_feature_list() if man.get_feature_list is not None else None
            if flist is None:
                off, s, sl = C.c_uint64(0), _U8P(), C.c_uint64()
                strs = []
                while _ck(lib.e2t_bytes_list_next(payload, plen.value, C.byref(off), C.byref(s), C.byref(sl))) == 1:
                    strs.append(C.string_at(s, sl.value))
                out[data_key] = np.asarray(strs, object).reshape(-1, 1)
            else:
                # one vocabulary per stream.  Keyed by the stream's name and checked against the list's content: a manifest
                # whose get_feature_list() builds a fresh list per call must neither rebuild the table per record nor hit
                # another stream's table through a recycled id()
                v = cache.get(data_key)
                if v is None or v.source != flist:
                    v = cache[data_key] = _Vocab(flist)
                res = np.empty(count.value + 1, np.int32)
                n = C.c_uint64()
                _ck(lib.e2t_tokens_to_indices(payload, plen.value, v.sorted, v.ids.ctypes.data_as(_P), v.n, man.OOV_id,
                                              man.EOS_id if man.APPEND_EOS else -1, res.ctypes.data_as(_P), res.size,
                                              C.byref(n)))
                out[data_key] = res[: n.value].copy()
        else:
            res = np.empty(count.value, np.int64)
            n = C.c_uint64()
            _ck(lib.e2t_int64_list_copy(payload, plen.value, res.ctypes.data_as(_P), res.size, C.byref(n)))
            out[data_key] = res
    return out


def read_examples(paths: Iterable[str], manifests: Dict[str, "object"], check_crc: bool = True):
    """All examples of the given TFRecord files, parsed (the reference's dataset.map(parse...) idiom)."""
    cache: dict = {}
    for p in paths:
        for rec in iter_records(p, check_crc):
            yield parse_example(rec, manifests, cache)


def pad_batch_f32(seqs: List[np.ndarray], T_pad: Optional[int] = None, out: Optional[np.ndarray] = None,
                  threads: int = 1) -> np.ndarray:
    """Zero-padded [B, T_pad, C] batch of variable-length [T_i, C] fp32 sequences (padding_value 0.0).  `out`: write into this
    buffer (e.g. a page-locked staging buffer); `threads` > 1 splits the utterances over native worker threads."""
    lib = load()
    B = len(seqs)
    C_ = int(seqs[0].shape[1])
    lens = np.asarray([s.shape[0] for s in seqs], np.int64)
    T_pad = int(T_pad or lens.max())
    if out is None:
        out = np.empty((B, T_pad, C_), np.float32)
    assert out.shape == (B, T_pad, C_) and out.dtype == np.float32 and out.flags.c_contiguous
    keep = [np.ascontiguousarray(s, np.float32) for s in seqs]
    ptrs = (_P * B)(*[k.ctypes.data for k in keep])
    if threads > 1:
        _ck(lib.e2t_pad_batch_f32_mt(out.ctypes.data_as(_P), B, T_pad, C_, ptrs, lens.ctypes.data_as(_P), int(threads)))
    else:
        _ck(lib.e2t_pad_batch_f32(out.ctypes.data_as(_P), B, T_pad, C_, ptrs, lens.ctypes.data_as(_P)))
    return out
