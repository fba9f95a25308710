// tfrecord_io.cpp -- native TFRecord framing + tf.train.Example decoding for the batch-assembly row
// (A1 / N1 of SURVEY.md section 8): the on-disk input contract of the hot path.
//
// Replaces, for this path only, what the reference reaches through TensorFlow:
//   writer : tf.io.TFRecordWriter + tfh.make_feature_example   (/root/reference/ecog2txt/data_generators.py:317-326)
//   reader : tf.data.TFRecordDataset + tf.io.parse_single_example with VarLenFeature(float32|string)
//            (/root/reference/ecog2txt/subjects.py:297-302,616-618; trainers.py:891-901)
// Format: little-endian  u64 len | u32 masked_crc32c(len) | bytes | u32 masked_crc32c(bytes);
// payload = Example{ features = 1: Features{ feature = 1: map<string, Feature{ bytes_list=1 | float_list=2 | int64_list=3 }> } }.
// C-ABI, host only (no CUDA): declared in include/e2t_io.h, bound by ecog2txt_b200/tfrecord.py via ctypes.
#include "../../include/e2t_io.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_io_err;

// ---- CRC32C (Castagnoli), slicing-by-8 ---------------------------------------------------------
uint32_t g_tab[8][256];
bool g_tab_ready = false;
void crc_init() {
  if (g_tab_ready) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
    g_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_tab[t][i] = (g_tab[t - 1][i] >> 8) ^ g_tab[0][g_tab[t - 1][i] & 0xFF];
  g_tab_ready = true;
}
uint32_t crc32c(const uint8_t* p, size_t n) {
  crc_init();
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = g_tab[7][lo & 0xFF] ^ g_tab[6][(lo >> 8) & 0xFF] ^ g_tab[5][(lo >> 16) & 0xFF] ^ g_tab[4][lo >> 24] ^
        g_tab[3][hi & 0xFF] ^ g_tab[2][(hi >> 8) & 0xFF] ^ g_tab[1][(hi >> 16) & 0xFF] ^ g_tab[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = g_tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
inline uint32_t mask_crc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xA282EAD8u; }

// ---- protobuf wire helpers -----------------------------------------------------------------------
struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  bool more() const { return ok && p < end; }
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (p < end && shift < 64) {
      uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
    }
    ok = false;
    return 0;
  }
  // length-delimited field: returns a sub-cursor and advances
  Cursor sub() {
    uint64_t n = varint();
    Cursor c{p, p, ok};
    if (!ok || n > (uint64_t)(end - p)) { ok = false; c.ok = false; return c; }
    c.end = p + n;
    p += n;
    return c;
  }
  void skip(uint32_t wire) {
    switch (wire) {
      case 0: varint(); break;
      case 1: if (end - p >= 8) p += 8; else ok = false; break;
      case 2: sub(); break;
      case 5: if (end - p >= 4) p += 4; else ok = false; break;
      default: ok = false;
    }
  }
};

void put_varint(std::string& s, uint64_t v) {
  while (v >= 0x80) { s.push_back((char)(v | 0x80)); v >>= 7; }
  s.push_back((char)v);
}
void put_len_field(std::string& s, uint32_t field, const std::string& body) {
  put_varint(s, (field << 3) | 2);
  put_varint(s, body.size());
  s += body;
}

}  // namespace

struct e2t_tfr_reader {
  FILE* f = nullptr;
  std::vector<uint8_t> buf;
  int check_crc = 1;
};
struct e2t_tfr_writer {
  FILE* f = nullptr;
};
struct e2t_example_builder {
  std::string features;  // serialized repeated map entries of Features.feature
};

extern "C" {

const char* e2t_io_last_error(void) { return g_io_err.c_str(); }
int e2t_io_abi_version(void) { return 2; }

uint32_t e2t_io_masked_crc32c(const void* data, uint64_t n) { return mask_crc(crc32c((const uint8_t*)data, (size_t)n)); }

// ---- reader ---------------------------------------------------------------------------------------
int e2t_tfr_reader_open(const char* path, int check_crc, e2t_tfr_reader** out) {
  if (!path || !out) { g_io_err = "e2t_io: NULL argument"; return -1; }
  FILE* f = fopen(path, "rb");
  if (!f) { g_io_err = std::string("e2t_io: cannot open '") + path + "'"; return -1; }
  e2t_tfr_reader* r = new e2t_tfr_reader();
  r->f = f;
  r->check_crc = check_crc;
  *out = r;
  return 0;
}
int e2t_tfr_reader_close(e2t_tfr_reader* r) {
  if (!r) return 0;
  if (r->f) fclose(r->f);
  delete r;
  return 0;
}
// returns 1 = record read (*data valid until the next call), 0 = clean end of file, <0 = error
int e2t_tfr_reader_next(e2t_tfr_reader* r, const uint8_t** data, uint64_t* len) {
  if (!r || !data || !len) { g_io_err = "e2t_io: NULL argument"; return -1; }
  uint8_t hdr[12];
  size_t got = fread(hdr, 1, 12, r->f);
  if (got == 0) return 0;
  if (got != 12) { g_io_err = "e2t_io: truncated record header"; return -2; }
  uint64_t n;
  uint32_t crc_len;
  memcpy(&n, hdr, 8);
  memcpy(&crc_len, hdr + 8, 4);
  if (r->check_crc && mask_crc(crc32c(hdr, 8)) != crc_len) { g_io_err = "e2t_io: length CRC mismatch (corrupt TFRecord)"; return -3; }
  if (n > (1ull << 34)) { g_io_err = "e2t_io: implausible record length"; return -3; }
  r->buf.resize((size_t)n + 4);
  if (fread(r->buf.data(), 1, (size_t)n + 4, r->f) != (size_t)n + 4) { g_io_err = "e2t_io: truncated record body"; return -2; }
  uint32_t crc_data;
  memcpy(&crc_data, r->buf.data() + n, 4);
  if (r->check_crc && mask_crc(crc32c(r->buf.data(), (size_t)n)) != crc_data) { g_io_err = "e2t_io: data CRC mismatch (corrupt TFRecord)"; return -3; }
  *data = r->buf.data();
  *len = n;
  return 1;
}

// ---- Example decoding -------------------------------------------------------------------------------
// Finds feature `key` in a serialized tf.train.Example.  kind: 0 = absent, 1 = bytes_list, 2 = float_list,
// 3 = int64_list.  For float_list with packed encoding *payload points at count little-endian floats
// (zero-copy); for bytes_list / int64_list *payload / *payload_len delimit the list message, to be walked
// with e2t_bytes_list_next / e2t_int64_list_copy.  count = number of elements.
int e2t_example_find(const uint8_t* rec, uint64_t len, const char* key, int* kind, const uint8_t** payload,
                     uint64_t* payload_len, uint64_t* count) {
  if (!rec || !key || !kind || !payload || !payload_len || !count) { g_io_err = "e2t_io: NULL argument"; return -1; }
  *kind = 0; *payload = nullptr; *payload_len = 0; *count = 0;
  const size_t klen = strlen(key);
  Cursor ex{rec, rec + len};
  while (ex.more()) {
    uint64_t tag = ex.varint();
    if ((tag >> 3) == 1 && (tag & 7) == 2) {           // Example.features
      Cursor fs = ex.sub();
      while (fs.more()) {
        uint64_t t2 = fs.varint();
        if ((t2 >> 3) == 1 && (t2 & 7) == 2) {         // Features.feature map entry
          Cursor ent = fs.sub();
          bool match = false;
          Cursor val{nullptr, nullptr};
          bool have_val = false;
          while (ent.more()) {
            uint64_t t3 = ent.varint();
            if ((t3 >> 3) == 1 && (t3 & 7) == 2) {     // key
              Cursor k = ent.sub();
              match = k.ok && (size_t)(k.end - k.p) == klen && memcmp(k.p, key, klen) == 0;
            } else if ((t3 >> 3) == 2 && (t3 & 7) == 2) {  // value = Feature
              val = ent.sub();
              have_val = val.ok;
            } else ent.skip((uint32_t)(t3 & 7));
          }
          if (!ent.ok) { g_io_err = "e2t_io: malformed Features.feature entry"; return -2; }
          if (match && have_val) {
            while (val.more()) {
              uint64_t t4 = val.varint();
              uint32_t field = (uint32_t)(t4 >> 3);
              if ((t4 & 7) == 2 && field >= 1 && field <= 3) {
                Cursor lst = val.sub();
                if (!lst.ok) { g_io_err = "e2t_io: malformed Feature"; return -2; }
                *kind = (int)field;
                *payload = lst.p;
                *payload_len = (uint64_t)(lst.end - lst.p);
                if (field == 2) {
                  // FloatList.value = 1: packed (wire 2) is what every writer emits; unpacked (wire 5) handled by copy API
                  Cursor fl = lst;
                  if (fl.more()) {
                    uint64_t t5 = fl.varint();
                    if ((t5 >> 3) == 1 && (t5 & 7) == 2) {
                      Cursor pk = fl.sub();
                      if (!pk.ok) { g_io_err = "e2t_io: malformed FloatList"; return -2; }
                      if (!fl.more()) {   // a single packed run: zero-copy
                        *payload = pk.p;
                        *payload_len = (uint64_t)(pk.end - pk.p);
                        *count = *payload_len / 4;
                        return 0;
                      }
                    }
                  } else { *count = 0; *payload_len = 0; return 0; }
                  g_io_err = "e2t_io: FloatList is not a single packed run";
                  return -3;
                } else if (field == 1) {
                  Cursor bl = lst;
                  uint64_t n = 0;
                  while (bl.more()) {
                    uint64_t t5 = bl.varint();
                    if ((t5 >> 3) == 1 && (t5 & 7) == 2) { bl.sub(); ++n; } else bl.skip((uint32_t)(t5 & 7));
                  }
                  if (!bl.ok) { g_io_err = "e2t_io: malformed BytesList"; return -2; }
                  *count = n;
                  return 0;
                } else {
                  Cursor il = lst;
                  uint64_t n = 0;
                  while (il.more()) {
                    uint64_t t5 = il.varint();
                    if ((t5 >> 3) == 1 && (t5 & 7) == 2) { Cursor pk = il.sub(); while (pk.more()) { pk.varint(); ++n; } }
                    else if ((t5 >> 3) == 1 && (t5 & 7) == 0) { il.varint(); ++n; }
                    else il.skip((uint32_t)(t5 & 7));
                  }
                  if (!il.ok) { g_io_err = "e2t_io: malformed Int64List"; return -2; }
                  *count = n;
                  return 0;
                }
              } else val.skip((uint32_t)(t4 & 7));
            }
            return 0;  // Feature present but empty oneof
          }
        } else fs.skip((uint32_t)(t2 & 7));
      }
      if (!fs.ok) { g_io_err = "e2t_io: malformed Features"; return -2; }
    } else ex.skip((uint32_t)(tag & 7));
  }
  if (!ex.ok) { g_io_err = "e2t_io: malformed Example"; return -2; }
  return 0;
}

// Walks a BytesList payload: *offset starts at 0; returns 1 and the next string, 0 at the end.
int e2t_bytes_list_next(const uint8_t* payload, uint64_t payload_len, uint64_t* offset, const uint8_t** str,
                        uint64_t* str_len) {
  if (!payload || !offset || !str || !str_len) { g_io_err = "e2t_io: NULL argument"; return -1; }
  Cursor c{payload + *offset, payload + payload_len};
  while (c.more()) {
    uint64_t t = c.varint();
    if ((t >> 3) == 1 && (t & 7) == 2) {
      Cursor s = c.sub();
      if (!s.ok) break;
      *str = s.p;
      *str_len = (uint64_t)(s.end - s.p);
      *offset = (uint64_t)(c.p - payload);
      return 1;
    }
    c.skip((uint32_t)(t & 7));
  }
  if (!c.ok) { g_io_err = "e2t_io: malformed BytesList"; return -2; }
  return 0;
}

int e2t_int64_list_copy(const uint8_t* payload, uint64_t payload_len, int64_t* out, uint64_t cap, uint64_t* n_out) {
  if (!payload || !out || !n_out) { g_io_err = "e2t_io: NULL argument"; return -1; }
  Cursor il{payload, payload + payload_len};
  uint64_t n = 0;
  while (il.more()) {
    uint64_t t = il.varint();
    if ((t >> 3) == 1 && (t & 7) == 2) { Cursor pk = il.sub(); while (pk.more()) { uint64_t v = pk.varint(); if (n < cap) out[n] = (int64_t)v; ++n; } }
    else if ((t >> 3) == 1 && (t & 7) == 0) { uint64_t v = il.varint(); if (n < cap) out[n] = (int64_t)v; ++n; }
    else il.skip((uint32_t)(t & 7));
  }
  if (!il.ok) { g_io_err = "e2t_io: malformed Int64List"; return -2; }
  *n_out = n;
  return n <= cap ? 0 : -3;
}

// Strings -> indices with EOS append and OOV fallback (tfh.string_seq_to_index_seq; subjects.py:344-361).
// vocab: n_vocab NUL-terminated strings concatenated; a linear scan is fine for the few tokens per utterance
// only with a hash: built by the caller once -> here we take sorted (by bytes) vocab + permutation.
int e2t_tokens_to_indices(const uint8_t* payload, uint64_t payload_len, const char* const* sorted_vocab,
                          const int32_t* sorted_ids, int32_t n_vocab, int32_t oov_id, int32_t eos_id_or_neg,
                          int32_t* out, uint64_t cap, uint64_t* n_out) {
  if (!payload || !sorted_vocab || !sorted_ids || !out || !n_out) { g_io_err = "e2t_io: NULL argument"; return -1; }
  uint64_t off = 0, n = 0;
  const uint8_t* s;
  uint64_t sl;
  int rc;
  while ((rc = e2t_bytes_list_next(payload, payload_len, &off, &s, &sl)) == 1) {
    int lo = 0, hi = n_vocab - 1, id = oov_id;
    while (lo <= hi) {
      int mid = (lo + hi) / 2;
      const char* v = sorted_vocab[mid];
      size_t vl = strlen(v);
      int c = memcmp(s, v, sl < vl ? sl : vl);
      if (c == 0) c = (sl < vl) ? -1 : (sl > vl ? 1 : 0);
      if (c == 0) { id = sorted_ids[mid]; break; }
      if (c < 0) hi = mid - 1; else lo = mid + 1;
    }
    if (n < cap) out[n] = id;
    ++n;
  }
  if (rc < 0) return rc;
  if (eos_id_or_neg >= 0) { if (n < cap) out[n] = eos_id_or_neg; ++n; }
  *n_out = n;
  return n <= cap ? 0 : -3;
}

// ---- writer ---------------------------------------------------------------------------------------
int e2t_tfr_writer_open(const char* path, e2t_tfr_writer** out) {
  if (!path || !out) { g_io_err = "e2t_io: NULL argument"; return -1; }
  FILE* f = fopen(path, "wb");
  if (!f) { g_io_err = std::string("e2t_io: cannot create '") + path + "'"; return -1; }
  e2t_tfr_writer* w = new e2t_tfr_writer();
  w->f = f;
  *out = w;
  return 0;
}
int e2t_tfr_writer_write(e2t_tfr_writer* w, const void* data, uint64_t n) {
  if (!w || (!data && n)) { g_io_err = "e2t_io: NULL argument"; return -1; }
  uint8_t hdr[12];
  memcpy(hdr, &n, 8);
  uint32_t c = mask_crc(crc32c(hdr, 8));
  memcpy(hdr + 8, &c, 4);
  uint32_t cd = mask_crc(crc32c((const uint8_t*)data, (size_t)n));
  if (fwrite(hdr, 1, 12, w->f) != 12 || fwrite(data, 1, (size_t)n, w->f) != (size_t)n || fwrite(&cd, 1, 4, w->f) != 4) {
    g_io_err = "e2t_io: short write";
    return -2;
  }
  return 0;
}
int e2t_tfr_writer_close(e2t_tfr_writer* w) {
  if (!w) return 0;
  int rc = 0;
  if (w->f && fclose(w->f) != 0) { g_io_err = "e2t_io: close failed"; rc = -2; }
  delete w;
  return rc;
}

// ---- Example builder (tfh.make_feature_example) -------------------------------------------------------
int e2t_example_builder_new(e2t_example_builder** out) {
  if (!out) { g_io_err = "e2t_io: NULL argument"; return -1; }
  *out = new e2t_example_builder();
  return 0;
}
int e2t_example_builder_free(e2t_example_builder* b) { delete b; return 0; }
static void add_entry(e2t_example_builder* b, const char* key, uint32_t field, const std::string& list_msg) {
  std::string feature;
  put_len_field(feature, field, list_msg);
  std::string entry;
  put_len_field(entry, 1, std::string(key));
  put_len_field(entry, 2, feature);
  put_len_field(b->features, 1, entry);
}
int e2t_example_builder_add_floats(e2t_example_builder* b, const char* key, const float* v, uint64_t n) {
  if (!b || !key || (!v && n)) { g_io_err = "e2t_io: NULL argument"; return -1; }
  std::string lst;
  if (n) {
    put_varint(lst, (1 << 3) | 2);
    put_varint(lst, n * 4);
    lst.append(reinterpret_cast<const char*>(v), (size_t)n * 4);
  }
  add_entry(b, key, 2, lst);
  return 0;
}
// strs: n strings, lens[i] bytes each, concatenated in `blob`
int e2t_example_builder_add_bytes(e2t_example_builder* b, const char* key, const uint8_t* blob, const uint64_t* lens,
                                  uint64_t n) {
  if (!b || !key || (n && (!blob || !lens))) { g_io_err = "e2t_io: NULL argument"; return -1; }
  std::string lst;
  const uint8_t* p = blob;
  for (uint64_t i = 0; i < n; ++i) {
    put_varint(lst, (1 << 3) | 2);
    put_varint(lst, lens[i]);
    lst.append(reinterpret_cast<const char*>(p), (size_t)lens[i]);
    p += lens[i];
  }
  add_entry(b, key, 1, lst);
  return 0;
}
int e2t_example_builder_add_int64s(e2t_example_builder* b, const char* key, const int64_t* v, uint64_t n) {
  if (!b || !key || (!v && n)) { g_io_err = "e2t_io: NULL argument"; return -1; }
  std::string packed;
  for (uint64_t i = 0; i < n; ++i) put_varint(packed, (uint64_t)v[i]);
  std::string lst;
  if (n) put_len_field(lst, 1, packed);
  add_entry(b, key, 3, lst);
  return 0;
}
// serializes Example{features{...}} into *out (valid until the builder is reused / freed) and resets the builder
int e2t_example_builder_finish(e2t_example_builder* b, const uint8_t** out, uint64_t* len) {
  if (!b || !out || !len) { g_io_err = "e2t_io: NULL argument"; return -1; }
  std::string ex;
  put_len_field(ex, 1, b->features);
  b->features.swap(ex);   // keep the serialized Example alive inside the builder
  *out = reinterpret_cast<const uint8_t*>(b->features.data());
  *len = b->features.size();
  return 0;
}
int e2t_example_builder_reset(e2t_example_builder* b) {
  if (!b) return -1;
  b->features.clear();
  return 0;
}

// ---- padded batch assembly (the tf.data padded_batch of the reference's pipeline) -------------------------
// Copies utterance i ([len_i, C] fp32, row-major) into dst[i, :len_i, :] of a [B, T_pad, C] buffer and zero-fills
// the tail (padding_value 0.0, subjects.py:386-390).  Rows longer than T_pad are an error.
int e2t_pad_batch_f32(float* dst, int64_t B, int64_t T_pad, int64_t C, const float* const* src, const int64_t* lens) {
  if (!dst || !src || !lens) { g_io_err = "e2t_io: NULL argument"; return -1; }
  for (int64_t i = 0; i < B; ++i) {
    if (lens[i] < 0 || lens[i] > T_pad) { g_io_err = "e2t_io: sequence longer than the padded length"; return -3; }
    float* d = dst + i * T_pad * C;
    if (lens[i]) memcpy(d, src[i], (size_t)(lens[i] * C) * sizeof(float));
    memset(d + lens[i] * C, 0, (size_t)((T_pad - lens[i]) * C) * sizeof(float));
  }
  return 0;
}

// Same with the utterances split over n_threads worker threads (a 256-utterance batch of config 2 is 105 MB: one thread
// copies it in ~20 ms, five times the GPU step; the destination is typically a page-locked staging buffer).
int e2t_pad_batch_f32_mt(float* dst, int64_t B, int64_t T_pad, int64_t C, const float* const* src, const int64_t* lens,
                         int n_threads) {
  if (!dst || !src || !lens) { g_io_err = "e2t_io: NULL argument"; return -1; }
  for (int64_t i = 0; i < B; ++i)
    if (lens[i] < 0 || lens[i] > T_pad) { g_io_err = "e2t_io: sequence longer than the padded length"; return -3; }
  if (n_threads > B) n_threads = (int)B;
  if (n_threads <= 1) return e2t_pad_batch_f32(dst, B, T_pad, C, src, lens);
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      float* d = dst + i * T_pad * C;
      if (lens[i]) memcpy(d, src[i], (size_t)(lens[i] * C) * sizeof(float));
      memset(d + lens[i] * C, 0, (size_t)((T_pad - lens[i]) * C) * sizeof(float));
    }
  };
  std::vector<std::thread> pool;
  const int64_t per = (B + n_threads - 1) / n_threads;
  for (int t = 1; t < n_threads; ++t) {
    const int64_t lo = t * per, hi = lo + per < B ? lo + per : B;
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  work(0, per < B ? per : B);
  for (auto& th : pool) th.join();
  return 0;
}

}  // extern "C"
