/* e2t_io.h -- C-ABI of the native TFRecord / tf.train.Example reader-writer (libe2t_io.so, host only).
 *
 * The on-disk input contract of the hot path.  Each entry point replaces a
 * TensorFlow call the reference makes for this path:
 *   e2t_tfr_writer_* + e2t_example_builder_*  <- tf.io.TFRecordWriter + tfh.make_feature_example
 *                                                (/root/reference/ecog2txt/data_generators.py:317-326)
 *   e2t_tfr_reader_* + e2t_example_find       <- tf.data.TFRecordDataset + tf.io.parse_single_example with
 *                                                VarLenFeature(float32|string)
 *                                                (/root/reference/ecog2txt/subjects.py:297-302,616-618;
 *                                                 /root/reference/ecog2txt/trainers.py:891-901)
 *   e2t_tokens_to_indices                     <- tfh.string_seq_to_index_seq (EOS append, OOV fallback;
 *                                                /root/reference/ecog2txt/subjects.py:344-361)
 *   e2t_pad_batch_f32                         <- the zero-padded batch of encoder_inputs (subjects.py:386-390)
 * Every call returns int: 0 = OK (reader_next / bytes_list_next: 1 = item, 0 = end), <0 = error with text in
 * e2t_io_last_error() (thread-local).  The caller owns every buffer it passes in; pointers handed out stay valid
 * until the next call on the same reader / builder.
 */
#ifndef E2T_IO_H_
#define E2T_IO_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct e2t_tfr_reader e2t_tfr_reader;
typedef struct e2t_tfr_writer e2t_tfr_writer;
typedef struct e2t_example_builder e2t_example_builder;

const char* e2t_io_last_error(void);
int e2t_io_abi_version(void);
uint32_t e2t_io_masked_crc32c(const void* data, uint64_t n);

int e2t_tfr_reader_open(const char* path, int check_crc, e2t_tfr_reader** out);
int e2t_tfr_reader_next(e2t_tfr_reader* r, const uint8_t** data, uint64_t* len);
int e2t_tfr_reader_close(e2t_tfr_reader* r);

#define E2T_FEATURE_ABSENT 0
#define E2T_FEATURE_BYTES 1
#define E2T_FEATURE_FLOAT 2
#define E2T_FEATURE_INT64 3
int e2t_example_find(const uint8_t* rec, uint64_t len, const char* key, int* kind, const uint8_t** payload,
                     uint64_t* payload_len, uint64_t* count);
int e2t_bytes_list_next(const uint8_t* payload, uint64_t payload_len, uint64_t* offset, const uint8_t** str,
                        uint64_t* str_len);
int e2t_int64_list_copy(const uint8_t* payload, uint64_t payload_len, int64_t* out, uint64_t cap, uint64_t* n_out);
int e2t_tokens_to_indices(const uint8_t* payload, uint64_t payload_len, const char* const* sorted_vocab,
                          const int32_t* sorted_ids, int32_t n_vocab, int32_t oov_id, int32_t eos_id_or_neg,
                          int32_t* out, uint64_t cap, uint64_t* n_out);

int e2t_tfr_writer_open(const char* path, e2t_tfr_writer** out);
int e2t_tfr_writer_write(e2t_tfr_writer* w, const void* data, uint64_t n);
int e2t_tfr_writer_close(e2t_tfr_writer* w);

int e2t_example_builder_new(e2t_example_builder** out);
int e2t_example_builder_free(e2t_example_builder* b);
int e2t_example_builder_reset(e2t_example_builder* b);
int e2t_example_builder_add_floats(e2t_example_builder* b, const char* key, const float* v, uint64_t n);
int e2t_example_builder_add_bytes(e2t_example_builder* b, const char* key, const uint8_t* blob, const uint64_t* lens,
                                  uint64_t n);
int e2t_example_builder_add_int64s(e2t_example_builder* b, const char* key, const int64_t* v, uint64_t n);
int e2t_example_builder_finish(e2t_example_builder* b, const uint8_t** out, uint64_t* len);

int e2t_pad_batch_f32(float* dst, int64_t B, int64_t T_pad, int64_t C, const float* const* src, const int64_t* lens);
/* the same with the utterances split over n_threads worker threads (dst is typically a page-locked staging buffer) */
int e2t_pad_batch_f32_mt(float* dst, int64_t B, int64_t T_pad, int64_t C, const float* const* src, const int64_t* lens,
                         int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* E2T_IO_H_ */
