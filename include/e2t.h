/* e2t.h -- C-ABI of the B200-native ECoG->text seq2seq engine (libe2t.so).
 *
 * This is the drop-in boundary for the hot path of jgmakin/ecog2txt: everything the reference
 * reaches through `machine_learning.neural_networks.sequence_networks.SequenceNetwork`
 * (imported at /root/reference/ecog2txt/trainers.py:33, constructed :126-135, driven
 * :318,355,367 (fit), :379 (restore_and_assess), :699,750 (get_weights_as_numpy_array),
 * :813-823 (_convolve_sequences/_encode_sequences), :933-937 (online predictor)).
 * The reference has no FFI of its own (pure Python on TF1.15); the Python class
 * ecog2txt_b200.SequenceNetwork binds these symbols through ctypes.
 *
 * Conventions
 *   - every call returns int: 0 = OK, <0 = error; text via e2t_last_error() (thread-local).
 *   - plain pointers and sizes only.  `loc` says where x / lens / y live: E2T_HOST (the library
 *     stages them to the device on its stream) or E2T_DEVICE (used in place).
 *   - the caller owns every buffer it passes; the library owns all device memory behind the
 *     opaque handle.  One handle <-> one GPU <-> one CUDA stream; not thread-safe per handle.
 *   - all floats are fp32, all indices int32, all arrays C-contiguous.
 *   - tensor names / shapes follow the TF checkpoint convention the reference parses in
 *     MultiSubjectTrainer.recover_model_sizes (/root/reference/ecog2txt/trainers.py:444-554).
 */
#ifndef E2T_H_
#define E2T_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define E2T_MAX_SUBNETS 16
#define E2T_MAX_LAYERS 8
#define E2T_HOST 0
#define E2T_DEVICE 1
#define E2T_STAGED0 2   /* inputs were copied ahead of time into staging slot 0 / 1 by e2t_stage_inputs */
#define E2T_STAGED1 3

/* which copy of a parameter tensor */
#define E2T_VALUE 0   /* trained variable                                      */
#define E2T_GRAD 1    /* gradient of the summed loss (before 1/ntok)           */
#define E2T_ADAM_M 2
#define E2T_ADAM_V 3
#define E2T_EMA 4     /* .../ExponentialMovingAverage shadow (trainers.py:466) */
/* e2t_flat_buffer only: the gradient buffer followed by 4 floats [unmasked-token count of the last training step, 0, 0, 0].
 * A data-parallel caller all-reduces THIS range once per step (gradients and the loss normaliser in one collective) and then
 * calls e2t_adam_ema_step_dev(h, subnet, NULL). */
#define E2T_GRAD_AND_COUNT 5

/* activations */
#define E2T_ACT_LINEAR 0
#define E2T_ACT_RELU 1

/* decoder attention (A7) */
#define E2T_ATTN_NONE 0
#define E2T_ATTN_LUONG 1    /* q = h Wq^T; alpha = softmax_s(q . enc_s); h~ = tanh(Wc [ctx; h] + bc); logits from h~ */
#define E2T_ATTN_BAHDANAU 2 /* additive score: alpha = softmax_s(v . tanh(Wq h + Wk enc_s)); context / combine as above */

/* encoder-targets head (A6) */
#define E2T_AUX_GAUSSIAN 0     /* float targets [B,T,F], 0.5 * squared error (subjects.py:369-380: MFCC-type streams) */
#define E2T_AUX_CATEGORICAL 1  /* int32 targets [B,T] (class index, 0 = pad), cross-entropy (phoneme streams) */

/* GEMM backends */
#define E2T_GEMM_AUTO 0     /* tcgen05 where shapes allow, SIMT otherwise */
#define E2T_GEMM_SIMT 1     /* fp32 CUDA-core tiles only (validation path) */
#define E2T_GEMM_TCGEN05 2  /* same as AUTO but reports an error if nothing could use tensor cores */

typedef struct e2t_handle e2t_handle;

/* Model geometry + optimiser constants; mirrors the manifest keys SequenceNetwork consumes
 * (layer_sizes, FF_dropout, RNN_dropout, EMA_decay: mochastar_word_sequence.yaml:3-4,11,62-75). */
typedef struct e2t_config {
  int32_t n_subnets;                     /* subjects; private conv weights (trainers.py:337-338) */
  int32_t subnet_id[E2T_MAX_SUBNETS];    /* ECoGSubject.subnet_id, subjects.py:106-108 */
  int32_t subnet_C[E2T_MAX_SUBNETS];     /* encoder_inputs num_features */
  int32_t subnet_W[E2T_MAX_SUBNETS];     /* decimation_factor = conv width = stride, subjects.py:144-157 */
  int32_t E;                             /* layer_sizes['encoder_embedding'][0] */
  int32_t n_enc_layers;
  int32_t H[E2T_MAX_LAYERS];             /* layer_sizes['encoder_rnn'] (per direction) */
  int32_t D;                             /* layer_sizes['decoder_embedding'][0] */
  int32_t Hd;                            /* layer_sizes['decoder_rnn'][0]; must equal 2*H[last] */
  int32_t V;                             /* decoder_targets num_features (vocabulary) */
  int32_t conv_act, emb_act;             /* E2T_ACT_* */
  int32_t pad_id, eos_id, start_id;      /* indices of <pad>, <EOS>; first decoder input */
  int32_t max_B, max_T, max_L;           /* workspace capacity: utterances/call, frames, target len */
  int32_t max_beam;                      /* capacity for beam search (>=1) */
  float ff_dropout, rnn_dropout;         /* training only */
  float lr, beta1, beta2, eps;           /* TF1 AdamOptimizer */
  float ema_decay;                       /* 0 disables the shadow copy */
  float penalty_scale;                   /* decoder_targets penalty_scale, subjects.py:289 */
  int32_t gemm_backend;                  /* E2T_GEMM_* */
  int32_t device;                        /* CUDA device ordinal */
  int32_t attention;                     /* E2T_ATTN_*: optional Luong / Bahdanau attention over the encoder outputs (not in the
                                          * reference model; default NONE = reference behaviour) */
  /* A6: auxiliary loss on the outputs of encoder layer `aux_layer` ('encoder_1_targets' -> 1: trainers.py:798-799,
   * mochastar_word_sequence.yaml:54,68-69,81).  aux_F == 0: no head (the minimal data_mapping has none, README.md:61). */
  int32_t aux_layer;
  int32_t aux_hidden;                    /* layer_sizes['encoder_1_projection'][0] (225); 0 = output layer only */
  int32_t aux_F;                         /* num_features of the encoder targets; 0 disables the head */
  int32_t aux_kind;                      /* E2T_AUX_* */
  float aux_penalty;                     /* encoder_1_targets_penalty_scale */
  /* layer_sizes['decoder_projection'] (mochastar_word_sequence.yaml:65; empty in every shipped manifest): width of ONE optional
   * hidden FF layer (relu + FF dropout) between the decoder output and the vocabulary projection; 0 = none. */
  int32_t proj_hidden;
} e2t_config;

const char* e2t_last_error(void);
/* ABI version of this header; bump on any signature change. */
int e2t_abi_version(void);

/* replaces SequenceNetwork.__init__ (trainers.py:126-135) */
int e2t_create(const e2t_config* cfg, e2t_handle** out);
int e2t_destroy(e2t_handle* h);
/* run on the caller's CUDA stream (cudaStream_t as void*; NULL = the library's own stream) */
int e2t_set_stream(e2t_handle* h, void* cuda_stream);
int e2t_sync(e2t_handle* h);

/* ---- parameters: replaces get_weights_as_numpy_array / the TF Saver (trainers.py:699-700,240-252) */
int e2t_param_count(e2t_handle* h);
int64_t e2t_param_total(e2t_handle* h); /* number of fp32 elements in the flat buffer */
/* name (NUL-terminated, truncated to name_cap), shape[<=4], ndim, offset into the flat buffer */
int e2t_param_info(e2t_handle* h, int index, char* name, int name_cap, int64_t* shape, int* ndim,
                   int64_t* offset);
int e2t_get_tensor(e2t_handle* h, const char* name, int which, float* host_out);
int e2t_set_tensor(e2t_handle* h, const char* name, int which, const float* host_in);
/* device pointer + length of one flat buffer (E2T_GRAD for the data-parallel all-reduce) */
int e2t_flat_buffer(e2t_handle* h, int which, void** dev_ptr, int64_t* n);
/* train_vars_scope (trainers.py:312,352,366): per-tensor trainable flag, default 1 */
int e2t_set_trainable(e2t_handle* h, const char* name, int trainable);
/* Adam step counter (for checkpoint/resume) */
int e2t_get_step(e2t_handle* h, int64_t* step);
int e2t_set_step(e2t_handle* h, int64_t step);

/* ---- training: replaces one sess.run(train_op) inside SequenceNetwork.fit (trainers.py:318) ----
 * x [B,T,C_subnet] fp32 zero-padded at the tail; lens [B] or NULL (inferred from the zero padding,
 * trainers.py:806-807); y [B,L] int32 targets with <EOS> appended and pad_id after it.
 * Runs forward + backward; gradients of  penalty_scale * sum_tokens CE  are left in the E2T_GRAD
 * buffer (zeroed first).  loss_sum / ntok (nullable, host) receive the summed loss and the number
 * of unmasked target tokens; reading them synchronises the stream. */
int e2t_train_step_grads(e2t_handle* h, int subnet, const float* x, const int32_t* lens,
                         const int32_t* y, int loc, int B, int T, int L, uint32_t dropout_seed,
                         float* loss_sum, int32_t* ntok);
/* Input pipeline (the tf.data prefetch of the reference, trainers.py:891-901): copy the NEXT minibatch host -> device on
 * the library's copy stream while the current step computes.  Two slots; a slot is overwritten only after the step that
 * consumed it has finished (event-ordered, no host synchronisation).  Consume with loc = E2T_STAGED0 + slot and the same
 * subnet / B / T / L (x, lens, y arguments are then ignored).  Host buffers should be page-locked for the copy to overlap. */
int e2t_stage_inputs(e2t_handle* h, int slot, int subnet, const float* x, const int32_t* lens,
                     const int32_t* y, int B, int T, int L);
/* Gradient buckets: lets a data-parallel caller overlap the all-reduce with the backward pass.  With
 * bucketing on, e2t_train_step_grads completes E2T_GRAD in flat ranges -- decoder-side tensors first, then the encoder layers
 * top to bottom, the subject-private conv last -- and records a CUDA event after each (deferred bias sums / un-permutes are
 * flushed per bucket: a few more small launches per step, so leave it off on one GPU).  After the call has returned (work
 * enqueued, nothing synchronised): bucket i = [offset, offset + n) of the flat buffer, final once e2t_grad_bucket_wait's
 * event has fired; buckets are disjoint, in completion order, and cover every tensor.  Off: one bucket = the whole buffer. */
int e2t_set_grad_buckets(e2t_handle* h, int on);
int e2t_grad_bucket_count(e2t_handle* h);
int e2t_grad_bucket_info(e2t_handle* h, int i, int64_t* offset, int64_t* n);
/* make `cuda_stream` (cudaStream_t as void*) wait on the device until bucket i of the most recent step is final */
int e2t_grad_bucket_wait(e2t_handle* h, int i, void* cuda_stream);
/* A6: encoder targets of the NEXT e2t_train_step_grads / e2t_eval_loss / e2t_input_saliency call (consumed by it):
 * float [B,T,aux_F] (E2T_AUX_GAUSSIAN) or int32 [B,T] (E2T_AUX_CATEGORICAL) at the input's frame rate, zero / pad-index
 * padded.  The library reverses them within the utterance length and keeps every W-th frame (trainers.py:791-795).
 * A step without targets skips the head (its parameters get zero gradients). */
int e2t_set_encoder_targets(e2t_handle* h, const void* targets, int loc, int B, int T);
/* the two terms of the most recent loss: penalty_scale * sum CE over `ntok` tokens, aux_penalty * sum over
 * `aux_frames` frames (0 when the head did not run); loss_sum of the calls above is their sum.  Synchronises. */
int e2t_last_losses(e2t_handle* h, float* decoder_sum, int32_t* ntok, float* aux_sum, int32_t* aux_frames);
/* The same four numbers without a synchronisation inside the step loop: e2t_post_losses enqueues their copy into
 * page-locked host memory (ring slot 0..3) behind the step just enqueued; e2t_fetch_losses waits for THAT copy only and
 * returns them -- typically one step later, while the device already works on the next step (the reference fetches the
 * loss with every session.run, trainers.py:806-823; a pipelined loop reads it with a lag of one step). */
int e2t_post_losses(e2t_handle* h, int slot);
int e2t_fetch_losses(e2t_handle* h, int slot, float* decoder_sum, int32_t* ntok, float* aux_sum, int32_t* aux_frames);
/* Running sums over the training steps since the last reset: out4 = [decoder loss, unmasked tokens, encoder-targets loss,
 * encoder-target frames].  Lets a training loop report the epoch loss with ONE host synchronisation per epoch instead of
 * one per step.  Synchronises. */
int e2t_read_loss_accumulators(e2t_handle* h, double* out4, int reset);
/* block the host until the most recent e2t_stage_inputs copy into `slot` has completed (its host buffer may be reused);
 * bounds how far a sync-free training loop runs ahead of the device */
int e2t_wait_staged(e2t_handle* h, int slot);
/* page-locked host memory for the caller's staging buffers: e2t_stage_inputs copies from it without an intermediate
 * pageable -> pinned bounce, i.e. asynchronously with respect to the host and to the compute stream */
int e2t_host_alloc(void** out, int64_t bytes);
int e2t_host_free(void* p);
/* Adam + EMA on the trainable tensors of `subnet` (private) and the shared ones, using
 * grad * grad_scale (1 / global token count).  subnet < 0: every subnet. */
int e2t_adam_ema_step(e2t_handle* h, int subnet, float grad_scale);
/* same with grad_scale = 1 / max(*token_count_dev, 1) read on the device (fp32 scalar in device memory, e.g. the all-reduced
 * token count): no host read-back between the backward pass and the optimiser.  token_count_dev == NULL: the count slot
 * behind the gradient buffer (E2T_GRAD_AND_COUNT). */
int e2t_adam_ema_step_dev(e2t_handle* h, int subnet, const float* token_count_dev);
/* forward only (assessment loss): same inputs, no dropout, weights = value or EMA */
int e2t_eval_loss(e2t_handle* h, int subnet, const float* x, const int32_t* lens, const int32_t* y,
                  int loc, int B, int T, int L, int use_ema, float* loss_sum, int32_t* ntok);

/* ---- saliency: replaces restore_and_get_saliencies (trainers.py:703-732) ----
 * d(loss)/d(encoder_inputs) of one batch, dropout off, weights = value or EMA.  decoder_penalty / aux_penalty
 * override the configured penalty scales for this call (get_saliencies sets every *_targets penalty to 0 except the
 * one under study).  dx [B,T,C] (nullable, host or device per `loc`): the input gradient (zero past each utterance's last conv window).
 * sq_norms [B,C] (nullable): sum over time of dx^2 per electrode (assessment_type 'norms'; the caller takes the
 * square root / averages).  Overwrites the E2T_GRAD buffer. */
int e2t_input_saliency(e2t_handle* h, int subnet, const float* x, const int32_t* lens, const int32_t* y, int loc,
                       int B, int T, int L, int use_ema, float decoder_penalty, float aux_penalty, float* dx,
                       float* sq_norms);

/* ---- decoding: replaces restore_and_assess's decode and the online predictor
 * ('decoder_outputs:0', 'decoder_probs:0'; trainers.py:379,933-937) ----
 * tokens [B,max_len] int32: argmax tokens up to and including <EOS>, then pad_id.
 * logp   [B,max_len] fp32 (nullable): log softmax(logits/temperature) of each emitted token. */
int e2t_greedy_decode(e2t_handle* h, int subnet, const float* x, const int32_t* lens, int loc, int B,
                      int T, int max_len, int use_ema, float temperature, int32_t* tokens,
                      float* logp);
/* tokens [B,beam,max_len] best-first (trainers.py:952-963); scores [B,beam] summed log-probs. */
int e2t_beam_decode(e2t_handle* h, int subnet, const float* x, const int32_t* lens, int loc, int B,
                    int T, int beam, int max_len, int use_ema, float temperature, int32_t* tokens,
                    float* scores);

/* ---- introspection (get_internal_activations, trainers.py:757-859) -------------------------
 * Copies an activation of the most recent forward pass to the host.  names: "lens", "lens2"
 * (int32 [B]), "conv_out" [T',B,E], "enc<l>_out" [T',B,2H], "final_h", "final_c" [B,Hd],
 * "logits" [L,B,V], "aux_out" [T',B,aux_F] (A6 head outputs; in training they hold d(loss)/d(out)).  n_cap = capacity of host_out in elements; *n_out = elements written. */
int e2t_get_activation(e2t_handle* h, const char* name, void* host_out, int64_t n_cap,
                       int64_t* n_out);

/* ---- accounting -------------------------------------------------------------------------- */
/* kernels launched by this handle since creation (all / those that used tcgen05) */
int e2t_launch_counts(e2t_handle* h, int64_t* total, int64_t* tensor_core);
/* named counters: "launches", "tcgen05_launches", "persistent_rnn_launches" (whole-layer recurrent kernels),
 * "decode_graph_replays" (greedy decodes of <= 8 host utterances replayed as one CUDA graph) */
int e2t_counter(e2t_handle* h, const char* name, int64_t* value);
/* Per-category device timing (CUDA events around every launch of the category; for bench.py's
 * roofline leg only -- the events perturb the step, so it is never on during a timed region).
 * categories: 0 recurrent steps (h Wh GEMM + gate kernel, fwd and bwd), 1 bulk GEMMs (input
 * projections, output projection, weight/input gradients), 2 temporal conv, 3 everything else. */
#define E2T_CAT_RECURRENT 0
#define E2T_CAT_BULK_GEMM 1
#define E2T_CAT_CONV 2
#define E2T_CAT_OTHER 3
#define E2T_CAT_REC_FWD 4   /* whole-layer persistent recurrent kernel, forward (both directions) */
#define E2T_CAT_REC_BWD 5   /* whole-layer persistent recurrent kernel, BPTT */
int e2t_profile_enable(e2t_handle* h, int on);
int e2t_profile_read(e2t_handle* h, int category, double* ms_total, int64_t* launches);
/* per-kernel breakdown of the same records: lines "label<TAB>launches<TAB>ms_total\n" (NUL-terminated, truncated to cap) */
int e2t_profile_report(e2t_handle* h, char* buf, int64_t cap);
/* self-test of the tcgen05 GEMM against the SIMT GEMM on random data; returns max |diff| */
int e2t_selftest_gemm(e2t_handle* h, int M, int N, int K, float* max_abs_diff);
/* diagnostic: average device time (ms) of one tcgen05 GEMM launch of this shape (tn: C = A^T B), zero operands */
int e2t_bench_gemm(e2t_handle* h, int M, int N, int K, int tn, float beta, int iters, float* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* E2T_H_ */
