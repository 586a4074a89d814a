/*
 * s2vt.h -- C ABI of the B200-native S2VT caption hot path (libs2vt_b200.so).
 *
 * The reference (adwardlee/multitask-end-to-end-video-captioning, TensorFlow 1.1 scripts) has no FFI or plugin
 * interface: its seam is the feed/fetch contract of the graphs built by class Video_Caption_Generator and the
 * host helpers around them.  Every entry point below replaces one such contract; the citation names the
 * reference lines it stands in for (paths relative to the reference root).
 *
 * Conventions
 *   - plain C types, raw DEVICE pointers unless a parameter says "host"; the caller owns every buffer
 *     (persistent state block and scratch workspace are sized by the library and allocated by the caller --
 *     PyTorch in this repo); nothing is allocated after s2vt_create.
 *   - every call is asynchronous on the given cudaStream_t and never synchronises;
 *   - return 0 on success, a negative S2VT_E* code otherwise; s2vt_last_error() gives the text;
 *   - one handle per GPU / process; a handle is not thread-safe; distinct handles are independent;
 *   - rows are "sample-major" like the reference's np.vstack (reinforcement_multisampling_tf_s2vt.py:764-782):
 *     row n of an [N, ...] tensor belongs to video n % B.
 */
#ifndef S2VT_H_
#define S2VT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct s2vt_handle s2vt_handle;
typedef void* s2vt_stream; /* cudaStream_t */

enum {
    S2VT_OK = 0,
    S2VT_EINVAL = -1,     /* bad argument / shape */
    S2VT_ENOTFOUND = -2,  /* unknown variable name (optimistic_restore skips these silently) */
    S2VT_ESHAPE = -3,     /* variable known, shape differs (also skipped by optimistic_restore) */
    S2VT_ECUDA = -4,      /* CUDA runtime error, see s2vt_last_error */
    S2VT_ENOSPACE = -5,   /* bound state / workspace block too small */
    S2VT_ESTATE = -6      /* call order violated (e.g. backward without forward) */
};

enum { S2VT_PREC_BF16 = 0, S2VT_PREC_FP32 = 1 };
enum { S2VT_GEMM_AUTO = 0, S2VT_GEMM_MMA_SYNC = 1, S2VT_GEMM_TCGEN05 = 2,
       /* debugging / A-B variants of the tcgen05 path (tests/test_gpu_backends.py): */
       S2VT_GEMM_TC_N128 = 3,        /* 128-wide batched tiles */
       S2VT_GEMM_TC_MC2X2 = 4,       /* + 2x2 TMA multicast clusters */
       S2VT_GEMM_STEP_MC8 = 5, S2VT_GEMM_STEP_MC4 = 6,   /* multicast the activations of the >128-row step kernels */
       S2VT_GEMM_STEP_N64 = 7,       /* 64-wide tiles for the >128-row step kernels */
       S2VT_GEMM_WGRAD_TRANSPOSED = 8,   /* weight gradients through explicit transposes instead of MN-major descriptors */
       S2VT_GEMM_PER_STEP = 9,       /* one launch per time step instead of persistent chains */
       S2VT_GEMM_CHAIN_RING = 10,    /* persistent chains without the weights-stationary variant */
       S2VT_GEMM_CHAIN_MC4 = 11,     /* weights-stationary chains: activation multicast over clusters of 4 instead of 8 */
       S2VT_GEMM_CHAIN_NOMC = 12,    /* weights-stationary chains without activation multicast */
       S2VT_GEMM_CHAIN_PLAIN = 13,   /* > 128-row forward chains on the plain ring kernel instead of the pipelined weights-stationary one */
       S2VT_GEMM_SINGLE_CTA = 14,    /* large batched GEMMs on single-CTA 128 x 256 tiles instead of cta_group::2 pairs (256 x 256) */
       S2VT_GEMM_PAIR_ALL = 15,      /* every large batched GEMM on pairs (default: only where the isolated timings say the pair wins) */
       S2VT_GEMM_CHAIN_WS2_BWD = 16 }; /* > 128-row BPTT chain on the weights-stationary pipelined kernel (partial tiles through L2) instead of the ring chain: measured slower */

/* Model dimensions: the "Train Parameters" constants of reinforcement_multisampling_tf_s2vt.py:505-511
 * (dim_image, word_dim, lstm_dim, n_video_lstm_step, n_caption_lstm_step) and n_words = len(wordtoix) (:617). */
typedef struct s2vt_config {
    int32_t dim_image;       /* 1536 */
    int32_t word_dim;        /* 500  */
    int32_t lstm_dim;        /* 1000 */
    int32_t n_words;         /* 9972 */
    int32_t n_video_steps;   /* T_v  */
    int32_t n_caption_steps; /* 35   */
    int32_t n_attributes;    /* 0, or 400 for the attribute head (reinforce_multitask_e2e_attribute_loss.py:112-114) */
    int32_t precision;       /* S2VT_PREC_* : compute dtype of the GEMM operands (accumulation is always fp32) */
    int32_t gemm_backend;    /* S2VT_GEMM_* for the large batched GEMMs */
    float dropout_keep;      /* DropoutWrapper(output_keep_prob), 0.9 in the reference (:66,85-87); 1 disables */
} s2vt_config;

int s2vt_create(const s2vt_config* cfg, s2vt_handle** out);
void s2vt_destroy(s2vt_handle* h);
const char* s2vt_last_error(const s2vt_handle* h);

/* ---- memory: Video_Caption_Generator.__init__ (:64-98) creates the variables; here the caller provides the block.
 * state block  = fp32 parameters | fp32 gradients | Adam m | Adam v | compute-dtype weight copies | embedding table.
 * workspace    = scratch for the largest call with at most n_videos videos / n_rows caption rows / beam k. */
size_t s2vt_num_params(const s2vt_handle* h); /* floats in the flat fp32 parameter (= gradient = Adam slot) vector */
size_t s2vt_state_bytes(const s2vt_handle* h);
size_t s2vt_workspace_bytes(const s2vt_handle* h, int n_videos, int n_rows, int beam);
int s2vt_bind(s2vt_handle* h, void* state, size_t state_bytes, void* workspace, size_t workspace_bytes);
float* s2vt_params(const s2vt_handle* h);    /* device pointers into the state block (flat, TF layouts) */
float* s2vt_grads(const s2vt_handle* h);
float* s2vt_adam_m(const s2vt_handle* h);
float* s2vt_adam_v(const s2vt_handle* h);

/* Variable table keyed by the TF checkpoint names (:79-98; "s2vt/LSTM{1,2}/basic_lstm_cell/{weights,biases}" and
 * their TF>=1.2 / TF<=0.12 aliases).  offset is in floats inside the flat vector. */
int s2vt_num_variables(const s2vt_handle* h);
int s2vt_variable_info(const s2vt_handle* h, int index, const char** tf_name, int64_t* offset, int64_t shape[2], int* ndim);
/* optimistic_restore (:47-61): copy iff name and shape match.  src is a HOST pointer to fp32 data. */
int s2vt_load_param(s2vt_handle* h, const char* tf_name, const float* src_host, const int64_t* shape, int ndim, s2vt_stream st);
/* Re-derive the compute-dtype copies and the Wemb.W2[emb rows] gate table from the fp32 master parameters.
 * Must follow any direct write to s2vt_params(); s2vt_optimizer_step calls it itself. */
int s2vt_refresh(s2vt_handle* h, s2vt_stream st);

/* Opt-in: let the next s2vt_rl_backward / s2vt_xe_backward reuse the frame projection and LSTM1 forward that the
 * preceding s2vt_rollout computed for the SAME video buffer (LSTM1 never sees a word and its state is not affected by the
 * output dropout, so the two passes are identical).
 * HAZARD: the cache is keyed on the (pointer, B) pair only -- the library cannot see the buffer's contents.  A caller that
 * REWRITES the same device buffer between the rollout and the backward call (e.g. a staging buffer refilled in place) would train
 * on the stale LSTM1 pass.  With the option enabled the caller promises not to modify the buffer in between; any call that takes
 * a different pointer or B, s2vt_refresh / s2vt_optimizer_step, and s2vt_set_reuse_frontend itself drop the cache.
 * trainer.ReinforceTrainer.step is the one user: rollout and backward run back to back on one tensor. */
int s2vt_set_reuse_frontend(s2vt_handle* h, int enable);

/* ---- decoding -------------------------------------------------------------------------------------------------
 * build_sampler (:342-391): greedy argmax decode, no early stop.   video fp32 [B, T_v, dim_image] -> ids int32 [B, T_c] */
int s2vt_greedy(s2vt_handle* h, const float* video, int B, int32_t* ids_out, s2vt_stream st);
/* build_multinomial_sampler (:294-339) called K times (:743-753) + build_sampler, fused: the frames are encoded once.
 * sampled int32 [K*B, T_c] (row k*B+j = sample k of video j), greedy int32 [B, T_c] (nullable).
 * Categorical draws use Philox4x32-10(seed; row_base + row, step, vocab index) Gumbel-max (tf.multinomial [lib]). */
int s2vt_rollout(s2vt_handle* h, const float* video, int B, int K, uint64_t seed, uint32_t row_base, int32_t* sampled_out,
                 int32_t* greedy_out, s2vt_stream st);
/* decode_captions_masks (cider_evaluation.py:145-172), the mask half: 1 through the first <eos>. lengths nullable. */
int s2vt_caption_masks(s2vt_handle* h, const int32_t* ids, int N, float* mask_out, int32_t* lengths_out, s2vt_stream st);

/* ---- training ---------------------------------------------------------------------------------------------------
 * build_loss (:227-292) / build_model (:100-177) forward, teacher forced, with DropoutWrapper masks drawn from
 * Philox(drop_seed; row_base + n, step, unit) (drop_seed == 0 or dropout_keep == 1: no dropout).
 * video fp32 [B, T_v, D]; captions int32 [N, T_c]; row n uses video n % B (B == N: one video per row, the literal feed).
 * logp_out   fp32 [N, T_c]  log_softmax(logits)[n, t, captions[n,t]]      (un-masked; nullable)
 * logits_out fp32 [T_c, N, n_words] the `probs` list of build_model (:157)  (nullable, debugging / parity only) */
int s2vt_teacher_forward(s2vt_handle* h, const float* video, int B, const int32_t* captions, int N, uint64_t drop_seed,
                         uint32_t row_base, float* logp_out, float* logits_out, s2vt_stream st);
/* REINFORCE objective and its gradient (:643-650): sum_loss = -sum(logp*mask*(rewards-base_line)) / norm.
 * norm <= 0 means sum(mask) of this call; pass the global sum for data-parallel runs.  Overwrites s2vt_grads()
 * (scaled by grad_scale, 1 for the plain objective, (1-lambda) for the stage-3 mixes) and writes
 * loss_out[0] = grad_scale * sum_loss (device); with accumulate != 0 both gradients and loss_out[0] are added to. */
int s2vt_rl_backward(s2vt_handle* h, const float* video, int B, const int32_t* captions, const float* mask, const float* rewards,
                     const float* base_line, int N, float norm, float grad_scale, int accumulate, uint64_t drop_seed, uint32_t row_base,
                     float* loss_out, s2vt_stream st);
/* build_model XE objective (:153-166 = tf_s2vt.py:143-166): label-smoothed batch-mean CE, /sum(mask), + decay * L2.
 * loss_out[0] = total loss, loss_out[1] = weight-decay part. */
int s2vt_xe_backward(s2vt_handle* h, const float* video, int B, const int32_t* captions, const float* mask, int N, float label_smoothing,
                     float decay, float norm, float grad_scale, int accumulate, uint64_t drop_seed, uint32_t row_base, float* loss_out,
                     s2vt_stream st);
/* The same objective for a SHARD of the batch (data parallel): Q3 couples the rows of a batch -- the step loss is
 * mean_b(CE_b) * sum_b mask[b, i] -- so the shard needs the global per-step mask sums mask_colsum_global fp32 [T_c] (device), the
 * global row count and norm = the global sum(mask) (> 0, required).  Per-rank gradients and losses then ADD up to the single-process
 * ones (sum all-reduce, no averaging); pass decay on one rank only. */
int s2vt_xe_backward_sharded(s2vt_handle* h, const float* video, int B, const int32_t* captions, const float* mask, int N, float label_smoothing,
                             float decay, float norm, const float* mask_colsum_global, int n_rows_global, float grad_scale, int accumulate,
                             uint64_t drop_seed, uint32_t row_base, float* loss_out, s2vt_stream st);
/* Attribute head (reinforce_multitask_e2e_attribute_loss.py:375-380): sigmoid CE of mean_t(video) . attr_W + attr_b
 * against labels fp32 {0,1} [B, n_attributes], / (n_attributes * B).  Adds grad_scale * gradient to attr_W / attr_b. */
int s2vt_attribute_backward(s2vt_handle* h, const float* video, int B, const float* labels, float grad_scale, float* loss_out,
                            s2vt_stream st);
/* tf.clip_by_global_norm + AdamOptimizer.apply_gradients (:650-652; tf_s2vt.py:446-448), TF-1.1 Adam arithmetic.
 * step = 1-based Adam time step; lr already includes exponential_decay.
 * flags: S2VT_OPT_WEMB_SLICE_NORM uses the IndexedSlices norm of the Wemb gradient;
 *        S2VT_OPT_NORMALIZE divides the gradients (accumulated with norm = 1 and summed over ranks) by the global
 *        sum(mask) that s2vt_rl_backward left in the gradient block's aux slot -- the data-parallel form of `/ norm`.
 * out (device, nullable, 2 floats): out[0] = global gradient norm, out[1] = loss (aux slot, normalised likewise). */
enum { S2VT_OPT_WEMB_SLICE_NORM = 1, S2VT_OPT_NORMALIZE = 2 };
int s2vt_optimizer_step(s2vt_handle* h, float lr, float clip_norm, int64_t step, int flags, float* out, s2vt_stream st);

/* ---- instrumentation used by bench.py: kernels launched by this handle so far; CUDA-event brackets around every GEMM
 * launch of class 0 (batched GEMMs) and around every CHAIN of class 1 launches (recurrent-step GEMMs: bracketing each one
 * would defeat their programmatic dependent launch overlap), with algorithmic FLOPs and bytes.  profile_read synchronises
 * the device, fills ms[2], flops_out[4] = {flops class 0, flops class 1, bytes class 0, bytes class 1}, launches[2] and
 * clears the records. */
long long s2vt_launch_count(const s2vt_handle* h);
int s2vt_profile(s2vt_handle* h, int enable);
int s2vt_profile_read(s2vt_handle* h, double* ms_out, double* flops_out, long long* launches_out);
/* per (class, M, N, K) totals of the current records (call before s2vt_profile_read); returns the number of rows.
 * count = GEMMs executed (recurrent steps for class 1), launches = kernel launches that executed them (a persistent
 * chain runs many steps per launch), bytes = algorithmic bytes; bytes / launches may be NULL */
int s2vt_profile_shapes(s2vt_handle* h, int cap, int* cls, int* M, int* N, int* K, double* ms, long long* count, double* bytes,
                        long long* launches);
/* Data-parallel overlap hook.  After s2vt_rl_backward (plain REINFORCE objective: grad_scale 1, no accumulation) three contiguous
 * ranges of the flat gradient block are final long before the call's last kernel:
 *   segment 0: embed_word_W, embed_word_b (a third of the bytes) -- they only need d logits, the BPTT chains follow;
 *   segment 1: Wemb;   segment 2: the LSTM2 weights and biases -- final while the LSTM1 BPTT chain still runs on the side stream.
 * The call makes `stream` wait for exactly that point and returns the range (in floats), so the host can launch its all-reduce under
 * the remaining kernels and reduce the rest afterwards.  S2VT_ESTATE if the last backward call gives no such guarantee (XE / mixed
 * objectives add weight decay or a second pass) or the segment was already handed out. */
int s2vt_grad_segment_ready(s2vt_handle* h, int segment, s2vt_stream stream, int64_t* offset, int64_t* count);
/* ---- data-parallel gradient exchange over NVLink peer memory (replaces the NCCL all-reduce of trainer.allreduce_gradients; the reference
 * is single-GPU).  One process per GPU, all GPUs of one NVSwitch node.  Every rank exports its bound state block and a small flag
 * block as CUDA IPC handles (64 bytes each; state_offset = position of the state block inside its device allocation), the host exchanges them
 * by any means (trainer.py: torch.distributed all_gather_object), s2vt_peer_connect maps the peers.  s2vt_peer_allreduce then sums the flat
 * gradient block (+ aux slots) over the ranks with ONE kernel per rank: slice r is summed by rank r in rank order from all blocks
 * (reduce-scatter by 128-bit loads over NVLink) and written from the registers into every rank's block (all-gather by remote stores), one
 * flag barrier before and one behind -- identical bits on every rank, independent of timing.  Collective: every rank makes the call, in the same order.  Errors:
 * S2VT_ECUDA when the memory cannot be shared or mapped (the caller keeps NCCL), S2VT_ESTATE before connect.  Teardown: every rank calls
 * s2vt_peer_disconnect (or destroys its handle) BEFORE any rank frees its state block -- CUDA IPC forbids freeing memory a peer still maps. */
#define S2VT_PEER_HANDLE_BYTES 64
int s2vt_peer_export(s2vt_handle* h, unsigned char* state_handle, int64_t* state_offset, unsigned char* comm_handle);
int s2vt_peer_connect(s2vt_handle* h, int rank, int world, const unsigned char* state_handles, const int64_t* state_offsets, const unsigned char* comm_handles);
int s2vt_peer_allreduce(s2vt_handle* h, s2vt_stream st);
/* The exchange fused with s2vt_optimizer_step (same arguments, same arithmetic per element): reduce-scatter, global norm from per-rank partial
 * sums (added in rank order: the same value everywhere), clip + TF Adam on this rank's slice only, the updated PARAMETERS pushed to every rank.  The
 * Adam slots are then current in the own slice only: s2vt_optimizer_step refuses (S2VT_ESTATE) and a checkpoint writer calls
 * s2vt_peer_gather_state (collective) first, which makes them whole on every rank again. */
int s2vt_peer_optimizer_step(s2vt_handle* h, float lr, float clip_norm, int64_t step, int flags, float* out, s2vt_stream st);
int s2vt_peer_gather_state(s2vt_handle* h, s2vt_stream st);
int s2vt_peer_disconnect(s2vt_handle* h);
/* tuning: which independent pieces use the library's internal side stream (bit 0: late half of s2vt_refresh, bit 1: the
 * vocabulary weight gradient, bit 2: the LSTM1 backward chain, bit 7 (128): the part of dout1 and of the two large LSTM2 weight gradients that
 * belongs to the time steps the LSTM2 BPTT chain finishes first runs beside the rest of that chain, behind a watcher of its grid-barrier counter);
 * default 135.  Bit 3 (debug) makes s2vt_beam_search use its un-fused
 * step -- materialised logits and separate top-k / bookkeeping / state-gather launches -- the checker of the fused one.  Bit 6 (64) moves the
 * vocabulary weight gradient from beside the LSTM2 BPTT chain to after it (measured slower: 8.80 vs 8.63 ms per iteration). */
int s2vt_set_overlap(s2vt_handle* h, int mask);
/* debug / measurement: one GEMM of the given (padded) shape on scratch operands inside the bound workspace, through the engine's own
 * dispatch (tile shape, CTA pairs, gemm_backend) -- scripts/gemm_shapes.py times every shape of an iteration in isolation with it.
 * mn_major != 0: C[M,N] = X^T . Y over K rows (weight-gradient form); fp32_out: fp32 or fp16 result store. */
int s2vt_debug_gemm(s2vt_handle* h, int M, int N, int K, int mn_major, int fp32_out, s2vt_stream st);
/* debug: per-launch phase timestamps (%globaltimer) of CTA (0,0) of every tcgen05 GEMM; NULL disables */
int s2vt_debug_probe(void* device_buffer);

/* ---- beam search: beam_probability + the host loop of final_beam_search.py:202-294 / e2e_beam_search.py:235-344,
 * batched over B videos on the device, semantics B1-B7 of SURVEY.md.
 * sentences int32 [B, T_c] (0-padded after the final token), lengths int32 [B], logprob / score fp32 [B]. */
int s2vt_beam_search(s2vt_handle* h, const float* video, int B, int beam_size, float length_normalization_factor, int32_t* sentences_out,
                     int32_t* lengths_out, float* logprob_out, float* score_out, s2vt_stream st);
/* The single-hypothesis contracts of the reference, for drop-in callers of beam_probability:
 * beam_init: encode one video -> state1, state2 ([1, 2*lstm_dim] = concat(c, h), state_is_tuple=False)
 * beam_step: (state2, state1, word) -> top-k ids, probs, state2', state1'. */
int s2vt_beam_init(s2vt_handle* h, const float* video, float* state1_out, float* state2_out, s2vt_stream st);
int s2vt_beam_step(s2vt_handle* h, const float* state2, const float* state1, const int32_t* word, int beam_size, int32_t* idx_out,
                   float* prob_out, float* state2_out, float* state1_out, s2vt_stream st);

/* ---- CIDEr-D reward: CiderD_scorer.compute_score behind evaluate_captions_cider (cider_evaluation.py:33-39, 60-87).
 * Tokens are int32 ids (ids >= n_words may be used for out-of-vocabulary reference words); n-grams are exact 64-bit
 * keys of <= 4 ids of 16 bits.  Build (host, once per corpus): */
typedef struct ciderd_corpus ciderd_corpus;
/* ref_tokens: concatenated reference sentences; ref_offsets[n_refs+1]; video_ref_offsets[n_videos+1] indexes refs.
 * Document frequencies are counted over the videos with video_in_df[v] != 0 (all when NULL) -- the CiderD df corpus
 * (`CiderD(df='msvd')`, cider_evaluation.py:12); ref_len = ln(#df videos).  Videos outside the df corpus can still be
 * scored against (e.g. test-set references). */
int ciderd_corpus_create(const int32_t* ref_tokens, const int64_t* ref_offsets, int64_t n_refs, const int64_t* video_ref_offsets,
                         int64_t n_videos, const uint8_t* video_in_df, ciderd_corpus** out);
void ciderd_corpus_destroy(ciderd_corpus* c);
size_t ciderd_corpus_device_bytes(const ciderd_corpus* c);
/* Serialise the tables into a caller-allocated HOST buffer of ciderd_corpus_device_bytes(); the caller copies it to
 * the device and passes that device pointer to ciderd_score. */
int ciderd_corpus_serialize(const ciderd_corpus* c, void* host_buffer);
/* Score N hypotheses: hyp int32 [N, T_c] (tokens before the first 0 count, R2), video_of_row int32 [N] indexes the
 * corpus videos.  scores_out float64 [N].  counts_out (nullable) uint64 [N, 140, 2] (key, count) n-gram tables. */
int ciderd_score(const void* corpus_device, const int32_t* hyp, const int32_t* video_of_row, int N, int Tc, double* scores_out,
                 unsigned long long* counts_out, s2vt_stream st);

/* ---- BLEU-4 / ROUGE-L rewards: the evaluate_captions_cider (sic) of bleu_evaluation.py:60-87 (Bleu(4).compute_score,
 * reward = per-sentence scores[3]) and rouge_evaluation.py:60-87 (Rouge().compute_score per-sentence scores), used by
 * bleu4_/rouge_reinforcement_multisampling_tf_s2vt.py:792-808.  pycocoevalcap arithmetic: BleuScorer(n=4) with
 * option='closest', tiny 1e-15 / small 1e-9 smoothing and brevity penalty; Rouge with beta = 1.2 over LCS.
 * Same token / key conventions as the CIDEr-D corpus; one corpus serves both scores. */
typedef struct s2vt_reward_corpus s2vt_reward_corpus;
int s2vt_reward_corpus_create(const int32_t* ref_tokens, const int64_t* ref_offsets, int64_t n_refs, const int64_t* video_ref_offsets,
                              int64_t n_videos, s2vt_reward_corpus** out);
void s2vt_reward_corpus_destroy(s2vt_reward_corpus* c);
size_t s2vt_reward_corpus_device_bytes(const s2vt_reward_corpus* c);
int s2vt_reward_corpus_serialize(const s2vt_reward_corpus* c, void* host_buffer);
/* hyp int32 [N, T_c] (words before the first 0), video_of_row int32 [N]; bleu_out float64 [N, 4] = BLEU_1..BLEU_4.
 * comps_out (nullable) int32 [N, 10] = cook_test's {correct[4], guess[4], testlen, closest reflen}: their sums over a test set
 * give the corpus-level BLEU of score_all (bleu_evaluation.py:14-30). */
int s2vt_bleu_score(const void* corpus_device, const int32_t* hyp, const int32_t* video_of_row, int N, int Tc, double* bleu_out, int32_t* comps_out,
                    s2vt_stream st);
/* rouge_out float64 [N].  empty_token: the id the caller gave the empty word ''.split(" ") yields (an empty
 * hypothesis is ONE empty token in the reference's Rouge); any id no reference uses if references hold no empty words. */
int s2vt_rouge_score(const void* corpus_device, const int32_t* hyp, const int32_t* video_of_row, int N, int Tc, int32_t empty_token,
                     double* rouge_out, s2vt_stream st);

/* ---- temporal-attention decoder (BASELINE config 3): class Video_Caption_Generator of
 * original_attention.py:54-251 -- frame projection, additive attention over the n frame embeddings, one LSTM3, tanh MLP head.
 * A separate handle with the same conventions as s2vt_handle (caller-owned state / workspace, explicit stream, codes).
 * Variables by TF name: Wemb [V,H], encode_image_W [D,H], encode_image_b, embed_att_w [H,1], embed_att_Wa, embed_att_Ua,
 * embed_att_ba, embed_word_W [H,V], embed_word_b, embed_nn_Wp [3H,H], embed_nn_bp, s2vt/LSTM3/basic_lstm_cell/{weights
 * [3H,4H], biases}.  One row per video (the reference has no multi-sample path for this model). */
typedef struct s2vt_att_handle s2vt_att_handle;
typedef struct s2vt_att_config {
    int32_t dim_image;        /* 1536 (:296) */
    int32_t dim_hidden;       /* 1000 (:297): LSTM3 width = word embedding width = attention width */
    int32_t n_words;
    int32_t n_video_steps;    /* n frames attended over (5 / 32), <= 128 */
    int32_t n_caption_steps;  /* 35 */
    int32_t precision;        /* S2VT_PREC_BF16 | S2VT_PREC_FP32 */
    float dropout_keep;       /* DropoutWrapper(output_keep_prob) of build_model: 0.9 (:417) */
    float hinge_beta;         /* beta = 10 (:300) */
    float hinge_m;            /* m = 0.5 (:299) */
    int32_t reg_frames;       /* alphas_1 = temp_alphas[:, 0:8] (:123) */
} s2vt_att_config;
int s2vt_att_create(const s2vt_att_config* cfg, s2vt_att_handle** out);
void s2vt_att_destroy(s2vt_att_handle* h);
const char* s2vt_att_last_error(const s2vt_att_handle* h);
size_t s2vt_att_num_params(const s2vt_att_handle* h);
size_t s2vt_att_state_bytes(const s2vt_att_handle* h);
/* train != 0: room for the per-step operands the backward pass needs */
size_t s2vt_att_workspace_bytes(const s2vt_att_handle* h, int n_videos, int train);
int s2vt_att_bind(s2vt_att_handle* h, void* state, size_t state_bytes, void* workspace, size_t workspace_bytes);
float* s2vt_att_params(const s2vt_att_handle* h);
float* s2vt_att_grads(const s2vt_att_handle* h);   /* num_params + 8 floats: gradients | [slice sq-norm, loss, regulariser, sum(mask)] */
float* s2vt_att_adam_m(const s2vt_att_handle* h);
float* s2vt_att_adam_v(const s2vt_att_handle* h);
int s2vt_att_num_variables(const s2vt_att_handle* h);
int s2vt_att_variable_info(const s2vt_att_handle* h, int index, const char** tf_name, int64_t* offset, int64_t shape[2], int* ndim);
int s2vt_att_load_param(s2vt_att_handle* h, const char* tf_name, const float* src_host, const int64_t* shape, int ndim, s2vt_stream st);
int s2vt_att_refresh(s2vt_att_handle* h, s2vt_stream st);
/* build_generator (:155-199) / build_sampler (:201-251): video [B, n, D] fp32 -> ids int32 [B, T_c] (arg-max words, no early
 * stop) and, if alphas_out != NULL, saved_alphas float32 [T_c, n, B]. */
int s2vt_att_greedy(s2vt_att_handle* h, const float* video, int B, int32_t* ids_out, float* alphas_out, s2vt_stream st);
/* build_model (:88-152) forward: captions int32 [B, T_c], mask fp32 [B, T_c] -> loss_out[0] = loss, loss_out[1] = its
 * regulariser part; logits_out (nullable) fp32 [T_c, B, V].  Dropout stream as in s2vt.h (drop_seed 0 = keep everything). */
int s2vt_att_xe_loss(s2vt_att_handle* h, const float* video, int B, const int32_t* captions, const float* mask, uint64_t drop_seed, uint32_t row_base,
                     float* loss_out, float* logits_out, s2vt_stream st);
/* build_model loss AND its gradients w.r.t. the 13 variables (what optimizer.compute_gradients(tf_loss) returns, :432), into the
 * flat gradient block in TF layouts. */
int s2vt_att_xe_backward(s2vt_att_handle* h, const float* video, int B, const int32_t* captions, const float* mask, uint64_t drop_seed, uint32_t row_base,
                         float* loss_out, s2vt_stream st);
/* tf.clip_by_global_norm(gradients, clip_norm) + AdamOptimizer.apply_gradients (:431-435; clip 10, lr 1e-4 halved every 10000
 * steps by the caller); the Wemb term of the global norm is the IndexedSlices norm as in TF.  out[0] = global norm, out[1] = loss.
 * Re-packs the operand copies (no separate refresh needed). */
int s2vt_att_optimizer_step(s2vt_att_handle* h, float lr, float clip_norm, int64_t step, float* out, s2vt_stream st);
long long s2vt_att_launch_count(const s2vt_att_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* S2VT_H_ */
