/*
 * s2vt_io.h -- host-side C ABI either side of the S2VT caption hot path (libs2vt_b200.so): the reference's on-disk
 * formats read at speed into the in-memory layouts the device entry points of s2vt.h take.  SURVEY.md section 8(f)
 * rows N3 (feature-file ingest) and N2 (TensorFlow checkpoint reader).  Host pointers only; no CUDA call is made by
 * anything in this header.  Return codes are the S2VT_E* values of s2vt.h; s2vt_io_last_error() gives the text of the
 * last failure on the calling thread.
 */
#ifndef S2VT_IO_H_
#define S2VT_IO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* s2vt_io_last_error(void);

/* ---- feature files -------------------------------------------------------------------------------------------------
 * Replaces the feature half of get_video_feature_caption_pair (tf_s2vt.py:332-342; writer tf_feature_extract.py:153-154):
 * a text file with one frame per line, "vid<id>_frame_<k>,f_1,...,f_D".  Frames are grouped by the text before the first
 * '_' of the first field, videos are numbered in order of first appearance, the frames of a video keep file order, and
 * every video must have the same number of frames (the reference asserts it: S2VT_EINVAL here).  Numbers are converted
 * like the reference's feed does (NumPy: decimal string -> correctly rounded double -> float32).
 */
typedef struct s2vt_feature_file s2vt_feature_file;

/* Map `path` and index its lines with `n_threads` workers (<= 0: all cores). */
int32_t s2vt_features_open(const char* path, int32_t n_threads, s2vt_feature_file** out);
/* Same over text already in memory (e.g. a decompressed .gz); `text` must stay valid until s2vt_features_close. */
int32_t s2vt_features_open_memory(const void* text, size_t len, int32_t n_threads, s2vt_feature_file** out);
void s2vt_features_close(s2vt_feature_file* f);
int64_t s2vt_features_num_videos(const s2vt_feature_file* f);
int32_t s2vt_features_num_frames(const s2vt_feature_file* f);  /* T_v */
int32_t s2vt_features_dim(const s2vt_feature_file* f);         /* D = fields after the first, of the first line */
/* NUL-terminated id of video i ("vid1234"); owned by the handle. */
const char* s2vt_features_video_id(const s2vt_feature_file* f, int64_t i);
/* Index of a video id, or -1 (a KeyError in the reference's `train_features[x]`, tf_s2vt.py:487). */
int64_t s2vt_features_find(const s2vt_feature_file* f, const char* video_id);
/* Parse the frames of videos video_index[0..n) into out[n][T_v][D] float32 (host memory, ideally pinned): the batch
 * `[train_features[x] for x in vid]` of tf_s2vt.py:487 in its device layout.  A line with a different field count or a
 * field that is not a number gives S2VT_ESHAPE / S2VT_EINVAL (ValueError in the reference's feed). */
int32_t s2vt_features_read(const s2vt_feature_file* f, const int64_t* video_index, int64_t n, float* out, int32_t n_threads);

/* ---- TensorFlow checkpoints ----------------------------------------------------------------------------------------
 * Replaces tf.train.Saver.restore / tf.train.NewCheckpointReader as used by optimistic_restore
 * (reinforcement_multisampling_tf_s2vt.py:47-61, 663, 883; tf_s2vt.py:440, 560) without TensorFlow: reads both
 * checkpoint formats a TF-1.x Saver writes,
 *   V2 (default since TF 0.12): "<prefix>.index" (an SSTable of BundleEntryProto) + "<prefix>.data-?????-of-?????"
 *   V1 (write_version=1):       one SSTable file "<prefix>" of SavedTensorSlices protos,
 * and exposes {variable name -> dtype, shape, data}.  Only what a Saver writes for dense variables is handled (full
 * tensors, no partitioned slices, block compression none); anything else is reported as S2VT_EINVAL, never guessed.
 */
typedef struct s2vt_ckpt s2vt_ckpt;

enum { S2VT_DT_FLOAT = 1, S2VT_DT_DOUBLE = 2, S2VT_DT_INT32 = 3, S2VT_DT_INT64 = 9 };  /* tensorflow.DataType values */

/* `prefix`: what the reference passes to saver.restore (e.g. "models/s2vt_model-10"). */
int32_t s2vt_ckpt_open(const char* prefix, s2vt_ckpt** out);
void s2vt_ckpt_close(s2vt_ckpt* c);
int32_t s2vt_ckpt_format(const s2vt_ckpt* c);                   /* 1 or 2 */
int32_t s2vt_ckpt_num_tensors(const s2vt_ckpt* c);
/* Name (owned by the handle), dtype, rank and dims (up to 8) of tensor i; tensors are in key order. */
int32_t s2vt_ckpt_tensor_info(const s2vt_ckpt* c, int32_t i, const char** name, int32_t* dtype, int32_t* ndim, int64_t* dims);
int32_t s2vt_ckpt_find(const s2vt_ckpt* c, const char* name);   /* index or -1 */
/* Copy tensor i as float32 into out[capacity] (DT_FLOAT verbatim; DT_DOUBLE / integer types converted); verifies the
 * masked CRC32C the bundle records (V2).  S2VT_ENOSPACE if capacity is smaller than the element count. */
int32_t s2vt_ckpt_read_f32(const s2vt_ckpt* c, int32_t i, float* out, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* S2VT_IO_H_ */
