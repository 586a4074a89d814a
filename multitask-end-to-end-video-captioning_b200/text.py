"""Host text / data glue with the reference's function names and observable behaviour (tf_s2vt.py:324-401,
cider_evaluation.py:122-172), written array-first: ids and masks come back as NumPy arrays ready for one pinned
host->device copy instead of nested Python lists."""
import gzip

import numpy as np

EOS, BOS, UNK = 0, 1, 2


def _open(path):
    return gzip.open(path, 'rt') if str(path).endswith('.gz') else open(path, 'r')


def read_vocabulary(vocabulary_file):
    """tf_s2vt.py:411-413."""
    with _open(vocabulary_file) as f:
        return [line.rstrip() for line in f]


def preProBuildWordVocab(vocabulary, word_count_threshold=0):
    """tf_s2vt.py:347-368: <eos>=0, <bos>=1, vocabulary line i -> id i+2 (so '<en_unk>' on line 0 is id 2)."""
    wordtoix = {'<eos>': EOS, '<bos>': BOS}
    wordtoix.update((w, i + 2) for i, w in enumerate(vocabulary))
    ixtoword = {EOS: '<eos>', BOS: '<bos>'}
    ixtoword.update((i + 2, w) for i, w in enumerate(vocabulary))
    return wordtoix, ixtoword


def read_sentences(sent_file):
    """The sentence half of get_video_feature_caption_pair (tf_s2vt.py:327-331): ndarray [n, 2] of (vid, sentence)."""
    rows = []
    with _open(sent_file) as f:
        for line in f:
            vid, sent = line.strip().split('\t')[:2]
            rows.append((vid, sent))
    return np.array(rows, dtype=object)


def read_features(feature_file, dtype=np.float32, n_threads=0):
    """The feature half (tf_s2vt.py:332-342): lines 'vid<id>_frame_<k>,f_1,...,f_D' grouped by the text before the first
    '_'; every video must have the same number of frames (the reference asserts it).  Returns {vid: float32 [T_v, D]},
    parsed by the native multi-threaded reader (ingest.FeatureFile; `ingest.FeatureFile(path)` itself gives lazy
    per-batch parsing into pinned memory)."""
    from . import _lib, ingest
    try:
        ff = ingest.FeatureFile(feature_file, n_threads)
    except _lib.S2vtError as e:
        if 'frame counts' in str(e):
            raise AssertionError(str(e))
        raise ValueError(str(e))
    try:
        feats = ff.to_dict()
    except _lib.S2vtError as e:
        raise ValueError(str(e))
    finally:
        ff.close()
    return feats if dtype == np.float32 else {v: a.astype(dtype) for v, a in feats.items()}


def get_video_feature_caption_pair(sent_file, feature_file):
    """tf_s2vt.py:324-344 -> (sents ndarray [n,2], features {vid: float32 [T_v, D]})."""
    return read_sentences(sent_file), read_features(feature_file)


def sentence_padding_toix(captions_batch, wordtoix, n_caption_lstm_step=35):
    """tf_s2vt.py:371-401.  ids int32 [B, T_c], mask float32 [B, T_c]: words (split on single spaces, lower-cased),
    then <eos> padding; mask is 1 through the first <eos>; captions with >= T_c words keep T_c-1 words + <eos> and an
    all-ones mask; out-of-vocabulary words become '<en_unk>'."""
    B, Tc = len(captions_batch), n_caption_lstm_step
    ids = np.zeros((B, Tc), dtype=np.int32)
    mask = np.zeros((B, Tc), dtype=np.float32)
    unk = wordtoix['<en_unk>']
    for b, cap in enumerate(captions_batch):
        words = cap.lower().split(' ')
        keep = len(words) if len(words) < Tc else Tc - 1
        ids[b, :keep] = [wordtoix.get(w, unk) for w in words[:keep]]
        mask[b, :keep + 1] = 1.0
    return ids, mask


def decode_captions(captions, idx_to_word):
    """cider_evaluation.py:122-143: join the words before the first <eos>."""
    caps = np.asarray(captions)
    caps = caps[None, :] if caps.ndim == 1 else caps
    out = []
    for row in caps:
        stop = np.flatnonzero(row == EOS)
        n = int(stop[0]) if stop.size else row.shape[0]
        out.append(' '.join(idx_to_word[int(w)] for w in row[:n]))
    return out


def decode_captions_masks(captions, idx_to_word):
    """cider_evaluation.py:145-172 -> (masks [N, T] of 0/1 lists, decoded strings); mask covers the first <eos> (R1)."""
    caps = np.asarray(captions)
    caps = caps[None, :] if caps.ndim == 1 else caps
    T = caps.shape[1]
    is_eos = caps == EOS
    first = np.where(is_eos.any(1), is_eos.argmax(1), T - 1)
    masks = (np.arange(T)[None, :] <= first[:, None]).astype(np.int64)
    return masks.tolist(), decode_captions(caps, idx_to_word)


def get_captions(captions, vid):
    """reinforcement_multisampling_tf_s2vt.py:600-601."""
    captions = np.asarray(captions, dtype=object)
    return captions[captions[:, 0] == vid, 1].tolist()


def group_by_video(sents):
    """{vid: [sentences]} and the video order of first appearance (what get_captions yields for every vid, in one pass)."""
    by, order = {}, []
    for vid, s in sents:
        if vid not in by:
            by[vid] = []
            order.append(vid)
        by[vid].append(s)
    return by, order


def get_multilabel(vid_sentence, vocabulary):
    """reinforce_multitask_e2e_attribute_loss.py:874-893: {vid: int64 [len(vocabulary)]} with label[k] = 1 iff attribute word k
    occurs (as a whole `str.split()` token) in any sentence of the video."""
    dup = {}                                                            # the reference marks EVERY position holding the word: duplicates share it
    for k, w in enumerate(vocabulary):
        dup.setdefault(w, []).append(k)
    out = {}
    for vid, sents in vid_sentence.items():
        lab = np.zeros(len(vocabulary), dtype=np.int64)
        for s in sents:
            for w in set(s.split()):
                if w in dup:
                    lab[dup[w]] = 1
        out[vid] = lab
    return out
