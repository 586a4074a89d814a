"""Oracle restatement of the reference's pure-Python text glue.  TEST INFRASTRUCTURE.

Each function follows the reference statement by statement (py2 -> py3 only);
these are exact, not approximations.
"""
import numpy as np


def build_word_vocab(vocabulary):
    """tf_s2vt.py:347-368 (preProBuildWordVocab): <eos>=0, <bos>=1, vocab line i -> id i+2."""
    ixtoword = {1: '<bos>', 0: '<eos>'}
    wordtoix = {'<bos>': 1, '<eos>': 0}
    for idx, w in enumerate(vocabulary):
        wordtoix[w] = idx + 2
        ixtoword[idx + 2] = w
    return wordtoix, ixtoword


def read_vocabulary(path):
    """tf_s2vt.py:411-413: one word per line, rstrip()."""
    import gzip
    op = gzip.open if str(path).endswith('.gz') else open
    with op(path, 'rt') as f:
        return [line.rstrip() for line in f]


def read_sentences(path):
    """tf_s2vt.py:327-331: 'vid<id>\\t<sentence>' per line -> list of (vid, sentence)."""
    import gzip
    op = gzip.open if str(path).endswith('.gz') else open
    sents = []
    with op(path, 'rt') as f:
        for line in f:
            line = line.strip()
            id_sent = line.split('\t')
            sents.append((id_sent[0], id_sent[1]))
    return sents


def sentence_padding_toix(captions_batch, wordtoix, n_caption_lstm_step=35):
    """tf_s2vt.py:371-401.  Returns (ids list-of-lists, mask float [B, n_caption_lstm_step]).

    Quirks kept: split(' ') (not split()), pad with ' <eos>' up to n steps, the mask is 1
    through the first <eos> and 0 after, long captions are cut to n-1 words + <eos>,
    OOV words map to '<en_unk>'.
    """
    captions_batch = list(captions_batch)
    captions_mask = []
    for idx, each_cap in enumerate(captions_batch):
        one_caption_mask = np.ones(n_caption_lstm_step)
        word = each_cap.lower().split(' ')
        if len(word) < n_caption_lstm_step:
            for i in range(len(word), n_caption_lstm_step):
                captions_batch[idx] = captions_batch[idx] + ' <eos>'
                if i != len(word):
                    one_caption_mask[i] = 0
        else:
            new_word = ''
            for i in range(n_caption_lstm_step - 1):
                new_word = new_word + word[i] + ' '
            captions_batch[idx] = new_word + '<eos>'
        captions_mask.append(one_caption_mask)
    captions_mask = np.reshape(captions_mask, (-1, n_caption_lstm_step))
    caption_batch_ind = []
    for cap in captions_batch:
        current_word_ind = []
        for word in cap.lower().split(' '):
            if word in wordtoix:
                current_word_ind.append(wordtoix[word])
            else:
                current_word_ind.append(wordtoix['<en_unk>'])
        caption_batch_ind.append(current_word_ind)
    return caption_batch_ind, captions_mask


def decode_captions(captions, idx_to_word):
    """cider_evaluation.py:122-143: ids -> string of the words before the first <eos>."""
    captions = np.asarray(captions)
    if captions.ndim == 1:
        T = captions.shape[0]
        N = 1
    else:
        N, T = captions.shape
    decoded = []
    for i in range(N):
        words = []
        for t in range(T):
            word = idx_to_word[int(captions[t])] if captions.ndim == 1 else idx_to_word[int(captions[i, t])]
            if word == '<eos>':
                break
            words.append(word)
        decoded.append(' '.join(words))
    return decoded


def decode_captions_masks(captions, idx_to_word):
    """cider_evaluation.py:145-172: as decode_captions plus mask = 1 through the first <eos> (R1)."""
    captions = np.asarray(captions)
    if captions.ndim == 1:
        T = captions.shape[0]
        N = 1
    else:
        N, T = captions.shape
    decoded = []
    masks = []
    for i in range(N):
        words = []
        mask = []
        for t in range(T):
            word = idx_to_word[int(captions[t])] if captions.ndim == 1 else idx_to_word[int(captions[i, t])]
            if word == '<eos>':
                mask.append(1)
                break
            words.append(word)
            mask.append(1)
        decoded.append(' '.join(words))
        mask.extend([0] * (T - len(mask)))
        masks.append(mask)
    return masks, decoded


def get_captions(captions, vid):
    """reinforcement_multisampling_tf_s2vt.py:600-601: all sentences of one video id (linear scan)."""
    return [y for x, y in captions if x == vid]


def read_features(path):
    """tf_s2vt.py:332-342 + the float32 feed of :487-497: group the lines 'vid<id>_frame_<k>,f_1,...,f_D' by the text
    before the first '_', keep the STRING fields like the reference does, and convert them the way the feed does
    (np.asarray(list of str lists, dtype=float32)).  -> ({vid: float32 [T_v, D]}, video order of first appearance)."""
    import gzip
    import numpy as np
    op = gzip.open if str(path).endswith('.gz') else open
    features = {}
    with op(path, 'rt', newline='\n') as f:
        for line in f:
            splits = line.split(',')
            video_id = splits[0].split('_')[0]
            if video_id not in features:
                features[video_id] = []
            features[video_id].append(splits[1:])
    feature_length = [len(v) for v in features.values()]
    assert len(set(feature_length)) == 1
    return {v: np.asarray(rows, dtype=np.float32) for v, rows in features.items()}, list(features)
