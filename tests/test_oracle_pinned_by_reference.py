"""The oracle (and the product's host mirror) against outputs of EXECUTED reference code.

tests/golden/reference_exec_golden.json.gz was written by scripts/make_reference_fixtures.py, which loads the pure-Python pieces of
the hot path from /root/reference at run time and executes them under Python 3 (beam_search.py Caption / TopN, the beam host loops of
final_beam_search.py:248-294 and e2e_beam_search.py:301-344, cider_evaluation.py decode_captions[_masks], tf_s2vt.py
preProBuildWordVocab / sentence_padding_toix, get_captions, get_multilabel).  These rows of SURVEY 8(a) -- a11, a14, a16, a17 and the
label half of a18 -- are therefore pinned by reference outputs, not by a restatement."""
import gzip
import json
import math
import os

import numpy as np
import pytest

import s2vt_b200
import synthetic_beam_step as S
from oracle import beam as obeam
from oracle import text as otext

G = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def gold():
    with gzip.open(os.path.join(G, 'reference_exec_golden.json.gz'), 'rt') as f:
        return json.load(f)


@pytest.fixture(scope='module')
def vocab():
    v = otext.read_vocabulary(os.path.join(G, 'msvd_vocabulary1.txt.gz'))
    return v, otext.build_word_vocab(v)


def test_caption_and_topn_match_reference(gold):
    c = obeam.Caption([1], None, None, -1.0, -2.0)
    d = obeam.Caption([2], None, None, -5.0, -2.0)
    assert {'lt': c < d, 'eq': c == d, 'lt_lower': obeam.Caption([3], None, None, 0, -3.0) < c} == gold['caption_cmp']
    for trace in gold['topn']:
        for case in trace:
            t = obeam.TopN(case['n'])
            for i, s in enumerate(case['scores']):
                t.push(obeam.Caption([i], None, None, s, s))
            assert t.size() == case['size']
            got = t.extract(sort=True)
            t.reset()
            assert [x.score for x in got] == case['sorted_scores']
            # which of several equal-score items survive depends on the heap's insertion order: ids must match too
            assert sorted(((x.score, x.sentence[0]) for x in got), reverse=True) == [tuple(x) for x in case['sorted_ids_by_score']]
            assert t.size() == case['size_after_reset']


def test_beam_host_loop_matches_reference(gold):
    """B1-B7 (exclude_num shrinking the expansion, finals leaving the beam, length normalisation, running out of steps)."""
    assert len(gold['beam_loop']) == len(S.CASES)
    finished = 0
    for case in gold['beam_loop']:
        calls = [0]
        step = S.make_step(case['seed'], case['beam_size'], case['eos_ramp'])

        def counted(s1, s2, w, step=step):
            calls[0] += 1
            return step(s1, s2, w)

        s1, s2 = S.initial_states()
        sent, lp, sc = obeam.beam_search(counted, s1, s2, case['beam_size'], case['Tc'], case['lnf'])
        assert [int(w) for w in sent] == case['sentence'], case
        assert lp == case['logprob'] and sc == case['score'], case       # same float64 operations in the same order
        assert calls[0] == case['step_calls']
        finished += case['sentence'][-1] == 0
    assert 0 < finished < len(S.CASES)                                   # both endings are exercised


def test_vocab_padding_and_decoding_match_reference(gold, vocab):
    v, (w2i, i2w) = vocab
    pw2i, pi2w = s2vt_b200.text.preProBuildWordVocab(v)
    for m in ((w2i, i2w), (pw2i, pi2w)):
        assert len(m[0]) == gold['vocab']['n_words']
        assert all(m[0][w] == i for w, i in gold['vocab']['probe'].items())
        assert all(m[1][int(i)] == w for i, w in gold['vocab']['ixtoword_probe'].items())
    sp = gold['sentence_padding_toix']
    want_ids, want_mask = np.array(sp['ids']), np.array(sp['mask'])
    ids, mask = otext.sentence_padding_toix(list(sp['captions']), w2i, 35)
    np.testing.assert_array_equal(np.array(ids), want_ids)
    np.testing.assert_array_equal(np.asarray(mask).astype(int), want_mask)
    ids, mask = s2vt_b200.text.sentence_padding_toix(list(sp['captions']), w2i, 35)
    np.testing.assert_array_equal(ids, want_ids)
    np.testing.assert_array_equal(mask.astype(int), want_mask)
    d = gold['decode']
    caps = np.array(d['captions'])
    for T in (otext, s2vt_b200.text):
        masks, dec = T.decode_captions_masks(caps, i2w)
        assert masks == d['masks'] and dec == d['decoded']
        assert T.decode_captions(caps, i2w) == d['decoded_plain']
        assert T.decode_captions(caps[3], i2w) == d['decoded_1d']
        assert T.decode_captions_masks(caps[5], i2w)[0] == d['masks_1d']


def test_get_captions_and_multilabel_match_reference(gold):
    sents = s2vt_b200.text.read_sentences(os.path.join(G, 'msvd_sents_train_noval_lc_nopunc.txt.gz'))
    for vid, want in gold['get_captions'].items():
        assert otext.get_captions([tuple(x) for x in sents], vid) == want
        assert s2vt_b200.text.get_captions(sents, vid) == want
    attr = otext.read_vocabulary(os.path.join(G, 'train_most_freq_vocab_400_truncated.txt.gz'))
    assert len(attr) == gold['get_multilabel']['n_attributes']
    by, _ = s2vt_b200.text.group_by_video(sents)
    lab = s2vt_b200.text.get_multilabel({v: by[v] for v in gold['get_multilabel']['videos']}, attr)
    for v, want in gold['get_multilabel']['labels'].items():
        assert lab[v].tolist() == want and 0 < sum(want) < len(want)


@pytest.mark.gpu
def test_device_masks_match_reference(gold):
    """caption_mask_kernel (s2vt_caption_masks) against the masks the reference's decode_captions_masks produced."""
    import torch
    d = gold['decode']
    caps = np.array(d['captions'], dtype=np.int32)
    m = s2vt_b200.Video_Caption_Generator(dim_image=32, n_words=64, word_dim=16, lstm_dim=24, batch_size=4, n_video_lstm_step=2,
                                          n_caption_lstm_step=caps.shape[1], precision='fp32', max_videos=4, max_rows=caps.shape[0])
    mask, lens = m.caption_masks(torch.from_numpy(caps).cuda())
    np.testing.assert_array_equal(mask.cpu().numpy().astype(int), np.array(d['masks']))
    assert math.isclose(float(mask.sum().item()), float(np.array(d['masks']).sum()))
