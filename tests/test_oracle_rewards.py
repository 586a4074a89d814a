"""Oracle checks (CPU) of the BLEU-4 / ROUGE-L reward restatements (oracle/bleu_rouge.py): hand-derived known answers
(Papineni et al. 2002 clipping example, Lin 2004 ROUGE-L example) and structural properties.  The reference itself holds
no test or golden vector for these rewards (parity unpinned)."""
import math

import numpy as np

from oracle import bleu_rouge as R


def test_bleu_clipping_example():
    # modified unigram precision of "the x7" against the two classic references is 2/7 (Papineni et al., sec. 2.1)
    refs = ['the cat is on the mat', 'there is a cat on the mat']
    b = R.sentence_bleu('the the the the the the the', refs)
    assert abs(b[0] - 2.0 / 7.0) < 1e-8
    # no bigram matches: BLEU_2 = sqrt(2/7 * tiny/6) (the smoothing constants of bleu_scorer.compute_score)
    assert abs(b[1] - math.sqrt((2 + R.TINY) / (7 + R.SMALL) * R.TINY / (6 + R.SMALL))) < 1e-15


def test_bleu_identity_brevity_and_closest_length():
    refs = ['a man is playing a guitar', 'a man plays the guitar on stage tonight']
    b = R.sentence_bleu('a man is playing a guitar', refs)
    assert all(abs(x - 1.0) < 1e-8 for x in b)
    # 3 words against reference lengths {6, 8}: closest is 6, BP = exp(1 - 6/3)
    b = R.sentence_bleu('a man is', refs)
    assert abs(b[0] - math.exp(1 - 1 / ((3 + R.TINY) / (6 + R.SMALL)))) < 1e-8 and abs(b[2] / b[0] - 1.0) < 1e-8
    # tie in |l - testlen| picks the shorter reference (min over (distance, length) pairs): lengths {6, 8}, test length 7
    b7 = R.sentence_bleu('a man is playing a guitar tonight', refs)
    p1 = (7 + R.TINY) / (7 + R.SMALL)
    assert abs(b7[0] - p1) < 1e-12                                # ratio 7/6 > 1: no brevity penalty
    assert R.sentence_bleu('', refs)[3] < 1e-3                   # empty hypothesis is finite and ~0
    got = R.evaluate_captions_bleu({0: refs, 1: refs}, ['a man is playing a guitar', 'zzz'])
    assert got.shape == (2,) and got[0] > 0.999 and got[1] < 1e-4


def test_rouge_l_lin_2004_example():
    # Lin (2004), sec. 3.1: S1 "police killed the gunman"; S2 "police kill the gunman" (LCS 3), S3 "the gunman kill police" (LCS 2)
    assert R.my_lcs('police kill the gunman'.split(), 'police killed the gunman'.split()) == 3
    assert R.my_lcs('the gunman kill police'.split(), 'police killed the gunman'.split()) == 2
    assert abs(R.rouge_l('police kill the gunman', ['police killed the gunman']) - 0.75) < 1e-12
    assert abs(R.rouge_l('the gunman kill police', ['police killed the gunman']) - 0.5) < 1e-12
    # max precision and max recall may come from different references
    s = R.rouge_l('a b c d', ['a b', 'a x c y d z w'])
    p, r = 3 / 4.0, 2 / 2.0
    assert abs(s - (1 + 1.44) * p * r / (r + 1.44 * p)) < 1e-12
    assert R.rouge_l('', ['a b']) == 0.0 and R.rouge_l('q', ['a b']) == 0.0
    assert R.rouge_l('a b', ['a b']) == 1.0


def test_lcs_properties():
    rng = np.random.RandomState(0)
    for _ in range(200):
        a = rng.randint(0, 5, rng.randint(0, 12)).tolist()
        b = rng.randint(0, 5, rng.randint(0, 12)).tolist()
        l = R.my_lcs(a, b)
        assert l == R.my_lcs(b, a) and 0 <= l <= min(len(a), len(b))
        assert R.my_lcs(a, a) == len(a)
