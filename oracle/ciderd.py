"""Oracle CIDEr-D.  TEST INFRASTRUCTURE.

Restates the algorithm of the un-vendored third-party scorer the reference calls:
``pyciderevalcap.ciderD.ciderD.CiderD(df='msvd')`` (cider_evaluation.py:9, 12, 33-39),
i.e. vrama91/cider ``ciderD_scorer.py`` (no version pinned by the reference, not a
submodule; its ``data/msvd.p`` document-frequency pickle is absent).  The published
algorithm (n=4, sigma=6):

    precook(s)      : s.split() -> counts of all 1..4-grams
    df[g]           : number of reference *sets* (videos) containing g;  ref_len = ln(#videos)
    counts2vec      : vec_n[g] = tf(g) * (ref_len - ln(max(1, df[g])));  norm_n = ||vec_n||
                      length = sum of tf over BIGRAMS only (the library's `if n == 1` after
                      `n = len(ngram)-1`)
    sim(h, r)       : val_n = sum_{g in h} min(vec_h[g], vec_r[g]) * vec_r[g]
                      val_n /= norm_h*norm_r  (only if both non-zero)
                      val_n *= e ** (-(len_h - len_r)^2 / (2 sigma^2))
    score(h, refs)  : 10 * mean_n( sum_r val_n ) / len(refs)

Document frequencies are rebuilt from the training references.
Pinned (softly) by the reference artefact msvd_best_captions, see tests/test_oracle_ciderd.py.
"""
import math
from collections import defaultdict

import numpy as np

N_GRAM = 4
SIGMA = 6.0


def precook(s, n=N_GRAM):
    words = s.split()
    counts = defaultdict(int)
    for k in range(1, n + 1):
        for i in range(len(words) - k + 1):
            counts[tuple(words[i:i + k])] += 1
    return counts


class CiderD:
    """Drop-in for ``CiderD(df=<corpus pickle>)``: df / ref_len fixed at construction."""

    def __init__(self, ref_sets, n=N_GRAM, sigma=SIGMA):
        """ref_sets: iterable of lists of reference strings, one list per video (the df corpus)."""
        self.n = n
        self.sigma = sigma
        self.document_frequency = defaultdict(float)
        count = 0
        for refs in ref_sets:
            count += 1
            cooked = [precook(r, n) for r in refs]
            for ngram in set(g for ref in cooked for g in ref.keys()):
                self.document_frequency[ngram] += 1
        self.ref_len = np.log(float(count))

    def counts2vec(self, cnts):
        vec = [defaultdict(float) for _ in range(self.n)]
        length = 0
        norm = [0.0 for _ in range(self.n)]
        for ngram, term_freq in cnts.items():
            df = np.log(max(1.0, self.document_frequency.get(ngram, 0.0)))
            n = len(ngram) - 1
            vec[n][ngram] = float(term_freq) * (self.ref_len - df)
            norm[n] += pow(vec[n][ngram], 2)
            if n == 1:
                length += term_freq
        norm = [np.sqrt(x) for x in norm]
        return vec, norm, length

    def sim(self, vec_hyp, vec_ref, norm_hyp, norm_ref, length_hyp, length_ref):
        delta = float(length_hyp - length_ref)
        val = np.array([0.0 for _ in range(self.n)])
        for n in range(self.n):
            for ngram in vec_hyp[n].keys():
                val[n] += min(vec_hyp[n][ngram], vec_ref[n].get(ngram, 0.0)) * vec_ref[n].get(ngram, 0.0)
            if norm_hyp[n] != 0 and norm_ref[n] != 0:
                val[n] /= (norm_hyp[n] * norm_ref[n])
            assert not math.isnan(val[n])
            val[n] *= np.e ** (-(delta ** 2) / (2 * self.sigma ** 2))
        return val

    def score_one(self, hyp, refs, cooked_refs=None):
        vec, norm, length = self.counts2vec(precook(hyp, self.n))
        score = np.array([0.0 for _ in range(self.n)])
        if cooked_refs is None:
            cooked_refs = [self.counts2vec(precook(r, self.n)) for r in refs]
        for vec_ref, norm_ref, length_ref in cooked_refs:
            score += self.sim(vec, vec_ref, norm, norm_ref, length, length_ref)
        score_avg = np.mean(score)
        score_avg /= len(cooked_refs)
        score_avg *= 10.0
        return score_avg

    def compute_score(self, gts, res):
        """Same call shape as CiderD.compute_score(gts: {id: [refs]}, res: [{'image_id', 'caption': [s]}])."""
        scores = []
        cache = {}
        for r in res:
            iid = r['image_id']
            refs = gts[iid]
            key = id(refs)
            if key not in cache:
                cache[key] = [self.counts2vec(precook(x, self.n)) for x in refs]
            scores.append(self.score_one(r['caption'][0], refs, cache[key]))
        scores = np.array(scores)
        return float(np.mean(scores)) if len(scores) else 0.0, scores


def evaluate_captions_cider(scorer, ref, cand):
    """cider_evaluation.py:60-87: ref {i: [str]}, cand [str] -> float64 [N] CIDEr-D per hypothesis."""
    hypo = []
    refe = {}
    for i, caption in enumerate(cand):
        hypo.append({'image_id': i, 'caption': [caption]})
        refe[i] = ref[i]
    _, scores = scorer.compute_score(refe, hypo)
    return scores


def ngram_count_table(tokens, n=N_GRAM):
    """Exact n-gram multiset of a token-id list: dict {tuple(ids): count} (for the bit-exact count test)."""
    counts = defaultdict(int)
    for k in range(1, n + 1):
        for i in range(len(tokens) - k + 1):
            counts[tuple(int(x) for x in tokens[i:i + k])] += 1
    return dict(counts)
