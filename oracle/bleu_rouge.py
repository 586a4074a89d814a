"""TEST INFRASTRUCTURE -- CPU restatement of the BLEU-4 and ROUGE-L rewards of the reference's alternative RL scripts.

Call sites: bleu_evaluation.py:60-87 (`Bleu(4).compute_score(refe, hypo)`, reward = per-sentence `scores[3]`) and
rouge_evaluation.py:60-87 (`Rouge().compute_score(refe, hypo)`, reward = per-sentence `scores`), used by
bleu4_reinforcement_multisampling_tf_s2vt.py:792-808 / rouge_reinforcement_multisampling_tf_s2vt.py:792-808.

The arithmetic lives in the third-party `pycocoevalcap` package (tylin/coco-caption; path-hacked import
bleu_evaluation.py:4-6, NOT vendored in the reference, no version pinned).  Its published algorithm is restated here:
  * bleu/bleu_scorer.py  BleuScorer(n=4): precook / cook_refs / cook_test, compute_score(option='closest'), per-sentence
    list with tiny = 1e-15, small = 1e-9 smoothing and the brevity penalty exp(1 - 1/ratio) when ratio < 1;
  * rouge/rouge.py       Rouge(beta=1.2): LCS-based F-measure, max precision and max recall over the references.
PARITY UNPINNED: the reference holds no test or golden vector for these rewards.
"""
import math

import numpy as np

TINY, SMALL = 1e-15, 1e-9
BETA = 1.2


# ---- BLEU ------------------------------------------------------------------------------------------------------------
def precook(s, n=4):
    """bleu_scorer.precook: (len(words), {ngram tuple: count}) over s.split()."""
    words = s.split()
    counts = {}
    for k in range(1, n + 1):
        for i in range(len(words) - k + 1):
            g = tuple(words[i:i + k])
            counts[g] = counts.get(g, 0) + 1
    return len(words), counts


def cook_refs(refs, n=4):
    """bleu_scorer.cook_refs with eff=None: (list of reference lengths, {ngram: max count over the references})."""
    reflen, maxcounts = [], {}
    for ref in refs:
        rl, counts = precook(ref, n)
        reflen.append(rl)
        for g, c in counts.items():
            maxcounts[g] = max(maxcounts.get(g, 0), c)
    return reflen, maxcounts


def cook_test(test, cooked_refs, n=4):
    """bleu_scorer.cook_test with eff=None."""
    reflen, refmaxcounts = cooked_refs
    testlen, counts = precook(test, n)
    correct = [0] * n
    for g, c in counts.items():
        correct[len(g) - 1] += min(refmaxcounts.get(g, 0), c)
    return {'reflen': reflen, 'testlen': testlen, 'guess': [max(0, testlen - k + 1) for k in range(1, n + 1)], 'correct': correct}


def sentence_bleu(test, refs, n=4):
    """The per-sentence entries BleuScorer.compute_score(option='closest') appends to bleu_list: [BLEU_1 .. BLEU_n]."""
    comps = cook_test(test, cook_refs(refs, n), n)
    testlen = comps['testlen']
    reflen = min((abs(l - testlen), l) for l in comps['reflen'])[1]          # _single_reflen, option 'closest'
    out, bleu = [], 1.0
    for k in range(n):
        bleu *= (float(comps['correct'][k]) + TINY) / (float(comps['guess'][k]) + SMALL)
        out.append(bleu ** (1.0 / (k + 1)))
    ratio = (testlen + TINY) / (reflen + SMALL)
    if ratio < 1:
        out = [b * math.exp(1 - 1 / ratio) for b in out]
    return out


def evaluate_captions_bleu(ref, cand):
    """bleu_evaluation.evaluate_captions_cider (sic, :60-87): ref {i: [refs]}, cand [str] -> list of per-sentence BLEU-4."""
    return np.array([sentence_bleu(c, ref[i])[3] for i, c in enumerate(cand)], dtype=np.float64)


def bleu_all_orders(ref, cand):
    return np.array([sentence_bleu(c, ref[i]) for i, c in enumerate(cand)], dtype=np.float64)


def corpus_bleu(ref, cand, n=4):
    """BleuScorer.compute_score(option='closest')[0] -- the corpus-level [Bleu_1..Bleu_n] of score_all
    (cider_evaluation.py:14-30): per-sentence components summed, then the same smoothed product and brevity penalty."""
    tot_c, tot_g, testlen, reflen = [0] * n, [0] * n, 0, 0
    for i, c in enumerate(cand):
        comps = cook_test(c, cook_refs(ref[i], n), n)
        testlen += comps['testlen']
        reflen += min((abs(l - comps['testlen']), l) for l in comps['reflen'])[1]
        for k in range(n):
            tot_c[k] += comps['correct'][k]
            tot_g[k] += comps['guess'][k]
    bleus, bleu = [], 1.0
    for k in range(n):
        bleu *= float(tot_c[k] + TINY) / (tot_g[k] + SMALL)
        bleus.append(bleu ** (1.0 / (k + 1)))
    ratio = (testlen + TINY) / (reflen + SMALL)
    if ratio < 1:
        bleus = [b * math.exp(1 - 1 / ratio) for b in bleus]
    return bleus


# ---- ROUGE-L ---------------------------------------------------------------------------------------------------------
def my_lcs(string, sub):
    """rouge.my_lcs: length of the longest common subsequence of two token lists (dynamic programme)."""
    if len(string) < len(sub):
        sub, string = string, sub
    lengths = [[0] * (len(sub) + 1) for _ in range(len(string) + 1)]
    for j in range(1, len(sub) + 1):
        for i in range(1, len(string) + 1):
            if string[i - 1] == sub[j - 1]:
                lengths[i][j] = lengths[i - 1][j - 1] + 1
            else:
                lengths[i][j] = max(lengths[i - 1][j], lengths[i][j - 1])
    return lengths[len(string)][len(sub)]


def rouge_l(candidate, refs, beta=BETA):
    """Rouge.calc_score: tokens by split(" ") (so '' is one empty token), max precision / recall over the references."""
    token_c = candidate.split(" ")
    prec, rec = [], []
    for reference in refs:
        token_r = reference.split(" ")
        lcs = my_lcs(token_r, token_c)
        prec.append(lcs / float(len(token_c)))
        rec.append(lcs / float(len(token_r)))
    prec_max, rec_max = max(prec), max(rec)
    if prec_max != 0 and rec_max != 0:
        return ((1 + beta ** 2) * prec_max * rec_max) / float(rec_max + beta ** 2 * prec_max)
    return 0.0


def evaluate_captions_rouge(ref, cand):
    """rouge_evaluation.evaluate_captions_cider (sic, :60-87): -> float64 [N] per-sentence ROUGE-L."""
    return np.array([rouge_l(c, ref[i]) for i, c in enumerate(cand)], dtype=np.float64)
