"""GPU parity: BLEU-4 and ROUGE-L reward kernels (csrc/rewards.cu) against oracle/bleu_rouge.py on the real MSVD
references: the committed CIDEr golden hypotheses, every training reference of 60 videos scored against its own video,
and edge cases (empty, all-OOV, no <eos>, repeated words)."""
import gzip
import json
import os

import numpy as np
import pytest
import torch

from oracle import bleu_rouge as R
from oracle import text as otext

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def msvd():
    import s2vt_b200
    sents = otext.read_sentences(os.path.join(G, 'msvd_sents_train_noval_lc_nopunc.txt.gz'))
    by, vids = {}, []
    for v, s in sents:
        if v not in by:
            by[v] = []; vids.append(v)
        by[v].append(s)
    vocab = otext.read_vocabulary(os.path.join(G, 'msvd_vocabulary1.txt.gz'))
    w2i, i2w = otext.build_word_vocab(vocab)
    bleu = s2vt_b200.rewards.Bleu4([by[v] for v in vids], w2i)
    rouge = s2vt_b200.rewards.RougeL([by[v] for v in vids], w2i)
    return by, vids, w2i, i2w, bleu, rouge


def _cases(by, vids):
    with gzip.open(os.path.join(G, 'ciderd_golden.json.gz'), 'rt') as f:
        gold = json.load(f)
    vidx = {v: i for i, v in enumerate(vids)}
    hyps, rows = list(gold['hyps']), [vidx[v] for v in gold['vids']]
    for i, v in enumerate(vids[:60]):
        for s in by[v]:
            hyps.append(s); rows.append(i)
    edge = ['', 'zzzunknownword', 'a a a a a a a a a a', 'a man', 'the', ' '.join(['a man is playing a guitar'] * 5),
            'a man is playing a zzzunknownword guitar qqq']
    for j, h in enumerate(edge):
        hyps.append(h); rows.append(j % 7)
    return hyps, np.asarray(rows, dtype=np.int32)


def test_bleu_matches_oracle(msvd):
    by, vids, w2i, i2w, bleu, rouge = msvd
    hyps, rows = _cases(by, vids)
    want = R.bleu_all_orders({i: by[vids[r]] for i, r in enumerate(rows)}, hyps)
    L = max(len(h.split()) for h in hyps) + 1
    ids = np.zeros((len(hyps), L), dtype=np.int32)
    for i, h in enumerate(hyps):
        t = bleu._ids(h)
        ids[i, :len(t)] = t
    got = bleu.score_all_orders(torch.from_numpy(ids), rows).cpu().numpy()
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
    print('\n[bleu] %d hypotheses, max rel err %.3e, mean BLEU-4 %.4f' % (len(hyps), rel.max(), want[:, 3].mean()))
    assert rel.max() < 1e-12
    np.testing.assert_allclose(bleu.score_strings(hyps, rows).cpu().numpy(), want[:, 3], rtol=1e-12, atol=0)
    ref = {i: by[vids[r]] for i, r in enumerate(rows[:50])}
    np.testing.assert_allclose(bleu.evaluate_captions_cider(ref, hyps[:50]), R.evaluate_captions_bleu(ref, hyps[:50]), rtol=1e-12, atol=0)


def test_rouge_matches_oracle(msvd):
    by, vids, w2i, i2w, bleu, rouge = msvd
    hyps, rows = _cases(by, vids)
    want = R.evaluate_captions_rouge({i: by[vids[r]] for i, r in enumerate(rows)}, hyps)
    got = rouge.score_strings(hyps, rows).cpu().numpy()
    err = np.abs(got - want).max()
    print('\n[rouge] %d hypotheses, max |gpu - oracle| = %.3e, mean ROUGE-L %.4f' % (len(hyps), err, want.mean()))
    assert err < 1e-15 or np.array_equal(got, want)             # same fp64 operations in the same order: bit-exact
    assert got[len(hyps) - 7] == 0.0                              # empty hypothesis


def test_rewards_at_rollout_shape_without_eos(msvd):
    """[320, 35] id tensors as the trainer passes them, including rows that never emit <eos> (35 words)."""
    by, vids, w2i, i2w, bleu, rouge = msvd
    rng = np.random.RandomState(3)
    common = [w2i[w] for w in 'a man is the woman playing and in on with of dog cat guitar'.split()]
    ids = rng.choice(common, size=(320, 35)).astype(np.int32)
    lens = rng.randint(0, 36, 320)
    for i, l in enumerate(lens):
        ids[i, l:] = 0
    rows = (np.arange(320) % 64).astype(np.int32)
    strs = [' '.join(i2w[int(t)] for t in row[:l]) for row, l in zip(ids, lens)]
    ref = {i: by[vids[r]] for i, r in enumerate(rows)}
    gb = bleu.score_ids(torch.from_numpy(ids).cuda(), rows).cpu().numpy()
    gr = rouge.score_ids(torch.from_numpy(ids).cuda(), rows).cpu().numpy()
    np.testing.assert_allclose(gb, R.evaluate_captions_bleu(ref, strs), rtol=1e-12, atol=0)
    np.testing.assert_allclose(gr, R.evaluate_captions_rouge(ref, strs), rtol=0, atol=1e-15)


def test_reinforce_trainer_accepts_the_alternative_rewards(msvd):
    import s2vt_b200
    by, vids, w2i, i2w, bleu, rouge = msvd
    m = s2vt_b200.Video_Caption_Generator(dim_image=64, n_words=len(w2i), word_dim=32, lstm_dim=64, batch_size=4, n_video_lstm_step=3,
                                          n_caption_lstm_step=35, precision='fp32', max_videos=4, max_rows=8)
    feats = torch.rand(4, 3, 64, device='cuda')
    for scorer in (bleu, rouge):
        tr = s2vt_b200.trainer.ReinforceTrainer(m, scorer, n_samples=2, start_learning_rate=1e-3)
        out = tr.step(feats, np.arange(4, dtype=np.int32))
        assert np.isfinite(out.cpu().numpy()).all()
        r = tr.last['rewards'].cpu().numpy()
        assert r.shape == (8,) and (r >= 0).all() and (r <= 1.0 + 1e-6).all()


def test_rewards_match_committed_golden(msvd):
    by, vids, w2i, i2w, bleu, rouge = msvd
    g = np.load(os.path.join(G, 'next_rows_golden.npz'))
    vidx = {v: i for i, v in enumerate(vids)}
    hyps = [str(h) for h in g['reward_hyps']]
    rows = np.array([vidx[str(v)] for v in g['reward_vids']], dtype=np.int32)
    np.testing.assert_allclose(bleu.score_strings(hyps, rows).cpu().numpy(), g['bleu'][:, 3], rtol=1e-12)
    np.testing.assert_allclose(rouge.score_strings(hyps, rows).cpu().numpy(), g['rouge'], rtol=0, atol=1e-15)


def test_full_metric_evaluator_matches_oracle(msvd):
    """score_all (cider_evaluation.py:14-30) on 40 videos: corpus BLEU_1..4, mean ROUGE_L, CIDEr with df from the scored refs."""
    import s2vt_b200
    from oracle import ciderd as ocider
    by, vids, w2i, i2w, bleu, rouge = msvd
    keys = vids[100:140]
    cand = {k: [by[k][3 % len(by[k])] if j % 3 else 'a man is ' + by[k][0]] for j, k in enumerate(keys)}
    ref = {k: by[k] for k in keys}
    got = s2vt_b200.rewards.evaluate_for_particular_captions(cand, ref, w2i)
    hyps = [cand[k][0] for k in keys]
    oref = {i: by[k] for i, k in enumerate(keys)}
    want_b = R.corpus_bleu(oref, hyps)
    for m, w in zip(('Bleu_1', 'Bleu_2', 'Bleu_3', 'Bleu_4'), want_b):
        assert abs(got[m] - w) < 1e-12 * max(1.0, w), m
    assert abs(got['ROUGE_L'] - R.evaluate_captions_rouge(oref, hyps).mean()) < 1e-12
    sc = ocider.CiderD([by[k] for k in keys])
    want_c = ocider.evaluate_captions_cider(sc, oref, hyps).mean()
    assert abs(got['CIDEr'] - want_c) < 1e-9
    print('\n[score_all] %s' % {k: round(v, 4) for k, v in got.items()})
