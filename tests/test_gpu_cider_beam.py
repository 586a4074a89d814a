"""GPU parity: CIDEr-D reward kernel and batched beam search against the oracle / the reference's own artefact."""
import gzip
import json
import os

import numpy as np
import pytest
import torch

from oracle import beam as obeam
from oracle import ciderd as ocider
from oracle import s2vt_numpy as M
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
    scorer = s2vt_b200.cider.CiderD([by[v] for v in vids], w2i)
    return by, vids, w2i, i2w, scorer


def test_ciderd_matches_oracle_golden_and_counts_bit_exact(msvd):
    import s2vt_b200
    by, vids, w2i, i2w, scorer = msvd
    with gzip.open(os.path.join(G, 'ciderd_golden.json.gz'), 'rt') as f:
        gold = json.load(f)
    vidx = {v: i for i, v in enumerate(vids)}
    rows = np.array([vidx[v] for v in gold['vids']], dtype=np.int32)
    got = scorer.score_strings(gold['hyps'], rows).cpu().numpy()
    err = np.abs(got - np.array(gold['scores'])).max()
    print('\n[ciderd] %d hypotheses, max |gpu - oracle| = %.3e, mean score %.4f' % (len(got), err, got.mean()))
    assert err < 1e-5                                            # north-star tolerance; observed ~1e-15
    # exact n-gram multisets from token ids
    ids = np.zeros((len(gold['hyps']), 64), dtype=np.int32)
    for i, h in enumerate(gold['hyps']):
        t = [w2i[w] for w in h.split()]
        ids[i, :len(t)] = t
    scores, counts = scorer.score_ids(torch.from_numpy(ids), rows, want_counts=True)
    counts = counts.cpu().numpy()
    for i in range(len(ids)):
        toks = ids[i][:list(ids[i]).index(0)] if 0 in ids[i] else ids[i]
        ref = ocider.ngram_count_table(toks)
        mine = {s2vt_b200.cider.decode_ngram_key(k): int(c) for k, c in counts[i] if k != 0}
        assert mine == ref, i
    np.testing.assert_allclose(scores.cpu().numpy(), got, rtol=0, atol=1e-12)


def test_ciderd_reproduces_msvd_best_captions_on_gpu(msvd):
    """Full-size known-answer test: 48 774 hypotheses (every training reference scored against its own video)."""
    by, vids, w2i, i2w, scorer = msvd
    with gzip.open(os.path.join(G, 'msvd_best_captions.gz'), 'rt') as f:
        best = dict(line.rstrip('\n').split('\t') for line in f)
    hyps, rows = [], []
    for i, v in enumerate(vids):
        for s in by[v]:
            hyps.append(s); rows.append(i)
    scores = scorer.score_strings(hyps, np.array(rows, dtype=np.int32)).cpu().numpy()
    ok, pos = 0, 0
    for i, v in enumerate(vids):
        n = len(by[v])
        sc = scores[pos:pos + n]
        cider_score, one_best = 0.0, None
        for j in range(n):
            if sc[j] > cider_score:
                cider_score, one_best = sc[j], by[v][j]
        ok += int(one_best is not None and one_best.strip() == best[v].strip())
        pos += n
    print('\n[ciderd KAT] %d / 1200 lines of msvd_best_captions reproduced on the GPU' % ok)
    assert ok >= 1100
    # a slice against the oracle scorer
    osc = ocider.CiderD([by[v] for v in vids[:1200]])
    for i in (0, 17, 400):
        v = vids[i]
        start = sum(len(by[u]) for u in vids[:i])
        ref = [osc.score_one(s, by[v]) for s in by[v][:5]]
        np.testing.assert_allclose(scores[start:start + 5], ref, rtol=0, atol=1e-9)


def test_evaluate_captions_cider_drop_in_and_edge_cases(msvd):
    by, vids, w2i, i2w, scorer = msvd
    ref = {0: by[vids[3]], 1: by[vids[3]], 2: by[vids[10]], 3: by[vids[10]]}
    cand = [by[vids[3]][0], '', 'zzzunseenword qqqanother zzzunseenword', by[vids[3]][1]]
    got = scorer.evaluate_captions_cider(ref, cand)
    osc = ocider.CiderD([by[v] for v in vids])
    want = ocider.evaluate_captions_cider(osc, ref, cand)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)
    assert got[1] == 0.0 and got[2] == 0.0
    with pytest.raises(KeyError):
        scorer.evaluate_captions_cider({0: ['not a corpus video']}, ['a man'])


def _peaked_model(precision, dims, Tv, Tc, bias, scale, k):
    import s2vt_b200
    p = M.init_params(seed=4, dtype=np.float32, peaked_bias=bias, logit_scale=scale, **dims)
    m = s2vt_b200.Video_Caption_Generator(dim_image=dims['D'], n_words=dims['V'], word_dim=dims['E'], lstm_dim=dims['H'], batch_size=8,
                                          n_video_lstm_step=Tv, n_caption_lstm_step=Tc, dropout_rate=1.0, precision=precision, beam_size=k,
                                          max_videos=8, max_rows=16)
    m.load_variables(p)
    return m, p


@pytest.mark.parametrize('k,lnf', [(3, 0.0), (3, 1.0), (5, 0.0), (5, 1.0)])
def test_beam_search_matches_oracle_full_dims(k, lnf):
    g = np.load(os.path.join(G, 'oracle_golden.npz'))
    dims = dict(D=1536, E=500, H=1000, V=9972)
    m, p = _peaked_model('fp32', dims, 5, 35, g['peaked_bias'], 3.0, k)
    video = M.synthetic_features(4, 5)
    sent, lens, lp, sc = [x.cpu().numpy() for x in m.beam_search(video, k, lnf)]
    gold = json.loads(bytes(g['beam_k%d_lnf%d' % (k, int(lnf))]).decode())
    for v in range(2):
        s = sent[v, :lens[v]].tolist()
        print('\n[beam k=%d lnf=%g video %d] gpu %s lp %.5f | oracle %s lp %.5f' % (k, lnf, v, s, lp[v], gold[v]['sentence'], gold[v]['logprob']))
        assert s == gold[v]['sentence']
        assert abs(lp[v] - gold[v]['logprob']) < 1e-3 and abs(sc[v] - gold[v]['score']) < 1e-3


def test_beam_search_small_dims_many_videos_and_drop_in_step():
    """Path-dependent bookkeeping (B1-B3) on 8 videos with early and late <eos>, plus the single-hypothesis contracts."""
    dims = dict(D=64, E=40, H=48, V=60)
    Tv, Tc, k = 3, 12, 4
    rng = np.random.RandomState(2)
    bias = rng.uniform(-1, 1, dims['V']); bias[0] = 1.5
    m, p = _peaked_model('fp32', dims, Tv, Tc, bias, 20.0, k)
    p64 = {kk: v.astype(np.float64) for kk, v in p.items()}
    video = M.synthetic_features(8, Tv, dims['D'])
    n_final = 0
    for lnf in (0.0, 1.0):
        sent, lens, lp, sc = [x.cpu().numpy() for x in m.beam_search(video, k, lnf)]
        for v in range(8):
            s1, s2 = M.beam_initial_states(p64, video[v:v + 1].astype(np.float64))
            ref_sent, ref_lp, ref_sc = obeam.beam_search(M.beam_step_fn(p64, k), s1, s2, k, Tc, lnf)
            got = sent[v, :lens[v]].tolist()
            assert got == [int(x) for x in ref_sent], (lnf, v, got, ref_sent)
            assert abs(lp[v] - ref_lp) < 1e-3 and abs(sc[v] - ref_sc) < 1e-3
            n_final += int(got[-1] == 0)
    print('\n[beam small] 16 searches identical to the oracle, %d ended with <eos>' % n_final)
    assert 0 < n_final
    # drop-in: the reference's host loop (oracle.beam restates it) driven by the GPU beam_probability
    s1, s2 = m.beam_init(video[:1])
    o1, o2 = M.beam_initial_states(p64, video[:1].astype(np.float64))
    np.testing.assert_allclose(s1.cpu().numpy(), o1, atol=1e-5); np.testing.assert_allclose(s2.cpu().numpy(), o2, atol=1e-5)

    def step(state1, state2, word):
        idx, pr, n2, n1 = m.beam_probability(state2, state1, np.asarray(word, dtype=np.int32), k)
        return idx.cpu().numpy(), pr.cpu().numpy(), n2, n1

    got = obeam.beam_search(step, s1, s2, k, Tc, 0.0)
    ref = obeam.beam_search(M.beam_step_fn(p64, k), o1, o2, k, Tc, 0.0)
    assert [int(x) for x in got[0]] == [int(x) for x in ref[0]] and abs(got[1] - ref[1]) < 1e-3


def test_beam_search_bf16_runs_and_mostly_agrees():
    g = np.load(os.path.join(G, 'oracle_golden.npz'))
    dims = dict(D=1536, E=500, H=1000, V=9972)
    m, p = _peaked_model('bf16', dims, 5, 35, g['peaked_bias'], 3.0, 5)
    video = M.synthetic_features(4, 5)
    sent, lens, lp, sc = [x.cpu().numpy() for x in m.beam_search(video, 5, 0.0)]
    gold = json.loads(bytes(g['beam_k5_lnf0']).decode())
    same = sum(sent[v, :lens[v]].tolist() == gold[v]['sentence'] for v in range(2))
    print('\n[beam bf16] %d / 2 sentences identical to the fp64 oracle; logprob diff %s' % (same, [float(lp[v] - gold[v]['logprob']) for v in range(2)]))
    # identical sentences -> the log-probability is a sum of <= 35 log-softmax values, each within the bf16-mode tolerance
    assert all(abs(lp[v] - gold[v]['logprob']) < (2e-2 if sent[v, :lens[v]].tolist() == gold[v]['sentence'] else 0.2) for v in range(2))


@pytest.mark.parametrize('precision,Tv,B', [('fp32', 5, 8), ('bf16', 5, 8), ('bf16', 5, 1), ('bf16', 5, 3), ('bf16', 80, 64)])
def test_fused_beam_step_equals_unfused_launches(precision, Tv, B):
    """The fused step (vocabulary GEMM with the per-part top-k epilogue + one merge / bookkeeping / state-gather kernel per step)
    against the un-fused launches it replaces (materialised logits, topk_rows_kernel, beam_update_kernel, two gathers;
    s2vt_set_overlap bit 3).  Same GEMM mainloop -> the candidate logits are bit-identical; lse is summed in a different order."""
    import s2vt_b200
    g = np.load(os.path.join(G, 'oracle_golden.npz'))
    dims = dict(D=1536, E=500, H=1000, V=9972)
    p = M.init_params(seed=4, dtype=np.float32, peaked_bias=g['peaked_bias'], logit_scale=3.0, **dims)
    m = s2vt_b200.Video_Caption_Generator(dim_image=dims['D'], n_words=dims['V'], word_dim=dims['E'], lstm_dim=dims['H'], batch_size=B,
                                          n_video_lstm_step=Tv, n_caption_lstm_step=35, dropout_rate=1.0, precision=precision, beam_size=8,
                                          max_videos=B, max_rows=B)
    m.load_variables(p)
    video = M.synthetic_features(B, Tv)
    for k, lnf in ((5, 1.0), (5, 0.0), (3, 1.0), (8, 0.0), (1, 0.0)):
        m.lib.s2vt_set_overlap(m.h, 7)
        fused = [x.cpu().numpy() for x in m.beam_search(video, k, lnf)]
        m.lib.s2vt_set_overlap(m.h, 7 | 8)
        plain = [x.cpu().numpy() for x in m.beam_search(video, k, lnf)]
        m.lib.s2vt_set_overlap(m.h, 7)
        same = [v for v in range(B) if fused[1][v] == plain[1][v] and (fused[0][v] == plain[0][v]).all()]
        print('\n[beam fused vs un-fused %s T_v=%d k=%d lnf=%g] %d / %d sentences identical, max |logprob diff| %.2e'
              % (precision, Tv, k, lnf, len(same), B, np.abs(fused[2][same] - plain[2][same]).max()))
        assert len(same) >= B - B // 32          # a reordered fp32 lse may flip an exact near-tie between two hypotheses
        np.testing.assert_allclose(fused[2][same], plain[2][same], atol=2e-5)
        np.testing.assert_allclose(fused[3][same], plain[3][same], atol=2e-5)
        assert (fused[1] > 0).all()
