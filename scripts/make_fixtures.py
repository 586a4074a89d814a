#!/usr/bin/env python
"""Generate the committed fixtures under tests/golden/ (run in the build container, where /root/reference exists).

1. Data fixtures (the reference's entry-point *data* surface, not sources): the vocabulary file, the MSVD
   training sentences and the msvd_best_captions artefact are stored gzip-compressed so the GPU box -- which has
   no /root/reference -- can run the parity tests, smoke() and bench.py.
2. Oracle golden vectors (seeded; the reference ships no golden vectors and cannot run here, so these pin the
   oracle against regressions -- they are NOT reference outputs): teacher-forced logits / log-probs, greedy ids,
   Philox-sampled ids, RL loss + gradient checksums + post-Adam parameters, XE loss, beam sentences, CIDEr-D.
"""
import gzip
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = '/root/reference'
OUT = os.path.join(ROOT, 'tests', 'golden')


def gz_copy(name, dst=None):
    dst = os.path.join(OUT, (dst or name) + '.gz')
    with open(os.path.join(REF, name), 'rb') as f, gzip.GzipFile(dst, 'wb', mtime=0) as g:
        shutil.copyfileobj(f, g)
    print('wrote', dst, os.path.getsize(dst))


def data_fixtures():
    gz_copy('msvd_vocabulary1.txt')
    gz_copy('msvd_sents_train_noval_lc_nopunc.txt')
    gz_copy('msvd_best_captions')
    gz_copy('train_most_freq_vocab_400_truncated.txt')


def golden_vectors():
    from oracle import s2vt_numpy as M, philox, beam, ciderd, text
    dims = dict(D=1536, E=500, H=1000, V=9972)
    out = {}
    for dt, tag in ((np.float64, 'f64'), (np.float32, 'f32')):
        p = M.init_params(seed=4, dtype=dt, **dims)
        video = M.synthetic_features(4, 5, dtype=dt)
        rng = np.random.RandomState(11)
        cap = rng.randint(2, dims['V'], size=(4, 35))
        for n, L in enumerate((3, 9, 35, 1)):
            cap[n, L - 1:] = 0 if L < 35 else cap[n, L - 1:]
        mask = np.array(text.decode_captions_masks(cap, {i: ('<eos>' if i == 0 else 'w%d' % i) for i in range(dims['V'])})[0], dtype=dt)
        logits, _ = M.teacher_forward(p, video, cap, keep_cache=False)
        logp, _ = M.rl_logprobs(logits, cap, mask)
        out['tf_logits_strided_' + tag] = logits[:, :, ::97].astype(dt)
        out['tf_logits_sum_' + tag] = logits.sum(-1)
        out['tf_logp_' + tag] = logp
        if tag == 'f64':
            out['caption'] = cap.astype(np.int32)
            out['mask'] = mask.astype(np.float32)
            r = np.array([1.0, 0.2, 0.7, 0.05]); b = np.array([0.5, 0.5, 0.1, 0.3])
            loss, grads, aux = M.rl_objective(p, video, cap, mask, r, b)
            out['rl_loss_f64'] = np.array(loss)
            out['rl_grad_sqnorm_f64'] = np.array([float((grads[k] ** 2).sum()) for k in M.PARAM_NAMES])
            out['rl_grad_sum_f64'] = np.array([float(grads[k].sum()) for k in M.PARAM_NAMES])
            out['rl_emb_slice_sqnorm_f64'] = np.array(aux['emb_slice_sqnorm'])
            xl, xg, xa = M.xe_objective(p, video, cap, mask)
            out['xe_loss_f64'] = np.array(xl)
            out['xe_grad_sqnorm_f64'] = np.array([float((xg[k] ** 2).sum()) for k in M.PARAM_NAMES])
            out['rewards'] = r; out['base_line'] = b
    # peaked 'set B' weights: greedy / sampled / beam ids
    sents = text.read_sentences(os.path.join(REF, 'msvd_sents_train_noval_lc_nopunc.txt'))
    vocab = text.read_vocabulary(os.path.join(REF, 'msvd_vocabulary1.txt'))
    w2i, i2w = text.build_word_vocab(vocab)
    counts = np.zeros(dims['V'])
    for _, s in sents:
        for w in s.split(' '):
            counts[w2i.get(w, 2)] += 1
        counts[0] += 1
    bias = np.log(counts + 1.0)
    pB = M.init_params(seed=4, dtype=np.float64, peaked_bias=bias, logit_scale=3.0, **dims)
    video = M.synthetic_features(4, 5, dtype=np.float64)
    out['peaked_bias'] = bias
    out['greedy_ids'] = M.greedy_sampler(pB, video).astype(np.int32)
    out['sampled_ids'] = M.multinomial_sampler(pB, video, 2024, np.arange(4)).astype(np.int32)
    for k in (3, 5):
        for lnf in (0.0, 1.0):
            res = []
            for v in range(2):
                s1, s2 = M.beam_initial_states(pB, video[v:v + 1])
                sent, lp, sc = beam.beam_search(M.beam_step_fn(pB, k), s1, s2, k, 35, lnf)
                res.append(dict(sentence=[int(x) for x in sent], logprob=lp, score=sc))
            out['beam_k%d_lnf%d' % (k, int(lnf))] = np.frombuffer(json.dumps(res).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, 'oracle_golden.npz'), **out)
    print('wrote oracle_golden.npz', os.path.getsize(os.path.join(OUT, 'oracle_golden.npz')))


def cider_golden():
    from oracle import ciderd, text
    sents = text.read_sentences(os.path.join(REF, 'msvd_sents_train_noval_lc_nopunc.txt'))
    by, vids = {}, []
    for v, s in sents:
        if v not in by:
            by[v] = []; vids.append(v)
        by[v].append(s)
    sc = ciderd.CiderD([by[v] for v in vids])
    rng = np.random.RandomState(5)
    hyps, hv, scores = [], [], []
    for j in range(64):
        v = vids[j]
        pool = by[v]
        for k in range(5):
            if k == 0:
                h = pool[rng.randint(len(pool))]
            elif k == 1:
                w = pool[rng.randint(len(pool))].split(); rng.shuffle(w); h = ' '.join(w)
            elif k == 2:
                h = ' '.join(pool[rng.randint(len(pool))].split()[:2] + by[vids[(j + 7) % 1200]][0].split()[1:])
            elif k == 3:
                h = '' if j % 8 == 0 else 'a ' * (1 + j % 5) + pool[0].split()[-1]
            else:
                h = by[vids[(j + 13) % 1200]][rng.randint(3)]
            hyps.append(h); hv.append(v)
            scores.append(sc.score_one(h, by[v]))
    with gzip.open(os.path.join(OUT, 'ciderd_golden.json.gz'), 'wt') as f:
        json.dump(dict(hyps=hyps, vids=hv, scores=scores), f)
    print('wrote ciderd_golden.json.gz', np.mean(scores))


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    what = sys.argv[1:] or ['data', 'golden', 'cider']
    if 'data' in what:
        data_fixtures()
    if 'golden' in what:
        golden_vectors()
    if 'cider' in what:
        cider_golden()
