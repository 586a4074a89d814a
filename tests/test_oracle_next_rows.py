"""Oracle pinning (CPU) for the SURVEY 8(f) rows: the committed golden vectors of the attention decoder and of the BLEU-4 /
ROUGE-L rewards are reproduced by the oracle, plus structural checks of the attention restatement."""
import os

import numpy as np

from oracle import attention_numpy as A
from oracle import bleu_rouge as R
from oracle import philox, text

G = os.path.join(os.path.dirname(__file__), 'golden')


def test_attention_oracle_reproduces_golden():
    g = np.load(os.path.join(G, 'next_rows_golden.npz'))
    D, H, V, n, Tc, B = [int(x) for x in g['att_dims']]
    p = A.init_params(D, H, V, seed=16)
    drop = np.stack([philox.dropout_mask(5, philox.STREAM_DROP1, np.arange(B), t, H, 0.9) for t in range(Tc)])
    loss, reg, logits = A.build_model_loss(p, g['att_video'], g['att_cap'], g['att_mask'], drop, return_logits=True)
    np.testing.assert_allclose([loss, reg], g['att_loss'], rtol=1e-12)
    np.testing.assert_allclose(logits, g['att_logits'], rtol=1e-5, atol=1e-6)
    ids, alphas = A.build_sampler(p, g['att_video'], Tc)
    assert (ids == g['att_ids']).all()
    np.testing.assert_allclose(alphas, g['att_alphas'], atol=1e-6)


def test_attention_oracle_structure():
    p = A.init_params(32, 24, 50, seed=1)
    rng = np.random.RandomState(0)
    v = rng.rand(3, 5, 32)
    ids, al = A.build_sampler(p, v, 6)
    np.testing.assert_allclose(al.sum(1), 1.0, atol=1e-12)          # alphas are a distribution over the frames
    cap = ids.astype(np.int32); mask = np.ones((3, 6))
    loss, reg = A.build_model_loss(p, v, cap, mask)
    assert reg == 0.0                                                # 5 frames: alphas[:, 0:8] covers them all, hinge = max(0, 0.5 - 1)
    # teacher forcing on the greedy ids reproduces the sampler's logits (same recurrence, no dropout)
    _, _, lg = A.build_model_loss(p, v, cap, mask, return_logits=True)
    _, _, ls = A.build_sampler(p, v, 6, return_logits=True)
    np.testing.assert_allclose(lg, ls, rtol=1e-12)
    # permuting the frames permutes the alphas of step 0 (h_prev = 0: the scores depend on the frame only)
    perm = np.array([2, 0, 4, 1, 3])
    _, al2 = A.build_sampler(p, v[:, perm], 6)
    np.testing.assert_allclose(al2[0], al[0][perm], atol=1e-12)
    # 32 frames: the regulariser is active and equals beta * mean hinge
    v32 = rng.rand(2, 32, 32)
    loss32, reg32 = A.build_model_loss(p, v32, cap[:2], mask[:2])
    assert reg32 > 0 and loss32 > reg32


def test_reward_oracle_reproduces_golden():
    g = np.load(os.path.join(G, 'next_rows_golden.npz'))
    sents = text.read_sentences(os.path.join(G, 'msvd_sents_train_noval_lc_nopunc.txt.gz'))
    by = {}
    for v, s in sents:
        by.setdefault(v, []).append(s)
    hyps = [str(h) for h in g['reward_hyps']]
    ref = {i: by[str(v)] for i, v in enumerate(g['reward_vids'])}
    np.testing.assert_allclose(R.bleu_all_orders(ref, hyps), g['bleu'], rtol=1e-13)
    np.testing.assert_array_equal(R.evaluate_captions_rouge(ref, hyps), g['rouge'])
    assert ((g['bleu'] >= 0) & (g['bleu'] <= 1 + 1e-9)).all() and ((g['rouge'] >= 0) & (g['rouge'] <= 1)).all()


def test_attention_torch_restatement_equals_numpy_and_finite_differences():
    from oracle import attention_torch as AT
    D, H, V, n, Tc, B = 20, 12, 30, 24, 4, 3
    p = A.init_params(D, H, V, seed=3)
    rng = np.random.RandomState(2)
    video = rng.rand(B, n, D); cap = rng.randint(0, V, (B, Tc)); mask = np.array([[1, 1, 1, 0], [1, 1, 0, 0], [1, 1, 1, 1]], dtype=np.float64)
    drop = np.stack([philox.dropout_mask(8, philox.STREAM_DROP1, np.arange(B), t, H, 0.9) for t in range(Tc)])
    want = A.build_model_loss(p, video, cap, mask, drop)
    loss, reg, grads, slice_sq = AT.loss_and_grads(p, video, cap, mask, drop)
    assert abs(loss - want[0]) < 1e-12 and abs(reg - want[1]) < 1e-12 and reg > 0
    # central finite differences of the NumPy restatement on a few coordinates of every variable
    for name in p:
        flat = p[name].reshape(-1)
        for idx in rng.choice(flat.size, size=min(3, flat.size), replace=False):
            old = flat[idx]
            flat[idx] = old + 1e-6; up = A.build_model_loss(p, video, cap, mask, drop)[0]
            flat[idx] = old - 1e-6; dn = A.build_model_loss(p, video, cap, mask, drop)[0]
            flat[idx] = old
            fd = (up - dn) / 2e-6
            assert abs(fd - grads[name].reshape(-1)[idx]) < 1e-6 + 1e-4 * abs(fd), (name, idx, fd, grads[name].reshape(-1)[idx])
    assert slice_sq > 0
