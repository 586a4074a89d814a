"""Oracle self-checks (CPU): hand-derived BPTT vs torch.autograd, TF-Adam / clip semantics, sampler streams."""
import numpy as np
import pytest
import torch

from oracle import s2vt_numpy as M
from oracle import philox

DIMS = dict(D=24, E=12, H=16, V=50)


def _torch_forward(tp, video, caption, drop1, drop2):
    N, Tv, D = video.shape
    Tc = caption.shape[1]
    E = tp['encode_image_W'].shape[1]
    H = tp['embed_word_W'].shape[0]
    img = (video.reshape(-1, D) @ tp['encode_image_W'] + tp['encode_image_b']).reshape(N, Tv, E)
    z = lambda n: torch.zeros(N, n, dtype=torch.float64)
    c1, h1, c2, h2 = z(H), z(H), z(H), z(H)

    def cell(x, c, h, W, b):
        gates = torch.cat([x, h], 1) @ W + b
        i, j, f, o = gates.chunk(4, 1)
        c = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
        return torch.tanh(c) * torch.sigmoid(o), c

    logits = []
    for t in range(Tv + Tc):
        x1 = img[:, t] if t < Tv else z(E)
        h1, c1 = cell(x1, c1, h1, tp[M.LSTM1_W], tp[M.LSTM1_B])
        out1 = h1 * drop1[t]
        if t < Tv:
            emb = z(E)
        else:
            i = t - Tv
            ids = torch.ones(N, dtype=torch.long) if i == 0 else caption[:, i - 1]
            emb = tp['Wemb'][ids]
        h2, c2 = cell(torch.cat([out1, emb], 1), c2, h2, tp[M.LSTM2_W], tp[M.LSTM2_B])
        out2 = h2 * drop2[t]
        if t >= Tv:
            logits.append(out2 @ tp['embed_word_W'] + tp['embed_word_b'])
    return torch.stack(logits, 0)


def _setup(N=5, Tv=3, Tc=6, seed=0):
    rng = np.random.RandomState(seed)
    p = M.init_params(seed=4, dtype=np.float64, **DIMS)
    for k in p:   # non-zero biases so their gradients are exercised
        if p[k].ndim == 1:
            p[k] = rng.uniform(-0.1, 0.1, p[k].shape)
    video = M.synthetic_features(N, Tv, DIMS['D'], dtype=np.float64)
    caption = rng.randint(0, DIMS['V'], size=(N, Tc))
    caption[0, 2:] = 0
    mask = np.zeros((N, Tc))
    for n in range(N):
        eos = np.where(caption[n] == 0)[0]
        L = (eos[0] + 1) if len(eos) else Tc
        mask[n, :L] = 1
    keep = 0.9
    drop1 = np.stack([philox.dropout_mask(7, philox.STREAM_DROP1, np.arange(N), t, DIMS['H'], keep) for t in range(Tv + Tc)]).astype(np.float64)
    drop2 = np.stack([philox.dropout_mask(7, philox.STREAM_DROP2, np.arange(N), t, DIMS['H'], keep) for t in range(Tv + Tc)]).astype(np.float64)
    return p, video, caption, mask, drop1, drop2


def test_rl_gradients_match_autograd():
    p, video, caption, mask, drop1, drop2 = _setup()
    rng = np.random.RandomState(1)
    r, b = rng.uniform(0, 2, len(video)), rng.uniform(0, 2, len(video))
    loss, grads, aux = M.rl_objective(p, video, caption, mask, r, b, drop1, drop2)
    tp = {k: torch.tensor(v, requires_grad=True) for k, v in p.items()}
    logits = _torch_forward(tp, torch.tensor(video), torch.tensor(caption), torch.tensor(drop1), torch.tensor(drop2))
    lsm = torch.log_softmax(logits, -1)
    onehot = torch.nn.functional.one_hot(torch.tensor(caption).T, DIMS['V']).double()
    lossT = (lsm * onehot * torch.tensor(mask).T[:, :, None])                      # [Tc,N,V]  (:288)
    tl = -(lossT * torch.tensor(r - b)[None, :, None]).sum() / torch.tensor(mask).sum()
    tl.backward()
    assert abs(loss - tl.item()) < 1e-12
    for k in p:
        np.testing.assert_allclose(grads[k], tp[k].grad.numpy(), rtol=1e-8, atol=1e-12, err_msg=k)


def test_xe_gradients_match_autograd():
    p, video, caption, mask, drop1, drop2 = _setup(seed=3)
    loss, grads, aux = M.xe_objective(p, video, caption, mask, drop1, drop2)
    tp = {k: torch.tensor(v, requires_grad=True) for k, v in p.items()}
    logits = _torch_forward(tp, torch.tensor(video), torch.tensor(caption), torch.tensor(drop1), torch.tensor(drop2))
    V = DIMS['V']
    tmask = torch.tensor(mask)
    total = 0.0
    for i in range(caption.shape[1]):
        onehot = torch.nn.functional.one_hot(torch.tensor(caption[:, i]), V).double()
        target = onehot * 0.95 + 0.05 / V
        ce = -(target * torch.log_softmax(logits[i], -1)).sum(-1).mean()            # batch-mean scalar (Q3)
        total = total + (ce * tmask[:, i]).sum()
    wd = sum((v ** 2).sum() / 2 for k, v in tp.items() if 'bias' not in k) * 5e-5
    tl = total / tmask.sum() + wd
    tl.backward()
    assert abs(loss - tl.item()) < 1e-12
    for k in p:
        np.testing.assert_allclose(grads[k], tp[k].grad.numpy(), rtol=1e-8, atol=1e-12, err_msg=k)


def test_l2_set_follows_bias_substring():
    # Q4: only the LSTM '.../biases' are excluded; encode_image_b / embed_word_b are decayed.
    decayed = [k for k in M.PARAM_NAMES if 'bias' not in k]
    assert 'encode_image_b' in decayed and 'embed_word_b' in decayed and 'Wemb' in decayed
    assert M.LSTM1_B not in decayed and M.LSTM2_B not in decayed


def test_tf_adam_first_step_and_epsilon_placement():
    p = {'w': np.array([1.0, -2.0])}
    g = {'w': np.array([0.5, -1e-9])}
    opt = M.TFAdam(p)
    out = opt.apply(dict(p), g, lr=0.1)
    lr_t = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    m = 0.1 * g['w']; v = 0.001 * g['w'] ** 2
    np.testing.assert_allclose(out['w'], p['w'] - lr_t * m / (np.sqrt(v) + 1e-8), rtol=1e-15)
    # differs from torch.optim.Adam (eps added after bias correction) when |g| ~ eps
    tw = torch.tensor(p['w'], requires_grad=True)
    topt = torch.optim.Adam([tw], lr=0.1)
    tw.grad = torch.tensor(g['w']); topt.step()
    assert abs(out['w'][1] - tw.detach().numpy()[1]) > 1e-3


def test_clip_by_global_norm_and_slice_norm():
    g = {'Wemb': np.array([[3.0, 0.0], [0.0, 0.0]]), 'b': np.array([4.0])}
    c, gn = M.clip_by_global_norm(g, 2.5)
    assert abs(gn - 5.0) < 1e-12
    np.testing.assert_allclose(c['b'], [2.0])
    # R6: two duplicate slices [1.5,0]+[1.5,0] scatter-add to [3,0] but their slice norm is sqrt(4.5)
    c2, gn2 = M.clip_by_global_norm(g, 2.5, emb_slice_sqnorm=4.5)
    assert abs(gn2 - np.sqrt(4.5 + 16)) < 1e-12
    c3, gn3 = M.clip_by_global_norm(g, 100.0)
    np.testing.assert_allclose(c3['b'], g['b'])


def test_exponential_decay_staircase():
    assert M.exponential_decay(1e-6, 0, 1000) == 1e-6
    assert M.exponential_decay(1e-6, 999, 1000) == 1e-6
    assert M.exponential_decay(1e-6, 2000, 1000) == 0.25e-6


def test_greedy_equals_teacher_forced_argmax():
    p = M.init_params(seed=4, dtype=np.float64, **DIMS)
    video = M.synthetic_features(4, 3, DIMS['D'], dtype=np.float64)
    ids, lg = M.greedy_sampler(p, video, Tc=7, return_logits=True)
    logits, _ = M.teacher_forward(p, video, ids, keep_cache=False)
    np.testing.assert_allclose(logits, lg, rtol=1e-10, atol=1e-12)
    assert (np.argmax(logits, -1).T == ids).all()


def test_multinomial_sampler_distribution_and_reproducibility():
    p = M.init_params(seed=4, dtype=np.float32, logit_scale=5.0, **DIMS)
    video = M.synthetic_features(1, 2, DIMS['D'])
    n = 4000
    vid = np.repeat(video, n, axis=0)
    ids1, lg = M.multinomial_sampler(p, vid, seed=2024, global_rows=np.arange(n), Tc=1, return_logits=True)
    ids2 = M.multinomial_sampler(p, vid, seed=2024, global_rows=np.arange(n), Tc=1)
    assert (ids1 == ids2).all()
    prob = np.exp(M.log_softmax(lg[0, 0].astype(np.float64)))
    cnt = np.bincount(ids1[:, 0], minlength=DIMS['V'])
    exp = prob * n
    sel = exp > 5
    chi2 = ((cnt[sel] - exp[sel]) ** 2 / exp[sel]).sum()
    dof = sel.sum() - 1
    assert chi2 < dof + 5 * np.sqrt(2 * dof), (chi2, dof)


def test_philox_known_answer():
    # Random123 KAT for philox4x32-10: counter = key = 0 and the all-ones vector.
    o = philox.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in o] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xFFFFFFFF
    o = philox.philox4x32_10(f, f, f, f, f, f)
    assert [int(x) for x in o] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    u = philox.u32_to_uniform(np.array([0, 0xFFFFFFFF], dtype=np.uint32))
    assert 0.0 < u[0] < u[1] < 1.0


def test_attribute_head_grads():
    rng = np.random.RandomState(0)
    feats = rng.rand(3, 4, 10)
    labels = (rng.rand(3, 7) > 0.5).astype(np.float64)
    W, b = rng.randn(10, 7) * 0.1, rng.randn(7) * 0.1
    loss, g = M.attribute_loss(feats, labels, W, b)
    tW, tb = torch.tensor(W, requires_grad=True), torch.tensor(b, requires_grad=True)
    z = torch.tensor(feats).mean(1) @ tW + tb
    tl = torch.nn.functional.binary_cross_entropy_with_logits(z, torch.tensor(labels), reduction='sum') / (7 * 3)
    tl.backward()
    assert abs(loss - tl.item()) < 1e-12
    np.testing.assert_allclose(g['attr_W'], tW.grad.numpy(), rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(g['attr_b'], tb.grad.numpy(), rtol=1e-9, atol=1e-14)
