"""GPU parity tests of the model entry points (through the C ABI) against the NumPy oracle.

Tolerances (BASELINE.json north_star): teacher-forced logits / log-probs within 1e-3 relative in bf16 and 1e-5 in
fp32, measured norm-wise: max|gpu - oracle| / max|oracle| over the tensor; token ids exact wherever the oracle's
top-2 margin exceeds the logit tolerance.
"""
import os

import numpy as np
import pytest
import torch

from oracle import philox
from oracle import s2vt_numpy as M

pytestmark = pytest.mark.gpu

SMALL = dict(D=200, E=60, H=72, V=301)
FULL = dict(D=1536, E=500, H=1000, V=9972)
G = os.path.join(os.path.dirname(__file__), 'golden')


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def make(dims, Tv, Tc, precision, keep=1.0, params=None, **kw):
    import s2vt_b200
    m = s2vt_b200.Video_Caption_Generator(dim_image=dims['D'], n_words=dims['V'], word_dim=dims['E'], lstm_dim=dims['H'], batch_size=8,
                                          n_video_lstm_step=Tv, n_caption_lstm_step=Tc, dropout_rate=keep, precision=precision,
                                          max_videos=8, max_rows=16, **kw)
    if params is not None:
        restored = m.load_variables(params)
        assert len(restored) == 9
    return m


def rand_params(dims, seed=4, dtype=np.float32, bias_scale=0.1, **kw):
    p = M.init_params(seed=seed, dtype=dtype, **dims, **kw)
    rng = np.random.RandomState(seed + 1)
    for k in p:
        if p[k].ndim == 1 and not np.any(p[k]):
            p[k] = rng.uniform(-bias_scale, bias_scale, p[k].shape).astype(dtype)
    return p


def captions_and_mask(N, Tc, V, seed=3):
    rng = np.random.RandomState(seed)
    cap = rng.randint(2, V, size=(N, Tc))
    mask = np.ones((N, Tc), dtype=np.float32)
    for n in range(N):
        L = 1 + (n * 5) % Tc
        if n % 4 != 3:
            cap[n, L - 1:] = 0
            mask[n, L:] = 0
    return cap.astype(np.int32), mask


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-5), ('bf16', 1e-3)])
@pytest.mark.parametrize('dims,Tv', [(SMALL, 3), (FULL, 5)])
def test_teacher_forced_logits_and_logprobs(precision, tol, dims, Tv):
    Tc = 7 if dims is SMALL else 35
    N = 6
    p64 = rand_params(dims, dtype=np.float64)
    p32 = {k: v.astype(np.float32) for k, v in p64.items()}
    m = make(dims, Tv, Tc, precision, params=p32)
    video = M.synthetic_features(N, Tv, dims['D'])
    cap, mask = captions_and_mask(N, Tc, dims['V'])
    logp, logits = m.teacher_forward(video, cap, want_logits=True)
    ref_logits, _ = M.teacher_forward(p64, video.astype(np.float64), cap, keep_cache=False)
    ref_logp, _ = M.rl_logprobs(ref_logits, cap, np.ones_like(mask))
    e1 = rel_err(logits.cpu().numpy(), ref_logits)
    e2 = rel_err(logp.cpu().numpy(), ref_logp)
    print('\n[teacher_forced %s H=%d] logits rel err %.3e, logp rel err %.3e (tol %.0e)' % (precision, dims['H'], e1, e2, tol))
    # north-star tolerance, verbatim: 1e-5 (fp32 mode) / 1e-3 (bf16 mode) on logits AND log-probs.  The bf16 mode multiplies its
    # forward GEMMs on fp16 operands (10 mantissa bits, same tensor rate; bf16 operands gave 3.8e-3 .. 4.4e-3 on the logits).
    assert e2 < tol
    assert e1 < tol


def test_bf16_mode_kernels_match_fp16_rounded_oracle():
    """Isolates kernel correctness from operand quantisation: oracle run on weights / features rounded to fp16, the forward operand type of
    the bf16 mode (activations are rounded too inside the kernels, hence the remaining 1e-3)."""
    dims, Tv, Tc, N = SMALL, 3, 7, 6
    p = rand_params(dims)
    rnd = lambda a: torch.tensor(a).to(torch.float16).to(torch.float32).numpy()
    pq = {k: (rnd(v) if v.ndim == 2 else v) for k, v in p.items()}
    m = make(dims, Tv, Tc, 'bf16', params=p)
    video = M.synthetic_features(N, Tv, dims['D'])
    cap, mask = captions_and_mask(N, Tc, dims['V'])
    _, logits = m.teacher_forward(video, cap, want_logits=True)
    ref, _ = M.teacher_forward({k: v.astype(np.float64) for k, v in pq.items()}, rnd(video).astype(np.float64), cap, keep_cache=False)
    e = rel_err(logits.cpu().numpy(), ref)
    print('\n[bf16 mode vs fp16-rounded-weight oracle] rel err %.3e' % e)
    assert e < 1e-3


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_dropout_masks_match_philox_oracle(precision):
    dims, Tv, Tc, N, B = SMALL, 3, 7, 8, 4
    keep, seed, row_base = 0.9, 77, 5
    p = rand_params(dims, dtype=np.float64)
    m = make(dims, Tv, Tc, precision, keep=keep, params={k: v.astype(np.float32) for k, v in p.items()})
    video = M.synthetic_features(B, Tv, dims['D'])
    cap, mask = captions_and_mask(N, Tc, dims['V'])
    logp, logits = m.teacher_forward(video, cap, drop_seed=seed, row_base=row_base, want_logits=True)
    rows = row_base + np.arange(N)
    d1 = np.stack([philox.dropout_mask(seed, philox.STREAM_DROP1, rows, t, dims['H'], keep) for t in range(Tv + Tc)]).astype(np.float64)
    d2 = np.stack([philox.dropout_mask(seed, philox.STREAM_DROP2, rows, t, dims['H'], keep) for t in range(Tv + Tc)]).astype(np.float64)
    vid_rows = np.concatenate([video] * (N // B), 0).astype(np.float64)     # row n uses video n % B
    ref, _ = M.teacher_forward(p, vid_rows, cap, d1, d2, keep_cache=False)
    e = rel_err(logits.cpu().numpy(), ref)
    print('\n[dropout %s] logits rel err %.3e' % (precision, e))
    assert e < (1e-5 if precision == 'fp32' else 1e-3)
    logp0, _ = m.teacher_forward(video, cap, drop_seed=0)
    assert rel_err(logp0.cpu().numpy(), logp.cpu().numpy()) > 1e-3           # dropout really changes the result


def _margin_ok(logits, tol):
    s = np.sort(logits, axis=-1)
    return (s[..., -1] - s[..., -2]) > tol


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_greedy_and_sampled_ids(precision):
    g = np.load(os.path.join(G, 'oracle_golden.npz'))
    dims, Tv, Tc, B = FULL, 5, 35, 4
    pB = M.init_params(seed=4, dtype=np.float32, peaked_bias=g['peaked_bias'], logit_scale=3.0, **dims)
    m = make(dims, Tv, Tc, precision, params=pB)
    video = M.synthetic_features(B, Tv)
    ids = m.greedy(video).cpu().numpy()
    # oracle greedy (fp64 golden) + margins from the oracle's own logits
    ref_ids, ref_logits = M.greedy_sampler({k: v.astype(np.float64) for k, v in pB.items()}, video.astype(np.float64), Tc, return_logits=True)
    assert (ref_ids == g['greedy_ids']).all()
    # ids exact wherever the oracle's top-2 margin exceeds the logit tolerance (north star): 1e-5 / 1e-3 relative, x2 because a
    # margin is a difference of two logits
    tol = (2e-5 if precision == 'fp32' else 2e-3) * np.abs(ref_logits).max()
    # compare position by position until the first divergence of each row (after that the inputs differ)
    n_checked = n_match = 0
    for b in range(B):
        for t in range(Tc):
            ok = _margin_ok(ref_logits[t, b], tol)
            if ids[b, t] != ref_ids[b, t]:
                assert not ok, 'greedy id differs at a decisive margin (row %d step %d)' % (b, t)
                break
            n_checked += 1; n_match += 1
    print('\n[greedy %s] %d/%d positions identical to the oracle' % (precision, n_match, B * Tc))
    assert n_match >= 0.8 * B * Tc
    # K-sample rollout: Philox Gumbel-max stream shared with the oracle
    K, seed = 2, 2024
    samp, gr = m.rollout(video, K, seed)
    assert (gr.cpu().numpy() == ids).all()
    samp = samp.cpu().numpy()
    ref = M.multinomial_sampler(pB, np.concatenate([video] * K, 0), seed, np.arange(K * B), Tc)
    same_prefix = 0
    for r in range(K * B):
        d = np.nonzero(samp[r] != ref[r])[0]
        same_prefix += (d[0] if len(d) else Tc)
    print('[sample %s] identical prefix tokens %d / %d' % (precision, same_prefix, K * B * Tc))
    assert same_prefix >= (0.9 if precision == 'fp32' else 0.5) * K * B * Tc
    assert (samp[:B] != samp[B:]).any()                                     # different streams per sample row


def test_sampler_distribution_chi2():
    dims, Tv, Tc = SMALL, 2, 1
    p = rand_params(dims, logit_scale=8.0)
    m = make(dims, Tv, Tc, 'fp32', params=p, )
    video = M.synthetic_features(1, Tv, dims['D'])
    B = 8
    vid = np.repeat(video, B, axis=0)
    counts = np.zeros(dims['V'])
    n = 0
    for it in range(250):
        s = m.sample(vid, seed=1000 + it).cpu().numpy()
        counts += np.bincount(s[:, 0], minlength=dims['V']); n += B
    logits, _ = M.teacher_forward({k: v.astype(np.float64) for k, v in p.items()}, video.astype(np.float64), np.zeros((1, 1), np.int32), keep_cache=False)
    prob = np.exp(M.log_softmax(logits[0, 0]))
    exp = prob * n
    sel = exp > 5
    chi2 = ((counts[sel] - exp[sel]) ** 2 / exp[sel]).sum(); dof = sel.sum() - 1
    print('\n[chi2] %.1f for %d dof over %d draws' % (chi2, dof, n))
    assert chi2 < dof + 5 * np.sqrt(2 * dof)


def _grad_report(m, grads, tol, tag):
    worst = 0.0
    for k in M.PARAM_NAMES:
        got = m.variable(k, grad=True).cpu().numpy()
        e = rel_err(got, grads[k])
        worst = max(worst, e)
        print('   %-40s rel err %.3e  (|g|max %.3e)' % (k, e, np.abs(grads[k]).max()))
        assert e < tol, (tag, k, e)
    return worst


@pytest.mark.parametrize('precision,tol', [('fp32', 2e-4), ('bf16', 2e-2)])
@pytest.mark.parametrize('keep', [1.0, 0.9])
def test_rl_backward_gradients(precision, tol, keep):
    dims, Tv, Tc, B, K = SMALL, 3, 7, 4, 2
    N = B * K
    seed = 0 if keep == 1.0 else 99
    p = rand_params(dims, dtype=np.float64)
    m = make(dims, Tv, Tc, precision, keep=keep, params={k: v.astype(np.float32) for k, v in p.items()})
    video = M.synthetic_features(B, Tv, dims['D'])
    cap, mask = captions_and_mask(N, Tc, dims['V'])
    rng = np.random.RandomState(8)
    r, b = rng.uniform(0, 2, N).astype(np.float32), np.tile(rng.uniform(0, 2, B), K).astype(np.float32)
    loss = m.rl_backward(video, cap, mask, r, b, drop_seed=seed).item()
    d1 = d2 = None
    if keep < 1:
        d1 = np.stack([philox.dropout_mask(seed, philox.STREAM_DROP1, np.arange(N), t, dims['H'], keep) for t in range(Tv + Tc)]).astype(np.float64)
        d2 = np.stack([philox.dropout_mask(seed, philox.STREAM_DROP2, np.arange(N), t, dims['H'], keep) for t in range(Tv + Tc)]).astype(np.float64)
    vid_rows = np.concatenate([video] * K, 0).astype(np.float64)
    ref_loss, grads, aux = M.rl_objective(p, vid_rows, cap, mask, r, b, d1, d2)
    print('\n[rl_backward %s keep=%.1f] loss gpu %.6f oracle %.6f' % (precision, keep, loss, ref_loss))
    assert abs(loss - ref_loss) < tol * max(1.0, abs(ref_loss))
    _grad_report(m, grads, tol, 'rl')
    slice_sq = m.grads[m.n_params].item()
    assert abs(slice_sq - aux['emb_slice_sqnorm']) < max(tol, 1e-3) * aux['emb_slice_sqnorm']
    # the literal feed (one video row per caption row) gives the same gradients as the de-duplicated one
    g_dedup = m.grads[:m.n_params].clone()
    m.rl_backward(np.concatenate([video] * K, 0), cap, mask, r, b, drop_seed=seed)
    assert rel_err(m.grads[:m.n_params].cpu().numpy(), g_dedup.cpu().numpy()) < (1e-5 if precision == 'fp32' else 2e-2)


@pytest.mark.parametrize('precision,tol', [('fp32', 2e-4), ('bf16', 2e-2)])
def test_xe_backward_gradients(precision, tol):
    dims, Tv, Tc, N = SMALL, 3, 7, 6
    p = rand_params(dims, dtype=np.float64)
    m = make(dims, Tv, Tc, precision, params={k: v.astype(np.float32) for k, v in p.items()})
    video = M.synthetic_features(N, Tv, dims['D'])
    cap, mask = captions_and_mask(N, Tc, dims['V'])
    out = m.xe_backward(video, cap, mask).cpu().numpy()
    ref_loss, grads, aux = M.xe_objective(p, video.astype(np.float64), cap, mask)
    print('\n[xe_backward %s] loss gpu %.6f (wd %.6f) oracle %.6f (wd %.6f)' % (precision, out[0], out[1], ref_loss, aux['weight_decay']))
    assert abs(out[0] - ref_loss) < tol * abs(ref_loss) and abs(out[1] - aux['weight_decay']) < 1e-4 * aux['weight_decay']
    _grad_report(m, grads, tol, 'xe')


def test_optimizer_step_matches_tf_adam():
    dims, Tv, Tc, B, K = SMALL, 3, 7, 4, 2
    N = B * K
    p = rand_params(dims, dtype=np.float64)
    m = make(dims, Tv, Tc, 'fp32', params={k: v.astype(np.float32) for k, v in p.items()})
    video = M.synthetic_features(B, Tv, dims['D'])
    cap, mask = captions_and_mask(N, Tc, dims['V'])
    rng = np.random.RandomState(8)
    r, b = rng.uniform(0, 2, N).astype(np.float32), np.tile(rng.uniform(0, 2, B), K).astype(np.float32)
    opt = M.TFAdam(p)
    vid_rows = np.concatenate([video] * K, 0).astype(np.float64)
    clip = 0.05     # small enough that clipping is active
    for step in range(2):
        lr = M.exponential_decay(1e-3, step, 1)
        m.rl_backward(video, cap, mask, r, b)
        gn = m.optimizer_step(lr, clip, wemb_slice_norm=True)[0].item()
        _, grads, aux = M.rl_objective(p, vid_rows, cap, mask, r, b)
        clipped, ref_gn = M.clip_by_global_norm(grads, clip, emb_slice_sqnorm=aux['emb_slice_sqnorm'])
        assert ref_gn > clip
        p = opt.apply(p, clipped, lr)
        assert abs(gn - ref_gn) < 1e-3 * ref_gn, (gn, ref_gn)
        for k in M.PARAM_NAMES:
            e = rel_err(m.variable(k).cpu().numpy(), p[k])
            assert e < 1e-5, (step, k, e)
    print('\n[adam] two clipped TF-Adam steps reproduce the oracle parameters (global norm %.4f)' % gn)


def test_optimistic_restore_semantics():
    dims = SMALL
    m = make(dims, 3, 7, 'fp32')
    before = m.variable('Wemb').clone()
    p = rand_params(dims)
    bad = dict(p)
    bad['Wemb'] = p['Wemb'][:, :-1]            # wrong shape -> skipped silently
    bad['Variable'] = np.zeros(1, np.float32)  # unknown name -> skipped silently
    bad['s2vt/LSTM1/basic_lstm_cell/kernel'] = bad.pop(M.LSTM1_W)   # TF >= 1.2 alias accepted
    restored = m.load_variables(bad)
    assert 'Wemb' not in restored and 'Variable' not in restored and 's2vt/LSTM1/basic_lstm_cell/kernel' in restored
    assert torch.equal(m.variable('Wemb'), before)
    np.testing.assert_array_equal(m.variable(M.LSTM1_W).cpu().numpy(), p[M.LSTM1_W])


def test_caption_masks_match_reference_rule():
    m = make(SMALL, 3, 7, 'fp32')
    ids = np.array([[5, 6, 0, 7, 7, 0, 0], [0, 5, 5, 5, 5, 5, 5], [5, 5, 5, 5, 5, 5, 5]], dtype=np.int32)
    mask, lens = m.caption_masks(ids)
    assert mask.cpu().numpy().tolist() == [[1, 1, 1, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0], [1] * 7]
    assert lens.cpu().numpy().tolist() == [2, 0, 7]
