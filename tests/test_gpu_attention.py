"""GPU parity: temporal-attention decoder (csrc/attention.cu, original_attention.py) against oracle/attention_numpy.py --
teacher-forced loss, logits, greedy ids and attention weights; fp32 mode to 1e-5, bf16 mode to the operand-rounding bound."""
import numpy as np
import pytest
import torch

from oracle import attention_numpy as A
from oracle import philox

pytestmark = pytest.mark.gpu


def _setup(D, H, V, n, Tc, B, precision, keep=1.0, seed=16, peaked=True):
    import s2vt_b200
    p = A.init_params(D, H, V, seed=seed, dtype=np.float64)
    rng = np.random.RandomState(7)
    if peaked:   # separate the top-2 logits so that greedy ids are comparable across precisions
        p['embed_word_b'] = rng.normal(0, 2.0, size=V)
        p['embed_word_W'] = p['embed_word_W'] * 3
    m = s2vt_b200.attention.Video_Caption_Generator(dim_image=D, n_words=V, dim_hidden=H, batch_size=B, n_video_lstm_steps=n, n_caption_lstm_steps=Tc,
                                                    drop_out_rate=keep, precision=precision)
    restored = m.load_variables(p)
    assert len(restored) == 13
    video = np.maximum(0.0, rng.normal(0.25, 0.5, size=(B, n, D))).astype(np.float32)
    cap = rng.randint(0, V, size=(B, Tc)).astype(np.int32)
    lens = rng.randint(1, Tc + 1, size=B)
    mask = (np.arange(Tc)[None, :] < lens[:, None]).astype(np.float32)
    return p, m, video, cap, mask


@pytest.mark.parametrize('n', [5, 32])
def test_attention_fp32_matches_oracle(n):
    D, H, V, Tc, B = 96, 72, 300, 9, 6
    p, m, video, cap, mask = _setup(D, H, V, n, Tc, B, 'fp32')
    want_loss, want_reg, want_logits = A.build_model_loss(p, video, cap, mask, None, return_logits=True)
    out, logits = m.build_model(video, cap, mask, want_logits=True)
    out = out.cpu().numpy(); logits = logits.cpu().numpy()
    rel = np.abs(logits - want_logits).max() / np.abs(want_logits).max()
    print('\n[attention fp32 n=%d] logits rel err %.2e, loss %.6f vs %.6f, reg %.6f vs %.6f' % (n, rel, out[0], want_loss, out[1], want_reg))
    assert rel < 1e-5 and abs(out[0] - want_loss) < 1e-5 * max(1.0, abs(want_loss)) and abs(out[1] - want_reg) < 1e-5 * max(1.0, want_reg)
    ids, alphas = m.build_sampler(video)
    want_ids, want_alphas, want_gl = A.build_sampler(p, video, Tc, return_logits=True)
    assert (ids.cpu().numpy() == want_ids).all()
    np.testing.assert_allclose(alphas.cpu().numpy(), want_alphas, rtol=0, atol=1e-5)
    np.testing.assert_allclose(alphas.cpu().numpy().sum(1), 1.0, atol=1e-5)                     # attention weights of a row sum to 1
    assert (m.build_generator(video).cpu().numpy() == want_ids).all()


def test_attention_dropout_and_regulariser_fp32():
    D, H, V, n, Tc, B = 64, 40, 120, 32, 7, 5
    p, m, video, cap, mask = _setup(D, H, V, n, Tc, B, 'fp32', keep=0.9)
    seed, row_base = 99, 3
    drop = np.stack([philox.dropout_mask(seed, philox.STREAM_DROP1, row_base + np.arange(B), t, H, 0.9) for t in range(Tc)])
    want_loss, want_reg = A.build_model_loss(p, video, cap, mask, drop)
    out, _ = m.build_model(video, cap, mask, drop_seed=seed, row_base=row_base)
    out = out.cpu().numpy()
    assert want_reg > 0                                              # 32 frames: the first 8 rarely hold half of the attention mass
    assert abs(out[0] - want_loss) < 1e-5 * abs(want_loss) and abs(out[1] - want_reg) < 1e-5 * want_reg
    no_drop, _ = A.build_model_loss(p, video, cap, mask, None)
    assert abs(no_drop - want_loss) > 1e-4                           # the masks matter
    with pytest.raises(ValueError):
        m.build_model(video, cap, mask, drop_seed=0)


def test_attention_bf16_reference_dims():
    """dim_hidden 1000 / V 9972 / 32 frames (the shapes config 3 names), tcgen05 path."""
    D, H, V, n, Tc, B = 1536, 1000, 9972, 32, 6, 16
    p, m, video, cap, mask = _setup(D, H, V, n, Tc, B, 'bf16')
    want_loss, want_reg, want_logits = A.build_model_loss(p, video, cap, mask, None, return_logits=True)
    out, logits = m.build_model(video, cap, mask, want_logits=True)
    out = out.cpu().numpy(); logits = logits.cpu().numpy()
    rel = np.abs(logits - want_logits).max() / np.abs(want_logits).max()
    print('\n[attention bf16] logits rel err %.2e, loss %.5f vs %.5f' % (rel, out[0], want_loss))
    assert rel < 8e-3 and abs(out[0] - want_loss) < 5e-3 * abs(want_loss)
    ids, alphas = m.build_sampler(video)
    want_ids, want_alphas, wl = A.build_sampler(p, video, Tc, return_logits=True)
    ids = ids.cpu().numpy()
    top2 = np.sort(wl, axis=2)[:, :, -2:]
    margin = (top2[:, :, 1] - top2[:, :, 0]).T                         # [B, Tc]
    # ids exact wherever the oracle's top-2 margin exceeds the bf16 tolerance, up to the first disagreement of a row
    for b in range(B):
        for t in range(Tc):
            if ids[b, t] != want_ids[b, t]:
                assert margin[b, t] < 8e-3 * np.abs(wl).max(), (b, t, margin[b, t])
                break
    assert (ids == want_ids).mean() > 0.9
    np.testing.assert_allclose(alphas.cpu().numpy()[0], want_alphas[0], rtol=3e-2, atol=2e-3)    # step 0: h_prev = 0, only operand rounding


def test_attention_matches_committed_golden():
    """tests/golden/next_rows_golden.npz (scripts/make_fixtures_next_rows.py): loss with dropout, logits, greedy ids, alphas."""
    import os
    import s2vt_b200
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'next_rows_golden.npz'))
    D, H, V, n, Tc, B = [int(x) for x in g['att_dims']]
    m = s2vt_b200.attention.Video_Caption_Generator(dim_image=D, n_words=V, dim_hidden=H, batch_size=B, n_video_lstm_steps=n, n_caption_lstm_steps=Tc,
                                                    drop_out_rate=0.9, precision='fp32')
    m.load_variables(A.init_params(D, H, V, seed=16))
    out, logits = m.build_model(g['att_video'], g['att_cap'], g['att_mask'], drop_seed=5, want_logits=True)
    np.testing.assert_allclose(out.cpu().numpy(), g['att_loss'], rtol=1e-5)
    np.testing.assert_allclose(logits.cpu().numpy(), g['att_logits'], rtol=1e-4, atol=1e-5)
    ids, alphas = m.build_sampler(g['att_video'])
    assert (ids.cpu().numpy() == g['att_ids']).all()
    np.testing.assert_allclose(alphas.cpu().numpy(), g['att_alphas'], atol=1e-5)


@pytest.mark.parametrize('precision,tol', [('fp32', 2e-4), ('bf16', 5e-2)])
@pytest.mark.parametrize('n,keep', [(5, 1.0), (32, 0.9)])
def test_attention_gradients_match_autograd_oracle(precision, tol, n, keep):
    """optimizer.compute_gradients(tf_loss) (original_attention.py:432): all 13 gradients against torch.autograd of the float64
    restatement (max-norm relative error per variable), plus the IndexedSlices square norm of the Wemb gradient."""
    from oracle import attention_torch as AT
    D, H, V, Tc, B = 96, 72, 300, 8, 6
    p, m, video, cap, mask = _setup(D, H, V, n, Tc, B, precision, keep=keep)
    seed, row_base = 31, 2
    drop = None if keep >= 1.0 else np.stack([philox.dropout_mask(seed, philox.STREAM_DROP1, row_base + np.arange(B), t, H, keep) for t in range(Tc)])
    loss, reg, grads, slice_sq = AT.loss_and_grads(p, video, cap, mask, drop)
    out = m.xe_backward(video, cap, mask, drop_seed=seed if keep < 1.0 else 0, row_base=row_base).cpu().numpy()
    assert abs(out[0] - loss) < (1e-5 if precision == 'fp32' else 5e-3) * abs(loss)
    worst = {}
    for name in m.variables:
        got = m.variable(name, grad=True).cpu().numpy().astype(np.float64)
        want = grads[name].reshape(got.shape)
        worst[name] = np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)
    print('\n[attention grads %s n=%d keep=%.1f] ' % (precision, n, keep) + ', '.join('%s %.1e' % (k.split('/')[-1], v) for k, v in worst.items()))
    assert max(worst.values()) < tol, worst
    if n == 32:
        assert reg > 0                                               # the hinge regulariser contributes to the gradients checked above


def test_attention_optimizer_step_matches_tf_adam_with_slice_norm():
    from oracle import attention_torch as AT
    D, H, V, n, Tc, B = 64, 40, 50, 12, 7, 5                        # small vocabulary: duplicate tokens -> slice norm != dense norm
    p, m, video, cap, mask = _setup(D, H, V, n, Tc, B, 'fp32', keep=1.0)
    state = {'t': 0, 'm': {}, 'v': {}}
    params = {k: v.copy() for k, v in p.items()}
    for step in range(2):
        loss, reg, grads, slice_sq = AT.loss_and_grads(params, video, cap, mask, None)
        dense_sq = float((grads['Wemb'] ** 2).sum())
        assert abs(slice_sq - dense_sq) > 1e-3 * dense_sq
        params, gn = AT.clip_and_adam(params, grads, slice_sq, state, lr=1e-2, clip_norm=0.05)      # clip active: gn >> 0.05
        m.xe_backward(video, cap, mask)
        out = m.optimizer_step(1e-2, clip_norm=0.05).cpu().numpy()
        assert gn > 0.05 and abs(out[0] - gn) < 1e-4 * gn and abs(out[1] - loss) < 1e-5 * abs(loss)
        for name in m.variables:
            got = m.variable(name).cpu().numpy()
            np.testing.assert_allclose(got, params[name].reshape(got.shape), rtol=0, atol=2e-5 * max(1.0, np.abs(params[name]).max()))
    # the refreshed operand copies are in use: the next forward sees the updated weights
    want_loss, _ = A.build_model_loss(params, video, cap, mask, None)
    got_loss = m.build_model(video, cap, mask)[0].cpu().numpy()[0]
    assert abs(got_loss - want_loss) < 1e-4 * abs(want_loss)
    assert m.train_step(video, cap, mask, global_step=2, drop_seed=0).shape == (2,)
