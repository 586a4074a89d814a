"""GPU tests of the reference-named entry points on a tiny synthetic dataset written in the reference's file formats,
the checkpoint round trip, and the config-4 extras (attribute head, RL + XE mix)."""
import os

import numpy as np
import pytest
import torch

from oracle import s2vt_numpy as M

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def tiny(tmp_path_factory):
    d = tmp_path_factory.mktemp('tiny')
    rng = np.random.RandomState(0)
    vocab = ['<en_unk>'] + ['w%d' % i for i in range(40)]
    (d / 'vocab.txt').write_text('\n'.join(vocab) + '\n')
    D, Tv = 32, 3

    def write(split, nvid):
        with open(d / ('%s_feats.txt' % split), 'w') as f, open(d / ('%s_sents.txt' % split), 'w') as g:
            for v in range(1, nvid + 1):
                for k in range(Tv):
                    f.write('vid%d_frame_%d,' % (v, k) + ','.join('%.5f' % x for x in rng.rand(D)) + '\n')
                for r in range(4):
                    g.write('vid%d\t%s\n' % (v, ' '.join(vocab[1 + (v + r + j) % 40] for j in range(2 + (v + r) % 5))))
    write('train', 12); write('test', 5)
    return d, D, Tv


def _args(script_defaults, d, D, Tv, **kw):
    import s2vt_b200
    ap = s2vt_b200.cli.build_parser('t', script_defaults)
    a = ap.parse_args([])
    a.video_train_feature_file = str(d / 'train_feats.txt'); a.video_train_sent_file = str(d / 'train_sents.txt')
    a.video_test_feature_file = str(d / 'test_feats.txt'); a.video_test_sent_file = str(d / 'test_sents.txt')
    a.vocabulary_file = str(d / 'vocab.txt'); a.model_path = str(d / 'models'); a.out_file = str(d / 'out.txt')
    a.dim_image, a.lstm_dim, a.word_dim, a.n_video_lstm_step, a.n_caption_lstm_step = D, 24, 16, Tv, 8
    a.batch_size, a.n_epochs, a.max_iters, a.precision = 4, 1, 3, 'fp32'
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def test_three_stage_recipe_and_beam_scripts(tiny, monkeypatch):
    import s2vt_b200
    d, D, Tv = tiny
    monkeypatch.chdir(d)
    cli = s2vt_b200.cli
    cli.run_xe(_args(dict(model_name='xe'), d, D, Tv, start_learning_rate=1e-2))                       # README step 1
    ck = str(d / 'models' / 'xe-0')
    assert os.path.exists(ck + '.npz') and os.path.exists(d / 'new_vocab1_data' / 'wordtoix.npy')
    cli.run_rl(_args(dict(model_name='rl', n_samples=3, start_learning_rate=1e-4, clip_norm=5.0), d, D, Tv, restore=ck))   # step 2
    cli.run_stage3(_args(dict(model_name='s3', alpha=0.5, clip_norm=5.0), d, D, Tv, restore=str(d / 'models' / 'rl-0')))    # step 3
    score = cli.run_xe(_args(dict(), d, D, Tv, task='evaluate', restore=str(d / 'models' / 's3-0')))
    assert np.isfinite(score)
    lines = open(d / 'out.txt').read().split('\n')[:-1]
    assert len(lines) == 5 and all(l.startswith('vid') and '\t' in l for l in lines)   # 'vid<id>\t<caption>' (caption may be empty)
    hyps = cli.run_beam(_args(dict(task='test', beam_size=3, length_normalization_factor=0.0), d, D, Tv, task='test', restore=ck))
    assert len(hyps) == 5
    sc = cli.run_beam(_args(dict(task='test', beam_size=4, length_normalization_factor=1.0), d, D, Tv, task='evaluate', restore=ck))
    assert np.isfinite(sc)


def test_checkpoint_round_trip_and_optimizer_state():
    import s2vt_b200
    mk = lambda seed: s2vt_b200.Video_Caption_Generator(dim_image=32, n_words=50, word_dim=16, lstm_dim=24, batch_size=4, n_video_lstm_step=2,
                                                        n_caption_lstm_step=5, precision='fp32', max_videos=4, max_rows=4, seed=seed)
    a, b = mk(1), mk(2)
    a.adam_m.normal_(); a.adam_v.uniform_(); a.adam_step = 7
    import tempfile
    with tempfile.TemporaryDirectory() as t:
        path = s2vt_b200.checkpoint.save(a, os.path.join(t, 'ck'), global_step=42)
        restored, step = s2vt_b200.checkpoint.optimistic_restore(b, path, with_optimizer=True)
    assert len(restored) == 9 and step == 42 and b.adam_step == 7
    for name, (off, shp) in a.variables.items():                 # (the flat blocks have alignment gaps between variables)
        n = int(np.prod(shp))
        for blk in ('params', 'adam_m', 'adam_v'):
            assert torch.equal(getattr(a, blk)[off:off + n], getattr(b, blk)[off:off + n]), (name, blk)
    video = M.synthetic_features(4, 2, 32)
    assert torch.equal(a.greedy(video), b.greedy(video))


@pytest.mark.parametrize('precision,tol', [('fp32', 2e-4), ('bf16', 5e-2)])
def test_attribute_head_and_stage3_mix(precision, tol):
    """Config 4: sigmoid-CE attribute head (reinforce_multitask_e2e_attribute_loss.py:375-380) and -(1-l) RL + l XE."""
    import s2vt_b200
    dims = dict(D=200, E=60, H=72, V=301)
    A, B, Tv, Tc = 37, 6, 3, 7
    p = M.init_params(seed=4, dtype=np.float64, **dims)
    rng = np.random.RandomState(3)
    aW, ab = rng.uniform(-0.1, 0.1, (dims['D'], A)), rng.uniform(-0.1, 0.1, A)
    m = s2vt_b200.Video_Caption_Generator(dim_image=dims['D'], n_words=dims['V'], word_dim=dims['E'], lstm_dim=dims['H'], batch_size=B,
                                          n_video_lstm_step=Tv, n_caption_lstm_step=Tc, dropout_rate=1.0, n_attributes=A, precision=precision,
                                          max_videos=B, max_rows=B)
    named = {k: v.astype(np.float32) for k, v in p.items()}
    named['attr_W'] = aW.astype(np.float32); named['attr_b'] = ab.astype(np.float32)
    assert len(m.load_variables(named)) == 11
    video = M.synthetic_features(B, Tv, dims['D'])
    labels = (rng.rand(B, A) > 0.7).astype(np.float32)
    cap = rng.randint(2, dims['V'], size=(B, Tc)).astype(np.int32); cap[:, 4:] = 0
    mask = np.zeros((B, Tc), np.float32); mask[:, :5] = 1
    r, b = rng.uniform(0, 2, B).astype(np.float32), rng.uniform(0, 2, B).astype(np.float32)
    lam, alpha = 0.5, 0.05
    rl = m.rl_backward(video, cap, mask, r, b, grad_scale=1 - lam).item()
    xe = m.xe_backward(video, cap, mask, grad_scale=lam, accumulate=True).cpu().numpy()
    at = m.attribute_backward(video, labels, grad_scale=alpha).item()
    rl_ref, g_rl, _ = M.rl_objective(p, video.astype(np.float64), cap, mask, r, b)
    xe_ref, g_xe, _ = M.xe_objective(p, video.astype(np.float64), cap, mask)
    at_ref, g_at = M.attribute_loss(video.astype(np.float64), labels, aW, ab)
    assert abs(rl - (1 - lam) * rl_ref) < tol * max(1, abs(rl_ref)) and abs(at - at_ref) < tol * at_ref
    assert abs(xe[0] - ((1 - lam) * rl_ref + lam * xe_ref)) < tol * abs(xe_ref)      # loss_out accumulates the mix
    rel = lambda x, y: float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-30))
    for k in M.PARAM_NAMES:
        want = (1 - lam) * g_rl[k] + lam * g_xe[k]
        assert rel(m.variable(k, grad=True).cpu().numpy(), want) < tol, k
    assert rel(m.variable('attr_W', grad=True).cpu().numpy(), alpha * g_at['attr_W']) < tol
    assert rel(m.variable('attr_b', grad=True).cpu().numpy(), alpha * g_at['attr_b']) < tol


def test_attention_and_alternative_reward_scripts(tiny, monkeypatch):
    """original_attention.py (train -> checkpoint -> evaluate with the full-metric evaluator) and the bleu4_ / rouge_ RL scripts."""
    import s2vt_b200
    d, D, Tv = tiny
    monkeypatch.chdir(d)
    cli = s2vt_b200.cli
    a = _args(dict(model_name='_att', start_learning_rate=1e-2, decay_steps=10000, clip_norm=10.0, seed_num=16), d, D, Tv, max_iters=4)
    cli.run_attention(a)
    ck = str(d / 'models' / 'batch_size4_att-0')
    assert os.path.exists(ck + '.npz')
    rec = open(d / 'out.txt').read()
    assert 'Epoch 0' in rec and 'Bleu_4:' in rec and 'CIDEr:' in rec                                  # per-epoch metric record (:481-520)
    scores = cli.run_attention(_args(dict(), d, D, Tv, task='evaluate', restore=ck))
    assert set(scores) >= {'Bleu_1', 'Bleu_4', 'ROUGE_L', 'CIDEr'} and all(np.isfinite(v) for v in scores.values())
    lines = open(d / 'out.txt').read().split('\n')[:-1]
    assert len(lines) == 5 and all(l.startswith('vid') and '\t' in l for l in lines)
    for reward in ('bleu4', 'rouge'):
        cli.run_rl(_args(dict(model_name='rl_' + reward, n_samples=2, start_learning_rate=1e-4, clip_norm=5.0, reward=reward), d, D, Tv))
        assert os.path.exists(str(d / 'models' / ('rl_%s-0.npz' % reward)))


def test_feature_pipe_stages_batches_ahead():
    """trainer.FeaturePipe: slots are filled on the copy stream, handed out in order, and refuse to be over-filled."""
    import s2vt_b200
    pipe = s2vt_b200.trainer.FeaturePipe('cuda:0', 4, 3, 8)
    hosts = [torch.full((4, 3, 8), float(i)).pin_memory() for i in range(5)]
    idx = [torch.full((4,), i, dtype=torch.int32).pin_memory() for i in range(5)]
    pipe.put(hosts[0], idx[0]); pipe.put(hosts[1], idx[1])
    with pytest.raises(RuntimeError):
        pipe.put(hosts[2], idx[2])
    for i in range(5):
        v, vi, slot = pipe.get()
        out = (v.sum() + vi.sum()).item()                       # consume on the compute stream
        pipe.release(slot)
        assert out == i * 96 + i * 4
        if i + 2 < 5:
            pipe.put(hosts[i + 2], idx[i + 2])
    with pytest.raises(RuntimeError):
        pipe.get()
    short = torch.ones(2, 3, 8).pin_memory()                     # ragged last batch
    pipe.put(short, torch.zeros(2, dtype=torch.int32))
    v, vi, slot = pipe.get()
    assert tuple(v.shape) == (2, 3, 8) and tuple(vi.shape) == (2,)


def test_feature_pipe_fp16_host_cache_feeds_identical_numbers():
    """trainer.FeaturePipe staging from an fp16 host cache: in the tensor-core mode the frames are rounded to fp16 before the first product anyway, so the
    rollout is bit-identical to the float32 feed -- at half the host -> device bytes."""
    import s2vt_b200
    B, Tv, D = 8, 4, 96
    m = s2vt_b200.Video_Caption_Generator(dim_image=D, n_words=120, word_dim=32, lstm_dim=48, batch_size=B, n_video_lstm_step=Tv, n_caption_lstm_step=6,
                                          precision='bf16', max_videos=B, max_rows=2 * B, seed=3)
    video = torch.from_numpy(M.synthetic_features(B, Tv, D))
    pipe = s2vt_b200.trainer.FeaturePipe(m.device, B, Tv, D)
    idx = torch.arange(B, dtype=torch.int32)
    out = []
    for feed in (video.pin_memory(), video.to(torch.float16).pin_memory()):
        pipe.put(feed, idx)
        v, vi, slot = pipe.get()
        samp, greedy = m.rollout(v, 2, seed=5)
        pipe.release(slot)
        out.append((samp.cpu(), greedy.cpu(), v.clone().cpu()))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert torch.equal(out[1][2], video.to(torch.float16).to(torch.float32))
