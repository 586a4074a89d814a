"""Size-independent properties of the hot path AT BASELINE.json's full sizes (config 2: 64 videos x K=5, T_v=80, T_c=35,
D=1536, E=500, H=1000, V=9972, bf16 -- the shape bench.py times; config 3 shape [B, 32, 1536] for the decode paths), where
the NumPy oracle would need minutes per case:
  * determinism: the same seeds give bit-identical samples and losses, gradients to atomic summation order;
  * row independence: a video's greedy / beam caption does not depend on which batch it is decoded in;
  * linearity of the REINFORCE gradient in (reward - baseline), and its sign symmetry;
  * encode -> decode round trip of the reward glue: masks cut at the first <eos>, CIDEr-D of a reference against its own video
    is the maximum over the batch of candidates.
"""
import gzip
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')
DIMS = dict(D=1536, E=500, H=1000, V=9972)
B, K, Tv, Tc = 64, 5, 80, 35


@pytest.fixture(scope='module')
def full():
    import s2vt_b200
    m = s2vt_b200.Video_Caption_Generator(dim_image=DIMS['D'], n_words=DIMS['V'], word_dim=DIMS['E'], lstm_dim=DIMS['H'], batch_size=B,
                                          n_video_lstm_step=Tv, n_caption_lstm_step=Tc, dropout_rate=0.9, precision='bf16', max_videos=B,
                                          max_rows=K * B, seed=4)
    rng = np.random.RandomState(0)
    m.variable('embed_word_b').copy_(torch.as_tensor(rng.normal(0, 2.0, DIMS['V']).astype(np.float32)))   # peaked logits: stable arg-max
    m.variable('embed_word_W').mul_(3.0)
    m.refresh()
    video = torch.from_numpy(np.maximum(0.0, rng.normal(0.25, 0.5, size=(B, Tv, DIMS['D']))).astype(np.float32)).cuda()
    return m, video


def _grads(m, video, samp, mask, r, b, seed):
    loss = m.rl_backward(video, samp, mask, r, b, drop_seed=seed).item()
    return loss, m.grads[:m.n_params].clone()


def test_determinism_at_full_size(full):
    m, video = full
    s1, g1 = m.rollout(video, K, seed=77)
    s2, g2 = m.rollout(video, K, seed=77)
    assert torch.equal(s1, s2) and torch.equal(g1, g2)
    s3, _ = m.rollout(video, K, seed=78)
    assert not torch.equal(s1, s3)                                  # another seed gives other samples
    assert not torch.equal(s1[:B], s1[B:2 * B])                     # the K samples of a video differ
    mask, lens = m.caption_masks(s1)
    r = torch.rand(K * B, device='cuda'); b = torch.rand(B, device='cuda').repeat(K)
    l1, gr1 = _grads(m, video, s1, mask, r, b, 5)
    l2, gr2 = _grads(m, video, s1, mask, r, b, 5)
    assert l1 == l2
    # the forward is bit-reproducible; the bias-gradient column sums and the embedding scatter combine partial sums with fp32
    # atomics, so gradients repeat to summation-order noise only
    assert ((gr1 - gr2).abs().max() / gr1.abs().max()).item() < 1e-5
    _, gr3 = _grads(m, video, s1, mask, r, b, 6)                    # another dropout seed changes the gradient
    assert ((gr1 - gr3).abs().max() / gr1.abs().max()).item() > 1e-3


def test_rows_are_independent_at_full_size(full):
    m, video = full
    whole = m.greedy(video)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).cuda()
    assert torch.equal(m.greedy(video[perm]), whole[perm])          # a caption does not depend on its batch neighbours
    part = m.greedy(video[:7].contiguous())                         # ragged batch (7 videos): other tile shapes, fp32 summation order
    agree = (part == whole[:7]).float().mean().item()
    assert agree > 0.97, agree
    sent, lens, lp, sc = m.beam_search(video[:16].contiguous(), 5, 1.0)
    sent2, lens2, lp2, sc2 = m.beam_search(video[8:24].contiguous(), 5, 1.0)
    assert torch.equal(sent[8:16], sent2[:8]) and torch.equal(lens[8:16], lens2[:8])
    torch.testing.assert_close(sc[8:16], sc2[:8], rtol=1e-5, atol=1e-6)
    assert (lens >= 1).all() and (lens <= Tc).all() and (sc <= 0).all()


def test_reinforce_gradient_is_linear_in_the_advantage(full):
    m, video = full
    samp, greedy = m.rollout(video, K, seed=3)
    mask, _ = m.caption_masks(samp)
    rng = torch.Generator(device='cuda').manual_seed(2)
    r = torch.rand(K * B, device='cuda', generator=rng); b = torch.rand(K * B, device='cuda', generator=rng)
    l1, g1 = _grads(m, video, samp, mask, r, b, 9)
    l2, g2 = _grads(m, video, samp, mask, 2 * r, 2 * b, 9)         # advantage doubled (exact in floating point)
    assert abs(l2 - 2 * l1) <= 1e-6 * abs(l1) + 1e-9
    rel = ((g2 - 2 * g1).abs().max() / g2.abs().max()).item()
    assert rel < 2e-3, rel                                           # bf16 rounding of the per-step gate gradients is not scale-free
    l3, g3 = _grads(m, video, samp, mask, b, r, 9)                  # advantage negated
    assert abs(l3 + l1) <= 1e-6 * abs(l1) + 1e-9
    assert ((g3 + g1).abs().max() / g1.abs().max()).item() < 2e-3
    lz, gz = _grads(m, video, samp, mask, r, r, 9)                  # zero advantage: zero gradient
    assert lz == 0.0 and gz.abs().max().item() == 0.0


def test_reward_glue_round_trip_at_full_size(full):
    import s2vt_b200
    from oracle import text as otext
    m, video = full
    sents = otext.read_sentences(os.path.join(G, 'msvd_sents_train_noval_lc_nopunc.txt.gz'))
    by, vids = {}, []
    for v, s in sents:
        if v not in by:
            by[v] = []; vids.append(v)
        by[v].append(s)
    w2i, i2w = otext.build_word_vocab(otext.read_vocabulary(os.path.join(G, 'msvd_vocabulary1.txt.gz')))
    scorer = s2vt_b200.cider.CiderD([by[v] for v in vids], w2i)
    # encode references of 64 videos as id rows -> masks -> decode: text survives, masks end at the first <eos>
    caps = [by[vids[j]][0] for j in range(B)]
    ids, mask = s2vt_b200.text.sentence_padding_toix(caps, w2i, Tc)
    dmask, lens = m.caption_masks(torch.from_numpy(ids).cuda())
    assert np.array_equal(dmask.cpu().numpy(), mask)
    assert s2vt_b200.text.decode_captions(ids, i2w) == [' '.join(c.split(' ')[:Tc - 1]) for c in caps]
    # a reference scored against its own video beats the same sentence scored against the other 63 videos (CIDEr-D is a
    # consensus score; ties only if two videos share the sentence)
    own = scorer.score_ids(torch.from_numpy(ids).cuda(), torch.arange(B, dtype=torch.int32)).cpu().numpy()
    other = scorer.score_ids(torch.from_numpy(ids).cuda(), ((torch.arange(B) + 1) % B).to(torch.int32)).cpu().numpy()
    assert (own > 0).all() and (own >= other).mean() > 0.95
