"""The bf16 kernel variants must agree with each other: the tcgen05 paths (persistent chains, weights-stationary chains,
per-step launches, split-K clusters, MN-major weight gradients, 128/256-wide tiles, TMA multicast) against the warp-level
mma.sync checker path, on the same inputs.  Differences are only fp32 summation order."""
import numpy as np
import pytest
import torch

from oracle import s2vt_numpy as M

pytestmark = pytest.mark.gpu

DIMS = dict(D=200, E=60, H=72, V=301)
Tv, Tc, B, K = 3, 7, 4, 2


def _run(backend, dims=DIMS, tv=Tv, tc=Tc, b=B, k=K, keep=0.9, overlap=None):
    import s2vt_b200
    p = M.init_params(seed=4, dtype=np.float32, **dims)
    m = s2vt_b200.Video_Caption_Generator(dim_image=dims['D'], n_words=dims['V'], word_dim=dims['E'], lstm_dim=dims['H'], batch_size=b,
                                          n_video_lstm_step=tv, n_caption_lstm_step=tc, dropout_rate=keep, precision='bf16',
                                          gemm_backend=backend, max_videos=b, max_rows=k * b)
    m.load_variables(p)
    if overlap is not None:
        m.lib.s2vt_set_overlap(m.h, overlap)
    video = M.synthetic_features(b, tv, dims['D'])
    samp, greedy = m.rollout(video, k, seed=5)
    mask, _ = m.caption_masks(samp)
    rng = np.random.RandomState(1)
    r = torch.tensor(rng.uniform(0, 2, k * b), dtype=torch.float32); base = torch.tensor(rng.uniform(0, 2, k * b), dtype=torch.float32)
    logp, logits = m.teacher_forward(video, samp, drop_seed=9, want_logits=True)
    loss = m.rl_backward(video, samp, mask, r, base, drop_seed=9).item()
    grads = m.grads[:m.n_params].clone()
    w2 = m.variable('s2vt/LSTM2/basic_lstm_cell/weights', grad=True)
    w2_big = torch.cat([w2[:dims['H']], w2[dims['H'] + dims['E']:]]).clone().cpu()      # the [out1 rows] and [h rows] blocks: plain (non-atomic) sums
    m.optimizer_step(1e-3, 5.0)
    return dict(samp=samp.cpu(), greedy=greedy.cpu(), logits=logits.cpu(), loss=loss, grads=grads.cpu(), params=m.params.cpu().clone(), w2_big=w2_big)


@pytest.fixture(scope='module')
def reference_run():
    return _run('mma_sync')


@pytest.mark.parametrize('backend', ['auto', 'tcgen05_n128', 'tcgen05_mc2x2', 'step_mc8', 'step_n64', 'wgrad_transposed', 'per_step', 'chain_ring', 'chain_mc4',
                                     'chain_nomc'])
def test_tcgen05_variants_match_mma_sync(backend, reference_run):
    ref, got = reference_run, _run(backend)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert torch.equal(got['greedy'], ref['greedy'])
    same = (got['samp'] == ref['samp']).float().mean().item()
    assert same > 0.9, same
    if same == 1.0:      # identical sampled captions -> the training pass saw identical inputs
        assert rel(got['logits'], ref['logits']) < 2e-3
        assert abs(got['loss'] - ref['loss']) < 2e-3 * max(1.0, abs(ref['loss']))
        assert rel(got['grads'], ref['grads']) < 2e-2
        # one clipped Adam step moves every element by about +-lr (1e-3) whatever its gradient's size: an element whose gradient is at the
        # rounding noise of the two mainloops may move the other way (2 lr); anything systematic would show up everywhere
        assert rel(got['params'], ref['params']) < 5e-3
        assert float(((got['params'] - ref['params']).abs() > 1e-4).float().mean()) < 1e-3


def test_full_size_multicast_chain_equals_plain_chain():
    """64 rows, H=1000: weights-stationary chains with the activation rows multicast over clusters of 4 vs without."""
    dims = dict(D=1536, E=500, H=1000, V=9972)
    a = _run('chain_mc4', dims, 5, 35, 64, 2)
    b = _run('chain_nomc', dims, 5, 35, 64, 2)
    assert torch.equal(a['greedy'], b['greedy']) and torch.equal(a['samp'], b['samp']) and torch.equal(a['logits'], b['logits'])


def test_full_size_chain_equals_per_step_launches():
    """B=64 rows, H=1000: persistent (weights-stationary + ring + split-K cluster) chains vs one launch per step."""
    dims = dict(D=1536, E=500, H=1000, V=9972)
    a = _run('auto', dims, 5, 35, 64, 2)
    b = _run('per_step', dims, 5, 35, 64, 2)
    assert torch.equal(a['greedy'], b['greedy']) and torch.equal(a['samp'], b['samp'])
    assert torch.equal(a['logits'], b['logits'])            # same tiles, same K order -> bit-identical forward
    rel = float((a['grads'] - b['grads']).abs().max() / b['grads'].abs().max())
    assert rel < 1e-5, rel


@pytest.mark.parametrize('b,k,tv', [(64, 5, 5), (64, 3, 5), (40, 4, 3), (64, 6, 2)])
def test_pipelined_weights_stationary_chain_equals_plain_chain(b, k, tv):
    """> 128 caption rows, H=1000: the teacher-forced LSTM2 chain on the pipelined weights-stationary kernel (gemm_tcgen05_ws2.cuh: 43 x 3
    CTAs, resident weight slabs, two M=64 halves per row group, per-half flags) against the plain ring chain (gemm_backend
    chain_plain).  One accumulator, K ascending in both -> bit-identical logits; row counts 320 / 192 / 160 / 384 exercise full,
    short and ragged row groups."""
    dims = dict(D=1536, E=500, H=1000, V=9972)
    a = _run('auto', dims, tv, 35, b, k)
    p = _run('chain_plain', dims, tv, 35, b, k)
    assert torch.equal(a['greedy'], p['greedy']) and torch.equal(a['samp'], p['samp'])
    assert torch.equal(a['logits'], p['logits'])
    rel = float((a['grads'] - p['grads']).abs().max() / p['grads'].abs().max())
    assert rel < 1e-5, rel
    assert abs(a['loss'] - p['loss']) < 1e-6 * max(1.0, abs(p['loss']))
    # the opt-in weights-stationary BPTT chain (partial tiles summed through L2 in K-slice order) gives the same gradients
    w = _run('chain_ws2_bwd', dims, tv, 35, b, k)
    assert torch.equal(w['logits'], p['logits'])
    rel = float((w['grads'] - p['grads']).abs().max() / p['grads'].abs().max())
    assert rel < 1e-5, rel


def test_cta_pair_gemms_equal_single_cta_gemms():
    """Large batched GEMMs (projections, G2x, logits, d logits . Wo^T, MN-major weight gradients) on 256 x 256 cta_group::2 pair tiles
    (gemm_tcgen05_pair.cuh) against the single-CTA 128 x 256 tiles (gemm_backend single_cta): one accumulator, K ascending in both."""
    dims = dict(D=1536, E=500, H=1000, V=9972)
    a = _run('pair_all', dims, 5, 35, 64, 3)
    p = _run('single_cta', dims, 5, 35, 64, 3)
    assert torch.equal(a['greedy'], p['greedy']) and torch.equal(a['samp'], p['samp'])
    assert torch.equal(a['logits'], p['logits'])
    rel = float((a['grads'] - p['grads']).abs().max() / p['grads'].abs().max())
    assert rel < 1e-5, rel


@pytest.mark.parametrize('b,k,tv', [(64, 5, 5), (40, 4, 11)])
def test_gated_consumption_under_the_bptt_chain_equals_the_sequential_order(b, k, tv):
    """> 128 caption rows: 40 % of dout1 and of the two large LSTM2 weight gradients (the time steps the LSTM2 BPTT chain finishes first) are computed on
    the side stream WHILE the chain runs, behind a watcher of its grid-barrier counter (s2vt_set_overlap bit 7, default on), against everything after the
    chain (mask 7).  Same products; the weight gradients are summed in two parts instead of one -> fp32 summation order only."""
    dims = dict(D=1536, E=500, H=1000, V=9972)
    a = _run('auto', dims, tv, 35, b, k, overlap=135)
    p = _run('auto', dims, tv, 35, b, k, overlap=7)
    assert torch.equal(a['samp'], p['samp']) and torch.equal(a['logits'], p['logits'])
    rel = float((a['grads'] - p['grads']).abs().max() / p['grads'].abs().max())
    assert rel < 1e-5, rel
    assert a['loss'] == p['loss']
    again = _run('auto', dims, tv, 35, b, k, overlap=135)
    # deterministic: the split point is fixed, never "whatever the chain had finished" (the rest of the block holds atomically summed gradients)
    assert torch.equal(a['w2_big'], again['w2_big'])
    assert float((a['grads'] - again['grads']).abs().max() / a['grads'].abs().max()) < 1e-6
