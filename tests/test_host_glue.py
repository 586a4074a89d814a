"""CPU tests of the host-side mirror: text glue against the oracle restatement, feature-file parsing, batching rule,
learning-rate schedule and the world_size-2 gradient all-reduce plumbing (gloo)."""
import os

import numpy as np
import pytest
import torch

import s2vt_b200
from oracle import text as otext

G = os.path.join(os.path.dirname(__file__), 'golden')
T = s2vt_b200.text


@pytest.fixture(scope='module')
def msvd():
    sents = T.read_sentences(os.path.join(G, 'msvd_sents_train_noval_lc_nopunc.txt.gz'))
    vocab = T.read_vocabulary(os.path.join(G, 'msvd_vocabulary1.txt.gz'))
    return sents, vocab


def test_vocab_matches_oracle(msvd):
    sents, vocab = msvd
    w2i, i2w = T.preProBuildWordVocab(vocab)
    ow2i, oi2w = otext.build_word_vocab(vocab)
    assert w2i == ow2i and i2w == oi2w and len(w2i) == 9972


def test_sentence_padding_toix_matches_oracle(msvd):
    sents, vocab = msvd
    w2i, _ = T.preProBuildWordVocab(vocab)
    batch = [s for _, s in sents[:300]] + [' '.join(['a'] * 34), ' '.join(['man'] * 35), ' '.join(['is'] * 50), 'a  double space', 'Unknownword here', '']
    ids, mask = T.sentence_padding_toix(list(batch), w2i, 35)
    oids, omask = otext.sentence_padding_toix(list(batch), w2i, 35)
    assert ids.dtype == np.int32 and ids.shape == (len(batch), 35)
    np.testing.assert_array_equal(ids, np.array(oids))
    np.testing.assert_array_equal(mask, omask)


def test_decode_captions_and_masks_match_oracle(msvd):
    sents, vocab = msvd
    _, i2w = T.preProBuildWordVocab(vocab)
    rng = np.random.RandomState(0)
    caps = rng.randint(0, 60, size=(200, 35))
    caps[::7, 0] = 0
    caps[1::7] = np.where(caps[1::7] == 0, 5, caps[1::7])      # rows without any <eos>
    masks, dec = T.decode_captions_masks(caps, i2w)
    omasks, odec = otext.decode_captions_masks(caps, i2w)
    assert masks == omasks and dec == odec
    assert T.decode_captions(caps, i2w) == otext.decode_captions(caps, i2w)
    assert T.decode_captions(caps[3], i2w) == otext.decode_captions(caps[3], i2w)


def test_feature_file_round_trip(tmp_path):
    rng = np.random.RandomState(1)
    feats = {'vid%d' % v: rng.rand(3, 8).astype(np.float32) for v in (1, 2, 10)}
    p = tmp_path / 'feats.txt'
    with open(p, 'w') as f:                                      # writer format of tf_feature_extract.py:153-154
        for v, a in feats.items():
            for k, row in enumerate(a):
                f.write('%s_frame_%d,' % (v, k) + ','.join(repr(float(x)) for x in row) + '\n')
    got = T.read_features(str(p))
    assert set(got) == set(feats)
    for v in feats:
        np.testing.assert_allclose(got[v], feats[v], rtol=1e-6)
    with open(p, 'a') as f:
        f.write('vid99_frame_0,' + ','.join(['0.5'] * 8) + '\n')
    with pytest.raises(AssertionError):                          # ragged frame counts are rejected (tf_s2vt.py:342)
        T.read_features(str(p))


def test_get_captions_and_grouping(msvd):
    sents, _ = msvd
    by, order = T.group_by_video(sents)
    assert order[0] == 'vid1' and len(order) == 1200
    assert T.get_captions(sents, 'vid7') == by['vid7'] == otext.get_captions([tuple(x) for x in sents], 'vid7')


def test_batching_and_schedule():
    from s2vt_b200 import cli, trainer
    assert cli._batches(10, 3, 0, 1) == [0, 3, 6]               # zip(range(0, n-bs, bs), ...) drops the tail (Q6)
    assert cli._batches(9, 3, 0, 1) == [0, 3]
    assert cli._batches(20, 3, 1, 2) == [3, 9, 15]
    # a batch count that does not divide by the world size: every rank gets the same number of batches (each one issues collectives)
    # and together they are a prefix of the single-process batch list
    for n, bs, world in ((10, 3, 2), (48774, 256, 8), (48774, 64, 8), (100, 7, 3)):
        parts = [cli._batches(n, bs, r, world) for r in range(world)]
        assert len({len(x) for x in parts}) == 1
        merged = sorted(sum(parts, []))
        assert merged == cli._batches(n, bs, 0, 1)[:len(merged)] and len(cli._batches(n, bs, 0, 1)) - len(merged) < world
    assert trainer.exponential_decay(1e-6, 999, 1000) == 1e-6 and trainer.exponential_decay(1e-6, 3000, 1000) == 1.25e-7
    # gradient ranges left for the final all-reduce once the early-final segments (s2vt_grad_segment_ready) have been handed out
    assert trainer.complement_ranges([], 10) == [(0, 10)]
    assert trainer.complement_ranges([(7, 10), (0, 2), (4, 5)], 12) == [(2, 4), (5, 7), (10, 12)]
    assert trainer.complement_ranges([(0, 12)], 12) == []


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)

    class Fake(object):
        pass
    m = Fake()
    n = 1000
    m.grads = torch.arange(n + 8, dtype=torch.float32) * (rank + 1)
    m.grads[n + 2] = 10.0 * (rank + 1)                           # aux slot: local sum(mask)
    from s2vt_b200 import trainer
    assert trainer.connect_peers(m) is False                      # host memory cannot be mapped by peers: the NCCL / gloo all-reduce stays
    trainer.allreduce_gradients(m, bucket_bytes=1024)            # several buckets
    if rank == 0:
        torch.save(m.grads, out)
    dist.destroy_process_group()


def test_gradient_allreduce_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / 'g.pt')
    mp.spawn(_dp_worker, args=(2, 29611, out), nprocs=2, join=True)
    g = torch.load(out)
    exp = torch.arange(1008, dtype=torch.float32) * 3
    exp[1002] = 30.0
    assert torch.equal(g, exp)


def _uneven_worker(rank, world, port, out):
    """The CLI's epoch loop shape: one all-reduce per batch, batch list from cli._batches with a count that does not divide evenly."""
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from s2vt_b200 import cli
    total = torch.zeros(1)
    for epoch in range(3):
        for start in cli._batches(23, 3, rank, world):            # 7 batches for 2 ranks
            t = torch.tensor([float(start)])
            dist.all_reduce(t)
            total += t
    dist.barrier()
    if rank == 0:
        torch.save(total, out)
    dist.destroy_process_group()


def test_uneven_batch_count_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / 't.pt')
    mp.spawn(_uneven_worker, args=(2, 29613, out), nprocs=2, join=True)     # would dead-lock if the ranks ran 4 and 3 iterations
    assert torch.load(out).item() == 3 * (0 + 3 + 6 + 9 + 12 + 15)


class _FakeModel(object):
    """Host stand-in with the attributes checkpoint.save / optimistic_restore touch."""

    def __init__(self, seed):
        g = torch.Generator().manual_seed(seed)
        self.variables = {'Wemb': (0, (6, 4)), 'encode_image_b': (24, (4,))}
        self.params = torch.randn(28, generator=g)
        self.adam_m = torch.randn(28, generator=g); self.adam_v = torch.rand(28, generator=g)
        self.adam_step = 0

    def state_dict(self):
        return {k: self.params[o:o + int(np.prod(s))].view(*s).numpy().copy() for k, (o, s) in self.variables.items()}

    def load_variables(self, named):
        done = []
        for k, (o, s) in self.variables.items():
            if k in named and tuple(named[k].shape) == tuple(s):
                self.params[o:o + int(np.prod(s))] = torch.from_numpy(np.asarray(named[k], np.float32).reshape(-1)); done.append(k)
        return done


def test_restore_brings_adam_state_and_matching_step_counter(tmp_path):
    """optimistic_restore runs over tf.global_variables() (reinforcement_multisampling_tf_s2vt.py:47-61): Adam slots and beta powers
    follow the variables; the step counter only when its name matches (Variable vs g_step)."""
    from s2vt_b200 import checkpoint as ck
    assert ck.adam_step_from_beta1_power(0.9) == 0                      # fresh optimiser: beta1_power initialised to beta1
    assert ck.adam_step_from_beta1_power(0.9 ** 8) == 7                 # 7 applies -> beta1^(7+1)
    a, b = _FakeModel(1), _FakeModel(2)
    a.adam_step = 7
    path = ck.save(a, str(tmp_path / 'xe-3'), global_step=1234, step_name='Variable')
    restored, step = ck.optimistic_restore(b, path, with_optimizer=True, step_name='g_step')     # stage 2 restoring a stage-1 file
    assert sorted(restored) == ['Wemb', 'encode_image_b'] and step == 0
    assert b.adam_step == 7 and torch.equal(b.adam_m, a.adam_m) and torch.equal(b.adam_v, a.adam_v) and torch.equal(b.params, a.params)
    c = _FakeModel(3)
    _, step = ck.optimistic_restore(c, path, with_optimizer=True, step_name='Variable')          # stage 1 resuming itself
    assert step == 1234
    # a TF-written file carries no `adam_step`: it comes from beta1_power
    d = dict(np.load(path)); del d['adam_step']
    np.savez(str(tmp_path / 'tf_like.npz'), **d)
    e = _FakeModel(4)
    ck.optimistic_restore(e, str(tmp_path / 'tf_like.npz'), with_optimizer=True)
    assert e.adam_step == 7


class _FakeRLModel(object):
    """Records what ReinforceTrainer.step asks of the library (host logic only; CPU tensors)."""

    def __init__(self, rank):
        self.device = torch.device('cpu')
        self.grads = torch.zeros(16 + 8)
        self.calls = []
        self.rank = rank

    def set_reuse_frontend(self, enable=True):
        self.calls.append(('reuse', enable))

    def rollout(self, v, K, seed, row_base=0):
        self.calls.append(('rollout', K, seed, row_base))
        B = v.shape[0]
        return torch.full((K * B, 35), 3, dtype=torch.int32), torch.full((B, 35), 4, dtype=torch.int32)

    def caption_masks(self, ids):
        return torch.ones(ids.shape, dtype=torch.float32), None

    def rl_backward(self, v, samp, mask, r, b, norm=0.0, drop_seed=0, row_base=0):
        self.calls.append(('backward', norm, drop_seed, row_base, float(r.sum()), float(b.sum())))
        self.grads[:16] = float(self.rank + 1)
        self.grads[16 + 2] = float(mask.sum())              # aux slot: local sum(mask)

    def optimizer_step(self, lr, clip, wemb_slice_norm=True, normalize=False):
        self.calls.append(('adam', lr, clip, wemb_slice_norm, normalize, self.grads.clone()))
        return torch.zeros(2)


class _FakeScorer(object):
    def score_ids(self, ids, vi):
        return torch.arange(ids.shape[0], dtype=torch.float64) * 0.01


def _trainer_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from s2vt_b200 import trainer
    m = _FakeRLModel(rank)
    tr = trainer.ReinforceTrainer(m, _FakeScorer(), n_samples=3, start_learning_rate=1e-3, decay_steps=2, clip_norm=5.0, seed=7)
    assert tr.peer_exchange is False                             # host tensors: the collective path
    for _ in range(3):
        tr.step(np.zeros((2, 4, 8), np.float32), np.array([0, 1], np.int32))
    if rank == 1:
        torch.save(m.calls, out)
    dist.barrier()
    dist.destroy_process_group()


def test_reinforce_trainer_step_host_logic_world_size_2_gloo(tmp_path):
    """Rank 1 of 2: Philox row base = rank * K * B, per-iteration seeds, norm deferred to the optimiser (normalize=True), gradients and the
    sum(mask) aux slot summed over the ranks before the optimiser step, staircase learning-rate decay."""
    import torch.multiprocessing as mp
    out = str(tmp_path / 'calls.pt')
    mp.spawn(_trainer_worker, args=(2, 29617, out), nprocs=2, join=True)
    calls = torch.load(out, weights_only=False)
    assert calls[0] == ('reuse', True)
    K, B = 3, 2
    steps = [calls[1 + 3 * i: 4 + 3 * i] for i in range(3)]
    for it, (ro, bw, ad) in enumerate(steps):
        assert ro == ('rollout', K, 7 + it, 1 * K * B)
        assert bw[:4] == ('backward', 1.0, 7 * 7919 + it + 1, 1 * K * B)
        assert abs(bw[4] - sum(0.01 * j for j in range(K * B))) < 1e-6                       # rewards: scores of the K * B sampled rows
        assert abs(bw[5] - K * sum(0.01 * (K * B + j) for j in range(B))) < 1e-6             # baseline: the greedy scores, repeated K times
        assert ad[0] == 'adam' and abs(ad[1] - 1e-3 * 0.5 ** (it // 2)) < 1e-12 and ad[2] == 5.0 and ad[3] is True and ad[4] is True
        g = ad[5]
        assert torch.equal(g[:16], torch.full((16,), 3.0))                                  # 1 + 2 over the two ranks
        assert g[18].item() == 2 * K * B * 35                                                # global sum(mask)


class _FakeXEModel(object):
    def __init__(self, rank):
        self.device = torch.device('cpu')
        self.grads = torch.zeros(8 + 8)
        self.rank, self.calls = rank, []

    def _f32(self, x):
        return torch.as_tensor(x, dtype=torch.float32)

    def xe_backward_sharded(self, video, captions, mask, colsum, n_global, norm, decay=None, drop_seed=0, row_base=0):
        self.calls.append(('xe', colsum.clone(), n_global, norm, decay, drop_seed, row_base))
        self.grads[:8] = float(10 * (self.rank + 1))
        return torch.tensor([1.0 + self.rank, 0.5 if decay is None else 0.0])

    def optimizer_step(self, lr, clip, wemb_slice_norm=True, normalize=False):
        self.calls.append(('adam', lr, clip, wemb_slice_norm, normalize, self.grads.clone()))


def _xe_trainer_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from s2vt_b200 import trainer
    m = _FakeXEModel(rank)
    tr = trainer.XETrainer(m, start_learning_rate=1e-3, decay_steps=5000, clip_norm=10.0, seed=3)
    n = 2 + rank                                                  # ragged shards: 2 rows on rank 0, 3 on rank 1
    mask = np.zeros((n, 6), np.float32); mask[:, :2 + rank] = 1
    loss = tr.step(np.zeros((n, 4, 8), np.float32), np.zeros((n, 6), np.int32), mask)
    torch.save((m.calls, loss), out + str(rank))
    dist.barrier()
    dist.destroy_process_group()


def test_xe_trainer_step_host_logic_world_size_2_gloo(tmp_path):
    """Q3 couples the rows of a batch: the ranks exchange the per-step mask sums and the row count FIRST, back-propagate with the global
    statistics (weight decay on rank 0 only), then sum gradients and loss parts."""
    import torch.multiprocessing as mp
    out = str(tmp_path / 'xe')
    mp.spawn(_xe_trainer_worker, args=(2, 29619, out), nprocs=2, join=True)
    for rank in range(2):
        calls, loss = torch.load(out + str(rank), weights_only=False)
        xe, ad = calls
        colsum = torch.tensor([5., 5., 3., 0., 0., 0.])           # 2 rows with 2 ones + 3 rows with 3 ones
        assert xe[0] == 'xe' and torch.equal(xe[1], colsum) and xe[2] == 5 and xe[3] == 13.0
        assert xe[4] == (None if rank == 0 else 0.0)
        assert xe[5] == 3 * 7919 + 1 and xe[6] == rank * (2 + rank)
        assert ad[0] == 'adam' and ad[1] == 1e-3 and ad[2] == 10.0 and ad[3] is False
        assert torch.equal(ad[5][:8], torch.full((8,), 30.0))
        assert torch.equal(loss, torch.tensor([3.0, 0.5]))        # CE parts add up, the decay part comes from rank 0 alone
