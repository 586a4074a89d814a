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
    assert trainer.exponential_decay(1e-6, 999, 1000) == 1e-6 and trainer.exponential_decay(1e-6, 3000, 1000) == 1.25e-7


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
