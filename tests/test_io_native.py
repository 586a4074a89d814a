"""CPU tests of the host-side C ABI of include/s2vt_io.h: native feature-file ingest (N3) against the oracle's restatement
of the reference reader, and the TensorFlow checkpoint reader (N2) against files written by tests/tf_ckpt_writer.py."""
import gzip
import os
import sys
import time

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tf_ckpt_writer as W                                       # noqa: E402
from oracle import text as otext                                  # noqa: E402


@pytest.fixture(scope='module')
def pkg():
    import __graft_entry__ as ge
    ge.build()
    import s2vt_b200
    return s2vt_b200


def _write_features(path, n_videos, T, D, seed, fmt='repr', shuffle=False):
    rng = np.random.RandomState(seed)
    lines = []
    for v in range(n_videos):
        vals = np.maximum(0.0, rng.normal(0.25, 0.5, size=(T, D)))
        for k in range(T):
            if fmt == 'repr':
                fields = [repr(float(x)) for x in vals[k]]
            elif fmt == 'short':
                fields = ['%.6g' % x for x in vals[k]]
            else:                                                 # a mix of spellings float() accepts
                fields = []
                for j, x in enumerate(vals[k]):
                    fields.append(['%e' % x, ' %r' % float(x), '+%.3f' % x, '-%.10f' % x, '%d' % int(x * 10), '1e-42', '3.4028236e38', '1e39',
                                   '0.1', '16777217', '1_0.5', 'inf', '-Infinity', '.5', '5.'][(j + k) % 15])
            lines.append('vid%d_frame_%d,' % (v + 1, k) + ','.join(fields))
    if shuffle:                                                   # frames of different videos interleaved; per-video order kept
        order = np.random.RandomState(seed + 1).permutation(len(lines))
        lines = [lines[i] for i in sorted(order, key=lambda i: (i % T, i // T))]
    with open(path, 'w') as f:
        f.write('\n'.join(lines) + '\n')


@pytest.mark.parametrize('fmt,shuffle', [('repr', False), ('short', True), ('mixed', False)])
def test_feature_ingest_bit_exact(pkg, tmp_path, fmt, shuffle):
    p = str(tmp_path / 'f.txt')
    _write_features(p, 7, 5, 24, seed=3, fmt=fmt, shuffle=shuffle)
    want, order = otext.read_features(p)
    ff = pkg.ingest.FeatureFile(p, n_threads=3)
    assert ff.ids == order and (ff.n_videos, ff.n_frames, ff.dim) == (7, 5, 24)
    got = ff.to_dict()
    for v in order:
        assert got[v].dtype == np.float32
        np.testing.assert_array_equal(got[v].view(np.uint32), want[v].view(np.uint32))       # bit-exact, incl. inf / subnormals
    pick = ['vid3', 'vid1', 'vid3', 'vid7']
    b = ff.batch(pick)
    np.testing.assert_array_equal(b.view(np.uint32), np.stack([want[v] for v in pick]).view(np.uint32))
    np.testing.assert_array_equal(ff['vid2'].view(np.uint32), want['vid2'].view(np.uint32))
    with pytest.raises(KeyError):
        ff.batch(['vid999'])
    d = pkg.text.read_features(p)
    assert list(d) == order
    ff.close()


def test_feature_ingest_gz_no_trailing_newline_and_single_thread(pkg, tmp_path):
    p = str(tmp_path / 'f.txt')
    _write_features(p, 3, 2, 5, seed=9)
    raw = open(p, 'rb').read().rstrip(b'\n')
    gz = str(tmp_path / 'g.txt.gz')
    with gzip.open(gz, 'wb') as f:
        f.write(raw)
    want, order = otext.read_features(gz)
    got = pkg.ingest.FeatureFile(gz, n_threads=1).to_dict()
    assert list(got) == order
    for v in order:
        np.testing.assert_array_equal(got[v], want[v])


def test_feature_ingest_errors(pkg, tmp_path):
    p = str(tmp_path / 'f.txt')
    _write_features(p, 3, 2, 5, seed=9)
    base = open(p).read()
    with open(p, 'w') as f:
        f.write(base + 'vid9_frame_0,' + ','.join(['0.5'] * 5) + '\n')
    with pytest.raises(AssertionError):                           # ragged frame counts (tf_s2vt.py:342)
        pkg.text.read_features(p)
    with pytest.raises(AssertionError):
        otext.read_features(p)
    with open(p, 'w') as f:
        f.write(base.replace('\n', ',abc\n', 1))
    with pytest.raises(ValueError):                               # not a number / extra field: ValueError in the reference's feed
        pkg.text.read_features(p)
    with open(p, 'w') as f:
        lines = base.split('\n')
        lines[1] = ','.join(lines[1].split(',')[:-1])
        f.write('\n'.join(lines))
    with pytest.raises(ValueError):                               # a short line
        pkg.text.read_features(p)
    with pytest.raises(Exception):
        pkg.ingest.FeatureFile(str(tmp_path / 'missing.txt'))


def test_feature_ingest_is_faster_than_the_python_reader(pkg, tmp_path):
    p = str(tmp_path / 'big.txt')
    _write_features(p, 40, 20, 1536, seed=5)                      # 800 lines x 1536 floats, ~25 MB
    t0 = time.perf_counter(); want, order = otext.read_features(p); t_py = time.perf_counter() - t0
    t0 = time.perf_counter(); ff = pkg.ingest.FeatureFile(p); got = ff.read(np.arange(ff.n_videos)); t_nat = time.perf_counter() - t0
    np.testing.assert_array_equal(got, np.stack([want[v] for v in order]))
    print('feature ingest: python %.3fs, native %.3fs (%.0f MB/s)' % (t_py, t_nat, os.path.getsize(p) / 1e6 / t_nat))
    assert t_nat < t_py


# ---- TensorFlow checkpoints ----------------------------------------------------------------------------------------
def _tensors(seed=0):
    rng = np.random.RandomState(seed)
    t = {'Wemb': rng.randn(60, 12).astype(np.float32), 'encode_image_W': rng.randn(32, 12).astype(np.float32),
         'encode_image_b': rng.randn(12).astype(np.float32), 'embed_word_W': rng.randn(16, 60).astype(np.float32),
         'embed_word_b': rng.randn(60).astype(np.float32),
         's2vt/LSTM1/basic_lstm_cell/weights': rng.randn(28, 64).astype(np.float32), 's2vt/LSTM1/basic_lstm_cell/biases': rng.randn(64).astype(np.float32),
         's2vt/LSTM2/basic_lstm_cell/weights': rng.randn(44, 64).astype(np.float32), 's2vt/LSTM2/basic_lstm_cell/biases': rng.randn(64).astype(np.float32),
         'Variable': np.asarray(1234, dtype=np.int32), 'beta1_power': np.asarray(0.9 ** 7, dtype=np.float32),
         'beta2_power': np.asarray(0.999 ** 7, dtype=np.float32), 'some/double': rng.randn(3, 2, 2), 'some/int64': np.arange(-3, 4, dtype=np.int64)}
    for k in list(t):
        if t[k].ndim >= 1 and t[k].dtype == np.float32:
            t[k + '/Adam'] = rng.randn(*t[k].shape).astype(np.float32)
            t[k + '/Adam_1'] = (rng.rand(*t[k].shape) ** 2).astype(np.float32)
    return t


@pytest.mark.parametrize('fmt', [1, 2])
@pytest.mark.parametrize('block_size', [64, 4096, 1 << 20])
def test_tf_checkpoint_round_trip(pkg, tmp_path, fmt, block_size):
    t = _tensors()
    prefix = str(tmp_path / 's2vt_model-10')
    (W.write_v1 if fmt == 1 else W.write_v2)(prefix, t, block_size=block_size)
    r = pkg.checkpoint.TFCheckpointReader(prefix)
    assert r.format == fmt
    shapes = r.get_variable_to_shape_map()
    assert sorted(shapes) == sorted(t)
    for k, a in t.items():
        assert shapes[k] == list(a.shape)
        got = r.get_tensor(k)
        assert got.dtype == np.float32 and got.shape == a.shape
        np.testing.assert_array_equal(got, a.astype(np.float32))
    assert r.has_tensor('Wemb') and not r.has_tensor('nope')
    r.close()
    loaded = pkg.checkpoint.load_tf_checkpoint(prefix)
    assert set(loaded) == set(t)


def test_tf_checkpoint_corruption_is_detected(pkg, tmp_path):
    t = _tensors()
    prefix = str(tmp_path / 'm')
    W.write_v2(prefix, t)
    data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    data[100] ^= 0x40
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
    r = pkg.checkpoint.TFCheckpointReader(prefix)
    bad = 0
    for k in t:
        try:
            r.get_tensor(k)
        except pkg._lib.S2vtError as e:
            assert 'CRC32C' in str(e)
            bad += 1
    assert bad == 1
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[10] ^= 0x01
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(pkg._lib.S2vtError):
        pkg.checkpoint.TFCheckpointReader(prefix)
    with pytest.raises(pkg._lib.S2vtError):
        pkg.checkpoint.TFCheckpointReader(str(tmp_path / 'absent'))
    v1 = str(tmp_path / 'v1')
    W.write_v1(v1, t)
    raw = bytearray(open(v1, 'rb').read())
    raw[len(raw) // 2] ^= 0x10
    open(v1, 'wb').write(bytes(raw))
    with pytest.raises(pkg._lib.S2vtError):
        pkg.checkpoint.TFCheckpointReader(v1)


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors (the CRC the table / bundle formats use)
    assert W.crc32c(b'\x00' * 32) == 0x8A9136AA and W.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert W.crc32c(bytes(range(32))) == 0x46DD794E and W.crc32c(b'123456789') == 0xE3069283
