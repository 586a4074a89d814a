"""Feature-file ingest on the native reader (include/s2vt_io.h, csrc/ingest.cpp): the feature half of
get_video_feature_caption_pair (tf_s2vt.py:332-342) and the per-batch `[train_features[x] for x in vid]` (:487), with the
file mapped once and the numbers parsed by a thread pool straight into a pinned [n, T_v, D] float32 batch."""
import ctypes as C
import gzip

import numpy as np

from . import _lib


class FeatureFile(object):
    """Mapping-like view of a feature text file: `ff[vid]` -> float32 [T_v, D], `ff.batch(vids)` -> float32 [n, T_v, D].

    Iteration / `in` / `len` follow the reference's `features` dict (keys in order of first appearance)."""

    def __init__(self, path, n_threads=0):
        self.lib = _lib.load()
        self.path, self.n_threads = str(path), int(n_threads)
        h = C.c_void_p()
        self._text = None
        if self.path.endswith('.gz'):
            self._text = gzip.open(self.path, 'rb').read()       # kept alive: the handle indexes into it
            _lib.check_io(self.lib.s2vt_features_open_memory(self._text, len(self._text), self.n_threads, C.byref(h)))
        else:
            _lib.check_io(self.lib.s2vt_features_open(self.path.encode(), self.n_threads, C.byref(h)))
        self.h = h
        self.n_videos = int(self.lib.s2vt_features_num_videos(h))
        self.n_frames = int(self.lib.s2vt_features_num_frames(h))
        self.dim = int(self.lib.s2vt_features_dim(h))
        self.ids = [self.lib.s2vt_features_video_id(h, i).decode() for i in range(self.n_videos)]
        self._index = {v: i for i, v in enumerate(self.ids)}

    def close(self):
        if getattr(self, 'h', None):
            self.lib.s2vt_features_close(self.h)
            self.h = None

    __del__ = close

    def __len__(self):
        return self.n_videos

    def __iter__(self):
        return iter(self.ids)

    def __contains__(self, vid):
        return vid in self._index

    def keys(self):
        return list(self.ids)

    def index_of(self, vids):
        """Video ids -> int64 indices; KeyError for an unknown id, as `train_features[x]` (tf_s2vt.py:487)."""
        return np.fromiter((self._index[v] for v in vids), dtype=np.int64)

    def read(self, index, out=None):
        """Parse the videos `index` (int64 [n]) into `out` (float32 [n, T_v, D]; a pinned torch tensor or an ndarray)."""
        idx = np.ascontiguousarray(index, dtype=np.int64)
        n = idx.shape[0]
        if out is None:
            out = np.empty((n, self.n_frames, self.dim), dtype=np.float32)
        if hasattr(out, 'data_ptr'):                             # torch tensor (pinned host memory)
            assert tuple(out.shape) == (n, self.n_frames, self.dim) and out.is_contiguous() and out.device.type == 'cpu'
            assert str(out.dtype) == 'torch.float32'
            ptr = out.data_ptr()
        else:
            assert out.shape == (n, self.n_frames, self.dim) and out.dtype == np.float32 and out.flags['C_CONTIGUOUS']
            ptr = out.ctypes.data
        _lib.check_io(self.lib.s2vt_features_read(self.h, idx.ctypes.data_as(C.c_void_p), n, C.c_void_p(ptr), self.n_threads))
        return out

    def batch(self, vids, out=None):
        return self.read(self.index_of(vids), out)

    def __getitem__(self, vid):
        return self.read(self.index_of([vid]))[0]

    def to_dict(self):
        """{vid: float32 [T_v, D]} of the whole file (what get_video_feature_caption_pair returns), parsed in one call."""
        all_ = self.read(np.arange(self.n_videos, dtype=np.int64))
        return {v: all_[i] for i, v in enumerate(self.ids)}

    def cache_fp16(self, pin=True):
        """The whole file as ONE float16 tensor [n_videos, T_v, D] in (pinned) host memory -- the feed cache for `trainer.FeaturePipe`: the tensor-core
        mode consumes the frames in fp16, so batches gathered from this cache give bit-identical results at half the host -> device traffic.
        Values beyond fp16's range (|x| > 65504) would become inf: pooled CNN features are O(1); checked here."""
        import torch
        all_ = self.read(np.arange(self.n_videos, dtype=np.int64))
        if not np.isfinite(all_).all() or np.abs(all_).max() > 65504.0:
            raise ValueError('features outside the fp16 range: keep the float32 feed')
        t = torch.from_numpy(all_.astype(np.float16))
        return t.pin_memory() if pin and torch.cuda.is_available() else t
