"""Checkpoint save / restore keyed by the reference's TF variable names.

The reference uses tf.train.Saver (tf_s2vt.py:440, 560) and `optimistic_restore` (reinforcement_multisampling_tf_s2vt.py:
47-61: restore every variable whose name AND shape match, skip the rest silently).  Here a checkpoint is an .npz whose
keys are those TF names (plus `<name>/Adam`, `<name>/Adam_1` slots and `global_step`); genuine TensorFlow checkpoints
(V1 single-file tables as written by the reference's `Saver(write_version=1)`, and V2 .index/.data bundles) are read
without TensorFlow by the native reader of include/s2vt_io.h (csrc/tfckpt.cpp).
"""
import ctypes as C
import os

import numpy as np

IGNORED_PREFIXES = ('InceptionResnetV2/',)
IGNORED_NAMES = ('Variable', 'g_step', 'beta1_power', 'beta2_power')


def save(model, path, global_step=0, with_optimizer=True):
    out = dict(model.state_dict())
    if with_optimizer:
        for name, (off, shp) in model.variables.items():
            n = int(np.prod(shp))
            out[name + '/Adam'] = model.adam_m[off:off + n].view(*shp).cpu().numpy()
            out[name + '/Adam_1'] = model.adam_v[off:off + n].view(*shp).cpu().numpy()
        out['adam_step'] = np.asarray(model.adam_step, dtype=np.int64)
    out['global_step'] = np.asarray(global_step, dtype=np.int64)
    os.makedirs(os.path.dirname(os.path.abspath(path)) or '.', exist_ok=True)
    np.savez(path, **out)
    return path if path.endswith('.npz') else path + '.npz'


class _Arrays(object):
    """np.load-like view ({name: array}, .files) of a dict."""

    def __init__(self, d):
        self._d, self.files = d, list(d)

    def __getitem__(self, k):
        return self._d[k]


class TFCheckpointReader(object):
    """tf.train.NewCheckpointReader stand-in: get_variable_to_shape_map(), get_tensor(name), has_tensor(name)."""

    def __init__(self, prefix):
        from . import _lib
        self._lib, self.lib = _lib, _lib.load()
        h = C.c_void_p()
        _lib.check_io(self.lib.s2vt_ckpt_open(str(prefix).encode(), C.byref(h)))
        self.h = h
        self.format = int(self.lib.s2vt_ckpt_format(h))
        self._info = {}
        for i in range(self.lib.s2vt_ckpt_num_tensors(h)):
            name, dt, nd, dims = C.c_char_p(), C.c_int32(), C.c_int32(), (C.c_int64 * 8)()
            _lib.check_io(self.lib.s2vt_ckpt_tensor_info(h, i, C.byref(name), C.byref(dt), C.byref(nd), C.byref(dims)))
            self._info[name.value.decode()] = (i, dt.value, tuple(dims[k] for k in range(nd.value)))

    def close(self):
        if getattr(self, 'h', None):
            self.lib.s2vt_ckpt_close(self.h)
            self.h = None

    __del__ = close

    def get_variable_to_shape_map(self):
        return {k: list(v[2]) for k, v in self._info.items()}

    def has_tensor(self, name):
        return name in self._info

    def get_tensor(self, name):
        """Tensor as float32 (integer / double variables such as global_step are converted)."""
        i, dt, shape = self._info[name]
        out = np.empty(int(np.prod(shape)) if shape else 1, dtype=np.float32)
        self._lib.check_io(self.lib.s2vt_ckpt_read_f32(self.h, i, out.ctypes.data_as(C.c_void_p), out.size))
        return out.reshape(shape)


def is_tf_checkpoint(path):
    return os.path.exists(path + '.index') or (os.path.isfile(path) and not path.endswith('.npz'))


def load_tf_checkpoint(prefix, skip_unreadable=True):
    """{name: float32 array} of every dense variable in a TF V1 / V2 checkpoint (string tensors etc. are skipped)."""
    from . import _lib
    r = TFCheckpointReader(prefix)
    out = {}
    try:
        for name in r.get_variable_to_shape_map():
            try:
                out[name] = r.get_tensor(name)
            except _lib.S2vtError:
                if not skip_unreadable:
                    raise
    finally:
        r.close()
    return out


def optimistic_restore(model, path, with_optimizer=False):
    """Load what matches by name and shape; return (restored names, global_step).  `path` is an .npz written by save()
    or a TensorFlow checkpoint prefix (what the reference hands to saver.restore)."""
    if not path.endswith('.npz') and os.path.exists(path + '.npz'):
        path = path + '.npz'
    if is_tf_checkpoint(path):
        tf_vars = load_tf_checkpoint(path)
        for alias in ('Variable', 'g_step'):                    # the reference's global_step variables (tf_s2vt.py:441)
            if alias in tf_vars and 'global_step' not in tf_vars:
                tf_vars['global_step'] = tf_vars[alias].astype(np.int64).reshape(())
        if 'beta1_power' in tf_vars and 0.0 < float(tf_vars['beta1_power'].reshape(-1)[0]) < 1.0:
            tf_vars['adam_step'] = np.asarray(int(round(np.log(float(tf_vars['beta1_power'].reshape(-1)[0])) / np.log(0.9))), dtype=np.int64)
        data = _Arrays(tf_vars)
    else:
        data = np.load(path)
    named = {k: data[k] for k in data.files
             if not k.endswith('/Adam') and not k.endswith('/Adam_1') and k not in IGNORED_NAMES + ('global_step', 'adam_step')
             and not k.startswith(IGNORED_PREFIXES)}
    restored = model.load_variables(named)
    if with_optimizer and 'adam_step' in data.files:
        import torch
        for name in restored:
            off, shp = model.variables[name]
            n = int(np.prod(shp))
            if name + '/Adam' in data.files and data[name + '/Adam'].shape == tuple(shp):
                model.adam_m[off:off + n].copy_(torch.from_numpy(data[name + '/Adam'].reshape(-1)))
                model.adam_v[off:off + n].copy_(torch.from_numpy(data[name + '/Adam_1'].reshape(-1)))
        model.adam_step = int(data['adam_step'])
    step = int(data['global_step']) if 'global_step' in data.files else 0
    return restored, step
