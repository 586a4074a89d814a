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


def save(model, path, global_step=0, with_optimizer=True, step_name='global_step'):
    """step_name: the TF name of the script's step counter -- `Variable` (tf_s2vt.py:441, an unnamed tf.Variable) or `g_step`
    (reinforcement_multisampling_tf_s2vt.py:638) -- so that, as in the reference, a restore picks the counter up only when the
    checkpoint was written by the same kind of script."""
    out = dict(model.state_dict())
    if with_optimizer:
        for name, (off, shp) in model.variables.items():
            n = int(np.prod(shp))
            out[name + '/Adam'] = model.adam_m[off:off + n].view(*shp).cpu().numpy()
            out[name + '/Adam_1'] = model.adam_v[off:off + n].view(*shp).cpu().numpy()
        out['adam_step'] = np.asarray(model.adam_step, dtype=np.int64)
        # TF's own bookkeeping: beta powers start at beta and are multiplied after every apply -> beta^(t+1) after t updates
        out['beta1_power'] = np.asarray(0.9 ** (model.adam_step + 1), dtype=np.float32)
        out['beta2_power'] = np.asarray(0.999 ** (model.adam_step + 1), dtype=np.float32)
    out[step_name] = np.asarray(global_step, dtype=np.int64)
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


def adam_step_from_beta1_power(beta1_power, beta1=0.9):
    """Number of Adam updates already applied.  TF initialises beta1_power to beta1 and multiplies it by beta1 AFTER every apply
    (AdamOptimizer._finish), so after t updates it holds beta1^(t+1)."""
    b = float(np.asarray(beta1_power).reshape(-1)[0])
    if not (0.0 < b < 1.0):
        return 0
    return max(0, int(round(np.log(b) / np.log(beta1))) - 1)


def optimistic_restore(model, path, with_optimizer=False, step_name=None):
    """Load what matches by name and shape; return (restored names, global_step).  `path` is an .npz written by save()
    or a TensorFlow checkpoint prefix (what the reference hands to saver.restore).

    The reference restores over tf.global_variables() (:47-61), which besides the model variables holds the optimiser's
    `<var>/Adam`, `<var>/Adam_1` slots and `beta1_power` / `beta2_power`: with_optimizer=True restores those too (slots by name and
    shape, the Adam time step from beta1_power).  The step counter is a global variable as well but its NAME differs between the
    scripts (`Variable` in tf_s2vt.py:441, `g_step` in the RL scripts :638), so it is restored only under `step_name`
    (None: any of `global_step`, `Variable`, `g_step`)."""
    if not path.endswith('.npz') and os.path.exists(path + '.npz'):
        path = path + '.npz'
    if is_tf_checkpoint(path):
        data = _Arrays(load_tf_checkpoint(path))
    else:
        data = np.load(path)
    skip = IGNORED_NAMES + ('global_step', 'adam_step')
    named = {k: data[k] for k in data.files
             if not k.endswith('/Adam') and not k.endswith('/Adam_1') and k not in skip and not k.startswith(IGNORED_PREFIXES)}
    restored = model.load_variables(named)
    if with_optimizer:
        import torch
        for name in restored:
            off, shp = model.variables[name]
            n = int(np.prod(shp))
            for slot, dst in (('/Adam', model.adam_m), ('/Adam_1', model.adam_v)):      # each slot is a variable of its own: name + shape rule
                if name + slot in data.files and tuple(data[name + slot].shape) == tuple(shp):
                    dst[off:off + n].copy_(torch.from_numpy(np.ascontiguousarray(data[name + slot], dtype=np.float32).reshape(-1)))
        if 'adam_step' in data.files:
            model.adam_step = int(data['adam_step'])
        elif 'beta1_power' in data.files:
            model.adam_step = adam_step_from_beta1_power(data['beta1_power'])
    step = 0
    for k in ((step_name,) if step_name else ('global_step', 'Variable', 'g_step')):
        if k in data.files and np.asarray(data[k]).size == 1:
            step = int(np.asarray(data[k]).reshape(-1)[0])
            break
    return restored, step
