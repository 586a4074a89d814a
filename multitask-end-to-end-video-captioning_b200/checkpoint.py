"""Checkpoint save / restore keyed by the reference's TF variable names.

The reference uses tf.train.Saver (tf_s2vt.py:440, 560) and `optimistic_restore` (reinforcement_multisampling_tf_s2vt.py:
47-61: restore every variable whose name AND shape match, skip the rest silently).  Here a checkpoint is an .npz whose
keys are those TF names (plus `<name>/Adam`, `<name>/Adam_1` slots and `global_step`); reading genuine TF bundles
without TensorFlow is the SURVEY "next" row N2.
"""
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


def optimistic_restore(model, path, with_optimizer=False):
    """Load what matches by name and shape; return (restored names, global_step)."""
    if not path.endswith('.npz') and os.path.exists(path + '.npz'):
        path = path + '.npz'
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
