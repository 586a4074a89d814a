"""Host-side mirror of `class Video_Caption_Generator` of original_attention.py:54-251 (temporal-attention decoder,
SURVEY 8(f) N1 / BASELINE config 3) over the C ABI `s2vt_att_*` of include/s2vt.h.  Same constructor arguments as the
reference; `build_sampler` / `build_generator` / `build_model` are calls instead of graph builders."""
import ctypes as C

import numpy as np
import torch

from . import _lib

LSTM3_W = 's2vt/LSTM3/basic_lstm_cell/weights'
LSTM3_B = 's2vt/LSTM3/basic_lstm_cell/biases'


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Video_Caption_Generator(object):
    def __init__(self, dim_image=1536, n_words=9972, dim_hidden=1000, batch_size=64, n_video_lstm_steps=5, n_caption_lstm_steps=35, drop_out_rate=0.9,
                 bias_init_vector=None, beta=10.0, m=0.5, precision='bf16', device=None, seed=16, train=True):
        if not torch.cuda.is_available():
            raise RuntimeError('multitask-end-to-end-video-captioning_b200 needs a CUDA device (sm_100a); there is no CPU path')
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self.dim_image, self.n_words, self.dim_hidden, self.batch_size = dim_image, n_words, dim_hidden, batch_size
        self.n_video_lstm_steps, self.n_caption_lstm_steps, self.drop_out_rate = n_video_lstm_steps, n_caption_lstm_steps, drop_out_rate
        cfg = _lib.S2vtAttConfig(dim_image, dim_hidden, n_words, n_video_lstm_steps, n_caption_lstm_steps,
                                 {'bf16': _lib.PREC_BF16, 'fp32': _lib.PREC_FP32}[precision], float(drop_out_rate), float(beta), float(m), 8)
        h = C.c_void_p()
        rc = self.lib.s2vt_att_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise _lib.S2vtError(rc, 's2vt_att_create rejected the configuration')
        self.h = h
        with torch.cuda.device(self.device):
            self._state = torch.zeros(self.lib.s2vt_att_state_bytes(h) + 256, dtype=torch.uint8, device=self.device)
            self._ws = torch.empty(self.lib.s2vt_att_workspace_bytes(h, batch_size, 1 if train else 0) + 256, dtype=torch.uint8, device=self.device)
            so, wo = (-self._state.data_ptr()) % 256, (-self._ws.data_ptr()) % 256
            self._check(self.lib.s2vt_att_bind(h, C.c_void_p(self._state.data_ptr() + so), self._state.numel() - so,
                                               C.c_void_p(self._ws.data_ptr() + wo), self._ws.numel() - wo))
        n = self.lib.s2vt_att_num_params(h)
        off = self.lib.s2vt_att_params(h) - self._state.data_ptr()
        self.params = self._state[off:off + 4 * n].view(torch.float32)
        self.n_params = n
        base = self._state.data_ptr()
        view = lambda ptr, cnt: self._state[ptr - base:ptr - base + 4 * cnt].view(torch.float32)
        self.grads = view(self.lib.s2vt_att_grads(h), n + 8)
        self.adam_m, self.adam_v = view(self.lib.s2vt_att_adam_m(h), n), view(self.lib.s2vt_att_adam_v(h), n)
        self.adam_step = 0
        self.variables = {}
        for i in range(self.lib.s2vt_att_num_variables(h)):
            name, o, shape, nd = C.c_char_p(), C.c_int64(), (C.c_int64 * 2)(), C.c_int32()
            self._check(self.lib.s2vt_att_variable_info(h, i, C.byref(name), C.byref(o), C.byref(shape), C.byref(nd)))
            self.variables[name.value.decode()] = (o.value, (shape[0], shape[1]) if nd.value == 2 else (shape[0],))
        self.initialize(seed, bias_init_vector)

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.s2vt_att_last_error(self.h)
            raise _lib.S2vtError(rc, msg.decode() if msg else '')

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.lib.s2vt_att_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def variable(self, name, grad=False):
        off, shp = self.variables[name]
        return (self.grads if grad else self.params)[off:off + int(np.prod(shp))].view(*shp)

    def refresh(self):
        self._check(self.lib.s2vt_att_refresh(self.h, _stream()))

    def initialize(self, seed=16, bias_init_vector=None):
        """tf.global_variables_initializer() for :64-86: U(-0.1, 0.1) matrices, zero biases, glorot-uniform LSTM3 kernel."""
        g = torch.Generator(device='cpu').manual_seed(int(seed))
        for name, (off, shp) in self.variables.items():
            if len(shp) == 1:
                v = torch.zeros(shp)
            elif name == LSTM3_W:
                v = (torch.rand(shp, generator=g) * 2 - 1) * float(np.sqrt(6.0 / (shp[0] + shp[1])))
            else:
                v = (torch.rand(shp, generator=g) * 2 - 1) * 0.1
            self.variable(name).copy_(v)
        if bias_init_vector is not None:
            self.variable('embed_word_b').copy_(torch.as_tensor(np.asarray(bias_init_vector), dtype=torch.float32))
        self.refresh()

    def load_variables(self, named_arrays, strict=False):
        """optimistic_restore semantics: copy what matches by TF name and shape; returns the restored names."""
        restored = []
        for name, arr in named_arrays.items():
            a = np.ascontiguousarray(np.asarray(arr, dtype=np.float32))
            shape = (C.c_int64 * max(a.ndim, 1))(*a.shape)
            rc = self.lib.s2vt_att_load_param(self.h, name.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim, _stream())
            torch.cuda.current_stream().synchronize()
            if rc == 0:
                restored.append(name)
            elif rc in (_lib.S2VT_ENOTFOUND, _lib.S2VT_ESHAPE) and not strict:
                continue
            else:
                self._check(rc)
        self.refresh()
        return restored

    def state_dict(self):
        return {k: self.variable(k).detach().cpu().numpy().copy() for k in self.variables}

    def _video(self, video):
        v = torch.as_tensor(video, dtype=torch.float32, device=self.device) if not torch.is_tensor(video) else video.to(self.device, torch.float32)
        if v.dim() != 3 or v.shape[1] != self.n_video_lstm_steps or v.shape[2] != self.dim_image or v.shape[0] > self.batch_size:
            raise ValueError('video must be [B <= %d, %d, %d], got %s' % (self.batch_size, self.n_video_lstm_steps, self.dim_image, tuple(v.shape)))
        return v.contiguous()

    def build_sampler(self, video, want_alphas=True):
        """sess.run([greedy_captions, saved_alphas]) (:606, 638): ids int32 [B, T_c], alphas fp32 [T_c, n, B]."""
        v = self._video(video)
        B = v.shape[0]
        ids = torch.empty(B, self.n_caption_lstm_steps, dtype=torch.int32, device=self.device)
        al = torch.empty(self.n_caption_lstm_steps, self.n_video_lstm_steps, B, dtype=torch.float32, device=self.device) if want_alphas else None
        self._check(self.lib.s2vt_att_greedy(self.h, _ptr(v), B, _ptr(ids), _ptr(al), _stream()))
        return ids, al

    def build_generator(self, video):
        """sess.run(generated_words) (:155-199): ids int32 [B, T_c]."""
        return self.build_sampler(video, want_alphas=False)[0]

    greedy = build_generator

    def build_model(self, video, caption, caption_mask, drop_seed=0, row_base=0, want_logits=False):
        """sess.run(tf_loss, {tf_video, tf_caption, tf_caption_mask}) (:420): (device tensor [loss, regulariser part], logits
        [T_c, B, V] or None).  drop_seed = 0 disables the DropoutWrapper."""
        v = self._video(video)
        B = v.shape[0]
        cap = torch.as_tensor(caption).to(self.device, torch.int32).contiguous()
        mk = torch.as_tensor(caption_mask).to(self.device, torch.float32).contiguous()
        out = torch.empty(2, dtype=torch.float32, device=self.device)
        logits = torch.empty(self.n_caption_lstm_steps, B, self.n_words, dtype=torch.float32, device=self.device) if want_logits else None
        if drop_seed == 0 and self.drop_out_rate < 1.0:
            raise ValueError('drop_seed must be non-zero when drop_out_rate < 1 (use a model built with drop_out_rate=1 for evaluation)')
        self._check(self.lib.s2vt_att_xe_loss(self.h, _ptr(v), B, _ptr(cap), _ptr(mk), int(drop_seed), int(row_base), _ptr(out), _ptr(logits), _stream()))
        return out, logits

    def xe_backward(self, video, caption, caption_mask, drop_seed=0, row_base=0):
        """optimizer.compute_gradients(tf_loss) (:432): loss [2] and the gradients of the 13 variables in `self.grads`."""
        v = self._video(video)
        cap = torch.as_tensor(caption).to(self.device, torch.int32).contiguous()
        mk = torch.as_tensor(caption_mask).to(self.device, torch.float32).contiguous()
        out = torch.empty(2, dtype=torch.float32, device=self.device)
        if drop_seed == 0 and self.drop_out_rate < 1.0:
            raise ValueError('drop_seed must be non-zero when drop_out_rate < 1')
        self._check(self.lib.s2vt_att_xe_backward(self.h, _ptr(v), v.shape[0], _ptr(cap), _ptr(mk), int(drop_seed), int(row_base), _ptr(out), _stream()))
        return out

    def optimizer_step(self, lr, clip_norm=10.0):
        """clip_by_global_norm(., 10) + Adam apply (:433-435) -> device tensor [global gradient norm, loss]."""
        self.adam_step += 1
        out = torch.empty(2, dtype=torch.float32, device=self.device)
        self._check(self.lib.s2vt_att_optimizer_step(self.h, float(lr), float(clip_norm), self.adam_step, _ptr(out), _stream()))
        return out

    def train_step(self, video, caption, caption_mask, global_step, start_learning_rate=1e-4, drop_seed=1):
        """sess.run([train_op, tf_loss]) (:468-476): lr = exponential_decay(1e-4, global_step, 10000, 0.5, staircase) (:429-430)."""
        loss = self.xe_backward(video, caption, caption_mask, drop_seed=drop_seed).clone()
        self.optimizer_step(start_learning_rate * 0.5 ** (global_step // 10000), 10.0)
        return loss

    def launch_count(self):
        return int(self.lib.s2vt_att_launch_count(self.h))
