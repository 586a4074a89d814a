"""Host-side mirror of the reference's `class Video_Caption_Generator` over the B200 C ABI.

The reference builds TF-1.1 graphs (`build_model`, `build_loss`, `build_multinomial_sampler`, `build_sampler`,
`beam_probability`, reinforcement_multisampling_tf_s2vt.py:63-466, final_beam_search.py:202-294) and drives them
with `sess.run(fetches, feed_dict)`.  Here each graph is one call into libs2vt_b200.so; the feed / fetch
shapes are the same.  PyTorch only owns device memory and streams.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

LSTM1_W = 's2vt/LSTM1/basic_lstm_cell/weights'
LSTM1_B = 's2vt/LSTM1/basic_lstm_cell/biases'
LSTM2_W = 's2vt/LSTM2/basic_lstm_cell/weights'
LSTM2_B = 's2vt/LSTM2/basic_lstm_cell/biases'


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Video_Caption_Generator(object):
    """Same constructor arguments as the reference (:64-66, final_beam_search.py adds beam_size)."""

    def __init__(self, dim_image=1536, n_words=9972, word_dim=500, lstm_dim=1000, batch_size=64, n_lstm_steps=40,
                 n_video_lstm_step=5, n_caption_lstm_step=35, bias_init_vector=None, loss_weight=1, decay_value=0.00005,
                 dropout_rate=0.9, beam_size=3, n_attributes=0, precision='bf16', gemm_backend='auto', device=None,
                 max_videos=None, max_rows=None, seed=4):
        if not torch.cuda.is_available():
            raise RuntimeError('multitask-end-to-end-video-captioning_b200 needs a CUDA device (sm_100a); there is no CPU path')
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self.dim_image, self.n_words, self.word_dim, self.lstm_dim = dim_image, n_words, word_dim, lstm_dim
        self.batch_size, self.n_lstm_steps = batch_size, n_lstm_steps
        self.n_video_lstm_step, self.n_caption_lstm_step = n_video_lstm_step, n_caption_lstm_step
        self.loss_weight, self.decay_value, self.dropout_rate, self.beam_size = loss_weight, decay_value, dropout_rate, beam_size
        self.n_attributes = n_attributes
        cfg = _lib.S2vtConfig(dim_image, word_dim, lstm_dim, n_words, n_video_lstm_step, n_caption_lstm_step, n_attributes,
                              {'bf16': _lib.PREC_BF16, 'fp32': _lib.PREC_FP32}[precision],
                              {'auto': _lib.GEMM_AUTO, 'mma_sync': _lib.GEMM_MMA_SYNC, 'tcgen05': _lib.GEMM_TCGEN05, 'tcgen05_n128': 3, 'tcgen05_mc2x2': 4, 'step_mc8': 5, 'step_mc4': 6, 'step_n64': 7, 'wgrad_transposed': 8, 'per_step': 9, 'chain_ring': 10, 'chain_mc4': 11, 'chain_nomc': 12, 'chain_plain': 13, 'single_cta': 14, 'pair_all': 15, 'chain_ws2_bwd': 16}[gemm_backend],
                              float(dropout_rate))
        self.precision = precision
        h = C.c_void_p()
        rc = self.lib.s2vt_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise _lib.S2vtError(rc, 's2vt_create rejected the configuration')
        self.h = h
        self.max_videos = int(max_videos or batch_size)
        self.max_rows = int(max_rows or 8 * batch_size)
        with torch.cuda.device(self.device):
            self._state = torch.zeros(self.lib.s2vt_state_bytes(h) + 256, dtype=torch.uint8, device=self.device)
            self._ws = torch.empty(self.lib.s2vt_workspace_bytes(h, self.max_videos, self.max_rows, max(beam_size, 1)) + 256,
                                   dtype=torch.uint8, device=self.device)
            so, wo = (-self._state.data_ptr()) % 256, (-self._ws.data_ptr()) % 256
            self._check(self.lib.s2vt_bind(h, C.c_void_p(self._state.data_ptr() + so), self._state.numel() - so,
                                           C.c_void_p(self._ws.data_ptr() + wo), self._ws.numel() - wo))
        self.n_params = self.lib.s2vt_num_params(h)
        base = self._state.data_ptr()
        self.params = self._view(self.lib.s2vt_params(h) - base, self.n_params)
        self.grads = self._view(self.lib.s2vt_grads(h) - base, self.n_params + 8)     # + aux: [slice sq-norm, loss, ...]
        self.adam_m = self._view(self.lib.s2vt_adam_m(h) - base, self.n_params)
        self.adam_v = self._view(self.lib.s2vt_adam_v(h) - base, self.n_params)
        self.variables = {}
        for i in range(self.lib.s2vt_num_variables(h)):
            name, off, shape, nd = C.c_char_p(), C.c_int64(), (C.c_int64 * 2)(), C.c_int32()
            self._check(self.lib.s2vt_variable_info(h, i, C.byref(name), C.byref(off), C.byref(shape), C.byref(nd)))
            shp = (shape[0], shape[1]) if nd.value == 2 else (shape[0],)
            self.variables[name.value.decode()] = (off.value, shp)
        self.adam_step = 0
        self.peer_world = 0          # > 0 once peer_connect mapped the other ranks' gradient blocks
        self.optimizer_sharded = False
        self._loss = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.initialize(seed, bias_init_vector)

    # ---- plumbing --------------------------------------------------------------------------------------------
    def _view(self, byte_off, n):
        return self._state[byte_off:byte_off + 4 * n].view(torch.float32)

    def _check(self, rc):
        _lib.check(self.h, rc)

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.lib.s2vt_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def variable(self, name, grad=False):
        """Tensor view (TF layout) of a variable or of its gradient inside the flat blocks."""
        off, shp = self.variables[name]
        n = int(np.prod(shp))
        return (self.grads if grad else self.params)[off:off + n].view(*shp)

    def refresh(self):
        self._check(self.lib.s2vt_refresh(self.h, _stream()))

    def initialize(self, seed=4, bias_init_vector=None):
        """tf.global_variables_initializer() for the variables of :79-98: U(-0.1, 0.1) matrices, zero biases,
        glorot-uniform LSTM kernels (BasicLSTMCell default initializer [lib])."""
        g = torch.Generator(device='cpu').manual_seed(int(seed))
        for name, (off, shp) in self.variables.items():
            if len(shp) == 1:
                v = torch.zeros(shp)
            elif name in (LSTM1_W, LSTM2_W):
                lim = float(np.sqrt(6.0 / (shp[0] + shp[1])))
                v = (torch.rand(shp, generator=g) * 2 - 1) * lim
            else:
                v = (torch.rand(shp, generator=g) * 2 - 1) * 0.1
            self.variable(name).copy_(v)
        if bias_init_vector is not None:
            self.variable('embed_word_b').copy_(torch.as_tensor(np.asarray(bias_init_vector), dtype=torch.float32))
        self.adam_m.zero_(); self.adam_v.zero_(); self.adam_step = 0
        self.refresh()

    def load_variables(self, named_arrays, strict=False):
        """optimistic_restore (:47-61): copy every array whose TF name AND shape match, skip the rest silently.
        Returns the list of restored names."""
        restored = []
        for name, arr in named_arrays.items():
            a = np.ascontiguousarray(np.asarray(arr, dtype=np.float32))
            shape = (C.c_int64 * max(a.ndim, 1))(*a.shape)
            rc = self.lib.s2vt_load_param(self.h, name.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim, _stream())
            torch.cuda.current_stream().synchronize()   # the host array may be freed after return
            if rc == 0:
                restored.append(name)
            elif rc in (_lib.S2VT_ENOTFOUND, _lib.S2VT_ESHAPE) and not strict:
                continue
            else:
                self._check(rc)
        self.refresh()
        return restored

    def state_dict(self):
        """{TF variable name: fp32 numpy array} -- what tf.train.Saver would have written for the trainables."""
        return {k: self.variable(k).detach().cpu().numpy().copy() for k in self.variables}

    def _video(self, video):
        v = torch.as_tensor(video, dtype=torch.float32, device=self.device) if not torch.is_tensor(video) else video.to(self.device, torch.float32)
        if v.dim() != 3 or v.shape[1] != self.n_video_lstm_step or v.shape[2] != self.dim_image:
            raise ValueError('video must be [B, %d, %d], got %s' % (self.n_video_lstm_step, self.dim_image, tuple(v.shape)))
        return v.contiguous()

    def _i32(self, x):
        t = torch.as_tensor(x) if not torch.is_tensor(x) else x
        return t.to(self.device, torch.int32).contiguous()

    def _f32(self, x):
        t = torch.as_tensor(x) if not torch.is_tensor(x) else x
        return t.to(self.device, torch.float32).contiguous()

    # ---- decoding graphs -------------------------------------------------------------------------------------
    def greedy(self, video):
        """build_sampler + sess.run(greedy_captions) (:342-391, :975): video [B,T_v,D] -> int32 [B, T_c] (device)."""
        v = self._video(video)
        out = torch.empty(v.shape[0], self.n_caption_lstm_step, dtype=torch.int32, device=self.device)
        self._check(self.lib.s2vt_greedy(self.h, _ptr(v), v.shape[0], _ptr(out), _stream()))
        return out

    def rollout(self, video, K, seed, row_base=0, with_greedy=True):
        """K runs of build_multinomial_sampler (:294-339, :743-753) plus the greedy baseline, frames encoded once.
        Returns (sampled int32 [K*B, T_c] sample-major, greedy int32 [B, T_c] or None)."""
        v = self._video(video)
        B = v.shape[0]
        samp = torch.empty(K * B, self.n_caption_lstm_step, dtype=torch.int32, device=self.device)
        gr = torch.empty(B, self.n_caption_lstm_step, dtype=torch.int32, device=self.device) if with_greedy else None
        self._check(self.lib.s2vt_rollout(self.h, _ptr(v), B, K, int(seed), int(row_base), _ptr(samp), _ptr(gr), _stream()))
        return samp, gr

    def sample(self, video, seed, row_base=0):
        """One build_multinomial_sampler run: video [B,T_v,D] -> int32 [B, T_c]."""
        return self.rollout(video, 1, seed, row_base, with_greedy=False)[0]

    def caption_masks(self, ids):
        """The mask half of decode_captions_masks (cider_evaluation.py:145-172) on the device: (mask f32, lengths i32)."""
        ids = self._i32(ids)
        N = ids.shape[0]
        mask = torch.empty(N, self.n_caption_lstm_step, dtype=torch.float32, device=self.device)
        lens = torch.empty(N, dtype=torch.int32, device=self.device)
        self._check(self.lib.s2vt_caption_masks(self.h, _ptr(ids), N, _ptr(mask), _ptr(lens), _stream()))
        return mask, lens

    # ---- training graphs -------------------------------------------------------------------------------------
    def teacher_forward(self, video, captions, drop_seed=0, row_base=0, want_logits=False):
        """build_loss / build_model forward: returns (logp [N,T_c] un-masked, logits [T_c,N,V] or None)."""
        v = self._video(video); cap = self._i32(captions)
        B, N = v.shape[0], cap.shape[0]
        logp = torch.empty(N, self.n_caption_lstm_step, dtype=torch.float32, device=self.device)
        logits = torch.empty(self.n_caption_lstm_step, N, self.n_words, dtype=torch.float32, device=self.device) if want_logits else None
        self._check(self.lib.s2vt_teacher_forward(self.h, _ptr(v), B, _ptr(cap), N, int(drop_seed), int(row_base), _ptr(logp), _ptr(logits), _stream()))
        return logp, logits

    def rl_backward(self, video, captions, mask, rewards, base_line, norm=0.0, grad_scale=1.0, accumulate=False, drop_seed=0, row_base=0):
        """sum_loss and tf.gradients(sum_loss, ...) of :643-650.  Returns the loss as a 1-element device tensor."""
        v = self._video(video); cap = self._i32(captions); m = self._f32(mask); r = self._f32(rewards); b = self._f32(base_line)
        self._check(self.lib.s2vt_rl_backward(self.h, _ptr(v), v.shape[0], _ptr(cap), _ptr(m), _ptr(r), _ptr(b), cap.shape[0], float(norm),
                                              float(grad_scale), int(bool(accumulate)), int(drop_seed), int(row_base), _ptr(self._loss), _stream()))
        return self._loss[:1]

    def xe_backward(self, video, captions, mask, norm=0.0, grad_scale=1.0, accumulate=False, drop_seed=0, row_base=0, label_smoothing=0.05):
        """tf_loss and its gradients (build_model :153-166).  Returns device tensor [total loss, weight-decay part]."""
        v = self._video(video); cap = self._i32(captions); m = self._f32(mask)
        self._check(self.lib.s2vt_xe_backward(self.h, _ptr(v), v.shape[0], _ptr(cap), _ptr(m), cap.shape[0], float(label_smoothing),
                                              float(self.decay_value), float(norm), float(grad_scale), int(bool(accumulate)), int(drop_seed),
                                              int(row_base), _ptr(self._loss), _stream()))
        return self._loss[:2]

    def xe_backward_sharded(self, video, captions, mask, mask_colsum_global, n_rows_global, norm, decay=None, grad_scale=1.0, accumulate=False,
                            drop_seed=0, row_base=0, label_smoothing=0.05):
        """This rank's share of tf_loss and its gradients when the batch is split over ranks (Q3 couples the rows: the global per-step
        mask sums [T_c], the global row count and the global sum(mask) are inputs).  Shares ADD up over the ranks."""
        v = self._video(video); cap = self._i32(captions); m = self._f32(mask); cs = self._f32(mask_colsum_global)
        d = self.decay_value if decay is None else decay
        self._check(self.lib.s2vt_xe_backward_sharded(self.h, _ptr(v), v.shape[0], _ptr(cap), _ptr(m), cap.shape[0], float(label_smoothing), float(d),
                                                      float(norm), _ptr(cs), int(n_rows_global), float(grad_scale), int(bool(accumulate)), int(drop_seed),
                                                      int(row_base), _ptr(self._loss), _stream()))
        return self._loss[:2]

    def attribute_backward(self, video, labels, grad_scale=1.0):
        v = self._video(video); y = self._f32(labels)
        out = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._check(self.lib.s2vt_attribute_backward(self.h, _ptr(v), v.shape[0], _ptr(y), float(grad_scale), _ptr(out), _stream()))
        return out

    def optimizer_step(self, lr, clip_norm, wemb_slice_norm=True, normalize=False):
        """clip_by_global_norm + Adam apply (:650-652).  Returns device tensor [global gradient norm, loss aux slot].
        normalize=True: gradients were accumulated with norm=1 (and all-reduced); divide by the global sum(mask)."""
        self.adam_step += 1
        out = torch.empty(2, dtype=torch.float32, device=self.device)
        flags = (1 if wemb_slice_norm else 0) | (2 if normalize else 0)
        self._check(self.lib.s2vt_optimizer_step(self.h, float(lr), float(clip_norm), self.adam_step, flags, _ptr(out), _stream()))
        return out

    def grad_segment_ready(self, stream, segment=0):
        """Make `stream` (torch.cuda.Stream) wait until the early-final gradient segment of the last rl_backward is complete; returns
        (offset, count) in floats inside `self.grads`, or None when the last backward gives no early segment."""
        off, cnt = C.c_int64(), C.c_int64()
        rc = self.lib.s2vt_grad_segment_ready(self.h, int(segment), C.c_void_p(stream.cuda_stream), C.byref(off), C.byref(cnt))
        if rc == _lib.S2VT_ESTATE:
            return None
        self._check(rc)
        return off.value, cnt.value

    # ---- data-parallel exchange over NVLink peer memory (include/s2vt.h: s2vt_peer_*) -----------------------------------
    def peer_export(self):
        """(state handle, offset of the state block in its allocation, flag-block handle) to hand to the other ranks."""
        hs, hc, off = (C.c_ubyte * 64)(), (C.c_ubyte * 64)(), C.c_int64()
        self._check(self.lib.s2vt_peer_export(self.h, hs, C.byref(off), hc))
        return bytes(hs), off.value, bytes(hc)

    def peer_connect(self, rank, exports):
        """exports: the peer_export() tuples of all ranks, in rank order."""
        world = len(exports)
        hs = (C.c_ubyte * (64 * world)).from_buffer_copy(b''.join(e[0] for e in exports))
        hc = (C.c_ubyte * (64 * world)).from_buffer_copy(b''.join(e[2] for e in exports))
        offs = (C.c_int64 * world)(*[e[1] for e in exports])
        self._check(self.lib.s2vt_peer_connect(self.h, int(rank), world, hs, offs, hc))
        self.peer_world = world

    def peer_allreduce(self):
        self._check(self.lib.s2vt_peer_allreduce(self.h, _stream()))

    def peer_optimizer_step(self, lr, clip_norm, wemb_slice_norm=True, normalize=True):
        """allreduce_gradients + optimizer_step as one kernel per rank (collective): every rank clips and applies Adam to its own slice of the
        flat vector, the updated parameters are gathered.  The Adam slots stay sharded until gather_optimizer_state()."""
        self.adam_step += 1
        out = torch.empty(2, dtype=torch.float32, device=self.device)
        flags = (1 if wemb_slice_norm else 0) | (2 if normalize else 0)
        self._check(self.lib.s2vt_peer_optimizer_step(self.h, float(lr), float(clip_norm), self.adam_step, flags, _ptr(out), _stream()))
        self.optimizer_sharded = True
        return out

    def gather_optimizer_state(self):
        """Collective when the Adam slots are sharded (after peer_optimizer_step): make them whole on every rank (before saving them)."""
        if getattr(self, 'optimizer_sharded', False):
            self._check(self.lib.s2vt_peer_gather_state(self.h, _stream()))
            self.optimizer_sharded = False

    def peer_disconnect(self):
        self.gather_optimizer_state()
        self.peer_world = 0
        self._check(self.lib.s2vt_peer_disconnect(self.h))

    def set_reuse_frontend(self, enable=True):
        """Share the LSTM1 forward of a rollout with the training call that follows on the same video tensor."""
        self._check(self.lib.s2vt_set_reuse_frontend(self.h, int(bool(enable))))

    # ---- instrumentation -----------------------------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.s2vt_launch_count(self.h))

    def profile(self, enable):
        self._check(self.lib.s2vt_profile(self.h, int(bool(enable))))

    def profile_shapes(self, cap=256):
        """[(cls, M, N, K, total ms, GEMMs, algorithmic bytes, kernel launches)] recorded since the last profile_read."""
        I = C.c_int * cap
        cls, M, N, K, ms, cnt = I(), I(), I(), I(), (C.c_double * cap)(), (C.c_longlong * cap)()
        by, ln = (C.c_double * cap)(), (C.c_longlong * cap)()
        n = self.lib.s2vt_profile_shapes(self.h, cap, cls, M, N, K, ms, cnt, by, ln)
        return [(cls[i], M[i], N[i], K[i], ms[i], cnt[i], by[i], ln[i]) for i in range(max(n, 0))]

    def profile_read(self):
        """-> {'batched': (ms, flops, launches), 'step': (...)} of the GEMM launches since the last read."""
        ms, fl, n = (C.c_double * 2)(), (C.c_double * 4)(), (C.c_longlong * 2)()
        self._check(self.lib.s2vt_profile_read(self.h, ms, fl, n))
        return {'batched': (ms[0], fl[0], n[0], fl[2]), 'step': (ms[1], fl[1], n[1], fl[3])}

    # ---- drop-in step functions (the literal sess.run contracts) ------------------------------------------------
    def rl_step(self, mask, captions, video, rewards, base_line, lr, clip_norm=5.0, drop_seed=0, n_videos=None):
        """sess.run([train_op, sum_loss], {loss_masks, loss_captions, loss_features, rewards, base_line}) (:823-826).
        `video` is either the literal [N, T_v, D] feed (features replicated per sample, :779-782) or the
        de-duplicated [B, T_v, D] with N % B == 0."""
        loss = self.rl_backward(video, captions, mask, rewards, base_line, drop_seed=drop_seed).clone()
        self.optimizer_step(lr, clip_norm)
        return loss

    def xe_step(self, video, captions, mask, lr, clip_norm=10.0, drop_seed=0):
        """sess.run([train_op, tf_loss], {tf_video, tf_caption, tf_caption_mask}) (tf_s2vt.py:491-497)."""
        loss = self.xe_backward(video, captions, mask, drop_seed=drop_seed).clone()
        self.optimizer_step(lr, clip_norm, wemb_slice_norm=False)
        return loss[:1]

    # ---- beam search ---------------------------------------------------------------------------------------------
    def beam_search(self, video, beam_size=None, length_normalization_factor=0.0):
        """build_generator of final_beam_search.py:225-294 for a whole batch.
        Returns (sentences int32 [B,T_c], lengths int32 [B], logprob f32 [B], score f32 [B]) on the device."""
        v = self._video(video)
        B = v.shape[0]
        k = int(beam_size or self.beam_size)
        sent = torch.zeros(B, self.n_caption_lstm_step, dtype=torch.int32, device=self.device)
        lens = torch.zeros(B, dtype=torch.int32, device=self.device)
        lp = torch.zeros(B, dtype=torch.float32, device=self.device)
        sc = torch.zeros(B, dtype=torch.float32, device=self.device)
        self._check(self.lib.s2vt_beam_search(self.h, _ptr(v), B, k, float(length_normalization_factor), _ptr(sent), _ptr(lens), _ptr(lp), _ptr(sc), _stream()))
        return sent, lens, lp, sc

    def build_generator(self, video, length_normalization_factor=0.0):
        """Literal return shape of the reference's build_generator: (sentence list incl. the final <eos>, logprob, score)."""
        sent, lens, lp, sc = self.beam_search(video, self.beam_size, length_normalization_factor)
        n = int(lens[0].item())
        return [int(x) for x in sent[0, :n].tolist()], float(lp[0].item()), float(sc[0].item())

    def beam_init(self, video):
        """get_init_state / the encoder half of build_generator (:226-252): -> (state1, state2) each [1, 2*lstm_dim]."""
        v = self._video(video)
        s1 = torch.empty(1, 2 * self.lstm_dim, dtype=torch.float32, device=self.device)
        s2 = torch.empty(1, 2 * self.lstm_dim, dtype=torch.float32, device=self.device)
        self._check(self.lib.s2vt_beam_init(self.h, _ptr(v), _ptr(s1), _ptr(s2), _stream()))
        return s1, s2

    def beam_probability(self, state2, state1, word, beam_size=None):
        """One beam_probability run (:202-223): -> (word_index [k], probs [k], state2', state1')."""
        k = int(beam_size or self.beam_size)
        s2 = self._f32(state2); s1 = self._f32(state1); w = self._i32(word)
        idx = torch.empty(k, dtype=torch.int32, device=self.device)
        pr = torch.empty(k, dtype=torch.float32, device=self.device)
        n2 = torch.empty_like(s2); n1 = torch.empty_like(s1)
        self._check(self.lib.s2vt_beam_step(self.h, _ptr(s2), _ptr(s1), _ptr(w), k, _ptr(idx), _ptr(pr), _ptr(n2), _ptr(n1), _stream()))
        return idx, pr, n2, n1
