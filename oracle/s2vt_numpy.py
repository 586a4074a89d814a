"""NumPy oracle of the S2VT graphs.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, in the dtype of the parameter arrays (float64 reference mode or float32 mode):
  * Video_Caption_Generator.__init__         reinforcement_multisampling_tf_s2vt.py:64-98
  * build_model  (XE, teacher forced)        :100-177   (= tf_s2vt.py:90-167)
  * build_loss   (RL log-prob tensor)        :227-292
  * build_multinomial_sampler                :294-339
  * build_sampler (greedy)                   :342-391
  * RL objective / clip / Adam               :638-652 ; XE optimiser tf_s2vt.py:442-448
  * beam_probability                         final_beam_search.py:202-223
TF-1.1 library semantics ([lib] in SURVEY.md: BasicLSTMCell gate order i,j,f,o with
forget_bias 1 inside the sigmoid, DropoutWrapper on the cell *output* only, batch-mean
label-smoothed CE, TF Adam with epsilon outside the bias correction) are restated from
the published r1.1 sources -- parity unpinned, no TF install exists here.

Gradients are hand-derived BPTT; tests/test_oracle_model.py checks them against
torch.autograd on CPU.
"""
import numpy as np

from . import philox

LSTM1_W = 's2vt/LSTM1/basic_lstm_cell/weights'
LSTM1_B = 's2vt/LSTM1/basic_lstm_cell/biases'
LSTM2_W = 's2vt/LSTM2/basic_lstm_cell/weights'
LSTM2_B = 's2vt/LSTM2/basic_lstm_cell/biases'
PARAM_NAMES = ['Wemb', 'encode_image_W', 'encode_image_b', 'embed_word_W', 'embed_word_b',
               LSTM1_W, LSTM1_B, LSTM2_W, LSTM2_B]


def param_shapes(D=1536, E=500, H=1000, V=9972):
    """:64-98 plus the BasicLSTMCell kernels (Q7): LSTM1 rows [x(E); h1(H)], LSTM2 rows [out1(H); emb(E); h2(H)]."""
    return {'Wemb': (V, E), 'encode_image_W': (D, E), 'encode_image_b': (E,),
            'embed_word_W': (H, V), 'embed_word_b': (V,),
            LSTM1_W: (E + H, 4 * H), LSTM1_B: (4 * H,),
            LSTM2_W: (H + E + H, 4 * H), LSTM2_B: (4 * H,)}


def init_params(D=1536, E=500, H=1000, V=9972, seed=4, dtype=np.float32, peaked_bias=None, logit_scale=1.0):
    """Synthetic 'set A' of SURVEY.md 8(d): U(-0.1,0.1) matrices (:79-98), glorot-uniform LSTM kernels, zero biases.

    peaked_bias / logit_scale give 'set B' (embed_word_b = bias_init_vector hook :95-96, embed_word_W x scale).
    """
    rng = np.random.RandomState(seed)
    shp = param_shapes(D, E, H, V)
    p = {}
    for name in PARAM_NAMES:
        s = shp[name]
        if name in (LSTM1_W, LSTM2_W):
            lim = np.sqrt(6.0 / (s[0] + s[1]))
            p[name] = rng.uniform(-lim, lim, size=s)
        elif len(s) == 2:
            p[name] = rng.uniform(-0.1, 0.1, size=s)
        else:
            p[name] = np.zeros(s)
    p['embed_word_W'] = p['embed_word_W'] * logit_scale
    if peaked_bias is not None:
        p['embed_word_b'] = np.asarray(peaked_bias, dtype=np.float64).copy()
    return {k: v.astype(dtype) for k, v in p.items()}


def synthetic_features(B, Tv, D=1536, seed=1234, dtype=np.float32):
    """max(0, N(0.25, 0.5^2)) -- non-negative like post-ReLU pooled CNN features."""
    rng = np.random.RandomState(seed)
    return np.maximum(0.0, rng.normal(0.25, 0.5, size=(B, Tv, D))).astype(dtype)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_cell(x, c, h, W, b):
    """tf.contrib.rnn.BasicLSTMCell r1.1 [lib] (Q7), state_is_tuple=False handled by the caller.

    concat = [x, h] W + b ; i, j, f, o = split(concat, 4) ;
    new_c = c * sigmoid(f + forget_bias) + sigmoid(i) * tanh(j) ; new_h = tanh(new_c) * sigmoid(o)
    """
    one = x.dtype.type(1.0)
    gates = np.concatenate([x, h], axis=1) @ W + b
    i, j, f, o = np.split(gates, 4, axis=1)
    si, tj, sf, so = _sigmoid(i), np.tanh(j), _sigmoid(f + one), _sigmoid(o)
    new_c = c * sf + si * tj
    tc = np.tanh(new_c)
    new_h = tc * so
    return new_h, new_c, (si, tj, sf, so, tc)


def _lstm_cell_bwd(dh, dc_next, cache, c_prev):
    """Backward of lstm_cell w.r.t. the pre-activation gates; returns (dgates [N,4H], dc_prev)."""
    si, tj, sf, so, tc = cache
    one = dh.dtype.type(1.0)
    do = dh * tc
    dc = dc_next + dh * so * (one - tc * tc)
    di = dc * tj
    dj = dc * si
    df = dc * c_prev
    dc_prev = dc * sf
    dgates = np.concatenate([di * si * (one - si), dj * (one - tj * tj), df * sf * (one - sf), do * so * (one - so)], axis=1)
    return dgates, dc_prev


def log_softmax(x):
    m = x.max(axis=-1, keepdims=True)
    z = x - m
    return z - np.log(np.exp(z).sum(axis=-1, keepdims=True))


def encode(p, video):
    """Encoding stage without dropout (:306-316 / :361-370): returns (c1, h1, c2, h2) after T_v frames."""
    N, Tv, D = video.shape
    dt = p['Wemb'].dtype
    E = p['encode_image_W'].shape[1]
    H = p['embed_word_W'].shape[0]
    img = (video.reshape(-1, D).astype(dt) @ p['encode_image_W'] + p['encode_image_b']).reshape(N, Tv, E)
    c1 = np.zeros((N, H), dt); h1 = np.zeros((N, H), dt)
    c2 = np.zeros((N, H), dt); h2 = np.zeros((N, H), dt)
    padding = np.zeros((N, E), dt)
    for i in range(Tv):
        h1, c1, _ = lstm_cell(img[:, i, :], c1, h1, p[LSTM1_W], p[LSTM1_B])
        h2, c2, _ = lstm_cell(np.concatenate([h1, padding], 1), c2, h2, p[LSTM2_W], p[LSTM2_B])
    return c1, h1, c2, h2


def decode_step(p, state, words):
    """One bare-cell decoder step (:321-332): emb(word) -> lstm1(0) -> lstm2([out1, emb]) -> logits."""
    c1, h1, c2, h2 = state
    N = h1.shape[0]
    E = p['Wemb'].shape[1]
    padding = np.zeros((N, E), h1.dtype)
    emb = p['Wemb'][np.asarray(words, dtype=np.int64)]
    h1, c1, _ = lstm_cell(padding, c1, h1, p[LSTM1_W], p[LSTM1_B])
    h2, c2, _ = lstm_cell(np.concatenate([h1, emb], 1), c2, h2, p[LSTM2_W], p[LSTM2_B])
    logits = h2 @ p['embed_word_W'] + p['embed_word_b']
    return logits, (c1, h1, c2, h2)


def greedy_sampler(p, video, Tc=35, return_logits=False):
    """build_sampler :342-391: argmax every step, no early stop.  Returns int64 [N, Tc]."""
    N = video.shape[0]
    state = encode(p, video)
    words = np.ones(N, dtype=np.int64)
    out, all_logits = [], []
    for i in range(Tc):
        logits, state = decode_step(p, state, words)
        words = np.argmax(logits, axis=1).astype(np.int64)
        out.append(words)
        if return_logits:
            all_logits.append(logits)
    ids = np.stack(out, axis=1)
    return (ids, np.stack(all_logits, 0)) if return_logits else ids


def multinomial_sampler(p, video, seed, global_rows, Tc=35, return_logits=False):
    """build_multinomial_sampler :294-339 with tf.multinomial [lib] realised as Gumbel-max over the shared
    Philox stream (oracle/philox.py): word = argmax_v(log_softmax(logits)_v - log(-log(u_v)))."""
    N = video.shape[0]
    V = p['embed_word_b'].shape[0]
    state = encode(p, video)
    words = np.ones(N, dtype=np.int64)
    out, all_logits = [], []
    for i in range(Tc):
        logits, state = decode_step(p, state, words)
        g = philox.gumbel(seed, global_rows, i, V)
        words = np.argmax(logits.astype(np.float32) + g, axis=1).astype(np.int64)
        out.append(words)
        if return_logits:
            all_logits.append(logits)
    ids = np.stack(out, axis=1)
    return (ids, np.stack(all_logits, 0)) if return_logits else ids


def teacher_forward(p, video, caption, drop1=None, drop2=None, keep_cache=True):
    """Shared body of build_model (:100-165) and build_loss (:227-289).

    video [N,Tv,D], caption int [N,Tc]; drop1/drop2: None (keep_prob 1) or float arrays [Tv+Tc, N, H] holding
    0 or 1/keep -- the DropoutWrapper(output_keep_prob) masks of lstm1_dropout / lstm2_dropout (Q2; the
    tf.layers.dropout calls at :117,:246 are inert, Q1).  Returns (logits [Tc,N,V], cache).
    """
    N, Tv, D = video.shape
    Tc = caption.shape[1]
    dt = p['Wemb'].dtype
    E = p['encode_image_W'].shape[1]
    H = p['embed_word_W'].shape[0]
    V = p['embed_word_b'].shape[0]
    X = video.reshape(-1, D).astype(dt)
    img = (X @ p['encode_image_W'] + p['encode_image_b']).reshape(N, Tv, E)
    c1 = np.zeros((N, H), dt); h1 = np.zeros((N, H), dt)
    c2 = np.zeros((N, H), dt); h2 = np.zeros((N, H), dt)
    padding = np.zeros((N, E), dt)
    steps = []
    logits = np.zeros((Tc, N, V), dt)
    for t in range(Tv + Tc):
        x1 = img[:, t, :] if t < Tv else padding
        c1_prev, h1_prev = c1, h1
        h1, c1, k1 = lstm_cell(x1, c1, h1, p[LSTM1_W], p[LSTM1_B])
        out1 = h1 * drop1[t] if drop1 is not None else h1
        if t < Tv:
            emb, ids = padding, None
        else:
            i = t - Tv
            ids = np.ones(N, dtype=np.int64) if i == 0 else caption[:, i - 1].astype(np.int64)
            emb = p['Wemb'][ids]
        c2_prev, h2_prev = c2, h2
        h2, c2, k2 = lstm_cell(np.concatenate([out1, emb], 1), c2, h2, p[LSTM2_W], p[LSTM2_B])
        out2 = h2 * drop2[t] if drop2 is not None else h2
        if t >= Tv:
            logits[t - Tv] = out2 @ p['embed_word_W'] + p['embed_word_b']
        if keep_cache:
            steps.append(dict(x1=x1, h1_prev=h1_prev, c1_prev=c1_prev, k1=k1, out1=out1, emb=emb, ids=ids,
                              h2_prev=h2_prev, c2_prev=c2_prev, k2=k2, out2=out2))
    cache = dict(steps=steps, X=X, N=N, Tv=Tv, Tc=Tc, drop1=drop1, drop2=drop2)
    return logits, cache


def teacher_backward(p, cache, dlogits):
    """BPTT of teacher_forward given dL/dlogits [Tc,N,V].  Returns (grads dict, emb_slice_sqnorm).

    grads['Wemb'] is the dense scatter-added gradient; emb_slice_sqnorm is the sum of squares of the
    un-deduplicated per-step embedding slices (what clip_by_global_norm sees for the IndexedSlices
    gradient of the CPU-pinned gather, R6)."""
    steps = cache['steps']; N = cache['N']; Tv = cache['Tv']; Tc = cache['Tc']
    drop1, drop2 = cache['drop1'], cache['drop2']
    dt = p['Wemb'].dtype
    E = p['encode_image_W'].shape[1]
    H = p['embed_word_W'].shape[0]
    g = {k: np.zeros_like(v) for k, v in p.items()}
    W1, W2, Wo = p[LSTM1_W], p[LSTM2_W], p['embed_word_W']
    dh1 = np.zeros((N, H), dt); dc1 = np.zeros((N, H), dt)
    dh2 = np.zeros((N, H), dt); dc2 = np.zeros((N, H), dt)
    dimg = np.zeros((N, Tv, E), dt)
    emb_sq = 0.0
    for t in range(Tv + Tc - 1, -1, -1):
        s = steps[t]
        dh2_t = dh2
        if t >= Tv:
            dl = dlogits[t - Tv]
            g['embed_word_W'] += s['out2'].T @ dl
            g['embed_word_b'] += dl.sum(0)
            dout2 = dl @ Wo.T
            dh2_t = dh2_t + (dout2 * drop2[t] if drop2 is not None else dout2)
        dg2, dc2 = _lstm_cell_bwd(dh2_t, dc2, s['k2'], s['c2_prev'])
        in2 = np.concatenate([s['out1'], s['emb'], s['h2_prev']], 1)
        g[LSTM2_W] += in2.T @ dg2
        g[LSTM2_B] += dg2.sum(0)
        din2 = dg2 @ W2.T
        dout1, demb, dh2 = din2[:, :H], din2[:, H:H + E], din2[:, H + E:]
        if t >= Tv:
            np.add.at(g['Wemb'], s['ids'], demb)
            emb_sq += float((demb.astype(np.float64) ** 2).sum())
        dh1_t = dh1 + (dout1 * drop1[t] if drop1 is not None else dout1)
        dg1, dc1 = _lstm_cell_bwd(dh1_t, dc1, s['k1'], s['c1_prev'])
        in1 = np.concatenate([s['x1'], s['h1_prev']], 1)
        g[LSTM1_W] += in1.T @ dg1
        g[LSTM1_B] += dg1.sum(0)
        din1 = dg1 @ W1.T
        dx1, dh1 = din1[:, :E], din1[:, E:]
        if t < Tv:
            dimg[:, t, :] = dx1
    dimg2 = dimg.reshape(-1, E)
    g['encode_image_W'] += cache['X'].T @ dimg2
    g['encode_image_b'] += dimg2.sum(0)
    return g, emb_sq


def rl_logprobs(logits, caption, mask):
    """build_loss :283-291 reduced over the vocabulary axis: logp[n,t] = log_softmax(logits)[t,n,w_nt] * mask[n,t]."""
    lsm = log_softmax(logits)                      # [Tc,N,V]
    Tc, N, _ = logits.shape
    picked = lsm[np.arange(Tc)[:, None], np.arange(N)[None, :], caption.T.astype(np.int64)]
    return picked.T * mask.astype(logits.dtype), lsm


def rl_objective(p, video, caption, mask, rewards, base_line, drop1=None, drop2=None, want_grads=True):
    """:643-646: sum_loss = -sum(loss * (rewards - base_line)[:,None,None]) / sum(mask).  Returns (loss, grads, aux)."""
    dt = p['Wemb'].dtype
    logits, cache = teacher_forward(p, video, caption, drop1, drop2, keep_cache=want_grads)
    mask = mask.astype(dt)
    logp, lsm = rl_logprobs(logits, caption, mask)
    norm = mask.sum()
    residual = (np.asarray(rewards, dt) - np.asarray(base_line, dt))
    sum_loss = -(logp * residual[:, None]).sum() / norm
    aux = dict(logits=logits, logp=logp, norm=norm)
    if not want_grads:
        return sum_loss, None, aux
    Tc, N, V = logits.shape
    coef = (residual[:, None] * mask / norm).T            # [Tc,N]   (R4)
    dlogits = np.exp(lsm) * coef[:, :, None]
    dlogits[np.arange(Tc)[:, None], np.arange(N)[None, :], caption.T.astype(np.int64)] -= coef
    grads, emb_sq = teacher_backward(p, cache, dlogits)
    aux['emb_slice_sqnorm'] = emb_sq
    return sum_loss, grads, aux


def xe_objective(p, video, caption, mask, drop1=None, drop2=None, label_smoothing=0.05, decay=5e-5,
                 loss_weight=1.0, want_grads=True):
    """build_model :153-166 (tf_s2vt.py:143-166).  Q3: tf.losses.softmax_cross_entropy returns the batch MEAN, which
    is then multiplied by mask[:,i] and summed, i.e. step loss = mean_b(CE_b) * sum_b mask[b,i].  Q4: L2 on every
    trainable variable whose name lacks the substring 'bias' (the LSTM '.../biases' only)."""
    dt = p['Wemb'].dtype
    logits, cache = teacher_forward(p, video, caption, drop1, drop2, keep_cache=want_grads)
    Tc, N, V = logits.shape
    mask = mask.astype(dt)
    lsm = log_softmax(logits)
    ls = dt.type(label_smoothing)
    picked = lsm[np.arange(Tc)[:, None], np.arange(N)[None, :], caption.T.astype(np.int64)]     # [Tc,N]
    ce = -((1 - ls) * picked + (ls / V) * lsm.sum(-1))                                           # [Tc,N]
    S = mask.sum(0)                                                                              # [Tc]
    norm = mask.sum()
    loss = loss_weight * (ce.mean(1) * S).sum() / norm
    wd = 0.0
    for name, v in p.items():
        if 'bias' not in name:
            wd = wd + (v.astype(np.float64) ** 2).sum() / 2.0
    total = loss + decay * wd
    aux = dict(logits=logits, xe=loss, weight_decay=decay * wd)
    if not want_grads:
        return total, None, aux
    coef = (loss_weight * S / (N * norm)).astype(dt)                                             # [Tc]
    dlogits = (np.exp(lsm) - ls / V) * coef[:, None, None]
    dlogits[np.arange(Tc)[:, None], np.arange(N)[None, :], caption.T.astype(np.int64)] -= (1 - ls) * coef[:, None]
    grads, _ = teacher_backward(p, cache, dlogits)
    for name, v in p.items():
        if 'bias' not in name:
            grads[name] = grads[name] + dt.type(decay) * v
    return total, grads, aux


def clip_by_global_norm(grads, clip_norm, emb_slice_sqnorm=None):
    """tf.clip_by_global_norm [lib]: scale = clip_norm * min(1/global_norm, 1/clip_norm).  With emb_slice_sqnorm the
    Wemb term of the norm is the un-deduplicated IndexedSlices norm (R6), else the dense norm."""
    sq = 0.0
    for k, v in grads.items():
        if k == 'Wemb' and emb_slice_sqnorm is not None:
            sq += emb_slice_sqnorm
        else:
            sq += float((v.astype(np.float64) ** 2).sum())
    gn = np.sqrt(sq)
    scale = clip_norm * min(1.0 / gn, 1.0 / clip_norm) if gn > 0 else 1.0
    return {k: (v * v.dtype.type(scale)) for k, v in grads.items()}, gn


class TFAdam:
    """tf.train.AdamOptimizer r1.1 [lib] (R7): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps)."""

    def __init__(self, params, beta1=0.9, beta2=0.999, eps=1e-8):
        self.m = {k: np.zeros_like(v) for k, v in params.items()}
        self.v = {k: np.zeros_like(v) for k, v in params.items()}
        self.b1, self.b2, self.eps = beta1, beta2, eps
        self.t = 0

    def apply(self, params, grads, lr):
        self.t += 1
        lr_t = lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k in params:
            dt = params[k].dtype.type
            g = grads[k]
            self.m[k] = dt(self.b1) * self.m[k] + dt(1 - self.b1) * g
            self.v[k] = dt(self.b2) * self.v[k] + dt(1 - self.b2) * g * g
            params[k] = params[k] - dt(lr_t) * self.m[k] / (np.sqrt(self.v[k]) + dt(self.eps))
        return params


def exponential_decay(lr0, global_step, decay_steps, rate=0.5):
    """tf.train.exponential_decay(staircase=True) (:639-640, tf_s2vt.py:442-443)."""
    return lr0 * rate ** (global_step // decay_steps)


def beam_step_fn(p, beam_size):
    """beam_probability final_beam_search.py:202-223 as a callable for oracle.beam.beam_search.  States are the
    state_is_tuple=False layout concat([c, h], 1) of width 2H; probabilities come from the un-stabilised
    exp(l)/sum(exp(l)) (B6); top_k ties resolve to the lower index [lib]."""
    H = p['embed_word_W'].shape[0]

    def step(state1, state2, word):
        c1, h1 = state1[:, :H], state1[:, H:]
        c2, h2 = state2[:, :H], state2[:, H:]
        logits, (c1, h1, c2, h2) = decode_step(p, (c1, h1, c2, h2), np.asarray(word))
        e = np.exp(logits[0])
        prob = e / e.sum()
        idx = np.argsort(-prob, kind='stable')[:beam_size]
        return idx.astype(np.int32), prob[idx].astype(np.float32), np.concatenate([c2, h2], 1), np.concatenate([c1, h1], 1)

    return step


def beam_initial_states(p, video1):
    """final_beam_search.py:226-252: encode one video; returns (state1, state2) as [1, 2H] arrays."""
    c1, h1, c2, h2 = encode(p, video1)
    return np.concatenate([c1, h1], 1), np.concatenate([c2, h2], 1)


def attribute_loss(features, labels, attr_W, attr_b, want_grads=True):
    """reinforce_multitask_e2e_attribute_loss.py:375-380: mean over frames -> xw_plus_b -> sigmoid CE, / (400*B)."""
    dt = attr_W.dtype
    B = features.shape[0]
    A = attr_W.shape[1]
    pooled = features.astype(dt).mean(axis=1)
    z = pooled @ attr_W + attr_b
    y = labels.astype(dt)
    ce = np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z)))       # tf.nn.sigmoid_cross_entropy_with_logits [lib]
    loss = ce.sum() / (A * B)
    if not want_grads:
        return loss, None
    dz = (_sigmoid(z) - y) / (A * B)
    return loss, {'attr_W': pooled.T @ dz, 'attr_b': dz.sum(0)}
